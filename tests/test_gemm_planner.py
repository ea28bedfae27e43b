"""The tile / split-K plan of the batched GEMMs (csrc/runtime.cuh:plan_gemm_full_bf16) is host code: pin its decisions for the shapes of the
MSVD train step to what the sweeps on the B200 found best (profiles/r2_h_gemm_sweep.md, r1_g_gemm_sweep.md).  No GPU needed."""
import ctypes as C

import pytest

import recnet_b200  # noqa: F401
from recnet_b200 import _lib as L

B, T, E, H, A, V, Lr, R = 100, 28, 1536, 512, 128, 4188, 31, 1536
LB, BT, SB, GR = Lr * B, B * T, 28 * B, 4 * R


def plan(M, N, K, prec=L.PREC_BF16):
    bn, splits = C.c_int32(), C.c_int32()
    L.check(L.lib().recnet_plan_batched_gemm(prec, M, N, K, C.byref(bn), C.byref(splits)), "recnet_plan_batched_gemm")
    return bn.value, splits.value


@pytest.mark.parametrize("name,shape,expect", [
    ("rec.dW_hh: the one GEMM with >= 1.5 rounds of pair tiles and a long K loop", (GR, R, SB), 2256),
    ("dec.logits: K = 512, many tiles", (LB, V, H), 1256),
    ("dec.Gx", (LB, 4 * H, 512), 1256),
    ("dec.VW", (BT, 4 * H, E), 1256),
    ("rec.out", (SB, R, R), 1256),
    ("rec.dW_ih: 96 tiles of 256", (GR, H, SB), 1256),
    ("dec.dW_ctx", (4 * H, E, LB), 1256),
    ("rec.out_w: 72 tiles of 256 -> 128-wide", (R, R, SB), 1128),
    ("dec.out_w", (V, H, LB), 1128),
    ("dec.dHext", (LB, H, 4192), 1128),
    ("dec.dXe", (LB, 468, 4 * H), 1128),
])
def test_persistent_kernel_plans(name, shape, expect):
    bn, splits = plan(*shape)
    assert (bn, splits) == (expect, 1), name


@pytest.mark.parametrize("name,shape", [
    ("dec.dW_a", (A, H, LB)), ("dec.dU", (A, E, BT)), ("rec.attn_W", (A, R, SB)), ("rec.attn_U", (A, H, LB)),
])
def test_skinny_attention_gradients_stay_on_split_k(name, shape):
    bn, splits = plan(*shape)
    assert bn == 64 and splits >= 4, (name, bn, splits)


@pytest.mark.parametrize("shape", [(BT, A, E), (LB, A, H)])
def test_128_column_key_projections_stay_on_the_one_tile_kernel(shape):
    bn, splits = plan(*shape)
    assert bn in (64, 128) and splits == 1


def test_background_budget_limits_split_k_plans():
    lib = L.lib()
    free = plan(A, R, SB)
    L.check(lib.recnet_set_background_ctas(48), "recnet_set_background_ctas")
    try:
        bn, splits = plan(A, R, SB)
        assert bn in (64, 128) and -(-R // bn) * splits <= 48, (bn, splits)   # N-tiles x splits within the lane's budget
        assert plan(GR, R, SB) == (2256, 1)                                  # persistent kernels are capped at launch, not re-planned
    finally:
        lib.recnet_set_background_ctas(0)
    assert plan(A, R, SB) == free
    assert lib.recnet_set_background_ctas(-1) < 0


def test_rejects_empty_shapes():
    bn, splits = C.c_int32(), C.c_int32()
    assert L.lib().recnet_plan_batched_gemm(L.PREC_BF16, 0, 8, 8, C.byref(bn), C.byref(splits)) < 0
