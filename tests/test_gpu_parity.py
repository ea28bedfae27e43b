"""GPU parity tests (run on the B200 box with `pytest -m gpu`).  Everything goes through the C ABI
(librecnet_b200.so); the checker is the CPU oracle and the reference-generated golden fixtures.

Tolerances (BASELINE.json north_star): fp32 1e-3 relative, bf16 2e-2 relative, greedy ids bit-exact in fp32.
"relative" = max|a-b| / max|b| per tensor.
"""
import pytest
import torch

import recnet_b200
from recnet_b200 import _lib as L, ops
from recnet_b200 import functional as Fn
from recnet_b200 import train as T
from recnet_b200 import eval as E
from oracle import recnet_oracle as O
from tests.golden_util import load_golden

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-3, "bf16": 2e-2}


def dev():
    return torch.device("cuda:0")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def configure(m, precision, kind="local"):
    C = T.C
    C.decoder_model, C.reconstructor_model = m["dec_model"], m["rec_model"]
    C.batch_size, C.caption_max_len = m["B"], m["cap_len"]
    C.encoder_output_len, C.encoder_output_size = m["T"], m["E"]
    C.decoder_n_layers, C.decoder_hidden_size, C.decoder_attn_size = m["dec_layers"], m["H"], m["A"]
    C.embedding_size = m["EMB"]
    C.reconstructor_n_layers, C.reconstructor_hidden_size, C.reconstructor_attn_size = m["rec_layers"], m["E"], m["A"]
    C.reconstructor_type = kind if kind != "none" else "local"
    C.precision, C.device = precision, "cuda"


def build(m, precision, kind, dec_sd, rec_sd):
    configure(m, precision, kind)
    dec = T.build_decoder(m["V"])
    dec["model"].load_state_dict({k: v.float() for k, v in dec_sd.items()})
    dec["model"].eval()
    rec = None
    if kind != "none":
        rec = T.build_reconstructor()
        rec["model"].load_state_dict({k: v.float() for k, v in rec_sd.items()})
        rec["model"].eval()
    return dec, rec


# ---------------------------------------------------------------------------------------------------------------
# kernels
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", [L.PREC_FP32, L.PREC_BF16])
@pytest.mark.parametrize("tA", [False, True])
@pytest.mark.parametrize("tB", [False, True])
@pytest.mark.parametrize("shape", [(100, 2048, 2048), (100, 6144, 2048), (100, 128, 1536), (3100, 4188, 512), (2048, 2048, 3100),
                                   (8, 8, 8), (300, 72, 200), (128, 1536, 2800), (1312, 264, 136)])
def test_gemm_all_operand_layouts(prec, tA, tB, shape):
    M, N, K = shape
    if prec == L.PREC_BF16 and (((M if tA else K) % 8) or ((N if tB else K) % 8)):
        pytest.skip("TMA needs a 16-byte row pitch")
    dt = torch.bfloat16 if prec == L.PREC_BF16 else torch.float32
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn((K, M) if tA else (M, K), generator=g).to(dev()).to(dt)
    B = torch.randn((K, N) if tB else (N, K), generator=g).to(dev()).to(dt)
    bias = torch.randn(N, generator=g).to(dev())
    ref = (A.double().t() if tA else A.double()) @ (B.double() if tB else B.double().t()) + bias.double()
    out = ops.gemm(prec, A, tA, B, tB, bias=bias)
    assert rel(out, ref) < 1e-5          # same (bf16-rounded) operands on both sides: only accumulation order differs
    nsplit = 2 if K >= 128 else 1
    if nsplit > 1:
        parts = ops.gemm(prec, A, tA, B, tB, bias=bias, splits=nsplit)
        assert rel(parts.sum(0), ref) < 1e-5
    if prec == L.PREC_BF16:
        for bn in (64, 128):
            assert rel(ops.gemm(prec, A, tA, B, tB, bias=bias, bn_hint=bn), ref) < 1e-5
        for bn in (1128, 1256, 2128, 2256):          # gemm_tc2.cuh: persistent kernel, single CTAs / CTA pairs (cta_group::2)
            assert rel(ops.gemm(prec, A, tA, B, tB, bias=bias, bn_hint=bn), ref) < 1e-5, bn
            acc = torch.ones(M, N, device=dev())
            ops.gemm(prec, A, tA, B, tB, bias=bias, out=acc, accumulate=True, bn_hint=bn)
            assert rel(acc, ref + 1.0) < 1e-5, bn
            if N % 8 == 0:                           # operand-typed (bf16) output straight from the epilogue
                cb = torch.full((M, N), 7.0, dtype=torch.bfloat16, device=dev())
                L.check(L.lib().recnet_gemm(prec, A.data_ptr(), A.stride(0), int(tA), B.data_ptr(), B.stride(0), int(tB), None, 0,
                                            cb.data_ptr(), N, bias.data_ptr(), M, N, K, 1, 0, 0, bn,
                                            torch.cuda.current_stream().cuda_stream), "recnet_gemm")
                assert rel(cb.double(), ref) < 4e-3, bn


def test_gemm_rejects_misaligned_bf16_pitch():
    A = torch.randn(16, 36, device=dev()).to(torch.bfloat16)
    B = torch.randn(8, 36, device=dev()).to(torch.bfloat16)
    with pytest.raises(RuntimeError, match="ALIGNMENT"):
        ops.gemm(L.PREC_BF16, A, False, B, False)


@pytest.mark.parametrize("prec,tol", [(L.PREC_FP32, 1e-5), (L.PREC_BF16, 2e-2)])
@pytest.mark.parametrize("B,Tn,A,D", [(100, 28, 128, 1536), (7, 31, 16, 64), (3, 1, 8, 8)])
def test_attention_kernel_fwd_bwd(prec, tol, B, Tn, A, D):
    g = torch.Generator().manual_seed(B * Tn)
    mk = lambda *s: torch.randn(*s, generator=g).to(dev())
    Wh, Uv, b, w, V = mk(B, A), mk(B, Tn, A), mk(A), mk(1, A) * 0.3, mk(B, Tn, D)
    ins = [t.clone().requires_grad_(True) for t in (Wh, Uv, b, w, V)]
    out = ops.additive_attention(*ins, prec)
    gout = mk(B, D)
    out.backward(gout)
    refs = [t.clone().double().requires_grad_(True) for t in (Wh, Uv, b, w, V)]
    rWh, rUv, rb, rw, rV = refs
    e = torch.tanh(rWh.unsqueeze(1) + rUv + rb) @ rw.t()
    ref = (e * rV).mean(1)                                   # decoder.py:56-61: mean, no softmax
    ref.backward(gout.double())
    assert rel(out, ref) < tol
    for a, r, name in zip(ins, refs, ("Wh", "Uv", "b", "w", "V")):
        assert rel(a.grad, r.grad) < tol, name


@pytest.mark.parametrize("prec,tol", [(L.PREC_FP32, 1e-5), (L.PREC_BF16, 2e-2)])
@pytest.mark.parametrize("B,H", [(100, 512), (100, 1536), (5, 8)])
def test_lstm_cell_kernel_fwd_bwd(prec, tol, B, H):
    g = torch.Generator().manual_seed(H)
    pre = torch.randn(B, 4 * H, generator=g).to(dev()).requires_grad_(True)
    c0 = torch.randn(B, H, generator=g).to(dev()).requires_grad_(True)
    h, c = ops.lstm_cell(pre, c0, prec)
    gh, gc = torch.randn(B, H, generator=g).to(dev()), torch.randn(B, H, generator=g).to(dev())
    torch.autograd.backward([h, c], [gh, gc])
    p2, c2 = pre.detach().double().requires_grad_(True), c0.detach().double().requires_grad_(True)
    i, f, gg, o = p2.chunk(4, dim=1)                         # PyTorch LSTM gate order i,f,g,o
    cr = torch.sigmoid(f) * c2 + torch.sigmoid(i) * torch.tanh(gg)
    hr = torch.sigmoid(o) * torch.tanh(cr)
    torch.autograd.backward([hr, cr], [gh.double(), gc.double()])
    fwd_tol = 1e-5 if prec == L.PREC_FP32 else 2e-3        # bf16 build: MUFU tanh.approx (~2^-11) in the activations
    assert rel(h, hr) < fwd_tol and rel(c, cr) < fwd_tol
    assert rel(pre.grad, p2.grad) < tol and rel(c0.grad, c2.grad) < tol


# ---------------------------------------------------------------------------------------------------------------
# sequence level vs the reference-generated golden fixtures
# ---------------------------------------------------------------------------------------------------------------
GOLDEN = [("tiny_lstm", "fp32"), ("tiny_lstm_ragged", "fp32"), ("small_lstm", "fp32"), ("tiny_lstm", "bf16"), ("small_lstm", "bf16")]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_gru_decoder_matches_reference_golden(precision):
    """The reference's DEFAULT decoder cell (config.py:31 decoder_model = "GRU"): loss, hiddens, gradients, greedy ids and
    the per-step Decoder.forward against the reference-generated fixture."""
    g = load_golden("tiny_gru")
    m = dict(g["meta"], rec_model="LSTM")
    tol = TOL[precision]
    dec, _ = build(m, precision, "none", g["dec"], {})
    feats, targets = g["feats"].float().to(dev()), g["targets"].to(dev())
    dloss, hiddens, _ = T.forward_decoder(dec, feats, targets, targets > 0, 1.0)
    assert rel(dloss, torch.tensor(g["dec_loss"])) < tol and rel(hiddens, g["hiddens"]) < tol
    dloss.backward()
    for k, ref in g["grads"]["none"].items():
        assert rel(dict(dec["model"].named_parameters())[k[4:]].grad, ref) < tol, k
    B, H = feats.shape[0], m["H"]
    tok = torch.full((1, B), 1, dtype=torch.long, device=dev())
    with torch.no_grad():
        logits, h1 = dec["model"](tok, torch.zeros(1, B, H, device=dev()), feats)      # GRU hidden is a single tensor
    assert h1.shape == (1, B, H) and rel(logits, g["step0_logits"]) < tol
    if precision == "fp32":
        ids, n = dec["model"].greedy(feats, m["cap_len"] + 1)
        assert torch.equal(ids[: int(n)].cpu(), g["greedy_ids"])


@pytest.mark.parametrize("kind", ["none", "global", "local"])
@pytest.mark.parametrize("name,precision", GOLDEN)
def test_losses_hiddens_grads_match_reference_golden(name, precision, kind):
    g = load_golden(name)
    tol = TOL[precision]
    dec, rec = build(g["meta"], precision, kind, g["dec"], g.get(kind, {}))
    feats, targets = g["feats"].float().to(dev()), g["targets"].to(dev())
    dloss, hiddens, _ = T.forward_decoder(dec, feats, targets, targets > 0, 1.0)
    assert rel(dloss, torch.tensor(g["dec_loss"])) < tol
    assert hiddens.shape == g["hiddens"].shape and rel(hiddens, g["hiddens"]) < tol     # (L, 1, B, H), early break honoured
    loss = dloss
    if rec is not None:
        rloss = T.forward_reconstructor_for(kind)(hiddens, feats, rec)
        assert rel(rloss, torch.tensor(g[kind + "_loss"])) < tol
        loss = dloss + 1.0 * rloss                                                      # train.py:260
    loss.backward()
    Fn.check_loop_status()
    for k, ref in g["grads"][kind].items():
        owner, key = k.split(".", 1)
        p = dict((dec if owner == "dec" else rec)["model"].named_parameters())[key]
        assert rel(p.grad, ref) < tol, k


@pytest.mark.parametrize("name", ["tiny_lstm", "tiny_lstm_ragged", "small_lstm"])
def test_greedy_ids_bit_exact_fp32(name):
    g = load_golden(name)
    m = g["meta"]
    dec, _ = build(m, "fp32", "none", g["dec"], {})
    feats = g["feats"].float().to(dev())
    B = feats.shape[0]
    T.C.batch_size = B
    tok = torch.full((1, B), 1, dtype=torch.long, device=dev())
    hid = (torch.zeros(1, B, m["H"], device=dev()), torch.zeros(1, B, m["H"], device=dev()))
    ids = E.greedy_search(T.C, dec["model"], tok, hid, feats)              # eval.greedy_search signature
    assert torch.equal(torch.tensor(ids), g["greedy_ids"])                  # same ids AND same stop step (eval.py:30)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_single_step_module_forward_matches_reference(precision):
    g = load_golden("small_lstm")
    m = g["meta"]
    dec, _ = build(m, precision, "none", g["dec"], {})
    feats = g["feats"].float().to(dev())
    B = feats.shape[0]
    tok = torch.full((1, B), 1, dtype=torch.long, device=dev())
    hid = (torch.zeros(1, B, m["H"], device=dev()), torch.zeros(1, B, m["H"], device=dev()))
    with torch.no_grad():
        logits, (h, c) = dec["model"](tok, hid, feats)                      # Decoder.forward(input, hidden, encoder_outputs)
    assert logits.shape == (B, m["V"]) and h.shape == (1, B, m["H"])
    assert rel(logits, g["step0_logits"]) < TOL[precision]


def test_per_step_api_loop_is_differentiable_and_agrees_with_sequence_path():
    """Drive our modules exactly like the reference's train.forward_decoder does (Python loop over Decoder.forward +
    nn.CrossEntropyLoss) and compare loss and gradients with the one-call sequence path."""
    g = load_golden("tiny_lstm")
    m = g["meta"]
    feats, targets = g["feats"].float().to(dev()), g["targets"].to(dev())
    masks = targets > 0
    dec, _ = build(m, "fp32", "none", g["dec"], {})
    model = dec["model"]
    B = feats.shape[0]
    tok = torch.full((1, B), 1, dtype=torch.long, device=dev())
    hid = (torch.zeros(1, B, m["H"], device=dev()), torch.zeros(1, B, m["H"], device=dev()))
    loss, n = 0, 0
    for t in range(m["cap_len"] + 1):
        out, hid = model(tok, hid, feats)
        tok = targets[t].view(1, -1)
        loss = loss + dec["loss"](out[masks[t]], targets[t][masks[t]])
        n = n + masks[t].sum()
        if t == m["cap_len"] or not bool(masks[t + 1].any()):
            break
    loss = loss / n + dec["lambda_reg"] * sum(torch.norm(p) for p in model.parameters())
    loss.backward()
    assert rel(loss, torch.tensor(g["dec_loss"])) < 1e-3
    for k, ref in g["grads"]["none"].items():
        assert rel(dict(model.named_parameters())[k[4:]].grad, ref) < 1e-3, k


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json sizes: parity against the oracle + size-independent properties
# ---------------------------------------------------------------------------------------------------------------
FULL = dict(B=100, T=28, E=1536, H=512, A=128, EMB=468, V=4188, cap_len=30, dec_layers=1, rec_layers=1, dec_model="LSTM", rec_model="LSTM")


def _full_inputs(B=100, seed=1234):
    return O.synthetic_batch(B, FULL["T"], FULL["E"], FULL["V"], 30, seed=seed)


@pytest.fixture(scope="module")
def full_oracle():
    out = {}
    feats, targets, masks = _full_inputs()
    P = O.init_decoder_params(FULL["V"], FULL["EMB"], FULL["E"], FULL["H"], FULL["A"], seed=0)
    for kind in ("local", "global"):
        Q = O.init_reconstructor_params(kind, FULL["H"], FULL["E"], FULL["A"], seed=1)
        Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
        Qr = {k: v.clone().requires_grad_(True) for k, v in Q.items()}
        dl, hid, _, aux = O.forward_decoder(Pr, feats, targets, masks)
        fn = O.forward_local_reconstructor if kind == "local" else O.forward_global_reconstructor
        rl, _ = fn(Qr, hid, feats)
        (dl + rl).backward()
        out[kind] = dict(P=P, Q=Q, dl=dl.detach(), rl=rl.detach(), hid=hid.detach(), logits=aux["logits"].detach(),
                         gP={k: v.grad for k, v in Pr.items()}, gQ={k: v.grad for k, v in Qr.items()})
    out["inputs"] = (feats, targets, masks)
    return out


@pytest.mark.parametrize("kind", ["local", "global"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_full_size_parity_against_oracle(full_oracle, precision, kind):
    o = full_oracle[kind]
    feats, targets, masks = (t.to(dev()) for t in full_oracle["inputs"])
    dec, rec = build(FULL, precision, kind, o["P"], o["Q"])
    tol = TOL[precision]
    dloss, hiddens, _ = T.forward_decoder(dec, feats, targets, masks, 1.0)
    rloss = T.forward_reconstructor_for(kind)(hiddens, feats, rec)
    (dloss + rloss).backward()
    Fn.check_loop_status()
    assert rel(dloss, o["dl"]) < tol and rel(rloss, o["rl"]) < tol and rel(hiddens, o["hid"]) < tol
    for k, p in dec["model"].named_parameters():
        assert rel(p.grad, o["gP"][k]) < tol, k
    for k, p in rec["model"].named_parameters():
        assert rel(p.grad, o["gQ"][k]) < tol, k
    tokens_in = torch.cat((torch.ones(1, 100, dtype=torch.long, device=dev()), targets[:30]), 0)
    logits, _ = dec["model"].teacher_forced_logits(tokens_in, feats)
    assert rel(logits, o["logits"]) < tol


# shape variants that walk the generic branches of the fused decoder kernels and the lean attention kernels:
#   msrvtt : 40 frames x 3584-d (BASELINE config 5 feature shape) -> second frame chunk / NF = 4 frame groups, 3584-wide VW GEMM
#   wide_a : attention size 256 -> two float4 chunks per lane (NCH = 2)
#   odd_h  : hidden 328 (not a multiple of the 128-unit tile), 19 frames, short captions
VARIANTS = {
    "msrvtt": dict(B=12, T=40, E=3584, H=512, A=128, EMB=468, V=1000, cap_len=12),
    "wide_a": dict(B=9, T=28, E=256, H=256, A=256, EMB=64, V=400, cap_len=8),
    "odd_h": dict(B=7, T=19, E=264, H=328, A=72, EMB=52, V=333, cap_len=6),
}


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_shape_variants_against_oracle(variant, precision):
    m = dict(VARIANTS[variant], dec_layers=1, rec_layers=1, dec_model="LSTM", rec_model="LSTM")
    feats, targets, masks = O.synthetic_batch(m["B"], m["T"], m["E"], m["V"], m["cap_len"], seed=5)
    P = O.init_decoder_params(m["V"], m["EMB"], m["E"], m["H"], m["A"], seed=2)
    Q = O.init_reconstructor_params("local", m["H"], m["E"], m["A"], seed=3)
    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    Qr = {k: v.clone().requires_grad_(True) for k, v in Q.items()}
    dl, hid, _, _ = O.forward_decoder(Pr, feats, targets, masks, caption_max_len=m["cap_len"])
    rl, _ = O.forward_local_reconstructor(Qr, hid, feats)
    (dl + rl).backward()
    dec, rec = build(m, precision, "local", P, Q)
    tol = TOL[precision]
    f, t, k = feats.to(dev()), targets.to(dev()), masks.to(dev())
    dloss, hiddens, _ = T.forward_decoder(dec, f, t, k, 1.0)
    rloss = T.forward_local_reconstructor(hiddens, f, rec)
    (dloss + rloss).backward()
    assert rel(dloss, dl.detach()) < tol and rel(rloss, rl.detach()) < tol and rel(hiddens, hid.detach()) < tol
    for name, p in dec["model"].named_parameters():
        assert rel(p.grad, Pr[name].grad) < tol, name
    for name, p in rec["model"].named_parameters():
        assert rel(p.grad, Qr[name].grad) < tol, name


def test_full_size_greedy_bit_exact_fp32_batch1024_shape():
    """BASELINE config 4 itself: greedy, decoder-only, batch 1024, 28 frames, max len 30 -- every id and the stop step, bit-exact."""
    B = 1024
    feats, _, _ = _full_inputs(B, seed=77)
    P = O.init_decoder_params(FULL["V"], FULL["EMB"], FULL["E"], FULL["H"], FULL["A"], seed=0)
    P["out.bias"][0] += 3.0                     # random weights never emit <PAD>; nudge it so the PAD-stop rule is exercised
    ref = O.greedy_search(P, feats)
    dec, _ = build(dict(FULL, B=B), "fp32", "none", P, {})
    ids, n = dec["model"].greedy(feats.to(dev()), 31)
    assert int(n) == ref.shape[0]
    assert torch.equal(ids[: int(n)].cpu(), ref)


def test_samples_are_independent_and_permutation_equivariant_at_full_size():
    """Size-independent property at BASELINE sizes: the path shards by sample (no cross-sample arithmetic anywhere),
    so permuting the batch permutes the decoder states bit-exactly."""
    feats, targets, masks = (t.to(dev()) for t in _full_inputs())
    P = O.init_decoder_params(FULL["V"], FULL["EMB"], FULL["E"], FULL["H"], FULL["A"], seed=0)
    dec, _ = build(FULL, "bf16", "none", P, {})
    perm = torch.randperm(100, generator=torch.Generator().manual_seed(1)).to(dev())
    _, h1, _ = T.forward_decoder(dec, feats, targets, masks, 1.0)
    _, h2, _ = T.forward_decoder(dec, feats[perm], targets[:, perm], masks[:, perm], 1.0)
    assert torch.equal(h1[:, :, perm], h2)      # bit-exact: no cross-sample arithmetic anywhere on the path


def test_gradient_scales_linearly_with_upstream_gradient():
    g = load_golden("small_lstm")
    dec, rec = build(g["meta"], "fp32", "local", g["dec"], g["local"])
    feats, targets = g["feats"].float().to(dev()), g["targets"].to(dev())

    def grads(scale):
        for mod in (dec["model"], rec["model"]):
            mod.zero_grad(set_to_none=True)
        dl, hid, _ = T.forward_decoder(dec, feats, targets, targets > 0, 1.0)
        rl = T.forward_local_reconstructor(hid, feats, rec)
        ((dl + rl) * scale).backward()
        return [p.grad.clone() for mod in (dec["model"], rec["model"]) for p in mod.parameters()]

    g1, g3 = grads(1.0), grads(3.0)
    for a, b in zip(g1, g3):
        assert rel(b, 3.0 * a) < 1e-5


def test_train_mode_dropout_is_seeded_and_statistically_sane():
    g = load_golden("small_lstm")
    dec, rec = build(g["meta"], "bf16", "local", g["dec"], g["local"])
    feats, targets = g["feats"].float().to(dev()), g["targets"].to(dev())
    dec["model"].train(); rec["model"].train()

    def run(seed):
        dec["model"].seed_dropout(seed); rec["model"].seed_dropout(seed)
        dl, hid, _ = T.forward_decoder(dec, feats, targets, targets > 0, 1.0)
        rl = T.forward_local_reconstructor(hid, feats, rec)
        return float(dl), float(rl)

    a, b, c = run(1), run(1), run(2)
    assert a == b and a != c                    # Philox masks: same seed -> same masks, different seed -> different
    assert all(map(lambda x: x == x and abs(x) < 1e3, a + c))
    eval_loss = g["dec_loss"]
    assert abs(a[0] - eval_loss) / eval_loss < 0.5      # inverted dropout keeps the expectation: same ballpark as eval


def test_data_parallel_gradient_equals_mean_of_shard_gradients():
    """'Fake world' on one GPU: all-reduced gradient (average) == mean over shards of per-shard gradients, which is what
    parallel.GradAllReducer computes; checked against the oracle on the shards."""
    m = dict(FULL, B=8, E=64, H=32, A=16, EMB=24, V=101, T=6, cap_len=7)
    feats, targets, masks = O.synthetic_batch(16, m["T"], m["E"], m["V"], m["cap_len"], seed=3, full_length_first=False)
    P = O.init_decoder_params(m["V"], m["EMB"], m["E"], m["H"], m["A"], seed=0)
    Q = O.init_reconstructor_params("local", m["H"], m["E"], m["A"], seed=1)
    acc_ours, acc_ref = None, None
    for r in range(2):
        sl = slice(r * 8, (r + 1) * 8)
        f, t, mk = feats[sl], targets[:, sl], masks[:, sl]
        Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
        Qr = {k: v.clone().requires_grad_(True) for k, v in Q.items()}
        dl, hid, _, _ = O.forward_decoder(Pr, f, t, mk, caption_max_len=m["cap_len"])
        rl, _ = O.forward_local_reconstructor(Qr, hid, f)
        (dl + rl).backward()
        ref = [Pr[k].grad for k in Pr] + [Qr[k].grad for k in Qr]
        dec, rec = build(m, "fp32", "local", P, Q)
        dloss, hiddens, _ = T.forward_decoder(dec, f.to(dev()), t.to(dev()), mk.to(dev()), 1.0)
        (dloss + T.forward_local_reconstructor(hiddens, f.to(dev()), rec)).backward()
        ours = [dict(dec["model"].named_parameters())[k].grad for k in P] + [dict(rec["model"].named_parameters())[k].grad for k in Q]
        acc_ours = ours if acc_ours is None else [a + b for a, b in zip(acc_ours, ours)]
        acc_ref = ref if acc_ref is None else [a + b for a, b in zip(acc_ref, ref)]
    for a, b in zip(acc_ours, acc_ref):
        assert rel(a / 2, b / 2) < 1e-3


def test_whole_train_step_runs_under_cuda_graph_and_updates_weights():
    g = load_golden("small_lstm")
    dec, rec = build(g["meta"], "bf16", "local", g["dec"], g["local"])
    feats, targets = g["feats"].float().to(dev()), g["targets"].to(dev())
    L_steps = g["hiddens"].shape[0]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            T.train_step(dec, rec, feats, targets, n_steps=L_steps)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    out = torch.zeros((), device=dev())
    with torch.cuda.graph(graph):
        loss, _, _ = T.train_step(dec, rec, feats, targets, n_steps=L_steps)
        out.copy_(loss.detach())
    w0 = dec["model"].out.weight.detach().clone()
    losses = []
    for _ in range(5):
        graph.replay()
        losses.append(float(out))
    assert all(x == x for x in losses) and len(set(losses)) > 1            # fresh dropout masks every replay
    assert not torch.equal(w0, dec["model"].out.weight)                     # Adam step inside the graph moved the weights


@pytest.mark.parametrize("kind", ["global", "local"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_gru_decoder_plus_gru_reconstructor_match_reference_golden(precision, kind):
    """GRU everywhere (decoder_model = reconstructor_model = "GRU"): joint loss and every gradient against the fixture."""
    g = load_golden("tiny_gru")
    tol = TOL[precision]
    dec, rec = build(g["meta"], precision, kind, g["dec"], g[kind])
    feats, targets = g["feats"].float().to(dev()), g["targets"].to(dev())
    dloss, hiddens, _ = T.forward_decoder(dec, feats, targets, targets > 0, 1.0)
    rloss = T.forward_reconstructor_for(kind)(hiddens, feats, rec)
    assert rel(rloss, torch.tensor(g[kind + "_loss"])) < tol
    (dloss + rloss).backward()
    for k, ref in g["grads"][kind].items():
        owner, key = k.split(".", 1)
        p = dict((dec if owner == "dec" else rec)["model"].named_parameters())[key]
        assert rel(p.grad, ref) < tol, k


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_layer_decoder_matches_reference_golden(precision):
    """Stacked decoder (decoder_n_layers = 2, BASELINE config 5 family): loss, hiddens (L,2,B,H), gradients of every layer,
    greedy ids and the per-step forward against the reference-generated fixture."""
    g = load_golden("tiny_lstm_2layer")
    m = g["meta"]
    tol = TOL[precision]
    dec, _ = build(m, precision, "none", g["dec"], {})
    feats, targets = g["feats"].float().to(dev()), g["targets"].to(dev())
    dloss, hiddens, _ = T.forward_decoder(dec, feats, targets, targets > 0, 1.0)
    assert hiddens.shape == g["hiddens"].shape
    assert rel(dloss, torch.tensor(g["dec_loss"])) < tol and rel(hiddens, g["hiddens"]) < tol
    dloss.backward()
    for k, ref in g["grads"]["none"].items():
        assert rel(dict(dec["model"].named_parameters())[k[4:]].grad, ref) < tol, k
    B, H = feats.shape[0], m["H"]
    tok = torch.full((1, B), 1, dtype=torch.long, device=dev())
    z = torch.zeros(2, B, H, device=dev())
    with torch.no_grad():
        logits, (h1, c1) = dec["model"](tok, (z, z.clone()), feats)
    assert h1.shape == (2, B, H) and rel(logits, g["step0_logits"]) < tol
    if precision == "fp32":
        ids, n = dec["model"].greedy(feats, m["cap_len"] + 1)
        assert torch.equal(ids[: int(n)].cpu(), g["greedy_ids"])


@pytest.mark.parametrize("kind", ["global", "local"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_reconstructors_over_two_layer_decoder_match_reference_golden(precision, kind):
    """Reconstructors fed by a stacked decoder's (L,2,B,H) hiddens: the global one pools over time AND layers and reads
    layer 0 (global_reconstructor.py:33-40); the local one runs one LSTM pseudo-step per decoder layer and projects the
    first (local_reconstructor.py:42-54).  Joint loss and every gradient (decoder layers included)."""
    g = load_golden("tiny_lstm_2layer")
    tol = TOL[precision]
    dec, rec = build(g["meta"], precision, kind, g["dec"], g[kind])
    feats, targets = g["feats"].float().to(dev()), g["targets"].to(dev())
    dloss, hiddens, _ = T.forward_decoder(dec, feats, targets, targets > 0, 1.0)
    assert hiddens.shape[1] == 2
    rloss = T.forward_reconstructor_for(kind)(hiddens, feats, rec)
    assert rel(rloss, torch.tensor(g[kind + "_loss"])) < tol
    (dloss + rloss).backward()
    Fn.check_loop_status()
    for k, ref in g["grads"][kind].items():
        owner, key = k.split(".", 1)
        p = dict((dec if owner == "dec" else rec)["model"].named_parameters())[key]
        assert rel(p.grad, ref) < tol, k


class _Vocab:
    def __init__(self, n):
        self.n_vocabs = n
        self.word2idx = {'<PAD>': 0, '<SOS>': 1, '<EOS>': 2}


@pytest.mark.parametrize("name", ["tiny_lstm", "tiny_gru"])
def test_beam_search_matches_oracle_restatement(name):
    """eval.beam_search (CUDA-only in the reference, eval.py:39,57) against the oracle's CPU restatement of the same
    algorithm; fp32 build, beam 3 and 5: identical top-1 sequences."""
    g = load_golden(name)
    m = dict(g["meta"], rec_model="LSTM")
    P = {k: v.float() for k, v in g["dec"].items()}
    P["out.bias"] = P["out.bias"].clone()
    P["out.bias"][2] += 1.5                      # make <EOS> likely enough that the length-normalisation path is exercised
    dec, _ = build(m, "fp32", "none", P, {})
    feats = g["feats"].float().to(dev())
    B, H = feats.shape[0], m["H"]
    T.C.batch_size = B
    for width in (3, 5):
        ref = O.beam_search(P, g["feats"].float(), width, model_name=m["dec_model"], n_layers=1, caption_max_len=m["cap_len"])
        tok = torch.full((1, B), 1, dtype=torch.long, device=dev())
        z = torch.zeros(1, B, H, device=dev())
        hid = (z, z.clone()) if m["dec_model"] == "LSTM" else z
        got = E.beam_search(T.C, width, _Vocab(m["V"]), dec["model"], tok, hid, feats)
        assert got == ref, (width, got, ref)


def test_checkpoint_roundtrip_reference_layout(tmp_path):
    g = load_golden("tiny_lstm")
    dec, rec = build(g["meta"], "fp32", "local", g["dec"], g["local"])
    path = str(tmp_path / "5_checkpoint.tar")
    E.save_checkpoint(path, 5, dec, rec, loss=torch.tensor(1.0), config=None)
    blob = torch.load(path, weights_only=False)
    assert set(blob) == {'iteration', 'dec', 'rec', 'dec_opt', 'rec_opt', 'loss', 'config'}       # train.py:404-412
    dec2, rec2 = build(g["meta"], "fp32", "local", {k: torch.zeros_like(v) for k, v in g["dec"].items()},
                       {k: torch.zeros_like(v) for k, v in g["local"].items()})
    assert E.load_checkpoint(path, dec2, rec2) == 5
    for k, v in dec["model"].state_dict().items():
        assert torch.equal(v, dec2["model"].state_dict()[k])
