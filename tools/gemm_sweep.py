"""Sweep (N-tile width, split-K count) for the batched GEMMs of one train step (decoder + local reconstructor, MSVD shape) on
the B200: device time of GEMM (+ split-K reduce) per configuration, cold L2 (a 256 MB write between repetitions).  Output: one JSON
line per shape with every configuration's microseconds -- the data behind runtime.cuh:plan_gemm_full."""
import ctypes as C
import json
import sys

import torch

sys.path.insert(0, ".")
import recnet_b200
from recnet_b200 import _lib as L, ops

B, T, E, H, A, EMBp, V, Lr, R = 100, 28, 1536, 512, 128, 512, 4188, 31, 1536
LB, BT, SB, GR = Lr * B, B * T, 28 * B, 4 * R
SHAPES = {   # name: (M, N, K, transA, transB)
    "dec.Uv": (BT, A, E, 0, 0), "dec.Gx": (LB, 4 * H, EMBp, 0, 0), "dec.VW": (BT, 4 * H, E, 0, 0), "dec.logits": (LB, V, H, 0, 0),
    "dec.dHext": (LB, H, 4192, 0, 1), "dec.out_w": (V, H, LB, 1, 1), "dec.dW_ctx": (4 * H, E, BT, 1, 1), "dec.dW_hh": (4 * H, H, LB, 1, 1),
    "dec.dW_emb": (4 * H, 468, LB, 1, 1), "dec.dXe": (LB, 468, 4 * H, 0, 1), "dec.dW_a": (A, H, LB, 1, 1), "dec.dU": (A, E, BT, 1, 1),
    "rec.Uv": (LB, A, H, 0, 0), "rec.out": (SB, R, R, 0, 0), "rec.dHext": (SB, R, R, 0, 1), "rec.out_w": (R, R, SB, 1, 1),
    "rec.dW_ih": (GR, H, SB, 1, 1), "rec.dW_hh": (GR, R, SB, 1, 1), "rec.attn_W": (A, R, SB, 1, 1), "rec.attn_U": (A, H, LB, 1, 1),
    "rec.g_hid": (LB, H, A, 0, 1),
}
SCRATCH = 148 * 2 * 128 * 128
# per-step GEMMs of the two time loops (M = batch = 100): split-K partials are left for the consumer kernel (no reduce pass);
# timed warm (operands L2-resident, as inside the loop), 20 back-to-back launches per event pair
STEP_SHAPES = {
    "rec.gate": (100, 4 * R, H + R, 0, 0), "rec.dX": (100, H + R, 4 * R, 0, 1), "rec.Wh": (100, A, R, 0, 0), "rec.dQ": (100, R, A, 0, 1),
    "dec.gate": (100, A + 4 * H, H, 0, 0), "dec.dh": (100, H, A + 4 * H, 0, 1),
}


def sweep_steps(lib, dev, stream):
    for name, (M, N, K, tA, tB) in STEP_SHAPES.items():
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        b = (torch.randn(K, N, device=dev) if tB else torch.randn(N, K, device=dev)).to(torch.bfloat16)
        res = {}
        for bn in (64, 128, 256):
            if bn > 64 and N <= bn // 2:
                continue
            for splits in (1, 2, 3, 4, 6, 8, 9, 12, 16):
                if (K + 63) // 64 < splits:
                    continue
                part = torch.empty(splits, M, N, dtype=torch.float32, device=dev)

                def run():
                    L.check(lib.recnet_gemm(L.PREC_BF16, a.data_ptr(), a.stride(0), tA, b.data_ptr(), b.stride(0), tB, part.data_ptr(), N,
                                            None, 0, None, M, N, K, splits, M * N, 0, bn, stream), "gemm")
                for _ in range(3):
                    run()
                ts = []
                for _ in range(5):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(20):
                        run()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3 / 20)
                ts.sort()
                res[f"{bn}x{splits}"] = round(ts[2], 2)
        best = min((v, k) for k, v in res.items())
        print(json.dumps({"shape": name, "MNK": [M, N, K], "tA": tA, "tB": tB, "best": best[1], "best_us": best[0], "us": res}), flush=True)


def main():
    dev = torch.device("cuda:0")
    lib = L.lib()
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    only = sys.argv[1:]
    if only and only[0] == "--steps":
        return sweep_steps(lib, dev, stream)
    for name, (M, N, K, tA, tB) in SHAPES.items():
        if only and name not in only:
            continue
        Kp = (K + 7) // 8 * 8
        a = torch.randn((K, (M + 7) // 8 * 8) if tA else (M, Kp), device=dev).to(torch.bfloat16)
        b = torch.randn((K, (N + 7) // 8 * 8) if tB else (N, Kp), device=dev).to(torch.bfloat16)
        a = a[:, :M] if tA else a[:, :K]
        b = b[:, :N] if tB else b[:, :K]
        out = torch.empty(M, N, dtype=torch.float32, device=dev)
        res = {}
        for bn in (64, 128, 256):
            for splits in (1, 2, 3, 4, 6, 8):
                Np = (N + 3) // 4 * 4
                if splits > 1 and (splits * M * Np > SCRATCH or (K + 63) // 64 < splits):
                    continue
                part = torch.empty(splits, M, Np, dtype=torch.float32, device=dev) if splits > 1 else None

                def run():
                    if splits == 1:
                        L.check(lib.recnet_gemm(L.PREC_BF16, a.data_ptr(), a.stride(0), tA, b.data_ptr(), b.stride(0), tB, out.data_ptr(), N,
                                                None, 0, None, M, N, K, 1, 0, 0, bn, stream), "gemm")
                    else:
                        L.check(lib.recnet_gemm(L.PREC_BF16, a.data_ptr(), a.stride(0), tA, b.data_ptr(), b.stride(0), tB, part.data_ptr(), Np,
                                                None, 0, None, M, N, K, splits, M * Np, 0, bn, stream), "gemm")
                        L.check(lib.recnet_splitk_reduce(part.data_ptr(), splits, M * Np, Np, out.data_ptr(), N, M, N, 0, stream), "reduce")
                try:
                    for _ in range(2):
                        run()
                    ts = []
                    for _ in range(8):
                        flush.zero_()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(); run(); e1.record()
                        torch.cuda.synchronize()
                        ts.append(e0.elapsed_time(e1) * 1e3)
                    ts.sort()
                    res[f"{bn}x{splits}"] = round(ts[len(ts) // 2], 1)
                except RuntimeError as ex:
                    res[f"{bn}x{splits}"] = str(ex)[:40]
        for code in (1128, 1256, 2128, 2256):              # gemm_tc2.cuh: persistent, 1 CTA / CTA pair per tile
            def run2():
                L.check(lib.recnet_gemm(L.PREC_BF16, a.data_ptr(), a.stride(0), tA, b.data_ptr(), b.stride(0), tB, out.data_ptr(), N,
                                        None, 0, None, M, N, K, 1, 0, 0, code, stream), "gemm")
            try:
                for _ in range(2):
                    run2()
                ts = []
                for _ in range(8):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); run2(); e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                ts.sort()
                res[f"p{code // 1000}x{code % 1000}"] = round(ts[len(ts) // 2], 1)
            except RuntimeError as ex:
                res[f"p{code // 1000}x{code % 1000}"] = str(ex)[:40]
        best = min((v, k) for k, v in res.items() if isinstance(v, float))
        print(json.dumps({"shape": name, "MNK": [M, N, K], "tA": tA, "tB": tB, "best": best[1], "best_us": best[0], "us": res}), flush=True)


if __name__ == "__main__":
    main()
