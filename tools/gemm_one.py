"""Run ONE batched GEMM configuration of tools/gemm_sweep.py a few times (target of ncu captures):
    python tools/gemm_one.py dec.logits 2256 [reps]"""
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
import recnet_b200  # noqa: F401
from recnet_b200 import _lib as L
from gemm_sweep import SHAPES


def main():
    name, code = sys.argv[1], int(sys.argv[2])
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    M, N, K, tA, tB = SHAPES[name]
    dev = torch.device("cuda:0")
    lib = L.lib()
    Kp = (K + 7) // 8 * 8
    a = torch.randn((K, (M + 7) // 8 * 8) if tA else (M, Kp), device=dev).to(torch.bfloat16)
    b = torch.randn((K, (N + 7) // 8 * 8) if tB else (N, Kp), device=dev).to(torch.bfloat16)
    a = a[:, :M] if tA else a[:, :K]
    b = b[:, :N] if tB else b[:, :K]
    out = torch.empty(M, N, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(reps):
        L.check(lib.recnet_gemm(L.PREC_BF16, a.data_ptr(), a.stride(0), tA, b.data_ptr(), b.stride(0), tB, out.data_ptr(), N,
                                None, 0, None, M, N, K, 1, 0, 0, code, stream), "gemm")
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
