"""Measure the per-phase overhead of the persistent loop kernel (empty phases, with / without grid barriers)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recnet_b200
from recnet_b200 import _lib as L
lib = L.lib()
dev = torch.device("cuda:0")
N = 400
scratch = torch.zeros(N * 1024 + 4096, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
for sync in (0, 1):
    for _ in range(3):
        L.check(lib.recnet_debug_loop_overhead(N, sync, scratch.data_ptr(), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        L.check(lib.recnet_debug_loop_overhead(N, sync, scratch.data_ptr(), st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"sync_after={sync}: {ms*1e3:.1f} us per launch of {N} phases -> {ms*1e3/N:.3f} us per phase; err flag {int(scratch[256:260].view(torch.int32))}")
