"""2+ ranks (torchrun): the symmetric-memory gradient all-reduce must give the same averaged gradients as the NCCL path."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recnet_b200  # noqa: E402
from recnet_b200.parallel import GradAllReducer  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))


class M(torch.nn.Module):
    def __init__(self, n):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(n, device="cuda"))


mods = [M(15_000_003), M(9_700_001)]
g = torch.Generator(device="cuda").manual_seed(100 + rank)
flats = [torch.randn(m.w.numel(), device="cuda", generator=g) for m in mods]
ref = [f.clone() for f in flats]
for r in ref:
    dist.all_reduce(r, op=dist.ReduceOp.AVG)
for mode in sys.argv[1:] or ["multimem", "two_shot"]:
    os.environ["RECNET_DP_SYMM"] = mode
    red = GradAllReducer(mods)
    bufs = [f.clone() for f in flats]
    red._symm_allreduce(bufs)
    torch.cuda.synchronize()
    err = max(float((b - r).abs().max() / r.abs().max()) for b, r in zip(bufs, ref))
    # timing of the bare collective path (pack + kernel + unpack)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        red._symm_allreduce(bufs)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"{mode}: max rel err vs NCCL AVG {err:.2e}; {e0.elapsed_time(e1) / 20 * 1e3:.0f} us per call (pack + all-reduce + unpack, 99 MB)", flush=True)
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
