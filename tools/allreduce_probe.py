"""Time NCCL all-reduce (AVG) of the two gradient buffers (61 MB local reconstructor, 38 MB decoder) alone."""
import os, sys, torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
dist.init_process_group("nccl", device_id=dev)
for mb in (61, 38, 99, 8):
    x = torch.randn(mb * 1000 * 1000 // 4, device=dev)
    for _ in range(5):
        dist.all_reduce(x, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_reduce(x, op=dist.ReduceOp.AVG)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    if rank == 0:
        print(f"all_reduce AVG {mb} MB x{world}: {ms*1e3:.1f} us  algbw {mb/ms:.1f} GB/s  busbw {mb/ms*2*(world-1)/world:.1f} GB/s", flush=True)
# the same two collectives captured in a CUDA graph
a = torch.randn(61 * 1000 * 1000 // 4, device=dev); b = torch.randn(38 * 1000 * 1000 // 4, device=dev)
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    dist.all_reduce(a, op=dist.ReduceOp.AVG); dist.all_reduce(b, op=dist.ReduceOp.AVG)
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, capture_error_mode="thread_local"):
    w1 = dist.all_reduce(a, op=dist.ReduceOp.AVG, async_op=True)
    w2 = dist.all_reduce(b, op=dist.ReduceOp.AVG, async_op=True)
    w1.wait(); w2.wait()
for _ in range(3):
    g.replay()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    g.replay()
e1.record(); torch.cuda.synchronize()
if rank == 0:
    print(f"all_reduce graph(61 MB + 38 MB): {e0.elapsed_time(e1)/20*1e3:.1f} us per replay", flush=True)
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
