#!/usr/bin/env python
"""GPU timeline of ONE captured train step (CUDA-graph replay) from torch.profiler / CUPTI: every kernel with stream, start and
duration, so that what runs concurrently (background lane, side streams, all-reduce) can be read off.
    python tools/step_trace.py [--recon local] [--out gpurun_out/step_trace.txt]
Prints kernels of the last replay in start order: t_start_us  dur_us  stream  name."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--recon", default="local")
    ap.add_argument("--out", default="gpurun_out/step_trace.txt")
    args = ap.parse_args()
    import bench
    from recnet_b200 import train as T
    from recnet_b200.data import synthetic_batch
    from torch.profiler import ProfilerActivity, profile
    s = bench.SHAPE
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:                                   # under torchrun: the data-parallel step, rank 0's timeline
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    C = T.C
    C.decoder_model = C.reconstructor_model = "LSTM"
    C.batch_size, C.caption_max_len, C.encoder_output_len, C.encoder_output_size = s["B"], s["cap"], s["T"], s["E"]
    C.decoder_n_layers, C.decoder_hidden_size, C.decoder_attn_size, C.embedding_size = 1, s["H"], s["A"], s["EMB"]
    C.reconstructor_n_layers, C.reconstructor_hidden_size, C.reconstructor_attn_size = 1, s["R"], s["A"]
    C.use_recon = args.recon != "none"
    C.reconstructor_type = args.recon if C.use_recon else "local"
    C.precision, C.device = "bf16", f"cuda:{local}"
    torch.manual_seed(0)
    dec = T.build_decoder(s["V"])
    rec = T.build_reconstructor() if C.use_recon else None
    L = s["cap"] + 1
    feats, targets, _ = synthetic_batch(s["B"], s["T"], s["E"], s["V"], s["cap"], seed=1234 + rank)
    feats, targets = feats.to(dev), targets.to(dev)
    reducer = None
    if world > 1:
        from recnet_b200.parallel import broadcast_parameters, make_reducer
        broadcast_parameters([dec["model"]] + ([rec["model"]] if rec else []))
        reducer = make_reducer(([rec["model"]] if rec else []) + [dec["model"]])

    def step():
        if reducer is not None:
            reducer.start_iteration()
        T.train_step(dec, rec, feats, targets, n_steps=L, reducer=reducer)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    prio = int(os.environ.get("RECNET_CAPTURE_PRIO", "0"))
    with torch.cuda.graph(g, stream=torch.cuda.Stream(priority=prio), capture_error_mode="thread_local" if world > 1 else "global"):
        step()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
    if rank != 0 and os.environ.get("RECNET_TRACE_ALL_RANKS", "0") != "1":
        torch.distributed.barrier()
        os._exit(0)
    if rank != 0:
        args.out = args.out.replace(".txt", f"_rank{rank}.txt")
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    if not evs:
        print("no CUDA events recorded")
        return
    # split into replays by the largest gaps
    n = len(evs) // 3
    last = evs[2 * n:]
    t0 = last[0].time_range.start if os.environ.get("RECNET_TRACE_ABS", "0") != "1" else 0
    lines = []
    for e in last:
        stream = getattr(e, "stream", None)
        lines.append(f"{e.time_range.start - t0:10.1f} {e.time_range.end - e.time_range.start:8.1f} {stream!s:>4} {e.name[:110]}")
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(f"{len(last)} kernels, span {last[-1].time_range.end - t0:.1f} us -> {args.out}")
    if world > 1:
        torch.distributed.barrier()
        os._exit(0)       # NCCL objects captured in a live graph: see bench.py


if __name__ == "__main__":
    main()
