"""Correctness + timing of the own NVLink all-reduce kernel (csrc/allreduce.cuh) under torchrun (>= 2 GPUs of one box):
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dp_nvlink_check.py
1. gradients of one decoder + local-reconstructor step on per-rank data, averaged by recnet_allreduce_avg in place in symmetric
   memory, against NCCL's all_reduce(AVG) of a copy -- multicast (NVLS) path and peer-pointer path; 2. kernel time for the 99 MB."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    import bench
    import recnet_b200
    from recnet_b200 import train as T
    from recnet_b200.data import synthetic_batch
    from recnet_b200.parallel import NvlinkAllReducer, broadcast_parameters
    s = bench.SHAPE
    bench._configure(T.C, s, "local", "bf16", 1, f"cuda:{local}")
    torch.manual_seed(0)
    dec, rec = T.build_decoder(s["V"]), T.build_reconstructor()
    broadcast_parameters([dec["model"], rec["model"]])
    feats, targets, _ = synthetic_batch(s["B"], s["T"], s["E"], s["V"], s["cap"], seed=1234 + rank)
    feats, targets = feats.to(dev), targets.to(dev)
    out = {"world": world}
    for mode, ctas in (("multicast", 32), ("peer", 32), ("multicast16", 16), ("multicast64", 64)):
        os.environ["RECNET_AR_MULTICAST"] = "1" if mode.startswith("multicast") else "0"
        red = NvlinkAllReducer([rec["model"], dec["model"]], overlap=False, ctas=ctas)
        out[mode + "_has_multicast"] = bool(red.multicast)
        dec["model"].seed_dropout(5); rec["model"].seed_dropout(6)
        T.train_step(dec, rec, feats, targets, n_steps=s["cap"] + 1, optimizer_step=False)
        params = list(rec["model"].parameters()) + list(dec["model"].parameters())
        ref = [p.grad.detach().clone() for p in params]
        for g in ref:
            dist.all_reduce(g, op=dist.ReduceOp.AVG)
        red.start_iteration()
        red.wait()
        red.check()
        worst = 0.0
        for p, g in zip(params, ref):
            worst = max(worst, float((p.grad - g).abs().max() / (g.abs().max() + 1e-30)))
        out[mode + "_worst_rel_err_vs_nccl"] = worst
        # all ranks hold bitwise identical results?
        flat = red.buf.clone()
        mx, mn = flat.clone(), flat.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        out[mode + "_identical_on_all_ranks"] = bool(torch.equal(mx, mn))
        # timing: the kernel alone, 20 launches back to back
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            red.start_iteration(); red.wait()
        dist.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            red.start_iteration(); red.wait()
        e1.record()
        torch.cuda.synchronize()
        red.check()
        t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[mode + "_us_per_allreduce"] = round(float(t) * 1e3, 1)
        out[mode + "_bytes"] = sum(n for _, n in red.ranges) * 4
        red.remove()
        del red
    # NCCL reference timing for the same bytes
    n = sum(p.numel() for p in rec["model"].parameters()) + sum(p.numel() for p in dec["model"].parameters())
    x = torch.randn(n, device=dev)
    for _ in range(3):
        dist.all_reduce(x, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_reduce(x, op=dist.ReduceOp.AVG)
    e1.record()
    torch.cuda.synchronize()
    out["nccl_us_per_allreduce"] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
