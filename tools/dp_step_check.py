"""Whole data-parallel train steps under torchrun (>= 2 GPUs of one box): the default flow (own NVLink all-reduce kernel in
symmetric memory, reconstructor slice reduced underneath the decoder's backward, its optimiser step underneath the decoder's slice)
against NCCL after backward (RECNET_DP_IMPL=nccl) -- eagerly and as a CUDA graph.
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_step_check.py
Prints one JSON line: worst relative parameter difference after 3 steps, and whether every rank holds identical parameters."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(mode, graph, s, dev, rank):
    import bench
    from recnet_b200 import train as T
    from recnet_b200.data import synthetic_batch
    from recnet_b200.parallel import broadcast_parameters, make_reducer
    os.environ["RECNET_DP_IMPL"] = "nvlink" if mode == "nvlink" else "nccl"
    bench._configure(T.C, s, "local", "bf16", 1, f"cuda:{dev.index}")
    torch.manual_seed(0)
    dec, rec = T.build_decoder(s["V"]), T.build_reconstructor()
    broadcast_parameters([dec["model"], rec["model"]])
    dec["model"].seed_dropout(5); rec["model"].seed_dropout(6)
    feats, targets, _ = synthetic_batch(s["B"], s["T"], s["E"], s["V"], s["cap"], seed=1234 + rank)
    feats, targets = feats.to(dev), targets.to(dev)
    red = make_reducer([rec["model"], dec["model"]])

    def step():
        red.start_iteration()
        T.train_step(dec, rec, feats, targets, n_steps=s["cap"] + 1, reducer=red)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
        if not graph:
            step(); step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    if graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            step()
        g.replay(); g.replay()
        torch.cuda.synchronize()
    if hasattr(red, "check"):
        red.check()
    params = [p.detach().clone() for p in list(dec["model"].parameters()) + list(rec["model"].parameters())]
    same = True
    for p in params:
        mx, mn = p.clone(), p.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        same = same and bool(torch.equal(mx, mn))
    name = type(red).__name__
    if hasattr(red, "remove"):
        red.remove()
    del red
    return params, same, name


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    import bench
    import recnet_b200  # noqa: F401
    s = bench.SHAPE
    out = {"world": world}
    for graph in (False, True):
        ref, same_ref, n_ref = run("plain", graph, s, dev, rank)
        got, same_got, n_got = run("nvlink", graph, s, dev, rank)
        worst = 0.0
        for a, b in zip(got, ref):
            worst = max(worst, float((a - b).abs().max() / (b.abs().max() + 1e-30)))
        k = "graph" if graph else "eager"
        out[k] = {"worst_rel_param_diff_after_3_steps": worst, "ranks_identical_plain": same_ref, "ranks_identical_nvlink": same_got,
                  "reducers": [n_ref, n_got]}
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)       # NCCL captured in a live graph: destroy_process_group() would hang (see bench.py)


if __name__ == "__main__":
    main()
