#!/usr/bin/env python
"""Where the captured train step spends its time: CUDA-graph replays of growing prefixes of the step
(decoder fwd | + reconstructor fwd | + backward | + clip/Adam), differences = per-segment cost under graph replay.
    python tools/segments.py [--recon local|global|none] [--steps 40]
Prints one JSON object (ms per replay of each prefix and the differences)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--recon", default="local")
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--precision", default="bf16")
    args = ap.parse_args()
    import bench
    from recnet_b200 import train as T
    from recnet_b200.data import synthetic_batch
    s = bench.SHAPE
    dev = torch.device("cuda", 0)
    C = T.C
    C.decoder_model = C.reconstructor_model = "LSTM"
    C.batch_size, C.caption_max_len, C.encoder_output_len, C.encoder_output_size = s["B"], s["cap"], s["T"], s["E"]
    C.decoder_n_layers, C.decoder_hidden_size, C.decoder_attn_size, C.embedding_size = 1, s["H"], s["A"], s["EMB"]
    C.reconstructor_n_layers, C.reconstructor_hidden_size, C.reconstructor_attn_size = 1, s["R"], s["A"]
    C.use_recon = args.recon != "none"
    C.reconstructor_type = args.recon if C.use_recon else "local"
    C.precision, C.device = args.precision, "cuda:0"
    torch.manual_seed(0)
    dec = T.build_decoder(s["V"])
    rec = T.build_reconstructor() if C.use_recon else None
    L = s["cap"] + 1
    feats, targets, _ = synthetic_batch(s["B"], s["T"], s["E"], s["V"], s["cap"], seed=1234)
    feats, targets = feats.to(dev), targets.to(dev)
    dec["model"].train()
    if rec:
        rec["model"].train()
    fwd_rec = T.forward_reconstructor_for(C.reconstructor_type) if rec else None

    def seg_dec_fwd():
        T.forward_decoder(dec, feats, targets, None, 1.0, n_steps=L)

    def seg_fwd():
        dl, hid, _ = T.forward_decoder(dec, feats, targets, None, 1.0, n_steps=L)
        if rec:
            fwd_rec(hid, feats, rec)

    def seg_fwd_bwd():
        T.train_step(dec, rec, feats, targets, n_steps=L, optimizer_step=False)

    def seg_full():
        T.train_step(dec, rec, feats, targets, n_steps=L)

    out = {}
    for name, fn in (("dec_fwd", seg_dec_fwd), ("fwd", seg_fwd), ("fwd_bwd", seg_fwd_bwd), ("full", seg_full)):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.set_grad_enabled(name in ("fwd_bwd", "full")):
                for _ in range(2):
                    fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            with torch.set_grad_enabled(name in ("fwd_bwd", "full")):
                fn()
        for _ in range(5):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        out[name] = round(e0.elapsed_time(e1) / args.steps, 4)
        del g
    out["rec_fwd"] = round(out["fwd"] - out["dec_fwd"], 4)
    out["bwd"] = round(out["fwd_bwd"] - out["fwd"], 4)
    out["optim"] = round(out["full"] - out["fwd_bwd"], 4)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
