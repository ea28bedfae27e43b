"""Greedy decoding throughput (BASELINE config 4: decoder-only, batch 1024, 28 frames, max len 30) on 1 GPU, device time.
   Usage: python tools/bench_greedy.py [batch] [precision]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recnet_b200  # noqa: E402
from recnet_b200 import train as T  # noqa: E402
from recnet_b200.data import synthetic_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
C = T.C
C.decoder_model, C.precision, C.device, C.batch_size = "LSTM", prec, "cuda:0", B
torch.manual_seed(0)
dec = T.build_decoder(4188)["model"].eval()
feats, _, _ = synthetic_batch(B, 28, 1536, 4188, 30, seed=3)
feats = feats.cuda()
for _ in range(3):
    ids, n = dec.greedy(feats, 31)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
e0.record()
for _ in range(K):
    ids, n = dec.greedy(feats, 31)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"greedy decode: batch {B}, {prec}, 31 steps: {ms:.3f} ms per batch, {B / ms * 1e3:.0f} captions/s, steps decoded {int(n)}")
