"""Per-replay duration distribution of the captured train-step graph (1 GPU): is there step-to-step jitter that a
data-parallel lock-step would have to pay for?  Usage: python tools/step_jitter.py [n_replays]"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recnet_b200  # noqa: E402
from recnet_b200 import train as T  # noqa: E402
from recnet_b200.data import synthetic_batch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
C = T.C
C.decoder_model = C.reconstructor_model = "LSTM"
C.precision, C.device, C.batch_size = "bf16", "cuda:0", 100
C.reconstructor_type = "local"
torch.manual_seed(0)
dec, rec = T.build_decoder(4188), T.build_reconstructor()
feats, targets, _ = synthetic_batch(100, 28, 1536, 4188, 30, seed=1)
feats, targets = feats.cuda(), targets.cuda()


def step():
    T.train_step(dec, rec, feats, targets, n_steps=31)


s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
for _ in range(5):
    g.replay()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
ev[0].record()
for i in range(n):
    g.replay()
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
q = statistics.quantiles(ms, n=100)
print(f"replays {n}: mean {statistics.mean(ms):.4f} ms, stdev {statistics.pstdev(ms):.4f}, min {min(ms):.4f}, p50 {q[49]:.4f}, p90 {q[89]:.4f}, "
      f"p99 {q[98]:.4f}, max {max(ms):.4f}")
