"""%globaltimer timeline of the cluster-resident decoder loops (csrc/seq_decoder_cluster.cuh), block 0 / thread 0.
forward tags: 20 step start, 21 GEMM done, 27 W.h all-gather issued, 22 barrier 1 passed, 23 scores + all-gather issued,
24 barrier 2 passed, 25 context + cell + h all-gather issued, 26 barrier 3 passed."""
import os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recnet_b200
from recnet_b200 import _lib as L, train as T
from recnet_b200.data import synthetic_batch
lib = L.lib(); dev = torch.device("cuda:0")
C = T.C
C.decoder_model = C.reconstructor_model = "LSTM"; C.reconstructor_type = "local"; C.precision = "bf16"; C.device = "cuda"
dec = T.build_decoder(4188)
dec["model"].train()
feats, targets, masks = synthetic_batch(100, 28, 1536, 4188, 30, seed=1)
f, t, m = feats.to(dev), targets.to(dev), masks.to(dev)
bwd = len(sys.argv) > 1 and sys.argv[1] == "bwd"
for _ in range(3):
    loss, hid, _ = T.forward_decoder(dec, f, t, m, 1.0, n_steps=31)
    if bwd:
        (loss + hid.sum() * 1e-3).backward()
torch.cuda.synchronize()
buf = torch.zeros(4096, dtype=torch.int64, device=dev)
if not bwd:
    L.check(lib.recnet_debug_set_timeline(buf.data_ptr()))
loss, hid, _ = T.forward_decoder(dec, f, t, m, 1.0, n_steps=31)
torch.cuda.synchronize()
if bwd:
    L.check(lib.recnet_debug_set_timeline(buf.data_ptr()))
    (loss + hid.sum() * 1e-3).backward()
    torch.cuda.synchronize()
L.check(lib.recnet_debug_set_timeline(None))
st = buf.cpu().numpy().astype("uint64")
st = st[st != 0]
tag = (st >> 56).astype(int); ns = (st & ((1 << 56) - 1)).astype("int64")
order = ns.argsort(kind="stable"); tag, ns = tag[order], ns[order]
print("records", len(st), "total us", (ns[-1] - ns[0]) / 1e3)
agg = collections.defaultdict(list)
for i in range(len(st) - 1):
    agg[(int(tag[i]), int(tag[i + 1]))].append((ns[i + 1] - ns[i]) / 1e3)
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    v2 = sorted(v)
    print(f"  {k[0]} -> {k[1]}  n={len(v):3d} mean {sum(v)/len(v):6.2f} us median {v2[len(v2)//2]:6.2f} total {sum(v):8.1f}")
