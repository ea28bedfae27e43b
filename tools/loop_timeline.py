"""Tagged %globaltimer timeline of the decoder-forward persistent loop kernel (RECNET_MEGA=1), block 0 / thread 0."""
import os, sys, collections
os.environ["RECNET_MEGA"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recnet_b200
from recnet_b200 import _lib as L, train as T
from recnet_b200.data import synthetic_batch
lib = L.lib(); dev = torch.device("cuda:0")
C = T.C
C.decoder_model = C.reconstructor_model = "LSTM"; C.reconstructor_type = "local"; C.precision = "bf16"; C.device = "cuda"
dec = T.build_decoder(4188)
feats, targets, masks = synthetic_batch(100, 28, 1536, 4188, 30, seed=1)
f, t, m = feats.to(dev), targets.to(dev), masks.to(dev)
names = {0: "end", 1: "gemm", 2: "attn_fwd", 3: "cell_fwd", 4: "cell_bwd", 5: "attn_bwd", 100: "barrier"}
buf = torch.zeros(4096, dtype=torch.int64, device=dev)
for _ in range(3):
    T.forward_decoder(dec, f, t, m, 1.0, n_steps=31)
torch.cuda.synchronize()
L.check(lib.recnet_debug_set_timeline(buf.data_ptr()))
T.forward_decoder(dec, f, t, m, 1.0, n_steps=31)
torch.cuda.synchronize()
L.check(lib.recnet_debug_set_timeline(None))
st = buf.cpu().numpy().astype("uint64")
st = st[st != 0]
tag = (st >> 56).astype(int); ns = (st & ((1 << 56) - 1)).astype("int64")
print("records", len(st), "total us", (ns[-1] - ns[0]) / 1e3)
agg = collections.defaultdict(list)
for i in range(len(st) - 1):
    agg[(int(tag[i]), int(tag[i + 1]))].append((ns[i + 1] - ns[i]) / 1e3)
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    v2 = sorted(v)
    print(f"  {names.get(k[0], k[0])!s:>9} -> {names.get(k[1], k[1])!s:<9} n={len(v):3d} mean {sum(v)/len(v):6.2f} us median {v2[len(v2)//2]:6.2f} total {sum(v):8.1f}")
