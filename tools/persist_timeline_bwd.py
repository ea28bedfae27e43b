"""%globaltimer timeline of the weight-resident local-reconstructor BPTT loop (csrc/seq_recon_persist.cuh:local_bwd_kernel), block 0:
1 dG ready seen by the producer, 4 accumulator ready, 5 partials-ready arrive, 12 partials ready seen by the attention warps,
3 dWh-ready arrive, 7 dWh ready seen by the cell phase, 8 cell math + stores done, 9 dG-ready arrive."""
import os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recnet_b200
from recnet_b200 import _lib as L, train as T
from recnet_b200.data import synthetic_batch
lib = L.lib(); dev = torch.device("cuda:0")
C = T.C
C.decoder_model = C.reconstructor_model = "LSTM"; C.reconstructor_type = "local"; C.precision = "bf16"; C.device = "cuda"
dec = T.build_decoder(4188); rec = T.build_reconstructor()
dec["model"].train(); rec["model"].train()
feats, targets, masks = synthetic_batch(100, 28, 1536, 4188, 30, seed=1)
f, t, m = feats.to(dev), targets.to(dev), masks.to(dev)
with torch.no_grad():
    _, hid, _ = T.forward_decoder(dec, f, t, m, 1.0, n_steps=31)
hid = hid.detach().requires_grad_(True)
for _ in range(3):
    T.forward_local_reconstructor(hid, f, rec).backward()
torch.cuda.synchronize()
loss = T.forward_local_reconstructor(hid, f, rec)
torch.cuda.synchronize()
buf = torch.zeros(4096, dtype=torch.int64, device=dev)
L.check(lib.recnet_debug_set_timeline(buf.data_ptr()))
loss.backward()
torch.cuda.synchronize()
L.check(lib.recnet_debug_set_timeline(None))
st = buf.cpu().numpy().astype("uint64")
st = st[st != 0]
tag = (st >> 56).astype(int); ns = (st & ((1 << 56) - 1)).astype("int64")
order = ns.argsort(); tag, ns = tag[order], ns[order]
print("records", len(st), "total us", (ns[-1] - ns[0]) / 1e3)
agg = collections.defaultdict(list)
for i in range(len(st) - 1):
    agg[(int(tag[i]), int(tag[i + 1]))].append((ns[i + 1] - ns[i]) / 1e3)
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    v2 = sorted(v)
    print(f"  {k[0]} -> {k[1]}  n={len(v):3d} mean {sum(v)/len(v):6.2f} us median {v2[len(v2)//2]:6.2f} total {sum(v):8.1f}")
