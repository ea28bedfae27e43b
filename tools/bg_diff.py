import os, sys, torch
sys.path.insert(0, ".")
import recnet_b200
from recnet_b200 import train as T
from recnet_b200.data import synthetic_batch
from tests.test_gpu_parity import FULL, configure, dev
m = dict(FULL, B=100)
feats, targets, _ = synthetic_batch(m["B"], m["T"], m["E"], m["V"], m["cap_len"], seed=77)
feats, targets = feats.to(dev()), targets.to(dev())
L = m["cap_len"] + 1
res = {}
for bg in ("0", "1", "1"):
    os.environ["RECNET_BG_WGRAD"] = bg
    configure(m, "bf16", "local"); T.C.batch_size = 100
    torch.manual_seed(5)
    dec = T.build_decoder(FULL["V"]); rec = T.build_reconstructor()
    T.train_step(dec, rec, feats, targets, n_steps=L, optimizer_step=(os.environ.get("OPT", "1") == "1"))
    torch.cuda.synchronize()
    names = [("dec." + k, p) for k, p in dec["model"].named_parameters()] + [("rec." + k, p) for k, p in rec["model"].named_parameters()]
    cur = {k: (p.detach().clone(), None if p.grad is None else p.grad.detach().clone()) for k, p in names}
    if bg in res:
        tag = "1 vs 1"
        ref = res[bg]
    else:
        res[bg] = cur
        if bg == "0":
            continue
        tag = "0 vs 1"; ref = res["0"]
    for k in cur:
        dp = (cur[k][0] - ref[k][0]).abs().max().item()
        dg = (cur[k][1] - ref[k][1]).abs().max().item() if cur[k][1] is not None else -1
        if dp or dg:
            print(tag, k, "param diff", dp, "grad diff", dg, "grad max", ref[k][1].abs().max().item())
print("done")
