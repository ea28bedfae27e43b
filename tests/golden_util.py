"""Load the reference-generated golden fixtures (tests/golden/*.npz)."""
import ast
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name, dtype=torch.float64):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = ast.literal_eval(str(z["meta"]))

    def group(prefix):
        out = {}
        for k in z.files:
            if k.startswith(prefix):
                t = torch.from_numpy(z[k])
                out[k[len(prefix):]] = t.to(dtype) if t.is_floating_point() else t
        return out

    g = {
        "meta": meta,
        "feats": torch.from_numpy(z["feats"]).to(dtype),
        "targets": torch.from_numpy(z["targets"]),
        "dec": group("dec."), "global": group("global."), "local": group("local."),
        "dec_loss": float(z["dec_loss"]), "global_loss": float(z["global_loss"]), "local_loss": float(z["local_loss"]),
        "hiddens": torch.from_numpy(z["hiddens"]).to(dtype),
        "greedy_ids": torch.from_numpy(z["greedy_ids"]),
        "step0_logits": torch.from_numpy(z["step0_logits"]).to(dtype),
        "grads": {kind: group(f"grad_{kind}.") for kind in ("none", "global", "local")},
    }
    return g
