"""The background lane of train.train_step (functional.deferred_weight_grads): the reconstructor's weight gradients, its optimiser step
and the decoder's vocabulary-projection gradients run on a second stream underneath the decoder's backward loop.  Same arithmetic in
the same order per tensor, so training steps must leave the same parameters and optimiser state, eagerly and as a CUDA graph."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

import recnet_b200  # noqa: E402,F401
from recnet_b200 import train as T  # noqa: E402
from recnet_b200.data import synthetic_batch  # noqa: E402
from tests.test_gpu_parity import FULL, configure, dev  # noqa: E402


def _models(kind, seed=5):
    configure(dict(FULL, B=100), "bf16", kind)
    T.C.batch_size = 100
    torch.manual_seed(seed)
    dec = T.build_decoder(FULL["V"])
    rec = T.build_reconstructor() if kind != "none" else None
    return dec, rec


def _state(dec, rec):
    out = [p.detach().clone() for p in dec["model"].parameters()]
    if rec is not None:
        out += [p.detach().clone() for p in rec["model"].parameters()]
        out += [v.detach().clone() for st in rec["optimizer"].state_dict()["state"].values() for v in st.values() if torch.is_tensor(v)]
    out += [v.detach().clone() for st in dec["optimizer"].state_dict()["state"].values() for v in st.values() if torch.is_tensor(v)]
    return out


@pytest.mark.parametrize("kind", ["local", "global", "none"])
@pytest.mark.parametrize("graph", [False, True])
def test_background_lane_leaves_the_same_parameters_as_the_plain_order(kind, graph, monkeypatch):
    m = dict(FULL, B=100)
    feats, targets, _ = synthetic_batch(m["B"], m["T"], m["E"], m["V"], m["cap_len"], seed=77)
    feats, targets = feats.to(dev()), targets.to(dev())
    L = m["cap_len"] + 1
    results = []
    for bg in ("0", "1"):
        monkeypatch.setenv("RECNET_BG_WGRAD", bg)
        dec, rec = _models(kind)
        dec0 = copy.deepcopy(dec["model"].state_dict())

        def step():
            return T.train_step(dec, rec, feats, targets, n_steps=L)

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()                                       # eager step 1 (builds the optimiser tables)
            if graph:
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step()
            g.replay()
            g.replay()
            torch.cuda.synchronize()
        results.append(_state(dec, rec))
        assert any(not torch.equal(a, b) for a, b in zip(dec0.values(), dec["model"].state_dict().values()))   # it did train
    assert len(results[0]) == len(results[1])
    for a, b in zip(*results):
        # bit-identical except what the embedding gradient's float atomics (misc::embed_scatter_kernel) leave open: ~1e-7 relative
        assert (a.double() - b.double()).abs().max().item() <= 1e-6 * a.double().abs().max().item() + 1e-12
