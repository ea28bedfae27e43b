"""GPU tests of the fused clip + Adam kernels (recnet_adam_step via optim.ClipAdam) against
torch.nn.utils.clip_grad_norm_ + torch.optim.Adam -- the two calls of the reference iteration (train.py:269-273)."""
import copy

import pytest
import torch

import recnet_b200
from recnet_b200 import train as T
from recnet_b200.optim import ClipAdam
from tests.golden_util import load_golden
from tests.test_gpu_parity import build, dev, rel

pytestmark = pytest.mark.gpu


def assert_same_update(p, q, p_before, ulps=4, what=""):
    """p (ours) vs q (torch) after the same optimiser steps from p_before.  The weights are O(1) and the updates O(lr), so the
    comparison is made where fp32 can resolve it: |p - q| within a few ulps of the weights, and the update itself to 1 %."""
    p, q, z = p.detach().double(), q.detach().double(), p_before.detach().double()
    ulp = 2.0 ** -23 * float(q.abs().max())
    assert float((p - q).abs().max()) <= ulps * ulp, (what, float((p - q).abs().max()), ulp)
    upd = float((q - z).abs().max())
    assert upd > 0 and float(((p - z) - (q - z)).abs().max()) <= 1e-2 * upd + ulps * ulp, what

SHAPES = [(4188, 468), (128, 512), (128,), (1, 128), (2048, 2004), (2048,), (7, 3), (1,), (16385,), (33, 5, 7)]


def _make(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g).to(dev()).requires_grad_(True) for s in SHAPES]


def _flat_grads(params, seed, scale):
    """Gradients as views of one flat buffer, in an order different from the parameter order (like the sequence Functions)."""
    g = torch.Generator().manual_seed(seed)
    order = list(range(len(params)))[::-1]
    total = sum((params[i].numel() + 63) // 64 * 64 for i in order)
    flat = torch.zeros(total, device=dev())
    off, views = 0, {}
    for i in order:
        n = params[i].numel()
        flat[off: off + n] = (torch.randn(n, generator=g) * scale).to(dev())
        views[i] = flat[off: off + n].view_as(params[i])
        off += (n + 63) // 64 * 64
    return flat, [views[i] for i in range(len(params))]


@pytest.mark.parametrize("amsgrad", [False, True])
@pytest.mark.parametrize("max_norm,gscale", [(None, 1.0), (50.0, 1.0), (50.0, 1e-3)])      # no clip / clip active / clip inactive
@pytest.mark.parametrize("flat", [True, False])
def test_clip_adam_matches_torch_clip_plus_adam(amsgrad, max_norm, gscale, flat):
    ours, ref = _make(1), _make(1)
    kw = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5, amsgrad=amsgrad)
    opt = ClipAdam(ours, max_grad_norm=max_norm, **kw)
    opt_ref = torch.optim.Adam(ref, **kw)
    p0 = [p.detach().clone() for p in ref]
    for it in range(6):
        if flat:
            _, grads = _flat_grads(ours, 100 + it, gscale)
        else:
            g = torch.Generator().manual_seed(100 + it)
            grads = [(torch.randn(p.shape, generator=g) * gscale).to(dev()) for p in ours]
        for p, q, gr in zip(ours, ref, grads):
            p.grad = gr
            q.grad = gr.detach().clone()
        if max_norm is not None:
            total = torch.nn.utils.clip_grad_norm_(ref, max_norm)
        opt.step()
        opt_ref.step()
        if max_norm is not None:
            assert abs(float(opt.last_grad_norm) - float(total)) <= 1e-5 * float(total)
            for p, q in zip(ours, ref):                       # clip_grad_norm_ scales .grad in place; so do we
                assert rel(p.grad, q.grad) < 1e-5
    for p, q, z in zip(ours, ref, p0):
        assert_same_update(p, q, z)
    sd, sd_ref = opt.state_dict(), opt_ref.state_dict()
    for i in sd_ref["state"]:
        assert float(sd["state"][i]["step"]) == float(sd_ref["state"][i]["step"]) == 6.0
        for k in ("exp_avg", "exp_avg_sq") + (("max_exp_avg_sq",) if amsgrad else ()):
            assert rel(sd["state"][i][k], sd_ref["state"][i][k]) < 1e-5


def test_clip_adam_state_dict_round_trips_with_torch_adam():
    ours, ref = _make(2), _make(2)
    kw = dict(lr=1e-3, weight_decay=1e-5, amsgrad=True)
    opt_ref = torch.optim.Adam(ref, **kw)
    for it in range(3):
        g = torch.Generator().manual_seed(it)
        for q in ref:
            q.grad = torch.randn(q.shape, generator=g).to(dev())
        opt_ref.step()
    for p, q in zip(ours, ref):
        p.data.copy_(q.data)
    opt = ClipAdam(ours, **kw)
    opt.load_state_dict(copy.deepcopy(opt_ref.state_dict()))          # reference-style checkpoint -> ours
    g = torch.Generator().manual_seed(99)
    for p, q in zip(ours, ref):
        p.grad = torch.randn(p.shape, generator=g).to(dev())
        q.grad = p.grad.clone()
    before = [q.detach().clone() for q in ref]
    opt.step(); opt_ref.step()
    for p, q, z in zip(ours, ref, before):
        assert_same_update(p, q, z)
    fresh = _make(2)
    opt2 = torch.optim.Adam(fresh, **kw)
    opt2.load_state_dict(copy.deepcopy(opt.state_dict()))              # ours -> torch.optim.Adam
    assert float(opt2.state_dict()["state"][0]["step"]) == 4.0


def test_clip_adam_rejects_cpu_and_missing_grads():
    with pytest.raises(RuntimeError):
        o = ClipAdam([torch.zeros(4, requires_grad=True)], lr=1e-3)
        o.param_groups[0]["params"][0].grad = torch.zeros(4)
        o.step()
    ps = _make(3)
    o = ClipAdam(ps, lr=1e-3)
    ps[0].grad = torch.zeros_like(ps[0])
    with pytest.raises(NotImplementedError):
        o.step()


def test_train_step_with_own_optimizer_matches_torch_optimizer_and_graph_replay_counts_steps(monkeypatch):
    """Same model, same batch, eval-mode dropout off is not available in train_step, so compare on identical dropout seeds."""
    g = load_golden("small_lstm")
    feats, targets = g["feats"].float().to(dev()), g["targets"].to(dev())
    L_steps = g["hiddens"].shape[0]
    results = {}
    for impl in ("torch", "recnet"):
        monkeypatch.setenv("RECNET_OPTIMIZER", impl)
        dec, rec = build(g["meta"], "fp32", "local", g["dec"], g["local"])
        assert isinstance(dec["optimizer"], ClipAdam) == (impl == "recnet")
        dec["model"].seed_dropout(7); rec["model"].seed_dropout(8)
        w0 = [p.detach().clone() for p in list(dec["model"].parameters()) + list(rec["model"].parameters())]
        for _ in range(3):
            T.train_step(dec, rec, feats, targets, n_steps=L_steps)
        torch.cuda.synchronize()
        results[impl] = [p.detach().clone() for p in list(dec["model"].parameters()) + list(rec["model"].parameters())]
    for a, b, z in zip(results["recnet"], results["torch"], w0):
        assert_same_update(a, b, z, ulps=8)
    # whole step incl. the own optimizer under CUDA-graph capture: the device-side step counter advances per replay
    monkeypatch.setenv("RECNET_OPTIMIZER", "recnet")
    dec, rec = build(g["meta"], "bf16", "local", g["dec"], g["local"])
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            T.train_step(dec, rec, feats, targets, n_steps=L_steps)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        T.train_step(dec, rec, feats, targets, n_steps=L_steps)
    w0 = dec["model"].out.weight.detach().clone()
    for _ in range(4):
        graph.replay()
    torch.cuda.synchronize()
    assert float(dec["optimizer"].state_dict()["state"][0]["step"]) == 3 + 4       # capture records, it does not run
    assert float(rec["optimizer"].state_dict()["state"][0]["step"]) == 3 + 4
    assert not torch.equal(w0, dec["model"].out.weight)
    assert all(bool(torch.isfinite(p).all()) for p in dec["model"].parameters())


@pytest.mark.parametrize("name", ["tiny_lstm_ragged", "small_lstm"])
def test_teacher_forcing_prep_matches_the_reference_formulas(name):
    """recnet_teacher_forcing_prep: <SOS> row + shifted targets (train.py:25,44-45) and CE weights mask / (n_t * sum n_t) (train.py:54-60,68)."""
    from recnet_b200 import functional as Fn
    g = load_golden(name)
    targets = g["targets"].to(dev())
    L_steps, B = g["hiddens"].shape[0], targets.shape[1]
    tok, w = Fn.teacher_forcing_inputs(targets, L_steps, 0, 1)
    m = (targets > 0)[:L_steps].float()
    n_t = m.sum(dim=1, keepdim=True)
    ref_w = m / (n_t.clamp_min(1.0) * n_t.sum())
    ref_tok = torch.cat((torch.ones(1, B, dtype=torch.long, device=dev()), targets[: L_steps - 1]), dim=0)
    assert torch.equal(tok, ref_tok)
    assert rel(w, ref_w) < 1e-6 and abs(float(w.sum()) - float(ref_w.sum())) < 1e-5
    with pytest.raises(RuntimeError):
        Fn.teacher_forcing_inputs(g["targets"], L_steps, 0, 1)          # CPU tensor: no CPU path
