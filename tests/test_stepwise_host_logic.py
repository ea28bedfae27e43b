"""CPU tests of the host-side composition in models.py: the per-step ``forward`` of the three modules and the step-wise
``forward_sequence`` fall-back (stacked GRU decoder, GRU reconstructor over a stacked decoder, multi-layer reconstructors)
against the reference-generated golden fixtures.

The operator kernels have no CPU implementation, so the four ops (and the norm regulariser) are replaced by plain-torch
TEST DOUBLES here; what is under test is everything around them -- layer stacking, which state feeds the attention query,
the decoder-layer pseudo-time-steps of the local reconstructor (local_reconstructor.py:52), the mean-pool rescale of the global
one (global_reconstructor.py:33-37), loss assembly (train.py:54-70,96-104,125-130).  The GPU suite runs the same fixtures
through the real kernels (tests/test_gpu_parity.py).
"""
import pytest
import torch

import recnet_b200
from recnet_b200 import models as M
from recnet_b200 import train as T
from tests.golden_util import golden_cases, load_golden

TOL = 2e-6          # train.forward_decoder builds the CE weights in fp32 (as the CUDA path wants them); logic errors are O(1)


def _linear(x, W, bias, precision):
    y = x @ W.t()
    return y if bias is None else y + bias


def _additive_attention(Wh, Uv, attn_b, attn_w, V, precision):
    e = torch.tanh(Wh.unsqueeze(1) + Uv + attn_b) @ attn_w.view(-1, 1)          # (B,T,1)
    return (e * V).mean(dim=1)


def _lstm_cell(pre, c_prev, precision):
    i, f, g, o = pre.chunk(4, dim=1)
    c = torch.sigmoid(f) * c_prev + torch.sigmoid(i) * torch.tanh(g)
    return torch.sigmoid(o) * torch.tanh(c), c


def _gru_cell(gi, gh, h_prev, precision):
    ir, iz, in_ = gi.chunk(3, dim=1)
    hr, hz, hn = gh.chunk(3, dim=1)
    r, z = torch.sigmoid(ir + hr), torch.sigmoid(iz + hz)
    n = torch.tanh(in_ + r * hn)
    return (1 - z) * n + z * h_prev


@pytest.fixture
def doubles(monkeypatch):
    monkeypatch.setattr(M.ops, "linear", _linear)
    monkeypatch.setattr(M.ops, "additive_attention", _additive_attention)
    monkeypatch.setattr(M.ops, "lstm_cell", _lstm_cell)
    monkeypatch.setattr(M.ops, "gru_cell", _gru_cell)
    monkeypatch.setattr(M.Fn, "param_norm_sum", lambda params: sum(p.norm() for p in params))
    monkeypatch.setattr(M.Decoder, "uses_fused_sequence", property(lambda self: False))
    monkeypatch.setattr(M._ReconstructorBase, "_fused_ok", lambda self, h: False)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)


def _close(a, b, tol=TOL):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item() if a.numel() else 0.0
    assert err <= tol * (1.0 + b.abs().max().item()), err


def _modules(g, kind):
    m = g["meta"]
    C = T.C
    C.decoder_model, C.reconstructor_model = m["dec_model"], m["rec_model"]
    C.batch_size, C.caption_max_len, C.encoder_output_len, C.encoder_output_size = m["B"], m["cap_len"], m["T"], m["E"]
    dec = recnet_b200.Decoder(m["dec_model"], m["dec_layers"], m["E"], m["EMB"], 1, m["H"], m["A"], m["V"], 0.5, 0.5, 0.5).double()
    dec.load_state_dict(g["dec"])
    dec.eval()
    rec = None
    if kind == "global":
        rec = recnet_b200.GlobalReconstructor(m["rec_model"], m["rec_layers"], m["H"], m["E"], 0.5, 0.5, m["cap_len"]).double()
    elif kind == "local":
        rec = recnet_b200.LocalReconstructor(m["rec_model"], m["rec_layers"], m["H"], m["E"], 0.5, 0.5, m["A"]).double()
    if rec is not None:
        rec.load_state_dict(g[kind])
        rec.eval()
    d = {"model": dec, "lambda_reg": torch.tensor(0.001)}
    r = None if rec is None else {"model": rec, "lambda_reg": torch.tensor(0.01)}
    return d, r


@pytest.mark.parametrize("kind", ["none", "global", "local"])
@pytest.mark.parametrize("name", golden_cases())
def test_stepwise_sequence_path_matches_reference_golden(doubles, name, kind):
    g = load_golden(name)
    dec, rec = _modules(g, kind)
    masks = g["targets"] > 0
    dloss, hiddens, _ = T.forward_decoder(dec, g["feats"], g["targets"], masks, 1.0)
    _close(dloss.item(), g["dec_loss"])
    _close(hiddens.detach(), g["hiddens"])
    loss = dloss
    if kind != "none":
        rloss = T.forward_reconstructor_for(kind)(hiddens, g["feats"], rec)
        _close(rloss.item(), g[f"{kind}_loss"])
        loss = dloss + 1.0 * rloss
    loss.backward()
    for k, ref in g["grads"][kind].items():
        owner, key = k.split(".", 1)
        mod = dec["model"] if owner == "dec" else rec["model"]
        _close(dict(mod.named_parameters())[key].grad, ref)


@pytest.mark.parametrize("name", golden_cases())
def test_per_step_decoder_forward_and_stepwise_greedy(doubles, name):
    g = load_golden(name)
    m = g["meta"]
    dec, _ = _modules(g, "none")
    model = dec["model"]
    B = g["feats"].shape[0]
    tok = torch.ones(1, B, dtype=torch.long)
    hid = M._zero_state(m["dec_model"], m["dec_layers"], B, m["H"], g["feats"].device)
    with torch.no_grad():
        logits, new = model(tok, hid, g["feats"])
    _close(logits, g["step0_logits"])
    top = new[0] if m["dec_model"] == "LSTM" else new
    assert tuple(top.shape) == (m["dec_layers"], B, m["H"])
    if m["dec_layers"] > 1:                      # single-layer greedy is one C call (GPU suite); stacked decoders loop over forward()
        ids, n = model.greedy(g["feats"], m["cap_len"] + 1)
        assert torch.equal(ids[: int(n)], g["greedy_ids"])
        lg, hs = model.teacher_forced_logits(torch.cat((tok, g["targets"][: g["hiddens"].shape[0] - 1])), g["feats"])
        _close(hs, g["hiddens"])
        _close(lg[0], g["step0_logits"])


def test_which_variants_use_the_fused_sequence_drivers():
    D = recnet_b200.Decoder
    mk = lambda name, nl: D(name, nl, 16, 8, 1, 8, 8, 11, 0.5, 0.5, 0.5)
    assert mk("LSTM", 1).uses_fused_sequence and mk("LSTM", 4).uses_fused_sequence and mk("GRU", 1).uses_fused_sequence
    assert not mk("GRU", 2).uses_fused_sequence and not mk("LSTM", 5).uses_fused_sequence
    one, two = torch.zeros(3, 1, 2, 8), torch.zeros(3, 2, 2, 8)
    L = recnet_b200.LocalReconstructor
    assert L("LSTM", 1, 8, 16, 0.5, 0.5, 8)._fused_ok(one) and L("LSTM", 1, 8, 16, 0.5, 0.5, 8)._fused_ok(two)
    assert L("GRU", 1, 8, 16, 0.5, 0.5, 8)._fused_ok(one) and not L("GRU", 1, 8, 16, 0.5, 0.5, 8)._fused_ok(two)
    assert not L("LSTM", 2, 8, 16, 0.5, 0.5, 8)._fused_ok(one)
    G = recnet_b200.GlobalReconstructor
    assert G("LSTM", 1, 8, 16, 0.5, 0.5, 30)._fused_ok(torch.zeros(3, 2, 8)) and not G("GRU", 2, 8, 16, 0.5, 0.5, 30)._fused_ok(one)
