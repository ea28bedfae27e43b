"""Pin the CPU oracle (oracle/recnet_oracle.py) to the reference-run golden vectors.

The fixtures were produced by tests/golden/make_golden.py calling the real
reference (train.forward_decoder, forward_{global,local}_reconstructor,
eval.greedy_search, Decoder.forward) in fp64.  fp64-vs-fp64 tolerance: 1e-9.
"""
import pytest
import torch

from oracle import recnet_oracle as O
from tests.golden_util import golden_cases, load_golden

TOL = 1e-9


def _req(P):
    return {k: v.clone().requires_grad_(True) for k, v in P.items()}


def _close(a, b, tol=TOL):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item() if a.numel() else 0.0
    assert err <= tol * (1.0 + b.abs().max().item()), err


@pytest.mark.parametrize("name", golden_cases())
def test_decoder_loss_hiddens_grads(name):
    g = load_golden(name)
    m = g["meta"]
    P = _req(g["dec"])
    masks = g["targets"] > 0
    loss, hiddens, _, aux = O.forward_decoder(P, g["feats"], g["targets"], masks, model_name=m["dec_model"],
                                              n_layers=m["dec_layers"], caption_max_len=m["cap_len"])
    _close(loss.item(), g["dec_loss"])
    _close(hiddens.detach(), g["hiddens"])
    loss.backward()
    for k, ref in g["grads"]["none"].items():
        _close(P[k[len("dec."):]].grad, ref)


@pytest.mark.parametrize("kind", ["global", "local"])
@pytest.mark.parametrize("name", golden_cases())
def test_reconstructor_loss_and_joint_grads(name, kind):
    g = load_golden(name)
    m = g["meta"]
    P, Q = _req(g["dec"]), _req(g[kind])
    masks = g["targets"] > 0
    dloss, hiddens, _, _ = O.forward_decoder(P, g["feats"], g["targets"], masks, model_name=m["dec_model"],
                                             n_layers=m["dec_layers"], caption_max_len=m["cap_len"])
    if kind == "global":
        rloss, _ = O.forward_global_reconstructor(Q, hiddens, g["feats"], model_name=m["rec_model"],
                                                  n_layers=m["rec_layers"], caption_max_len=m["cap_len"])
    else:
        rloss, _ = O.forward_local_reconstructor(Q, hiddens, g["feats"], model_name=m["rec_model"],
                                                 n_layers=m["rec_layers"])
    _close(rloss.item(), g[f"{kind}_loss"])
    (dloss + 1.0 * rloss).backward()
    for k, ref in g["grads"][kind].items():
        owner, key = k.split(".", 1)
        _close((P if owner == "dec" else Q)[key].grad, ref)


@pytest.mark.parametrize("name", golden_cases())
def test_greedy_ids_bit_exact_and_step_logits(name):
    g = load_golden(name)
    m = g["meta"]
    ids = O.greedy_search(g["dec"], g["feats"], model_name=m["dec_model"], n_layers=m["dec_layers"],
                          caption_max_len=m["cap_len"])
    assert torch.equal(ids, g["greedy_ids"])
    B, H = g["feats"].shape[0], m["H"]
    tok = torch.full((1, B), O.SOS, dtype=torch.long)
    hid = O.zero_hidden(m["dec_model"], m["dec_layers"], B, H, g["feats"])
    logits, _ = O.decoder_step(g["dec"], tok, hid, g["feats"], model_name=m["dec_model"], n_layers=m["dec_layers"])
    _close(logits, g["step0_logits"])


def test_synthetic_batch_shape_contract():
    feats, targets, masks = O.synthetic_batch(6, 28, 32, 50, caption_max_len=30, seed=7)
    assert feats.shape == (6, 28, 32) and targets.shape == (31, 6)
    assert int(masks[:, 0].sum()) == 31                     # sample 0 is full length -> L = 31
    assert bool(((targets == O.EOS).sum(0) == 1).all())     # exactly one <EOS> per caption
    assert not bool((targets == O.SOS).any())               # no <SOS> inside targets
