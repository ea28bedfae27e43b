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


# ---------------------------------------------------------------------------------------------------------------
# train mode (dropout on) and beam search: fixtures of tests/golden/make_golden_train.py
# ---------------------------------------------------------------------------------------------------------------
from tests.golden_util import load_golden_beam, load_golden_train, philox_scales      # noqa: E402
from tests.philox_ref import SITE_EMB, SITE_GLOBAL_MP, SITE_LOCAL_X, SITE_LOGITS      # noqa: E402


@pytest.mark.parametrize("name", ["small_lstm", "tiny_lstm_ragged"])
@pytest.mark.parametrize("kind", ["none", "global", "local"])
def test_oracle_train_mode_with_philox_masks_matches_reference(name, kind):
    """Dropout on: the same Philox masks (tests/philox_ref.py) injected into the real reference's nn.Dropout modules and passed to the
    oracle give the same losses and the same gradient for every parameter of dec_loss + 1.0 * rec_loss."""
    g = load_golden_train(name)
    m = g["meta"]
    feats, targets = g["feats"], g["targets"]
    Lmax, B = m["cap_len"] + 1, m["B"]
    P = {k: v.clone().requires_grad_(True) for k, v in g["dec"].items()}
    de = philox_scales(g["seed_dec"], 1, SITE_EMB, (Lmax, B, m["EMB"]), g["p_emb"])
    dl_ = philox_scales(g["seed_dec"], 1, SITE_LOGITS, (Lmax, B, m["V"]), g["p_out"])
    dloss, hid, _, _ = O.forward_decoder(P, feats, targets, targets > 0, caption_max_len=m["cap_len"], drop_emb=de, drop_logits=dl_)
    assert abs(float(dloss) - g["dec_loss"]) < 1e-9 * max(1.0, abs(g["dec_loss"]))
    assert torch.allclose(hid, g["hiddens"], rtol=1e-9, atol=1e-11)
    loss = dloss
    Q = None
    if kind != "none":
        Q = {k: v.clone().requires_grad_(True) for k, v in g[kind].items()}
        L = hid.shape[0]
        if kind == "local":
            sc = philox_scales(g["seed_local"], 1, SITE_LOCAL_X, (m["T"], B, m["H"]), g["p_rec"])
            rloss, _ = O.forward_local_reconstructor(Q, hid, feats, drop_x=sc)
        else:
            sc = philox_scales(g["seed_global"], 1, SITE_GLOBAL_MP, (L, B, m["H"]), g["p_rec"])
            rloss, _ = O.forward_global_reconstructor(Q, hid, feats, caption_max_len=m["cap_len"], drop_mp=sc)
        assert abs(float(rloss) - g[f"{kind}_loss"]) < 1e-9 * max(1.0, abs(g[f"{kind}_loss"]))
        loss = loss + rloss
    loss.backward()
    for k, ref in g["grads"][kind].items():
        mod, key = k.split(".", 1)
        got = (P if mod == "dec" else Q)[key].grad
        assert torch.allclose(got, ref, rtol=1e-8, atol=1e-11), k


@pytest.mark.parametrize("name", ["tiny_lstm", "tiny_gru", "small_lstm"])
def test_oracle_beam_search_matches_the_references_own_beam_search(name):
    """eval.beam_search run by the reference itself (CPU alias for its torch.cuda.FloatTensor constructor) vs the oracle restatement."""
    g = load_golden_beam(name)
    m = g["meta"]
    for width in (3, 5):
        got = O.beam_search(g["dec"], g["feats"], width, model_name=m["dec_model"], n_layers=1, caption_max_len=m["cap_len"])
        assert got == g["beams"][width], (width, got, g["beams"][width])
