"""CPU oracle for the RecNet hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file is a plain-torch (CPU, fp32 or fp64) restatement of the algorithm the
reference implements with nn.Module objects.  It exists only to check the CUDA
path: it may be imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` and by nothing
else.  The product package never imports it and has no CPU path.

How it is pinned ("parity pinned by reference-run fixtures"): the reference has
no tests or golden vectors of its own, so
``tests/golden/make_golden.py`` imports the real reference modules from
/root/reference in the build container, runs ``train.forward_decoder`` /
``forward_*_reconstructor`` / ``eval.greedy_search`` on seeded inputs and
commits inputs + outputs + gradients as ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks every function below against them.

All arithmetic the reference delegates to PyTorch (nn.LSTM / nn.GRU /
nn.Linear / CrossEntropyLoss / MSELoss, pinned torch-nightly 1.0.0.dev20181113
in the reference's requirements.txt:4) is restated here from the published
formulas: LSTM gate order i,f,g,o, GRU gate order r,z,n.

Weights are passed as a dict keyed exactly like the reference ``state_dict``
(SURVEY.md section 8b).  Gradients come from torch autograd over these explicit
ops (no nn.LSTM, no cuDNN/oneDNN fused RNN on this path).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]

PAD, SOS, EOS = 0, 1, 2  # config.py:56  init_word2idx


# ----------------------------------------------------------------------------
# cells (PyTorch nn.LSTM / nn.GRU published formulas; call sites
# models/decoder.py:36-40, models/global_reconstructor.py:22-26,
# models/local_reconstructor.py:29-33)
# ----------------------------------------------------------------------------
def lstm_cell(x: Tensor, h: Tensor, c: Tensor, w_ih: Tensor, w_hh: Tensor,
              b_ih: Tensor, b_hh: Tensor) -> Tuple[Tensor, Tensor]:
    """One LSTM step. x (B,I), h,c (B,H). Row blocks of w_* are i,f,g,o."""
    gates = x @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
    H = h.shape[1]
    i = torch.sigmoid(gates[:, 0 * H:1 * H])
    f = torch.sigmoid(gates[:, 1 * H:2 * H])
    g = torch.tanh(gates[:, 2 * H:3 * H])
    o = torch.sigmoid(gates[:, 3 * H:4 * H])
    c_new = f * c + i * g
    h_new = o * torch.tanh(c_new)
    return h_new, c_new


def gru_cell(x: Tensor, h: Tensor, w_ih: Tensor, w_hh: Tensor,
             b_ih: Tensor, b_hh: Tensor) -> Tensor:
    """One GRU step. Row blocks of w_* are r,z,n (PyTorch convention)."""
    H = h.shape[1]
    gi = x @ w_ih.t() + b_ih
    gh = h @ w_hh.t() + b_hh
    r = torch.sigmoid(gi[:, 0:H] + gh[:, 0:H])
    z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = torch.tanh(gi[:, 2 * H:3 * H] + r * gh[:, 2 * H:3 * H])
    return (1.0 - z) * n + z * h


def _rnn_layers(P: Params, prefix: str, model_name: str, n_layers: int, xs: Sequence[Tensor], hidden):
    """Run a (seq_len = len(xs)) x n_layers stacked RNN exactly as nn.LSTM/nn.GRU
    would for an input of shape (len(xs), B, I) in eval mode (inter-layer dropout
    off).  Returns (outputs of the top layer per pseudo-timestep, new hidden)."""
    is_lstm = model_name == "LSTM"
    if is_lstm:
        h_all = [hidden[0][l] for l in range(n_layers)]
        c_all = [hidden[1][l] for l in range(n_layers)]
    else:
        h_all = [hidden[l] for l in range(n_layers)]
    layer_in = list(xs)
    for l in range(n_layers):
        w_ih, w_hh = P[f"{prefix}.weight_ih_l{l}"], P[f"{prefix}.weight_hh_l{l}"]
        b_ih, b_hh = P[f"{prefix}.bias_ih_l{l}"], P[f"{prefix}.bias_hh_l{l}"]
        outs = []
        h = h_all[l]
        c = c_all[l] if is_lstm else None
        for x in layer_in:
            if is_lstm:
                h, c = lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh)
            else:
                h = gru_cell(x, h, w_ih, w_hh, b_ih, b_hh)
            outs.append(h)
        h_all[l] = h
        if is_lstm:
            c_all[l] = c
        layer_in = outs
    if is_lstm:
        new_hidden = (torch.stack(h_all), torch.stack(c_all))
    else:
        new_hidden = torch.stack(h_all)
    return layer_in, new_hidden


# ----------------------------------------------------------------------------
# additive "attention" shared by decoder and local reconstructor.
# NOTE: the reference builds nn.Softmax but never calls
# it; the raw tanh score multiplies the values and the context is a MEAN.
# ----------------------------------------------------------------------------
def additive_scores(query: Tensor, keys_proj: Tensor, W: Tensor, b: Tensor, w: Tensor) -> Tensor:
    """query (B,Hq); keys_proj = U(keys) with a leading set of key dims whose
    last two dims broadcast with (B,A).  Returns scores with trailing dim 1.
    models/decoder.py:50-58, models/local_reconstructor.py:38-46."""
    Wh = query @ W.t()                                   # (B,A)
    s = torch.tanh(Wh + keys_proj + b)                   # broadcast over key dims
    return s @ w.t()                                     # (...,1)


# ----------------------------------------------------------------------------
# Decoder  (models/decoder.py:45-70)
# ----------------------------------------------------------------------------
def decoder_step(P: Params, tok: Tensor, hidden, feats: Tensor, *, model_name: str = "LSTM",
                 n_layers: int = 1, embedding_scale: float = 1.0, drop_emb: Optional[Tensor] = None,
                 drop_logits: Optional[Tensor] = None):
    """tok (1,B) int64; hidden ((NL,B,H),(NL,B,H)) or (NL,B,H); feats (B,T,E).
    Eval mode by default (all dropouts identity); train mode = the caller supplies this step's inverted-dropout scale tensors
    (0 or 1/(1-p)): drop_emb (B,EMB) for decoder.py:48, drop_logits (B,V) for decoder.py:69.  Returns (logits (B,V), new hidden)."""
    emb = P["embedding.weight"][tok[0]] * embedding_scale             # decoder.py:46-47
    if drop_emb is not None:
        emb = emb * drop_emb                                          # decoder.py:48
    top_h = hidden[0][-1] if model_name == "LSTM" else hidden[-1]    # decoder.py:50-53
    Uv = feats @ P["attn_U.weight"].t()                               # decoder.py:54 (B,T,A), recomputed per step
    Wh = (top_h @ P["attn_W.weight"].t()).unsqueeze(1)                # decoder.py:51,55
    e = torch.tanh(Wh + Uv + P["attn_b"]) @ P["attn_w.weight"].t()    # decoder.py:56-58 (B,T,1)
    ctx = (e * feats).mean(dim=1)                                     # decoder.py:59-61 -- mean, no softmax
    x = torch.cat((emb, ctx), dim=1)                                  # decoder.py:64
    outs, hidden = _rnn_layers(P, "rnn", model_name, n_layers, [x], hidden)  # decoder.py:66
    logits = outs[0] @ P["out.weight"].t() + P["out.bias"]            # decoder.py:68
    if drop_logits is not None:
        logits = logits * drop_logits                                 # decoder.py:69: the LOGITS are dropped, before the loss
    return logits, hidden


def zero_hidden(model_name: str, n_layers: int, B: int, H: int, like: Tensor):
    z = lambda: torch.zeros(n_layers, B, H, dtype=like.dtype, device=like.device)
    return (z(), z()) if model_name == "LSTM" else z()


def param_norm_sum(P: Params) -> Tensor:
    """sum of UN-squared L2 norms over every parameter tensor (train.py:69,101,127)."""
    return sum(torch.sqrt((p * p).sum()) for p in P.values())


def forward_decoder(P: Params, feats: Tensor, targets: Tensor, masks: Tensor, *,
                    model_name: str = "LSTM", n_layers: int = 1, embedding_scale: float = 1.0,
                    caption_max_len: int = 30, lambda_reg: float = 1e-3,
                    teacher_forcing: bool = True, drop_emb: Optional[Tensor] = None, drop_logits: Optional[Tensor] = None):
    """train.py:17-75.  targets (caption_max_len+1, B) int64 ; masks = targets > 0.
    Train mode: drop_emb (L,B,EMB) / drop_logits (L,B,V) hold the per-step inverted-dropout scales (None = eval mode).
    Returns (loss, hiddens (L,NL,B,H), output_indices list, aux dict)."""
    B = feats.shape[0]
    H = P["rnn.weight_hh_l0"].shape[1]
    tok = torch.full((1, B), SOS, dtype=torch.long)                    # train.py:25
    hidden = zero_hidden(model_name, n_layers, B, H, feats)            # train.py:28-35
    ce_sum = feats.new_zeros(())
    n_sum = 0
    hiddens: List[Tensor] = []
    out_idx: List[Tensor] = []
    logits_all: List[Tensor] = []
    for t in range(caption_max_len + 1):                               # train.py:41
        logits, hidden = decoder_step(P, tok, hidden, feats, model_name=model_name,
                                      n_layers=n_layers, embedding_scale=embedding_scale,
                                      drop_emb=None if drop_emb is None else drop_emb[t],
                                      drop_logits=None if drop_logits is None else drop_logits[t])
        logits_all.append(logits)
        if teacher_forcing:
            tok = targets[t].view(1, -1)                               # train.py:44-45
        else:
            tok = logits.argmax(dim=1).view(1, -1)                     # train.py:47-51 (topk(1) == argmax, lowest index on ties)
            out_idx.append(tok[0].clone())
        m = masks[t]
        n_t = int(m.sum())
        if n_t > 0:
            lp = torch.log_softmax(logits[m], dim=1)                   # CrossEntropyLoss() mean over the masked rows, train.py:54-57
            ce_t = -(lp.gather(1, targets[t][m].view(-1, 1))).mean()
        else:
            ce_t = feats.new_tensor(float("nan"))                      # reference would produce nan too (mean over 0 rows)
        ce_sum = ce_sum + ce_t                                         # train.py:59
        n_sum += n_t                                                   # train.py:60
        hiddens.append(hidden[0] if model_name == "LSTM" else hidden)  # train.py:61-64
        if t == caption_max_len or not bool(masks[t + 1].any()):       # train.py:66
            break
    ce = ce_sum / n_sum                                                # train.py:68
    reg = param_norm_sum(P)
    loss = ce + lambda_reg * reg                                       # train.py:69-70
    hiddens_t = torch.stack(hiddens)                                   # train.py:73  (L,NL,B,H)
    return loss, hiddens_t, out_idx, {"ce": ce, "reg": reg, "logits": torch.stack(logits_all)}


# ----------------------------------------------------------------------------
# Global reconstructor (models/global_reconstructor.py:30-46, train.py:78-105)
# ----------------------------------------------------------------------------
def global_reconstructor_step(P: Params, inp: Tensor, hidden, decoder_hiddens: Tensor, *,
                              model_name: str = "LSTM", n_layers: int = 1, caption_max_len: int = 30,
                              drop_mp: Optional[Tensor] = None):
    """inp = decoder_hiddens[t] (NLdec,B,H) ; decoder_hiddens (L,NLdec,B,H); drop_mp (B,H): this step's dropout scales (train mode)."""
    L = decoder_hiddens.shape[0]
    mp = decoder_hiddens.mean(dim=0).mean(dim=0)                        # global_reconstructor.py:33-36 (mean over L then layers)
    mp = mp / L * caption_max_len                                       # global_reconstructor.py:37
    if drop_mp is not None:
        mp = mp * drop_mp                                               # global_reconstructor.py:38
    x = torch.cat((inp[0], mp), dim=1)                                  # global_reconstructor.py:40 (layer-0 state only)
    outs, hidden = _rnn_layers(P, "rnn", model_name, n_layers, [x], hidden)   # :43
    out = outs[0] @ P["out.weight"].t() + P["out.bias"]                 # :45
    return out, hidden


def forward_global_reconstructor(P: Params, decoder_hiddens: Tensor, feats: Tensor, *,
                                 model_name: str = "LSTM", n_layers: int = 1, caption_max_len: int = 30,
                                 lambda_reg: float = 1e-2, drop_mp: Optional[Tensor] = None):
    """drop_mp (L,B,H): per-step dropout scales of the mean-pooled state (train mode), None = eval mode."""
    B = feats.shape[0]
    R = P["rnn.weight_hh_l0"].shape[1]
    hidden = zero_hidden(model_name, n_layers, B, R, feats)             # train.py:82-89
    outs = []
    L = decoder_hiddens.shape[0]
    for t in range(L):                                                  # train.py:93
        o, hidden = global_reconstructor_step(P, decoder_hiddens[t], hidden, decoder_hiddens,
                                              model_name=model_name, n_layers=n_layers,
                                              caption_max_len=caption_max_len,
                                              drop_mp=None if drop_mp is None else drop_mp[t])
        outs.append(o)
    outs_t = torch.stack(outs)
    mse = ((outs_t.mean(0) - feats.mean(1)) ** 2).mean()                # train.py:96-99 MSELoss()
    rec = mse / L                                                       # train.py:100
    reg = param_norm_sum(P)
    return rec + lambda_reg * reg, {"rec": rec, "reg": reg, "outputs": outs_t}   # train.py:101-103


# ----------------------------------------------------------------------------
# Local reconstructor (models/local_reconstructor.py:37-55, train.py:108-131)
# ----------------------------------------------------------------------------
def local_reconstructor_step(P: Params, hidden, decoder_hiddens: Tensor, *,
                             model_name: str = "LSTM", n_layers: int = 1, drop_x: Optional[Tensor] = None):
    """decoder_hiddens (L,NLdec,B,H).  The attended input keeps the decoder-layer
    axis, and the RNN treats that axis as TIME (NLdec pseudo-steps) -- SURVEY 8a A7."""
    top_h = hidden[0][-1] if model_name == "LSTM" else hidden[-1]       # local_reconstructor.py:38-41
    Uv = decoder_hiddens @ P["attn_U.weight"].t()                        # :42 (L,NLdec,B,A)
    Wh = top_h @ P["attn_W.weight"].t()                                  # (B,A) broadcast over (L,NLdec)
    beta = torch.tanh(Wh + Uv + P["attn_b"]) @ P["attn_w.weight"].t()    # :44-46 (L,NLdec,B,1)
    x = (beta * decoder_hiddens).mean(dim=0)                             # :47-49 (NLdec,B,H) -- mean over L, no softmax
    if drop_x is not None:
        x = x * drop_x                                                   # :50 (train mode: this step's scales, (NLdec,B,H) or (B,H))
    outs, hidden = _rnn_layers(P, "rnn", model_name, n_layers, list(x.unbind(0)), hidden)  # :52
    out = outs[0] @ P["out.weight"].t() + P["out.bias"]                  # :54 first pseudo-step of the top layer
    return out, hidden


def forward_local_reconstructor(P: Params, decoder_hiddens: Tensor, feats: Tensor, *,
                                model_name: str = "LSTM", n_layers: int = 1, lambda_reg: float = 1e-2,
                                drop_x: Optional[Tensor] = None):
    """drop_x (T,B,H): per-step dropout scales of the attended input (train mode), None = eval mode."""
    B, T, _ = feats.shape
    R = P["rnn.weight_hh_l0"].shape[1]
    hidden = zero_hidden(model_name, n_layers, B, R, feats)             # train.py:112-119
    outs = []
    for t in range(T):                                                  # train.py:122 (encoder_output_len steps)
        o, hidden = local_reconstructor_step(P, hidden, decoder_hiddens, model_name=model_name, n_layers=n_layers,
                                             drop_x=None if drop_x is None else drop_x[t])
        outs.append(o)
    outs_t = torch.stack(outs)                                          # (T,B,R)
    rec = ((outs_t.transpose(0, 1) - feats) ** 2).mean()                # train.py:126-128
    reg = param_norm_sum(P)
    return rec + lambda_reg * reg, {"rec": rec, "reg": reg, "outputs": outs_t}   # train.py:129-130


# ----------------------------------------------------------------------------
# Greedy search (eval.py:19-33)
# ----------------------------------------------------------------------------
@torch.no_grad()
def greedy_search(P: Params, feats: Tensor, *, model_name: str = "LSTM", n_layers: int = 1,
                  embedding_scale: float = 1.0, caption_max_len: int = 30) -> Tensor:
    """Returns (n_steps, B) int64 token ids; stops when every fed-back token is <PAD>."""
    B = feats.shape[0]
    H = P["rnn.weight_hh_l0"].shape[1]
    tok = torch.full((1, B), SOS, dtype=torch.long)
    hidden = zero_hidden(model_name, n_layers, B, H, feats)
    ids = []
    for t in range(caption_max_len + 1):
        logits, hidden = decoder_step(P, tok, hidden, feats, model_name=model_name, n_layers=n_layers,
                                      embedding_scale=embedding_scale)
        tok = logits.argmax(dim=1).view(1, -1)                          # eval.py:24 topk(1): lowest index wins ties
        ids.append(tok[0].clone())
        if t == caption_max_len or bool((tok == PAD).all()):            # eval.py:30
            break
    return torch.stack(ids)


@torch.no_grad()
def beam_search(P: Params, feats: Tensor, beam_width: int, *, model_name: str = "LSTM", n_layers: int = 1,
                embedding_scale: float = 1.0, caption_max_len: int = 30) -> List[List[int]]:
    """eval.py:36-120 restated on CPU (the reference body is CUDA-only because it
    builds torch.cuda.FloatTensor at eval.py:39,57).  Scores are log(sigmoid(logit))
    (eval.py:61), the running score is divided by len**0.7 EVERY step before the
    new term is added (eval.py:53-59), and the loop stops when all fed tokens are
    <PAD> (eval.py:116).  Returns the top-1 id list per sample."""
    B = feats.shape[0]
    H = P["rnn.weight_hh_l0"].shape[1]
    V = P["out.weight"].shape[0]
    is_lstm = model_name == "LSTM"
    inputs = [torch.full((1, B), SOS, dtype=torch.long)]
    hiddens = [zero_hidden(model_name, n_layers, B, H, feats)]
    cum = [torch.zeros(B, dtype=feats.dtype)]                           # log(1.)
    outputs: List[List[List[int]]] = [[[]] for _ in range(B)]
    for t in range(caption_max_len + 1):
        cand = []
        nxt_h = []
        for i, (tok, hid, cp) in enumerate(zip(inputs, hiddens, cum)):
            logits, nh = decoder_step(P, tok, hid, feats, model_name=model_name, n_layers=n_layers,
                                      embedding_scale=embedding_scale)
            nxt_h.append(nh)
            seq_len = torch.full((B,), float(t + 1), dtype=feats.dtype)
            for b in range(B):
                seq = outputs[b][i]
                if EOS in seq:   # np.where + fancy assignment at eval.py:51-54: the LAST <EOS> position wins
                    seq_len[b] = len(seq) - seq[::-1].index(EOS)
            cp = cp / seq_len ** 0.7
            cand.append(torch.log(torch.sigmoid(logits)) + cp.unsqueeze(1))
        flat = torch.cat(cand, dim=1)                                   # (B, beams*V)
        top_p, top_i = flat.topk(beam_width, dim=1)
        top_p, top_i = top_p.t(), top_i.t()                             # (k,B)
        tok_ids, src = top_i % V, top_i // V
        new_hiddens = []
        for k in range(beam_width):
            if is_lstm:
                h = torch.stack([nxt_h[int(src[k, b])][0][:, b] for b in range(B)], dim=1)
                c = torch.stack([nxt_h[int(src[k, b])][1][:, b] for b in range(B)], dim=1)
                new_hiddens.append((h, c))
            else:
                new_hiddens.append(torch.stack([nxt_h[int(src[k, b])][:, b] for b in range(B)], dim=1))
        new_outputs = [[outputs[b][int(src[k, b])] + [int(tok_ids[k, b])] for k in range(beam_width)] for b in range(B)]
        inputs = [tok_ids[k].view(1, -1) for k in range(beam_width)]
        hiddens, cum, outputs = new_hiddens, [top_p[k] for k in range(beam_width)], new_outputs
        if t == caption_max_len or bool((torch.cat(inputs) == PAD).all()):
            break
    return [o[0] for o in outputs]


# ----------------------------------------------------------------------------
# helpers shared by tests / bench (synthetic MSVD-shaped inputs)
# ----------------------------------------------------------------------------
def synthetic_batch(B: int, T: int, E: int, V: int, caption_max_len: int = 30, seed: int = 1234,
                    full_length_first: bool = True, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, T, E, generator=g, dtype=torch.float32).to(dtype)
    lo = min(3, caption_max_len)
    lens = torch.randint(lo, caption_max_len + 1, (B,), generator=g)
    if full_length_first:
        lens[0] = caption_max_len
    targets = torch.zeros(caption_max_len + 1, B, dtype=torch.long)
    for b in range(B):
        n = int(lens[b])
        targets[:n, b] = torch.randint(3, V, (n,), generator=g)
        targets[n, b] = EOS
    return feats, targets, targets > PAD


def init_decoder_params(V: int, EMB: int, E: int, H: int, A: int, n_layers: int = 1, model_name: str = "LSTM",
                        seed: int = 0, dtype=torch.float32) -> Params:
    """Same distributions as the reference's default nn.Module initialisers
    (not the same draws; parity tests load identical tensors on both sides)."""
    g = torch.Generator().manual_seed(seed)
    G = 4 if model_name == "LSTM" else 3
    u = lambda shape, k: ((torch.rand(*shape, generator=g) * 2 - 1) * k).to(dtype)
    P = {
        "attn_b": torch.ones(A, dtype=dtype),
        "embedding.weight": torch.randn(V, EMB, generator=g).to(dtype),
        "attn_W.weight": u((A, H), 1 / math.sqrt(H)),
        "attn_U.weight": u((A, E), 1 / math.sqrt(E)),
        "attn_w.weight": u((1, A), 1 / math.sqrt(A)),
    }
    for l in range(n_layers):
        I = EMB + E if l == 0 else H
        k = 1 / math.sqrt(H)
        P[f"rnn.weight_ih_l{l}"] = u((G * H, I), k)
        P[f"rnn.weight_hh_l{l}"] = u((G * H, H), k)
        P[f"rnn.bias_ih_l{l}"] = u((G * H,), k)
        P[f"rnn.bias_hh_l{l}"] = u((G * H,), k)
    P["out.weight"] = u((V, H), 1 / math.sqrt(H))
    P["out.bias"] = u((V,), 1 / math.sqrt(H))
    return P


def init_reconstructor_params(kind: str, Hdec: int, R: int, A: int = 128, n_layers: int = 1,
                              model_name: str = "LSTM", seed: int = 1, dtype=torch.float32) -> Params:
    g = torch.Generator().manual_seed(seed)
    G = 4 if model_name == "LSTM" else 3
    u = lambda shape, k: ((torch.rand(*shape, generator=g) * 2 - 1) * k).to(dtype)
    P: Params = {}
    if kind == "local":
        P["attn_b"] = torch.ones(A, dtype=dtype)
        P["attn_W.weight"] = u((A, R), 1 / math.sqrt(R))
        P["attn_U.weight"] = u((A, Hdec), 1 / math.sqrt(Hdec))
        P["attn_w.weight"] = u((1, A), 1 / math.sqrt(A))
        I0 = Hdec
    elif kind == "global":
        I0 = 2 * Hdec
    else:
        raise NotImplementedError("Unknown reconstructor: {}".format(kind))
    for l in range(n_layers):
        I = I0 if l == 0 else R
        k = 1 / math.sqrt(R)
        P[f"rnn.weight_ih_l{l}"] = u((G * R, I), k)
        P[f"rnn.weight_hh_l{l}"] = u((G * R, R), k)
        P[f"rnn.bias_ih_l{l}"] = u((G * R,), k)
        P[f"rnn.bias_hh_l{l}"] = u((G * R,), k)
    P["out.weight"] = u((R, R), 1 / math.sqrt(R))
    P["out.bias"] = u((R,), 1 / math.sqrt(R))
    return P
