"""Drop-in mirrors of the reference's three nn.Modules.

Same constructor kwargs, same per-step ``forward`` signatures, same ``state_dict`` keys / shapes / gate order
as models/decoder.py, models/global_reconstructor.py and models/local_reconstructor.py of the reference, so a
reference checkpoint loads with ``load_state_dict``.  The arithmetic runs in librecnet_b200.so:

* ``forward`` (one timestep, models/decoder.py:45) -> operator-level kernels (``ops.py``)
* ``forward_sequence`` (whole teacher-forced loop, what train.forward_* use) -> one C call (``functional.py``)

One extra constructor kwarg, ``precision`` ("bf16" | "fp32"); everything else is positional-compatible.
Built: Decoder and both reconstructors with LSTM or GRU cells (GRU is the reference's default decoder, config.py:31),
n_layers == 1.  Anything else raises NotImplementedError (never a silent fallback).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from . import functional as Fn
from . import ops
from .rnn_params import RNNParams


def _precision_id(p) -> int:
    if isinstance(p, int):
        return p
    if p not in L.PRECISIONS:
        raise ValueError(f"precision must be one of {list(L.PRECISIONS)}, got {p!r}")
    return L.PRECISIONS[p]


class _RngMixin:
    """(seed, offset) pair living on the device; the offset is bumped once per forward so that every
    iteration draws fresh Philox dropout masks, also under CUDA-graph replay."""

    def _init_rng(self, seed: int = 0x5EED):
        self.register_buffer("_rng", torch.tensor([seed, 0], dtype=torch.int64), persistent=False)

    def _next_rng(self) -> torch.Tensor:
        if self.training:
            self._rng[1] += 1
        return self._rng

    def seed_dropout(self, seed: int):
        self._rng[0] = int(seed)
        self._rng[1] = 0


def _require_supported(model_name: str, n_layers: int, what: str, gru_ok: bool = False, max_layers: int = 1):
    if model_name != "LSTM" and not gru_ok:
        raise NotImplementedError(f"{what}: model_name={model_name!r} (GRU) is not built yet in recnet_b200; use 'LSTM'")
    if n_layers > max_layers or (n_layers > 1 and model_name != "LSTM"):
        raise NotImplementedError(f"{what}: n_layers={n_layers} with {model_name} cells is not built yet in recnet_b200")


class Decoder(nn.Module, _RngMixin):
    """models/decoder.py:6-70."""

    def __init__(self, model_name, n_layers, encoder_size, embedding_size, embedding_scale, hidden_size,
                 attn_size, output_size, embedding_dropout, dropout, out_dropout, precision="bf16"):
        super().__init__()
        self.model_name = model_name
        self.n_layers = n_layers
        self.encoder_size = encoder_size
        self.embedding_size = embedding_size
        self.embedding_scale = embedding_scale
        self.hidden_size = hidden_size
        self.attn_size = attn_size
        self.output_size = output_size
        self.embedding_dropout_p = embedding_dropout
        self.dropout_p = dropout
        self.out_dropout_p = out_dropout
        self.precision = precision

        # parameter holders, created in the reference's order (decoder.py:22-42) so equal seeds give equal draws
        self.embedding = nn.Embedding(output_size, embedding_size)
        self.attn_W = nn.Linear(hidden_size, attn_size, bias=False)
        self.attn_U = nn.Linear(encoder_size, attn_size, bias=False)
        self.attn_b = nn.Parameter(torch.ones(attn_size), requires_grad=True)
        self.attn_w = nn.Linear(attn_size, 1, bias=False)
        self.rnn = RNNParams(model_name, embedding_size + encoder_size, hidden_size, n_layers, dropout)
        self.out = nn.Linear(hidden_size, output_size)
        self._init_rng(0xDEC0)
        self._uv_cache = None

    # ---- helpers ----
    def _params(self):
        w_ih, w_hh, b_ih, b_hh = self.rnn.layer(0)
        base = (self.embedding.weight, self.attn_W.weight, self.attn_U.weight, self.attn_b, self.attn_w.weight,
                w_ih, w_hh, b_ih, b_hh, self.out.weight, self.out.bias)
        extra = tuple(t for l in range(1, self.n_layers) for t in self.rnn.layer(l))       # stacked layers 1..NL-1
        return base + extra

    def _meta(self):
        return dict(H=self.hidden_size, A=self.attn_size, EMB=self.embedding_size, V=self.output_size,
                    precision=_precision_id(self.precision), train=self.training, embedding_scale=self.embedding_scale,
                    p_emb=self.embedding_dropout_p, p_out=self.out_dropout_p, p_layer=self.dropout_p,
                    cell=L.CELL_LSTM if self.model_name == "LSTM" else L.CELL_GRU)       # decoder.py:32-35

    # ---- whole teacher-forced loop: one C call ----
    def forward_sequence(self, tokens_in, targets, ce_weight, encoder_outputs):
        """tokens_in/targets (L,B) int64, ce_weight (L,B) f32 -> (ce scalar, hiddens (L,NL,B,H), reg = sum_p ||p||)."""
        _require_supported(self.model_name, self.n_layers, "Decoder", gru_ok=True, max_layers=L.MAX_LAYERS)
        return Fn.DecoderSequenceFn.apply(self._meta(), encoder_outputs, tokens_in, targets, ce_weight, self._next_rng(),
                                          *self._params())

    @torch.no_grad()
    def teacher_forced_logits(self, tokens_in, encoder_outputs):
        _require_supported(self.model_name, self.n_layers, "Decoder", gru_ok=True, max_layers=L.MAX_LAYERS)
        return Fn.decoder_teacher_forced_logits(self._meta(), encoder_outputs, tokens_in, self._rng, self._params())

    @torch.no_grad()
    def greedy(self, encoder_outputs, max_steps):
        """eval.greedy_search (eval.py:19-33) on device: returns (ids (n,B) int64 on device, n)."""
        import ctypes as C
        _require_supported(self.model_name, self.n_layers, "Decoder", gru_ok=True, max_layers=L.MAX_LAYERS)
        if self.n_layers > 1:       # stacked decoder: step-by-step through Decoder.forward (our kernels), feedback on device
            B = encoder_outputs.shape[0]
            dev = encoder_outputs.device
            z = lambda: torch.zeros(self.n_layers, B, self.hidden_size, device=dev)
            hid, tok = (z(), z()), torch.ones(1, B, dtype=torch.long, device=dev)
            was_training = self.training
            self.eval()
            ids = torch.zeros(max_steps, B, dtype=torch.long, device=dev)
            n = max_steps
            for t in range(max_steps):
                logits, hid = self.forward(tok, hid, encoder_outputs)
                tok = logits.argmax(dim=1).view(1, -1)
                ids[t] = tok[0]
                if bool((tok == 0).all()):
                    n = t + 1
                    break
            self.train(was_training)
            return ids, torch.tensor([n], dtype=torch.int32, device=dev)
        lib = L.lib()
        feats = Fn._f32c(encoder_outputs, "encoder_outputs")
        B, T, E = feats.shape
        meta = self._meta()
        d = L.decoder_desc(B=B, T=T, E=E, H=meta["H"], A=meta["A"], EMB=meta["EMB"], V=meta["V"], L=1,
                           precision=meta["precision"], train=0, embedding_scale=float(meta["embedding_scale"]),
                           p_emb_drop=0.0, p_out_drop=0.0, cell=meta["cell"])
        nbytes = lib.recnet_greedy_workspace_bytes(C.byref(d))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=feats.device)
        ids = torch.empty(max_steps, B, dtype=torch.int64, device=feats.device)
        n = torch.zeros(1, dtype=torch.int32, device=feats.device)
        params = tuple(Fn._f32c(p, "param") for p in self._params())
        w = Fn._pack(L.decoder_tensors, params)
        L.check(lib.recnet_decoder_greedy(C.byref(d), C.byref(w), feats.data_ptr(), max_steps, ws.data_ptr(), nbytes,
                                          ids.data_ptr(), n.data_ptr(), Fn._stream()), "recnet_decoder_greedy")
        return ids, n

    # ---- single timestep (reference API) ----
    def forward(self, input, hidden, encoder_outputs):
        """input (1,B) int64; hidden ((NL,B,H),(NL,B,H)) [LSTM] or (1,B,H) [GRU]; encoder_outputs (B,T,E) -> (logits (B,V), hidden)."""
        _require_supported(self.model_name, self.n_layers, "Decoder", gru_ok=True, max_layers=L.MAX_LAYERS)
        p = _precision_id(self.precision)
        is_lstm = self.model_name == "LSTM"
        h, c = (hidden[0][-1], hidden[1][-1]) if is_lstm else (hidden[-1], None)
        emb = torch.nn.functional.embedding(input[0], self.embedding.weight) * self.embedding_scale      # decoder.py:46-47
        emb = torch.nn.functional.dropout(emb, self.embedding_dropout_p, self.training)                   # decoder.py:48
        # U.v is time-invariant (the reference recomputes it every step, decoder.py:54).  Without autograd (greedy / beam
        # loops over this method) it is cached across the steps of one sequence; with autograd it is recomputed so that
        # every step owns its graph.  The training hot path (forward_sequence) hoists it out of the loop altogether.
        B, T, E = encoder_outputs.shape
        need_grad = torch.is_grad_enabled() and self.attn_U.weight.requires_grad
        key = (encoder_outputs.data_ptr(), encoder_outputs._version, self.attn_U.weight._version, (B, T, E), p)
        if need_grad or self._uv_cache is None or self._uv_cache[0] != key:
            Uv = ops.linear(encoder_outputs.reshape(B * T, E), self.attn_U.weight, None, p).view(B, T, -1)
            self._uv_cache = None if need_grad else (key, Uv)
        else:
            Uv = self._uv_cache[1]
        Wh = ops.linear(h, self.attn_W.weight, None, p)                                                   # decoder.py:51
        ctx = ops.additive_attention(Wh, Uv, self.attn_b, self.attn_w.weight, encoder_outputs, p)          # decoder.py:55-61
        w_ih, w_hh, b_ih, b_hh = self.rnn.layer(0)
        gi, gh = ops.linear(torch.cat((emb, ctx), dim=1), w_ih, b_ih, p), ops.linear(h, w_hh, b_hh, p)     # decoder.py:64-66
        if self.n_layers > 1:          # stacked LSTM: layer 0 takes [emb ; ctx], layer l the (dropped-out) output of layer l-1
            hs, cs, x = [], [], torch.cat((emb, ctx), dim=1)
            for l in range(self.n_layers):
                w_ih, w_hh, b_ih, b_hh = self.rnn.layer(l)
                pre = ops.linear(x, w_ih, b_ih, p) + ops.linear(hidden[0][l], w_hh, b_hh, p)
                hl, cl = ops.lstm_cell(pre, hidden[1][l], p)
                hs.append(hl); cs.append(cl)
                x = torch.nn.functional.dropout(hl, self.dropout_p, self.training) if l < self.n_layers - 1 else hl
            logits = ops.linear(hs[-1], self.out.weight, self.out.bias, p)
            logits = torch.nn.functional.dropout(logits, self.out_dropout_p, self.training)
            return logits, (torch.stack(hs), torch.stack(cs))
        if is_lstm:
            h2, c2 = ops.lstm_cell(gi + gh, c, p)
        else:
            h2 = ops.gru_cell(gi, gh, h, p)
        logits = ops.linear(h2, self.out.weight, self.out.bias, p)                                        # decoder.py:68
        logits = torch.nn.functional.dropout(logits, self.out_dropout_p, self.training)                    # decoder.py:69
        return logits, ((h2.unsqueeze(0), c2.unsqueeze(0)) if is_lstm else h2.unsqueeze(0))


class GlobalReconstructor(nn.Module, _RngMixin):
    """models/global_reconstructor.py:6-46."""

    def __init__(self, model_name, n_layers, decoder_hidden_size, hidden_size, dropout, decoder_dropout, caption_max_len,
                 precision="bf16"):
        super().__init__()
        self.model_name = model_name
        self.n_layers = n_layers
        self.decoder_hidden_size = decoder_hidden_size
        self.hidden_size = hidden_size
        self.dropout_p = dropout
        self.decoder_dropout_p = decoder_dropout
        self.caption_max_len = caption_max_len
        self.precision = precision
        self.rnn = RNNParams(model_name, decoder_hidden_size * 2, hidden_size, n_layers, dropout)
        self.out = nn.Linear(hidden_size, hidden_size)
        self._init_rng(0x610B)

    def _params(self):
        w_ih, w_hh, b_ih, b_hh = self.rnn.layer(0)
        return (w_ih, w_hh, b_ih, b_hh, self.out.weight, self.out.bias)

    def forward_sequence(self, decoder_hiddens, encoder_outputs):
        """decoder_hiddens (L,1,B,H) or (L,B,H); encoder_outputs (B,T,R) -> (MSE(mean_t out, mean_tau feats) / L, reg = sum_p ||p||)."""
        _require_supported(self.model_name, self.n_layers, "GlobalReconstructor", gru_ok=True)
        hid = _sequence_hiddens(decoder_hiddens, self.model_name)
        meta = dict(precision=_precision_id(self.precision), train=self.training, p_drop=self.decoder_dropout_p,
                    caption_max_len=self.caption_max_len, cell=L.CELL_LSTM if self.model_name == "LSTM" else L.CELL_GRU)
        return Fn.GlobalReconstructorFn.apply(meta, hid, encoder_outputs, self._next_rng(), *self._params())

    def forward(self, input, hidden, decoder_hiddens):
        """input (1,B,H) = decoder_hiddens[t]; hidden ((1,B,R),(1,B,R)) [LSTM] or (1,B,R) [GRU]; decoder_hiddens (L,1,B,H)."""
        _require_supported(self.model_name, self.n_layers, "GlobalReconstructor", gru_ok=True)
        p = _precision_id(self.precision)
        is_lstm = self.model_name == "LSTM"
        h_prev = hidden[0][-1] if is_lstm else hidden[-1]
        Lsteps = decoder_hiddens.size(0)
        mp = decoder_hiddens.mean(0).mean(0) / Lsteps * self.caption_max_len            # global_reconstructor.py:33-37
        mp = torch.nn.functional.dropout(mp, self.decoder_dropout_p, self.training)      # :38
        x = torch.cat((input[0], mp), 1)                                                 # :40
        w_ih, w_hh, b_ih, b_hh = self.rnn.layer(0)
        gi, gh = ops.linear(x, w_ih, b_ih, p), ops.linear(h_prev, w_hh, b_hh, p)         # :43
        if is_lstm:
            h2, c2 = ops.lstm_cell(gi + gh, hidden[1][-1], p)
        else:
            h2 = ops.gru_cell(gi, gh, h_prev, p)
        out = ops.linear(h2, self.out.weight, self.out.bias, p)                          # :45
        return out, ((h2.unsqueeze(0), c2.unsqueeze(0)) if is_lstm else h2.unsqueeze(0))


class LocalReconstructor(nn.Module, _RngMixin):
    """models/local_reconstructor.py:6-55."""

    def __init__(self, model_name, n_layers, decoder_hidden_size, hidden_size, dropout, decoder_dropout, attn_size,
                 precision="bf16"):
        super().__init__()
        self.model_name = model_name
        self.n_layers = n_layers
        self.decoder_hidden_size = decoder_hidden_size
        self.hidden_size = hidden_size
        self.dropout_p = dropout
        self.decoder_dropout_p = decoder_dropout
        self.attn_size = attn_size
        self.precision = precision
        self.attn_W = nn.Linear(hidden_size, attn_size, bias=False)
        self.attn_U = nn.Linear(decoder_hidden_size, attn_size, bias=False)
        self.attn_b = nn.Parameter(torch.ones(attn_size), requires_grad=True)
        self.attn_w = nn.Linear(attn_size, 1, bias=False)
        self.rnn = RNNParams(model_name, decoder_hidden_size, hidden_size, n_layers, dropout)
        self.out = nn.Linear(hidden_size, hidden_size)
        self._init_rng(0x10CA)

    def _params(self):
        w_ih, w_hh, b_ih, b_hh = self.rnn.layer(0)
        return (self.attn_W.weight, self.attn_U.weight, self.attn_b, self.attn_w.weight, w_ih, w_hh, b_ih, b_hh,
                self.out.weight, self.out.bias)

    def forward_sequence(self, decoder_hiddens, encoder_outputs):
        """decoder_hiddens (L,1,B,H) or (L,B,H); encoder_outputs (B,S,R) -> (MSELoss(outputs^T, encoder_outputs), reg = sum_p ||p||)."""
        _require_supported(self.model_name, self.n_layers, "LocalReconstructor", gru_ok=True)
        hid = _sequence_hiddens(decoder_hiddens, self.model_name)
        meta = dict(A=self.attn_size, precision=_precision_id(self.precision), train=self.training,
                    p_drop=self.decoder_dropout_p, cell=L.CELL_LSTM if self.model_name == "LSTM" else L.CELL_GRU)
        return Fn.LocalReconstructorFn.apply(meta, hid, encoder_outputs, self._next_rng(), *self._params())

    def forward(self, hidden, decoder_hiddens):
        """hidden ((1,B,R),(1,B,R)) [LSTM] or (1,B,R) [GRU]; decoder_hiddens (L,1,B,H) -> (out (B,R), hidden)."""
        _require_supported(self.model_name, self.n_layers, "LocalReconstructor", gru_ok=True)
        p = _precision_id(self.precision)
        is_lstm = self.model_name == "LSTM"
        h_prev = hidden[0][-1] if is_lstm else hidden[-1]
        hid = _squeeze_layers(decoder_hiddens)                                           # (L,B,H)
        Lsteps, B, H = hid.shape
        Uv = ops.linear(hid.reshape(Lsteps * B, H), self.attn_U.weight, None, p).view(Lsteps, B, -1).transpose(0, 1)
        Wh = ops.linear(h_prev, self.attn_W.weight, None, p)                             # local_reconstructor.py:39
        x = ops.additive_attention(Wh, Uv.contiguous(), self.attn_b, self.attn_w.weight, hid.transpose(0, 1).contiguous(), p)
        x = torch.nn.functional.dropout(x, self.decoder_dropout_p, self.training)        # :50
        w_ih, w_hh, b_ih, b_hh = self.rnn.layer(0)
        gi, gh = ops.linear(x, w_ih, b_ih, p), ops.linear(h_prev, w_hh, b_hh, p)         # :52
        if is_lstm:
            h2, c2 = ops.lstm_cell(gi + gh, hidden[1][-1], p)
        else:
            h2 = ops.gru_cell(gi, gh, h_prev, p)
        out = ops.linear(h2, self.out.weight, self.out.bias, p)                          # :54
        return out, ((h2.unsqueeze(0), c2.unsqueeze(0)) if is_lstm else h2.unsqueeze(0))


def _sequence_hiddens(decoder_hiddens: torch.Tensor, model_name: str) -> torch.Tensor:
    """(L,1,B,H) -> (L,B,H); a stacked decoder's (L,NLdec,B,H) passes through (LSTM reconstructor cells only)."""
    if decoder_hiddens.dim() == 4 and decoder_hiddens.size(1) != 1:
        if model_name != "LSTM":
            raise NotImplementedError("GRU reconstructors over a stacked decoder are not built in recnet_b200")
        return decoder_hiddens
    return _squeeze_layers(decoder_hiddens)


def _squeeze_layers(decoder_hiddens: torch.Tensor) -> torch.Tensor:
    if decoder_hiddens.dim() == 4:
        if decoder_hiddens.size(1) != 1:
            raise NotImplementedError("per-step reconstructor forward over a stacked decoder is not built in recnet_b200; use forward_sequence")
        return decoder_hiddens[:, 0]
    return decoder_hiddens
