"""torch.autograd wrappers over the C ABI (sequence level).

PyTorch supplies device memory, the current stream and the autograd graph; every kernel that runs is ours
(librecnet_b200.so).  Each Function owns a workspace tensor (raw bytes) that carries the stashed activations
from forward to backward, like cuDNN's reserve space would for nn.LSTM -- except nothing here calls cuDNN.

Each sequence Function also returns ``reg = sum_p ||p||_2`` over its module's parameters (the reference adds
``lambda_reg * reg`` to every loss, train.py:69,101,127) and writes ALL parameter gradients of the module --
BPTT gradients plus the regulariser's g * p / ||p|| -- into ONE flat buffer whose views become ``p.grad``.
That single contiguous buffer per module is what the data-parallel all-reduce sends (parallel.py).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib as L


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_cuda:
        raise RuntimeError(f"{name}: expected a float32 CUDA tensor, got {t.dtype} on {t.device} (recnet_b200 has no CPU path)")
    return t if t.is_contiguous() else t.contiguous()


def _pack(struct_cls, tensors: Sequence[torch.Tensor]):
    """Fill a C tensor-table struct.  Decoder: the first 11 tensors are layer 0 (+ embedding / attention / out); every
    further group of 4 is (weight_ih, weight_hh, bias_ih, bias_hh) of one extra stacked layer."""
    s = struct_cls()
    nbase = len(struct_cls.FIELDS)
    for name, t in zip(struct_cls.FIELDS, tensors):
        setattr(s, name, t.data_ptr())
    extra = tensors[nbase:]
    if extra:
        assert hasattr(struct_cls, "EXTRA") and len(extra) % 4 == 0 and len(extra) // 4 <= L.MAX_LAYERS - 1
        for li in range(len(extra) // 4):
            for k, name in enumerate(struct_cls.EXTRA):
                getattr(s, name)[li] = extra[4 * li + k].data_ptr()
    return s


# ---- multi-tensor tables for the L2-norm regulariser ------------------------------------------------------------
class _NormTable:
    """Device-side (pointer, size, block map) tables for the multi-tensor norm kernels, cached per parameter list."""
    CHUNK = 16384

    def __init__(self, params: Sequence[torch.Tensor]):
        dev = params[0].device
        self.n = len(params)
        self.ptrs = torch.tensor([p.data_ptr() for p in params], dtype=torch.int64, device=dev)
        self.sizes = torch.tensor([p.numel() for p in params], dtype=torch.int64, device=dev)
        bt, bc, offs, o = [], [], [], 0
        for i, p in enumerate(params):
            for c in range((p.numel() + self.CHUNK - 1) // self.CHUNK):
                bt.append(i)
                bc.append(c)
            offs.append(o)
            o += (p.numel() + 63) // 64 * 64          # keep every gradient view 256-byte aligned
        self.total, self.offsets = o, offs
        self.offset_bytes = torch.tensor([x * 4 for x in offs], dtype=torch.int64, device=dev)
        self.blk_tensor = torch.tensor(bt, dtype=torch.int32, device=dev)
        self.blk_chunk = torch.tensor(bc, dtype=torch.int32, device=dev)
        self.n_blocks = len(bt)


_norm_tables: Dict[tuple, _NormTable] = {}


def _table_for(params) -> _NormTable:
    key = tuple((p.data_ptr(), p.numel()) for p in params)
    t = _norm_tables.get(key)
    if t is None:
        t = _norm_tables[key] = _NormTable(params)
    return t


def _norms_alloc(params, lambda_dev=None):
    """Output buffers of _norms_fwd (sumsq, reg, fused or None, partial), allocated on the CURRENT stream -- a caller that runs _norms_fwd
    on another stream allocates them first, where they are used and freed."""
    tab = _table_for(params)
    dev = params[0].device
    return (torch.empty(tab.n, dtype=torch.float32, device=dev), torch.empty((), dtype=torch.float32, device=dev),
            torch.empty((), dtype=torch.float32, device=dev) if lambda_dev is not None else None,
            torch.empty(tab.n_blocks, dtype=torch.float32, device=dev))


def _norms_fwd(params, base=None, lambda_dev=None, bufs=None):
    """reg = sum_p ||p||  (+ fused = base + lambda_dev * reg when ``lambda_dev`` is given: the loss assembly of train.py:70,102,128
    inside the finalize kernel).  Returns (reg, sumsq, fused or None)."""
    tab = _table_for(params)
    dev = params[0].device
    if bufs is None:
        bufs = _norms_alloc(params, lambda_dev)
    sumsq, reg, fused, spare = bufs
    pre = _prefetched_norms.pop(tuple((p.data_ptr(), p.numel()) for p in params), None)
    if pre is not None:                     # the squared-norm partials were computed ahead of the forward pass (prefetch_param_norms)
        partial, ready = pre
        torch.cuda.current_stream().wait_event(ready)
        L.check(L.lib().recnet_param_norms_finalize(partial.data_ptr(), tab.blk_tensor.data_ptr(), tab.n_blocks, tab.n, sumsq.data_ptr(),
                                                    reg.data_ptr(), _ptr(base), _ptr(lambda_dev), _ptr(fused), _stream()),
                "recnet_param_norms_finalize")
        _bg.keep.append(partial)
        return reg, sumsq, fused
    partial = spare
    L.check(L.lib().recnet_param_norms_fwd(tab.ptrs.data_ptr(), tab.sizes.data_ptr(), tab.n, tab.blk_tensor.data_ptr(),
                                           tab.blk_chunk.data_ptr(), tab.n_blocks, partial.data_ptr(), sumsq.data_ptr(),
                                           reg.data_ptr(), _ptr(base), _ptr(lambda_dev), _ptr(fused), _stream()),
            "recnet_param_norms_fwd")
    return reg, sumsq, fused


def _lambda_scalar(meta, dev):
    """meta['lambda_reg'] (the module dict's device scalar, train.py:151,188) as a float32 CUDA 0-dim tensor, or None."""
    lam = meta.get("lambda_reg")
    if lam is None:
        return None
    if not torch.is_tensor(lam):
        lam = torch.tensor(float(lam), dtype=torch.float32, device=dev)
    lam = lam.detach()
    if lam.dtype != torch.float32 or lam.device != dev:
        lam = lam.to(device=dev, dtype=torch.float32)
    return lam.reshape(())


def _flat_grads(params) -> (torch.Tensor, List[torch.Tensor], torch.Tensor):
    """One flat fp32 buffer + per-parameter views + device table of the views' addresses (computed with a device-side
    add so that it is valid under CUDA-graph capture, where the buffer address is fixed)."""
    tab = _table_for(params)
    key = tuple(p.data_ptr() for p in params)
    flat = _flat_placement.get(key)          # data-parallel runs: a fixed slice of symmetric memory (parallel.NvlinkAllReducer)
    if flat is None or flat.numel() != tab.total:
        flat = torch.empty(tab.total, dtype=torch.float32, device=params[0].device)
    views = [flat[o: o + p.numel()].view_as(p) for o, p in zip(tab.offsets, params)]
    _last_flat.pop(key, None)
    _last_flat[key] = flat                                          # see flat_buffer_of(); newest last
    while len(_last_flat) > 8:                                      # a process trains a handful of modules; do not pin old buffers
        _last_flat.pop(next(iter(_last_flat)))
    return flat, views, tab.offset_bytes + flat.data_ptr()


# The newest flat gradient buffer per parameter list.  autograd stores the views it is handed DETACHED (``p.grad._base`` is
# None although the memory is still the flat buffer's), so whoever wants the contiguous buffer back -- the data-parallel reducer,
# to send one all-reduce per module instead of one per tensor -- asks here.  Holding the reference also keeps the buffer from
# being recycled while the optimiser still reads its views.
_last_flat: Dict[tuple, torch.Tensor] = {}

# parameter list (tuple of data_ptrs) -> preallocated flat gradient buffer.  The NVLink all-reduce (csrc/allreduce.cuh) works in place
# on symmetric memory, so the data-parallel reducer places every module's flat gradient buffer there once and backward writes into it.
_flat_placement: Dict[tuple, torch.Tensor] = {}


def place_flat_grads(params, buffer: Optional[torch.Tensor]) -> int:
    """Make ``buffer`` (float32, >= flat_grad_numel(params) elements; None = back to fresh allocations) the flat gradient buffer of this
    parameter list from now on.  Returns the number of elements used."""
    params = [p for p in params]
    tab = _table_for(params)
    key = tuple(p.data_ptr() for p in params)
    if buffer is None:
        _flat_placement.pop(key, None)
        return tab.total
    if buffer.dtype != torch.float32 or not buffer.is_contiguous() or buffer.numel() < tab.total:
        raise ValueError("place_flat_grads: need a contiguous float32 buffer of at least flat_grad_numel(params) elements")
    _flat_placement[key] = buffer[: tab.total]
    return tab.total


def flat_grad_numel(params) -> int:
    """Elements of the flat gradient buffer of a parameter list (every view 256-byte aligned)."""
    return _table_for([p for p in params]).total


def flat_buffer_of(grads: Sequence[torch.Tensor]):
    """The flat buffer of the most recent backward that contains every tensor of ``grads`` (all of them, each entirely), or None."""
    if not grads:
        return None
    for flat in _last_flat.values():
        if flat.device != grads[0].device or flat.dtype != grads[0].dtype:
            continue
        lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * flat.element_size()
        if all(g.is_contiguous() and lo <= g.data_ptr() and g.data_ptr() + g.numel() * g.element_size() <= hi for g in grads):
            return flat
    return None


def _norms_bwd_into(params, sumsq, g_reg, gptrs, accumulate: bool, lambda_dev=None):
    tab = _table_for(params)
    L.check(L.lib().recnet_param_norms_bwd(tab.ptrs.data_ptr(), gptrs.data_ptr(), tab.sizes.data_ptr(), tab.n,
                                           tab.blk_tensor.data_ptr(), tab.blk_chunk.data_ptr(), tab.n_blocks, sumsq.data_ptr(),
                                           g_reg.data_ptr(), 1.0, _ptr(lambda_dev), int(accumulate), _stream()),
            "recnet_param_norms_bwd")


# (workspace, byte offset of the loop kernel's error flag) of the most recent sequence calls, for check_loop_status()
_status_slots = {}


def _remember_status(kind: str, ws: torch.Tensor, offset: int):
    _status_slots[kind] = (ws, int(offset))


def check_loop_status() -> None:
    """Synchronise and raise if any persistent loop kernel reported a protocol timeout (see include/recnet_b200.h)."""
    torch.cuda.synchronize()
    for kind, (ws, off) in _status_slots.items():
        code = int(ws[off: off + 4].view(torch.int32).item())
        if code != 0:
            raise RuntimeError(f"recnet_b200 loop kernel ({kind}) failed with device status {code} "
                               "(2 = mbarrier timeout, 3 = grid-barrier timeout)")


def _scalar(g, dev):
    if g is None:
        return torch.zeros((), dtype=torch.float32, device=dev)
    return g.contiguous().float()


# ----------------------------------------------------------------------------------------------------------------
def _decoder_fwd_raw(meta, feats, tokens_in, targets, ce_weight, rng, params, split=False):
    lib = L.lib()
    feats = _f32c(feats, "encoder_outputs")
    params = tuple(_f32c(p, f"decoder parameter {i}") for i, p in enumerate(params))
    NL = 1 + (len(params) - len(L.decoder_tensors.FIELDS)) // 4
    L.require_device(feats.device.index if feats.device.index is not None else torch.cuda.current_device())
    B, T, E = feats.shape
    Lsteps = tokens_in.shape[0]
    d = L.decoder_desc(B=B, T=T, E=E, H=meta["H"], A=meta["A"], EMB=meta["EMB"], V=meta["V"], L=Lsteps,
                       precision=meta["precision"], train=int(meta["train"]),
                       embedding_scale=float(meta["embedding_scale"]), p_emb_drop=float(meta["p_emb"]),
                       p_out_drop=float(meta["p_out"]), cell=int(meta.get("cell", L.CELL_LSTM)), n_layers=NL,
                       p_layer_drop=float(meta.get("p_layer", 0.0)))
    nbytes = lib.recnet_decoder_workspace_bytes(C.byref(d))
    if nbytes < 0:
        L.check(int(nbytes), "recnet_decoder_workspace_bytes")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=feats.device)
    hiddens = torch.empty(Lsteps, NL, B, meta["H"], dtype=torch.float32, device=feats.device)     # train.py:61-64,73
    # the CE kernels overwrite `ce`; only the loss-free inference call (no targets) needs it pre-cleared
    ce = (torch.empty if (targets is not None and ce_weight is not None) else torch.zeros)((), dtype=torch.float32, device=feats.device)
    tokens_in = tokens_in.contiguous()
    targets = targets.contiguous() if targets is not None else None
    ce_weight = ce_weight.contiguous() if ce_weight is not None else None
    w = _pack(L.decoder_tensors, params)

    def phase(bits):
        L.check(lib.recnet_decoder_fwd_phase(C.byref(d), C.byref(w), feats.data_ptr(), tokens_in.data_ptr(), _ptr(targets),
                                             _ptr(ce_weight), rng.data_ptr(), ws.data_ptr(), nbytes, hiddens.data_ptr(),
                                             ce.data_ptr(), bits, _stream()), "recnet_decoder_fwd_phase")

    saved = (feats, tokens_in, targets, ce_weight, rng, *params)
    _remember_status("decoder", ws, lib.recnet_decoder_error_offset(C.byref(d)))
    if split and targets is not None and ce_weight is not None and lib.recnet_decoder_bwd_is_split(C.byref(d)):
        phase(1)                              # ... the loop: `hiddens` is complete; the caller runs phase(2) where it likes
        return ce, hiddens, ws, d, nbytes, saved, (lambda: phase(2))
    phase(3)
    return ce, hiddens, ws, d, nbytes, saved, None


class DecoderSequenceFn(torch.autograd.Function):
    """Whole teacher-forced decoder loop (train.py:17-75 over models/decoder.py:45-70).

    inputs : meta dict, feats (B,T,E), tokens_in (L,B) i64, targets (L,B) i64, ce_weight (L,B) f32, rng (2,) i64,
             then the 11 parameters in decoder_tensors.FIELDS order.
    outputs: ce (scalar: sum_t CE_t / sum_t n_t), hiddens (L,NL,B,H), reg (scalar: sum_p ||p||)
    """

    @staticmethod
    def forward(ctx, meta: Dict, feats, tokens_in, targets, ce_weight, rng, *params):
        ce, hiddens, ws, d, nbytes, saved, tail = _decoder_fwd_raw(meta, feats, tokens_in, targets, ce_weight, rng, params,
                                                                   split=_bg.split_forward)
        lam = _lambda_scalar(meta, ce.device)
        if tail is None:
            reg, sumsq, fused = _norms_fwd(saved[5:], ce, lam)
        else:
            # vocabulary projection + CE + loss assembly on the lane, next to the reconstructor's staging: nothing of it feeds the
            # reconstructor.  The trainer waits (wait_forward_tail) before it touches the loss.  Output buffers allocated HERE (main stream).
            bufs = _norms_alloc(saved[5:], lam)
            done = torch.cuda.Event()

            def work():
                tail()
                out = _norms_fwd(saved[5:], ce, lam, bufs)
                done.record(torch.cuda.current_stream())
                return out
            reg, sumsq, fused = run_in_background(work, ws, ce, hiddens, *bufs, *saved)
            _bg.forward_tails.append(done)
        ctx.desc, ctx.nbytes = d, nbytes
        ctx.lam = lam                       # not an autograd input: a plain attribute keeps it alive for backward
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(*saved, ws, sumsq, hiddens)
        if fused is not None:               # output 0 is the assembled loss ce + lambda * reg (train.py:70); reg is informational
            ctx.mark_non_differentiable(reg)
            return fused, hiddens, reg
        return ce, hiddens, reg

    @staticmethod
    def backward(ctx, g_ce, g_hiddens, g_reg):
        lib = L.lib()
        feats, tokens_in, targets, ce_weight, rng, *params, ws, sumsq, hiddens = ctx.saved_tensors
        flat, grads, gptrs = _flat_grads(params)
        g_ce = _scalar(g_ce, feats.device)
        g_hid = g_hiddens.contiguous() if g_hiddens is not None else None
        w, g = _pack(L.decoder_tensors, params), _pack(L.decoder_tensors, grads)

        def phase(bits):
            L.check(lib.recnet_decoder_bwd_phase(C.byref(ctx.desc), C.byref(w), feats.data_ptr(), tokens_in.data_ptr(),
                                                 targets.data_ptr(), ce_weight.data_ptr(), rng.data_ptr(), ws.data_ptr(), ctx.nbytes,
                                                 g_ce.data_ptr(), _ptr(g_hid), hiddens.data_ptr(), C.byref(g), bits, _stream()),
                    "recnet_decoder_bwd_phase")

        if _bg.active and lib.recnet_decoder_bwd_is_split(C.byref(ctx.desc)):
            phase(1)                                   # CE backward, gradient wrt the states through the vocabulary projection
            vocab_done = torch.cuda.Event()

            def vocab():
                phase(4)                               # the vocabulary projection's own gradients: underneath the loop, on the lane
                vocab_done.record(torch.cuda.current_stream())
            background_gemms(vocab, feats, tokens_in, targets, ce_weight, rng, ws, g_ce, g_hid, hiddens, flat, gptrs, *params)
            phase(2)                                   # the BPTT loop
            _flush(_bg.late)                           # lane: from here on, next to the weight-gradient GEMMs below
            phase(8)
            torch.cuda.current_stream().wait_event(vocab_done)
        else:
            phase(15)
        if ctx.lam is not None:
            _norms_bwd_into(params, sumsq, g_ce, gptrs, accumulate=True, lambda_dev=ctx.lam)
        elif g_reg is not None:
            _norms_bwd_into(params, sumsq, _scalar(g_reg, feats.device), gptrs, accumulate=True)
        return (None, None, None, None, None, None, *grads)


@torch.no_grad()
def decoder_teacher_forced_logits(meta: Dict, feats, tokens_in, rng, params):
    """Inference helper: stacked logits (L,B,V) and hiddens (L,B,H) of the teacher-forced loop (no loss)."""
    ce, hiddens, ws, d, nbytes, _, _ = _decoder_fwd_raw(meta, feats, tokens_in, None, None, rng, params)
    ld = C.c_int64()
    ptr = L.lib().recnet_decoder_logits(C.byref(d), ws.data_ptr(), C.byref(ld))
    off = ptr - ws.data_ptr()
    Lsteps, B = tokens_in.shape
    flat = ws[off: off + Lsteps * B * ld.value * 4].view(torch.float32).view(Lsteps, B, ld.value)
    return flat[:, :, : meta["V"]], hiddens


def _layered_shape(hiddens):
    """decoder hiddens are (L,B,H) or, from a stacked decoder, (L,NLdec,B,H) -> (L, NLdec, B, H) sizes."""
    if hiddens.dim() == 4:
        return tuple(hiddens.shape)
    if hiddens.dim() != 3:
        raise ValueError(f"decoder_hiddens must be (L,B,H) or (L,n_layers,B,H), got {tuple(hiddens.shape)}")
    return hiddens.shape[0], 1, hiddens.shape[1], hiddens.shape[2]


# ----------------------------------------------------------------------------------------------------------------
# ---- background lane -------------------------------------------------------------------------------------------------
# Between the end of the reconstructor's BPTT loop and the decoder's optimiser step the critical path is the decoder's backward loop:
# a chain of 62 latency-bound kernels on 100 CTAs.  Everything else in that window only has to be finished by the end of the step.
# With ``deferred_weight_grads()`` active (train.train_step) that work goes to a second stream, the LANE, in this order:
#   1. the local reconstructor's batched weight-gradient GEMMs       (LocalReconstructorFn.backward, right after its loop)
#   2. the decoder's vocabulary-projection gradients                  (DecoderSequenceFn.backward, right after the CE backward)
#   -- the lane then waits for the decoder's loop --
#   3. ``late`` items: reconstructor regulariser gradient + optimiser step: machine-filling elementwise kernels that would stall the
#      loop (they take the register file), so they run next to the decoder's weight-gradient GEMMs instead.
# The persistent GEMMs of 1-2 are capped to ``ctas`` CTAs (recnet_set_background_ctas) so they never hold the SMs the loop's kernels
# are waiting for.  ``join_background()`` makes the main stream wait for the lane; everything the lane reads is kept alive until then.
# Plain cross-stream dependencies: works under CUDA-graph capture.  Measured timelines: profiles/r2_j_step_timeline.md.
class _Background:
    active = False
    ctas = 48
    stream: Optional[torch.cuda.Stream] = None
    pending = False
    keep: list = []
    late: list = []
    split_forward = False
    forward_tails: list = []


_bg = _Background()


def background_stream(device=None) -> torch.cuda.Stream:
    if _bg.stream is None or (device is not None and _bg.stream.device != torch.device(device)):
        _bg.stream = torch.cuda.Stream(device=device)
    return _bg.stream


class deferred_weight_grads:
    """Context manager (see above).  ``background_ctas`` = SMs the lane's GEMMs may hold at any time."""

    def __init__(self, enabled: bool = True, background_ctas: Optional[int] = None):
        import os
        self.enabled = enabled
        self.ctas = int(os.environ.get("RECNET_BG_CTAS", "48")) if background_ctas is None else int(background_ctas)

    def __enter__(self):
        self.prev = (_bg.active, _bg.ctas)
        _bg.active, _bg.ctas = bool(self.enabled), self.ctas
        return self

    def __exit__(self, *exc):
        _bg.active, _bg.ctas = self.prev
        if exc[0] is not None:
            _bg.late.clear()
            join_background()
        return False


def background_active() -> bool:
    return _bg.active


def run_in_background(fn, *keep):
    """Run ``fn()`` with the lane as the current stream, ordered after everything already on the lane AND after what the main stream has
    issued so far.  ``keep``: tensors that must outlive the lane's work."""
    main = torch.cuda.current_stream()
    bg = background_stream(main.device)
    bg.wait_stream(main)
    _bg.pending = True
    _bg.keep.extend(keep)
    with torch.cuda.stream(bg):
        return fn()


def background_gemms(fn, *keep):
    """``run_in_background`` with the persistent GEMMs capped to the lane's CTA budget."""
    def capped():
        lib = L.lib()
        L.check(lib.recnet_set_background_ctas(_bg.ctas), "recnet_set_background_ctas")
        try:
            return fn()
        finally:
            lib.recnet_set_background_ctas(0)
    return run_in_background(capped, *keep)


# parameter list -> (squared-norm partials, event): computed on the lane at the start of a step, consumed by _norms_fwd of the same step
_prefetched_norms: Dict[tuple, tuple] = {}


def prefetch_param_norms(*param_lists):
    """Squared-norm partials of the regularisers (train.py:69,101,127) on the lane, underneath the decoder's forward loop: they depend on
    the parameters only.  ``_norms_fwd`` of the same step picks them up (and waits for the lane's event); anything left over is dropped by
    the next call."""
    _prefetched_norms.clear()
    todo = [[p for p in pl] for pl in param_lists if pl]
    if not todo:
        return

    def work():
        for params in todo:
            tab = _table_for(params)
            partial = torch.empty(tab.n_blocks, dtype=torch.float32, device=params[0].device)
            L.check(L.lib().recnet_param_norms_partial(tab.ptrs.data_ptr(), tab.sizes.data_ptr(), tab.blk_tensor.data_ptr(),
                                                       tab.blk_chunk.data_ptr(), tab.n_blocks, partial.data_ptr(), _stream()),
                    "recnet_param_norms_partial")
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            _prefetched_norms[tuple((p.data_ptr(), p.numel()) for p in params)] = (partial, ev)
    run_in_background(work)


def background_late(fn, *keep):
    _bg.late.append(fn)
    _bg.keep.extend(keep)


def _flush(items: list):
    if items:
        todo = list(items)
        items.clear()
        run_in_background(lambda: [f() for f in todo])


class split_decoder_forward:
    """Context manager for a trainer: inside it DecoderSequenceFn.forward puts everything after its time loop (vocabulary projection, CE, loss
    assembly) on the lane.  The trainer MUST call ``wait_forward_tail()`` before it uses the decoder's loss."""

    def __init__(self, enabled: bool = True):
        self.enabled = bool(enabled)

    def __enter__(self):
        self.prev = _bg.split_forward
        _bg.split_forward = self.enabled
        return self

    def __exit__(self, *exc):
        _bg.split_forward = self.prev
        return False


def wait_forward_tail():
    """The main stream waits for the forward tails issued under ``split_decoder_forward``."""
    main = torch.cuda.current_stream() if _bg.forward_tails else None
    for ev in _bg.forward_tails:
        main.wait_event(ev)
    _bg.forward_tails.clear()


def background_pending() -> bool:
    return _bg.pending or bool(_bg.late)


def join_background():
    """Queue whatever is still waiting for its slot, then make the main stream wait for the lane; its results (reconstructor
    gradients, optimiser step, vocabulary-projection gradients) may be used afterwards."""
    _flush(_bg.late)
    if _bg.pending:
        torch.cuda.current_stream().wait_stream(_bg.stream)
        _bg.pending = False
    _bg.keep.clear()


class LocalReconstructorFn(torch.autograd.Function):
    """train.forward_local_reconstructor (train.py:108-131) over LocalReconstructor.forward.
    inputs: meta, hiddens (L,B,H) or (L,NLdec,B,H), feats (B,S,R), rng, then the 10 parameters in local_tensors.FIELDS order.
    outputs: mse (scalar), reg (scalar)."""

    @staticmethod
    def forward(ctx, meta: Dict, hiddens, feats, rng, *params):
        lib = L.lib()
        hiddens, feats = _f32c(hiddens, "decoder_hiddens"), _f32c(feats, "encoder_outputs")
        params = tuple(_f32c(p, n) for p, n in zip(params, L.local_tensors.FIELDS))
        Lsteps, NLd, B, H = _layered_shape(hiddens)
        _, S, R = feats.shape
        d = L.local_desc(B=B, S=S, R=R, H=H, A=meta["A"], L=Lsteps, precision=meta["precision"], train=int(meta["train"]),
                         p_drop=float(meta["p_drop"]), cell=int(meta.get("cell", L.CELL_LSTM)), dec_layers=NLd)
        nbytes = lib.recnet_local_workspace_bytes(C.byref(d))
        if nbytes < 0:
            L.check(int(nbytes), "recnet_local_workspace_bytes")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=feats.device)
        mse = torch.empty((), dtype=torch.float32, device=feats.device)
        w = _pack(L.local_tensors, params)
        L.check(lib.recnet_local_fwd(C.byref(d), C.byref(w), hiddens.data_ptr(), feats.data_ptr(), rng.data_ptr(), ws.data_ptr(),
                                     nbytes, mse.data_ptr(), _stream()), "recnet_local_fwd")
        _remember_status("local", ws, lib.recnet_local_error_offset(C.byref(d)))
        lam = _lambda_scalar(meta, mse.device)
        reg, sumsq, fused = _norms_fwd(params, mse, lam)
        ctx.desc, ctx.nbytes = d, nbytes
        ctx.lam = lam
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(hiddens, feats, rng, ws, sumsq, *params)
        if fused is not None:               # output 0 is the assembled loss mse + lambda * reg (train.py:128-130)
            ctx.mark_non_differentiable(reg)
            return fused, reg
        return mse, reg

    @staticmethod
    def backward(ctx, g_mse, g_reg):
        lib = L.lib()
        hiddens, feats, rng, ws, sumsq, *params = ctx.saved_tensors
        flat, grads, gptrs = _flat_grads(params)
        g_hid = torch.empty_like(hiddens)
        g_mse = _scalar(g_mse, feats.device)
        w, g = _pack(L.local_tensors, params), _pack(L.local_tensors, grads)

        def phase(bits):
            L.check(lib.recnet_local_bwd_phase(C.byref(ctx.desc), C.byref(w), hiddens.data_ptr(), feats.data_ptr(), rng.data_ptr(),
                                               ws.data_ptr(), ctx.nbytes, g_mse.data_ptr(), C.byref(g), g_hid.data_ptr(), bits,
                                               _stream()), "recnet_local_bwd_phase")

        def regulariser():
            if ctx.lam is not None:
                _norms_bwd_into(params, sumsq, g_mse, gptrs, accumulate=True, lambda_dev=ctx.lam)
            elif g_reg is not None:
                _norms_bwd_into(params, sumsq, _scalar(g_reg, feats.device), gptrs, accumulate=True)

        if _bg.active and ctx.desc.dec_layers == 1:
            phase(1)                                   # BPTT loop + g_hiddens: what the decoder's backward waits for
            # keep-alive list: NOT the gradient views themselves -- autograd only adopts a gradient tensor as ``p.grad`` when nobody
            # else holds it (otherwise it clones, on the main stream, before the lane has written it); `flat` owns their storage
            background_gemms(lambda: phase(2), hiddens, feats, rng, ws, sumsq, g_mse, flat, gptrs, g_reg, ctx.lam, *params)
            background_late(regulariser)
        else:
            phase(3)
            regulariser()
        return (None, g_hid, None, None, *grads)


class GlobalReconstructorFn(torch.autograd.Function):
    """train.forward_global_reconstructor (train.py:78-105) over GlobalReconstructor.forward.
    inputs: meta, hiddens (L,B,H) or (L,NLdec,B,H), feats (B,T,R), rng, then the 6 parameters in global_tensors.FIELDS order.
    outputs: MSE(mean_t out, mean_tau feats) / L  (scalar), reg (scalar)."""

    @staticmethod
    def forward(ctx, meta: Dict, hiddens, feats, rng, *params):
        lib = L.lib()
        hiddens, feats = _f32c(hiddens, "decoder_hiddens"), _f32c(feats, "encoder_outputs")
        params = tuple(_f32c(p, n) for p, n in zip(params, L.global_tensors.FIELDS))
        Lsteps, NLd, B, H = _layered_shape(hiddens)
        _, T, R = feats.shape
        d = L.global_desc(B=B, L=Lsteps, R=R, H=H, T=T, precision=meta["precision"], train=int(meta["train"]),
                          p_drop=float(meta["p_drop"]), caption_max_len=float(meta["caption_max_len"]),
                          cell=int(meta.get("cell", L.CELL_LSTM)), dec_layers=NLd)
        nbytes = lib.recnet_global_workspace_bytes(C.byref(d))
        if nbytes < 0:
            L.check(int(nbytes), "recnet_global_workspace_bytes")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=feats.device)
        loss = torch.empty((), dtype=torch.float32, device=feats.device)
        w = _pack(L.global_tensors, params)
        L.check(lib.recnet_global_fwd(C.byref(d), C.byref(w), hiddens.data_ptr(), feats.data_ptr(), rng.data_ptr(), ws.data_ptr(),
                                      nbytes, loss.data_ptr(), _stream()), "recnet_global_fwd")
        _remember_status("global", ws, lib.recnet_global_error_offset(C.byref(d)))
        lam = _lambda_scalar(meta, loss.device)
        reg, sumsq, fused = _norms_fwd(params, loss, lam)
        ctx.desc, ctx.nbytes = d, nbytes
        ctx.lam = lam
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(hiddens, feats, rng, ws, sumsq, *params)
        if fused is not None:               # output 0 is the assembled loss (train.py:100-102)
            ctx.mark_non_differentiable(reg)
            return fused, reg
        return loss, reg

    @staticmethod
    def backward(ctx, g_loss, g_reg):
        lib = L.lib()
        hiddens, feats, rng, ws, sumsq, *params = ctx.saved_tensors
        flat, grads, gptrs = _flat_grads(params)
        g_hid = torch.empty_like(hiddens)
        g_loss = _scalar(g_loss, feats.device)
        w, g = _pack(L.global_tensors, params), _pack(L.global_tensors, grads)
        L.check(lib.recnet_global_bwd(C.byref(ctx.desc), C.byref(w), hiddens.data_ptr(), feats.data_ptr(), rng.data_ptr(),
                                      ws.data_ptr(), ctx.nbytes, g_loss.data_ptr(), C.byref(g), g_hid.data_ptr(), _stream()),
                "recnet_global_bwd")
        if ctx.lam is not None:
            _norms_bwd_into(params, sumsq, g_loss, gptrs, accumulate=True, lambda_dev=ctx.lam)
        elif g_reg is not None:
            _norms_bwd_into(params, sumsq, _scalar(g_reg, feats.device), gptrs, accumulate=True)
        return (None, g_hid, None, None, *grads)


# ----------------------------------------------------------------------------------------------------------------
class ParamNormSumFn(torch.autograd.Function):
    """Stand-alone reg = sum_p ||p||_2 over a parameter list (train.py:69,101,127); grad_p = g * p / ||p||."""

    @staticmethod
    def forward(ctx, *params):
        params = tuple(_f32c(p, "parameter") for p in params)
        reg, sumsq, _ = _norms_fwd(params)
        ctx.save_for_backward(sumsq, *params)
        return reg

    @staticmethod
    def backward(ctx, g):
        sumsq, *params = ctx.saved_tensors
        flat, grads, gptrs = _flat_grads(params)
        _norms_bwd_into(params, sumsq, _scalar(g, params[0].device), gptrs, accumulate=False)
        return tuple(grads)


def param_norm_sum(params: Sequence[torch.Tensor]) -> torch.Tensor:
    return ParamNormSumFn.apply(*params)


def teacher_forcing_inputs(targets: torch.Tensor, n_steps: int, pad: int, sos: int):
    """tokens_in (L,B) int64 = (<SOS> row, targets[:L-1]) and ce_weight (L,B) f32 = mask / (max(n_t,1) * sum_t n_t) in one launch
    (train.py:25,44-45,54-60,68)."""
    if not targets.is_cuda or targets.dtype != torch.int64:
        raise RuntimeError(f"targets: expected an int64 CUDA tensor, got {targets.dtype} on {targets.device} (recnet_b200 has no CPU path)")
    targets = targets.contiguous()
    Lmax, B = targets.shape
    if not 1 <= n_steps <= Lmax:
        raise ValueError(f"n_steps={n_steps} outside 1..{Lmax}")
    tokens_in = torch.empty(n_steps, B, dtype=torch.int64, device=targets.device)
    ce_weight = torch.empty(n_steps, B, dtype=torch.float32, device=targets.device)
    L.check(L.lib().recnet_teacher_forcing_prep(targets.data_ptr(), n_steps, B, int(pad), int(sos), tokens_in.data_ptr(),
                                                ce_weight.data_ptr(), _stream()), "recnet_teacher_forcing_prep")
    return tokens_in, ce_weight
