"""Data-parallel plumbing: one process per GPU, batch sharded, weights replicated, ONE gradient
all-reduce (average) per iteration over NCCL / NVLink.

The reference has no distributed code at all.  The path shards cleanly: samples are independent through both
loops, only parameter gradients couple the replicas, so the only collective is the gradient all-reduce.

Every sequence Function writes its module's gradients into one flat buffer (functional.py); after backward the gradients
(61 MB local reconstructor + 38 MB decoder, fp32) go out as ONE fused NCCL group (ncclGroupStart/End), i.e. one rank
synchronisation per step, then clip + Adam.  autograd stores the views detached, so the group carries one all-reduce per
parameter tensor unless RECNET_DP_FLAT=1 recovers the two flat buffers (functional.flat_buffer_of).  Measured on 2 x B200 (profiles/r1_d_dp.md): fused, after
backward 4.67 ms/step; per-module all-reduces launched from grad hooks to overlap the decoder BPTT 5.00 ms/step --
the NCCL CTAs slow the latency-bound BPTT chain as much as they hide, and a second collective is a second rank
sync -- so overlap is opt-in (RECNET_DP_OVERLAP=1).
"""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
import torch.distributed as dist


class GradAllReducer:
    """Launches an async all-reduce of a module's flat gradient buffer as soon as that module's backward has
    produced it (post-accumulate-grad hook on its parameters); ``wait()`` joins before the optimiser."""

    def __init__(self, modules: Sequence[torch.nn.Module], group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.backend = dist.get_backend(group) if dist.is_initialized() else None
        self.pending: List = []
        self._fired = {}
        self._handles = []
        self.modules = list(modules)
        self.bytes_last = 0
        self.overlap = os.environ.get("RECNET_DP_OVERLAP", "0") == "1"
        # autograd stores gradient views detached (``p.grad._base`` is None), so by default each module's gradients go out as one
        # all-reduce per parameter tensor inside ONE NCCL group (21 tensors for decoder + local reconstructor).  RECNET_DP_FLAT=1
        # looks the contiguous buffer up in functional.flat_buffer_of() and sends one all-reduce per module (2 in the group);
        # CPU-tested over gloo, not yet measured on NVLink -> opt-in.
        self.flat_lookup = os.environ.get("RECNET_DP_FLAT", "0") == "1"
        self.force = os.environ.get("RECNET_DP_SELF") == "1" and dist.is_initialized()    # probe: run the collective even with one rank
        if self.world > 1 and self.overlap:
            for mi, m in enumerate(self.modules):
                params = [p for p in m.parameters() if p.requires_grad]
                for p in params:
                    self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(mi, params)))

    def _make_hook(self, mi, params):
        def hook(p):
            if self._fired.get(mi):
                return
            base = p.grad._base if p.grad is not None else None
            if base is None:
                return                        # not a flat-buffer view: reduced per tensor in wait()
            self._fired[mi] = base
            self._launch(base)
        return hook

    def _launch(self, buf: torch.Tensor):
        self.bytes_last += buf.numel() * buf.element_size()
        if os.environ.get("RECNET_DP_DRYRUN") == "1":      # developer probe: everything but the collective itself
            return
        if os.environ.get("RECNET_DP_DRYRUN") == "2":      # developer probe: synchronise the ranks, move no data
            buf = buf.view(-1)[:1]
        if self.backend == "nccl":
            self.pending.append(dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        else:                                  # gloo (CPU tests): no AVG
            w = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.pending.append((w, buf))

    def start_iteration(self):
        self._fired, self.pending, self.bytes_last = {}, [], 0

    def wait(self):
        """Join the outstanding all-reduces; modules whose grads were not flat views are reduced tensor by tensor."""
        if self.world <= 1 and not self.force:
            return
        late = []
        for mi, m in enumerate(self.modules):
            if self._fired.get(mi) is None:
                grads = [p.grad for p in m.parameters() if p.grad is not None]
                flat = _common_base(grads)
                if flat is None and self.flat_lookup:
                    from .functional import flat_buffer_of
                    flat = flat_buffer_of(grads)
                if flat is not None:
                    late.append(flat)                                # ONE flat buffer per module
                else:
                    late.extend(grads)
        if len(late) > 1 and self.backend == "nccl" and os.environ.get("RECNET_DP_DRYRUN") is None:
            # one NCCL group (ncclGroupStart/End): all buffers in a single fused collective launch -> one rank sync per step
            with dist._coalescing_manager(group=self.group, device=late[0].device, async_ops=True) as cm:
                for buf in late:
                    self.bytes_last += buf.numel() * buf.element_size()
                    dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group)
            self.pending.append(cm)
        else:
            for buf in late:
                self._launch(buf)
        for w in self.pending:
            if isinstance(w, tuple):
                w[0].wait()
                w[1].div_(self.world)
            else:
                w.wait()
        self.pending = []

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []


def _common_base(grads):
    """The tensor all ``grads`` are views of, or None (compared by storage address and size).  Gradients that went through
    autograd's AccumulateGrad are stored detached, so this only recognises views that were assigned to ``.grad`` by hand."""
    if not grads:
        return None
    bases = [g._base for g in grads]
    if any(b is None for b in bases):
        return None
    b0 = bases[0]
    key = (b0.data_ptr(), b0.numel(), b0.dtype)
    if any((b.data_ptr(), b.numel(), b.dtype) != key for b in bases[1:]):
        return None
    return b0 if b0.is_contiguous() else None


def broadcast_parameters(modules: Sequence[torch.nn.Module], src: int = 0, group=None):
    """Replicate rank-`src` weights (all replicas must start identical)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for m in modules:
        for p in m.parameters():
            dist.broadcast(p.data, src=src, group=group)
