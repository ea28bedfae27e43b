"""Data-parallel plumbing: one process per GPU, batch sharded, weights replicated, ONE gradient
all-reduce (average) per module per iteration over NCCL / NVLink, overlapped with the rest of backward.

The reference has no distributed code at all.  The path shards cleanly: samples are independent through both
loops, only parameter gradients couple the replicas, so the only collective is the gradient all-reduce.

Overlap: every sequence Function writes its module's gradients into one flat buffer (functional.py).  The
reconstructor's backward finishes before the decoder's BPTT starts, so its buffer (61 MB fp32 for the local
reconstructor) is all-reduced on NCCL's stream WHILE the decoder BPTT kernels run; the decoder's buffer (38 MB)
follows and is waited on just before clip + Adam.
"""
from __future__ import annotations

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
        if self.world > 1:
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
        if self.backend == "nccl":
            self.pending.append(dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        else:                                  # gloo (CPU tests): no AVG
            w = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.pending.append((w, buf))

    def start_iteration(self):
        self._fired, self.pending, self.bytes_last = {}, [], 0

    def wait(self):
        """Join the outstanding all-reduces; modules whose grads were not flat views are reduced tensor by tensor."""
        if self.world <= 1:
            return
        for mi, m in enumerate(self.modules):
            if self._fired.get(mi) is None:
                for p in m.parameters():
                    if p.grad is not None:
                        self._launch(p.grad)
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


def broadcast_parameters(modules: Sequence[torch.nn.Module], src: int = 0, group=None):
    """Replicate rank-`src` weights (all replicas must start identical)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for m in modules:
        for p in m.parameters():
            dist.broadcast(p.data, src=src, group=group)
