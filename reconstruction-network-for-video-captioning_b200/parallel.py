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
        self._done = False

    def wait_first(self):
        """Interface parity with NvlinkAllReducer: the NCCL group goes out in one piece."""
        self.wait()

    def wait(self):
        """Join the outstanding all-reduces; modules whose grads were not flat views are reduced tensor by tensor."""
        if (self.world <= 1 and not self.force) or getattr(self, "_done", False):
            return
        self._done = True
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


class NvlinkAllReducer:
    """Data-parallel gradient exchange as ONE kernel of our own over NVLink / NVSwitch (csrc/allreduce.cuh) -- no NCCL collective in
    the step.  The flat gradient buffers of all modules are placed in one symmetric-memory allocation (functional.place_flat_grads),
    so backward writes the gradients where every peer (and the switch's multicast alias) can reach them; ``wait()`` launches
    ``recnet_allreduce_avg`` in place on the current stream (a plain kernel node under CUDA-graph capture; its rank barriers count
    epochs on the device, so replays need no reset).

    ``modules`` in the order their gradients become final (reconstructor first).  With ``overlap=True`` the first module's slice is
    reduced on a side stream as soon as its backward has finished (grad hook), hidden behind the decoder's BPTT; the rest goes
    out in ``wait()``.  NCCL is only used at construction time (rendezvous of the symmetric allocations)."""

    MAX_CTAS, MAX_WORLD = 64, 16

    def __init__(self, modules: Sequence[torch.nn.Module], group=None, ctas: int = None, overlap: bool = None):
        import torch.distributed._symmetric_memory as symm
        from . import _lib as L
        from . import functional as Fn
        if not dist.is_initialized():
            raise RuntimeError("NvlinkAllReducer needs an initialised process group (one process per GPU)")
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > self.MAX_WORLD:
            raise NotImplementedError(f"NvlinkAllReducer covers one NVSwitch domain (<= {self.MAX_WORLD} ranks)")
        self.modules = list(modules)
        self.ctas = int(ctas if ctas is not None else os.environ.get("RECNET_AR_CTAS", "16"))
        self.overlap = (os.environ.get("RECNET_DP_OVERLAP", "1") == "1") if overlap is None else bool(overlap)
        # parameter lists in the order the sequence Functions save them (models.*._params): functional._flat_grads keys the
        # placement by exactly that tuple
        plists = []
        for m in self.modules:
            pl = list(m._params()) if hasattr(m, "_params") else [p for p in m.parameters()]
            if {id(p) for p in pl} != {id(p) for p in m.parameters()} or any(not p.requires_grad for p in pl):
                raise NotImplementedError("NvlinkAllReducer: every parameter of a module must take part in its fused backward")
            plists.append(pl)
        self._plists = plists
        dev = plists[0][0].device
        sizes = [(Fn.flat_grad_numel(pl) + 63) // 64 * 64 for pl in plists]
        total = sum(sizes)
        self.buf = symm.empty(total, dtype=torch.float32, device=dev)
        self.buf.zero_()
        self.flags = symm.empty(2 * self.MAX_CTAS * self.MAX_WORLD, dtype=torch.int32, device=dev)
        self.flags.zero_()
        torch.cuda.synchronize()
        name = self.group.group_name
        hb, hf = symm.rendezvous(self.buf, name), symm.rendezvous(self.flags, name)
        self.multicast = int(hb.multicast_ptr) if (hb.has_multicast_support and os.environ.get("RECNET_AR_MULTICAST", "1") == "1") else 0
        self.peer_ptrs = torch.tensor([int(x) for x in hb.buffer_ptrs], dtype=torch.int64, device=dev)
        self.peer_flags = torch.tensor([int(x) for x in hf.buffer_ptrs], dtype=torch.int64, device=dev)
        self._handles = (hb, hf)                      # keep the mappings alive
        self.epochs = torch.zeros(self.MAX_CTAS, dtype=torch.int32, device=dev)     # per-CTA epoch counters, advanced by the kernel
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.ranges, off = [], 0
        for pl, n in zip(plists, sizes):
            Fn.place_flat_grads(pl, self.buf[off: off + n])
            self.ranges.append((off, n))
            off += n
        self.bytes_last = 0
        self._lib, self._check, self._stream = L.lib(), L.check, Fn._stream
        self._side = torch.cuda.Stream(device=dev) if self.overlap and len(self.modules) > 1 else None
        self._early_done = False
        self._rest_launched = False
        self._all_done = False
        self._early_event = torch.cuda.Event()
        self._hooks = []
        if self._side is not None:
            first = plists[0]
            remaining = {"n": 0}

            def hook(_p, first=first, remaining=remaining):
                remaining["n"] += 1
                if remaining["n"] == len(first):      # every gradient of the first module has been accumulated
                    remaining["n"] = 0
                    self._reduce_early()
            for p in first:
                self._hooks.append(p.register_post_accumulate_grad_hook(hook))
        def fresh(_p):                                # a new backward has produced gradients: the next wait() has work to do
            self._all_done = False
        for pl in plists:
            self._hooks.append(pl[0].register_post_accumulate_grad_hook(fresh))
        dist.barrier(self.group)                      # every rank's flag block is zeroed and mapped before anyone signals into it
        torch.cuda.synchronize()

    def _launch(self, off: int, n: int):
        self._check(self._lib.recnet_allreduce_avg(
            self.buf.data_ptr(), self.multicast or None, self.peer_ptrs.data_ptr(), self.flags.data_ptr(), self.peer_flags.data_ptr(),
            self.epochs.data_ptr(), self.err.data_ptr(), off, n, self.rank, self.world, self.ctas, self._stream()),
            "recnet_allreduce_avg")
        self.bytes_last += n * 4

    def _reduce_early(self):
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            off, n = self.ranges[0]
            self._launch(off, n)
            self._early_event.record(self._side)
        self._early_done = True

    def start_iteration(self):
        self.bytes_last = 0
        self._early_done = False
        self._rest_launched = False
        self._all_done = False

    def _assert_placed(self):
        lo = self.buf.data_ptr()
        hi = lo + self.buf.numel() * 4
        for pl in self._plists:
            for p in pl:
                if p.grad is None or not (lo <= p.grad.data_ptr() < hi):
                    raise RuntimeError("NvlinkAllReducer: a gradient does not live in the symmetric buffer (backward did not go through "
                                       "the fused sequence Functions, or the parameter list changed)")

    def wait_first(self):
        """The FIRST module's averaged gradients are ready on the current stream when this returns; the remaining slices start on the
        side stream right away, so the caller can run the first module's optimiser underneath them and ``wait()`` afterwards."""
        self._assert_placed()
        main = torch.cuda.current_stream()
        if self._side is None or len(self.ranges) < 2:
            return self.wait()
        if not self._early_done:
            self._reduce_early()
        rest = self.ranges[1:]
        self._side.wait_stream(main)                  # the other modules' backward has finished (queued after the early kernel on the side stream)
        with torch.cuda.stream(self._side):
            self._launch(rest[0][0], sum(r[1] for r in rest))
        self._rest_launched = True
        # the first slice is final once the early kernel has finished; it precedes the `rest` kernel on the side stream, so wait for an
        # event recorded between them
        main.wait_event(self._early_event)

    def wait(self):
        """Reduce whatever has not gone out yet and join: after this every rank holds the averaged gradients."""
        if self._all_done:
            return
        self._assert_placed()
        self._all_done = True
        main = torch.cuda.current_stream()
        if self._rest_launched:
            main.wait_stream(self._side)
        elif self._early_done:
            rest = self.ranges[1:]
            main.wait_event(self._early_event)        # same flag set: the two kernels must not overlap each other
            self._launch(rest[0][0], sum(r[1] for r in rest))
        else:
            self._launch(0, sum(r[1] for r in self.ranges))
        self._rest_launched = False
        self._early_done = False                      # consumed: the next backward starts over

    def check(self):
        torch.cuda.synchronize()
        if int(self.err.item()) != 0:
            raise RuntimeError("recnet_allreduce_avg: a peer did not reach the rank barrier in time (device status 4)")

    def remove(self):
        from . import functional as Fn
        for h in self._hooks:
            h.remove()
        self._hooks = []
        for pl in self._plists:
            Fn.place_flat_grads(pl, None)


def make_reducer(modules: Sequence[torch.nn.Module], group=None):
    """The gradient reducer of a data-parallel run: our NVLink kernel on CUDA + NCCL process groups (RECNET_DP_IMPL=nccl falls back to
    the fused NCCL group), the NCCL / gloo path otherwise (CPU tests)."""
    impl = os.environ.get("RECNET_DP_IMPL", "nvlink")
    if dist.is_initialized() and dist.get_world_size(group) > 1 and dist.get_backend(group) == "nccl" and impl == "nvlink":
        return NvlinkAllReducer(modules, group)
    return GradAllReducer(modules, group)


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
