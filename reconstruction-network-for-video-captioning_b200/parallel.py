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
        # opt-in: raw NCCL communicator driven on the CURRENT stream through libnccl's C API (no ProcessGroup work objects, no side
        # stream, no fork/join in a captured graph).  Probe (1 x B200, 1-rank group, collective captured in the step graph,
        # RECNET_DP_SELF=1): a captured collective costs 0.1-0.3 ms of graph time even with one rank, and the raw path costs the
        # same as ProcessGroupNCCL -- so the overhead is not torch's stream handling.
        self.raw = os.environ.get("RECNET_DP_RAW", "0") == "1"
        self._raw_comm = None
        # symmetric-memory all-reduce over NVLink peer / NVSwitch multicast memory instead of an NCCL collective:
        # "multimem" (in-switch reduction, multimem.ld_reduce / multimem.st), "two_shot" (P2P reduce-scatter + all-gather), "" = NCCL
        self.symm = os.environ.get("RECNET_DP_SYMM", "")
        self._symm_buf = None
        self._symm_group_name = None
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
        if late and self.symm and self.backend == "nccl" and all(b.dtype == torch.float32 for b in late):
            self._symm_allreduce(late)
        elif late and self.raw and self.backend == "nccl" and os.environ.get("RECNET_DP_DRYRUN") is None:
            self._raw_allreduce(late)
        elif len(late) > 1 and self.backend == "nccl" and os.environ.get("RECNET_DP_DRYRUN") is None:
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

    def _raw_allreduce(self, bufs):
        """ncclAllReduce(float32, AVG) of every buffer on the CURRENT stream through a communicator of our own, driven straight through
        libnccl's C API (ctypes).  The communicator is created on first use, outside graph capture: all ranks must get here together."""
        import ctypes
        if self._raw_comm is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("GradAllReducer: the NCCL communicator must be created before graph capture (run an eager step first)")

            class NcclUniqueId(ctypes.Structure):
                _fields_ = [("internal", ctypes.c_byte * 128)]

            lib = ctypes.CDLL("libnccl.so.2")          # the copy PyTorch already loaded
            lib.ncclGetUniqueId.restype = ctypes.c_int
            lib.ncclCommInitRank.restype = ctypes.c_int
            lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, NcclUniqueId, ctypes.c_int]
            lib.ncclAllReduce.restype = ctypes.c_int
            lib.ncclAllReduce.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p]
            world = max(self.world, 1)
            rank = dist.get_rank(self.group)
            uid = NcclUniqueId()
            if rank == 0 and lib.ncclGetUniqueId(ctypes.byref(uid)) != 0:
                raise RuntimeError("ncclGetUniqueId failed")
            box = [bytes(uid.internal) if rank == 0 else None]
            if world > 1:
                dist.broadcast_object_list(box, src=0, group=self.group)
            ctypes.memmove(ctypes.byref(uid), box[0], 128)
            comm = ctypes.c_void_p()
            rc = lib.ncclCommInitRank(ctypes.byref(comm), world, uid, rank)
            if rc != 0:
                raise RuntimeError(f"ncclCommInitRank failed with {rc}")
            self._raw_lib, self._raw_comm = lib, comm
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        NCCL_FLOAT32, NCCL_AVG = 7, 4                  # ncclDataType_t / ncclRedOp_t
        self._raw_lib.ncclGroupStart()                 # one fused launch for all buffers
        for b in bufs:
            if b.dtype != torch.float32 or not b.is_contiguous():
                raise RuntimeError("GradAllReducer: raw NCCL path expects contiguous float32 gradient buffers")
            self.bytes_last += b.numel() * 4
            rc = self._raw_lib.ncclAllReduce(b.data_ptr(), b.data_ptr(), b.numel(), NCCL_FLOAT32, NCCL_AVG, self._raw_comm, stream)
            if rc != 0:
                self._raw_lib.ncclGroupEnd()
                raise RuntimeError(f"ncclAllReduce failed with {rc}")
        rc = self._raw_lib.ncclGroupEnd()
        if rc != 0:
            raise RuntimeError(f"ncclGroupEnd failed with {rc}")

    def _symm_allreduce(self, bufs):
        """Average `bufs` across ranks through ONE symmetric-memory buffer: pack -> all-reduce kernel over peer memory -> unpack * 1/N.
        The buffer is allocated and rendezvous-ed on first use (must happen outside CUDA-graph capture: run one eager step first)."""
        import torch.distributed._symmetric_memory as sm
        if torch.cuda.is_current_stream_capturing():
            # measured on 2 x B200: correct and 421 us per 99 MB eagerly (tools/dp_check.py), but the captured step graph never
            # completed its first replay (signal-pad barrier vs graph replay); refuse rather than hang
            raise RuntimeError("GradAllReducer: RECNET_DP_SYMM is an eager-mode experiment; it is not usable under CUDA-graph capture")
        group = self.group if self.group is not None else dist.group.WORLD
        total = sum(b.numel() for b in bufs)
        padded = (total + 4095) // 4096 * 4096
        if self._symm_buf is None or self._symm_buf.numel() < padded:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("GradAllReducer: the symmetric buffer must be created before graph capture (run an eager step first)")
            self._symm_buf = sm.empty(padded, dtype=torch.float32, device=bufs[0].device)
            self._symm_buf.zero_()
            sm.rendezvous(self._symm_buf, group.group_name)
            self._symm_group_name = group.group_name
        off = 0
        for b in bufs:
            n = b.numel()
            self._symm_buf[off:off + n].copy_(b.reshape(-1))
            off += n
            self.bytes_last += n * 4
        if self.symm == "two_shot":
            torch.ops.symm_mem.two_shot_all_reduce_(self._symm_buf, "sum", self._symm_group_name)
        elif self.symm == "one_shot":
            self._symm_buf.copy_(torch.ops.symm_mem.one_shot_all_reduce(self._symm_buf, "sum", self._symm_group_name))
        else:
            torch.ops.symm_mem.multimem_all_reduce_(self._symm_buf, "sum", self._symm_group_name)
        off = 0
        for b in bufs:
            n = b.numel()
            torch.mul(self._symm_buf[off:off + n], 1.0 / self.world, out=b.reshape(-1))
            off += n

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
