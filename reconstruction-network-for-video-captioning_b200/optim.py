"""Fused gradient clip + Adam: the tail of the reference iteration (train.py:269-273) as three kernels of our own.

``ClipAdam`` is a ``torch.optim.Optimizer`` with ``torch.optim.Adam``'s constructor arguments, update rule (L2 weight
decay added to the gradient, optional amsgrad) and ``state_dict`` layout -- a reference checkpoint's optimiser state
(train.py:404-420) loads into it and its own state loads into ``torch.optim.Adam`` -- plus ``max_grad_norm``: when set,
``torch.nn.utils.clip_grad_norm_(params, max_grad_norm)`` (train.py:269-270) is folded into the same pass (the total norm
is reduced on the device and the coefficient applied while the gradients are read for the update).

State lives in flat fp32 buffers laid out like the flat gradient buffer the sequence Functions write
(functional._flat_grads), so the kernels take four address tables and stream p, g, m, v (, vmax) exactly once.
The step counter is a device scalar advanced by the kernel: CUDA-graph replays keep counting.  No CPU path.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import _lib as L
from . import functional as Fn


class ClipAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False,
                 max_grad_norm: Optional[float] = None, write_clipped_grads: bool = True):
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"ClipAdam: invalid hyper-parameters lr={lr} betas={betas} eps={eps} weight_decay={weight_decay}")
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=bool(amsgrad),
                        max_grad_norm=max_grad_norm, write_clipped_grads=bool(write_clipped_grads))
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise NotImplementedError("ClipAdam handles one parameter group (the reference builds one optimiser per module, "
                                      "train.py:149,186)")
        self._ready = False
        self._gptr_cache: Dict[tuple, torch.Tensor] = {}

    # ---- lazily built flat state --------------------------------------------------------------------------------
    def _params(self) -> List[torch.Tensor]:
        return [p for p in self.param_groups[0]["params"] if p.requires_grad]

    def _build(self):
        params = self._params()
        if not params:
            raise ValueError("ClipAdam: no parameters require gradients")
        for p in params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError(f"ClipAdam: parameters must be contiguous float32 CUDA tensors, got {p.dtype} on {p.device} "
                                   "(recnet_b200 has no CPU path)")
        dev = params[0].device
        tab = Fn._table_for(params)
        self._tab = tab
        amsgrad = self.param_groups[0]["amsgrad"]
        self._flat = {k: torch.zeros(tab.total, dtype=torch.float32, device=dev)
                      for k in (("exp_avg", "exp_avg_sq") + (("max_exp_avg_sq",) if amsgrad else ()))}
        self._ptrs = {k: tab.offset_bytes + v.data_ptr() for k, v in self._flat.items()}
        self._dev_state = torch.zeros(8, dtype=torch.float32, device=dev)
        self._partial = torch.empty(tab.n_blocks, dtype=torch.float32, device=dev)
        for o, p in zip(tab.offsets, params):
            st = self.state[p]
            st["step"] = self._dev_state[0]                                   # shared device scalar (view)
            for k, v in self._flat.items():
                st[k] = v[o: o + p.numel()].view_as(p)
        self._ready = True

    def _grad_table(self, params) -> torch.Tensor:
        grads = [p.grad for p in params]
        if any(g is None for g in grads):
            missing = sum(g is None for g in grads)
            raise NotImplementedError(f"ClipAdam: {missing} parameter(s) have no gradient; every parameter of a module takes part "
                                      "in the reference's iteration (the norm regulariser touches all of them)")
        for g in grads:
            if not g.is_cuda or g.dtype != torch.float32 or not g.is_contiguous():
                raise RuntimeError("ClipAdam: gradients must be contiguous float32 CUDA tensors")
        # Address table = per-gradient byte offsets from the lowest address (built on the host once per layout) + that address
        # (a device-side add, legal under CUDA-graph capture).  The sequence Functions hand autograd views of ONE flat buffer
        # with a fixed layout, so after the first eager step every later step -- captured or not -- reuses the offset table.
        # (Addresses rather than ``g._base`` identity: the flat buffer's Python object is gone once backward has returned.)
        ptrs = [g.data_ptr() for g in grads]
        b0 = min(ptrs)
        rel = tuple(q - b0 for q in ptrs)
        t = self._gptr_cache.get((b0, rel))
        if t is None:
            rel_t = self._gptr_cache.get(("rel", rel))
            if rel_t is None:
                self._no_capture("the gradient address table (layout not seen in an eager step)")
                if len(self._gptr_cache) > 64:
                    self._gptr_cache.clear()
                rel_t = self._gptr_cache[("rel", rel)] = torch.tensor(rel, dtype=torch.int64, device=grads[0].device)
            if len(self._gptr_cache) > 64:
                self._gptr_cache = {("rel", rel): rel_t}
            t = self._gptr_cache[(b0, rel)] = rel_t + b0
        return t

    @staticmethod
    def _no_capture(what: str):
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError(f"ClipAdam: {what} is built on the host, which cannot happen during CUDA-graph capture "
                               "(run one eager step first)")

    # ---- public API -----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if not self._ready:
            self._build()
        g = self.param_groups[0]
        params = self._params()
        tab = self._tab
        gptrs = self._grad_table(params)
        x = self._ptrs.get("max_exp_avg_sq")
        mgn = g.get("max_grad_norm")
        L.check(L.lib().recnet_adam_step(
            tab.ptrs.data_ptr(), gptrs.data_ptr(), self._ptrs["exp_avg"].data_ptr(), self._ptrs["exp_avg_sq"].data_ptr(),
            None if x is None else x.data_ptr(), tab.sizes.data_ptr(), tab.n, tab.blk_tensor.data_ptr(), tab.blk_chunk.data_ptr(),
            tab.n_blocks, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]),
            float(mgn) if mgn else 0.0, self._partial.data_ptr(), self._dev_state.data_ptr(),
            int(bool(g.get("write_clipped_grads", True))), Fn._stream()), "recnet_adam_step")
        return loss

    @property
    def last_grad_norm(self) -> torch.Tensor:
        """Total gradient norm seen by the last step (device scalar; 0 when max_grad_norm is unset)."""
        if not self._ready:
            self._build()
        return self._dev_state[1]

    def state_dict(self):
        """torch.optim.Adam's layout.  Every parameter gets its OWN ``step`` tensor (a copy of the shared device counter): the
        live state aliases one scalar, and an aliased ``step`` loaded into torch.optim.Adam would be advanced once per
        parameter per iteration."""
        if not self._ready and any(p.requires_grad for p in self.param_groups[0]["params"]):
            self._build()
        sd = super().state_dict()
        for st in sd["state"].values():
            if "step" in st and torch.is_tensor(st["step"]):
                st["step"] = st["step"].detach().clone()
            for k in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"):
                if k in st:
                    st[k] = st[k].detach().clone()      # independent storage: not views of the flat buffers
        return sd

    def load_state_dict(self, state_dict):
        """Accepts torch.optim.Adam's layout; values are copied INTO the flat buffers (the views stay bound)."""
        if not self._ready:
            self._build()
        groups = state_dict["param_groups"]
        if len(groups) != 1:
            raise NotImplementedError("ClipAdam.load_state_dict: one parameter group expected")
        params = self.param_groups[0]["params"]
        ids = groups[0]["params"]
        if len(ids) != len(params):
            raise ValueError("ClipAdam.load_state_dict: parameter count mismatch")
        step = None
        for pid, p in zip(ids, params):
            st = state_dict["state"].get(pid)
            if st is None or not p.requires_grad:
                continue
            mine = self.state[p]
            for k in self._flat:
                if k in st:
                    mine[k].copy_(st[k].to(mine[k].device, torch.float32))
            if "step" in st:
                s = float(st["step"])
                step = s if step is None else max(step, s)
        if step is not None:
            self._dev_state[0] = step
        for k, v in groups[0].items():
            if k != "params" and k in self.param_groups[0]:
                self.param_groups[0][k] = tuple(v) if k == "betas" else v
