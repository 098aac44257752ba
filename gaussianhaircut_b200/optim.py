"""Fused multi-group Adam over the Gaussian parameters ("next" row 2 of SURVEY.md 8f).

Same arithmetic and hyper-parameter surface as the reference's `torch.optim.Adam(l, lr=0.0,
eps=1e-15)` with per-group learning rates (src/scene/gaussian_model.py:431-448); the reference's
NaN guard (src/train_gaussians.py:174-181, seven blocking `.isnan().any()` syncs per step) becomes a
device-side flag.  One kernel launch per step for all groups, through `gh_adam_step` of the C ABI.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List

import torch

from . import _capi


class FusedAdam:
    """Minimal optimiser: `param_groups` like torch (dicts with 'params', 'lr', optional 'name'),
    `.step()`, `.zero_grad()`.  All tensors float32 CUDA contiguous."""

    def __init__(self, param_groups: Iterable[Dict], betas=(0.9, 0.999), eps: float = 1e-15, nan_guard: bool = True):
        self.param_groups: List[Dict] = [dict(g) for g in param_groups]
        self.betas, self.eps = betas, eps
        self.step_count = 0
        self.state: Dict[int, Dict[str, torch.Tensor]] = {}
        flat = [p for g in self.param_groups for p in g["params"]]
        if len(flat) > 8:
            raise ValueError("at most 8 parameter tensors (GH_ADAM_MAX_GROUPS)")
        for p in flat:
            if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                raise ValueError("FusedAdam needs contiguous float32 CUDA parameters")
            self.state[id(p)] = {"exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
        dev = flat[0].device
        self.nan_flag = torch.zeros(1, dtype=torch.int32, device=dev) if nan_guard else None
        self.step_state = torch.zeros(2, dtype=torch.int32, device=dev)   # [steps taken, scratch], device resident

    def zero_grad(self, set_to_none: bool = True):
        for g in self.param_groups:
            for p in g["params"]:
                p.grad = None if set_to_none else (p.grad.zero_() if p.grad is not None else None)

    def step(self, grads: List[torch.Tensor] = None):
        """Update every parameter that has a gradient (`grads` overrides `.grad`, in parameter order)."""
        lib = _capi.load()
        ps, gs, ms, vs, ns, lrs = [], [], [], [], [], []
        it = iter(grads) if grads is not None else None
        for g in self.param_groups:
            for p in g["params"]:
                gr = next(it) if it is not None else p.grad
                if gr is None:
                    continue
                gr = gr.contiguous()
                st = self.state[id(p)]
                ps.append(p); gs.append(gr); ms.append(st["exp_avg"]); vs.append(st["exp_avg_sq"])
                ns.append(p.numel()); lrs.append(float(g["lr"]))
        if not ps:
            return
        self.step_count += 1
        n = len(ps)
        arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])   # noqa: E731
        dev = ps[0].device
        with torch.cuda.device(dev):
            _capi.check(lib.gh_adam_step(
                n, arr(ps), arr(gs), arr(ms), arr(vs), (C.c_ulonglong * n)(*ns), (C.c_float * n)(*lrs),
                float(self.betas[0]), float(self.betas[1]), float(self.eps), int(self.step_count),
                C.c_void_p(self.step_state.data_ptr()),
                C.c_void_p(self.nan_flag.data_ptr()) if self.nan_flag is not None else None,
                C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        self._keep = (ps, gs)   # keep the gradient tensors alive until the kernels have run
