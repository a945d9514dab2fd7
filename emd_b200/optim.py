"""Fused Adam -- drop-in for the optimizer the reference builds after the render path:
``torch.optim.Adam(groups, lr=0.0, eps=1e-15)`` (``OmniRe/models/trainers/base.py:226``,
``S3Gaussian/scene/gaussian_model.py:200``).  Same constructor, ``param_groups`` and per-parameter state keys
(``step``, ``exp_avg``, ``exp_avg_sq``), so the schedulers that write ``group["lr"]`` (``base.py:436-441``) and the
densification code that slices / replaces ``exp_avg`` and ``exp_avg_sq`` (``OmniRe/models/gaussians/basics.py:196-240``,
``gaussian_model.py:425-480``) work unchanged; ``state_dict`` round-trips with ``torch.optim.Adam``.

``step()`` is one C-ABI call per 32 tensors (``emd_adam_step``) instead of ~6 foreach launches per group.
``grad_scale`` folds the 1/world_size of a summed all-reduce into the update.
"""
from __future__ import annotations

import ctypes
from typing import List

import torch

from . import _C


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False, maximize=False):
        if amsgrad or maximize:
            raise NotImplementedError("emd_b200.FusedAdam: amsgrad / maximize are not used by the reference "
                                      "(base.py:226, gaussian_model.py:200)")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= weight_decay:
            raise ValueError("invalid lr / eps / weight_decay")
        if not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError(f"invalid betas {betas}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False))

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _C.lib()
        cap = L.emd_adam_max_tensors()
        rows: List[tuple] = []
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("emd_b200.FusedAdam does not support sparse gradients")
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise _C.EmdError("emd_b200.FusedAdam: parameters must be contiguous fp32 CUDA tensors")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)      # host scalar, as torch.optim.Adam keeps it
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                rows.append((p, g, st["exp_avg"], st["exp_avg_sq"], float(group["lr"]), float(b1), float(b2),
                             float(group["eps"]), float(group["weight_decay"]), int(st["step"].item())))
        for i in range(0, len(rows), cap):
            chunk = rows[i:i + cap]
            n = len(chunk)
            ptrs = [(_C.P * n)(*[_C.ptr(r[k], torch.float32, ("param", "grad", "exp_avg", "exp_avg_sq")[k]) for r in chunk])
                    for k in range(4)]
            numel = (_C.c_int64 * n)(*[r[0].numel() for r in chunk])
            dbl = [(ctypes.c_double * n)(*[r[k] for r in chunk]) for k in range(4, 9)]
            steps = (_C.c_int64 * n)(*[r[9] for r in chunk])
            _C.check(L.emd_adam_step(*ptrs, numel, *dbl, steps, n, float(grad_scale), _C.stream()), "emd_adam_step")
        return loss
