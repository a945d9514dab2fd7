"""Densification statistics: host-side mirror of what ``BasicTrainer.postprocess_per_train_step``
(``OmniRe/models/trainers/base.py:279-297``) feeds ``VanillaGaussians.after_train`` (``vanilla.py:163-191``) -- the
per-Gaussian running sum of screen-space gradient norms, visibility counts and largest screen radius that
``refinement_after`` (``vanilla.py:206``) thresholds.  One kernel launch per step for ALL Gaussian classes and cameras;
the per-class attributes the reference keeps (``xys_grad_norm``, ``vis_counts``, ``max_2Dsize``) are slices of the state."""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from . import _C


class DensifyStats:
    def __init__(self, num_points: int, device, class_slices: Optional[Dict[str, Tuple[int, int]]] = None):
        """``class_slices``: {class name: (first, last+1)} ranges of the concatenated Gaussian list
        (``collect_gaussians`` order, ``base.py:342-383``)."""
        self.n = int(num_points)
        self.state = torch.zeros(3, self.n, dtype=torch.float32, device=device)   # grad-norm sum | counts | max size
        self._synced = torch.zeros_like(self.state)     # the part of `state` every rank already agrees on
        self.first = True
        self.class_slices = dict(class_slices or {})

    xys_grad_norm = property(lambda self: self.state[0])
    vis_counts = property(lambda self: self.state[1])
    max_2Dsize = property(lambda self: self.state[2])

    def of(self, class_name: str) -> Dict[str, Tensor]:
        a, b = self.class_slices[class_name]
        return {"xys_grad_norm": self.state[0, a:b], "vis_counts": self.state[1, a:b], "max_2Dsize": self.state[2, a:b]}

    @torch.no_grad()
    def update(self, info: Dict, absgrad: bool = True, batch_size: int = 1) -> None:
        """``postprocess_per_train_step``: ``info`` is the rasterizer's meta dict after ``backward()``
        (``radii`` [C,N], ``means2d`` with ``.absgrad`` / ``.grad``, ``width``, ``height``)."""
        m2 = info["means2d"]
        grads = m2.absgrad if absgrad else m2.grad
        if grads is None:
            raise RuntimeError("DensifyStats.update: info['means2d'] carries no gradient yet (call after backward(); "
                               "absgrad=True needs rasterization(..., absgrad=True))")
        radii = info["radii"]
        C, N = radii.shape
        assert N == self.n, f"statistics hold {self.n} Gaussians, the render had {N}"
        W, H = int(info["width"]), int(info["height"])
        _C.check(_C.lib().emd_densify_stats(_C.ptr(radii.contiguous(), torch.int32), _C.ptr(grads.contiguous(), torch.float32),
                                            N, C, W / 2.0 * batch_size, H / 2.0 * batch_size, float(max(W, H)),
                                            1 if self.first else 0, _C.ptr(self.state[0]), _C.ptr(self.state[1]),
                                            _C.ptr(self.state[2]), _C.stream()), "emd_densify_stats")
        self.first = False

    @torch.no_grad()
    def sync(self) -> None:
        """Data-parallel replicas see different views: before a refinement step make the statistics agree -- sums of what
        every rank accumulated since the last sync for the gradient norms and counts, the maximum for the screen size
        (the reference is single-process and has no counterpart)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        delta = self.state[:2] - self._synced[:2]
        dist.all_reduce(delta)
        self.state[:2] = self._synced[:2] + delta
        dist.all_reduce(self.state[2], op=dist.ReduceOp.MAX)
        self._synced.copy_(self.state)

    def reset(self) -> None:
        """After a refinement (``vanilla.py:259-261`` sets the three attributes back to None)."""
        self.state.zero_()
        self._synced.zero_()
        self.first = True
