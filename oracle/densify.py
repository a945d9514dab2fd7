"""CPU oracle: densification statistics of a training step.  TEST INFRASTRUCTURE.

Restates ``BasicTrainer.postprocess_per_train_step`` (``OmniRe/models/trainers/base.py:279-297``: the gradient scaling and
the per-class split) and ``VanillaGaussians.after_train`` (``OmniRe/models/gaussians/vanilla.py:163-191``) with the
reference's own boolean-mask formulation.  Pinned by ``tests/golden/densify.npz`` (the reference's ``after_train`` run on
the CPU, ``tests/golden/make_golden.py --densify``)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor


def scale_grads(grads: Tensor, width: int, height: int, batch_size: int = 1) -> Tensor:
    """base.py:281-286."""
    g = grads.clone()
    g[..., 0] *= width / 2.0 * batch_size
    g[..., 1] *= height / 2.0 * batch_size
    return g


def after_train(state: Dict[str, Optional[Tensor]], radii: Tensor, xys_grad: Tensor, last_size: int) -> None:
    """vanilla.py:163-191 with ``filter_mask`` all-true (``state``: xys_grad_norm, vis_counts, max_2Dsize or None)."""
    visible_mask = (radii > 0).flatten()
    grads = xys_grad.norm(dim=-1)
    if state.get("xys_grad_norm") is None:
        state["xys_grad_norm"] = grads.clone()
        state["vis_counts"] = torch.ones_like(grads)
    else:
        state["vis_counts"][visible_mask] = state["vis_counts"][visible_mask] + 1
        state["xys_grad_norm"][visible_mask] = grads[visible_mask] + state["xys_grad_norm"][visible_mask]
    if state.get("max_2Dsize") is None:
        state["max_2Dsize"] = torch.zeros(radii.numel(), dtype=torch.float32)
    newradii = radii[visible_mask]
    state["max_2Dsize"][visible_mask] = torch.maximum(state["max_2Dsize"][visible_mask], newradii / float(last_size))
