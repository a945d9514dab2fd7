"""Voxel LBS-weight lookup of the SMPL nodes (K1f) -- host-side mirror of ``VoxelDeformer``
(``OmniRe/models/modules.py:459-640``) as ``SMPLTemplate.forward`` queries it (``OmniRe/models/human_body.py:174-179``,
config ``use_voxel_deformer: true``).

The reference keeps ``lbs_voxel_base`` / ``voxel_w_correction`` channel-major ``[B, J, D, H, W]``, adds the two full
volumes every step (``get_voxel_weight``) and calls the 5-D ``F.grid_sample``.  Here both volumes live channel-LAST
(``[B, D, H, W, J]``: one voxel corner = one 96-byte line), the sum is formed only at the corners a point touches, and
the lookup and its VJP are one C-ABI call each (``emd_voxel_lbs_fwd / _bwd``).  ``load_reference_* / reference_*``
convert from / to the reference layout (checkpoint keys ``lbs_voxel_base``, ``voxel_w_correction``, ``offset``,
``scale``).  The init-time KNN diffusion of the SMPL weights into the volume (``_query_weights_smpl``) is set-up code, not
on the per-step path: build the volume with the reference and load it here.  No CPU path.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
from torch import Tensor

from . import _C


def to_channel_last(v: Tensor) -> Tensor:
    """Reference layout [B,J,D,H,W] -> emd_b200 layout [B,D,H,W,J]."""
    return v.permute(0, 2, 3, 4, 1).contiguous()


def to_reference_layout(v: Tensor) -> Tensor:
    """emd_b200 layout [B,D,H,W,J] -> reference layout [B,J,D,H,W]."""
    return v.permute(0, 4, 1, 2, 3).contiguous()


class _VoxelLBS(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xc, corr, base, offset, scale, ratio, ratio_dim):
        L = _C.lib()
        xc = xc.float().contiguous()
        B, V = xc.shape[0], xc.shape[1]
        _, D, H, W, J = base.shape
        out = torch.empty(B, V, J, dtype=torch.float32, device=xc.device)
        _C.check(L.emd_voxel_lbs_fwd(_C.ptr(base, torch.float32, "lbs_voxel_base"), _C.ptr(corr, torch.float32, "voxel_w_correction"),
                                     _C.ptr(offset, torch.float32, "offset"), _C.ptr(scale, torch.float32, "scale"), float(ratio),
                                     int(ratio_dim), B, D, H, W, J, _C.ptr(xc), V, _C.ptr(out), _C.stream()),
                 "emd_voxel_lbs_fwd")
        ctx.save_for_backward(xc, corr, base, offset, scale)
        ctx.cfg = (float(ratio), int(ratio_dim))
        return out

    @staticmethod
    def backward(ctx, v_out):
        L = _C.lib()
        xc, corr, base, offset, scale = ctx.saved_tensors
        ratio, ratio_dim = ctx.cfg
        B, V = xc.shape[0], xc.shape[1]
        _, D, H, W, J = base.shape
        need_xc, need_corr = ctx.needs_input_grad[0], corr is not None and ctx.needs_input_grad[1]
        v_xc = torch.empty_like(xc) if need_xc else None
        v_corr = torch.zeros_like(corr) if need_corr else None   # the scatter ADDS into it
        _C.check(L.emd_voxel_lbs_bwd(_C.ptr(base), _C.ptr(corr), _C.ptr(offset), _C.ptr(scale), ratio, ratio_dim, B, D, H, W, J,
                                     _C.ptr(xc), V, _C.ptr(v_out.float().contiguous()), _C.ptr(v_corr), _C.ptr(v_xc), _C.stream()),
                 "emd_voxel_lbs_bwd")
        return v_xc, v_corr, None, None, None, None, None


class VoxelDeformer:
    """Tensors ``VoxelDeformer`` owns (same names and meaning, volumes channel-last) + the fused lookup.

    ``lbs_voxel_base`` / ``voxel_w_correction``: ``[B, D, H, W, J]``; ``offset[B,1,3]``, ``scale[B,1,1]``; ``ratio`` and
    ``ratio_dim`` as the reference stores them (``resolution[long] / resolution[short]`` and ``-1 - short_dim_dhw``)."""

    def __init__(self, lbs_voxel_base_ref: Tensor, offset: Tensor, scale: Tensor, resolution_dhw: Sequence[int],
                 short_dim_dhw: int = 0, long_dim_dhw: int = 1, voxel_w_correction_ref: Optional[Tensor] = None):
        assert tuple(lbs_voxel_base_ref.shape[2:]) == tuple(resolution_dhw), "lbs_voxel_base must be [B, J, D, H, W]"
        assert lbs_voxel_base_ref.shape[1] % 4 == 0 and lbs_voxel_base_ref.shape[1] <= 32, "J must be a multiple of 4, <= 32"
        self.resolution_dhw = list(resolution_dhw)
        self.ratio = float(resolution_dhw[long_dim_dhw]) / float(resolution_dhw[short_dim_dhw])   # modules.py:481-486
        self.ratio_dim = -1 - short_dim_dhw
        self.lbs_voxel_base = to_channel_last(lbs_voxel_base_ref.float())
        self.offset = offset.float().reshape(-1, 1, 3).contiguous()
        self.scale = scale.float().reshape(-1, 1, 1).contiguous()
        self.num_bones = lbs_voxel_base_ref.shape[1]
        self.voxel_w_correction: Optional[Tensor] = None
        if voxel_w_correction_ref is not None:
            self.voxel_w_correction = to_channel_last(voxel_w_correction_ref.float()).requires_grad_(True)

    # -- reference API -------------------------------------------------------------------------------------------
    def enable_voxel_correction(self):
        """modules.py:559-561 (zero-initialised trainable correction)."""
        self.voxel_w_correction = torch.zeros_like(self.lbs_voxel_base).requires_grad_(True)

    @property
    def get_voxel_weight(self) -> Tensor:
        """base + correction in the REFERENCE layout [B,J,D,H,W] (modules.py:575-582) -- for export / inspection; the
        lookup below never materialises it."""
        w = self.lbs_voxel_base if self.voxel_w_correction is None else self.lbs_voxel_base + self.voxel_w_correction
        return to_reference_layout(w)

    def normalize(self, x: Tensor) -> Tensor:
        """modules.py:627-632."""
        xn = (x - self.offset) / self.scale
        mul = torch.ones(3, dtype=x.dtype, device=x.device)
        mul[self.ratio_dim] = self.ratio
        return xn * mul

    def denormalize(self, x: Tensor) -> Tensor:
        """modules.py:634-639."""
        mul = torch.ones(3, dtype=x.dtype, device=x.device)
        mul[self.ratio_dim] = self.ratio
        return x / mul * self.scale + self.offset

    def forward(self, xc: Tensor) -> Tensor:
        """``xc[B, N, 3]`` canonical points -> skinning weights ``[B, N, J]`` (modules.py:612-625)."""
        assert xc.dim() == 3 and xc.shape[0] == self.lbs_voxel_base.shape[0] and xc.shape[-1] == 3
        return _VoxelLBS.apply(xc, self.voxel_w_correction, self.lbs_voxel_base, self.offset.reshape(-1, 3),
                               self.scale.reshape(-1), self.ratio, self.ratio_dim % 3)

    __call__ = forward

    # regularisers of smpl.py:639-650 (every few steps, outside the render path): plain tensor expressions on the
    # channel-last parameter -- dims (D, H, W) are 1, 2, 3 here, the channel norm is over the last dim
    def get_tv(self, name: str = "dc") -> Tensor:
        d = self.voxel_w_correction
        if name != "dc" or d is None:
            return torch.zeros((), device=self.lbs_voxel_base.device)
        tv_x = torch.abs(d[:, 1:] - d[:, :-1]).mean()
        tv_y = torch.abs(d[:, :, 1:] - d[:, :, :-1]).mean()
        tv_z = torch.abs(d[:, :, :, 1:] - d[:, :, :, :-1]).mean()
        return (tv_x + tv_y + tv_z) / 3.0

    def get_mag(self, name: str = "dc") -> Tensor:
        d = self.voxel_w_correction
        if name != "dc" or d is None:
            return torch.zeros((), device=self.lbs_voxel_base.device)
        return torch.norm(d, dim=-1).mean()

    # -- checkpoints in the reference layout ----------------------------------------------------------------------
    def reference_state(self):
        out = {"lbs_voxel_base": to_reference_layout(self.lbs_voxel_base), "offset": self.offset, "scale": self.scale}
        if self.voxel_w_correction is not None:
            out["voxel_w_correction"] = to_reference_layout(self.voxel_w_correction.detach())
        return out

    def load_reference_state(self, sd):
        self.lbs_voxel_base = to_channel_last(sd["lbs_voxel_base"].float())
        self.offset, self.scale = sd["offset"].float().reshape(-1, 1, 3), sd["scale"].float().reshape(-1, 1, 1)
        if "voxel_w_correction" in sd:
            self.voxel_w_correction = to_channel_last(sd["voxel_w_correction"].float()).requires_grad_(True)
