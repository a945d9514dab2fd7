"""HexPlane feature field feeding the S3Gaussian EMD deformation MLP (K1e) -- host-side mirror of
``HexPlaneField`` (``S3Gaussian/scene/hexplane.py:109-187``; constructed at ``deformation.py:41``, queried at
``deformation.py:187-199``).  Same constructor, ``set_aabb`` / ``get_aabb`` and ``forward(pts, timestamps)``.

The reference keeps 24 parameters ``grids.{s}.{p}`` of shape ``[1, F, H, W]`` and evaluates them with 24
``F.grid_sample`` calls.  Here the planes are ONE flat parameter in feature-last order (``[H][W][F]`` per plane,
so a bilinear corner is one 128-byte line) and one C-ABI call gathers all scales and planes
(``emd_hexplane_fwd`` / ``emd_hexplane_bwd``).  ``load_reference_grids`` / ``reference_grids`` convert from / to
the reference's layout (checkpoints load by key through ``load_reference_state_dict``).
"""
from __future__ import annotations

import itertools
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
from torch import Tensor, nn

from . import _C

COMBS = list(itertools.combinations(range(4), 2))


class _HexPlaneFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, planes, pts, t, field, tail=None):
        """``tail[N,E]`` (optional): the result is ``[features | tail]`` of shape [N, S*F+E], the gather writing its columns
        in place (row pitch S*F+E) -- the MLP input of deformation.py:205 without a concatenation pass."""
        L = _C.lib()
        planes_c = planes.detach().float().contiguous()
        pts_c = pts.detach().float().contiguous()
        N = pts_c.shape[0]
        t_c = t.detach().float().contiguous().reshape(-1)
        t_stride = 0 if t_c.numel() == 1 else 1
        if t_stride == 1 and t_c.numel() != N:
            raise ValueError(f"timestamps has {t_c.numel()} values for {N} points")
        S, F = field.num_scales, field.feat_per_plane
        E = 0 if tail is None else int(tail.shape[1])
        if tail is not None and (tail.shape[0] != N or (S * F + E) % 4 != 0):
            raise ValueError(f"tail must be [N, E] with S*F+E a multiple of 4, got {tuple(tail.shape)} for N={N}, S*F={S * F}")
        ld = S * F + E
        feat = torch.empty(N, ld, dtype=torch.float32, device=pts_c.device)
        aabb = field._aabb_host()
        _C.check(L.emd_hexplane_fwd_ld(_C.ptr(planes_c, torch.float32, "planes"), field._offsets_c, field._reso_c, S, F, aabb,
                                       _C.ptr(pts_c, torch.float32, "pts"), _C.ptr(t_c, torch.float32, "timestamps"), t_stride,
                                       N, _C.ptr(feat), ld, _C.stream()), "emd_hexplane_fwd")
        if E:
            feat[:, S * F:] = tail.detach()
        ctx.save_for_backward(planes_c, pts_c, t_c)
        ctx.field, ctx.aabb, ctx.t_stride, ctx.t_shape, ctx.ld = field, aabb, t_stride, t.shape, ld
        return feat

    @staticmethod
    def backward(ctx, v_feat):
        L = _C.lib()
        planes_c, pts_c, t_c = ctx.saved_tensors
        field = ctx.field
        N = pts_c.shape[0]
        dev = pts_c.device
        S, F = field.num_scales, field.feat_per_plane
        v_feat = v_feat.float().contiguous()
        v_planes = torch.zeros_like(planes_c)
        v_pts = torch.empty(N, 3, dtype=torch.float32, device=dev) if ctx.needs_input_grad[1] else None
        v_t = None
        if ctx.needs_input_grad[2]:
            v_t = torch.zeros(1, dtype=torch.float32, device=dev) if ctx.t_stride == 0 else \
                torch.empty(N, dtype=torch.float32, device=dev)
        ws_bytes = L.emd_hexplane_bwd_workspace_bytes(N)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _C.check(L.emd_hexplane_bwd_ld(_C.ptr(planes_c), field._offsets_c, field._reso_c, S, F, ctx.aabb, _C.ptr(pts_c),
                                       _C.ptr(t_c), ctx.t_stride, N, _C.ptr(v_feat), ctx.ld, _C.ptr(v_planes), _C.ptr(v_pts),
                                       _C.ptr(v_t), _C.ptr(ws), ws_bytes, _C.stream()), "emd_hexplane_bwd")
        v_tail = v_feat[:, S * F:] if (ctx.ld > S * F and ctx.needs_input_grad[4]) else None
        return v_planes, v_pts, (v_t.reshape(ctx.t_shape) if v_t is not None else None), None, v_tail


class HexPlaneField(nn.Module):
    """``HexPlaneField(bounds, planeconfig, multires)`` as in hexplane.py:109-146."""

    def __init__(self, bounds, planeconfig: Dict, multires: Sequence[int]) -> None:
        super().__init__()
        aabb = torch.tensor([[bounds, bounds, bounds], [-bounds, -bounds, -bounds]], dtype=torch.float32)
        self.aabb = nn.Parameter(aabb, requires_grad=False)
        self.grid_config = [planeconfig]
        self.multiscale_res_multipliers = list(multires)
        self.concat_features = True
        if planeconfig["grid_dimensions"] != 2 or planeconfig["input_coordinate_dim"] != 4:
            raise NotImplementedError("emd_b200 HexPlaneField: grid_dimensions=2, input_coordinate_dim=4 only "
                                      "(S3Gaussian/arguments/gaussian_options.py:138-143)")
        self.feat_per_plane = int(planeconfig["output_coordinate_dim"])
        self.num_scales = len(self.multiscale_res_multipliers)
        self.feat_dim = self.feat_per_plane * self.num_scales
        base = list(planeconfig["resolution"])
        self.reso: List[List[int]] = [[r * m for r in base[:3]] + base[3:] for m in self.multiscale_res_multipliers]
        offsets, total = [], 0
        for reso in self.reso:
            for (i, j) in COMBS:
                offsets.append(total)
                total += reso[j] * reso[i] * self.feat_per_plane
        self.plane_offsets = offsets
        self.planes = nn.Parameter(torch.empty(total, dtype=torch.float32))
        self._offsets_c = (_C.c_int64 * len(offsets))(*offsets)
        self._reso_c = (_C.c_int * (4 * self.num_scales))(*[r for reso in self.reso for r in reso])
        self._aabb_key, self._aabb_c = None, None
        self.reset_parameters()

    # ---- initialisation / layout conversion -------------------------------------------------------------
    def plane_view(self, s: int, p: int, flat: Optional[Tensor] = None) -> Tensor:
        """[H, W, F] view of plane (s, p) inside ``self.planes`` (or a same-shaped flat tensor, e.g. its grad)."""
        i, j = COMBS[p]
        H, W, F = self.reso[s][j], self.reso[s][i], self.feat_per_plane
        o = self.plane_offsets[s * 6 + p]
        return (self.planes if flat is None else flat)[o:o + H * W * F].view(H, W, F)

    def reset_parameters(self, a: float = 0.1, b: float = 0.5) -> None:
        """init_grid_param (hexplane.py:47-69): time planes 1, space planes U[a, b]."""
        with torch.no_grad():
            for s in range(self.num_scales):
                for p, comb in enumerate(COMBS):
                    v = self.plane_view(s, p)
                    v.fill_(1.0) if 3 in comb else v.uniform_(a, b)

    def load_reference_grids(self, grids: Sequence[Sequence[Tensor]]) -> None:
        """grids[s][p] in the reference layout [1, F, H, W] (``HexPlaneField.grids``)."""
        with torch.no_grad():
            for s in range(self.num_scales):
                for p in range(6):
                    self.plane_view(s, p).copy_(grids[s][p][0].permute(1, 2, 0))

    def reference_grids(self, flat: Optional[Tensor] = None) -> List[List[Tensor]]:
        """The planes (or a same-shaped flat tensor such as ``planes.grad``) in the reference layout [1, F, H, W]."""
        return [[self.plane_view(s, p, flat).permute(2, 0, 1)[None].contiguous() for p in range(6)]
                for s in range(self.num_scales)]

    def load_reference_state_dict(self, state: Dict[str, Tensor], prefix: str = "") -> None:
        """Keys ``{prefix}grids.{s}.{p}`` and ``{prefix}aabb`` of a reference checkpoint."""
        self.load_reference_grids([[state[f"{prefix}grids.{s}.{p}"] for p in range(6)] for s in range(self.num_scales)])
        if f"{prefix}aabb" in state:
            with torch.no_grad():
                self.aabb.copy_(state[f"{prefix}aabb"])

    def reference_state_dict(self, prefix: str = "") -> Dict[str, Tensor]:
        out = {f"{prefix}aabb": self.aabb.detach().clone()}
        for s, row in enumerate(self.reference_grids()):
            for p, g in enumerate(row):
                out[f"{prefix}grids.{s}.{p}"] = g.detach()
        return out

    # ---- reference interface ------------------------------------------------------------------------------
    @property
    def get_aabb(self):
        return self.aabb[0], self.aabb[1]

    def set_aabb(self, xyz_max, xyz_min) -> None:
        aabb = torch.from_numpy(np.array([xyz_max, xyz_min], dtype=np.float32)).to(self.aabb.device)
        self.aabb = nn.Parameter(aabb, requires_grad=False)

    def _aabb_host(self):
        key = (id(self.aabb), self.aabb._version)
        if key != self._aabb_key:
            vals = self.aabb.detach().cpu().reshape(-1).tolist()      # one readback per aabb change, not per step
            self._aabb_c, self._aabb_key = (_C.c_float * 6)(*vals), key
        return self._aabb_c

    def get_density(self, pts: Tensor, timestamps: Optional[Tensor] = None, tail: Optional[Tensor] = None) -> Tensor:
        """``tail[N,E]``: return ``cat([features, tail], -1)`` built in place (see ``_HexPlaneFn.forward``)."""
        if timestamps is None:
            raise ValueError("HexPlaneField needs timestamps (input_coordinate_dim = 4)")
        pts = pts.reshape(-1, pts.shape[-1])
        if timestamps.dim() >= 1 and timestamps.numel() > 1 and timestamps.stride(0) == 0:
            timestamps = timestamps.reshape(-1)[:1]                   # an expanded scalar: one shared time
        if tail is not None and (self.feat_dim + tail.shape[-1]) % 4 != 0:
            return torch.cat([_HexPlaneFn.apply(self.planes, pts, timestamps, self), tail], dim=-1)
        return _HexPlaneFn.apply(self.planes, pts, timestamps, self, tail)

    def forward(self, pts: Tensor, timestamps: Optional[Tensor] = None) -> Tensor:
        return self.get_density(pts, timestamps)
