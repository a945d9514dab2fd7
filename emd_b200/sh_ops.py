"""Spherical-harmonics colour and the fused node activation (K1b).

``spherical_harmonics`` mirrors ``gsplat.cuda._wrapper.spherical_harmonics`` as
the reference imports it (``OmniRe/models/gaussians/basics.py:16``; calls at
``vanilla.py:388``, ``rigid.py:584``, ``smpl.py:555``).  ``activate_gaussians``
is the single-pass equivalent of the tail of ``get_gaussians``
(``vanilla.py:378-414``, ``rigid.py:578-603``, ``smpl.py:549-576``).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
from torch import Tensor

from . import _C

_c = ctypes


class _SphericalHarmonics(torch.autograd.Function):
    @staticmethod
    def forward(ctx, degree: int, dirs: Tensor, coeffs: Tensor):
        L = _C.lib()
        dirs = dirs.float().contiguous()
        coeffs = coeffs.float().contiguous()
        N, K = coeffs.shape[0], coeffs.shape[1]
        out = torch.empty(N, 3, dtype=torch.float32, device=coeffs.device)
        _C.check(L.emd_sh_fwd(degree, _C.ptr(dirs, torch.float32, "dirs"), _C.ptr(coeffs, torch.float32, "coeffs"),
                              N, K, _C.ptr(out), _C.stream()), "emd_sh_fwd")
        ctx.save_for_backward(dirs)
        ctx.cfg = (degree, N, K)
        return out

    @staticmethod
    def backward(ctx, v_out):
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("emd_b200.spherical_harmonics: gradient w.r.t. dirs is not implemented "
                                      "(the reference passes detached view directions)")
        L = _C.lib()
        (dirs,) = ctx.saved_tensors
        degree, N, K = ctx.cfg
        v_out = v_out.float().contiguous()
        v_coeffs = torch.empty(N, K, 3, dtype=torch.float32, device=dirs.device)
        _C.check(L.emd_sh_bwd(degree, _C.ptr(dirs), N, K, _C.ptr(v_out), _C.ptr(v_coeffs), _C.stream()), "emd_sh_bwd")
        return None, None, v_coeffs


def spherical_harmonics(degrees_to_use: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor] = None) -> Tensor:
    """``dirs[...,3]`` (normalised inside), ``coeffs[...,K,3]`` -> ``[...,3]``."""
    if masks is not None:
        raise NotImplementedError("emd_b200.spherical_harmonics: masks= is not supported")
    assert coeffs.shape[-1] == 3 and dirs.shape[-1] == 3
    assert (degrees_to_use + 1) ** 2 <= coeffs.shape[-2], "coeffs K too small for degree"
    lead = coeffs.shape[:-2]
    out = _SphericalHarmonics.apply(int(degrees_to_use), dirs.reshape(-1, 3), coeffs.reshape(-1, coeffs.shape[-2], 3))
    return out.reshape(lead + (3,))


class _Activate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means_world, dc, rest, opac_logit, log_scales, quats, point_ids, inst_valid, cam_pos, degree):
        L = _C.lib()
        N = dc.shape[0]
        dev = dc.device
        K = 1 + (rest.shape[1] if rest is not None and rest.numel() > 0 else 0)
        means_world = means_world.detach().float().contiguous()
        dc, quats = dc.float().contiguous(), quats.float().contiguous()
        rest = rest.float().contiguous() if K > 1 else None
        opac_logit = opac_logit.float().contiguous().reshape(-1)
        log_scales = log_scales.float().contiguous()
        C = len(cam_pos) // 3
        cam = (_c.c_float * (3 * C))(*[float(v) for v in cam_pos])
        rgbs = torch.empty(C, N, 3, dtype=torch.float32, device=dev)
        opac = torch.empty(N, dtype=torch.float32, device=dev)
        scales = torch.empty(N, 3, dtype=torch.float32, device=dev)
        quats_n = torch.empty(N, 4, dtype=torch.float32, device=dev)
        clamp_pass = torch.empty(C, N, dtype=torch.uint8, device=dev)
        pid = point_ids.contiguous() if point_ids is not None else None
        iv = inst_valid.to(torch.uint8).contiguous() if inst_valid is not None else None
        _C.check(L.emd_activate_fwd(_C.ptr(means_world), _C.ptr(dc), _C.ptr(rest), _C.ptr(opac_logit),
                                    _C.ptr(log_scales), _C.ptr(quats), _C.ptr(pid, torch.int64, "point_ids"),
                                    _C.ptr(iv), cam, C, N, K, degree, _C.ptr(rgbs), _C.ptr(opac), _C.ptr(scales),
                                    _C.ptr(quats_n), _C.ptr(clamp_pass), _C.stream()), "emd_activate_fwd")
        ctx.save_for_backward(means_world, dc, rest if rest is not None else torch.empty(0, device=dev), opac_logit,
                              log_scales, quats, pid if pid is not None else torch.empty(0, device=dev),
                              iv if iv is not None else torch.empty(0, device=dev), clamp_pass, scales)
        ctx.cfg = (N, K, degree, tuple(float(v) for v in cam_pos), pid is not None, C)
        return rgbs, opac, scales, quats_n

    @staticmethod
    def backward(ctx, v_rgbs, v_opac, v_scales, v_quats_n):
        L = _C.lib()
        means_world, dc, rest, opac_logit, log_scales, quats, pid, iv, clamp_pass, scales = ctx.saved_tensors
        N, K, degree, cam_pos, has_ids, C = ctx.cfg
        dev = dc.device
        cam = (_c.c_float * (3 * C))(*cam_pos)

        def z(g, shape):
            return g.float().contiguous() if g is not None else torch.zeros(shape, dtype=torch.float32, device=dev)

        v_rgbs, v_opac = z(v_rgbs, (C, N, 3)), z(v_opac, (N,))
        v_scales, v_quats_n = z(v_scales, (N, 3)), z(v_quats_n, (N, 4))
        v_dc = torch.empty(N, 3, dtype=torch.float32, device=dev)
        v_rest = torch.empty(N, K - 1, 3, dtype=torch.float32, device=dev) if K > 1 else None
        v_logit = torch.empty(N, dtype=torch.float32, device=dev)
        v_ls = torch.empty(N, 3, dtype=torch.float32, device=dev)
        v_q = torch.empty(N, 4, dtype=torch.float32, device=dev)
        _C.check(L.emd_activate_bwd(_C.ptr(means_world), _C.ptr(dc), _C.ptr(rest) if K > 1 else None,
                                    _C.ptr(opac_logit), _C.ptr(log_scales), _C.ptr(quats),
                                    _C.ptr(pid) if has_ids else None, _C.ptr(iv) if has_ids else None, cam, C, N, K,
                                    degree, _C.ptr(clamp_pass), _C.ptr(scales), _C.ptr(v_rgbs), _C.ptr(v_opac),
                                    _C.ptr(v_scales), _C.ptr(v_quats_n), _C.ptr(v_dc), _C.ptr(v_rest), _C.ptr(v_logit),
                                    _C.ptr(v_ls), _C.ptr(v_q), _C.stream()), "emd_activate_bwd")
        return None, v_dc, v_rest, v_logit, v_ls, v_q, None, None, None, None


class _ActColors(torch.autograd.Function):
    """Colour half of the node activation (``emd_activate_fwd/bwd`` with the geometry group null)."""

    @staticmethod
    def forward(ctx, means_world, dc, rest, cam_pos, degree):
        L = _C.lib()
        N = dc.shape[0]
        dev = dc.device
        K = 1 + (rest.shape[1] if rest is not None and rest.numel() > 0 else 0)
        means_world = means_world.detach().float().contiguous()
        dc = dc.float().contiguous()
        rest = rest.float().contiguous() if K > 1 else None
        C = len(cam_pos) // 3
        cam = (_c.c_float * (3 * C))(*[float(v) for v in cam_pos])
        rgbs = torch.empty(C, N, 3, dtype=torch.float32, device=dev)
        clamp_pass = torch.empty(C, N, dtype=torch.uint8, device=dev)
        _C.check(L.emd_activate_fwd(_C.ptr(means_world), _C.ptr(dc), _C.ptr(rest), None, None, None, None, None, cam, C,
                                    N, K, degree, _C.ptr(rgbs), None, None, None, _C.ptr(clamp_pass), _C.stream()),
                 "emd_activate_fwd[colours]")
        ctx.save_for_backward(means_world, dc, rest if rest is not None else torch.empty(0, device=dev), clamp_pass)
        ctx.cfg = (N, K, degree, tuple(float(v) for v in cam_pos), C)
        return rgbs

    @staticmethod
    def backward(ctx, v_rgbs):
        L = _C.lib()
        means_world, dc, rest, clamp_pass = ctx.saved_tensors
        N, K, degree, cam_pos, C = ctx.cfg
        dev = dc.device
        cam = (_c.c_float * (3 * C))(*cam_pos)
        v_rgbs = v_rgbs.float().contiguous()
        v_dc = torch.empty(N, 3, dtype=torch.float32, device=dev)
        v_rest = torch.empty(N, K - 1, 3, dtype=torch.float32, device=dev) if K > 1 else None
        _C.check(L.emd_activate_bwd(_C.ptr(means_world), _C.ptr(dc), _C.ptr(rest) if K > 1 else None, None, None, None,
                                    None, None, cam, C, N, K, degree, _C.ptr(clamp_pass), None, _C.ptr(v_rgbs), None,
                                    None, None, _C.ptr(v_dc), _C.ptr(v_rest), None, None, None, _C.stream()),
                 "emd_activate_bwd[colours]")
        return None, v_dc, v_rest, None, None


class _ActGeom(torch.autograd.Function):
    """Geometry half of the node activation (``emd_activate_fwd/bwd`` with the colour group null)."""

    @staticmethod
    def forward(ctx, opac_logit, log_scales, quats, point_ids, inst_valid):
        L = _C.lib()
        N = quats.shape[0]
        dev = quats.device
        quats = quats.float().contiguous()
        opac_logit = opac_logit.float().contiguous().reshape(-1)
        log_scales = log_scales.float().contiguous()
        opac = torch.empty(N, dtype=torch.float32, device=dev)
        scales = torch.empty(N, 3, dtype=torch.float32, device=dev)
        quats_n = torch.empty(N, 4, dtype=torch.float32, device=dev)
        pid = point_ids.contiguous() if point_ids is not None else None
        iv = inst_valid.to(torch.uint8).contiguous() if inst_valid is not None else None
        cam = (_c.c_float * 3)(0.0, 0.0, 0.0)
        _C.check(L.emd_activate_fwd(None, None, None, _C.ptr(opac_logit), _C.ptr(log_scales), _C.ptr(quats),
                                    _C.ptr(pid, torch.int64, "point_ids"), _C.ptr(iv), cam, 1, N, 1, 0, None,
                                    _C.ptr(opac), _C.ptr(scales), _C.ptr(quats_n), None, _C.stream()),
                 "emd_activate_fwd[geometry]")
        e = torch.empty(0, device=dev)
        ctx.save_for_backward(opac_logit, log_scales, quats, pid if pid is not None else e, iv if iv is not None else e,
                              scales)
        ctx.cfg = (N, pid is not None)
        return opac, scales, quats_n

    @staticmethod
    def backward(ctx, v_opac, v_scales, v_quats_n):
        L = _C.lib()
        opac_logit, log_scales, quats, pid, iv, scales = ctx.saved_tensors
        N, has_ids = ctx.cfg
        dev = quats.device

        def z(g, shape):
            return g.float().contiguous() if g is not None else torch.zeros(shape, dtype=torch.float32, device=dev)

        v_opac, v_scales, v_quats_n = z(v_opac, (N,)), z(v_scales, (N, 3)), z(v_quats_n, (N, 4))
        v_logit = torch.empty(N, dtype=torch.float32, device=dev)
        v_ls = torch.empty(N, 3, dtype=torch.float32, device=dev)
        v_q = torch.empty(N, 4, dtype=torch.float32, device=dev)
        cam = (_c.c_float * 3)(0.0, 0.0, 0.0)
        _C.check(L.emd_activate_bwd(None, None, None, _C.ptr(opac_logit), _C.ptr(log_scales), _C.ptr(quats),
                                    _C.ptr(pid) if has_ids else None, _C.ptr(iv) if has_ids else None, cam, 1, N, 1, 0,
                                    None, _C.ptr(scales), None, _C.ptr(v_opac), _C.ptr(v_scales), _C.ptr(v_quats_n),
                                    None, None, _C.ptr(v_logit), _C.ptr(v_ls), _C.ptr(v_q), _C.stream()),
                 "emd_activate_bwd[geometry]")
        return v_logit, v_ls, v_q, None, None


def _flat_cams(cam_pos):
    if isinstance(cam_pos, Tensor):
        cam_pos = cam_pos.detach().cpu().tolist()
    multi = len(cam_pos) > 0 and isinstance(cam_pos[0], (list, tuple))
    flat = [float(v) for cp in cam_pos for v in cp] if multi else [float(v) for v in cam_pos]
    assert len(flat) % 3 == 0 and 3 <= len(flat) <= 24, "cam_pos: 1..8 camera centres"
    return flat, multi


def sh_colors(means_world: Tensor, features_dc: Tensor, features_rest: Optional[Tensor], cam_pos,
              sh_degree_to_use: int) -> Tensor:
    """Colour half of ``get_gaussians`` alone: ``clamp(SH(normalize(x - cam), [dc, rest]) + 0.5, 0, 1)`` ->
    ``[N,3]`` (one camera centre) or ``[C,N,3]``.  A separate autograd node, so a caller that evaluates it AFTER the
    projection (`pipeline.StreetScene.render`) gets its backward -- the largest gradient of the step -- BEFORE the
    projection's, and a gradient all-reduce started from a hook overlaps the rest of the backward pass."""
    flat, multi = _flat_cams(cam_pos)
    rgbs = _ActColors.apply(means_world, features_dc, features_rest, flat, int(sh_degree_to_use))
    return rgbs if multi else rgbs[0]


def activate_geometry(opacities: Tensor, scales: Tensor, quats: Tensor, point_ids: Optional[Tensor] = None,
                      inst_valid: Optional[Tensor] = None):
    """Geometry half of ``get_gaussians`` alone -> opacities[N] (sigmoid x frame-valid), scales (exp), unit quats."""
    return _ActGeom.apply(opacities.reshape(-1), scales, quats, point_ids, inst_valid)


def activate_gaussians(means_world: Tensor, features_dc: Tensor, features_rest: Optional[Tensor], opacities: Tensor,
                       scales: Tensor, quats: Tensor, cam_pos, sh_degree_to_use: int,
                       point_ids: Optional[Tensor] = None, inst_valid: Optional[Tensor] = None):
    """-> rgbs in [0,1], opacities[N] (sigmoid x frame-valid), scales[N,3] (exp), quats[N,4] (unit).

    ``cam_pos`` is one camera centre (3 numbers -> rgbs[N,3]) or a list of up to 8
    (-> rgbs[C,N,3]; the SH coefficients are read once for all cameras).
    ``opacities`` are logits ``[N,1]`` or ``[N]``."""
    if isinstance(cam_pos, Tensor):
        cam_pos = cam_pos.detach().cpu().tolist()
    multi = len(cam_pos) > 0 and isinstance(cam_pos[0], (list, tuple))
    flat = [float(v) for cp in cam_pos for v in cp] if multi else [float(v) for v in cam_pos]
    assert len(flat) % 3 == 0 and 3 <= len(flat) <= 24, "cam_pos: 1..8 camera centres"
    rgbs, opac, sc, qn = _Activate.apply(means_world, features_dc, features_rest, opacities.reshape(-1), scales, quats,
                                         point_ids, inst_valid, flat, int(sh_degree_to_use))
    return (rgbs if multi else rgbs[0]), opac, sc, qn
