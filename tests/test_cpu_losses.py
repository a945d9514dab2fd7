"""CPU: loss_math.cuh (the per-pixel arithmetic image_loss.cu includes) through the host build against the oracle --
loss terms and the cotangents of colour / depth / opacity / sky for the OmniRe and S3Gaussian flavours."""
import ctypes

import numpy as np
import pytest
import torch

from tests.loss_util import OMNIRE_CASES, loss_inputs, oracle_omnire, oracle_s3g

P = ctypes.c_void_p


@pytest.fixture(scope="module")
def hostlib():
    from emd_b200 import build
    return ctypes.CDLL(str(build.build_hostmath()))


def _p(a):
    return a.ctypes.data_as(P) if a is not None else None


def _np(t):
    return np.ascontiguousarray(t.detach().numpy(), dtype=np.float32) if t is not None else None


def _host_call(hostlib, cfg, lay, rgb_base, depth_off, alpha, sky, gt, valid, sky_mask, lidar, v_terms):
    """rgb_base: numpy buffer holding colour (and depth at float offset depth_off, or a separate buffer if tuple)."""
    from emd_b200 import losses as PL
    f = hostlib.emd_host_image_loss
    f.argtypes = [P] * 8 + [ctypes.c_int] * 3 + [P, ctypes.POINTER(ctypes.c_float), P, P, P, P, P, P]
    f.restype = None
    C, H, W = lay.C, lay.H, lay.W
    terms = np.zeros((C, 6), np.float32)
    if isinstance(rgb_base, tuple):
        rgb, depth = rgb_base
        v_rgb, v_depth = np.zeros_like(rgb), np.zeros_like(depth)
        rp, dp, vrp, vdp = _p(rgb), _p(depth), _p(v_rgb), _p(v_depth)
    else:
        rgb = rgb_base
        v_rgb = np.zeros_like(rgb)
        rp, vrp = rgb.ctypes.data, v_rgb.ctypes.data
        dp, vdp = rp + 4 * depth_off, vrp + 4 * depth_off
        v_depth = None
    v_alpha = np.zeros_like(alpha)
    v_sky = np.zeros_like(sky) if sky is not None else None
    cc = cfg._c(lay.strides)
    f(rp, dp, _p(alpha), _p(sky), _p(gt), _p(valid), _p(sky_mask), _p(lidar), C, H, W, ctypes.byref(cc), PL._window(),
      _p(terms), _p(v_terms), vrp, vdp, _p(v_alpha), _p(v_sky))
    return terms, v_rgb, v_depth, v_alpha, v_sky


def _close(a, b, tol, what):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    err = float((a - b).abs().max())
    ref = max(float(b.abs().max()), 1e-30)
    assert err <= tol * ref, f"{what}: max err {err:.3e} vs scale {ref:.3e}"


@pytest.mark.parametrize("case", OMNIRE_CASES, ids=[c[0] for c in OMNIRE_CASES])
def test_host_image_loss_omnire(hostlib, case):
    from emd_b200 import losses as PL
    name, okw, pkw, use_sky, use_ego = case
    H, W = 40, 56
    views = [loss_inputs(21, H, W), loss_inputs(22, H, W)]
    g = torch.Generator().manual_seed(1)
    v_terms = torch.rand(2, 6, generator=g) + 0.5
    terms_o, gr_o, ga_o, gs_o = oracle_omnire(views, v_terms, use_sky, use_ego, okw)
    cfg = PL.ImageLossConfig.omnire()
    for k, v in pkw.items():
        setattr(cfg, k, v)
    lay = PL._Layout(False, 2, H, W, True, 4)
    st = lambda k: _np(torch.stack([d[k] for d in views]))
    renders = _np(torch.stack([torch.cat([d["rgb"], d["depth"]], -1) for d in views]))
    valid = _np(torch.stack([1.0 - d["ego_mask"] for d in views])) if use_ego else None
    terms, v_r, _, v_a, v_s = _host_call(hostlib, cfg, lay, renders, 3, st("alpha"), st("sky") if use_sky else None, st("gt"),
                                         valid, st("sky_mask"), st("lidar"), _np(v_terms))
    _close(terms, terms_o, 2e-5, "terms")
    _close(v_r[..., :3], gr_o[..., :3], 2e-4, "v_rgb")
    _close(v_r[..., 3], gr_o[..., 3], 2e-4, "v_depth")
    _close(v_a, ga_o, 2e-4, "v_alpha")
    if use_sky:
        _close(v_s, gs_o, 2e-4, "v_sky")


@pytest.mark.parametrize("use_sky,use_mask", [(True, True), (False, False)])
def test_host_image_loss_s3g(hostlib, use_sky, use_mask):
    from emd_b200 import losses as PL
    H, W = 37, 50
    d = loss_inputs(31, H, W)
    d["rgb"] = d["rgb"].clamp(max=1.0)
    v_terms = torch.tensor([1.0, 0.7, 1.3, 0.9, 0.0, 0.0])
    t_o, gc_o, gd_o, gw_o, gs_o = oracle_s3g(d, v_terms, use_sky, use_mask)
    cfg = PL.ImageLossConfig.s3g()
    lay = PL._Layout(True, 1, H, W, True, 3)
    chw = lambda k: _np(d[k].permute(2, 0, 1))
    terms, v_c, v_d, v_w, v_s = _host_call(hostlib, cfg, lay, (chw("rgb"), chw("depth")), 0, chw("alpha"),
                                           chw("sky") if use_sky else None, chw("gt"), None,
                                           _np(d["sky_mask"][None]) if use_mask else None, _np(d["lidar"][None]),
                                           _np(v_terms[None]))
    _close(terms[0], t_o, 2e-5, "terms")
    _close(v_c, gc_o, 2e-4, "v_color")
    _close(v_d, gd_o, 2e-4, "v_depth")
    _close(v_w, gw_o, 2e-4, "v_weight")
    if use_sky:
        _close(v_s, gs_o, 2e-4, "v_sky")
