"""GPU parity of the fused image losses (csrc/image_loss.cu through emd_b200.losses) against the oracle: loss terms and
the cotangents of colour / depth / opacity / sky, OmniRe (HWC, C views) and S3Gaussian (CHW) flavours."""
import pytest
import torch

from tests.loss_util import OMNIRE_CASES, OMNIRE_KEYS, loss_inputs, oracle_omnire, oracle_s3g

pytestmark = pytest.mark.gpu


def _close(a, b, tol, what):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    err = float((a - b).abs().max())
    ref = max(float(b.abs().max()), 1e-30)
    assert err <= tol * ref, f"{what}: max err {err:.3e} vs scale {ref:.3e}"


def _omnire_gpu(views, cfg, use_sky, use_ego, v_terms):
    from emd_b200 import losses as PL
    dev = "cuda"
    st = lambda k: torch.stack([d[k] for d in views]).to(dev)
    renders = torch.stack([torch.cat([d["rgb"], d["depth"]], -1) for d in views]).to(dev).requires_grad_(True)
    alphas = st("alpha").requires_grad_(True)
    sky = st("sky").requires_grad_(True) if use_sky else None
    valid = (1.0 - st("ego_mask")) if use_ego else None
    terms, sums = PL.image_losses_hwc(renders, alphas, st("gt"), cfg, rgb_sky=sky, valid_mask=valid, sky_masks=st("sky_mask"),
                                      lidar_depth_map=st("lidar"))
    (terms * v_terms.to(dev)).sum().backward()
    return terms, renders.grad, alphas.grad, (sky.grad if use_sky else None)


@pytest.mark.parametrize("case", OMNIRE_CASES, ids=[c[0] for c in OMNIRE_CASES])
@pytest.mark.parametrize("hw", [(40, 56), (64, 96), (33, 17)])
def test_image_loss_omnire_vs_oracle(case, hw):
    from emd_b200 import losses as PL
    name, okw, pkw, use_sky, use_ego = case
    H, W = hw
    views = [loss_inputs(21, H, W), loss_inputs(22, H, W), loss_inputs(23, H, W)]
    g = torch.Generator().manual_seed(1)
    v_terms = torch.rand(3, 6, generator=g) + 0.5
    terms_o, gr_o, ga_o, gs_o = oracle_omnire(views, v_terms, use_sky, use_ego, okw)
    cfg = PL.ImageLossConfig.omnire()
    for k, v in pkw.items():
        setattr(cfg, k, v)
    terms, gr, ga, gs = _omnire_gpu(views, cfg, use_sky, use_ego, v_terms)
    _close(terms, terms_o, 2e-5, "terms")
    _close(gr[..., :3], gr_o[..., :3], 2e-4, "v_rgb")
    _close(gr[..., 3], gr_o[..., 3], 2e-4, "v_depth")
    _close(ga, ga_o, 2e-4, "v_alpha")
    if use_sky:
        _close(gs, gs_o, 2e-4, "v_sky")


@pytest.mark.parametrize("use_sky,use_mask", [(True, True), (False, False)])
def test_image_loss_s3g_vs_oracle(use_sky, use_mask):
    from emd_b200 import losses as PL
    H, W = 37, 50
    d = loss_inputs(31, H, W)
    d["rgb"] = d["rgb"].clamp(max=1.0)
    v_terms = torch.tensor([1.0, 0.7, 1.3, 0.9, 0.0, 0.0])
    t_o, gc_o, gd_o, gw_o, gs_o = oracle_s3g(d, v_terms, use_sky, use_mask)
    cfg = PL.ImageLossConfig.s3g()
    chw = lambda k: d[k].permute(2, 0, 1).contiguous().cuda()
    color, depth, weight = chw("rgb").requires_grad_(True), chw("depth").requires_grad_(True), chw("alpha").requires_grad_(True)
    sky = chw("sky").requires_grad_(True) if use_sky else None
    out = PL.s3g_image_losses(color, depth, weight, sky, chw("gt"), d["lidar"][None].cuda(),
                              d["sky_mask"][None].bool().cuda() if use_mask else None, cfg)
    keys = ["Ll1", "ssim_loss", "sky_loss", "depth_loss"]
    total = sum(out[k] * v_terms[i] for i, k in enumerate(keys) if k in out)
    total.backward()
    for i, k in enumerate(keys):
        if k in out:
            assert abs(float(out[k]) - float(t_o[i])) <= 2e-5 * max(1.0, abs(float(t_o[i]))), k
    _close(color.grad, gc_o, 2e-4, "v_color")
    _close(depth.grad, gd_o, 2e-4, "v_depth")
    _close(weight.grad, gw_o, 2e-4, "v_weight")
    if use_sky:
        _close(sky.grad, gs_o, 2e-4, "v_sky")


def test_image_loss_reference_keys_and_dict():
    """omnire_image_losses returns compute_losses' keys (base.py:542-584), summed over the views."""
    from emd_b200 import losses as PL
    H, W = 40, 56
    views = [loss_inputs(5, H, W), loss_inputs(6, H, W)]
    st = lambda k: torch.stack([d[k] for d in views]).cuda()
    renders = torch.stack([torch.cat([d["rgb"], d["depth"]], -1) for d in views]).cuda()
    infos = {"pixels": st("gt"), "sky_masks": st("sky_mask"), "egocar_masks": st("ego_mask"), "lidar_depth_map": st("lidar")}
    out = PL.omnire_image_losses(renders, st("alpha"), st("sky"), infos, PL.ImageLossConfig.omnire())
    assert tuple(out) == OMNIRE_KEYS
    terms_o, *_ = oracle_omnire(views, torch.ones(2, 6), True, True, {})
    for i, k in enumerate(OMNIRE_KEYS):
        assert abs(float(out[k]) - float(terms_o[:, i].sum())) <= 2e-5 * max(1.0, abs(float(terms_o[:, i].sum()))), k


def test_image_loss_full_size_properties():
    """BASELINE configs[1] image size (3 x 640 x 960): bit-reproducible, linear in the term cotangents, zero photometric
    terms and zero colour gradient for a perfect render, and one view checked against the oracle at full size."""
    from emd_b200 import losses as PL
    H, W, C = 640, 960, 3
    views = [loss_inputs(40 + c, H, W) for c in range(C)]
    cfg = PL.ImageLossConfig.omnire()
    v1 = torch.ones(C, 6)
    t1, gr1, ga1, gs1 = _omnire_gpu(views, cfg, True, False, v1)
    t2, gr2, ga2, gs2 = _omnire_gpu(views, cfg, True, False, v1)
    assert torch.equal(t1, t2) and torch.equal(gr1, gr2) and torch.equal(ga1, ga2) and torch.equal(gs1, gs2)
    assert torch.isfinite(t1).all() and torch.isfinite(gr1).all() and torch.isfinite(ga1).all()
    t3, gr3, ga3, _ = _omnire_gpu(views, cfg, True, False, 2.0 * v1)
    assert torch.equal(t3, t1)
    _close(gr3, 2.0 * gr1, 1e-6, "linearity v_renders")
    _close(ga3, 2.0 * ga1, 1e-6, "linearity v_alpha")
    # perfect render: gt := blended prediction
    perfect = []
    for d in views:
        e = dict(d)
        e["gt"] = (d["rgb"].clamp(max=1.0) + d["sky"] * (1.0 - d["alpha"]))
        perfect.append(e)
    tp, grp, _, _ = _omnire_gpu(perfect, cfg, True, False, v1)
    assert float(tp[:, 0].abs().max()) <= 1e-7 and float(tp[:, 1].abs().max()) <= 1e-6
    only_photo = torch.zeros(C, 6)
    only_photo[:, :2] = 1.0
    _, grq, _, _ = _omnire_gpu(perfect, cfg, True, False, only_photo)
    assert float(grq[..., :3].abs().max()) <= 1e-9
    # one full-size view against the oracle
    terms_o, gr_o, ga_o, gs_o = oracle_omnire(views[:1], v1[:1], True, False, {})
    _close(t1[:1], terms_o, 2e-5, "terms (full size)")
    _close(gr1[:1, ..., :3], gr_o[..., :3], 5e-4, "v_rgb (full size)")
    _close(gr1[:1, ..., 3], gr_o[..., 3], 5e-4, "v_depth (full size)")
    _close(ga1[:1], ga_o, 5e-4, "v_alpha (full size)")


def test_image_loss_errors():
    from emd_b200 import _C, losses as PL
    cfg = PL.ImageLossConfig.omnire()
    z = torch.zeros(1, 8, 8, 4, device="cuda")
    with pytest.raises(_C.EmdError):      # valid-window SSIM needs H, W > 10
        PL.image_losses_hwc(z, torch.zeros(1, 8, 8, 1, device="cuda"), torch.zeros(1, 8, 8, 3, device="cuda"), cfg)
    with pytest.raises(_C.EmdError):      # CPU tensors are refused
        PL.image_losses_hwc(torch.zeros(1, 16, 16, 4), torch.zeros(1, 16, 16, 1), torch.zeros(1, 16, 16, 3), cfg)
