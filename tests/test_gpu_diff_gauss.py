"""GPU parity: emd_b200.diff_gauss_api (the diff_gauss drop-in S3Gaussian calls) vs the oracle."""
import pytest
import torch

from tests.dg_util import s3g_camera
from tests.util import bits, rel_err, rel_l2

pytestmark = pytest.mark.gpu


def _scene(seed, n, W, H, K=16):
    from emd_b200 import scenes
    g = torch.Generator().manual_seed(seed)
    sc = scenes.simple_gaussians(n, g, W, H, scale=0.05)
    q = sc["quats"] / sc["quats"].norm(dim=-1, keepdim=True)  # rotation_activation = normalize
    shs = torch.cat([(torch.rand(n, 1, 3, generator=g) - 0.5) / 0.2820948, 0.1 * torch.randn(n, K - 1, 3, generator=g)], 1)
    return dict(means=sc["means"], scales=sc["scales"], rots=q, opac=sc["opacities"][:, None], shs=shs,
                colors=sc["colors"]), g


@pytest.mark.parametrize("seed,n,W,H,use_sh,deg,mod", [
    (0, 2500, 240, 160, True, 3, 1.0),
    (1, 1500, 200, 136, False, 0, 1.0),
    (2, 1200, 160, 96, True, 1, 0.7),
])
def test_diff_gauss_parity(seed, n, W, H, use_sh, deg, mod):
    from emd_b200.diff_gauss_api import GaussianRasterizationSettings, GaussianRasterizer
    from oracle import diff_gauss_ref as DG
    p, g = _scene(seed, n, W, H)
    cam = s3g_camera(10.0 * seed, W, H)
    bg = torch.tensor([0.2, 0.5, 0.1])
    names = ["means", "scales", "rots", "opac", "shs" if use_sh else "colors"]
    cpu = {k: p[k].clone().requires_grad_(True) for k in names}
    m2_c = torch.zeros(n, 3, requires_grad=True)
    s_c = DG.Settings(H, W, cam["tanfovx"], cam["tanfovy"], bg, mod, cam["viewmatrix"], cam["projmatrix"], deg, cam["campos"])
    rc, rd, rn, ra, rr, info = DG.rasterize(cpu["means"], m2_c, cpu["shs"] if use_sh else None,
                                            None if use_sh else cpu["colors"], cpu["opac"], cpu["scales"], cpu["rots"],
                                            s_c, return_unstable=True)
    dev = "cuda"
    gpu = {k: p[k].to(dev).requires_grad_(True) for k in names}
    m2_g = torch.zeros(n, 3, device=dev, requires_grad=True)
    s_g = GaussianRasterizationSettings(H, W, cam["tanfovx"], cam["tanfovy"], bg.to(dev), mod, cam["viewmatrix"].to(dev),
                                        cam["projmatrix"].to(dev), deg, cam["campos"].to(dev), False, False)
    gc, gd, gn, ga, gr, _ = GaussianRasterizer(raster_settings=s_g)(
        means3D=gpu["means"], means2D=m2_g, shs=gpu["shs"] if use_sh else None,
        colors_precomp=None if use_sh else gpu["colors"], opacities=gpu["opac"], scales=gpu["scales"],
        rotations=gpu["rots"], cov3Ds_precomp=None, extra_attrs=None)
    assert gc.shape == (3, H, W) and gd.shape == (1, H, W) and ga.shape == (1, H, W) and gn.shape == (3, H, W)
    assert torch.equal(gr.cpu(), rr), "radii differ"
    assert int((rr > 0).sum()) > n // 3
    ok = ~info["unstable"]
    assert float(info["unstable"].float().mean()) < 2e-3
    assert float((gc.detach().cpu() - rc.detach()).abs()[:, ok].max()) <= 1e-4
    assert float((ga.detach().cpu() - ra.detach()).abs()[:, ok].max()) <= 1e-4
    assert float((gd.detach().cpu() - rd.detach()).abs()[:, ok].max()) <= 1e-4 * max(1.0, float(rd.detach().max()))
    keep = ok.float()[None]
    vc = torch.randn(3, H, W, generator=g) * keep
    vd = 0.05 * torch.randn(1, H, W, generator=g) * keep
    va = torch.randn(1, H, W, generator=g) * keep
    ((rc * vc).sum() + (rd * vd).sum() + (ra * va).sum()).backward()
    ((gc * vc.to(dev)).sum() + (gd * vd.to(dev)).sum() + (ga * va.to(dev)).sum()).backward()
    for k in names:
        e, l2 = rel_err(gpu[k].grad, cpu[k].grad), rel_l2(gpu[k].grad, cpu[k].grad)
        assert e <= 1e-3 and l2 <= 1e-3, f"grad {k}: max-rel {e}, l2-rel {l2}"
    assert rel_err(m2_g.grad, m2_c.grad) <= 1e-3, "screen-space (means2D) gradient"
    assert float(m2_g.grad[:, 2].abs().max()) == 0.0


def test_diff_gauss_rejects_unsupported():
    from emd_b200.diff_gauss_api import GaussianRasterizationSettings, GaussianRasterizer
    p, g = _scene(5, 100, 64, 48)
    cam = s3g_camera(0.0, 64, 48)
    dev = "cuda"
    s = GaussianRasterizationSettings(48, 64, cam["tanfovx"], cam["tanfovy"], torch.zeros(3, device=dev), 1.0,
                                      cam["viewmatrix"].to(dev), cam["projmatrix"].to(dev), 0, cam["campos"].to(dev),
                                      False, False)
    r = GaussianRasterizer(s)
    kw = dict(means3D=p["means"].to(dev), means2D=torch.zeros(100, 3, device=dev), opacities=p["opac"].to(dev),
              scales=p["scales"].to(dev), rotations=p["rots"].to(dev))
    with pytest.raises(ValueError):
        r(shs=None, colors_precomp=None, **kw)
    with pytest.raises(NotImplementedError):
        r(shs=None, colors_precomp=p["colors"].to(dev), cov3Ds_precomp=torch.zeros(100, 6, device=dev), **kw)
