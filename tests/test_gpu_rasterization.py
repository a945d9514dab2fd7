"""GPU parity: emd_b200.rasterization (CUDA, through the C ABI) vs the CPU oracle.

Bars: radii / tiles-per-Gaussian / sort keys / sorted ids /
tile offsets bit-exact; images, depth, alpha <= 1e-4 absolute; gradients <= 1e-3
relative.  Pixels whose value hinges on a comparison that sits within rounding
distance of its threshold (alpha vs 1/255, T vs 1e-4, sigma vs 0) are identified
by the oracle, excluded, and bounded in number.
"""
import pytest
import torch

from tests.util import bits, raster_scene, rel_err, rel_l2

pytestmark = pytest.mark.gpu


def _run_both(seed, n, W, H, yaws, render_mode="RGB+ED", backgrounds=False, radius_clip=0.0, scale=0.05,
              near=0.1, grads=True):
    import emd_b200
    from oracle import gsplat_ref as G

    sc, viewmats, Ks, _, g = raster_scene(seed, n, W, H, yaws, scale=scale)
    C = viewmats.shape[0]
    D = {"RGB": 3, "D": 1, "ED": 1, "RGB+D": 4, "RGB+ED": 4}[render_mode]
    bg = torch.rand(C, 3, generator=g) if backgrounds else None
    names = ["means", "quats", "scales", "opacities", "colors"]
    cpu = {k: sc[k].clone().requires_grad_(grads) for k in names}
    rc, ra, meta = G.rasterization(cpu["means"], cpu["quats"], cpu["scales"], cpu["opacities"], cpu["colors"],
                                   viewmats, Ks, W, H, near_plane=near, radius_clip=radius_clip,
                                   backgrounds=bg, render_mode=render_mode, return_unstable=True)
    dev = torch.device("cuda")
    gpu = {k: sc[k].to(dev).requires_grad_(grads) for k in names}
    gc, ga, gmeta = emd_b200.rasterization(gpu["means"], gpu["quats"], gpu["scales"], gpu["opacities"], gpu["colors"],
                                           viewmats.to(dev), Ks.to(dev), W, H, near_plane=near,
                                           radius_clip=radius_clip, packed=False, absgrad=True,
                                           backgrounds=bg.to(dev) if bg is not None else None,
                                           render_mode=render_mode)
    return sc, cpu, gpu, (rc, ra, meta), (gc, ga, gmeta), g, D


def _check_integers(meta, gmeta):
    assert torch.equal(gmeta["radii"].cpu(), meta["radii"]), "radii differ"
    assert torch.equal(gmeta["tiles_per_gauss"].cpu(), meta["tiles_per_gauss"]), "tiles_per_gauss differ"
    assert torch.equal(bits(gmeta["means2d"]), bits(meta["means2d"])), "means2d bits differ"
    assert torch.equal(bits(gmeta["depths"]), bits(meta["depths"])), "depth bits differ"
    assert torch.equal(bits(gmeta["conics"]), bits(meta["conics"])), "conic bits differ"
    assert gmeta["isect_ids"].numel() == meta["isect_ids"].numel(), "n_isects differ"
    assert torch.equal(gmeta["isect_ids"].cpu(), meta["isect_ids"]), "sorted 64-bit keys differ"
    assert torch.equal(gmeta["flatten_ids"].cpu(), meta["flatten_ids"]), "sorted Gaussian ids differ"
    assert torch.equal(gmeta["isect_offsets"].cpu(), meta["isect_offsets"]), "tile offsets differ"


def _check_images(rc, ra, meta, gc, ga, tol=1e-4, max_unstable_frac=2e-3):
    unstable = meta["unstable"]
    frac = unstable.float().mean().item()
    assert frac <= max_unstable_frac, f"too many threshold-ambiguous pixels: {frac}"
    ok = ~unstable
    dc = (gc.detach().cpu() - rc.detach()).abs()
    da = (ga.detach().cpu() - ra.detach()).abs()
    # depth channel is metric (tens of metres): the absolute bar applies to colour/alpha, depth is checked relative
    assert float(da[ok].max()) <= tol, f"alpha differs by {float(da[ok].max())}"
    D = rc.shape[-1]
    for k in range(D):
        ref = rc.detach()[..., k]
        scale = max(1.0, float(ref.abs().max()))
        err = float(dc[..., k][ok].max())
        assert err <= tol * scale, f"channel {k} differs by {err} (scale {scale})"
    return frac


@pytest.mark.parametrize("seed,n,W,H,yaws,mode,bg", [
    (0, 3000, 240, 160, (0.0,), "RGB+ED", False),
    (1, 2000, 200, 136, (0.0, 20.0), "RGB", True),      # ragged tiles (136 = 8.5 tiles), 2 cameras, background
    (2, 1500, 160, 96, (0.0,), "RGB+D", False),
    (3, 800, 96, 64, (0.0, -30.0, 30.0), "ED", False),
    # several 1024-Gaussian segments per tile: transmittance pass + per-segment compositing + combine
    (4, 12000, 64, 48, (0.0,), "RGB+ED", False),
    (7, 9000, 64, 40, (0.0, 15.0), "RGB", True),
])
def test_forward_parity(seed, n, W, H, yaws, mode, bg):
    _, _, _, (rc, ra, meta), (gc, ga, gmeta), _, _ = _run_both(seed, n, W, H, yaws, mode, bg, grads=False)
    _check_integers(meta, gmeta)
    _check_images(rc, ra, meta, gc, ga)
    assert torch.equal(gmeta["last_ids"].cpu()[~meta["unstable"]], meta["last_ids"][~meta["unstable"]])


@pytest.mark.parametrize("seed,n,W,H,yaws,mode,bg", [
    (10, 2500, 208, 144, (0.0,), "RGB+ED", False),
    (11, 1200, 128, 96, (0.0, 25.0), "RGB", True),
    (12, 1000, 112, 80, (0.0,), "RGB+D", False),
    # long per-tile lists (several 1024-Gaussian segments per tile): the segment-parallel backward + checkpoints
    (13, 9000, 64, 48, (0.0,), "RGB+ED", False),
    (14, 7000, 48, 32, (0.0, 20.0), "RGB", True),
])
def test_backward_parity(seed, n, W, H, yaws, mode, bg):
    sc, cpu, gpu, (rc, ra, meta), (gc, ga, gmeta), g, D = _run_both(seed, n, W, H, yaws, mode, bg)
    _check_integers(meta, gmeta)
    _check_images(rc, ra, meta, gc, ga)
    keep = (~meta["unstable"]).float()[..., None]
    vc = torch.randn(rc.shape, generator=g) * keep
    va = torch.randn(ra.shape, generator=g) * keep
    if mode in ("RGB+ED", "RGB+D"):
        vc[..., 3] *= 0.05  # depth is in metres; keep its cotangent commensurate
    meta["means2d"].retain_grad()
    gmeta["means2d"].retain_grad()
    ((rc * vc).sum() + (ra * va).sum()).backward()
    ((gc * vc.cuda()).sum() + (ga * va.cuda()).sum()).backward()
    for k in ("means", "quats", "scales", "opacities", "colors"):
        e, l2 = rel_err(gpu[k].grad, cpu[k].grad), rel_l2(gpu[k].grad, cpu[k].grad)
        assert e <= 1e-3 and l2 <= 1e-3, f"grad {k}: max-rel {e}, l2-rel {l2}"
    e = rel_err(gmeta["means2d"].grad, meta["means2d"].grad)
    assert e <= 1e-3, f"means2d.grad: {e}"
    # absgrad: sum over pixels of |per-pixel grad|; not available from autograd -> check the invariant |grad| <= absgrad
    ab = gmeta["means2d"].absgrad
    assert ab.shape == gmeta["means2d"].shape
    assert bool((ab + 1e-6 * ab.abs().max() >= gmeta["means2d"].grad.abs()).all())


def test_determinism():
    import emd_b200
    sc, viewmats, Ks, _, g = raster_scene(5, 4000, 256, 160)
    dev = "cuda"
    outs = []
    for _ in range(2):
        p = {k: v.to(dev).requires_grad_(True) for k, v in sc.items()}
        c, a, m = emd_b200.rasterization(p["means"], p["quats"], p["scales"], p["opacities"], p["colors"],
                                         viewmats.to(dev), Ks.to(dev), 256, 160, packed=False, render_mode="RGB+ED")
        (c.sum() + a.sum()).backward()
        outs.append([c.detach().clone(), a.detach().clone()] + [p[k].grad.clone() for k in sorted(p)])
    for x, y in zip(*outs):
        assert torch.equal(x, y), "two identical calls gave different bits"


def test_zero_opacity_and_empty():
    import emd_b200
    sc, viewmats, Ks, _, g = raster_scene(6, 500, 128, 96)
    dev = "cuda"
    op = sc["opacities"].clone()
    op[::2] = 0.0  # class mask pre-multiplied (base.py:397)
    c, a, m = emd_b200.rasterization(sc["means"].to(dev), sc["quats"].to(dev), sc["scales"].to(dev), op.to(dev),
                                     sc["colors"].to(dev), viewmats.to(dev), Ks.to(dev), 128, 96, packed=False)
    assert torch.isfinite(c).all() and torch.isfinite(a).all()
    # nothing visible: camera looking away
    vm = viewmats.clone()
    vm[:, 2, 3] -= 1000.0
    c, a, m = emd_b200.rasterization(sc["means"].to(dev), sc["quats"].to(dev), sc["scales"].to(dev), op.to(dev),
                                     sc["colors"].to(dev), vm.to(dev), Ks.to(dev), 128, 96, packed=False)
    assert m["isect_ids"].numel() == 0 and float(a.abs().max()) == 0.0 and float(c.abs().max()) == 0.0
    # zero Gaussians
    z = torch.zeros(0, 3, device=dev)
    c, a, m = emd_b200.rasterization(z, torch.zeros(0, 4, device=dev), z, torch.zeros(0, device=dev), z,
                                     viewmats.to(dev), Ks.to(dev), 128, 96, packed=False)
    assert float(a.abs().max()) == 0.0
