"""GPU versions of SURVEY.md section 8(c)'s adversarial micro-cases, with the EXPECTED side of every threshold
asserted (not just agreement with the oracle), and a value check of ``absgrad``.

  * alpha exactly at 1/255, one ulp above and one ulp below (both compositing flavours),
  * transmittance landing EXACTLY on 1e-4 (gsplat stops at ``<=``, the Inria rule at ``<``),
  * a Gaussian whose centre sits exactly on a tile corner and whose radius ends exactly on tile edges,
  * depth ties through the whole pipeline (sort stability decides the compositing order),
  * ``info["means2d"].absgrad`` == sum over pixels of |per-pixel gradient| (the densification statistic the reference
    reads at ``OmniRe/models/trainers/base.py:282``).
"""
import pytest
import torch

from tests.util import bits, rel_err

pytestmark = pytest.mark.gpu

A255 = torch.tensor(1.0 / 255.0, dtype=torch.float32)


def _stage_forward(means2d, conics, colors, opac, depths, radii, W, H, flavour):
    """Hand-made per-Gaussian screen-space inputs through binning + sort + compositing (the C-ABI stages), forward only."""
    from emd_b200 import raster_ops as R
    from oracle import gsplat_ref as G
    dev = "cuda"
    tw, th, _ = R.tile_grid(W, H)
    x0, y0, x1, y1 = G.tile_rects(means2d[None], radii[None], 16, tw, th)
    tpg = ((x1 - x0) * (y1 - y0)).to(torch.int32)
    m2, cn, cl, op, dp, rd = (t.to(dev) for t in (means2d[None], conics[None], colors[None], opac[None], depths[None], radii[None]))
    _, ids, flat, cum = R.isect_tiles(m2, rd.contiguous(), dp, tpg.to(dev).contiguous(), W, H)
    offs = R.isect_offset_encode(ids, 1, W, H)
    out, alpha, last = R.rasterize_to_pixels(m2, cn, cl, op, None, None, rd.contiguous(), cum, offs, flat, ids, W, H,
                                             with_depth=False, ed_mode=False, absgrad=False, flavour=flavour)
    kw = dict(max_alpha=0.99, t_stop_inclusive=False, pixel_center=0.0) if flavour == 1 else {}
    rout, ralpha, rlast, unstable = G.rasterize_to_pixels(means2d[None], conics[None], colors[None], opac[None], W, H, 16,
                                                          offs.cpu(), flat.cpu(), return_unstable=True, **kw)
    return (out[0].cpu(), alpha[0, ..., 0].cpu(), last[0].cpu()), (rout[0], ralpha[0, ..., 0], rlast[0], unstable[0])


@pytest.mark.parametrize("flavour", [0, 1])
def test_alpha_threshold_one_ulp_either_side(flavour):
    """A Gaussian centred exactly on a pixel centre has alpha == opacity there (sigma = 0, exp(0) = 1 exactly):
    opacity = 1/255 and 1/255 + 1 ulp are blended, 1/255 - 1 ulp is skipped (``alpha < 1/255`` -> continue)."""
    W, H = 96, 32
    off = 0.5 if flavour == 0 else 0.0   # gsplat samples pixel centres, the Inria rule integer coordinates
    up, down = torch.nextafter(A255, torch.tensor(1.0)), torch.nextafter(A255, torch.tensor(0.0))
    opac = torch.stack([down, A255, up])
    cx = torch.tensor([8.0, 40.0, 72.0]) + off
    means2d = torch.stack([cx, torch.full((3,), 8.0 + off)], -1)
    conics = torch.tensor([[0.5, 0.0, 0.5]]).expand(3, 3).contiguous()
    colors = torch.tensor([[1.0, 0.5, 0.25]]).expand(3, 3).contiguous()
    depths = torch.tensor([1.0, 2.0, 3.0])
    radii = torch.full((3,), 5, dtype=torch.int32)
    (out, alpha, last), (rout, ralpha, rlast, unstable) = _stage_forward(means2d, conics, colors, opac, depths, radii, W, H, flavour)
    # expected side, stated: below -> nothing anywhere in its tile; at / above -> exactly the centre pixel
    assert float(alpha[:, :32].abs().max()) == 0.0 and float(out[:, :32].abs().max()) == 0.0
    for k, x in ((1, 40), (2, 72)):
        want = 1.0 - (1.0 - float(opac[k]))
        blk = alpha[:, x - 8:x + 24]
        assert int((blk > 0).sum()) == 1 and blk[8, 8] > 0
        assert abs(float(blk[8, 8]) - want) <= 1e-7
        assert abs(float(out[8, x, 0]) - float(opac[k])) <= 1e-7
    # and the oracle agrees on every pixel it does not flag (it flags exactly the three threshold pixels)
    ok = ~unstable
    assert float((alpha - ralpha).abs()[ok].max()) <= 1e-6
    assert float((alpha - ralpha).abs().max()) <= 1e-6, "the oracle's own threshold decisions agree too"


@pytest.mark.parametrize("flavour", [0, 1])
def test_transmittance_lands_exactly_on_the_stop_threshold(flavour):
    """T after k Gaussians is constructed exactly: alpha_0 = 1 - fl(1e-4) * 2^13 (Sterbenz-exact), then alpha = 1/2.
    Before Gaussian 13 the next transmittance would be EXACTLY fl(1e-4): gsplat (``<= 1e-4``) stops in front of it,
    the Inria rule (``< 1e-4``) blends it and stops in front of Gaussian 14."""
    W, H = 16, 16
    off = 0.5 if flavour == 0 else 0.0
    t = torch.tensor(1e-4, dtype=torch.float32)
    c = t * 8192.0                       # exact (power-of-two scaling), 0.8192...
    a0 = 1.0 - c                         # exact for c in [0.5, 1]
    assert float(1.0 - a0) == float(c)
    n = 16
    opac = torch.cat([a0.reshape(1), torch.full((n - 1,), 0.5)])
    means2d = torch.full((n, 2), 8.0 + off)
    conics = torch.tensor([[0.02, 0.0, 0.02]]).expand(n, 3).contiguous()
    colors = torch.rand(n, 3, generator=torch.Generator().manual_seed(3))
    depths = torch.arange(1, n + 1, dtype=torch.float32)
    radii = torch.full((n,), 40, dtype=torch.int32)
    (out, alpha, last), (rout, ralpha, rlast, unstable) = _stage_forward(means2d, conics, colors, opac, depths, radii, W, H, flavour)
    if flavour == 0:
        want_last, want_T = 12, float(c) / 4096.0          # 13 blended (indices 0..12), T = 2 * fl(1e-4)
    else:
        want_last, want_T = 13, float(t)                    # 14 blended, T = fl(1e-4) exactly
    assert int(last[8, 8]) == want_last, f"last blended index {int(last[8, 8])}, expected {want_last}"
    assert abs(float(alpha[8, 8]) - (1.0 - want_T)) <= 1.2e-7
    # the colour at that pixel is the exact front-to-back sum of the expected prefix
    T, acc = 1.0, torch.zeros(3, dtype=torch.float64)
    for k in range(want_last + 1):
        acc += colors[k].double() * float(opac[k]) * T
        T *= 1.0 - float(opac[k])
    assert float((out[8, 8].double() - acc).abs().max()) <= 2e-6
    ok = ~unstable
    assert float((out - rout).abs().amax(-1)[ok].max()) <= 1e-4 and float((alpha - ralpha).abs()[ok].max()) <= 1e-4


def test_gaussian_centred_on_a_tile_corner():
    """96x64 image (fx = 103, cx = 48, cy = 32): a Gaussian on the optical axis projects EXACTLY onto the corner shared
    by tiles (2,1),(3,1),(2,2),(3,2); with radius 16 its rectangle ends exactly on tile edges: floor(3 - 1) = 2 ..
    ceil(3 + 1) = 4 and floor(2 - 1) = 1 .. ceil(2 + 1) = 3 -> exactly 4 tiles, none of the neighbours."""
    import emd_b200
    from emd_b200 import scenes
    from oracle import gsplat_ref as G
    W, H = 96, 64
    viewmats, Ks, c2w = scenes.cameras((0.0,), W, H)
    means = torch.tensor([[10.0, 0.0, 1.6], [20.0, 0.0, 1.6], [5.0, 0.0, 1.6]])
    scales = torch.tensor([[0.5, 0.5, 0.5], [1.0, 1.0, 1.0], [0.02, 0.02, 0.02]])
    quats = torch.tensor([[1.0, 0.0, 0.0, 0.0]]).expand(3, 4).contiguous()
    opac = torch.tensor([0.6, 0.7, 0.9])
    colors = torch.tensor([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    rc, ra, rm = G.rasterization(means, quats, scales, opac, colors, viewmats, Ks, W, H, near_plane=0.1,
                                 render_mode="RGB+ED", return_unstable=True)
    assert torch.equal(rm["means2d"][0], torch.tensor([[48.0, 32.0]] * 3)), "construction: centres exactly on the corner"
    assert rm["radii"][0].tolist() == [16, 16, 3], rm["radii"]
    assert rm["tiles_per_gauss"][0].tolist() == [4, 4, 4]          # radius 3 around a corner still touches 4 tiles
    dev = "cuda"
    gc, ga, gm = emd_b200.rasterization(means.to(dev), quats.to(dev), scales.to(dev), opac.to(dev), colors.to(dev),
                                        viewmats.to(dev), Ks.to(dev), W, H, near_plane=0.1, packed=False,
                                        render_mode="RGB+ED")
    assert gm["radii"][0].tolist() == [16, 16, 3] and gm["tiles_per_gauss"][0].tolist() == [4, 4, 4]
    tiles = sorted(set((gm["isect_ids"].cpu() >> 32).tolist()))
    assert tiles == [1 * 6 + 2, 1 * 6 + 3, 2 * 6 + 2, 2 * 6 + 3], tiles
    assert torch.equal(gm["isect_ids"].cpu(), rm["isect_ids"]) and torch.equal(gm["flatten_ids"].cpu(), rm["flatten_ids"])
    assert torch.equal(gm["isect_offsets"].cpu(), rm["isect_offsets"])
    ok = ~rm["unstable"]
    assert float((gc.cpu() - rc)[..., :3].abs().amax(-1)[ok].max()) <= 1e-4
    assert float((ga.cpu() - ra).abs()[..., 0][ok].max()) <= 1e-4
    # nothing leaked into the neighbouring tiles beyond the rectangle
    assert float(ga[0, :16].abs().max()) == 0.0 and float(ga[0, 48:].abs().max()) == 0.0
    assert float(ga[0, :, :32].abs().max()) == 0.0 and float(ga[0, :, 64:].abs().max()) == 0.0


def test_depth_ties_through_the_whole_pipeline():
    """Every Gaussian sits at the same camera depth (and some are exact duplicates): the 64-bit keys tie inside a tile,
    so the compositing order is decided by the STABILITY of the radix sort (ascending Gaussian index, the emission
    order) -- checked explicitly, against the oracle's stable sort, and in the image."""
    import emd_b200
    from emd_b200 import scenes
    from oracle import gsplat_ref as G
    W, H = 96, 64
    viewmats, Ks, c2w = scenes.cameras((0.0,), W, H)
    g = torch.Generator().manual_seed(21)
    n = 300
    lat = (torch.rand(n, 2, generator=g) - 0.5) * torch.tensor([8.0, 5.0])
    means = torch.stack([torch.full((n,), 10.0), lat[:, 0], 1.6 + lat[:, 1]], -1)
    means[100:150] = means[50:100]                     # exact duplicates
    scales = 0.15 + 0.3 * torch.rand(n, 3, generator=g)
    quats = scenes.random_quats(n, g)
    opac = 0.1 + 0.5 * torch.rand(n, generator=g)
    colors = torch.rand(n, 3, generator=g)
    rc, ra, rm = G.rasterization(means, quats, scales, opac, colors, viewmats, Ks, W, H, near_plane=0.1,
                                 render_mode="RGB+ED", return_unstable=True)
    vis = rm["radii"][0] > 0
    assert int(vis.sum()) > 250
    assert len(set(bits(rm["depths"][0][vis]).tolist())) == 1, "construction: one depth bit pattern for every Gaussian"
    dev = "cuda"
    gc, ga, gm = emd_b200.rasterization(means.to(dev), quats.to(dev), scales.to(dev), opac.to(dev), colors.to(dev),
                                        viewmats.to(dev), Ks.to(dev), W, H, near_plane=0.1, packed=False,
                                        render_mode="RGB+ED")
    keys, ids = gm["isect_ids"].cpu(), gm["flatten_ids"].cpu()
    same = keys[1:] == keys[:-1]
    assert int(same.sum()) > 1000, "construction: long runs of tied keys"
    assert bool((ids[1:][same] > ids[:-1][same]).all()), "tied keys must keep emission (ascending index) order"
    assert torch.equal(keys, rm["isect_ids"]) and torch.equal(ids, rm["flatten_ids"])
    ok = ~rm["unstable"]
    assert float(ok.float().mean()) > 0.99
    assert float((gc.cpu() - rc)[..., :3].abs().amax(-1)[ok].max()) <= 1e-4
    assert torch.equal(gm["last_ids"].cpu()[ok], rm["last_ids"][ok])


def test_absgrad_values():
    """absgrad[c, n] = sum over pixels of |d l_p / d means2d[c, n]| with l_p the pixel's own loss term -- computed
    on the CPU from one oracle backward per pixel -- next to grad = |sum|.  RGB+ED, two cameras."""
    import emd_b200
    from oracle import gsplat_ref as G
    from tests.util import raster_scene
    W, H = 32, 32
    sc, viewmats, Ks, _, g = raster_scene(31, 60, W, H, yaws=(0.0, 10.0), depth=(2.0, 12.0), scale=0.08)
    C = 2
    q = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
    rc, ra, rm = G.rasterization(q["means"], q["quats"], q["scales"], q["opacities"], q["colors"], viewmats, Ks, W, H,
                                 near_plane=0.1, render_mode="RGB+ED", return_unstable=True)
    keep = (~rm["unstable"]).float()[..., None]
    vc = torch.randn(rc.shape, generator=g) * keep
    vc[..., 3] *= 0.05
    va = torch.randn(ra.shape, generator=g) * keep
    per_pixel = ((rc * vc).sum(-1) + (ra * va).sum(-1)).reshape(-1)      # [C*H*W]
    m2 = rm["means2d"]
    want_abs = torch.zeros_like(m2)
    want_sum = torch.zeros_like(m2)
    for p in range(per_pixel.numel()):
        (gp,) = torch.autograd.grad(per_pixel[p], m2, retain_graph=True)
        want_abs += gp.abs()
        want_sum += gp
    assert float(want_abs.max()) > 0 and float((want_abs - want_sum.abs()).max()) > 1e-3 * float(want_abs.max()), \
        "construction: cancellation must make absgrad differ from |grad|"
    dev = "cuda"
    p = {k: v.to(dev).requires_grad_(True) for k, v in sc.items()}
    gc, ga, gm = emd_b200.rasterization(p["means"], p["quats"], p["scales"], p["opacities"], p["colors"],
                                        viewmats.to(dev), Ks.to(dev), W, H, near_plane=0.1, packed=False, absgrad=True,
                                        render_mode="RGB+ED")
    gm["means2d"].retain_grad()
    ((gc * vc.to(dev)).sum() + (ga * va.to(dev)).sum()).backward()
    assert rel_err(gm["means2d"].grad, want_sum) <= 1e-3
    assert rel_err(gm["means2d"].absgrad, want_abs) <= 1e-3, "absgrad values"
    # as the reference consumes it (base.py:279-297): norm over xy of the visible Gaussians of camera 0
    sel = (gm["radii"][0] > 0).cpu()
    got = gm["means2d"].absgrad[0].cpu()[sel].norm(dim=-1)
    assert rel_err(got, want_abs[0][sel].norm(dim=-1)) <= 1e-3
