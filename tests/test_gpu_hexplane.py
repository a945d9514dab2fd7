"""GPU parity: the HexPlane gather (K1e, emd_hexplane_fwd/bwd through emd_b200.hexplane.HexPlaneField) against
the reference's own HexPlaneField outputs (tests/golden/hexplane.npz) and against the oracle."""
import os

import numpy as np
import pytest
import torch

from tests.hex_util import oracle_run
from tests.util import rel_err, rel_l2

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_t = lambda a: torch.from_numpy(np.asarray(a))  # noqa: E731


def _field(reso, mr, grids, aabb):
    from emd_b200.hexplane import HexPlaneField
    cfg = {"grid_dimensions": 2, "input_coordinate_dim": 4, "output_coordinate_dim": 32, "resolution": list(reso)}
    f = HexPlaneField(1.6, cfg, list(mr))
    f.set_aabb(aabb[0].tolist(), aabb[1].tolist())
    f.load_reference_grids(grids)
    return f.cuda()


@pytest.mark.parametrize("name", ["a", "b"])
def test_hexplane_golden(name):
    from oracle import hexplane as OH
    z = np.load(f"{G}/hexplane.npz")
    reso, mr = [int(v) for v in z[f"{name}_resolution"]], [int(v) for v in z[f"{name}_multires"]]
    grids = OH.hash_planes(reso, mr, salt=int(z[f"{name}_salt"]))
    f = _field(reso, mr, grids, _t(z[f"{name}_aabb"]))
    pts = _t(z[f"{name}_pts"]).cuda().requires_grad_(True)
    t = _t(z[f"{name}_t"]).cuda().requires_grad_(True)          # [N,1], one time per point (the reference's call shape)
    feat = f(pts, t)
    assert feat.shape == z[f"{name}_feat"].shape
    assert float((feat.detach().cpu() - _t(z[f"{name}_feat"])).abs().max()) <= 1e-6
    (feat * _t(z[f"{name}_cot"]).cuda()).sum().backward()
    assert rel_err(pts.grad, _t(z[f"{name}_v_pts"])) <= 1e-4
    assert rel_err(t.grad, _t(z[f"{name}_v_t"])) <= 1e-4
    got = [g for row in f.reference_grids(f.planes.grad) for g in row]
    sums = np.array([g.double().sum().item() for g in got])
    l2 = np.array([g.double().pow(2).sum().sqrt().item() for g in got])
    assert np.allclose(sums, z[f"{name}_v_plane_sum"], rtol=1e-4, atol=1e-4)
    assert np.allclose(l2, z[f"{name}_v_plane_l2"], rtol=1e-5)
    if name == "a":
        for k, g in enumerate(got):
            assert rel_err(g, _t(z[f"a_v_plane{k}"])) <= 1e-5, k


@pytest.mark.parametrize("shared_t", [True, False])
def test_hexplane_vs_oracle(shared_t):
    """Reference resolution [64,64,64,25] at two scales, 20 k points (25 % outside the box), shared or per-point time."""
    from oracle import hexplane as OH
    reso, mr = [64, 64, 64, 25], [1, 2]
    grids = OH.hash_planes(reso, mr, salt=9)
    aabb = torch.tensor([[40.0, 20.0, 10.0], [-30.0, -25.0, -4.0]])
    g = torch.Generator().manual_seed(5)
    N = 20000
    lo, hi = aabb.min(0).values, aabb.max(0).values
    pts = lo + (hi - lo) * (torch.rand(N, 3, generator=g) * 1.3 - 0.15)
    t = torch.full((N, 1), 0.61) if shared_t else torch.rand(N, 1, generator=g) * 2.4 - 1.2
    cot = torch.randn(N, 64, generator=g)
    feat_o, vp_o, vt_o, vg_o = oracle_run(grids, aabb, pts, t, cot)
    f = _field(reso, mr, grids, aabb)
    pg = pts.cuda().requires_grad_(True)
    tg = (torch.tensor([0.61]) if shared_t else t).cuda().requires_grad_(True)
    feat = f(pg, tg)
    # features are O(0.05); the kernel blends with FMAs, ATen with separate products: a few ulp of the 6-plane product
    assert float((feat.detach().cpu() - feat_o).abs().max()) <= 3e-6
    (feat * cot.cuda()).sum().backward()
    assert rel_err(pg.grad, vp_o) <= 1e-4
    if shared_t:
        assert abs(tg.grad.item() - vt_o.sum().item()) <= 1e-4 * max(1.0, abs(vt_o.sum().item()))
    else:
        assert rel_err(tg.grad, vt_o) <= 1e-4
    for row_g, row_o in zip(f.reference_grids(f.planes.grad), vg_o):
        for a, b in zip(row_g, row_o):
            assert rel_err(a, b) <= 1e-4 and rel_l2(a, b) <= 1e-5


def test_hexplane_full_size_properties():
    """BASELINE configs[2] size: resolution [64,64,64,25] x multires [1,2,4,8] (135 MB of planes), 1 M points.
    Size-independent properties: constant planes c_p give feature prod(c_p) everywhere; bilinear weights sum to one,
    so each plane's gradient sums to sum(cot) * prod_{q != p} c_q; the coordinate gradients vanish."""
    from emd_b200.hexplane import HexPlaneField
    cfg = {"grid_dimensions": 2, "input_coordinate_dim": 4, "output_coordinate_dim": 32, "resolution": [64, 64, 64, 25]}
    f = HexPlaneField(1.6, cfg, [1, 2, 4, 8]).cuda()
    consts = [0.5, 1.5, 0.75, 2.0, 1.25, 0.8]
    with torch.no_grad():
        for s in range(4):
            for p in range(6):
                f.plane_view(s, p).fill_(consts[p] + 0.01 * s)
    g = torch.Generator().manual_seed(1)
    N = 1_000_000
    pts = ((torch.rand(N, 3, generator=g) - 0.5) * 3.6).cuda().requires_grad_(True)
    t = torch.tensor([0.3], device="cuda", requires_grad=True)
    feat = f(pts, t)
    assert feat.shape == (N, 128)
    for s in range(4):
        want = float(np.prod([c + 0.01 * s for c in consts]))
        blk = feat[:, s * 32:(s + 1) * 32]
        assert float((blk - want).abs().max()) <= 2e-6 * want
    cot = torch.randn(N, 128, generator=g).cuda()
    (feat * cot).sum().backward()
    assert float(pts.grad.abs().max()) <= 1e-3 and abs(t.grad.item()) <= 1.0   # exact zero up to rounding of c - c
    for s in range(4):
        tot = cot[:, s * 32:(s + 1) * 32].double().sum().item()
        mag = cot[:, s * 32:(s + 1) * 32].double().abs().sum().item()
        for p in range(6):
            want = tot * float(np.prod([c + 0.01 * s for q, c in enumerate(consts) if q != p]))
            got = f.plane_view(s, p, f.planes.grad).double().sum().item()
            assert abs(got - want) <= 1e-5 * mag, (s, p, got, want)


def test_deformation_network_with_hexplane():
    """S3GDeformation querying K1e itself (hex_feat=None) against the oracle evaluating the same planes: forward values
    and the gradients that reach the planes, the points and time_offset through the HexPlane."""
    from emd_b200.emd_s3g import S3GDeformation
    from oracle import emd_s3g as S
    from oracle import hexplane as OH
    z = np.load(f"{G}/emd_s3g.npz")
    names = [k[len("w.deformation_net."):] for k in z.files if k.startswith("w.deformation_net.")]
    used = [n for n in names if not any(s in n for s in ("scales_deform", "rotations_deform"))]
    reso, mr = [16, 16, 16, 25], [1, 2, 4, 8]
    grids = OH.hash_planes(reso, mr, salt=2)
    aabb = torch.tensor([[1.6, 1.6, 1.6], [-1.6, -1.6, -1.6]])
    tm, it, cam = 0.8, 17000, 1
    w_c = {n: _t(z["w.deformation_net." + n]).clone().requires_grad_(True) for n in used}
    inp_c = {k: _t(z[k]).clone().requires_grad_(True) for k in ("point", "opacity", "shs", "embeddings")}
    G_c = [[x.clone().requires_grad_(True) for x in row] for row in grids]
    means, opac, shs, dd = S.deform(w_c, inp_c["point"], inp_c["opacity"], inp_c["shs"], inp_c["embeddings"], None,
                                    float(np.float32(tm)), it, cam, grids=G_c, aabb=aabb)
    dev = "cuda"
    f = _field(reso, mr, grids, aabb)
    w_g = {n: v.detach().to(dev).requires_grad_(True) for n, v in w_c.items()}
    inp_g = {k: v.detach().to(dev).requires_grad_(True) for k, v in inp_c.items()}
    net = S3GDeformation(w_g, hexplane=f)
    gm, _, _, go, gsh, gdd = net(inp_g["point"], _t(z["scales"]).to(dev), _t(z["rotations"]).to(dev), inp_g["opacity"],
                                 inp_g["shs"], float(np.float32(tm)), inp_g["embeddings"], it, cam)
    for a, b in ((gm, means), (go, opac), (gsh, shs)):
        assert float((a.detach().cpu() - b.detach()).abs().max()) <= 1e-5
    g = torch.Generator().manual_seed(3)
    cm, co, cs = (torch.randn(x.shape, generator=g) for x in (means, opac, shs))
    ((means * cm).sum() + (opac * co).sum() + (shs * cs).sum()).backward()
    ((gm * cm.to(dev)).sum() + (go * co.to(dev)).sum() + (gsh * cs.to(dev)).sum()).backward()
    assert rel_err(inp_g["point"].grad, inp_c["point"].grad) <= 1e-3
    assert rel_err(w_g["time_offset"].grad, w_c["time_offset"].grad) <= 1e-3
    assert rel_err(w_g["feature_out.0.weight"].grad, w_c["feature_out.0.weight"].grad) <= 1e-3
    for row_g, row_o in zip(f.reference_grids(f.planes.grad), G_c):
        for a, b in zip(row_g, row_o):
            assert rel_l2(a, b.grad) <= 1e-3


def test_hexplane_tail_columns_equal_concatenation():
    """``get_density(..., tail=emb)`` (row-pitched gather writing the MLP input [features | embedding] in place,
    deformation.py:205) is bit-identical to gather + torch.cat, forward and backward."""
    from oracle import hexplane as OH
    reso, mr = [16, 16, 16, 8], [1, 2, 4, 8]
    grids = OH.hash_planes(reso, mr, salt=3)
    aabb = torch.tensor([[9.0, 9.0, 9.0], [-9.0, -9.0, -9.0]])
    g = torch.Generator().manual_seed(11)
    N = 5003
    pts = (torch.rand(N, 3, generator=g) * 20 - 10)
    emb = torch.randn(N, 4, generator=g)
    cot = torch.randn(N, 132, generator=g).cuda()
    outs = []
    for tail in (False, True):
        f = _field(reso, mr, grids, aabb)
        pg, eg = pts.cuda().requires_grad_(True), emb.cuda().requires_grad_(True)
        tg = torch.tensor([0.3]).cuda().requires_grad_(True)
        x = f.get_density(pg, tg, tail=eg) if tail else torch.cat([f(pg, tg), eg], dim=-1)
        assert x.shape == (N, 132) and x.is_contiguous()
        (x * cot).sum().backward()
        outs.append((x.detach(), pg.grad, eg.grad, tg.grad))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    # the plane gradient is scattered with floating-point atomics: equal up to the order of the additions
    # (checked against the oracle in test_hexplane_vs_oracle); a pitch that is not a multiple of 4 falls back to cat
    f = _field(reso, mr, grids, aabb)
    x3 = f.get_density(pts.cuda(), torch.tensor([0.3]).cuda(), tail=emb[:, :3].cuda())
    assert x3.shape == (N, 131) and torch.equal(x3[:, :128], outs[0][0][:, :128])
