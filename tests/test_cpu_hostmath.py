"""CPU: the product's host-compilable math headers (proj_math.cuh, emd_math.cuh, built
with g++ -ffp-contract=off) against the oracle.  Catches op-order and VJP mistakes
without a GPU; the same headers are what the CUDA kernels include."""
import ctypes

import numpy as np
import pytest
import torch

from emd_b200 import build, scenes
from oracle import emd_rigid as ER
from oracle import gsplat_ref as G

P = ctypes.c_void_p


@pytest.fixture(scope="module")
def hostlib():
    return ctypes.CDLL(str(build.build_hostmath()))


def _fp(t):
    return t.detach().contiguous().numpy().ctypes.data_as(P)


@pytest.mark.parametrize("seed,yaws,near,clip", [(0, (0.0, 45.0, -45.0), 0.1, 0.0), (1, (10.0,), 0.01, 4.0)])
def test_projection_bit_exact(hostlib, seed, yaws, near, clip):
    g = torch.Generator().manual_seed(seed)
    W, H, N = 960, 640, 60000
    sc = scenes.simple_gaussians(N, g, W, H, depth=(0.05, 60.0))
    viewmats, Ks, _ = scenes.cameras(yaws, W, H)
    C = len(yaws)
    radii, m2d, depths, conics, comps = G.projection(sc["means"], sc["quats"], sc["scales"], viewmats, Ks, W, H, 0.3,
                                                     near, 1e10, clip)
    r2 = np.zeros((C, N), np.int32); m2 = np.zeros((C, N, 2), np.float32); d2 = np.zeros((C, N), np.float32)
    c2 = np.zeros((C, N, 3), np.float32); cp = np.zeros((C, N), np.float32); tp = np.zeros((C, N), np.int32)
    rc = np.zeros((C, N, 4), np.int32)
    f = hostlib.emd_host_projection_fwd
    f.argtypes = [P] * 5 + [ctypes.c_int64] * 2 + [ctypes.c_int] * 2 + [ctypes.c_float] * 4 + [ctypes.c_int] * 2 + [P] * 7
    f(_fp(sc["means"]), _fp(sc["quats"]), _fp(sc["scales"]), _fp(viewmats), _fp(Ks), N, C, W, H, 0.3, near, 1e10, clip,
      60, 40, *[a.ctypes.data_as(P) for a in (r2, m2, d2, c2, cp, tp, rc)])
    assert (radii > 0).sum() > N // 4
    assert np.array_equal(r2, radii.numpy())
    assert np.array_equal(m2.view(np.int32), m2d.numpy().view(np.int32))
    assert np.array_equal(d2.view(np.int32), depths.numpy().view(np.int32))
    assert np.array_equal(c2.view(np.int32), conics.numpy().view(np.int32))
    tpg, ids, flat, bits = G.isect_tiles(m2d, radii, depths, 16, 60, 40)
    assert np.array_equal(tp, tpg.numpy())
    x0, y0, x1, y1 = G.tile_rects(m2d, radii, 16, 60, 40)
    assert np.array_equal(rc, torch.stack([x0, y0, x1, y1], -1).numpy().astype(np.int32))


def test_projection_vjp(hostlib):
    g = torch.Generator().manual_seed(1)
    W, H, N, C = 960, 640, 8000, 2
    sc = scenes.simple_gaussians(N, g, W, H)
    viewmats, Ks, _ = scenes.cameras((0.0, 30.0), W, H)
    means = sc["means"].clone().requires_grad_(); quats = sc["quats"].clone().requires_grad_()
    scales = sc["scales"].clone().requires_grad_()
    radii, m2d, depths, conics, comps = G.projection(means, quats, scales, viewmats, Ks, W, H, 0.3, 0.1, 1e10, 0.0)
    vm = torch.randn(C, N, 2, generator=g); vd = torch.randn(C, N, generator=g); vc = torch.randn(C, N, 3, generator=g)
    ((m2d * vm).sum() + (depths * vd).sum() + (conics * vc).sum()).backward()
    gm = np.zeros((N, 3), np.float32); gq = np.zeros((N, 4), np.float32); gs = np.zeros((N, 3), np.float32)
    f = hostlib.emd_host_projection_bwd
    f.argtypes = [P] * 5 + [ctypes.c_int64] * 2 + [ctypes.c_int] * 2 + [ctypes.c_float] * 4 + [P] * 6
    f(_fp(means), _fp(quats), _fp(scales), _fp(viewmats), _fp(Ks), N, C, W, H, 0.3, 0.1, 1e10, 0.0, _fp(vm), _fp(vd),
      _fp(vc), *[a.ctypes.data_as(P) for a in (gm, gq, gs)])
    for got, ref in ((gm, means.grad), (gq, quats.grad), (gs, scales.grad)):
        got = torch.from_numpy(got)
        assert float((got - ref).abs().max() / ref.abs().max()) < 1e-4


def _rigid_problem(seed, I=5, step=7000, frame=37, empty_instance=False):
    g = torch.Generator().manual_seed(seed)
    rs = scenes.rigid_nodes(I, 40, g, num_frames=60)
    if empty_instance:
        rs.point_ids[rs.point_ids == 2] = 1  # instance 2 owns no points -> NaN mean -> offsets skipped
    p = ER.RigidEMD(point_ids=rs.point_ids[:, 0], embeddings=rs.embeddings, weight=rs.weight,
                    instances_quats=rs.instances_quats + 0.05 * torch.randn(rs.instances_quats.shape, generator=g),
                    instances_trans=rs.instances_trans, instances_fv=rs.instances_fv,
                    **{k: v for k, v in rs.track.items()})
    return rs, p, frame, step


@pytest.mark.parametrize("seed,step,empty", [(0, 0, False), (1, 7000, False), (2, 20000, True), (3, 31000, False)])
def test_rigid_instance_heads(hostlib, seed, step, empty):
    """emd_math.cuh per-instance forward + VJP vs the oracle's restatement of rigid.py:150-246."""
    rs, p, frame, step = _rigid_problem(seed, step=step, empty_instance=empty)
    I, E, d = p.weight.shape
    gdim = p.embeddings.shape[1]
    leaves = dict(weight=p.weight, iq=p.instances_quats, it=p.instances_trans, emb=p.embeddings,
                  **{k: getattr(p, k) for k in ("rot_c_w", "rot_c_b", "rot_f_w", "rot_f_b", "trans_c_w", "trans_c_b",
                                                "trans_f_w", "trans_f_b")})
    for v in leaves.values():
        v.requires_grad_(True)
    # oracle: per-instance pose (R, t, Q)
    from oracle.quat import quat_act, quat_mult, quat_to_rotmat
    dtrans, qoff, ok_t, ok_q = ER.track_offsets(p, frame, step)
    assert bool(ok_t.all()) == (not empty)
    R = quat_to_rotmat(quat_act(p.instances_quats[frame]))
    tt = p.instances_trans[frame] + dtrans
    Qg = torch.where(ok_q[:, None], quat_mult(p.instances_quats[frame], qoff), p.instances_quats[frame])
    Q = quat_act(Qg)
    ref = torch.cat([R.reshape(I, 9), tt, Q], dim=1)
    v = torch.randn(I, 16, generator=torch.Generator().manual_seed(seed + 100))
    (ref * v).sum().backward()

    mean_emb = torch.stack([p.embeddings[p.point_ids == i].mean(0) for i in range(I)]).detach()
    t = np.float32(frame / (p.num_frames - 1))
    cur_c, cur_f = 30, ER.int_lininterp(step, 30, 150, 20000)
    heads = [getattr(p, k).detach().contiguous() for k in ("rot_c_w", "rot_c_b", "rot_f_w", "rot_f_b", "trans_c_w",
                                                           "trans_c_b", "trans_f_w", "trans_f_b")]
    harr = (P * 8)(*[h.numpy().ctypes.data_as(P) for h in heads])
    out = np.zeros((I, 16), np.float32)
    pq = p.instances_quats[frame].detach().contiguous(); pt = p.instances_trans[frame].detach().contiguous()
    f = hostlib.emd_host_rigid_instance_fwd
    f.argtypes = [P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, P, ctypes.c_float, ctypes.c_int,
                  ctypes.c_int, P, P, P, P, P]
    f(_fp(p.weight), I, E, d, gdim, _fp(mean_emb), t, cur_c, cur_f, harr, _fp(pq), _fp(pq), _fp(pt),
      out.ctypes.data_as(P))
    assert np.allclose(out, ref.detach().numpy(), rtol=2e-5, atol=2e-6), np.abs(out - ref.detach().numpy()).max()

    pc = hostlib.emd_host_rigid_param_count(d, gdim)
    v_pqm = np.zeros((I, 4), np.float32); v_pqq = np.zeros((I, 4), np.float32); v_pt = np.zeros((I, 3), np.float32)
    v_params = np.zeros(pc, np.float32); v_table = np.zeros((I, E, d), np.float32); v_me = np.zeros((I, gdim), np.float32)
    b = hostlib.emd_host_rigid_instance_bwd
    b.argtypes = [P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, P, ctypes.c_float, ctypes.c_int,
                  ctypes.c_int, P, P, P, P, P, P, P, P, P, P, P]
    b(_fp(p.weight), I, E, d, gdim, _fp(mean_emb), t, cur_c, cur_f, harr, _fp(pq), _fp(pq), _fp(pt), _fp(v),
      *[a.ctypes.data_as(P) for a in (v_pqm, v_pqq, v_pt, v_params, v_table, v_me)])

    def close(a, b_, name):
        b_ = b_.detach().numpy()
        scale = max(np.abs(b_).max(), 1e-12)
        assert np.abs(a - b_).max() <= 2e-4 * scale, f"{name}: {np.abs(a - b_).max()} vs scale {scale}"

    close(v_pqm + v_pqq, p.instances_quats.grad[frame], "instances_quats")
    close(v_pt, p.instances_trans.grad[frame], "instances_trans")
    close(v_table, p.weight.grad, "weight")
    off = 0
    for k in ("rot_c_w", "rot_c_b", "rot_f_w", "rot_f_b", "trans_c_w", "trans_c_b", "trans_f_w", "trans_f_b"):
        ref_g = getattr(p, k).grad
        n = ref_g.numel()
        close(v_params[off:off + n].reshape(ref_g.shape), ref_g, k)
        off += n
    # mean-embedding gradient, spread over the instance's points
    cnt = torch.stack([(p.point_ids == i).sum() for i in range(I)]).float()
    per_point = torch.nan_to_num(torch.from_numpy(v_me))[p.point_ids] / cnt[p.point_ids][:, None]
    close(per_point.numpy(), p.embeddings.grad, "embeddings")


def test_dg_preprocess_bit_exact(hostlib):
    """dg_math.cuh (diff_gauss front end) vs oracle/diff_gauss_ref.preprocess: radii, means, depth bits, tile rects."""
    from oracle import diff_gauss_ref as DG
    from tests.dg_util import s3g_camera
    g = torch.Generator().manual_seed(4)
    W, H, N = 960, 640, 50000
    sc = scenes.simple_gaussians(N, g, W, H, depth=(0.05, 60.0))
    rots = (sc["quats"] / sc["quats"].norm(dim=-1, keepdim=True)).contiguous()
    cam = s3g_camera(15.0, W, H)
    s = DG.Settings(H, W, cam["tanfovx"], cam["tanfovy"], torch.zeros(3), 0.9, cam["viewmatrix"], cam["projmatrix"], 0,
                    cam["campos"])
    radii, m2d, depths, conics, rect = DG.preprocess(sc["means"], sc["scales"], rots, s)
    r2 = np.zeros(N, np.int32); m2 = np.zeros((N, 2), np.float32); d2 = np.zeros(N, np.float32)
    c2 = np.zeros((N, 3), np.float32); rc = np.zeros((N, 4), np.int32)
    f = hostlib.emd_host_dg_preprocess_fwd
    f.argtypes = [P] * 5 + [ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int64] + [P] * 5
    f(_fp(sc["means"]), _fp(sc["scales"]), _fp(rots), _fp(cam["viewmatrix"]), _fp(cam["projmatrix"]), cam["tanfovx"],
      cam["tanfovy"], W, H, 0.9, N, *[a.ctypes.data_as(P) for a in (r2, m2, d2, c2, rc)])
    assert (radii > 0).sum() > N // 4
    assert np.array_equal(r2, radii.numpy())
    assert np.array_equal(m2.view(np.int32), m2d.numpy().view(np.int32))
    assert np.array_equal(d2.view(np.int32), depths.numpy().view(np.int32))
    assert np.array_equal(c2.view(np.int32), conics.numpy().view(np.int32))
    assert np.array_equal(rc, torch.stack(rect, -1).numpy().astype(np.int32))


@pytest.mark.parametrize("per_point_t", [False, True])
def test_hexplane_host_math_matches_oracle(hostlib, per_point_t):
    """hexplane_math.cuh (the tap arithmetic the CUDA kernels include) against the oracle: features, plane /
    point / time gradients, with points outside the box, on grid nodes and on the box corners."""
    from oracle import hexplane as OH
    from tests.hex_util import flatten_planes, oracle_run, unflatten_like
    reso, mr = [8, 6, 10, 5], [1, 2, 4]
    grids = OH.hash_planes(reso, mr, salt=5)
    aabb = torch.tensor([[3.0, 2.0, 1.5], [-2.0, -2.5, -0.5]])
    g = torch.Generator().manual_seed(17)
    N = 400
    lo, hi = aabb.min(0).values, aabb.max(0).values
    pts = lo + (hi - lo) * (torch.rand(N, 3, generator=g) * 1.3 - 0.15)
    pts[0], pts[1], pts[2] = hi, lo, 0.5 * (lo + hi)
    t = torch.rand(N, 1, generator=g) * 2.4 - 1.2 if per_point_t else torch.full((N, 1), 0.37)
    cot = torch.randn(N, 32 * len(mr), generator=g)
    feat_o, vp_o, vt_o, vg_o = oracle_run(grids, aabb, pts, t, cot)
    flat, offsets, rs = flatten_planes(grids)
    feat = np.zeros((N, 32 * len(mr)), np.float32); v_planes = np.zeros(flat.numel(), np.float32)
    v_pts = np.zeros((N, 3), np.float32); v_t = np.zeros(N if per_point_t else 1, np.float32)
    t_in = t.reshape(-1).contiguous() if per_point_t else t[:1].reshape(-1).contiguous()
    f = hostlib.emd_host_hexplane
    f.argtypes = [P, P, P, ctypes.c_int, P, P, P, ctypes.c_int, ctypes.c_int64] + [P] * 5
    f.restype = None
    f(_fp(flat), offsets.ctypes.data_as(P), rs.ctypes.data_as(P), len(mr), _fp(aabb), _fp(pts), _fp(t_in),
      1 if per_point_t else 0, N, feat.ctypes.data_as(P), _fp(cot), v_planes.ctypes.data_as(P), v_pts.ctypes.data_as(P),
      v_t.ctypes.data_as(P))
    assert np.abs(feat - feat_o.numpy()).max() <= 1e-6
    assert np.abs(v_pts - vp_o.numpy()).max() <= 2e-5 * max(1.0, vp_o.abs().max().item())
    if per_point_t:
        assert np.abs(v_t - vt_o.reshape(-1).numpy()).max() <= 2e-5 * max(1.0, vt_o.abs().max().item())
    else:
        assert abs(v_t[0] - vt_o.sum().item()) <= 1e-4 * max(1.0, abs(vt_o.sum().item()))
    got = unflatten_like(torch.from_numpy(v_planes), grids, offsets)
    for row_g, row_o in zip(got, vg_o):
        for a, b in zip(row_g, row_o):
            assert (a - b).abs().max().item() <= 5e-6 * max(1.0, b.abs().max().item())


@pytest.mark.parametrize("wd,scale", [(0.0, 1.0), (0.01, 0.125)])
def test_adam_host_math_matches_torch(hostlib, wd, scale):
    """adam_math.cuh (what emd_adam_step runs per element) against torch.optim.Adam as the reference configures it
    (eps=1e-15, lr written per step by a scheduler) over 25 steps, including all-zero gradients (invisible Gaussians)."""
    g = torch.Generator().manual_seed(3)
    n = 5003
    p0 = torch.randn(n, generator=g)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=0.0, eps=1e-15, weight_decay=wd, foreach=False)
    p = p0.numpy().copy(); m = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
    f = hostlib.emd_host_adam_step
    f.argtypes = [P, P, P, P, ctypes.c_int64] + [ctypes.c_double] * 5 + [ctypes.c_int64, ctypes.c_double]
    f.restype = None
    for step in range(1, 26):
        grad = torch.randn(n, generator=g) * (10.0 ** float(torch.randint(-4, 2, (1,), generator=g)))
        grad[::7] = 0.0
        if step > 5:
            grad[:500] = 0.0
        lr = 1.6e-4 * 0.97 ** step
        opt.param_groups[0]["lr"] = lr
        ref.grad = (grad * scale).clone()
        opt.step()
        gn = grad.numpy().copy()
        f(p.ctypes.data_as(P), gn.ctypes.data_as(P), m.ctypes.data_as(P), v.ctypes.data_as(P), n, lr, 0.9, 0.999, 1e-15, wd,
          step, scale)
        st = opt.state[ref]
        assert np.abs(m - st["exp_avg"].numpy()).max() <= 2e-6 * max(1e-30, float(st["exp_avg"].abs().max()))
        assert np.abs(v - st["exp_avg_sq"].numpy()).max() <= 2e-6 * max(1e-30, float(st["exp_avg_sq"].abs().max()))
        assert np.abs(p - ref.detach().numpy()).max() <= 5e-7 * step


def test_voxel_lbs_host_math_matches_oracle(hostlib):
    """voxel_math.cuh (the tap arithmetic the voxel-LBS kernels include) against the oracle on the golden case: weights,
    correction-volume gradient, point gradient; channel-last layout in, reference layout compared."""
    import os
    from oracle import voxel_deformer as OV
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "omnire_modules.npz"))
    base, corr = torch.from_numpy(z["vox_base"]), torch.from_numpy(z["vox_corr"])
    B, J, D, H, W = base.shape
    xc, cot = torch.from_numpy(z["vox_xc"]), torch.from_numpy(z["vox_cot"])
    V = xc.shape[1]
    cl = lambda v: v.permute(0, 2, 3, 4, 1).contiguous()  # noqa: E731
    ratio_dim = int(z["vox_ratio_dim"]) % 3
    out = np.zeros((B, V, J), np.float32); v_corr = np.zeros((B, D, H, W, J), np.float32); v_xc = np.zeros((B, V, 3), np.float32)
    f = hostlib.emd_host_voxel_lbs
    f.argtypes = [P] * 4 + [ctypes.c_float] + [ctypes.c_int] * 6 + [P, ctypes.c_int64] + [P] * 4
    f.restype = None
    off, scl = torch.from_numpy(z["vox_offset"]).reshape(B, 3), torch.from_numpy(z["vox_scale"]).reshape(B)
    f(_fp(cl(base)), _fp(cl(corr)), _fp(off), _fp(scl), float(z["vox_ratio"]), ratio_dim, B, D, H, W, J, _fp(xc), V,
      out.ctypes.data_as(P), _fp(cot), v_corr.ctypes.data_as(P), v_xc.ctypes.data_as(P))
    assert np.abs(out - z["vox_w"]).max() <= 2e-6
    assert np.abs(v_corr.transpose(0, 4, 1, 2, 3) - z["vox_v_corr"]).max() <= 2e-6
    assert np.abs(v_xc - z["vox_v_xc"]).max() <= 2e-5 * max(1.0, np.abs(z["vox_v_xc"]).max())
    # and against the oracle on a second volume shape with the stretch on another axis
    g = torch.Generator().manual_seed(5)
    B, J, D, H, W, V = 2, 8, 6, 3, 5, 64
    base, corr = torch.randn(B, J, D, H, W, generator=g), torch.randn(B, J, D, H, W, generator=g).requires_grad_(True)
    off, scl = torch.randn(B, 1, 3, generator=g) * 0.2, torch.rand(B, 1, 1, generator=g) + 0.5
    xc = (torch.rand(B, V, 3, generator=g) * 2.6 - 1.3).requires_grad_(True)
    cot = torch.randn(B, V, J, generator=g)
    w = OV.voxel_weights(base + corr, off, scl, 2.0, -2, xc)
    (w * cot).sum().backward()
    out = np.zeros((B, V, J), np.float32); v_corr = np.zeros((B, D, H, W, J), np.float32); v_xc = np.zeros((B, V, 3), np.float32)
    f(_fp(cl(base)), _fp(cl(corr)), _fp(off.reshape(B, 3)), _fp(scl.reshape(B)), 2.0, 1, B, D, H, W, J, _fp(xc), V,
      out.ctypes.data_as(P), _fp(cot), v_corr.ctypes.data_as(P), v_xc.ctypes.data_as(P))
    assert np.abs(out - w.detach().numpy()).max() <= 2e-6
    assert np.abs(v_corr.transpose(0, 4, 1, 2, 3) - corr.grad.numpy()).max() <= 5e-6
    assert np.abs(v_xc - xc.grad.numpy()).max() <= 2e-5 * max(1.0, xc.grad.abs().max().item())


@pytest.mark.parametrize("M,K,Nout,ldx,ldy,relu", [(300, 100, 256, 100, 256, 1), (517, 356, 256, 356, 356, 1),
                                                    (130, 256, 7, 256, 7, 0), (1, 20, 33, 24, 40, 1)])
def test_dense_tile_logic_matches_torch(hostlib, M, K, Nout, ldx, ldy, relu):
    """dense_math.cuh (the per-thread tile logic of sgemm_kernel, run thread by thread on the host) against torch:
    forward with strided operands, data gradient on a column window with the producer's ReLU mask, split weight gradient,
    bias gradient."""
    g = torch.Generator().manual_seed(M + K)
    Xb = torch.randn(M, ldx, generator=g)
    Wt, b = torch.randn(Nout, K, generator=g) / K ** 0.5, torch.randn(Nout, generator=g)
    Yb = np.full((M, ldy), 123.0, np.float32)
    I64, I32 = ctypes.c_int64, ctypes.c_int
    f = hostlib.emd_host_dense_fwd
    f.argtypes = [P, I64, P, P, I64, I32, I32, I32, P, I64]
    f.restype = None
    off = ldy - Nout                                   # write into the LAST Nout columns of a wider buffer
    f(_fp(Xb), ldx, _fp(Wt), _fp(b), M, K, Nout, relu, ctypes.c_void_p(Yb.ctypes.data + 4 * off), ldy)
    ref = Xb[:, :K] @ Wt.T + b
    ref = torch.relu(ref) if relu else ref
    assert np.abs(Yb[:, off:] - ref.numpy()).max() <= 2e-5
    assert (Yb[:, :off] == 123.0).all()                # nothing outside the window is touched
    # backward
    dZ = torch.randn(M, Nout, generator=g)
    col0, ncols = (K - 16, 16) if K >= 32 else (0, K)
    mask = torch.randn(M, ncols, generator=g)
    dX = np.full((M, ncols + 3), -5.0, np.float32)
    dW = np.zeros((Nout, K), np.float32); db = np.zeros(Nout, np.float32)
    fb = hostlib.emd_host_dense_bwd
    fb.argtypes = [P, I64, P, P, I64, I64, I32, I32, P, I64, I32, I32, P, I64, P, P]
    fb.restype = None
    fb(_fp(Xb), ldx, _fp(Wt), _fp(dZ), Nout, M, K, Nout, dX.ctypes.data_as(P), ncols + 3, col0, ncols, _fp(mask), ncols,
       dW.ctypes.data_as(P), db.ctypes.data_as(P))
    ref_dX = (dZ @ Wt[:, col0:col0 + ncols]) * (mask > 0)
    assert np.abs(dX[:, :ncols] - ref_dX.numpy()).max() <= 2e-5 * max(1.0, ref_dX.abs().max().item())
    assert (dX[:, ncols:] == -5.0).all()
    ref_dW = dZ.T @ Xb[:, :K]
    assert np.abs(dW - ref_dW.numpy()).max() <= 2e-5 * max(1.0, ref_dW.abs().max().item())
    assert np.abs(db - dZ.sum(0).numpy()).max() <= 2e-5 * max(1.0, dZ.sum(0).abs().max().item())


def test_smpl_weight_grad_host_math(hostlib):
    """emd_math.cuh:smpl_point_weight_grad (d loss / d LBS weights, used when W comes from the voxel deformer) against
    autograd through the oracle's skinning arithmetic."""
    from oracle.quat import matrix_to_quaternion, quat_act, quat_mult, quat_to_rotmat
    g = torch.Generator().manual_seed(23)
    N, J = 64, 24
    # near-rigid blend: joints = small rotations about a common pose, sparse-ish positive weights
    qj = quat_act(torch.tensor([1.0, 0.0, 0.0, 0.0]) + 0.3 * torch.randn(J, 4, generator=g))
    A = torch.cat([quat_to_rotmat(qj).reshape(J, 9), 0.2 * torch.randn(J, 3, generator=g)], dim=1)   # [J,12]
    W = torch.rand(N, J, generator=g) ** 4
    W[:, 5:9] = 0.0
    W = (W / W.sum(-1, keepdim=True)).requires_grad_(True)
    x, q = torch.randn(N, 3, generator=g), torch.randn(N, 4, generator=g)
    cg, cq = torch.randn(N, 3, generator=g), torch.randn(N, 4, generator=g)
    T = W @ A
    R, t = T[:, :9].reshape(N, 3, 3), T[:, 9:]
    xw = torch.einsum("nij,nj->ni", R, x) + t
    qw = quat_mult(quat_act(matrix_to_quaternion(R)), quat_act(q))
    ((xw * cg).sum() + (qw * cq).sum()).backward()
    out = np.zeros((N, J), np.float32)
    f = hostlib.emd_host_smpl_weight_grad
    f.argtypes = [P] * 6 + [ctypes.c_int64, P]
    f.restype = None
    f(_fp(W), _fp(A), _fp(x), _fp(q), _fp(cg), _fp(cq), N, out.ctypes.data_as(P))
    ref = W.grad.numpy()
    assert np.abs(out - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


def test_projection_tile_edge_cases_bit_exact(hostlib):
    """Adversarial placements: projected means EXACTLY on tile corners / edges (the optical axis lands on
    pixel (480, 320) = tile corner (30, 20); offsets of whole tiles along each axis), radii that end exactly on a tile
    edge, Gaussians at the image border and just outside it, at the near plane, and with a zero scale axis.  Radii,
    means2d, depths, conics, tile rectangles and tiles-per-Gaussian must stay bit-identical to the oracle."""
    W, H = 960, 640
    viewmats, Ks, _ = scenes.cameras((0.0,), W, H)
    fx, cx, cy = float(Ks[0, 0, 0]), float(Ks[0, 0, 2]), float(Ks[0, 1, 2])
    assert cx == 480.0 and cy == 320.0
    c2w = torch.linalg.inv(viewmats[0])
    pts_cam, scl = [], []
    for z in (0.05, -1.0, 0.1, 0.10000001, 1.0, 5.15, 20.0, 103.0):   # 5.15 = fx / 200, 103 = fx / 10: whole-pixel offsets below
        for du in (0.0, 16.0, -16.0, 160.0, 479.0, 480.0, -480.0, -481.0, 496.0, 3000.0):
            for dv in (0.0, 16.0, -320.0, 319.0, 321.0):
                pts_cam.append([du * z / fx, dv * z / fx, z])
                # sigma chosen so that 3 sigma in pixels is an integer number of tiles for some of them
                scl.append([16.0 / 3.0 * z / fx, 32.0 / 3.0 * z / fx, 0.0 if du == 160.0 else 1e-3])
    pc = torch.tensor(pts_cam, dtype=torch.float32)
    means = (c2w[:3, :3] @ pc.T).T + c2w[:3, 3]
    N = means.shape[0]
    quats = torch.zeros(N, 4); quats[:, 0] = 1.0
    quats[::3] = torch.tensor([0.9238795, 0.0, 0.0, 0.3826834])
    scales = torch.tensor(scl, dtype=torch.float32)
    radii, m2d, depths, conics, comps = G.projection(means, quats, scales, viewmats, Ks, W, H, 0.3, 0.1, 1e10, 0.0)
    r2 = np.zeros((1, N), np.int32); m2 = np.zeros((1, N, 2), np.float32); d2 = np.zeros((1, N), np.float32)
    c2 = np.zeros((1, N, 3), np.float32); cp = np.zeros((1, N), np.float32); tp = np.zeros((1, N), np.int32)
    rc = np.zeros((1, N, 4), np.int32)
    f = hostlib.emd_host_projection_fwd
    f.argtypes = [P] * 5 + [ctypes.c_int64] * 2 + [ctypes.c_int] * 2 + [ctypes.c_float] * 4 + [ctypes.c_int] * 2 + [P] * 7
    f(_fp(means), _fp(quats), _fp(scales), _fp(viewmats), _fp(Ks), N, 1, W, H, 0.3, 0.1, 1e10, 0.0, 60, 40,
      *[a.ctypes.data_as(P) for a in (r2, m2, d2, c2, cp, tp, rc)])
    vis = radii[0] > 0
    assert 0.3 * N < int(vis.sum()) < N                      # both culled and visible placements are present
    on_edge = vis & ((m2d[0, :, 0] % 16 == 0) | (m2d[0, :, 1] % 16 == 0))
    assert int(on_edge.sum()) >= 20                          # means that really sit on tile edges after fp32 projection
    assert np.array_equal(r2, radii.numpy())
    assert np.array_equal(m2.view(np.int32), m2d.numpy().view(np.int32))
    assert np.array_equal(d2.view(np.int32), depths.numpy().view(np.int32))
    assert np.array_equal(c2.view(np.int32), conics.numpy().view(np.int32))
    tpg, ids, flat, bits = G.isect_tiles(m2d, radii, depths, 16, 60, 40)
    assert np.array_equal(tp, tpg.numpy())
    x0, y0, x1, y1 = G.tile_rects(m2d, radii, 16, 60, 40)
    assert np.array_equal(rc, torch.stack([x0, y0, x1, y1], -1).numpy().astype(np.int32))


@pytest.mark.parametrize("which,npad", [(0, 128), (1, 256), (1, 16), (1, 112), (2, 256), (2, 16), (2, 48), (3, 128)])
def test_tensor_core_staging_maps(hostlib, which, npad):
    """tc_stage_math.cuh (the thread -> shared-memory maps of the experimental tcgen05 GEMM, csrc/deform_net_tc.cu): every
    thread's writes of one chunk replayed on the host must tile the operand exactly once, and reading it back the way
    the UMMA k-step descriptors walk the canonical K-major layout must give the source matrix (A chunk; W chunk in the
    forward orientation [n][k]; W chunk in the data-gradient orientation [k][n]; dZ chunk of the weight gradient [k][m])."""
    rows = 128 if which in (0, 3) else npad
    g = torch.Generator().manual_seed(which * 1000 + npad)
    src = torch.randn(rows, 32, generator=g) if which in (0, 1) else torch.randn(32, rows, generator=g)
    out = np.zeros((rows, 32), np.float32)
    miss = ctypes.c_int(-1)
    f = hostlib.emd_host_tc_stage_replay
    f.argtypes = [ctypes.c_int, ctypes.c_int, P, P, ctypes.POINTER(ctypes.c_int)]
    f.restype = ctypes.c_int
    twice = f(which, npad, _fp(src), out.ctypes.data_as(P), ctypes.byref(miss))
    assert twice == 0 and miss.value == 0
    want = src.numpy() if which in (0, 1) else src.numpy().T
    assert np.array_equal(out, want)
