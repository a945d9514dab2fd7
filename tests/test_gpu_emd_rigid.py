"""GPU parity: EMD rigid deformation + fused activation/SH (K1a, K1b) vs the oracle."""
import pytest
import torch

from tests.util import rel_err, rel_l2

pytestmark = pytest.mark.gpu

HEADS = ("rot_c_w", "rot_c_b", "rot_f_w", "rot_f_b", "trans_c_w", "trans_c_b", "trans_f_w", "trans_f_b")


def _setup(seed, I, pts, frames, empty=False):
    from emd_b200 import scenes
    from oracle import emd_rigid as ER
    g = torch.Generator().manual_seed(seed)
    rs = scenes.rigid_nodes(I, pts, g, num_frames=frames)
    rs.instances_quats = rs.instances_quats + 0.05 * torch.randn(rs.instances_quats.shape, generator=g)
    if empty:
        rs.point_ids[rs.point_ids == 1] = 0
    names = ["means", "quats", "scales", "opacities", "features_dc", "features_rest", "embeddings", "weight",
             "instances_quats", "instances_trans"]
    cpu = {k: getattr(rs, k).clone().requires_grad_(True) for k in names}
    cpu.update({k: rs.track[k].clone().requires_grad_(True) for k in HEADS})
    p = ER.RigidEMD(point_ids=rs.point_ids[:, 0], embeddings=cpu["embeddings"], weight=cpu["weight"],
                    instances_quats=cpu["instances_quats"], instances_trans=cpu["instances_trans"],
                    instances_fv=rs.instances_fv, **{k: cpu[k] for k in HEADS})
    return rs, cpu, p, g


@pytest.mark.parametrize("seed,I,pts,frames,frame,step,empty", [
    (0, 6, 300, 50, 17, 3500, False),
    (1, 3, 2500, 150, 149, 20000, False),    # > 1 chunk per instance, last frame
    (2, 4, 100, 20, 0, 0, True),             # instance without points (NaN-skip), degree 0, first frame
])
def test_rigid_get_gaussians(seed, I, pts, frames, frame, step, empty):
    from emd_b200.emd_rigid import RigidNodesEMD
    from oracle import emd_rigid as ER
    rs, cpu, p, g = _setup(seed, I, pts, frames, empty)
    cam_pos = torch.tensor([0.3, -0.2, 1.6])
    ref = ER.get_gaussians(p, cpu["means"], cpu["quats"], cpu["scales"], cpu["opacities"], cpu["features_dc"],
                           cpu["features_rest"], frame, step, cam_pos)
    dev = "cuda"
    gpu = {k: v.detach().to(dev).requires_grad_(True) for k, v in cpu.items()}
    node = RigidNodesEMD(
        dict(_means=gpu["means"], _quats=gpu["quats"], _scales=gpu["scales"], _opacities=gpu["opacities"],
             _features_dc=gpu["features_dc"], _features_rest=gpu["features_rest"], _embeddings=gpu["embeddings"],
             point_ids=rs.point_ids.to(dev), weight=gpu["weight"], instances_quats=gpu["instances_quats"],
             instances_trans=gpu["instances_trans"], instances_fv=rs.instances_fv.to(dev)),
        {k: gpu[k] for k in HEADS})
    out = node.get_gaussians(cam_pos.tolist(), frame, step)
    cot = {}
    for k in ("_means", "_opacities", "_rgbs", "_scales", "_quats"):
        assert out[k].shape == ref[k].shape, k
        err = float((out[k].detach().cpu() - ref[k].detach()).abs().max())
        tol = 2e-5 * max(1.0, float(ref[k].detach().abs().max()))
        assert err <= tol, f"{k}: {err} > {tol}"
        cot[k] = torch.randn(ref[k].shape, generator=g)
    sum((ref[k] * cot[k]).sum() for k in cot).backward()
    sum((out[k] * cot[k].to(dev)).sum() for k in cot).backward()
    for k in cpu:
        gr, gg = cpu[k].grad, gpu[k].grad
        assert gg is not None, k
        if float(gr.abs().max()) == 0.0:
            assert float(gg.abs().max()) == 0.0, k
            continue
        e, l2 = rel_err(gg, gr), rel_l2(gg, gr)
        assert e <= 1e-3 and l2 <= 1e-3, f"grad {k}: max-rel {e}, l2-rel {l2}"


def test_rigid_test_set_interpolation():
    from emd_b200.emd_rigid import RigidNodesEMD
    from oracle import emd_rigid as ER
    rs, cpu, p, g = _setup(5, 4, 200, 30)
    frame, step = 11, 12000
    wm = ER.transform_means(p, cpu["means"], frame, step, in_test_set=True)
    wq = ER.transform_quats(p, cpu["quats"], frame, step)
    dev = "cuda"
    gpu = {k: v.detach().to(dev) for k, v in cpu.items()}
    node = RigidNodesEMD(
        dict(_means=gpu["means"], _quats=gpu["quats"], _scales=gpu["scales"], _opacities=gpu["opacities"],
             _features_dc=gpu["features_dc"], _features_rest=gpu["features_rest"], _embeddings=gpu["embeddings"],
             point_ids=rs.point_ids.to(dev), weight=gpu["weight"], instances_quats=gpu["instances_quats"],
             instances_trans=gpu["instances_trans"], instances_fv=rs.instances_fv.to(dev)),
        {k: gpu[k] for k in HEADS})
    node.in_test_set = True
    gm, gq = node.transform_means_and_quats(frame, step)
    assert float((gm.cpu() - wm.detach()).abs().max()) <= 2e-5 * float(wm.detach().abs().max())
    assert float((gq.cpu() - wq.detach()).abs().max()) <= 2e-5


@pytest.mark.parametrize("deg,K", [(0, 1), (1, 4), (2, 9), (3, 16), (1, 16)])
def test_spherical_harmonics(deg, K):
    import emd_b200
    from oracle import sh as SH
    g = torch.Generator().manual_seed(deg * 10 + K)
    N = 5003
    dirs = torch.randn(N, 3, generator=g) * 3.0
    coeffs = torch.randn(N, K, 3, generator=g).requires_grad_(True)
    ref = SH.spherical_harmonics(deg, dirs, coeffs)
    cg = coeffs.detach().cuda().requires_grad_(True)
    out = emd_b200.spherical_harmonics(deg, dirs.cuda(), cg)
    assert float((out.detach().cpu() - ref.detach()).abs().max()) <= 1e-5 * max(1.0, float(ref.detach().abs().max()))
    v = torch.randn(N, 3, generator=g)
    (ref * v).sum().backward()
    (out * v.cuda()).sum().backward()
    assert rel_err(cg.grad, coeffs.grad) <= 1e-5


def test_background_activation():
    from emd_b200 import scenes
    from emd_b200.sh_ops import activate_gaussians
    from oracle import emd_rigid as ER
    g = torch.Generator().manual_seed(3)
    bg = scenes.background(7001, g)
    bg["features_dc"][:50] = -0.5 / 0.28209479177387814  # RGB 0 -> clamp boundary
    bg["features_rest"][:50] = 0.0
    cpu = {k: v.clone().requires_grad_(True) for k, v in bg.items()}
    cam = torch.tensor([0.0, 0.0, 1.6])
    for step in (0, 1500, 9000):
        ref = ER.background_get_gaussians(cpu["means"], cpu["quats"], cpu["scales"], cpu["opacities"],
                                          cpu["features_dc"], cpu["features_rest"], step, cam)
        gpu = {k: v.detach().cuda().requires_grad_(True) for k, v in cpu.items()}
        rgbs, opac, scales, quats = activate_gaussians(gpu["means"], gpu["features_dc"], gpu["features_rest"],
                                                       gpu["opacities"], gpu["scales"], gpu["quats"], cam.tolist(),
                                                       min(step // 1000, 3))
        out = dict(_rgbs=rgbs, _opacities=opac[:, None], _scales=scales, _quats=quats)
        cot = {k: torch.randn(ref[k].shape, generator=g) for k in out}
        for k in out:
            assert float((out[k].detach().cpu() - ref[k].detach()).abs().max()) <= 2e-5 * max(1.0, float(ref[k].detach().abs().max())), k
        for v in cpu.values():
            v.grad = None
        sum((ref[k] * cot[k]).sum() for k in out).backward()
        sum((out[k] * cot[k].cuda()).sum() for k in out).backward()
        for k in ("quats", "scales", "opacities", "features_dc", "features_rest"):
            assert rel_err(gpu[k].grad, cpu[k].grad) <= 1e-4, (k, step)
