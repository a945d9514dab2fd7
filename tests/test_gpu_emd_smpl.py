"""GPU parity: EMD SMPL deformation (K1c) vs the oracle."""
import pytest
import torch

from tests.util import rel_err, rel_l2

pytestmark = pytest.mark.gpu

HEADS = ("smpl_c_w", "smpl_c_b", "smpl_f_w", "smpl_f_b")
GRAD = ["means", "quats", "scales", "opacities", "features_dc", "features_rest", "embeddings", "weight",
        "instances_quats", "smpl_qauts", "instances_trans"]


def _setup(seed, I, V, frames):
    from emd_b200 import scenes
    from oracle import emd_smpl as ES
    g = torch.Generator().manual_seed(seed)
    ss = scenes.smpl_nodes(I, g, V=V, num_frames=frames)
    cpu = {k: getattr(ss, k).clone().requires_grad_(True) for k in GRAD}
    cpu.update({k: ss.track[k].clone().requires_grad_(True) for k in HEADS})
    p = ES.SMPLEMD(point_ids=ss.point_ids[:, 0], embeddings=cpu["embeddings"], weight=cpu["weight"],
                   instances_quats=cpu["instances_quats"], smpl_quats=cpu["smpl_qauts"],
                   instances_trans=cpu["instances_trans"], instances_fv=ss.instances_fv,
                   J_canonical=ss.J_canonical, A0_inv=ss.A0_inv, W=ss.W, **{k: cpu[k] for k in HEADS})
    return ss, cpu, p, g


def _node(ss, gpu, dev):
    from emd_b200.emd_smpl import SMPLNodesEMD
    return SMPLNodesEMD(
        dict(_means=gpu["means"], _quats=gpu["quats"], _scales=gpu["scales"], _opacities=gpu["opacities"],
             _features_dc=gpu["features_dc"], _features_rest=gpu["features_rest"], _embeddings=gpu["embeddings"],
             point_ids=ss.point_ids.to(dev), weight=gpu["weight"], instances_quats=gpu["instances_quats"],
             smpl_qauts=gpu["smpl_qauts"], instances_trans=gpu["instances_trans"], instances_fv=ss.instances_fv.to(dev)),
        {k: gpu[k] for k in HEADS},
        dict(J_canonical=ss.J_canonical.to(dev), A0_inv=ss.A0_inv.to(dev), W=ss.W.to(dev)))


@pytest.mark.parametrize("seed,I,V,frames,frame,step", [
    (0, 3, 700, 40, 13, 4000),
    (1, 5, 1500, 60, 59, 20000),   # > 1 chunk per instance
    (2, 2, 300, 10, 0, 0),
])
def test_smpl_get_gaussians(seed, I, V, frames, frame, step):
    from oracle import emd_smpl as ES
    ss, cpu, p, g = _setup(seed, I, V, frames)
    ss.instances_fv[frame, 0] = True
    if I > 2:
        ss.instances_fv[frame, 1] = False  # an invisible instance: identity quats, means = trans
    cam_pos = torch.tensor([0.0, 0.0, 1.6])
    ref = ES.get_gaussians(p, cpu["means"], cpu["quats"], cpu["scales"], cpu["opacities"], cpu["features_dc"],
                           cpu["features_rest"], frame, step, cam_pos)
    dev = "cuda"
    gpu = {k: v.detach().to(dev).requires_grad_(True) for k, v in cpu.items()}
    out = _node(ss, gpu, dev).get_gaussians(cam_pos.tolist(), frame, step)
    cot = {}
    for k in ("_means", "_opacities", "_rgbs", "_scales", "_quats"):
        assert out[k].shape == ref[k].shape, k
        err = float((out[k].detach().cpu() - ref[k].detach()).abs().max())
        tol = 3e-5 * max(1.0, float(ref[k].detach().abs().max()))
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


def test_smpl_all_invisible_returns_none():
    ss, cpu, p, g = _setup(4, 2, 200, 8)
    ss.instances_fv[3] = False
    dev = "cuda"
    gpu = {k: v.detach().to(dev) for k, v in cpu.items()}
    assert _node(ss, gpu, dev).get_gaussians([0.0, 0.0, 1.6], 3, 100) is None
