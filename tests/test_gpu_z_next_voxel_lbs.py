"""GPU parity: voxel LBS-weight lookup (K1f) vs the golden vectors of the reference's own VoxelDeformer and
vs the oracle, alone and feeding the SMPL skinning (use_voxel_deformer)."""
import os

import numpy as np
import pytest
import torch

from tests.util import rel_err, rel_l2

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_voxel_lbs_golden():
    from emd_b200.voxel_deformer import VoxelDeformer, to_reference_layout
    z = np.load(f"{G}/omnire_modules.npz")
    t = lambda k: torch.from_numpy(z[k]).cuda()  # noqa: E731
    vd = VoxelDeformer(t("vox_base"), t("vox_offset"), t("vox_scale"), [int(v) for v in z["vox_res"]],
                       voxel_w_correction_ref=t("vox_corr"))
    xc = t("vox_xc").clone().requires_grad_(True)
    w = vd(xc)
    assert (w - t("vox_w")).abs().max().item() <= 2e-6
    (w * t("vox_cot")).sum().backward()
    assert (to_reference_layout(vd.voxel_w_correction.grad) - t("vox_v_corr")).abs().max().item() <= 5e-6
    ref = t("vox_v_xc")
    assert (xc.grad - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    assert torch.allclose(vd.get_tv("dc"), t("vox_tv"), rtol=1e-5) and torch.allclose(vd.get_mag("dc"), t("vox_mag"), rtol=1e-5)


@pytest.mark.parametrize("B,V,res,J,short,corr", [(8, 6890, [16, 64, 64], 24, 0, True), (2, 333, [5, 7, 20], 8, 1, False)])
def test_voxel_lbs_vs_oracle(B, V, res, J, short, corr):
    """Reference-sized volume (human_body.py:117-125: [res/4, res, res], 24 bones, 6890 template vertices) and a small
    odd-shaped one with the stretch on another axis and no correction enabled."""
    from emd_b200.voxel_deformer import VoxelDeformer, to_reference_layout
    from oracle import voxel_deformer as OV
    g = torch.Generator().manual_seed(B * 7 + V)
    base = torch.softmax(3.0 * torch.randn(B, J, *res, generator=g), dim=1)
    cor = (0.05 * torch.randn(B, J, *res, generator=g)).requires_grad_(True) if corr else None
    off, scl = 0.2 * torch.randn(B, 1, 3, generator=g), torch.rand(B, 1, 1, generator=g) + 0.7
    long_dim = 1 if short != 1 else 2
    ratio, ratio_dim = res[long_dim] / res[short], -1 - short
    xn = torch.rand(B, V, 3, generator=g) * 2.4 - 1.2                     # ~15 % outside [-1, 1]^3 (border clamp)
    mul = torch.ones(3); mul[ratio_dim] = ratio
    xc = (xn / mul * scl + off).requires_grad_(True)
    vol = base + cor if corr else base
    w_ref = OV.voxel_weights(vol, off, scl, ratio, ratio_dim, xc)
    cot = torch.randn(B, V, J, generator=g)
    (w_ref * cot).sum().backward()
    vd = VoxelDeformer(base.cuda(), off.cuda(), scl.cuda(), res, short_dim_dhw=short, long_dim_dhw=long_dim,
                       voxel_w_correction_ref=cor.detach().cuda() if corr else None)
    x = xc.detach().cuda().requires_grad_(True)
    w = vd(x)
    assert (w.detach().cpu() - w_ref.detach()).abs().max().item() <= 3e-6
    (w * cot.cuda()).sum().backward()
    gx = xc.grad
    assert (x.grad.cpu() - gx).abs().max().item() <= 1e-4 * max(1.0, gx.abs().max().item())
    assert rel_l2(x.grad, gx) <= 1e-4
    if corr:
        gc = to_reference_layout(vd.voxel_w_correction.grad).cpu()
        assert rel_err(gc, cor.grad) <= 1e-4 and rel_l2(gc, cor.grad) <= 1e-5
    else:
        assert vd.voxel_w_correction is None


def test_smpl_nodes_with_voxel_deformer():
    """SMPLNodes with use_voxel_deformer: W = VoxelDeformer(canonical means) feeds the skinning; gradients reach the voxel
    correction and (through both the skinning and the lookup) the means."""
    from emd_b200.voxel_deformer import VoxelDeformer, to_reference_layout
    from oracle import emd_smpl as ES
    from oracle import voxel_deformer as OV
    from tests.test_gpu_emd_smpl import GRAD, HEADS, _node, _setup
    I, V, frames, frame, step = 3, 700, 30, 11, 9000
    ss, cpu, p, g = _setup(7, I, V, frames)
    ss.instances_fv[frame] = True
    ss.instances_fv[frame, 2] = False
    res, J = [4, 16, 16], 24
    base = torch.softmax(2.0 * torch.randn(I, J, *res, generator=g), dim=1)
    cor = (0.02 * torch.randn(I, J, *res, generator=g)).requires_grad_(True)
    m = cpu["means"].detach().reshape(I, V, 3)
    lo, hi = m.min(1).values, m.max(1).values
    off = (0.5 * (lo + hi))[:, None]
    # the stretched axis (z, x4) of the synthetic figure is its longest: scale so that most points stay inside the volume
    # along z and a few fall outside (border clamp, zero coordinate gradient)
    scl = ((hi - lo).max(-1).values / 2 * 2.75)[:, None, None]
    ratio, ratio_dim = res[1] / res[0], -1
    p.W = OV.voxel_weights(base + cor, off, scl, ratio, ratio_dim, cpu["means"].reshape(I, V, 3))
    cam_pos = torch.tensor([0.0, 0.0, 1.6])
    ref = ES.get_gaussians(p, cpu["means"], cpu["quats"], cpu["scales"], cpu["opacities"], cpu["features_dc"],
                           cpu["features_rest"], frame, step, cam_pos)
    dev = "cuda"
    gpu = {k: v.detach().to(dev).requires_grad_(True) for k, v in cpu.items()}
    node = _node(ss, gpu, dev)
    vd = VoxelDeformer(base.to(dev), off.to(dev), scl.to(dev), res, voxel_w_correction_ref=cor.detach().to(dev))
    node.template["voxel_deformer"] = vd
    out = node.get_gaussians(cam_pos.tolist(), frame, step)
    cot = {}
    for k in ("_means", "_opacities", "_rgbs", "_scales", "_quats"):
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
    gc = to_reference_layout(vd.voxel_w_correction.grad).cpu()
    assert float(cor.grad.abs().max()) > 0
    assert rel_err(gc, cor.grad) <= 1e-3 and rel_l2(gc, cor.grad) <= 1e-3
