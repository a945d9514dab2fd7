"""GPU parity: the fused densification-statistics kernel vs the oracle's restatement of
``postprocess_per_train_step`` + ``after_train`` (base.py:279-297, vanilla.py:163-191), itself pinned to the reference's
own ``after_train``; then through the real pipeline: ``info`` of a render + backward."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_kernel_reproduces_reference_steps():
    from emd_b200.densify import DensifyStats
    z = np.load(os.path.join(G, "densify.npz"))
    n = z["radii"].shape[1]
    st = DensifyStats(n, "cuda")
    for s_ in range(z["radii"].shape[0]):
        m2 = torch.zeros(1, n, 2, device="cuda")
        m2.absgrad = torch.from_numpy(z["grads"][s_])[None].cuda() / torch.tensor([480.0, 320.0], device="cuda")
        info = {"radii": torch.from_numpy(z["radii"][s_])[None].cuda(), "means2d": m2, "width": 960, "height": 640}
        # postprocess_per_train_step scales by (W/2, H/2) before after_train: feed grads / (W/2, H/2) so the golden's
        # already-scaled gradients come out
        st.update(info, absgrad=True)
        for k in ("xys_grad_norm", "vis_counts", "max_2Dsize"):
            ref = torch.from_numpy(z[f"s{s_}_{k}"])
            got = getattr(st, k).cpu()
            assert float((got - ref).abs().max()) <= 1e-6 * max(1.0, float(ref.abs().max())), (s_, k)


def test_multi_camera_call_equals_successive_steps_and_class_slices():
    """C cameras in one call == C successive single-camera reference steps; per-class views are slices."""
    from emd_b200.densify import DensifyStats
    from oracle import densify as OD
    g = torch.Generator().manual_seed(3)
    C, n, W, H = 3, 4000, 960, 640
    slices = {"Background": (0, 3000), "RigidNodes": (3000, 3700), "SMPLNodes": (3700, 4000)}
    st = DensifyStats(n, "cuda", slices)
    states = {k: {} for k in slices}
    for step in range(3):
        radii = (torch.randint(0, 60, (C, n), generator=g) * (torch.rand(C, n, generator=g) < 0.5)).to(torch.int32)
        grads = 1e-4 * torch.randn(C, n, 2, generator=g)
        m2 = torch.zeros(C, n, 2, device="cuda")
        m2.absgrad = grads.abs().cuda()
        st.update({"radii": radii.cuda(), "means2d": m2, "width": W, "height": H})
        for c in range(C):
            sg = OD.scale_grads(grads.abs()[c], W, H)
            for k, (a, b) in slices.items():
                OD.after_train(states[k], radii[c, a:b], sg[a:b], max(W, H))
    for k in slices:
        got = st.of(k)
        for name in ("xys_grad_norm", "vis_counts", "max_2Dsize"):
            ref = states[k][name]
            assert float((got[name].cpu() - ref).abs().max()) <= 2e-6 * max(1.0, float(ref.abs().max())), (k, name)
    st.reset()
    assert float(st.state.abs().max()) == 0.0 and st.first


def test_statistics_from_a_real_step():
    import emd_b200
    from emd_b200.densify import DensifyStats
    from oracle import densify as OD
    from tests.util import raster_scene
    W, H = 160, 96
    sc, viewmats, Ks, _, g = raster_scene(8, 1500, W, H, yaws=(0.0, 15.0))
    dev = "cuda"
    p = {k: v.to(dev).requires_grad_(True) for k, v in sc.items()}
    st = DensifyStats(1500, dev)
    state = {}
    for it in range(2):
        c, a, info = emd_b200.rasterization(p["means"], p["quats"], p["scales"], p["opacities"], p["colors"], viewmats.to(dev),
                                            Ks.to(dev), W, H, packed=False, absgrad=True, render_mode="RGB+ED")
        info["means2d"].retain_grad()
        (c[..., :3].mean() + a.mean()).backward()
        st.update(info)
        ab = info["means2d"].absgrad.cpu()
        for cam in range(2):
            OD.after_train(state, info["radii"][cam].cpu(), OD.scale_grads(ab[cam], W, H), max(W, H))
    for k in ("xys_grad_norm", "vis_counts", "max_2Dsize"):
        assert float((getattr(st, k).cpu() - state[k]).abs().max()) <= 2e-6 * max(1.0, float(state[k].abs().max())), k
