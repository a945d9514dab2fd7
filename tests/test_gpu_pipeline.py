"""GPU parity of the whole step (EMD deform -> activation/SH -> rasterization, fwd + bwd)
on a small Background + Rigid + SMPL scene, multi-camera call vs per-camera oracle."""
import pytest
import torch

from tests.util import rel_err, rel_l2

pytestmark = pytest.mark.gpu


def test_street_scene_step():
    from emd_b200 import pipeline as P, scenes
    from oracle import pipeline_ref as PR
    W, H = 192, 128
    bg, rigid, smpl = P.make_street_scene(n_bg=6000, rigid_instances=3, pts_per_rigid=400, smpl_instances=2,
                                          smpl_V=500, seed=3, num_frames=20)
    # bring the street into view of the small image: shrink the world
    frame, step = 7, 2500
    viewmats, Ks, c2w = scenes.cameras((0.0, 35.0), W, H)
    C = 2
    g = torch.Generator().manual_seed(0)
    L = PR.leaves(bg, rigid, smpl)
    refs, unstable = [], []
    for c in range(C):
        rgb, depth, alpha, info = PR.render(L, rigid, smpl, c2w[c], Ks[c], W, H, frame, step, return_unstable=True)
        refs.append((rgb, depth, alpha))
        unstable.append(info["unstable"][0])
    keep = torch.stack([(~u).float() for u in unstable])[..., None]  # [C,H,W,1]
    v_rgb = torch.randn(C, H, W, 3, generator=g) * keep
    v_d = 0.05 * torch.randn(C, H, W, 1, generator=g) * keep
    v_a = torch.randn(C, H, W, 1, generator=g) * keep
    sum((refs[c][0] * v_rgb[c]).sum() + (refs[c][1] * v_d[c]).sum() + (refs[c][2] * v_a[c]).sum() for c in range(C)).backward()

    dev = torch.device("cuda")
    scene = P.StreetScene(bg, rigid, smpl, dev)
    rgb, depth, alpha, info = scene.render(c2w.to(dev), Ks.to(dev), W, H, frame, step)
    assert float(keep.mean()) > 0.995
    for c in range(C):
        ok = keep[c, ..., 0] > 0
        assert float((rgb[c].detach().cpu() - refs[c][0].detach()).abs()[ok].max()) <= 1e-4
        assert float((alpha[c].detach().cpu() - refs[c][2].detach()).abs()[ok].max()) <= 1e-4
        dref = refs[c][1].detach()
        assert float((depth[c].detach().cpu() - dref).abs()[ok].max()) <= 1e-4 * max(1.0, float(dref.max()))
    ((rgb * v_rgb.to(dev)).sum() + (depth * v_d.to(dev)).sum() + (alpha * v_a.to(dev)).sum()).backward()

    got = {}
    for k, v in scene.bg.items():
        got["bg." + k] = v
    ren = {"_means": "means", "_quats": "quats", "_scales": "scales", "_opacities": "opacities",
           "_features_dc": "features_dc", "_features_rest": "features_rest", "_embeddings": "embeddings"}
    for name, node in (("rigid", scene.rigid), ("smpl", scene.smpl)):
        for k, v in node.p.items():
            if isinstance(v, torch.Tensor) and v.is_floating_point():
                got[f"{name}.{ren.get(k, k)}"] = v
        for k, v in node.track.items():
            got[f"{name}.{k}"] = v
    assert set(got) == set(L), set(got) ^ set(L)
    for k in sorted(L):
        gr, gg = L[k].grad, got[k].grad
        assert gg is not None, k
        if gr is None or float(gr.abs().max()) == 0.0:
            assert float(gg.abs().max()) == 0.0, k
            continue
        e, l2 = rel_err(gg, gr), rel_l2(gg, gr)
        assert e <= 2e-3 and l2 <= 1e-3, f"grad {k}: max-rel {e}, l2-rel {l2}"
