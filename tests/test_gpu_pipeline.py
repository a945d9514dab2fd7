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
        assert e <= 1e-3 and l2 <= 1e-3, f"grad {k}: max-rel {e}, l2-rel {l2}"


def test_two_call_activation_is_the_fused_one():
    """``activate_geometry`` + ``sh_colors`` (two autograd nodes, colour backward first) == ``activate_gaussians``
    bit for bit, forward and backward; and the step with the colours evaluated after the projection == the step in
    the reference's order (``colors_after_projection=False``), bit for bit."""
    from emd_b200 import pipeline as P, scenes
    from emd_b200.sh_ops import activate_gaussians, activate_geometry, sh_colors
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(5)
    N = 5003
    mk = lambda *s: torch.randn(*s, generator=g).to(dev)  # noqa: E731
    means = mk(N, 3) * 10
    leaves_a = [mk(N, 3).requires_grad_(), (0.1 * mk(N, 15, 3)).requires_grad_(), mk(N, 1).requires_grad_(),
                (0.3 * mk(N, 3)).requires_grad_(), mk(N, 4).requires_grad_()]
    leaves_b = [t.detach().clone().requires_grad_() for t in leaves_a]
    cams = [[0.0, 0.0, 1.5], [3.0, -2.0, 1.0], [-4.0, 1.0, 2.0]]
    ids = torch.randint(0, 7, (N,), generator=g).to(dev)
    valid = (torch.rand(7, generator=g) > 0.3).to(dev)
    cot = [mk(3, N, 3), mk(N), mk(N, 3), mk(N, 4)]
    for deg in (0, 2, 3):
        for a in leaves_a + leaves_b:
            a.grad = None
        dc, rest, op, sc, q = leaves_a
        rgbs, opac, scl, qn = activate_gaussians(means, dc, rest, op, sc, q, cams, deg, point_ids=ids, inst_valid=valid)
        sum((o * c).sum() for o, c in zip((rgbs, opac, scl, qn), cot)).backward()
        dc2, rest2, op2, sc2, q2 = leaves_b
        opac2, scl2, qn2 = activate_geometry(op2, sc2, q2, point_ids=ids, inst_valid=valid)
        rgbs2 = sh_colors(means, dc2, rest2, cams, deg)
        for x, y in ((rgbs, rgbs2), (opac, opac2), (scl, scl2), (qn, qn2)):
            assert torch.equal(x, y)
        sum((o * c).sum() for o, c in zip((rgbs2, opac2, scl2, qn2), cot)).backward()
        for a, b in zip(leaves_a, leaves_b):
            assert torch.equal(a.grad, b.grad)

    W, H = 192, 128
    bg, rigid, smpl = P.make_street_scene(n_bg=6000, rigid_instances=3, pts_per_rigid=400, smpl_instances=2,
                                          smpl_V=500, seed=3, num_frames=20)
    viewmats, Ks, c2w = scenes.cameras((0.0, 35.0), W, H)
    outs = []
    for after in (False, True):
        scene = P.StreetScene(bg, rigid, smpl, dev)
        rgb, depth, alpha, info = scene.render(c2w.to(dev), Ks.to(dev), W, H, 7, 2500, colors_after_projection=after)
        (rgb.sum() + 0.1 * depth.sum() + alpha.sum()).backward()
        outs.append((rgb, depth, alpha, [p.grad for p in scene.parameters()]))
    for x, y in zip(outs[0][:3], outs[1][:3]):
        assert torch.equal(x, y)
    for x, y in zip(outs[0][3], outs[1][3]):
        assert (x is None and y is None) or torch.equal(x, y)
