"""CPU: the REFERENCE's own call-site code, unmodified, driven through this package's drop-in shims.

Runs only where ``/root/reference`` exists (the build container; the GPU box has no copy of the reference).  With
``emd_b200.compat.install()`` the reference's ``from gsplat.rendering import rasterization`` /
``from diff_gauss import GaussianRasterizationSettings, GaussianRasterizer`` bind to ``emd_b200.gsplat_api`` /
``emd_b200.diff_gauss_api``.  Then

  * ``BasicTrainer.render_gaussians`` (``OmniRe/models/trainers/base.py:385-432``) and
    ``BasicTrainer.postprocess_per_train_step`` (``:279-297``) are called as the reference's training loop calls them,
  * ``render`` of ``S3Gaussian/gaussian_renderer/__init__.py:27-303`` is called as ``train.py:207`` calls it,

and the test checks that every keyword they pass is accepted and every field they read back exists with the type /
shape they index it with (0-d tensor ``width`` / ``height``, ``info["means2d"].retain_grad()`` + ``.absgrad`` after
``backward()``, ``radii[0, mask]``, the 6-tuple of the rasterizer, ``viewspace_points.grad[:, :2]``).

There is no GPU here and the product has no CPU path, so the C-ABI STAGES behind the shims are replaced by
oracle-backed stand-ins for the duration of the test (test infrastructure; the stages themselves are what the ``-m gpu``
parity tests check).  What this test covers is the layer in between: the shims' own Python -- argument handling, meta
dict, autograd wiring -- against the reference's real consumers.
"""
import importlib
import os
import sys
import types
from unittest.mock import MagicMock

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")


def _stub(names):
    for n in names:
        parts = n.split(".")
        for i in range(1, len(parts) + 1):
            sub = ".".join(parts[:i])
            if sub not in sys.modules:
                m = MagicMock(name=sub)
                m.__path__ = []
                m.__spec__ = None
                sys.modules[sub] = m


@pytest.fixture(autouse=True)
def _restore_modules():
    """The stubs and the reference's modules must not outlive the test (other tests import torch internals lazily)."""
    before = dict(sys.modules)
    path = list(sys.path)
    yield
    keep = ("emd_b200", "oracle", "tests", "torch", "numpy")    # real packages imported on the way stay imported
    for k in [k for k in sys.modules if k not in before and k.split(".")[0] not in keep]:
        del sys.modules[k]
    sys.modules.update(before)
    sys.path[:] = path


@pytest.fixture()
def oracle_stages(monkeypatch):
    """Stand-ins for the C-ABI stages (projection, binning + sort, tile ranges, compositing; diff_gauss preprocess) built
    from the CPU oracle.  They keep the stage SIGNATURES of ``emd_b200.raster_ops``."""
    from emd_b200 import raster_ops as R, diff_gauss_api as DGA
    from oracle import gsplat_ref as G, diff_gauss_ref as DG

    def fully_fused_projection(means, quats, scales, viewmats, Ks, width, height, eps2d=0.3, near_plane=0.01,
                               far_plane=1e10, radius_clip=0.0, calc_compensations=False):
        radii, means2d, depths, conics, comps = G.projection(means, quats, scales, viewmats, Ks, width, height, eps2d,
                                                             near_plane, far_plane, radius_clip)
        tw, th, _ = R.tile_grid(width, height)
        x0, y0, x1, y1 = G.tile_rects(means2d.detach(), radii, 16, tw, th)
        return radii, means2d, depths, conics, (comps if calc_compensations else None), ((x1 - x0) * (y1 - y0)).to(torch.int32)

    def isect_tiles(means2d, radii, depths, tiles_per_gauss, width, height, sort=True, between=None, grad_enabled=True):
        if between is not None:
            with torch.set_grad_enabled(grad_enabled):
                between()
        tw, th, _ = R.tile_grid(width, height)
        tpg, keys, flat, _ = G.isect_tiles(means2d, radii, depths, 16, tw, th)
        keys, flat = G.sort_isects(keys, flat)
        return tpg, keys, flat, torch.cumsum(tpg.reshape(-1).long(), 0)

    def isect_offset_encode(isect_ids, C, width, height):
        tw, th, bits = R.tile_grid(width, height)
        return G.isect_offset_encode(isect_ids, C, tw, th, bits)

    def rasterize_to_pixels(means2d, conics, colors, opacities, depths, backgrounds, radii, cum_tiles, isect_offsets,
                            flatten_ids, isect_ids, width, height, with_depth=False, ed_mode=False, absgrad=False, flavour=0):
        C, N = radii.shape
        cols = colors if colors.dim() == 3 else colors[None].expand(C, -1, -1)
        if with_depth:
            cols = torch.cat([cols, depths[..., None]], -1)
        op = opacities if opacities.dim() == 2 else opacities[None].expand(C, -1)
        kw = dict(max_alpha=0.99, t_stop_inclusive=False, pixel_center=0.0) if flavour == 1 else {}
        out, alpha, last = G.rasterize_to_pixels(means2d, conics, cols, op, width, height, 16, isect_offsets, flatten_ids,
                                                 backgrounds, **kw)
        if ed_mode:
            out = torch.cat([out[..., :-1], out[..., -1:] / alpha.clamp(min=1e-10)], -1)
        if absgrad and means2d.requires_grad:      # the product sets .absgrad in its backward; mirror the timing
            means2d.register_hook(lambda g, t=means2d: setattr(t, "absgrad", g.abs()))
        return out, alpha, last

    for name, fn in (("fully_fused_projection", fully_fused_projection), ("isect_tiles", isect_tiles),
                     ("isect_offset_encode", isect_offset_encode), ("rasterize_to_pixels", rasterize_to_pixels)):
        monkeypatch.setattr(R, name, fn)

    def rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, extra_attrs,
                            s, info=None, cache=None):
        if (shs is None) == (colors_precomp is None):
            raise ValueError("Please provide exactly one of either SHs or precomputed colors!")
        so = DG.Settings(s.image_height, s.image_width, s.tanfovx, s.tanfovy, s.bg, s.scale_modifier, s.viewmatrix,
                         s.projmatrix, s.sh_degree, s.campos)
        color, depth, normal, alpha, radii, _ = DG.rasterize(means3D, means2D, shs, colors_precomp, opacities, scales,
                                                             rotations, so)
        return color, depth, normal, alpha, radii, torch.zeros(0, s.image_height, s.image_width)

    monkeypatch.setattr(DGA, "rasterize_gaussians", rasterize_gaussians)
    return True


def _cpuify(monkeypatch):
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    for fn in ("zeros", "ones", "tensor", "zeros_like", "full"):
        orig = getattr(torch, fn)

        def wrap(*a, __orig=orig, **k):
            if str(k.get("device", "")).startswith("cuda"):
                k.pop("device")
            return __orig(*a, **k)

        monkeypatch.setattr(torch, fn, wrap)


def _fresh_imports(prefixes):
    for k in [k for k in sys.modules if any(k == p or k.startswith(p + ".") for p in prefixes)]:
        del sys.modules[k]


def test_omnire_render_gaussians_and_postprocess_through_the_shim(oracle_stages, monkeypatch):
    import emd_b200.compat
    from emd_b200 import gsplat_api, scenes
    _stub(["open3d", "omegaconf", "pytorch3d", "pytorch3d.transforms", "pytorch3d.ops", "nvdiffrast", "nvdiffrast.torch",
           "imageio", "matplotlib", "matplotlib.pyplot", "trimesh", "kornia", "viser", "nerfview", "pytorch_msssim",
           "torchmetrics", "torchmetrics.image", "torchmetrics.image.lpip", "third_party", "third_party.smplx",
           "third_party.smplx.smplx", "third_party.smplx.smplx.lbs", "third_party.smplx.smplx.utils", "smplx", "sklearn",
           "sklearn.neighbors", "lpips", "wandb", "skimage", "skimage.metrics", "cv2", "tqdm"])
    _fresh_imports(["gsplat", "diff_gauss", "models", "utils", "datasets"])
    emd_b200.compat.install(force=True)
    monkeypatch.syspath_prepend(f"{REF}/OmniRe")
    basics = importlib.import_module("models.gaussians.basics")
    assert basics.rasterization is gsplat_api.rasterization, "the reference's import must bind to the drop-in"
    tr_pkg = types.ModuleType("models.trainers")       # skip the package __init__ (it pulls in the dataset stack)
    tr_pkg.__path__ = [f"{REF}/OmniRe/models/trainers"]
    sys.modules["models.trainers"] = tr_pkg
    base = importlib.import_module("models.trainers.base")

    W, H, n = 96, 64, 400
    g = torch.Generator().manual_seed(5)
    sc = scenes.simple_gaussians(n, g, W, H, scale=0.08)
    _, Ks, c2w = scenes.cameras((0.0,), W, H)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
    gs = basics.dataclass_gs(_means=leaves["means"], _scales=leaves["scales"], _quats=leaves["quats"], _rgbs=leaves["colors"],
                             _opacities=leaves["opacities"][:, None], detach_keys=[], extras=None)
    # width / height arrive as 0-d integer tensors (pixel_source.py:653-654 -> base.py:336-337)
    cam = basics.dataclass_camera(camtoworlds=c2w[0], camtoworlds_gt=c2w[0], Ks=Ks[0], H=torch.tensor(H), W=torch.tensor(W))
    seen = {}

    class _Model:
        def postprocess_per_train_step(self, step, optimizer, radii, xys_grad, last_size):
            seen.update(radii=radii, xys_grad=xys_grad, last_size=last_size)

    trainer = types.SimpleNamespace(
        render_cfg=types.SimpleNamespace(packed=False, absgrad=True, sparse_grad=False, antialiased=False, batch_size=1),
        training=True, viewer=None, optimizer=None, models={"Background": _Model()}, gaussian_classes={"Background": 0},
        pts_labels=torch.zeros(n, dtype=torch.long))
    # the kwargs of scene_graph.py:241-248
    results, render_fn = base.BasicTrainer.render_gaussians(trainer, gs, cam, near_plane=0.1, far_plane=1e10,
                                                            render_mode="RGB+ED", radius_clip=0.0)
    assert results["rgb_gaussians"].shape == (H, W, 3) and results["depth"].shape == (H, W, 1) and results["opacity"].shape == (H, W, 1)
    assert float(results["rgb_gaussians"].max()) <= 1.0
    info = trainer.info
    assert info["means2d"].shape == (1, n, 2) and info["radii"].shape == (1, n) and info["width"] == W and info["height"] == H
    (results["rgb_gaussians"].mean() + 0.01 * results["depth"].mean() + results["opacity"].mean()).backward()
    assert info["means2d"].grad is not None and info["means2d"].absgrad.shape == (1, n, 2)
    assert all(leaves[k].grad is not None for k in ("means", "quats", "scales", "opacities", "colors"))
    base.BasicTrainer.postprocess_per_train_step(trainer, step=10)
    assert seen["radii"].shape == (n,) and seen["xys_grad"].shape == (n, 2) and seen["last_size"] == max(W, H)
    vis = seen["radii"] > 0
    assert int(vis.sum()) > n // 4 and float(seen["xys_grad"][vis].abs().sum()) > 0
    assert torch.allclose(seen["xys_grad"][..., 0], info["means2d"].absgrad[0, :, 0] * W / 2.0)
    # the second (masked, no-grad) render of scene_graph.py:260-275: the closure is called again with an opacity mask
    with torch.no_grad():
        rgb2, depth2, op2 = render_fn(opaticy_mask=(torch.arange(n) % 2 == 0).float())
    assert rgb2.shape == (H, W, 3) and float(op2.mean()) < float(results["opacity"].mean())


def test_s3g_render_through_the_shim(oracle_stages, monkeypatch):
    import emd_b200.compat
    from emd_b200 import diff_gauss_api, scenes, s3g_render as SR
    _stub(["tkinter", "tinycudann", "open3d", "plyfile", "simple_knn", "simple_knn._C", "nvdiffrast", "nvdiffrast.torch", "lpips",
           "matplotlib", "matplotlib.pyplot", "imageio", "mmcv", "cv2", "tqdm", "sklearn", "sklearn.neighbors"])
    _fresh_imports(["gsplat", "diff_gauss", "scene", "utils", "arguments", "gaussian_renderer", "models"])
    emd_b200.compat.install(force=True)
    for name, sub in (("scene", "scene"), ("utils", "utils"), ("arguments", "arguments")):
        pkg = types.ModuleType(name)
        pkg.__path__ = [f"{REF}/S3Gaussian/{sub}"]
        sys.modules[name] = pkg
    sys.modules["scene.gaussian_model"] = MagicMock()       # only the GaussianModel type annotation is used
    _cpuify(monkeypatch)
    spec = importlib.util.spec_from_file_location("ref_gaussian_renderer", f"{REF}/S3Gaussian/gaussian_renderer/__init__.py")
    gr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gr)
    assert gr.GaussianRasterizer is diff_gauss_api.GaussianRasterizer
    assert gr.GaussianRasterizationSettings is diff_gauss_api.GaussianRasterizationSettings

    W, H, n = 96, 64, 300
    g = torch.Generator().manual_seed(6)
    sc = scenes.simple_gaussians(n, g, W, H, scale=0.08)
    P = lambda t: t.clone().requires_grad_(True)  # noqa: E731
    feat = {"coarse": P(torch.rand(n, 3, generator=g)), "fine": P(torch.rand(n, 3, generator=g))}

    calls = []

    def deformation(means3D, scales, rotations, opacity, shs, time, embeddings, it, cam_no, time_diff, is_train=False):
        # __init__.py:63,93 hand over the [N,1] repeat of the view's time (this package's mirror passes it once, [1,1])
        assert time.shape in ((n, 1), (1, 1)) and it == 7000 and cam_no == 1
        calls.append(tuple(time.shape))
        dd = {k: dict(dx=torch.zeros(n, 3), do=torch.zeros(n, 1), dshs=torch.zeros(n, 16, 3), feat=feat[k]) for k in feat}
        return means3D + 0.0, scales, rotations, opacity, shs, dd

    pc = types.SimpleNamespace(
        get_xyz=P(sc["means"]), _opacity=P(torch.logit(sc["opacities"].clamp(0.02, 0.98))[:, None]), _scaling=P(torch.log(sc["scales"])),
        _rotation=P(sc["quats"]), get_features=P(torch.cat([(torch.rand(n, 1, 3, generator=g) - 0.5) / 0.28, 0.1 * torch.randn(n, 15, 3, generator=g)], 1)),
        get_embedding=torch.zeros(n, 4), active_sh_degree=3, max_sh_degree=3, _deformation_table=None, _deformation=deformation,
        scaling_activation=torch.exp, rotation_activation=torch.nn.functional.normalize, opacity_activation=torch.sigmoid,
        _sky_model=lambda cam_, acc=None, is_train=False: torch.full((3, H, W), 0.5))
    cam0 = SR.make_camera(0.0, W, H, time=0.25, cam_no=1)
    cam = types.SimpleNamespace(FoVx=cam0.FoVx, FoVy=cam0.FoVy, image_height=H, image_width=W, time=0.25, cam_no=1, time_diff=0.0,
                                world_view_transform=cam0.world_view_transform, full_proj_transform=cam0.full_proj_transform,
                                camera_center=cam0.camera_center)
    args = types.SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False, combine_dynamic_static=False,
                                 no_coarse_deform=False, no_fine_deform=False)
    pkg = gr.render(args, cam, pc, torch.zeros(3), stage="fine", return_dx=True, render_feat=True, iter=7000, is_train=True)
    assert calls == [(n, 1)]
    for k in ("render", "depth", "weight", "feat_c", "feat_f", "sky_color"):
        assert pkg[k].shape[-2:] == (H, W), k
    assert pkg["radii"].shape == (n,) and pkg["visibility_filter"].dtype == torch.bool and "ddict" in pkg
    (pkg["render"].mean() + pkg["depth"].mean() * 0.01 + pkg["feat_c"].mean() + pkg["feat_f"].mean()).backward()
    # train.py:368, 407: the densification statistic
    vg = pkg["viewspace_points"].grad
    assert vg is not None and vg.shape == (n, 3) and float(vg[pkg["visibility_filter"], :2].norm(dim=-1).sum()) > 0
    assert pc.get_xyz.grad is not None and feat["coarse"].grad is not None and feat["fine"].grad is not None
    # and the package's own mirror of the same function returns the same keys for the same call
    mine = SR.render(SR.S3GOptions(), cam0, types.SimpleNamespace(**{**pc.__dict__}), torch.zeros(3), stage="fine", return_dx=True,
                     render_feat=True, iter=7000, is_train=True)
    assert set(pkg) <= set(mine) | {"normal"}
    for k in ("render", "depth", "weight", "feat_c", "feat_f"):
        assert torch.allclose(mine[k], pkg[k], atol=1e-6), k
