"""Generate the golden fixtures in tests/golden/ by running the REFERENCE's own Python
(/root/reference, read-only) on the CPU.  Run in the build container only:

    python tests/golden/make_golden.py

The reference cannot be imported as is -- it needs open3d / omegaconf / pytorch3d / gsplat /
nvdiffrast (absent) and hard-codes .cuda() -- so the missing third-party modules are stubbed with
MagicMock, ``Tensor.cuda`` becomes the identity and ``device="cuda"`` arguments are dropped.  Only
in-tree reference code produces the numbers: RigidNodes / SMPLNodes methods
(OmniRe/models/nodes/{rigid,smpl}.py), the quaternion helpers (OmniRe/models/gaussians/basics.py,
S3Gaussian/utils/graphics_utils.py), eval_sh (S3Gaussian/utils/sh_utils.py) and the S3Gaussian
deformation network (S3Gaussian/scene/deformation.py).  gsplat's ``spherical_harmonics`` (stubbed
module) is replaced by the reference's own ``eval_sh`` on normalised directions.
"""
import importlib
import importlib.util
import math
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


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


def _cpuify():
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    for fn in ("zeros", "ones", "tensor", "arange", "full", "eye", "empty", "rand", "randn"):
        orig = getattr(torch, fn)

        def wrap(*a, __orig=orig, **k):
            if str(k.get("device", "")).startswith("cuda"):
                k.pop("device")
            return __orig(*a, **k)

        setattr(torch, fn, wrap)


def load_file(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    _stub(["open3d", "omegaconf", "pytorch3d", "pytorch3d.transforms", "pytorch3d.ops", "gsplat", "gsplat.rendering",
           "gsplat.cuda", "gsplat.cuda._wrapper", "nvdiffrast", "nvdiffrast.torch", "imageio", "matplotlib",
           "matplotlib.pyplot", "trimesh", "kornia", "third_party", "third_party.smplx", "third_party.smplx.smplx",
           "third_party.smplx.smplx.lbs", "third_party.smplx.smplx.utils", "smplx", "tinycudann", "simple_knn",
           "simple_knn._C", "diff_gauss", "plyfile", "lpips", "wandb", "pytorch_msssim", "mmcv"])
    _cpuify()
    g = torch.Generator().manual_seed(20240917)

    # ---------------------------------------------------------------- SH (S3Gaussian/utils/sh_utils.py:57-112)
    sh_utils = load_file("ref_sh_utils", f"{REF}/S3Gaussian/utils/sh_utils.py")
    N = 64
    dirs = torch.randn(N, 3, generator=g)
    dirs_n = dirs / dirs.norm(dim=-1, keepdim=True)
    coeffs = torch.randn(N, 16, 3, generator=g)  # [N,K,3] gsplat layout
    sh_out = {f"deg{d}": sh_utils.eval_sh(d, coeffs.transpose(1, 2), dirs_n).numpy() for d in range(4)}
    np.savez(f"{HERE}/sh_eval.npz", dirs=dirs.numpy(), coeffs=coeffs.numpy(), **sh_out)

    # ---------------------------------------------------------------- quaternions
    sys.path.insert(0, f"{REF}/OmniRe")
    basics = importlib.import_module("models.gaussians.basics")
    gu = load_file("ref_graphics_utils", f"{REF}/S3Gaussian/utils/graphics_utils.py")
    q1 = torch.randn(50, 4, generator=g); q2 = torch.randn(50, 4, generator=g)
    q2[:5] = q1[:5] + 1e-3 * torch.randn(5, 4, generator=g)       # near-parallel branch
    q2[5:10] = -q1[5:10] + 0.3 * torch.randn(5, 4, generator=g)   # negative-dot branch
    np.savez(f"{HERE}/quat.npz", q1=q1.numpy(), q2=q2.numpy(),
             rotmat=basics.quat_to_rotmat(q1.clone()).numpy(),
             mult=basics.quat_mult(q1.clone(), q2.clone()).numpy(),
             slerp=basics.interpolate_quats(q1.clone(), q2.clone()).numpy(),
             bqm=gu.batch_quaternion_multiply(q1.clone(), q2.clone()).numpy())

    # ---------------------------------------------------------------- EMD rigid (OmniRe/models/nodes/rigid.py)
    basics.spherical_harmonics = lambda deg, d, c: sh_utils.eval_sh(deg, c.transpose(1, 2), d / d.norm(dim=-1, keepdim=True))
    rigid_mod = importlib.import_module("models.nodes.rigid")
    rigid_mod.spherical_harmonics = basics.spherical_harmonics
    from emd_b200 import scenes
    rs = scenes.rigid_nodes(3, 40, g, num_frames=12)
    rs.instances_quats = rs.instances_quats + 0.05 * torch.randn(rs.instances_quats.shape, generator=g)
    rs.point_ids[rs.point_ids == 2] = 1  # instance 2 owns no points: the NaN-skip path (rigid.py:528, :559)
    node = object.__new__(rigid_mod.RigidNodes)
    torch.nn.Module.__init__(node)
    P = torch.nn.Parameter
    node._means, node._quats, node._scales = P(rs.means.clone()), P(rs.quats.clone()), P(rs.scales.clone())
    node._opacities, node._features_dc, node._features_rest = P(rs.opacities.clone()), P(rs.features_dc.clone()), P(rs.features_rest.clone())
    node._embeddings, node.weight = P(rs.embeddings.clone()), P(rs.weight.clone())
    node.point_ids = rs.point_ids.clone()
    node.instances_quats, node.instances_trans = P(rs.instances_quats.clone()), P(rs.instances_trans.clone())
    node.instances_fv = rs.instances_fv.clone()
    node.instances_size = torch.tensor([[4.6, 2.0, 1.6]]).repeat(3, 1)
    for nm, out in (("rot_c", 1), ("rot_f", 1), ("trans_c", 3), ("trans_f", 3)):
        lin = torch.nn.Linear(36, out)
        lin.weight.data.copy_(rs.track[nm + "_w"]); lin.bias.data.copy_(rs.track[nm + "_b"])
        setattr(node, "track_" + nm, lin)
    node.temporal_embedding_dim, node.gaussian_embedding_dim = 32, 4
    node.max_embeddings, node.min_embeddings, node.c2f_temporal_iter = 150, 30, 20000
    for flag in ("no_temporal_embedding_dim", "no_gaussian_embedding_dim", "no_coarse_deform", "no_fine_deform",
                 "no_c2f_temporal_embedding", "no_apply_embed_shs", "no_apply_embed_track"):
        setattr(node, flag, False)
    node.ball_gaussians, node.gaussian_2d = False, False
    node.ctrl_cfg = types.SimpleNamespace(sh_degree_interval=1000, sh_degree=3)
    node.device = torch.device("cpu")
    out = dict(means=rs.means, quats=rs.quats, scales=rs.scales, opacities=rs.opacities, features_dc=rs.features_dc,
               features_rest=rs.features_rest, embeddings=rs.embeddings, point_ids=rs.point_ids, weight=rs.weight,
               instances_quats=rs.instances_quats, instances_trans=rs.instances_trans,
               instances_fv=rs.instances_fv, **{"track_" + k: v for k, v in rs.track.items()})
    out = {k: v.numpy() for k, v in out.items()}
    cam = types.SimpleNamespace(camtoworlds=torch.eye(4))
    cam.camtoworlds[:3, 3] = torch.tensor([0.3, -0.2, 1.6])
    cases = [(0, 0, False), (5, 3500, False), (11, 20000, False), (6, 12000, True), (3, 31000, False)]
    out["cases"] = np.array([(f, s, int(t)) for f, s, t in cases])
    with torch.no_grad():
        for ci, (frame, step, test_set) in enumerate(cases):
            node.cur_frame, node.step, node.in_test_set = frame, step, test_set
            tnorm = torch.tensor([[frame / (12 - 1)]]).float()
            out[f"c{ci}_temb_coarse"] = node.get_temporal_embed(tnorm, 30, weight=node.weight[0]).numpy()
            cur = node.int_lininterp(step, 30, 150, 20000)
            out[f"c{ci}_temb_fine"] = node.get_temporal_embed(tnorm, cur, weight=node.weight[0]).numpy()
            for ins in (0, 1):
                emb = node._embeddings[(node.point_ids == ins).squeeze(1), :]
                out[f"c{ci}_rot_off_{ins}"] = node.embedding_track_rot_offset(frame, 0, 11, emb, node.weight[ins]).numpy()
                out[f"c{ci}_trans_off_{ins}"] = node.embedding_track_trans_offset(frame, 0, 11, emb, node.weight[ins]).numpy()
            out[f"c{ci}_world_means"] = node.transform_means(node._means).numpy()
            out[f"c{ci}_world_quats"] = node.transform_quats(node._quats).numpy()
            gs = node.get_gaussians(cam)
            for k, v in gs.items():
                out[f"c{ci}_gs{k}"] = v.numpy()
    np.savez(f"{HERE}/emd_rigid.npz", cam_pos=cam.camtoworlds[:3, 3].numpy(), **out)

    # ---------------------------------------------------------------- EMD SMPL joint offsets (smpl.py:401-436)
    # (the skinning half needs SMPL_NEUTRAL.pkl + smplx, absent: only the EMD head is pinned here)
    try:
        smpl_mod = importlib.import_module("models.nodes.smpl")
        snode = object.__new__(smpl_mod.SMPLNodes)
        torch.nn.Module.__init__(snode)
        for flag in ("no_temporal_embedding_dim", "no_gaussian_embedding_dim", "no_coarse_deform", "no_fine_deform",
                     "no_c2f_temporal_embedding"):
            setattr(snode, flag, False)
        snode.temporal_embedding_dim, snode.max_embeddings, snode.c2f_temporal_iter = 32, 150, 20000
        cw, cb = 0.05 * torch.randn(24, 36, generator=g), 0.05 * torch.randn(24, generator=g)
        fw, fb = 0.05 * torch.randn(24, 36, generator=g), 0.05 * torch.randn(24, generator=g)
        snode.track_smpl_c, snode.track_smpl_f = torch.nn.Linear(36, 24), torch.nn.Linear(36, 24)
        snode.track_smpl_c.weight.data.copy_(cw); snode.track_smpl_c.bias.data.copy_(cb)
        snode.track_smpl_f.weight.data.copy_(fw); snode.track_smpl_f.bias.data.copy_(fb)
        emb = 0.1 * torch.randn(77, 4, generator=g)
        table = torch.randn(150, 32, generator=g) * (0.01 / math.sqrt(32))
        res = {}
        with torch.no_grad():
            for ci, (frame, step) in enumerate([(0, 0), (4, 9000), (9, 25000)]):
                snode.step = step
                res[f"c{ci}"] = snode.embedding_track_smpl_offset(frame, 0, 9, emb, table).numpy()
        np.savez(f"{HERE}/emd_smpl_offsets.npz", c_w=cw.numpy(), c_b=cb.numpy(), f_w=fw.numpy(), f_b=fb.numpy(),
                 embeddings=emb.numpy(), table=table.numpy(), cases=np.array([(0, 0), (4, 9000), (9, 25000)]), **res)
    except Exception as e:  # noqa: BLE001
        print("SMPL offsets golden skipped:", repr(e))
    print("golden fixtures written to", HERE)


if __name__ == "__main__" and not any(a in sys.argv for a in ("--s3g", "--hexplane", "--losses", "--modules", "--deformable", "--densify")):
    main()


def make_s3g_golden():
    """S3Gaussian deformation network (S3Gaussian/scene/deformation.py) run on the CPU with the flags of
    scripts/dynamic/run_dynamic_*.sh (--no_ds --no_dr --no_fine_hexplane_features) -> tests/golden/emd_s3g.npz."""
    _stub(["tkinter", "tinycudann", "open3d", "plyfile", "simple_knn", "simple_knn._C", "diff_gauss", "nvdiffrast",
           "nvdiffrast.torch", "lpips", "matplotlib", "matplotlib.pyplot", "imageio", "mmcv"])
    _cpuify()
    for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.") or k == "scene" or k.startswith("scene.")
              or k == "arguments" or k.startswith("arguments.")]:
        del sys.modules[k]
    sys.path.insert(0, f"{REF}/S3Gaussian")
    if f"{REF}/OmniRe" in sys.path:
        sys.path.remove(f"{REF}/OmniRe")
    # scene/__init__.py pulls the whole training stack; load the three modules we need directly
    pkg = types.ModuleType("scene"); pkg.__path__ = [f"{REF}/S3Gaussian/scene"]; sys.modules["scene"] = pkg
    upkg = types.ModuleType("utils"); upkg.__path__ = [f"{REF}/S3Gaussian/utils"]; sys.modules["utils"] = upkg
    apkg = types.ModuleType("arguments"); apkg.__path__ = [f"{REF}/S3Gaussian/arguments"]; sys.modules["arguments"] = apkg
    sys.modules["scene.encodings"] = MagicMock()
    opts = importlib.import_module("arguments.gaussian_options")
    deformation = importlib.import_module("scene.deformation")
    args = opts.BaseOptions()
    args.no_ds, args.no_dr, args.no_fine_hexplane_features = True, True, True
    torch.manual_seed(7)
    net = deformation.deform_network(args)
    net.deformation_net.time_offset.data = torch.tensor([[0.0], [0.013], [-0.02]])
    g = torch.Generator().manual_seed(11)
    N = 257
    point = (torch.rand(N, 3, generator=g) - 0.5) * 2.4
    scales = 0.1 * torch.randn(N, 3, generator=g); rots = torch.randn(N, 4, generator=g)
    opacity = torch.randn(N, 1, generator=g); shs = 0.3 * torch.randn(N, 16, 3, generator=g)
    emb = 0.1 * torch.randn(N, 4, generator=g)
    out = {"point": point, "scales": scales, "rotations": rots, "opacity": opacity, "shs": shs, "embeddings": emb}
    for k, v in net.state_dict().items():
        if "grid.grids" in k or "poc" in k or "aabb" in k:
            continue
        out["w." + k] = v.clone()
    cases = [(0.25, 6000, 0), (0.8, 17000, 1), (1.0, 40000, 2)]
    out["cases"] = torch.tensor([[t, it, c] for t, it, c in cases])
    grabbed = {}
    hook = net.deformation_net.grid.register_forward_hook(lambda m, i, o: grabbed.setdefault("hex", []).append(o.detach().clone()))
    with torch.no_grad():
        for ci, (t, it, cam_no) in enumerate(cases):
            grabbed.clear()
            times = torch.full((N, 1), t)
            m, s, r, o, sh, dd = net(point, scales, rots, opacity, shs, times, emb, it, cam_no, 1.0, is_train=False)
            out[f"c{ci}_hex"] = grabbed["hex"][0]   # hexplane features of the coarse pass (at `point`)
            out[f"c{ci}_means"], out[f"c{ci}_opacity"], out[f"c{ci}_shs"] = m, o, sh
            assert torch.equal(s, scales) and torch.equal(r, rots)
            for br in ("coarse", "fine"):
                for key in ("dx", "do", "dshs", "feat"):
                    out[f"c{ci}_{br}_{key}"] = dd[br][key]
                assert dd[br]["ds"] is None and dd[br]["dr"] is None
    hook.remove()
    np.savez(f"{HERE}/emd_s3g.npz", **{k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()})
    print("wrote emd_s3g.npz with", len(out), "arrays")


if __name__ == "__main__" and "--s3g" in sys.argv:
    make_s3g_golden()


def make_hexplane_golden():
    """HexPlaneField (S3Gaussian/scene/hexplane.py) run on the CPU: features and gradients w.r.t. points, times and
    planes -> tests/golden/hexplane.npz.  Plane contents come from oracle.hexplane.hash_planes (exact integer hash),
    so the fixture stores only points, cotangents and the reference's outputs."""
    ref = load_file("ref_hexplane", f"{REF}/S3Gaussian/scene/hexplane.py")
    from oracle import hexplane as OH
    out = {}
    cases = [
        # name, resolution, multires, bounds, set_aabb (max, min) or None, N, per-point times?
        ("a", [8, 8, 8, 5], [1, 2], 1.6, None, 300, True),
        ("b", [16, 16, 16, 25], [1, 2, 4, 8], 1.6, ([30.0, 12.0, 9.0], [-20.0, -14.0, -3.0]), 400, False),
    ]
    for name, reso, mr, bounds, box, N, per_point_t in cases:
        cfg = {"grid_dimensions": 2, "input_coordinate_dim": 4, "output_coordinate_dim": 32, "resolution": reso}
        field = ref.HexPlaneField(bounds, cfg, mr)
        if box is not None:
            field.set_aabb(*box)
        grids = OH.hash_planes(reso, mr, salt=len(name) + ord(name[0]))
        for s in range(len(mr)):
            for p in range(6):
                assert field.grids[s][p].shape == grids[s][p].shape
                field.grids[s][p].data = grids[s][p].clone()
        g = torch.Generator().manual_seed(3 + ord(name[0]))
        lo, hi = field.aabb.data.min(0).values, field.aabb.data.max(0).values
        pts = lo + (hi - lo) * (torch.rand(N, 3, generator=g) * 1.3 - 0.15)      # ~25 % outside the box (border clamp)
        pts[0] = hi; pts[1] = lo; pts[2] = 0.5 * (lo + hi)                         # exact corners / centre
        pts[3, 0] = lo[0] + (hi[0] - lo[0]) * 0.25                                 # lands on a grid node
        t = torch.full((N, 1), 0.37)
        if per_point_t:
            t[5:60, 0] = torch.rand(55, generator=g) * 2.6 - 1.3
            t[60, 0], t[61, 0] = 1.0, -1.0
        pts.requires_grad_(True); t.requires_grad_(True)
        feat = field(pts, t)
        cot = torch.randn(feat.shape, generator=g)
        (feat * cot).sum().backward()
        out[f"{name}_resolution"], out[f"{name}_multires"] = np.array(reso), np.array(mr)
        out[f"{name}_salt"] = np.array(len(name) + ord(name[0]))
        out[f"{name}_aabb"] = field.aabb.data.numpy().copy()
        out[f"{name}_pts"], out[f"{name}_t"], out[f"{name}_cot"] = pts.detach().numpy(), t.detach().numpy(), cot.numpy()
        out[f"{name}_feat"] = feat.detach().numpy()
        out[f"{name}_v_pts"], out[f"{name}_v_t"] = pts.grad.numpy(), t.grad.numpy()
        gg = [field.grids[s][p].grad for s in range(len(mr)) for p in range(6)]
        out[f"{name}_v_plane_sum"] = np.array([x.double().sum().item() for x in gg])
        out[f"{name}_v_plane_l2"] = np.array([x.double().pow(2).sum().sqrt().item() for x in gg])
        if name == "a":
            for k, x in enumerate(gg):
                out[f"a_v_plane{k}"] = x.numpy()
    np.savez_compressed(f"{HERE}/hexplane.npz", **out)
    print("wrote hexplane.npz with", len(out), "arrays")


if __name__ == "__main__" and "--hexplane" in sys.argv:
    make_hexplane_golden()


def make_losses_golden():
    """The reference's own loss code on the CPU -> tests/golden/losses.npz: OmniRe/models/losses.py (DepthLoss,
    binary_cross_entropy, safe_binary_cross_entropy; values and input gradients) and S3Gaussian/utils/loss_utils.py
    (l1_loss, ssim, compute_depth).  pytorch_msssim and kornia are absent, so the OmniRe SSIM / smoothness terms have no
    fixture (oracle/losses.py says so)."""
    _stub(["sklearn", "sklearn.cluster"])
    RL = load_file("ref_omnire_losses", f"{REF}/OmniRe/models/losses.py")
    SL = load_file("ref_s3g_loss_utils", f"{REF}/S3Gaussian/utils/loss_utils.py")
    from tests.loss_util import loss_inputs
    H, W = 40, 56
    d = loss_inputs(11, H, W)
    out = {"H": np.array(H), "W": np.array(W), "seed": np.array(11)}

    def grad_of(fn, x):
        x = x.clone().requires_grad_(True)
        y = fn(x)
        y.backward()
        return y.detach().numpy(), x.grad.numpy()

    valid = 1.0 - d["ego_mask"]
    hit = (d["lidar"] > 0).float() * valid
    for name, kw in (("l1_inv", dict(loss_type="l1", normalize=False, use_inverse_depth=True)),
                     ("l2_norm", dict(loss_type="l2", normalize=True, use_inverse_depth=False)),
                     ("sl1_norm_inv", dict(loss_type="smooth_l1", normalize=True, use_inverse_depth=True))):
        fn = RL.DepthLoss(**kw)
        out[f"depth_{name}"], out[f"depth_{name}_grad"] = grad_of(lambda x: fn(x, d["lidar"], hit), d["depth"])
    occ_t = (1.0 - d["sky_mask"]) * valid
    a = d["alpha"][..., 0]
    out["bce"], out["bce_grad"] = grad_of(lambda x: RL.binary_cross_entropy(x * valid, occ_t, reduction="mean"), a)
    out["safe_bce"], out["safe_bce_grad"] = grad_of(
        lambda x: RL.safe_binary_cross_entropy(x * valid, occ_t, limit=0.1, reduction="mean"), a)
    # S3Gaussian: CHW
    img = d["rgb"].permute(2, 0, 1).contiguous()
    gt = d["gt"].permute(2, 0, 1).contiguous()
    out["s3g_l1"], out["s3g_l1_grad"] = grad_of(lambda x: SL.l1_loss(x, gt), img)
    out["s3g_ssim"], out["s3g_ssim_grad"] = grad_of(lambda x: SL.ssim(x, gt), img)
    mask = (1.0 - d["sky_mask"])[None]
    dep = d["depth"].permute(2, 0, 1).contiguous()
    out["s3g_depth"], out["s3g_depth_grad"] = grad_of(lambda x: SL.compute_depth("l2", x * mask, d["lidar"][None] * mask), dep)
    np.savez_compressed(f"{HERE}/losses.npz", **out)
    print("wrote losses.npz with", len(out), "arrays")


if __name__ == "__main__" and "--losses" in sys.argv:
    make_losses_golden()


def make_modules_golden():
    """OmniRe/models/modules.py run on the CPU -> tests/golden/omnire_modules.npz: ``VoxelDeformer`` (weights at
    canonical points, gradients w.r.t. the voxel correction and the points, get_tv / get_mag) and
    ``ConditionalDeformNetwork`` (outputs and gradients w.r.t. every parameter and the condition)."""
    _stub(["pytorch3d", "pytorch3d.ops", "nvdiffrast", "nvdiffrast.torch", "utils", "utils.geometry"])
    M = load_file("ref_omnire_modules", f"{REF}/OmniRe/models/modules.py")
    out = {}
    # ---- VoxelDeformer: B = 3 instances, J = 24 bones, grid [D,H,W] = [4, 12, 10] (short axis z, as human_body.py:117-125)
    g = torch.Generator().manual_seed(41)
    B, V, J, res = 3, 200, 24, [4, 12, 10]
    vtx = torch.randn(B, 150, 3, generator=g) * torch.tensor([0.35, 0.25, 0.8]) + torch.tensor([0.1, -0.05, 0.2])
    feats = torch.softmax(torch.randn(B, 150, J, generator=g), dim=-1)
    torch.manual_seed(1234)
    vd = M.VoxelDeformer(vtx=vtx, vtx_features=feats, resolution_dhw=res, is_resume=True)   # random volume, no knn
    vd.enable_voxel_correction()
    vd.voxel_w_correction.data = 0.1 * torch.randn(vd.lbs_voxel_base.shape, generator=g)
    lo, hi = vd.bbox[:, 0], vd.bbox[:, 1]                                                    # [B,3]
    xc = lo[:, None] + (hi - lo)[:, None] * (torch.rand(B, V, 3, generator=g) * 1.3 - 0.15)  # ~25 % outside the box
    xc[:, 0], xc[:, 1], xc[:, 2] = hi, lo, 0.5 * (lo + hi)
    xc.requires_grad_(True)
    w = vd(xc)
    cot = torch.randn(w.shape, generator=g)
    (w * cot).sum().backward()
    out.update(vox_res=np.array(res), vox_base=vd.lbs_voxel_base.numpy(), vox_corr=vd.voxel_w_correction.detach().numpy(),
               vox_offset=vd.offset.numpy(), vox_scale=vd.scale.numpy(), vox_ratio=np.array(float(vd.ratio)),
               vox_ratio_dim=np.array(vd.ratio_dim), vox_xc=xc.detach().numpy(), vox_cot=cot.numpy(), vox_w=w.detach().numpy(),
               vox_v_xc=xc.grad.numpy(), vox_v_corr=vd.voxel_w_correction.grad.numpy(),
               vox_tv=vd.get_tv("dc").detach().numpy(), vox_mag=vd.get_mag("dc").detach().numpy())
    # ---- ConditionalDeformNetwork: the config's structure (D = 8, one skip, x/t multires 10, embed_dim 16, quaternion
    #      head, no scale head: omnire.yaml:159-166) at width 32 so the fixture stays small
    torch.manual_seed(77)
    net = M.ConditionalDeformNetwork(D=8, W=32, input_ch=3, embed_dim=16, x_multires=10, t_multires=10, deform_quat=True,
                                     deform_scale=False)
    N = 300
    x = torch.rand(N, 3, generator=g) * 2 - 1
    t = torch.full((N, 1), 0.4375)
    cond = torch.rand(N, 16, generator=g).requires_grad_(True)
    d_xyz, rot, scl = net(x, t, cond)
    assert scl is None
    c1, c2 = torch.randn(d_xyz.shape, generator=g), torch.randn(rot.shape, generator=g)
    ((d_xyz * c1).sum() + (rot * c2).sum()).backward()
    for k, v in net.state_dict().items():
        out[f"net_sd.{k}"] = v.numpy()
    for k, p in net.named_parameters():
        out[f"net_grad.{k}"] = p.grad.numpy()
    out.update(net_x=x.numpy(), net_t=t.numpy(), net_cond=cond.detach().numpy(), net_c1=c1.numpy(), net_c2=c2.numpy(),
               net_d_xyz=d_xyz.detach().numpy(), net_rot=rot.detach().numpy(), net_v_cond=cond.grad.numpy())
    np.savez_compressed(f"{HERE}/omnire_modules.npz", **out)
    print("wrote omnire_modules.npz with", len(out), "arrays")


if __name__ == "__main__" and "--modules" in sys.argv:
    make_modules_golden()


def make_densify_golden():
    """VanillaGaussians.after_train (OmniRe/models/gaussians/vanilla.py:163-191), the reference's own method, three
    successive training steps on the CPU -> tests/golden/densify.npz."""
    _stub(["open3d", "omegaconf", "pytorch3d", "pytorch3d.transforms", "pytorch3d.ops", "gsplat", "gsplat.rendering",
           "gsplat.cuda", "gsplat.cuda._wrapper", "nvdiffrast", "nvdiffrast.torch", "imageio", "matplotlib",
           "matplotlib.pyplot", "trimesh", "kornia", "third_party", "third_party.smplx", "third_party.smplx.smplx",
           "third_party.smplx.smplx.lbs", "third_party.smplx.smplx.utils", "smplx", "sklearn", "sklearn.neighbors"])
    _cpuify()
    sys.path.insert(0, f"{REF}/OmniRe")
    vanilla = importlib.import_module("models.gaussians.vanilla")
    g = torch.Generator().manual_seed(77)
    n, steps = 500, 3
    me = types.SimpleNamespace(num_points=n, filter_mask=torch.ones(n, dtype=torch.bool), xys_grad_norm=None, vis_counts=None,
                               max_2Dsize=None)
    radii = (torch.randint(0, 40, (steps, n), generator=g) * (torch.rand(steps, n, generator=g) < 0.6)).to(torch.int32)
    grads = 1e-3 * torch.randn(steps, n, 2, generator=g)
    out = dict(radii=radii.numpy(), grads=grads.numpy(), last_size=np.int64(960))
    for s_ in range(steps):
        vanilla.VanillaGaussians.after_train(me, radii[s_], grads[s_], 960)
        out[f"s{s_}_xys_grad_norm"] = me.xys_grad_norm.clone().numpy()
        out[f"s{s_}_vis_counts"] = me.vis_counts.clone().numpy()
        out[f"s{s_}_max_2Dsize"] = me.max_2Dsize.clone().numpy()
    np.savez_compressed(f"{HERE}/densify.npz", **out)
    print("wrote densify.npz")


def make_deformable_golden():
    """The reference's own ``DeformableNodes.get_gaussians`` (OmniRe/models/nodes/deformable.py:49-113, with
    ``get_deformation`` :35-47 and the ``ConditionalDeformNetwork`` of models/modules.py) run on the CPU ->
    tests/golden/deformable_nodes.npz: outputs and gradients w.r.t. the network, the instance embedding and the Gaussian
    parameters, for the shipped control flags (use_deformgs_for_nonrigid, use_deformgs_after = 3000,
    stop_optimizing_canonical_xyz) and for stop_optimizing_canonical_xyz = False."""
    _stub(["open3d", "omegaconf", "pytorch3d", "pytorch3d.transforms", "pytorch3d.ops", "gsplat", "gsplat.rendering",
           "gsplat.cuda", "gsplat.cuda._wrapper", "nvdiffrast", "nvdiffrast.torch", "imageio", "matplotlib",
           "matplotlib.pyplot", "trimesh", "kornia", "third_party", "third_party.smplx", "third_party.smplx.smplx",
           "third_party.smplx.smplx.lbs", "third_party.smplx.smplx.utils", "smplx", "tinycudann", "simple_knn",
           "simple_knn._C", "diff_gauss", "plyfile", "lpips", "wandb", "pytorch_msssim", "mmcv"])
    _cpuify()
    g = torch.Generator().manual_seed(20250611)
    sh_utils = load_file("ref_sh_utils", f"{REF}/S3Gaussian/utils/sh_utils.py")
    sys.path.insert(0, f"{REF}/OmniRe")
    basics = importlib.import_module("models.gaussians.basics")
    sh = lambda deg, d, c: sh_utils.eval_sh(deg, c.transpose(1, 2), d / d.norm(dim=-1, keepdim=True))  # noqa: E731
    basics.spherical_harmonics = sh
    rigid_mod = importlib.import_module("models.nodes.rigid")
    rigid_mod.spherical_harmonics = sh
    deform_mod = importlib.import_module("models.nodes.deformable")
    deform_mod.spherical_harmonics = sh
    from emd_b200 import scenes
    I, F, E = 3, 12, 16
    rs = scenes.rigid_nodes(I, 60, g, num_frames=F)
    rs.instances_quats = rs.instances_quats + 0.05 * torch.randn(rs.instances_quats.shape, generator=g)
    node = object.__new__(deform_mod.DeformableNodes)
    torch.nn.Module.__init__(node)
    P = torch.nn.Parameter
    node._means, node._quats, node._scales = P(rs.means.clone()), P(rs.quats.clone()), P(rs.scales.clone())
    node._opacities, node._features_dc, node._features_rest = P(rs.opacities.clone()), P(rs.features_dc.clone()), P(rs.features_rest.clone())
    node._embeddings, node.weight = P(rs.embeddings.clone()), P(rs.weight.clone())
    node.point_ids = rs.point_ids.clone()
    node.instances_quats, node.instances_trans = P(rs.instances_quats.clone()), P(rs.instances_trans.clone())
    node.instances_fv = rs.instances_fv.clone()
    node.instances_size = torch.tensor([0.8, 0.7, 1.7]) + 0.2 * torch.rand(I, 3, generator=g)
    node.instances_embedding = P(torch.rand(I, E, generator=g))
    node.normalized_timestamps = torch.linspace(0, 1, F)
    for nm, out in (("rot_c", 1), ("rot_f", 1), ("trans_c", 3), ("trans_f", 3)):
        lin = torch.nn.Linear(36, out)
        lin.weight.data.copy_(rs.track[nm + "_w"]); lin.bias.data.copy_(rs.track[nm + "_b"])
        setattr(node, "track_" + nm, lin)
    node.temporal_embedding_dim, node.gaussian_embedding_dim = 32, 4
    node.max_embeddings, node.min_embeddings, node.c2f_temporal_iter = 150, 30, 20000
    for flag in ("no_temporal_embedding_dim", "no_gaussian_embedding_dim", "no_coarse_deform", "no_fine_deform",
                 "no_c2f_temporal_embedding", "no_apply_embed_shs", "no_apply_embed_track"):
        setattr(node, flag, False)
    node.ball_gaussians, node.gaussian_2d, node.in_test_set = False, False, False
    node.device = torch.device("cpu")
    torch.manual_seed(99)
    node.deform_network = deform_mod.ConditionalDeformNetwork(input_ch=3, D=8, W=32, embed_dim=E, x_multires=10, t_multires=10,
                                                              deform_quat=True, deform_scale=False)
    # a freshly initialised network deforms by ~1e-2; scale the heads up so the deformation is visible in the outputs
    with torch.no_grad():
        node.deform_network.gaussian_warp.weight.mul_(4.0)
        node.deform_network.gaussian_rotation.weight.mul_(4.0)
    out = dict(means=rs.means, quats=rs.quats, scales=rs.scales, opacities=rs.opacities, features_dc=rs.features_dc,
               features_rest=rs.features_rest, embeddings=rs.embeddings, point_ids=rs.point_ids, weight=rs.weight,
               instances_quats=rs.instances_quats, instances_trans=rs.instances_trans, instances_fv=rs.instances_fv,
               instances_size=node.instances_size, instances_embedding=node.instances_embedding.detach(),
               normalized_timestamps=node.normalized_timestamps, **{"track_" + k: v for k, v in rs.track.items()})
    out = {k: v.numpy() for k, v in out.items()}
    for k, v in node.deform_network.state_dict().items():
        out[f"net_sd.{k}"] = v.numpy().copy()
    cam = types.SimpleNamespace(camtoworlds=torch.eye(4))
    cam.camtoworlds[:3, 3] = torch.tensor([0.3, -0.2, 1.6])
    out["cam_pos"] = cam.camtoworlds[:3, 3].numpy()
    cases = [(5, 8000, True), (11, 3001, False), (2, 3000, True)]      # (frame, step, stop_optimizing_canonical_xyz)
    out["cases"] = np.array([(f, s, int(t)) for f, s, t in cases])
    grad_of = dict(means=node._means, quats=node._quats, scales=node._scales, opacities=node._opacities,
                   features_dc=node._features_dc, features_rest=node._features_rest, embeddings=node._embeddings,
                   weight=node.weight, instances_quats=node.instances_quats, instances_trans=node.instances_trans,
                   instances_embedding=node.instances_embedding)
    for ci, (frame, step, stop) in enumerate(cases):
        node.cur_frame, node.step = frame, step
        node.ctrl_cfg = types.SimpleNamespace(sh_degree_interval=1000, sh_degree=3, use_deformgs_for_nonrigid=True,
                                              use_deformgs_after=3000, stop_optimizing_canonical_xyz=stop)
        node.zero_grad(set_to_none=True)
        gs = node.get_gaussians(cam)
        loss = 0.0
        for k in ("_means", "_opacities", "_rgbs", "_scales", "_quats"):
            out[f"c{ci}_gs{k}"] = gs[k].detach().numpy()
            cot = torch.randn(gs[k].shape, generator=g)
            out[f"c{ci}_cot{k}"] = cot.numpy()
            loss = loss + (gs[k] * cot).sum()
        loss.backward()
        for k, p in grad_of.items():
            if p.grad is not None:
                out[f"c{ci}_grad_{k}"] = p.grad.numpy().copy()
        for k, p in node.deform_network.named_parameters():
            if p.grad is not None:
                out[f"c{ci}_netgrad.{k}"] = p.grad.numpy().copy()
        lx = node._gs_cache["local_xyz_deformed"]
        if lx is not None:
            out[f"c{ci}_local_xyz_deformed"] = lx.detach().numpy()
    np.savez_compressed(f"{HERE}/deformable_nodes.npz", **out)
    print("wrote deformable_nodes.npz with", len(out), "arrays")


if __name__ == "__main__" and "--deformable" in sys.argv:
    make_deformable_golden()

if __name__ == "__main__" and "--densify" in sys.argv:
    make_densify_golden()
