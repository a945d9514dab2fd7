"""CPU: pin the oracle against golden vectors produced by the REFERENCE's own Python
(tests/golden/make_golden.py ran OmniRe/models/nodes/{rigid,smpl}.py, models/gaussians/basics.py,
S3Gaussian/utils/{sh_utils,graphics_utils}.py in the build container)."""
import os

import numpy as np
import pytest
import torch

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_sh_eval_golden():
    from oracle import sh as SH
    z = np.load(f"{G}/sh_eval.npz")
    for d in range(4):
        out = SH.spherical_harmonics(d, _t(z["dirs"]), _t(z["coeffs"]))
        assert torch.allclose(out, _t(z[f"deg{d}"]), rtol=1e-5, atol=1e-6), d


def test_quaternion_golden():
    from oracle import quat as Q
    z = np.load(f"{G}/quat.npz")
    q1, q2 = _t(z["q1"]), _t(z["q2"])
    assert torch.allclose(Q.quat_to_rotmat(q1), _t(z["rotmat"]), atol=1e-6)
    assert torch.allclose(Q.quat_mult(q1, q2), _t(z["mult"]), atol=1e-6)
    assert torch.allclose(Q.interpolate_quats(q1, q2), _t(z["slerp"]), atol=2e-6)
    assert torch.allclose(Q.batch_quaternion_multiply(q1, q2), _t(z["bqm"]), atol=1e-6)


def _rigid_from_golden(z):
    from oracle import emd_rigid as ER
    return ER.RigidEMD(point_ids=_t(z["point_ids"])[:, 0], embeddings=_t(z["embeddings"]), weight=_t(z["weight"]),
                       instances_quats=_t(z["instances_quats"]), instances_trans=_t(z["instances_trans"]),
                       instances_fv=_t(z["instances_fv"]),
                       **{k: _t(z["track_" + k]) for k in ("rot_c_w", "rot_c_b", "rot_f_w", "rot_f_b", "trans_c_w",
                                                           "trans_c_b", "trans_f_w", "trans_f_b")})


def test_emd_rigid_golden():
    """get_temporal_embed, track offsets, transform_means/quats, get_gaussians of RigidNodes (incl. the
    coarse-to-fine schedule, the eval-time neighbour interpolation and the empty-instance NaN skip)."""
    from oracle import emd_rigid as ER
    z = np.load(f"{G}/emd_rigid.npz")
    p = _rigid_from_golden(z)
    F = p.num_frames
    for ci, (frame, step, test_set) in enumerate(z["cases"].tolist()):
        t = frame / (F - 1)
        assert torch.allclose(ER.get_temporal_embed(t, 30, p.weight[0]), _t(z[f"c{ci}_temb_coarse"]), atol=1e-7)
        cur = ER.int_lininterp(step, 30, 150, 20000)
        assert torch.allclose(ER.get_temporal_embed(t, cur, p.weight[0]), _t(z[f"c{ci}_temb_fine"]), atol=1e-7)
        dtrans, qoff, ok_t, ok_q = ER.track_offsets(p, frame, step)
        for ins in (0, 1):
            assert torch.allclose(dtrans[ins], _t(z[f"c{ci}_trans_off_{ins}"]), atol=1e-6)
            assert torch.allclose(qoff[ins], _t(z[f"c{ci}_rot_off_{ins}"]), atol=1e-6)
        assert not bool(ok_t[2]) and not bool(ok_q[2])  # instance 2 owns no points
        wm = ER.transform_means(p, _t(z["means"]), frame, step, in_test_set=bool(test_set))
        wq = ER.transform_quats(p, _t(z["quats"]), frame, step)
        assert torch.allclose(wm, _t(z[f"c{ci}_world_means"]), rtol=1e-5, atol=1e-5), ci
        assert torch.allclose(wq, _t(z[f"c{ci}_world_quats"]), atol=2e-6), ci
        gs = ER.get_gaussians(p, _t(z["means"]), _t(z["quats"]), _t(z["scales"]), _t(z["opacities"]),
                              _t(z["features_dc"]), _t(z["features_rest"]), frame, step, _t(z["cam_pos"]),
                              in_test_set=bool(test_set))
        for k, v in gs.items():
            assert torch.allclose(v, _t(z[f"c{ci}_gs{k}"]), rtol=1e-5, atol=2e-5), (ci, k)


def test_emd_smpl_offsets_golden():
    from oracle import emd_smpl as ES
    z = np.load(f"{G}/emd_smpl_offsets.npz")
    n = z["embeddings"].shape[0]
    p = ES.SMPLEMD(point_ids=torch.zeros(n, dtype=torch.long), embeddings=_t(z["embeddings"]),
                   weight=_t(z["table"])[None], instances_quats=None, smpl_quats=None,
                   instances_trans=torch.zeros(10, 1, 3), instances_fv=None, smpl_c_w=_t(z["c_w"]),
                   smpl_c_b=_t(z["c_b"]), smpl_f_w=_t(z["f_w"]), smpl_f_b=_t(z["f_b"]), J_canonical=None, A0_inv=None,
                   W=torch.zeros(1, n, 24))
    for ci, (frame, step) in enumerate(z["cases"].tolist()):
        off = ES.track_smpl_offset(p, 0, frame, step)
        assert torch.allclose(off, _t(z[f"c{ci}"]), atol=1e-6), ci


def test_emd_s3g_golden():
    """The S3Gaussian deformation network (coarse + fine heads, c2f schedule, per-camera time offset)."""
    from oracle import emd_s3g as S
    z = np.load(f"{G}/emd_s3g.npz")
    w = {k[len("w.deformation_net."):]: _t(z[k]) for k in z.files if k.startswith("w.deformation_net.")}
    for ci, (t, it, cam) in enumerate(z["cases"].tolist()):
        means, opac, shs, dd = S.deform(w, _t(z["point"]), _t(z["opacity"]), _t(z["shs"]), _t(z["embeddings"]),
                                        _t(z[f"c{ci}_hex"]), float(np.float32(t)), int(it), int(cam))
        assert torch.allclose(means, _t(z[f"c{ci}_means"]), atol=2e-6), ci
        assert torch.allclose(opac, _t(z[f"c{ci}_opacity"]), atol=2e-6), ci
        assert torch.allclose(shs, _t(z[f"c{ci}_shs"]), atol=2e-6), ci
        for br in ("coarse", "fine"):
            for key in ("dx", "do", "dshs", "feat"):
                assert torch.allclose(dd[br][key], _t(z[f"c{ci}_{br}_{key}"]), atol=2e-6), (ci, br, key)


@pytest.mark.parametrize("name", ["a", "b"])
def test_hexplane_oracle_matches_reference(name):
    """oracle/hexplane.py against features and gradients of the reference's own HexPlaneField
    (S3Gaussian/scene/hexplane.py), tests/golden/hexplane.npz."""
    from oracle import hexplane as OH
    from tests.hex_util import oracle_run
    z = np.load(f"{G}/hexplane.npz")
    grids = OH.hash_planes(list(z[f"{name}_resolution"]), list(z[f"{name}_multires"]), salt=int(z[f"{name}_salt"]))
    feat, v_pts, v_t, v_g = oracle_run(grids, _t(z[f"{name}_aabb"]), _t(z[f"{name}_pts"]), _t(z[f"{name}_t"]),
                                       _t(z[f"{name}_cot"]))
    assert np.abs(feat.numpy() - z[f"{name}_feat"]).max() <= 1e-6
    assert np.abs(v_pts.numpy() - z[f"{name}_v_pts"]).max() <= 2e-5 * max(1.0, np.abs(z[f"{name}_v_pts"]).max())
    assert np.abs(v_t.numpy() - z[f"{name}_v_t"]).max() <= 2e-5 * max(1.0, np.abs(z[f"{name}_v_t"]).max())
    flat = [g for row in v_g for g in row]
    sums = np.array([g.double().sum().item() for g in flat])
    l2 = np.array([g.double().pow(2).sum().sqrt().item() for g in flat])
    assert np.allclose(sums, z[f"{name}_v_plane_sum"], rtol=1e-4, atol=1e-4)
    assert np.allclose(l2, z[f"{name}_v_plane_l2"], rtol=1e-5)
    if name == "a":
        for k, g in enumerate(flat):
            assert np.abs(g.numpy() - z[f"a_v_plane{k}"]).max() <= 2e-6 * max(1.0, np.abs(z[f"a_v_plane{k}"]).max())


def test_losses_golden():
    """oracle/losses.py against the reference's own OmniRe/models/losses.py and S3Gaussian/utils/loss_utils.py
    (values and input gradients)."""
    from oracle import losses as OL
    from tests.loss_util import loss_inputs
    z = np.load(f"{G}/losses.npz")
    d = loss_inputs(int(z["seed"]), int(z["H"]), int(z["W"]))

    def grad_of(fn, x):
        x = x.clone().requires_grad_(True)
        y = fn(x)
        y.backward()
        return y.detach(), x.grad

    def check(name, fn, x, tol=1e-6):
        y, g = grad_of(fn, x)
        assert abs(float(y) - float(z[name])) <= tol * max(1.0, abs(float(z[name]))), name
        ref = _t(z[name + "_grad"])
        assert float((g - ref).abs().max()) <= tol * max(1e-12, float(ref.abs().max())), name

    valid = 1.0 - d["ego_mask"]
    hit = (d["lidar"] > 0).float() * valid
    for name, kw in (("l1_inv", dict(loss_type="l1", normalize=False, use_inverse_depth=True)),
                     ("l2_norm", dict(loss_type="l2", normalize=True, use_inverse_depth=False)),
                     ("sl1_norm_inv", dict(loss_type="smooth_l1", normalize=True, use_inverse_depth=True))):
        check(f"depth_{name}", lambda x: OL.depth_loss(x, d["lidar"], hit, **kw), d["depth"])
    occ_t = (1.0 - d["sky_mask"]) * valid
    a = d["alpha"][..., 0]
    check("bce", lambda x: OL.binary_cross_entropy(x * valid, occ_t), a)
    check("safe_bce", lambda x: OL.safe_binary_cross_entropy(x * valid, occ_t, limit=0.1), a)
    img, gt = d["rgb"].permute(2, 0, 1).contiguous(), d["gt"].permute(2, 0, 1).contiguous()
    check("s3g_l1", lambda x: (x - gt).abs().mean(), img)
    check("s3g_ssim", lambda x: OL.ssim_s3g(x, gt), img, tol=1e-5)
    mask = (1.0 - d["sky_mask"])[None]
    check("s3g_depth", lambda x: OL.compute_depth_s3g("l2", x * mask, d["lidar"][None] * mask),
          d["depth"].permute(2, 0, 1).contiguous())
    # the two window constructions (loss_utils.py:56-58 vs pytorch_msssim) agree to an ulp
    assert float((OL.gaussian_window_s3g() - OL.gaussian_window_msssim()).abs().max()) <= 2e-8


def test_voxel_deformer_golden():
    """oracle/voxel_deformer.py against the reference's own VoxelDeformer (weights, gradients w.r.t. the correction
    volume and the canonical points, get_tv / get_mag)."""
    from oracle import voxel_deformer as OV
    z = np.load(f"{G}/omnire_modules.npz")
    base, corr = _t(z["vox_base"]), _t(z["vox_corr"]).requires_grad_(True)
    xc = _t(z["vox_xc"]).requires_grad_(True)
    w = OV.voxel_weights(base + corr, _t(z["vox_offset"]), _t(z["vox_scale"]), float(z["vox_ratio"]), int(z["vox_ratio_dim"]), xc)
    assert torch.allclose(w, _t(z["vox_w"]), atol=2e-6)
    (w * _t(z["vox_cot"])).sum().backward()
    assert torch.allclose(corr.grad, _t(z["vox_v_corr"]), atol=2e-6)
    ref = _t(z["vox_v_xc"])
    assert (xc.grad - ref).abs().max() <= 2e-5 * max(1.0, ref.abs().max().item())
    assert (ref == 0).any() and (ref != 0).any()      # clipped and interior points are both covered
    assert torch.allclose(OV.get_tv(corr.detach()), _t(z["vox_tv"]), rtol=1e-6)
    assert torch.allclose(OV.get_mag(corr.detach()), _t(z["vox_mag"]), rtol=1e-6)


def test_conditional_deform_network_golden():
    """oracle/deform_network.py against the reference's own ConditionalDeformNetwork: outputs, gradients w.r.t. every
    parameter and the condition (instance embedding)."""
    from oracle import deform_network as ON
    z = np.load(f"{G}/omnire_modules.npz")
    sd = {k[len("net_sd."):]: _t(z[k]).requires_grad_(True) for k in z.files if k.startswith("net_sd.")}
    cond = _t(z["net_cond"]).requires_grad_(True)
    d_xyz, rot, scl = ON.conditional_deform_network(sd, _t(z["net_x"]), _t(z["net_t"]), cond, D=8)
    assert scl is None
    assert torch.allclose(d_xyz, _t(z["net_d_xyz"]), atol=1e-6) and torch.allclose(rot, _t(z["net_rot"]), atol=1e-6)
    ((d_xyz * _t(z["net_c1"])).sum() + (rot * _t(z["net_c2"])).sum()).backward()
    assert torch.allclose(cond.grad, _t(z["net_v_cond"]), atol=1e-6)
    for k, p in sd.items():
        ref = _t(z[f"net_grad.{k}"])
        assert (p.grad - ref).abs().max() <= 1e-5 * max(1.0, ref.abs().max().item()), k


def test_deformable_nodes_golden():
    """The oracle's DeformableNodes composition (oracle/deform_network.py:deformed_canonical -> oracle/emd_rigid.py:
    get_gaussians) against the reference's own ``DeformableNodes.get_gaussians`` run on the CPU: outputs, the cached
    deformed canonical points and every gradient (network, instance embedding, Gaussian parameters, EMD tables, poses),
    for stop_optimizing_canonical_xyz on / off and for a step before use_deformgs_after (plain rigid route)."""
    from oracle import deform_network as ON
    from oracle import emd_rigid as ER
    z = np.load(f"{G}/deformable_nodes.npz")
    heads = ("rot_c_w", "rot_c_b", "rot_f_w", "rot_f_b", "trans_c_w", "trans_c_b", "trans_f_w", "trans_f_b")
    names = ("means", "quats", "scales", "opacities", "features_dc", "features_rest", "embeddings", "weight",
             "instances_quats", "instances_trans", "instances_embedding")
    ts = z["normalized_timestamps"].tolist()
    for ci, (frame, step, stop) in enumerate(z["cases"].tolist()):
        c = {k: _t(z[k]).clone().requires_grad_(True) for k in names}
        net = {k[len("net_sd."):]: _t(z[k]).clone().requires_grad_(True) for k in z.files if k.startswith("net_sd.")}
        p = ER.RigidEMD(point_ids=_t(z["point_ids"])[:, 0], embeddings=c["embeddings"], weight=c["weight"],
                        instances_quats=c["instances_quats"], instances_trans=c["instances_trans"],
                        instances_fv=_t(z["instances_fv"]), **{k: _t(z["track_" + k]) for k in heads})
        if step > 3000:
            m, q = ON.deformed_canonical(net, c["means"], c["quats"], _t(z["point_ids"]), _t(z["instances_size"]),
                                         c["instances_embedding"], ts[frame], D=8, stop_optimizing_canonical_xyz=bool(stop))
            assert torch.allclose(m, _t(z[f"c{ci}_local_xyz_deformed"]), atol=1e-6)
        else:
            m, q = c["means"], c["quats"]
            assert f"c{ci}_local_xyz_deformed" not in z.files
        out = ER.get_gaussians(p, m, q, c["scales"], c["opacities"], c["features_dc"], c["features_rest"], int(frame),
                               int(step), _t(z["cam_pos"]))
        loss = 0.0
        for k in ("_means", "_opacities", "_rgbs", "_scales", "_quats"):
            ref = _t(z[f"c{ci}_gs{k}"])
            assert out[k].shape == ref.shape, (ci, k)
            assert (out[k] - ref).abs().max() <= 3e-6 * max(1.0, ref.abs().max().item()), (ci, k)
            loss = loss + (out[k] * _t(z[f"c{ci}_cot{k}"])).sum()
        loss.backward()
        for k in names:
            key = f"c{ci}_grad_{k}"
            if key not in z.files:
                assert c[k].grad is None or float(c[k].grad.abs().max()) == 0.0, (ci, k)
                continue
            ref = _t(z[key])
            got = c[k].grad if c[k].grad is not None else torch.zeros_like(ref)
            assert (got - ref).abs().max() <= 2e-5 * max(1.0, ref.abs().max().item()), (ci, k)
        for k in net:
            key = f"c{ci}_netgrad.{k}"
            if key not in z.files:
                assert net[k].grad is None, (ci, k)
                continue
            ref = _t(z[key])
            assert (net[k].grad - ref).abs().max() <= 2e-5 * max(1.0, ref.abs().max().item()), (ci, k)
        if step > 3000:
            assert float(net["linear.0.weight"].grad.abs().max()) > 0 and float(c["instances_embedding"].grad.abs().max()) > 0
            assert (c["means"].grad is None) == bool(stop)     # stop_optimizing_canonical_xyz detaches the canonical means


def test_densify_oracle_matches_reference_after_train():
    """oracle.densify.after_train == VanillaGaussians.after_train (vanilla.py:163-191) over three successive steps."""
    from oracle import densify as OD
    z = np.load(os.path.join(G, "densify.npz"))
    state = {}
    for s_ in range(z["radii"].shape[0]):
        OD.after_train(state, torch.from_numpy(z["radii"][s_]), torch.from_numpy(z["grads"][s_]), int(z["last_size"]))
        for k in ("xys_grad_norm", "vis_counts", "max_2Dsize"):
            assert torch.equal(state[k], torch.from_numpy(z[f"s{s_}_{k}"])), (s_, k)
