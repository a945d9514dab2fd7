"""GPU parity: DeformableNodes deformation network (K1g) vs the golden vectors of the reference's own
ConditionalDeformNetwork and vs the oracle: the strided GEMM kernels alone, the whole network at the config's size
(D = 8, W = 256, embed 16), and DeformableNodes.get_gaussians end to end."""
import ctypes
import os

import numpy as np
import pytest
import torch

from tests.util import rel_err, rel_l2

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("M,K,Nout,ldx,ldy,relu", [(3000, 100, 256, 100, 256, 1), (5171, 356, 256, 356, 356, 1),
                                                    (1300, 256, 7, 256, 7, 0), (1, 20, 33, 24, 40, 1), (129, 17, 130, 17, 131, 0)])
def test_dense_kernels_vs_torch(M, K, Nout, ldx, ldy, relu):
    """emd_dense_fwd / emd_dense_bwd through the C ABI against fp64 torch: strided operands, output window, dgrad column
    window + ReLU mask, split-K weight gradient, bias gradient; the weight gradient twice (bit-reproducible)."""
    from emd_b200 import _C
    L = _C.lib()
    dev = "cuda"
    g = torch.Generator().manual_seed(M + K)
    Xb = torch.randn(M, ldx, generator=g).to(dev)
    W, b = (torch.randn(Nout, K, generator=g) / K ** 0.5).to(dev), torch.randn(Nout, generator=g).to(dev)
    Yb = torch.full((M, ldy), 123.0, device=dev)
    off = ldy - Nout
    st = _C.stream()
    _C.check(L.emd_dense_fwd(_C.ptr(Xb), ldx, _C.ptr(W), _C.ptr(b), M, K, Nout, relu, Yb.data_ptr() + 4 * off, ldy, st), "fwd")
    ref = Xb[:, :K].double() @ W.double().T + b.double()
    ref = torch.relu(ref) if relu else ref
    assert (Yb[:, off:].double() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    assert bool((Yb[:, :off] == 123.0).all())
    dZ = torch.randn(M, Nout, generator=g).to(dev)
    col0, ncols = (K - 16, 16) if K >= 32 else (0, K)
    mask = torch.randn(M, ncols, generator=g).to(dev)
    dX = torch.full((M, ncols + 3), -5.0, device=dev)
    ws_bytes = L.emd_dense_bwd_workspace_bytes(M, K, Nout)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    outs = []
    for _ in range(2):
        dW, db = torch.empty(Nout, K, device=dev), torch.empty(Nout, device=dev)
        _C.check(L.emd_dense_bwd(_C.ptr(Xb), ldx, _C.ptr(W), _C.ptr(dZ), Nout, M, K, Nout, _C.ptr(dX), ncols + 3, col0, ncols,
                                 _C.ptr(mask), ncols, _C.ptr(dW), _C.ptr(db), _C.ptr(ws), ws_bytes, st), "bwd")
        outs.append((dW.clone(), db.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    ref_dX = (dZ.double() @ W.double()[:, col0:col0 + ncols]) * (mask > 0)
    assert (dX[:, :ncols].double() - ref_dX).abs().max().item() <= 2e-5 * max(1.0, ref_dX.abs().max().item())
    assert bool((dX[:, ncols:] == -5.0).all())
    ref_dW = dZ.double().T @ Xb[:, :K].double()
    assert (dW.double() - ref_dW).abs().max().item() <= 3e-5 * max(1.0, ref_dW.abs().max().item())
    ref_db = dZ.double().sum(0)
    assert (db.double() - ref_db).abs().max().item() <= 3e-5 * max(1.0, ref_db.abs().max().item())


def _as_nodes_input(z):
    """The golden network takes (x, t, condition) directly; express them as DeformableNodes inputs: one instance per point
    would be wasteful, so instance height 2 (x = means / 2 * 2 = means) and a per-point 'instance' embedding."""
    x, cond = torch.from_numpy(z["net_x"]), torch.from_numpy(z["net_cond"])
    N = x.shape[0]
    return x, cond, torch.arange(N)[:, None], torch.full((N, 3), 2.0), float(z["net_t"][0, 0])


def test_conditional_deform_network_golden():
    from emd_b200.deformable import deform_canonical
    z = np.load(f"{G}/omnire_modules.npz")
    dev = "cuda"
    x, cond, ids, size, t = _as_nodes_input(z)
    sd = {k[len("net_sd."):]: torch.from_numpy(z[k]).to(dev).requires_grad_(True) for k in z.files if k.startswith("net_sd.")}
    emb = cond.to(dev).requires_grad_(True)
    quats = torch.zeros(x.shape[0], 4); quats[:, 0] = 1.0
    m, q = deform_canonical(x.to(dev), quats.to(dev), emb, ids.to(dev), size.to(dev), t, sd, D=8)
    d_xyz, rot = m - x.to(dev), q - quats.to(dev)
    assert (d_xyz.detach().cpu() - torch.from_numpy(z["net_d_xyz"])).abs().max().item() <= 5e-6
    assert (rot.detach().cpu() - torch.from_numpy(z["net_rot"])).abs().max().item() <= 5e-6
    ((m * torch.from_numpy(z["net_c1"]).to(dev)).sum() + (q * torch.from_numpy(z["net_c2"]).to(dev)).sum()).backward()
    ref = torch.from_numpy(z["net_v_cond"])
    assert (emb.grad.cpu() - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())
    for k, p in sd.items():
        ref = torch.from_numpy(z[f"net_grad.{k}"])
        assert (p.grad.cpu() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item()), k


def _network(D, Wd, E, g, quat_head=True):
    Kin = 3 + 60 + 1 + 20 + E
    sd = {}
    for i in range(D):
        K = Kin if i == 0 else (Kin + Wd if i - 1 == D // 2 else Wd)
        sd[f"linear.{i}.weight"] = torch.randn(Wd, K, generator=g) * (1.4 / K ** 0.5)
        sd[f"linear.{i}.bias"] = 0.1 * torch.randn(Wd, generator=g)
    sd["gaussian_warp.weight"], sd["gaussian_warp.bias"] = 0.3 * torch.randn(3, Wd, generator=g) / Wd ** 0.5, 0.01 * torch.randn(3, generator=g)
    if quat_head:
        sd["gaussian_rotation.weight"] = 0.3 * torch.randn(4, Wd, generator=g) / Wd ** 0.5
        sd["gaussian_rotation.bias"] = 0.01 * torch.randn(4, generator=g)
    return sd


@pytest.mark.parametrize("frame,step,stop_xyz", [(13, 8000, True), (0, 3001, False), (5, 100, True)])
def test_deformable_nodes_get_gaussians(frame, step, stop_xyz):
    """DeformableNodes.get_gaussians at the config's network size (omnire.yaml:159-166): deformation network -> rigid EMD
    transform -> activations + SH, values and every gradient (network, instance embedding, Gaussians, EMD heads, poses)
    against the oracle; step 100 <= use_deformgs_after takes the plain rigid route."""
    from emd_b200.deformable import DeformableNodesEMD
    from oracle import deform_network as ON
    from oracle import emd_rigid as ER
    from tests.test_gpu_emd_rigid import HEADS, _setup
    I, pts, frames, D, Wd, E = 6, 500, 40, 8, 256, 16
    rs, cpu, p, g = _setup(11, I, pts, frames)
    size = torch.tensor([0.8, 0.8, 1.7]) + 0.2 * torch.rand(I, 3, generator=g)
    cpu["instances_embedding"] = torch.rand(I, E, generator=g).requires_grad_(True)
    sd = _network(D, Wd, E, g)
    net_cpu = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ts = torch.linspace(0, 1, frames).tolist()
    cam_pos = torch.tensor([0.3, -0.2, 1.6])
    if step > 3000:
        m_ref, q_ref = ON.deformed_canonical(net_cpu, cpu["means"], cpu["quats"], rs.point_ids, size, cpu["instances_embedding"],
                                             ts[frame], D=D, stop_optimizing_canonical_xyz=stop_xyz)
    else:
        m_ref, q_ref = cpu["means"], cpu["quats"]
    ref = ER.get_gaussians(p, m_ref, q_ref, cpu["scales"], cpu["opacities"], cpu["features_dc"], cpu["features_rest"], frame,
                           step, cam_pos)
    dev = "cuda"
    gpu = {k: v.detach().to(dev).requires_grad_(True) for k, v in cpu.items()}
    net_gpu = {k: v.detach().to(dev).requires_grad_(True) for k, v in sd.items()}
    node = DeformableNodesEMD(
        dict(_means=gpu["means"], _quats=gpu["quats"], _scales=gpu["scales"], _opacities=gpu["opacities"],
             _features_dc=gpu["features_dc"], _features_rest=gpu["features_rest"], _embeddings=gpu["embeddings"],
             point_ids=rs.point_ids.to(dev), weight=gpu["weight"], instances_quats=gpu["instances_quats"],
             instances_trans=gpu["instances_trans"], instances_fv=rs.instances_fv.to(dev),
             instances_embedding=gpu["instances_embedding"], instances_size=size.to(dev)),
        {k: gpu[k] for k in HEADS}, net_gpu, ts, D=D, stop_optimizing_canonical_xyz=stop_xyz)
    out = node.get_gaussians(cam_pos.tolist(), frame, step)
    cot = {}
    for k in ("_means", "_opacities", "_rgbs", "_scales", "_quats"):
        assert out[k].shape == ref[k].shape, k
        err = float((out[k].detach().cpu() - ref[k].detach()).abs().max())
        tol = 3e-5 * max(1.0, float(ref[k].detach().abs().max()))
        assert err <= tol, f"{k}: {err} > {tol}"
        cot[k] = torch.randn(ref[k].shape, generator=g)
    sum((ref[k] * cot[k]).sum() for k in cot).backward()
    sum((out[k] * cot[k].to(dev)).sum() for k in cot).backward()
    pairs = [(k, cpu[k].grad, gpu[k].grad) for k in cpu] + [(k, net_cpu[k].grad, net_gpu[k].grad) for k in sd]
    for k, gr, gg in pairs:
        if gr is None or float(gr.abs().max()) == 0.0:
            assert gg is None or float(gg.abs().max()) == 0.0, k
            continue
        assert gg is not None, k
        e, l2 = rel_err(gg, gr), rel_l2(gg, gr)
        assert e <= 1e-3 and l2 <= 1e-3, f"grad {k}: max-rel {e}, l2-rel {l2}"
    if step > 3000:
        assert float(net_cpu["linear.0.weight"].grad.abs().max()) > 0 and float(cpu["instances_embedding"].grad.abs().max()) > 0
        assert node._gs_cache["local_xyz_deformed"] is not None
