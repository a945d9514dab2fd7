"""CPU: the host-side orchestration of the SURVEY 8f-4 components (emd_b200/deformable.py, voxel_deformer.py) -- layer
chaining, column windows of the skip buffer, which gradient goes where -- run against the oracle with the C-ABI entry
points replaced by host stand-ins: the GEMMs and the voxel lookup go through the SAME tile / tap code the CUDA kernels
include (libemd_b200_hostmath.so), the three small element-wise kernels are restated in numpy from their documented
contract.  This is test scaffolding only: the product path has no CPU route (``_C.ptr`` rejects CPU tensors), so the
test patches ``_C`` for its own duration.  The kernels themselves are checked on the GPU (tests/test_gpu_deformable.py).
"""
import ctypes

import numpy as np
import pytest
import torch

from emd_b200 import _C, build

P = ctypes.c_void_p
I64, I32, F32 = ctypes.c_int64, ctypes.c_int, ctypes.c_float


def _arr(ptr, shape, dtype=np.float32):
    """numpy view of raw memory"""
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype)
    ct = {np.float32: ctypes.c_float, np.int64: ctypes.c_int64, np.uint8: ctypes.c_uint8}[dtype]
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ct)), shape=(n,)).reshape(shape)


def _strided(ptr, rows, cols, ld):
    """[rows, cols] window of a row-major buffer with row stride ld"""
    flat = _arr(ptr, ((rows - 1) * ld + cols,)) if rows > 0 else np.zeros((0,), np.float32)
    return np.lib.stride_tricks.as_strided(flat, shape=(rows, cols), strides=(4 * ld, 4))


class FakeLib:
    """Host stand-ins for the entry points deformable.py / voxel_deformer.py call."""

    def __init__(self):
        self.h = ctypes.CDLL(str(build.build_hostmath()))
        self.h.emd_host_dense_fwd.argtypes = [P, I64, P, P, I64, I32, I32, I32, P, I64]
        self.h.emd_host_dense_fwd.restype = None
        self.h.emd_host_dense_bwd.argtypes = [P, I64, P, P, I64, I64, I32, I32, P, I64, I32, I32, P, I64, P, P]
        self.h.emd_host_dense_bwd.restype = None
        self.h.emd_host_voxel_lbs.argtypes = [P] * 4 + [F32] + [I32] * 6 + [P, I64] + [P] * 4
        self.h.emd_host_voxel_lbs.restype = None
        self.calls = []

    def emd_dense_fwd(self, X, ldx, W, b, M, K, Nout, relu, Y, ldy, stream):
        self.calls.append("dense_fwd")
        self.h.emd_host_dense_fwd(X, ldx, W, b, M, K, Nout, relu, Y, ldy)
        return 0

    def emd_dense_bwd_workspace_bytes(self, M, K, Nout):
        return 16

    def emd_dense_bwd(self, X, ldx, W, dZ, lddz, M, K, Nout, dX, lddx, col0, ncols, mask, ldmask, dW, db, ws, wsb, stream):
        self.calls.append("dense_bwd")
        self.h.emd_host_dense_bwd(X, ldx, W, dZ, lddz, M, K, Nout, dX, lddx, col0, ncols, mask, ldmask, dW, db)
        return 0

    def emd_deform_input_fwd(self, means, ids, size, emb, t, xm, tm, E, N, out0, ld0, out1, ld1, stream):
        m, idv = _arr(means, (N, 3)), _arr(ids, (N,), np.int64)
        I = int(idv.max()) + 1
        h = _arr(size, (I, 3))[idv, 2]
        x = (m / h[:, None] * np.float32(2)).astype(np.float32)
        cols = [x]
        for f in range(xm):
            a = x * np.float32(2 ** f)
            cols += [np.sin(a), np.cos(a)]
        tt = np.full((N, 1), t, np.float32)
        cols.append(tt)
        for f in range(tm):
            a = tt * np.float32(2 ** f)
            cols += [np.sin(a), np.cos(a)]
        if E:
            cols.append(_arr(emb, (I, E))[idv])
        row = np.concatenate(cols, axis=1).astype(np.float32)
        _strided(out0, N, row.shape[1], ld0)[:] = row
        if out1:
            _strided(out1, N, row.shape[1], ld1)[:] = row
        return 0

    def emd_deform_apply_fwd(self, means, quats, d, dcols, N, means_out, quats_out, stream):
        dd = _arr(d, (N, dcols))
        _arr(means_out, (N, 3))[:] = _arr(means, (N, 3)) + dd[:, :3]
        q = _arr(quats, (N, 4))
        qn = q / np.linalg.norm(q, axis=1, keepdims=True)
        _arr(quats_out, (N, 4))[:] = qn + (dd[:, 3:7] if dcols >= 7 else 0)
        return 0

    def emd_deform_apply_bwd(self, quats, v_mo, v_qo, dcols, N, v_d, v_means, v_quats, stream):
        vd = _arr(v_d, (N, dcols))
        vd[:] = 0
        vd[:, :3] = _arr(v_mo, (N, 3))
        go = _arr(v_qo, (N, 4)) if v_qo else np.zeros((N, 4), np.float32)
        if dcols >= 7:
            vd[:, 3:7] = go
        if v_means:
            _arr(v_means, (N, 3))[:] = _arr(v_mo, (N, 3))
        if v_quats:
            q = _arr(quats, (N, 4))
            nrm = np.linalg.norm(q, axis=1, keepdims=True)
            u = q / nrm
            _arr(v_quats, (N, 4))[:] = (go - u * (u * go).sum(1, keepdims=True)) / nrm
        return 0

    def emd_deform_embed_grad_workspace_bytes(self, I, E, max_pts):
        return 16

    def emd_deform_embed_grad(self, g0, g1, E, order, seg_start, I, max_pts, ws, wsb, v_emb, stream):
        seg = _arr(seg_start, (I + 1,), np.int64)
        N = int(seg[-1])
        od = _arr(order, (N,), np.int64)
        g = _arr(g0, (N, E)) + (_arr(g1, (N, E)) if g1 else 0)
        out = _arr(v_emb, (I, E))
        for i in range(I):
            assert seg[i + 1] - seg[i] <= max_pts
            out[i] = g[od[seg[i]:seg[i + 1]]].sum(0)
        return 0

    def emd_rigid_chunk_size(self):
        return 1024

    def emd_voxel_lbs_fwd(self, base, corr, offset, scale, ratio, ratio_dim, B, D, H, W, J, xc, V, out, stream):
        self.h.emd_host_voxel_lbs(base, corr, offset, scale, ratio, ratio_dim, B, D, H, W, J, xc, V, out, None, None, None)
        return 0

    def emd_voxel_lbs_bwd(self, base, corr, offset, scale, ratio, ratio_dim, B, D, H, W, J, xc, V, v_out, v_corr, v_xc, stream):
        scratch = np.zeros((B, V, J), np.float32)
        self.h.emd_host_voxel_lbs(base, corr, offset, scale, ratio, ratio_dim, B, D, H, W, J, xc, V, scratch.ctypes.data_as(P),
                                  v_out, v_corr, v_xc)
        return 0


@pytest.fixture
def fake_c(monkeypatch):
    fake = FakeLib()
    monkeypatch.setattr(_C, "lib", lambda: fake)
    monkeypatch.setattr(_C, "ptr", lambda t, dtype=None, name="tensor": None if t is None else t.data_ptr())
    monkeypatch.setattr(_C, "stream", lambda: 0)
    monkeypatch.setattr(_C, "check", lambda status, what: None if status == 0 else (_ for _ in ()).throw(RuntimeError(what)))
    return fake


def _network(D, Wd, E, g, quat_head=True):
    Kin = 3 + 60 + 1 + 20 + E
    sd = {}
    for i in range(D):
        K = Kin if i == 0 else (Kin + Wd if i - 1 == D // 2 else Wd)
        sd[f"linear.{i}.weight"] = torch.randn(Wd, K, generator=g) * (1.5 / K ** 0.5)
        sd[f"linear.{i}.bias"] = 0.1 * torch.randn(Wd, generator=g)
    Kh = Kin + Wd if D - 1 == D // 2 else Wd
    sd["gaussian_warp.weight"], sd["gaussian_warp.bias"] = torch.randn(3, Kh, generator=g) / Kh ** 0.5, 0.1 * torch.randn(3, generator=g)
    if quat_head:
        sd["gaussian_rotation.weight"] = torch.randn(4, Kh, generator=g) / Kh ** 0.5
        sd["gaussian_rotation.bias"] = 0.1 * torch.randn(4, generator=g)
    return sd


@pytest.mark.parametrize("D,Wd,E,quat_head,stop_xyz", [(8, 40, 16, True, True), (4, 24, 8, False, False), (1, 16, 4, True, True)])
def test_deform_canonical_orchestration(fake_c, D, Wd, E, quat_head, stop_xyz):
    from emd_b200.deformable import deform_canonical
    from oracle import deform_network as ON
    g = torch.Generator().manual_seed(D * 100 + Wd)
    N, I = 203, 5
    ids = torch.randint(0, I, (N, 1), generator=g)
    ids[:I, 0] = torch.arange(I)
    size = torch.rand(I, 3, generator=g) + 0.8
    means = (torch.rand(N, 3, generator=g) - 0.5) * size[ids[:, 0]]
    quats = torch.randn(N, 4, generator=g)
    emb = torch.rand(I, E, generator=g)
    sd = _network(D, Wd, E, g, quat_head)
    t = 0.3125
    leaves_o = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    mo, qo, eo = means.clone().requires_grad_(True), quats.clone().requires_grad_(True), emb.clone().requires_grad_(True)
    rm, rq = ON.deformed_canonical(leaves_o, mo, qo, ids, size, eo, t, D=D, stop_optimizing_canonical_xyz=stop_xyz)
    c1, c2 = torch.randn(N, 3, generator=g), torch.randn(N, 4, generator=g)
    ((rm * c1).sum() + (rq * c2).sum()).backward()
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    m, q, e = means.clone().requires_grad_(True), quats.clone().requires_grad_(True), emb.clone().requires_grad_(True)
    gm, gq = deform_canonical(m, q, e, ids, size, t, leaves, D=D, stop_optimizing_canonical_xyz=stop_xyz)
    assert (gm - rm).abs().max() <= 2e-5 and (gq - rq).abs().max() <= 2e-5
    ((gm * c1).sum() + (gq * c2).sum()).backward()
    assert fake_c.calls.count("dense_fwd") == D + 1

    def close(a, b, what):
        assert a is not None, what
        assert (a - b).abs().max() <= 1e-4 * max(1.0, b.abs().max().item()), what

    for k in sd:
        close(leaves[k].grad, leaves_o[k].grad, k)
    close(e.grad, eo.grad, "instances_embedding")
    close(q.grad, qo.grad, "quats")
    if stop_xyz:
        assert m.grad is None and mo.grad is None
    else:
        close(m.grad, mo.grad, "means")


def test_voxel_deformer_orchestration(fake_c):
    """VoxelDeformer mirror: layout conversion, ratio_dim convention, gradient routing (correction volume and points),
    regularisers, checkpoint round trip -- against the golden vectors of the reference's own VoxelDeformer."""
    import os
    from emd_b200.voxel_deformer import VoxelDeformer
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "omnire_modules.npz"))
    t = lambda k: torch.from_numpy(z[k])  # noqa: E731
    vd = VoxelDeformer(t("vox_base"), t("vox_offset"), t("vox_scale"), [int(v) for v in z["vox_res"]],
                       voxel_w_correction_ref=t("vox_corr"))
    assert abs(vd.ratio - float(z["vox_ratio"])) < 1e-7 and vd.ratio_dim == int(z["vox_ratio_dim"])
    xc = t("vox_xc").clone().requires_grad_(True)
    w = vd(xc)
    assert (w - t("vox_w")).abs().max() <= 2e-6
    (w * t("vox_cot")).sum().backward()
    from emd_b200.voxel_deformer import to_reference_layout
    assert (to_reference_layout(vd.voxel_w_correction.grad) - t("vox_v_corr")).abs().max() <= 2e-6
    assert (xc.grad - t("vox_v_xc")).abs().max() <= 2e-5 * max(1.0, t("vox_v_xc").abs().max().item())
    assert torch.allclose(vd.get_tv("dc"), t("vox_tv"), rtol=1e-5) and torch.allclose(vd.get_mag("dc"), t("vox_mag"), rtol=1e-5)
    assert torch.equal(vd.get_voxel_weight, t("vox_base") + t("vox_corr"))
    xn = vd.normalize(t("vox_xc"))
    assert (vd.denormalize(xn) - t("vox_xc")).abs().max() <= 1e-5
    sd = vd.reference_state()
    assert torch.equal(sd["lbs_voxel_base"], t("vox_base")) and torch.equal(sd["voxel_w_correction"], t("vox_corr"))
