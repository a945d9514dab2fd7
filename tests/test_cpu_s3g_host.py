"""CPU: host logic of the S3Gaussian step that needs no kernel -- the regulariser terms of ``training_losses`` computed
from the fused residual kernel's sums (``ddict["reg_sums"]``) equal the element-wise ``abs().mean()`` statement of
``S3Gaussian/train.py:240-305``, including their gradients through the sums; and the committed ncu launch lists parse."""
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ddict(n, g, with_sums):
    from emd_b200.emd_s3g import REG_KEYS
    dd = {b: dict(dx=torch.randn(n, 3, generator=g).requires_grad_(True), do=torch.randn(n, 1, generator=g).requires_grad_(True),
                  dshs=torch.randn(n, 16, 3, generator=g).requires_grad_(True)) for b in ("coarse", "fine")}
    if with_sums:   # what emd_s3g_apply_fwd returns, stated with torch
        dd["reg_sums"] = torch.stack([dd[b][k].abs().sum() for b, k in REG_KEYS])
    return dd


def test_regulariser_terms_from_fused_sums(monkeypatch):
    from emd_b200 import s3g_render as SR
    monkeypatch.setattr(SR, "s3g_image_losses", lambda *a, **k: {})
    args = SR.S3GOptions(lambda_dx=0.01, lambda_do=0.02, lambda_dshs=0.03, lambda_f2c=0.005, feat_head=False)
    pkg = dict(color=None, depth=None, weight=None, sky_color=None)
    outs = []
    for with_sums in (False, True):
        dd = _ddict(257, torch.Generator().manual_seed(3), with_sums)
        terms = SR.training_losses(args, dict(pkg, ddict=dd), None, None, None)
        assert set(terms) == {"dx_loss", "do_loss", "dshs_loss"}
        sum(terms.values()).backward()
        outs.append((terms, dd))
    (ta, da), (tb, db) = outs
    for k in ta:
        assert torch.allclose(ta[k], tb[k], rtol=1e-6, atol=0), k
    for b in ("coarse", "fine"):
        for k in ("dx", "do", "dshs"):
            assert torch.allclose(da[b][k].grad, db[b][k].grad, rtol=1e-6, atol=1e-12), (b, k)
    # a branch switched off drops its term in both statements
    args2 = SR.S3GOptions(no_fine_deform=True, feat_head=False)
    dd = _ddict(16, torch.Generator().manual_seed(4), True)
    t2 = SR.training_losses(args2, dict(pkg, ddict=dd), None, None, None)
    assert torch.allclose(t2["dshs_loss"], dd["coarse"]["dshs"].abs().mean() * args2.lambda_dshs)


def test_committed_launch_lists_parse():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import launch_table
    finally:
        sys.path.pop(0)
    for name in ("r03f_ncu_s3g_launches.csv", "r03k_ncu_s3g_launches.csv"):
        rows = launch_table.rows_of(os.path.join(ROOT, "profiles", name))
        assert len(rows) >= 300 and all(t > 0 for _, t in rows)
        names = {launch_table.short(n) for n, _ in rows}
        assert any("hexplane_bwd_kernel" in n for n in names) and any("raster_bwd_kernel" in n for n in names)
    after = {launch_table.short(n) for n, _ in launch_table.rows_of(os.path.join(ROOT, "profiles", "r03k_ncu_s3g_launches.csv"))}
    assert any("linear_wgrad_tc_bulk_kernel" in n for n in after) and any("s3g_apply_fwd_kernel" in n for n in after)


def test_s3g_mlp_layer_table_matches_the_network():
    """bench.py's layer table (the algorithmic bytes of the S3G roofline's HBM view) against the shapes
    ``S3GDeformation`` actually launches: same 20 (K, Nout, relu_in, relu_out), recorded by a stand-in ``linear``."""
    import importlib.util
    import numpy as np
    from emd_b200 import emd_s3g as E
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    z = np.load(os.path.join(ROOT, "tests", "golden", "emd_s3g.npz"))
    pre = "w.deformation_net."
    w = {k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)}
    seen = []

    def fake_linear(X, W, b, relu_in=False, relu_out=False):
        seen.append((X.shape[1], W.shape[0], bool(relu_in), bool(relu_out)))
        x = torch.relu(X) if relu_in else X
        y = torch.nn.functional.linear(x, W, b)
        return torch.relu(y) if relu_out else y

    def fake_temb(table, t, cur):
        return torch.zeros(table.shape[1])

    orig = (E.linear, E.temporal_embed)
    E.linear, E.temporal_embed = fake_linear, fake_temb
    try:
        n = 7
        net = E.S3GDeformation(w)
        net(torch.zeros(n, 3), torch.zeros(n, 3), torch.zeros(n, 4), torch.zeros(n, 1), torch.zeros(n, 16, 3), 0.1,
            torch.zeros(n, 4), 20000, 0, hex_feat=torch.zeros(n, 128))
    finally:
        E.linear, E.temporal_embed = orig
    assert sorted(seen) == sorted(bench.s3g_mlp_layers())
    fwd, bwd = bench.s3g_mlp_bytes_per_gaussian()
    assert fwd == sum(4 * (k + n_) for k, n_, _, _ in seen) and bwd > 2 * fwd
