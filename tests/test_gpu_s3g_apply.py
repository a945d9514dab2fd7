"""GPU parity of the fused residual application of the S3Gaussian deformation (``emd_s3g_apply_fwd`` / ``_bwd`` through
``emd_b200.emd_s3g.apply_residuals``) against the element-wise statement of the reference: the sums of
``deform_network.forward`` (``S3Gaussian/scene/deformation.py:439-481``), ``get_features``' concatenation and the
``abs().mean()`` regularisers of ``train.py:240-305`` -- plain torch, evaluated on the CPU in float64 for the sums."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(N, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    x = dict(point=r(N, 3), opacity=r(N, 1), dc=r(N, 1, 3), rest=r(N, 15, 3))
    dd = {b: dict(dx=0.1 * r(N, 3), do=0.1 * r(N, 1), dshs=0.1 * r(N, 16, 3)) for b in ("coarse", "fine")}
    if N:
        dd["coarse"]["dx"][0] = 0.0          # d|x|/dx at 0 is 0 (torch.abs' backward)
        dd["fine"]["dshs"][N // 2] = 0.0
    cot = (r(N, 3), r(N, 1), r(N, 16, 3))
    lam = torch.tensor([0.3, 0.7, 1.1, 0.2, 0.9, 0.5])
    return x, dd, cot, lam


@pytest.mark.parametrize("N", [0, 1, 15, 4097])
def test_apply_residuals_matches_elementwise(N):
    from emd_b200.emd_s3g import REG_KEYS, apply_residuals
    x, dd, cot, lam = _case(N, 3 + N)
    # reference statement (CPU)
    xr = {k: v.clone().requires_grad_(True) for k, v in x.items()}
    ddr = {b: {k: v.clone().requires_grad_(True) for k, v in d.items()} for b, d in dd.items()}
    means = xr["point"] + ddr["coarse"]["dx"] + ddr["fine"]["dx"]
    opac = xr["opacity"] + ddr["coarse"]["do"] + ddr["fine"]["do"]
    shs = torch.cat((xr["dc"], xr["rest"]), dim=1) + ddr["coarse"]["dshs"] + ddr["fine"]["dshs"]
    sums = torch.stack([ddr[b][k].double().abs().sum() for b, k in REG_KEYS]).float()
    ((means * cot[0]).sum() + (opac * cot[1]).sum() + (shs * cot[2]).sum() + (sums * lam).sum()).backward()
    # fused kernels
    xg = {k: v.cuda().requires_grad_(True) for k, v in x.items()}
    ddg = {b: {k: v.cuda().requires_grad_(True) for k, v in d.items()} for b, d in dd.items()}
    m, o, s, sm = apply_residuals(xg["point"], xg["opacity"], xg["dc"], xg["rest"], ddg)
    assert m.shape == (N, 3) and o.shape == (N, 1) and s.shape == (N, 16, 3) and sm.shape == (6,)
    assert torch.equal(m.cpu(), means.detach()) and torch.equal(o.cpu(), opac.detach()) and torch.equal(s.cpu(), shs.detach())
    assert torch.allclose(sm.cpu(), sums.detach(), rtol=2e-6, atol=1e-30)
    ((m * cot[0].cuda()).sum() + (o * cot[1].cuda()).sum() + (s * cot[2].cuda()).sum() + (sm * lam.cuda()).sum()).backward()
    for k in x:
        assert torch.equal(xg[k].grad.cpu(), xr[k].grad), k
    for b in dd:
        for k in dd[b]:
            assert torch.equal(ddg[b][k].grad.cpu(), ddr[b][k].grad), (b, k)


def test_apply_residuals_without_regulariser_cotangent():
    """Only the image path contributes (v_sums None / unused): the residual gradients equal the cotangents."""
    from emd_b200.emd_s3g import apply_residuals
    x, dd, cot, _ = _case(300, 9)
    xg = {k: v.cuda().requires_grad_(True) for k, v in x.items()}
    ddg = {b: {k: v.cuda().requires_grad_(True) for k, v in d.items()} for b, d in dd.items()}
    m, o, s, _ = apply_residuals(xg["point"], xg["opacity"], xg["dc"], xg["rest"], ddg)
    (s * cot[2].cuda()).sum().backward()       # means / opacity unused: NULL cotangents inside the kernel
    assert torch.equal(ddg["fine"]["dshs"].grad.cpu(), cot[2]) and torch.equal(xg["rest"].grad.cpu(), cot[2][:, 1:])
    assert torch.equal(ddg["coarse"]["dx"].grad.cpu(), torch.zeros(300, 3))
