"""GPU parity: emd_b200.FusedAdam (emd_adam_step) against torch.optim.Adam run on the CPU with the reference's settings."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_adam_matches_torch_adam():
    from emd_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(0)
    shapes = [(70001, 3), (70001, 4), (70001, 1), (70001, 15, 3), (150, 32), (3, 37), (1,), (0, 3)] + [(17, 5)] * 30
    lrs = [1.6e-4, 1e-3, 5e-2, 1.25e-4, 1e-3, 5e-4, 1e-2, 1e-3] + [1e-3] * 30
    ref_p = [torch.randn(*s, generator=g).requires_grad_(True) for s in shapes]
    gpu_p = [p.detach().clone().cuda().requires_grad_(True) for p in ref_p]
    mk = lambda ps: [dict(params=[p], lr=lr, weight_decay=(0.01 if i == 5 else 0.0)) for i, (p, lr) in enumerate(zip(ps, lrs))]
    ref = torch.optim.Adam(mk(ref_p), lr=0.0, eps=1e-15)
    opt = FusedAdam(mk(gpu_p), lr=0.0, eps=1e-15)
    for step in range(1, 13):
        for i, (a, b) in enumerate(zip(ref_p, gpu_p)):
            if i == 6 and step % 2 == 0:          # a parameter without a gradient this step is skipped, like torch
                a.grad, b.grad = None, None
                continue
            gr = torch.randn(a.shape, generator=g) * 10.0 ** float(torch.randint(-4, 2, (1,), generator=g))
            if gr.dim() > 1 and gr.shape[0] > 100:
                gr[::3] = 0.0                     # invisible Gaussians: exact zero gradients
            a.grad, b.grad = gr.clone(), gr.cuda()
        for grp_r, grp_g in zip(ref.param_groups, opt.param_groups):          # scheduler writes group["lr"]
            grp_r["lr"] = grp_g["lr"] = grp_r["lr"] * 0.98
        ref.step()
        opt.step()
        for i, (a, b) in enumerate(zip(ref_p, gpu_p)):
            assert float((b.detach().cpu() - a.detach()).abs().max()) <= 1e-6 * step * max(1.0, float(a.abs().max())) \
                if a.numel() else True, (step, i)
    for a, b in zip(ref_p, gpu_p):
        if a.numel() == 0 or a not in ref.state:
            continue
        sa, sb = ref.state[a], opt.state[b]
        assert int(sa["step"]) == int(sb["step"])
        for k in ("exp_avg", "exp_avg_sq"):
            assert float((sb[k].cpu() - sa[k]).abs().max()) <= 2e-6 * max(1e-30, float(sa[k].abs().max()))
    # state_dict round-trips into torch.optim.Adam (checkpoint compatibility, base.py:640-660)
    sd = opt.state_dict()
    t2 = torch.optim.Adam(mk([p.detach().clone().requires_grad_(True) for p in gpu_p]), lr=0.0, eps=1e-15)
    t2.load_state_dict(sd)


def test_fused_adam_grad_scale_and_densify_state_surgery():
    """grad_scale folds 1/world into the update; replacing exp_avg / exp_avg_sq with resized tensors (densification,
    basics.py:196-240) is picked up on the next step."""
    from emd_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(1)
    p = torch.randn(1000, 3, generator=g).cuda().requires_grad_(True)
    q = p.detach().clone().requires_grad_(True)
    a, b = FusedAdam([p], lr=1e-2, eps=1e-15), FusedAdam([q], lr=1e-2, eps=1e-15)
    gr = torch.randn(1000, 3, generator=g).cuda()
    p.grad, q.grad = gr * 8.0, gr.clone()
    a.step(grad_scale=0.125)
    b.step()
    assert torch.equal(p.detach(), q.detach())
    st = b.state[q]
    keep = torch.arange(1000, device="cuda") % 2 == 0
    new_q = torch.nn.Parameter(q.detach()[keep].clone())
    b.param_groups[0]["params"] = [new_q]
    b.state[new_q] = {"step": st["step"], "exp_avg": st["exp_avg"][keep].clone(), "exp_avg_sq": st["exp_avg_sq"][keep].clone()}
    del b.state[q]
    new_q.grad = torch.ones_like(new_q)
    b.step()
    assert int(b.state[new_q]["step"]) == 2 and torch.isfinite(new_q).all()
