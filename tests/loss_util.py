"""Seeded inputs shared by the loss fixture generator (tests/golden/make_golden.py --losses) and the loss tests."""
import torch


def loss_inputs(seed: int, H: int, W: int):
    """Seeded synthetic render outputs + supervision of one view (shared by the fixture and the tests: only the
    reference's OUTPUTS are stored).  Contains the corner cases: colours above 1 (clamp), opacity exactly 0 / 1 and
    outside [1e-6, 1-1e-6], lidar holes, lidar beyond 80 m, predicted depth below 1e-4, masked ego-car rows."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    d = {}
    d["rgb"] = r(H, W, 3) * 1.15
    d["gt"] = (0.6 * d["rgb"].clamp(max=1.0) + 0.4 * r(H, W, 3)).clamp(0, 1)
    d["depth"] = 0.5 + 60.0 * r(H, W, 1)
    d["depth"][0, :5, 0] = 0.0
    d["depth"][1, :5, 0] = 5e-5
    d["alpha"] = r(H, W, 1).clamp(0.0, 1.0)
    d["alpha"][2, :4, 0] = 0.0
    d["alpha"][3, :4, 0] = 1.0
    d["alpha"][4, :4, 0] = 5e-7
    d["alpha"][5, :4, 0] = 1.0 - 1e-7
    d["sky"] = r(H, W, 3)
    d["sky_mask"] = (r(H, W) < 0.3).float()
    d["ego_mask"] = torch.zeros(H, W)
    d["ego_mask"][-4:, :] = 1.0
    lidar = d["depth"][..., 0] * (0.8 + 0.4 * r(H, W))
    lidar[r(H, W) < 0.6] = 0.0
    lidar[6, :6] = 95.0
    lidar[7, :6] = 0.005
    d["lidar"] = lidar
    return d


OMNIRE_KEYS = ("rgb_loss", "ssim_loss", "sky_loss_opacity", "depth_loss", "opacity_entropy_loss", "inverse_depth_smoothness_loss")
S3G_KEYS = ("Ll1", "ssim_loss", "sky_loss", "depth_loss", None, None)

# OmniRe-flavour cases: (name, oracle kwargs, product ImageLossConfig kwargs, use sky, use ego mask)
OMNIRE_CASES = [
    ("paper", dict(), dict(), True, True),
    ("safe_bce_l2norm", dict(opacity_loss_type="safe_bce", depth_loss_type="l2", depth_normalize=True, depth_inverse=False,
                             w_depth=0.3),
     dict(opacity_loss="safe_bce", depth_type="l2", depth_normalize=True, depth_inverse=False, w_depth=0.3), False, False),
    ("smooth_l1_norm_inv", dict(depth_loss_type="smooth_l1", depth_normalize=True, depth_inverse=True),
     dict(depth_type="smooth_l1", depth_normalize=True, depth_inverse=True), True, False),
]


def oracle_omnire(views, v_terms, use_sky, use_ego, okw):
    """Oracle terms [C,6] and gradients (renders [C,H,W,4], alphas [C,H,W,1], sky [C,H,W,3]) of sum(v_terms * terms)."""
    from oracle import losses as OL
    terms, gr, ga, gs = [], [], [], []
    for c, d in enumerate(views):
        renders = torch.cat([d["rgb"], d["depth"]], -1).requires_grad_(True)
        alphas = d["alpha"].clone().requires_grad_(True)
        sky = d["sky"].clone().requires_grad_(True) if use_sky else None
        out = OL.omnire_losses(renders, alphas, sky, d["gt"], d["sky_mask"], d["ego_mask"] if use_ego else None, d["lidar"], **okw)
        t = torch.stack([out[k] for k in OMNIRE_KEYS])
        (t * v_terms[c]).sum().backward()
        terms.append(t.detach()); gr.append(renders.grad); ga.append(alphas.grad)
        gs.append(sky.grad if use_sky else None)
    return torch.stack(terms), torch.stack(gr), torch.stack(ga), (torch.stack(gs) if use_sky else None)


def oracle_s3g(d, v_terms, use_sky, use_mask):
    from oracle import losses as OL
    color = d["rgb"].permute(2, 0, 1).contiguous().requires_grad_(True)
    depth = d["depth"].permute(2, 0, 1).contiguous().requires_grad_(True)
    weight = d["alpha"].permute(2, 0, 1).contiguous().requires_grad_(True)
    sky = d["sky"].permute(2, 0, 1).contiguous().requires_grad_(True) if use_sky else None
    gt = d["gt"].permute(2, 0, 1).contiguous()
    out = OL.s3g_losses(color, depth, weight, sky, gt, d["lidar"][None], d["sky_mask"][None].bool() if use_mask else None)
    z = torch.zeros(())
    t = torch.stack([out.get(k, z) if k else z for k in S3G_KEYS])
    (t * v_terms).sum().backward()
    g0 = lambda x: x.grad if x is not None and x.grad is not None else (torch.zeros_like(x) if x is not None else None)
    return t.detach(), g0(color), g0(depth), g0(weight), g0(sky)
