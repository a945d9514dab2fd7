#!/usr/bin/env python
"""Host-side (Python / launch) profile of the bench step: cProfile over 30 steps after warm-up, top functions by own time.
The step is launch-bound as soon as the host needs longer than the GPU (8 ranks sharing one host's cores)."""
import cProfile
import io
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch  # noqa: E402

from emd_b200 import losses as LS, pipeline as P, scenes  # noqa: E402

dev = torch.device("cuda")
bg, rigid, smpl = P.make_street_scene(seed=0)
scene = P.StreetScene(bg, rigid, smpl, dev)
params = scene.parameters()
W, H = 960, 640
vm, Ks, c2w = scenes.cameras((0.0, 45.0, -45.0), W, H)
vm, Ks, c2w_d = vm.to(dev), Ks.to(dev), c2w.to(dev)
cams = c2w[:, :3, 3].tolist()
g = torch.Generator().manual_seed(1)
pix = torch.rand(3, H, W, 3, generator=g).to(dev)
sky = (torch.rand(3, H, W, generator=g) < 0.3).float().to(dev)
lidar = (40 * torch.rand(3, H, W, generator=g)).to(dev)
rgb_sky = torch.rand(3, H, W, 3, generator=g).to(dev).requires_grad_(True)
cfg = LS.ImageLossConfig.omnire(step=20000)


def step(i):
    for p in params:
        p.grad = None
    renders, alphas, info = scene.render_raw(c2w_d, Ks, W, H, (7 + 13 * i) % 150, 20000, viewmats=vm, cam_centers=cams)
    terms, _ = LS.image_losses_hwc(renders, alphas, pix, cfg, rgb_sky=rgb_sky, sky_masks=sky, lidar_depth_map=lidar)
    terms.sum().backward()


for i in range(8):
    step(i)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(30):
    step(8 + i)
torch.cuda.synchronize()
pr.disable()
out = io.StringIO()
st = pstats.Stats(pr, stream=out)
st.sort_stats("tottime").print_stats(45)
print(out.getvalue()[:9000])
