"""S3Gaussian camera conventions for the diff_gauss tests (S3Gaussian/scene/cameras.py:55-66,
S3Gaussian/utils/graphics_utils.py:72-92)."""
import math

import torch

from emd_b200 import scenes


def s3g_camera(yaw=0.0, W=240, H=160, znear=0.01, zfar=100.0):
    c2w, K = scenes.camera(yaw, W, H)
    fovx = 2 * math.atan(W / (2 * float(K[0, 0])))
    fovy = 2 * math.atan(H / (2 * float(K[1, 1])))
    w2c = torch.linalg.inv(c2w)
    world_view = w2c.transpose(0, 1).contiguous()
    tx, ty = math.tan(fovx / 2), math.tan(fovy / 2)
    top, right = ty * znear, tx * znear
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (2 * right)
    P[1, 1] = 2.0 * znear / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    proj = P.transpose(0, 1)
    full = world_view @ proj
    campos = torch.linalg.inv(world_view)[3, :3]
    return dict(viewmatrix=world_view, projmatrix=full.contiguous(), campos=campos.contiguous(), tanfovx=tx, tanfovy=ty,
                W=W, H=H)
