"""Shared helpers for the parity tests (oracle = checker, CUDA path = thing checked)."""
import math

import torch

from emd_b200 import scenes


def raster_scene(seed=0, n=3000, width=240, height=160, yaws=(0.0,), depth=(2.0, 40.0), scale=0.05):
    g = torch.Generator().manual_seed(seed)
    sc = scenes.simple_gaussians(n, g, width, height, depth=depth, scale=scale)
    viewmats, Ks, c2w = scenes.cameras(yaws, width, height)
    return sc, viewmats, Ks, c2w, g


def bits(t):
    return t.detach().cpu().contiguous().view(torch.int32)


def rel_err(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    denom = b.abs().max().clamp(min=1e-30)
    return float((a - b).abs().max() / denom)


def rel_l2(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))
