"""Shared helpers of the HexPlane tests: reference-layout planes -> the product's flat feature-last buffer."""
import itertools

import numpy as np
import torch

COMBS = list(itertools.combinations(range(4), 2))


def flatten_planes(grids):
    """grids[s][p] [1,F,H,W] -> (flat fp32 [total], offsets[S*6], reso[S*4]) in emd_b200's layout."""
    offsets, chunks, reso, total = [], [], [], 0
    for planes in grids:
        r = [0, 0, 0, 0]
        for (i, j), g in zip(COMBS, planes):
            _, F, H, W = g.shape
            r[i], r[j] = W, H
            offsets.append(total)
            chunks.append(g[0].permute(1, 2, 0).reshape(-1))
            total += H * W * F
        reso += r
    return torch.cat(chunks).contiguous(), np.asarray(offsets, np.int64), np.asarray(reso, np.int32)


def unflatten_like(flat, grids, offsets):
    """Flat feature-last buffer -> list of [1,F,H,W] tensors shaped like ``grids``."""
    out, k = [], 0
    for planes in grids:
        row = []
        for g in planes:
            _, F, H, W = g.shape
            o = int(offsets[k]); k += 1
            row.append(flat[o:o + H * W * F].view(H, W, F).permute(2, 0, 1)[None])
        out.append(row)
    return out


def oracle_run(grids, aabb, pts, t, cot):
    """Oracle features + gradients (points, times, planes)."""
    from oracle import hexplane as OH
    G = [[g.clone().requires_grad_(True) for g in row] for row in grids]
    p = pts.clone().requires_grad_(True)
    tt = t.clone().requires_grad_(True)
    feat = OH.hexplane_features(G, aabb, p, tt)
    (feat * cot).sum().backward()
    return feat.detach(), p.grad, tt.grad, [[g.grad for g in row] for row in G]
