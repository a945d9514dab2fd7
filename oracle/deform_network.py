"""ORACLE (test infrastructure, never on the product path): CPU restatement of the non-rigid deformation network of
OmniRe's DeformableNodes -- ``Embedder`` / ``get_embedder`` (``OmniRe/models/modules.py:318-366``),
``ConditionalDeformNetwork`` (``:411-457``), ``DeformableNodes.get_deformation`` and the deformation part of
``get_gaussians`` (``OmniRe/models/nodes/deformable.py:35-68``).

Pinned: ``tests/golden/omnire_modules.npz`` holds outputs and parameter gradients of the reference's own
``ConditionalDeformNetwork`` (``tests/golden/make_golden.py --modules``), and ``tests/golden/deformable_nodes.npz`` the
outputs and every gradient of the reference's own ``DeformableNodes.get_gaussians`` (``--deformable``: the class itself,
third-party imports stubbed); ``tests/test_cpu_golden.py`` checks this file -- composed with ``oracle/emd_rigid.py`` as
``deformable.py`` composes with ``RigidNodes`` -- against both.  Parameters are passed as the module's ``state_dict`` (``linear.{i}.weight``, ``gaussian_warp.weight`` ...).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor


def posenc(x: Tensor, multires: int) -> Tensor:
    """include_input + log-sampled bands 2^0 .. 2^(multires-1), [sin, cos] per band (modules.py:341-366)."""
    out = [x]
    freqs = 2.0 ** torch.linspace(0.0, float(multires - 1), steps=multires)
    for f in freqs:
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, dim=-1)


def conditional_deform_network(sd: Dict[str, Tensor], x: Tensor, t: Tensor, condition: Tensor, D: int = 8,
                               x_multires: int = 10, t_multires: int = 10
                               ) -> Tuple[Tensor, Optional[Tensor], Optional[Tensor]]:
    """``ConditionalDeformNetwork.forward`` (modules.py:436-457) -> (d_xyz, rotation | None, scaling | None)."""
    skips = [D // 2]
    inp = torch.cat([posenc(x, x_multires), posenc(t, t_multires), condition], dim=-1)
    h = inp
    for i in range(D):
        h = torch.relu(h @ sd[f"linear.{i}.weight"].T + sd[f"linear.{i}.bias"])
        if i in skips:
            h = torch.cat([inp, h], dim=-1)
    d_xyz = h @ sd["gaussian_warp.weight"].T + sd["gaussian_warp.bias"]
    rot = h @ sd["gaussian_rotation.weight"].T + sd["gaussian_rotation.bias"] if "gaussian_rotation.weight" in sd else None
    scl = h @ sd["gaussian_scaling.weight"].T + sd["gaussian_scaling.bias"] if "gaussian_scaling.weight" in sd else None
    return d_xyz, rot, scl


def get_deformation(sd, means: Tensor, point_ids: Tensor, instances_size: Tensor, instances_embedding: Tensor, t: float,
                    D: int = 8, x_multires: int = 10, t_multires: int = 10):
    """``DeformableNodes.get_deformation`` (deformable.py:35-47): the point is detached and scaled by the box height."""
    ids = point_ids.reshape(-1)
    emb = instances_embedding[ids]
    h = instances_size[ids][..., 2]
    x = means.detach() / h[:, None] * 2
    tt = torch.full((means.shape[0], 1), float(t), dtype=means.dtype)
    return conditional_deform_network(sd, x, tt, emb, D, x_multires, t_multires)


def deformed_canonical(sd, means, quats, point_ids, instances_size, instances_embedding, t, D=8, x_multires=10,
                       t_multires=10, stop_optimizing_canonical_xyz=True):
    """deformable.py:54-68: the canonical means / quaternions handed to ``transform_means`` / ``transform_quats``
    (``get_quats`` = ``quats / |quats|``, vanilla.py:142-146)."""
    d_xyz, d_quat, _ = get_deformation(sd, means, point_ids, instances_size, instances_embedding, t, D, x_multires, t_multires)
    m = (means.detach() if stop_optimizing_canonical_xyz else means) + d_xyz
    q = quats / quats.norm(dim=-1, keepdim=True)
    if d_quat is not None:
        q = q + d_quat
    return m, q
