"""CPU oracle: quaternion algebra of the reference (w-first).  TEST INFRASTRUCTURE.

Follows ``OmniRe/models/gaussians/basics.py:30-49`` (quat_to_rotmat),
``:100-110`` (quat_mult), ``:53-81`` (interpolate_quats) and
``S3Gaussian/utils/graphics_utils.py:172-195`` (batch_quaternion_multiply,
which normalises its result).  Pinned by ``tests/golden/quat_*.npz``.
"""
import torch
import torch.nn.functional as F
from torch import Tensor


def quat_act(x: Tensor) -> Tensor:
    """``x / x.norm(dim=-1, keepdim=True)`` (vanilla.py:143-144)."""
    return x / x.norm(dim=-1, keepdim=True)


def quat_to_rotmat(quats: Tensor) -> Tensor:
    quats = F.normalize(quats, p=2, dim=-1)
    w, x, y, z = torch.unbind(quats, dim=-1)
    R = torch.stack(
        [
            1 - 2 * (y**2 + z**2), 2 * (x * y - w * z), 2 * (x * z + w * y),
            2 * (x * y + w * z), 1 - 2 * (x**2 + z**2), 2 * (y * z - w * x),
            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x**2 + y**2),
        ],
        dim=-1,
    )
    return R.reshape(quats.shape[:-1] + (3, 3))


def quat_mult(q1: Tensor, q2: Tensor) -> Tensor:
    """Hamilton product q1 (x) q2, any leading shape."""
    w1, x1, y1, z1 = q1.unbind(-1)
    w2, x2, y2, z2 = q2.unbind(-1)
    w = w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2
    x = w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2
    y = w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2
    z = w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2
    return torch.stack([w, x, y, z], dim=-1)


def batch_quaternion_multiply(q1: Tensor, q2: Tensor) -> Tensor:
    q = quat_mult(q1, q2)
    return q / q.norm(dim=-1, keepdim=True)


def interpolate_quats(q1: Tensor, q2: Tensor, fraction: float = 0.5) -> Tensor:
    """Slerp with the reference's near-parallel fallback (basics.py:53-81)."""
    q1 = q1 / torch.norm(q1, dim=-1, keepdim=True)
    q2 = q2 / torch.norm(q2, dim=-1, keepdim=True)
    dot = (q1 * q2).sum(dim=-1).clamp(-1, 1)
    neg = dot < 0
    q2 = torch.where(neg[..., None], -q2, q2)
    dot = torch.where(neg, -dot, dot)
    similar = dot > 0.9995
    lin = q1 + fraction * (q2 - q1)
    theta_0 = torch.acos(dot)
    theta = theta_0 * fraction
    s2 = torch.sin(theta) / torch.sin(theta_0)
    s1 = torch.cos(theta) - dot * s2
    sl = s1[..., None] * q1 + s2[..., None] * q2
    return torch.where(similar[..., None], lin, sl)


def matrix_to_quaternion(matrix: Tensor) -> Tensor:
    """pytorch3d.transforms.matrix_to_quaternion (used at smpl.py:522): pick the
    best-conditioned of the four candidate quaternions; w-first; no sign
    standardisation is relied on by the reference (the result is normalised
    and multiplied, and q ~ -q render identically)."""
    m00, m01, m02 = matrix[..., 0, 0], matrix[..., 0, 1], matrix[..., 0, 2]
    m10, m11, m12 = matrix[..., 1, 0], matrix[..., 1, 1], matrix[..., 1, 2]
    m20, m21, m22 = matrix[..., 2, 0], matrix[..., 2, 1], matrix[..., 2, 2]

    def sqrt_pos(x):
        return torch.where(x > 0, torch.sqrt(torch.clamp(x, min=1e-30)), torch.zeros_like(x))

    q_abs = torch.stack(
        [
            sqrt_pos(1.0 + m00 + m11 + m22),
            sqrt_pos(1.0 + m00 - m11 - m22),
            sqrt_pos(1.0 - m00 + m11 - m22),
            sqrt_pos(1.0 - m00 - m11 + m22),
        ],
        dim=-1,
    )
    quat_by_rijk = torch.stack(
        [
            torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
            torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
            torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
            torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1),
        ],
        dim=-2,
    )
    flr = torch.tensor(0.1, dtype=q_abs.dtype)
    cand = quat_by_rijk / (2.0 * q_abs[..., None].max(flr))
    best = q_abs.argmax(dim=-1)
    out = torch.gather(cand, -2, best[..., None, None].expand(best.shape + (1, 4))).squeeze(-2)
    # pytorch3d >= 0.7.3 standardises to w >= 0
    return torch.where(out[..., 0:1] < 0, -out, out)
