"""emd_b200: B200-native (sm_100a) implementation of EMD's dense render hot path.

EMD motion-embedding deformation -> projection -> tile intersection + depth-keyed
radix sort + tile ranges -> alpha-composited rasterization (forward and
backward), as hand-written CUDA kernels behind a C ABI (``include/emd_b200.h``),
with the reference's two rasterizer call surfaces kept:
``emd_b200.gsplat_api.rasterization`` (``gsplat.rendering.rasterization``) and
``emd_b200.diff_gauss_api`` (``diff_gauss.GaussianRasterizer``).
``emd_b200.compat.install()`` registers them under those import names so the
reference's OmniRe / S3Gaussian code runs unchanged.

There is no CPU fallback: every op raises if ``libemd_b200.so`` is missing or a
tensor is not on a CUDA device.
"""
from . import _C  # noqa: F401
from .gsplat_api import rasterization  # noqa: F401
from .hexplane import HexPlaneField  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .sh_ops import activate_gaussians, spherical_harmonics  # noqa: F401

__all__ = ["rasterization", "spherical_harmonics", "activate_gaussians", "HexPlaneField", "FusedAdam"]
