"""``gsplat.rendering`` -- the symbol the reference imports at OmniRe/models/gaussians/basics.py:12."""
from emd_b200.gsplat_api import rasterization  # noqa: F401
