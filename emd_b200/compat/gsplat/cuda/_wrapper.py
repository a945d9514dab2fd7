"""``gsplat.cuda._wrapper`` -- the symbols the reference imports at OmniRe/models/gaussians/basics.py:16."""
from emd_b200.raster_ops import (  # noqa: F401
    fully_fused_projection,
    isect_offset_encode,
    isect_tiles,
    rasterize_to_pixels,
)
from emd_b200.sh_ops import spherical_harmonics  # noqa: F401
