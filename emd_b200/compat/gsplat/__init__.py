"""``gsplat`` import-name shim backed by emd_b200 (see emd_b200/compat/__init__.py)."""
__emd_b200__ = True
__version__ = "1.3.0+emd_b200"
from emd_b200.gsplat_api import rasterization  # noqa: E402,F401
from emd_b200.sh_ops import spherical_harmonics  # noqa: E402,F401
from . import rendering  # noqa: E402,F401
