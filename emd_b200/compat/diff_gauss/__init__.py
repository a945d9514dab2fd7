"""``diff_gauss`` import-name shim backed by emd_b200 (S3Gaussian/gaussian_renderer/__init__.py:14)."""
__emd_b200__ = True
from emd_b200.diff_gauss_api import GaussianRasterizationSettings, GaussianRasterizer  # noqa: F401
