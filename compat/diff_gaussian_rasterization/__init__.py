"""Import-compatible stand-in for the reference's `diff_gaussian_rasterization` package: put `compat/` on
PYTHONPATH ahead of the original and the reference's gaussian_renderer/__init__.py:14 import resolves to
libdgs_b200."""
from deblurgs_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                      rasterize_gaussians, _RasterizeGaussians)
