"""Stand-in for the reference's `gaussian_renderer` package (gaussian_renderer/__init__.py:18-90)."""
from deblurgs_b200.renderer import render, render_blurry  # noqa: F401
