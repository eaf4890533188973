"""Stand-in for the reference's `simple_knn._C` (scene/gaussian_model.py:20)."""
from deblurgs_b200.knn import distCUDA2  # noqa: F401
