"""Drop-in for the pip-git dependency ``diff_gauss`` (reference requirements.txt:14), the module
reference render/render.py:4 imports: the same rasterizer in its texture-less mode."""
from texture_gs_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer"]
