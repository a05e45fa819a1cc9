"""Drop-in for the pip-git dependency ``diff_gauss_uv_tex`` (reference requirements.txt:15), the
module reference render/uv_tex_render.py:4 imports. Backed by texture_gs_b200 (sm_100a kernels)."""
from texture_gs_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer"]
