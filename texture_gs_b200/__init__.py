"""texture_gs_b200 — B200-native (sm_100a) differentiable rasterizer for the Texture-GS hot path.

Public surface (mirrors what reference render/uv_tex_render.py and render/render.py use):
    GaussianRasterizationSettings, GaussianRasterizer, uv_tex_render, render
"""
from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, invalidate_packed_cache, invalidate_settings_cache,  # noqa: F401
                         last_stats, SpecSwitches, spec_switches)
from .render import render, uv_tex_render, uv_tex_render_dual, type2render_func  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "uv_tex_render", "uv_tex_render_dual", "render", "type2render_func",
           "last_stats", "SpecSwitches", "spec_switches"]
