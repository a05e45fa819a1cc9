"""The two render operators of the reference, re-stated over the B200 rasterizer.

``uv_tex_render`` has the signature, argument meaning and return dict of reference
``render/uv_tex_render.py:7-77``; ``render`` those of ``render/render.py:8-94``. A Texture-GS
checkout can also keep its own ``render/*.py`` verbatim and just put this repo on ``sys.path``:
the top-level packages ``diff_gauss_uv_tex`` / ``diff_gauss`` shadow the pip-git dependencies
(INTEGRATION.md).
"""
from __future__ import annotations

import math

import torch

from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer


def _settings(viewpoint_camera, gaussians, bg_color, scaling_modifier, debug):
    return GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5),
        tanfovy=math.tan(viewpoint_camera.FoVy * 0.5),
        bg=bg_color,
        scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=gaussians.active_sh_degree if hasattr(gaussians, "active_sh_degree") else 0,
        campos=viewpoint_camera.camera_center,
        prefiltered=False,
        debug=debug,
    )


def _screenspace_points(gaussians):
    xyz = gaussians.get_xyz
    pts = torch.zeros_like(xyz, dtype=xyz.dtype, requires_grad=True, device=xyz.device) + 0
    try:
        pts.retain_grad()
    except Exception:
        pass
    return pts


def uv_tex_render(viewpoint_camera, gaussians, cfg=None, bg_color=None, scaling_modifier=1.0, extra_attrs=None,
                  debug=False):
    """Textured render (reference render/uv_tex_render.py:7). Background tensor must be on the GPU."""
    screenspace_points = _screenspace_points(gaussians)
    rasterizer = GaussianRasterizer(raster_settings=_settings(viewpoint_camera, gaussians, bg_color, scaling_modifier, debug))
    image, depth, norm, alpha, radii, extra = rasterizer(
        means3D=gaussians.get_xyz, means2D=screenspace_points, shs=gaussians.get_shs,
        opacities=gaussians.get_opacity, scales=gaussians.get_scaling, rotations=gaussians.get_rotation,
        uvs=gaussians.get_uvs, gradient_uvs=gaussians.get_grad_uvs, texture=gaussians.get_texture,
        extra_attrs=extra_attrs)
    return {"render": image, "depth": depth, "norm": norm, "alpha": alpha,
            "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "extra": extra, "radii": radii}


def uv_tex_render_dual(viewpoint_camera, gaussians, cfg=None, bg_color=None, scaling_modifier=1.0, extra_attrs=None,
                       debug=False):
    """One pass for the reference's two renders per view: the normal textured render
    plus ``"render_no_sh"``, the image the reference obtains by setting ``active_sh_degree = 0`` and
    rendering again (models/texture_gaussian3d.py:375-389 in training, :505-511 in visual_step).
    Geometry, binning, sort, intersections and texel fetches are shared; depth / norm / alpha are the
    same for both and returned once."""
    screenspace_points = _screenspace_points(gaussians)
    rasterizer = GaussianRasterizer(raster_settings=_settings(viewpoint_camera, gaussians, bg_color, scaling_modifier, debug))
    image, depth, norm, alpha, radii, extra, image_no_sh = rasterizer(
        means3D=gaussians.get_xyz, means2D=screenspace_points, shs=gaussians.get_shs,
        opacities=gaussians.get_opacity, scales=gaussians.get_scaling, rotations=gaussians.get_rotation,
        uvs=gaussians.get_uvs, gradient_uvs=gaussians.get_grad_uvs, texture=gaussians.get_texture,
        extra_attrs=extra_attrs, dual_no_sh=True)
    return {"render": image, "render_no_sh": image_no_sh, "depth": depth, "norm": norm, "alpha": alpha,
            "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "extra": extra, "radii": radii}


def render(viewpoint_camera, gaussians, cfg=None, bg_color=None, scaling_modifier=1.0, override_color=None,
           extra_attrs=None, debug=False):
    """Plain 3DGS render (reference render/render.py:8): colour from full SH (``get_features``) or
    ``override_color``."""
    screenspace_points = _screenspace_points(gaussians)
    rasterizer = GaussianRasterizer(raster_settings=_settings(viewpoint_camera, gaussians, bg_color, scaling_modifier, debug))
    shs = gaussians.get_features if override_color is None else None
    image, depth, norm, alpha, radii, extra = rasterizer(
        means3D=gaussians.get_xyz, means2D=screenspace_points, shs=shs, colors_precomp=override_color,
        opacities=gaussians.get_opacity, scales=gaussians.get_scaling, rotations=gaussians.get_rotation,
        cov3Ds_precomp=None, extra_attrs=extra_attrs)
    return {"render": image, "depth": depth, "norm": norm, "alpha": alpha,
            "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "extra": extra, "radii": radii}


type2render_func = dict(render=render, uv_tex_render=uv_tex_render, uv_tex_render_dual=uv_tex_render_dual)   # reference render/__init__.py:4-7
