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


def _cfg_flag(cfg, name):
    if cfg is None:
        return False
    return bool(cfg.get(name, False)) if isinstance(cfg, dict) else bool(getattr(cfg, name, False))


def eval_sh(deg: int, sh: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """Real SH up to degree 3 with the reference's constants and sign pattern (utils/sh.py:26-112): ``sh`` is
    (..., C, (deg+1)^2), ``dirs`` (..., 3) unit vectors -> (..., C). Used only for ``cfg.convert_SHs_python``."""
    C0, C1 = 0.28209479177387814, 0.4886025119029199
    C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
    C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435)
    res = C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        res = res - C1 * y * sh[..., 1] + C1 * z * sh[..., 2] - C1 * x * sh[..., 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            res = (res + C2[0] * xy * sh[..., 4] + C2[1] * yz * sh[..., 5] + C2[2] * (2.0 * zz - xx - yy) * sh[..., 6]
                   + C2[3] * xz * sh[..., 7] + C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                res = (res + C3[0] * y * (3 * xx - yy) * sh[..., 9] + C3[1] * xy * z * sh[..., 10]
                       + C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12]
                       + C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + C3[5] * z * (xx - yy) * sh[..., 14]
                       + C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return res


def render(viewpoint_camera, gaussians, cfg=None, bg_color=None, scaling_modifier=1.0, override_color=None,
           extra_attrs=None, debug=False):
    """Plain 3DGS render (reference render/render.py:8-94): colour from full SH (``get_features``), from SH evaluated
    in Python (``cfg.convert_SHs_python``, :63-68) or from ``override_color``; covariance from scales + rotations or
    precomputed in Python (``cfg.compute_cov3D_python``, :52-53)."""
    screenspace_points = _screenspace_points(gaussians)
    rasterizer = GaussianRasterizer(raster_settings=_settings(viewpoint_camera, gaussians, bg_color, scaling_modifier, debug))
    scales = rotations = cov3D_precomp = None
    if _cfg_flag(cfg, "compute_cov3D_python"):
        cov3D_precomp = gaussians.get_covariance(scaling_modifier)
    else:
        scales, rotations = gaussians.get_scaling, gaussians.get_rotation
    shs = colors_precomp = None
    if override_color is None:
        if _cfg_flag(cfg, "convert_SHs_python"):
            feats = gaussians.get_features
            shs_view = feats.transpose(1, 2).view(-1, 3, (gaussians.max_sh_degree + 1) ** 2)
            dir_pp = gaussians.get_xyz - viewpoint_camera.camera_center.repeat(feats.shape[0], 1)
            sh2rgb = eval_sh(gaussians.active_sh_degree, shs_view, dir_pp / dir_pp.norm(dim=1, keepdim=True))
            colors_precomp = torch.clamp_min(sh2rgb + 0.5, 0.0)
        else:
            shs = gaussians.get_features
    else:
        colors_precomp = override_color
    image, depth, norm, alpha, radii, extra = rasterizer(
        means3D=gaussians.get_xyz, means2D=screenspace_points, shs=shs, colors_precomp=colors_precomp,
        opacities=gaussians.get_opacity, scales=scales, rotations=rotations,
        cov3Ds_precomp=cov3D_precomp, extra_attrs=extra_attrs)
    return {"render": image, "depth": depth, "norm": norm, "alpha": alpha,
            "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "extra": extra, "radii": radii}


type2render_func = dict(render=render, uv_tex_render=uv_tex_render, uv_tex_render_dual=uv_tex_render_dual)   # reference render/__init__.py:4-7
