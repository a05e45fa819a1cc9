"""TEST INFRASTRUCTURE: the ``diff_gauss_uv_tex`` / ``diff_gauss`` module surface on top of the emulated library, so that
CPU-only tests can drive the real kernel source through the SAME Python call the reference makes
(``GaussianRasterizer(raster_settings=...)(means3D=..., ...)`` -> 6-tuple, gradients through autograd)."""
from __future__ import annotations

import torch
from torch import nn

from texture_gs_b200.rasterizer import GaussianRasterizationSettings  # noqa: F401  (the product's own NamedTuple)

from . import emu

_NAMES = ("means3D", "means2D", "shs", "colors_precomp", "opacities", "scales", "rotations", "uvs", "gradient_uvs", "texture",
          "extra_attrs", "cov3Ds_precomp")


class _EmuRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, st, dual, *tensors):
        kw = dict(zip(_NAMES, tensors))
        ctx.st, ctx.dual, ctx.kw = st, dual, {k: (None if v is None else v.detach()) for k, v in kw.items()}
        res = _EmuRasterize._run(st, dual, ctx.kw, None)
        P = kw["means3D"].shape[0]
        ctx.mark_non_differentiable(res.radii)
        extra = res.extra if res.extra is not None else torch.zeros(0)
        nosh = res.image_nosh if res.image_nosh is not None else torch.zeros(0)
        return res.image, res.depth, res.norm, res.alpha, res.radii, extra, nosh

    @staticmethod
    def _run(st, dual, kw, cots):
        args = {k: v for k, v in kw.items() if k != "means2D"}
        extra_kw = {}
        if cots is not None:
            extra_kw = dict(cotangents=cots[:4], cot_extra=cots[4], cot_nosh=cots[5])
        return emu.rasterize(H=int(st.image_height), W=int(st.image_width), tanfovx=st.tanfovx, tanfovy=st.tanfovy,
                             bg=[float(x) for x in st.bg.reshape(-1).tolist()], scale_modifier=st.scale_modifier,
                             viewmatrix=st.viewmatrix.detach().cpu(), projmatrix=st.projmatrix.detach().cpu(),
                             campos=st.campos.detach().cpu(), sh_degree=int(st.sh_degree), debug=bool(st.debug), dual_no_sh=dual,
                             **args, **extra_kw)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_image, g_depth, g_norm, g_alpha, _g_radii, g_extra, g_nosh):
        H, W = int(ctx.st.image_height), int(ctx.st.image_width)
        z = lambda g, c: torch.zeros(c, H, W) if g is None else g
        cots = (z(g_image, 3), z(g_depth, 1), z(g_norm, 3), z(g_alpha, 1),
                g_extra if (ctx.kw["extra_attrs"] is not None) else None, g_nosh if ctx.dual else None)
        res = _EmuRasterize._run(ctx.st, ctx.dual, ctx.kw, cots)
        out = []
        for k in _NAMES:
            g = res.grads.get(k) if ctx.kw[k] is not None or k == "means2D" else None
            out.append(None if g is None else g.reshape(ctx.kw[k].shape) if ctx.kw[k] is not None else g)
        out[_NAMES.index("gradient_uvs")] = None
        return (None, None, *out)


class GaussianRasterizer(nn.Module):
    """Same constructor and call signature as ``texture_gs_b200.rasterizer.GaussianRasterizer``."""

    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3Ds_precomp=None, uvs=None, gradient_uvs=None, texture=None, extra_attrs=None, dual_no_sh=False):
        image, depth, norm, alpha, radii, extra, nosh = _EmuRasterize.apply(
            self.raster_settings, bool(dual_no_sh), means3D, means2D, shs, colors_precomp, opacities, scales, rotations, uvs, gradient_uvs,
            texture, extra_attrs, cov3Ds_precomp)
        extra = extra if extra_attrs is not None else None
        if dual_no_sh:
            return image, depth, norm, alpha, radii, extra, nosh
        return image, depth, norm, alpha, radii, extra
