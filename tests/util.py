"""Shared helpers for the test-suite (tests may import oracle/; the product package may not)."""
from __future__ import annotations

import math

import numpy as np
import torch

from oracle.raster_ref import RasterSettings, rasterize as oracle_rasterize
from texture_gs_b200.scene import SyntheticGaussians, orbit_cameras, output_cotangents, sphere_shell_scene

ABS_TOL = 1e-4      # BASELINE.json north_star: 1e-4 abs fp32 per pixel
GRAD_RTOL = 1e-3    # BASELINE.json north_star: grads within 1e-3 rel


def oracle_settings(cam, sh_degree, dtype=torch.float32, bg=(0.0, 0.0, 0.0), device="cpu", scale_modifier=1.0):
    return RasterSettings(
        image_height=cam.image_height, image_width=cam.image_width,
        tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5),
        bg=torch.tensor(bg, dtype=dtype, device=device), scale_modifier=scale_modifier,
        viewmatrix=cam.world_view_transform.to(device=device, dtype=dtype),
        projmatrix=cam.full_proj_transform.to(device=device, dtype=dtype),
        sh_degree=sh_degree, campos=cam.camera_center.to(device=device, dtype=dtype))


def run_oracle(g: SyntheticGaussians, cam, bg=(0.0, 0.0, 0.0), cot=None, dtype=torch.float32, sw=None,
               scale_modifier=1.0, **kw):
    """Forward (+ backward if ``cot`` given) of the oracle on CPU. Returns (outputs, aux, grads)."""
    from oracle.raster_ref import Switches
    gg = g.to(device="cpu", dtype=dtype, requires_grad=cot is not None)
    t = gg.tensors()
    st = oracle_settings(cam, g.active_sh_degree, dtype=dtype, bg=bg, scale_modifier=scale_modifier)
    m2 = torch.zeros_like(t["xyz"], requires_grad=cot is not None)
    out = oracle_rasterize(t["xyz"], m2, t["shs"], t["opacity"], t["scaling"], t["rotation"], t["uvs"],
                           t["grad_uvs"], t["texture"], st, sw or Switches(), return_aux=True, **kw)
    image, depth, norm, alpha, radii, extra, aux = out
    grads = None
    if cot is not None:
        c = [x.to(dtype) for x in cot]
        L = (image * c[0]).sum() + (depth * c[1]).sum() + (norm * c[2]).sum() + (alpha * c[3]).sum()
        L.backward()
        grads = {k: (v.grad if v is not None and v.grad is not None else None) for k, v in t.items()}
        grads["means2D"] = m2.grad
    return (image.detach(), depth.detach(), norm.detach(), alpha.detach(), radii), aux, grads


def run_cuda(g: SyntheticGaussians, cam, bg=(0.0, 0.0, 0.0), cot=None, debug=False, device="cuda", scale_modifier=1.0):
    """Forward (+ backward) through the product operator ``uv_tex_render`` (-> C-ABI)."""
    from texture_gs_b200 import uv_tex_render, last_stats
    gg = g.to(device=device, dtype=torch.float32, requires_grad=cot is not None)
    cam_d = cam.to(device)
    bg_t = torch.tensor(bg, dtype=torch.float32, device=device)
    pkg = uv_tex_render(cam_d, gg, None, bg_t, scaling_modifier=scale_modifier, debug=debug)
    grads = None
    if cot is not None:
        c = [x.to(device) for x in cot]
        L = (pkg["render"] * c[0]).sum() + (pkg["depth"] * c[1]).sum() + (pkg["norm"] * c[2]).sum() + (pkg["alpha"] * c[3]).sum()
        L.backward()
        t = gg.tensors()
        grads = {k: (v.grad.detach().cpu() if v is not None and v.grad is not None else None) for k, v in t.items()}
        grads["means2D"] = pkg["viewspace_points"].grad.detach().cpu()
    outs = tuple(pkg[k].detach().cpu() for k in ("render", "depth", "norm", "alpha", "radii"))
    return outs, last_stats(), grads


def compare_images(cuda_outs, ref_outs, ambiguous, names=("image", "depth", "norm", "alpha"), tol=ABS_TOL):
    """Per-pixel comparison. Returns a report dict; pixels flagged ambiguous by the oracle (a blend
    decision within a few ulp of its threshold) are reported separately."""
    rep = {}
    amb = ambiguous.bool()
    for n, a, b in zip(names, cuda_outs, ref_outs):
        d = (a.double() - b.double()).abs().amax(dim=0)
        rep[n] = dict(max_all=float(d.max()), max_clear=float(d[~amb].max()) if (~amb).any() else 0.0,
                      frac_over=float((d > tol).double().mean()), frac_over_clear=float((d[~amb] > tol).double().mean()) if (~amb).any() else 0.0)
    rep["ambiguous_frac"] = float(amb.double().mean())
    return rep


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max-norm relative error of two gradient tensors."""
    a, b = a.double(), b.double()
    den = float(b.abs().max())
    if den == 0.0:
        return float(a.abs().max())
    return float((a - b).abs().max()) / den
