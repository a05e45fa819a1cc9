"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the Texture-GS rasterizer.

    *** PARITY UNPINNED ***
    The reference's arithmetic for this path lives in the un-vendored, unpinned pip-git dependency
    ``diff_gauss_uv_tex`` (reference ``requirements.txt:15``; imported at
    ``render/uv_tex_render.py:4``) and its sibling ``diff_gauss`` (``requirements.txt:14``,
    ``render/render.py:4``). Neither source is under /root/reference, neither is installed, there
    is no network, and the reference holds no tests / golden vectors for the path.
    This file therefore restates (a) everything the reference *does* pin in-tree — operator
    surface, camera/pixel/quaternion/SH/cube-map conventions, output semantics — each cited below,
    and (b) the published 3DGS tile-rasterizer algorithm (graphdeco-inria/diff-gaussian-rasterization,
    Kerbl et al. 2023) plus the Texture-GS paper's ray–Gaussian intersection + first-order UV
    expansion (arXiv 2403.10050 §3), as the named spec choices E1–E13 of SURVEY.md §8c.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product package never does.

Pure torch, no custom ops, dtype-parametric (float64 for gradcheck; float32 for kernel parity),
differentiable by autograd. The per-tile lists are processed vectorised (padded lists), the heavy
per-contribution part (intersection, UV, cube lookup) sparsely.

Pinned in-tree conventions followed here (paths relative to /root/reference):
  * row-vector camera maths, ``p_view=[p,1]@world_view_transform``, ``p_clip=[p,1]@full_proj_transform``
    — utils/cameras.py:62-65, utils/graphics.py:38-71
  * pixel centres ``ndc=(2*pix+1)/S-1``  — losses/norm_reg_loss.py:25-30, models/texture_gaussian3d.py:303-304
  * view ray of a pixel ``(ndc_x*tanfovx, ndc_y*tanfovy, 1)`` — losses/norm_reg_loss.py:30
  * quaternion (r,x,y,z) -> R — utils/general.py:87-108 ; Sigma3D=(R S)(R S)^T — utils/general.py:110-119
  * SH basis / constants / sign pattern — utils/sh.py:26-112 ; colour = SH + 0.5 clamped at 0 — render/render.py:67-68
  * texture value -> rgb  C0*t+0.5 — models/texture_gaussian3d.py:16-21
  * Jacobian layout  J[3i+j]=d uv_i / d x_j — models/texture_gaussian3d.py:223-227
  * cube-map faces / texel centres / index order [face,row(y),col(x),rgb] —
    models/modules/NVDIFFREC/util.py:94-116, models/modules/NVDIFFREC/renderutils/c_src/cubemap.cu:32-61
  * output semantics (un-normalised depth / world normal / alpha, background only on colour) —
    models/texture_gaussian3d.py:299-309,322-368
"""
from __future__ import annotations

import math
from typing import NamedTuple, Optional

import numpy as np
import torch

TILE = 16
C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
      -1.0925484305920792, 0.5462742152960396)
C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
      -0.4570457994644658, 1.445305721320277, -0.5900435899266435)

ALPHA_MIN = 1.0 / 255.0
ALPHA_MAX = 0.99
T_STOP = 1e-4
NEAR_CULL = 0.2
ND_EPS = 1e-8
FACE_TIE_REL = 1e-4   # test-side conditioning flag: |u'| major/second-major tie (cube face chosen by rounding)
TEXEL_TIE = 2e-6      # test-side conditioning flag: a texel-space coordinate within TEXEL_TIE * R of an integer — the
                      # bilinear VALUE is continuous there, its derivative w.r.t. u' is not (the cell is chosen by rounding;
                      # fp32 resolves f = (s+1)R/2 - 1/2 to about 5e-7 * R)
DEPTH_TIE_REL = 1e-6  # test-side conditioning flag: two splats that both blend into a pixel, adjacent in its list, with view-space
                      # depths within ~8 ulp: their ORDER (spec E4) is decided by the rounding of z = p.V[:,2]
CLAMP_TIE = 4e-6      # test-side conditioning flag (gradients only): a colour channel within rounding of the clamp max(0, .)
THRESH_ULPS = 256     # test-side conditioning flag: alpha within THRESH_ULPS * 4e-6 (relative, fp32) of 1/255, T of 1e-4. 64 was
                      # enough for the small test scenes; at 500 k splats the fp32 conic of thin splats (det = ac - b^2 cancels)
                      # moves alpha by up to ~1e-3 relative (measured: float32 vs float64 build of oracle/raster_c.c)
GRAZING_COS = 0.05    # test-side conditioning flag only (never changes the rendered values); calibrated so that
                      # the fp32 and fp64 oracles agree to 1e-3 on every gradient once flagged pixels carry no cotangent


class RasterSettings(NamedTuple):
    """Same 12 fields as the reference builds at render/uv_tex_render.py:25-38."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool = False
    debug: bool = False


class Switches(NamedTuple):
    """Named spec choices that the reference does not pin (SURVEY §8c [EXT])."""
    depth_of_intersection: bool = False   # E7: False -> z of the Gaussian centre
    seamless_cube: bool = False           # E11: False -> clamp-to-edge inside the face; True -> taps beyond a face edge come from the adjacent face
    stopgrad_delta: bool = False          # E13: False -> full derivative through the intersection
    normalize_quat: bool = False          # upstream CUDA uses the quaternion as given
    upstream_clamp_grad: bool = False     # E2-alt: a view-space x/z (y/z) clamped at 1.3 tan(fov/2) passes NO gradient (the 3DGS lineage
                                          # zeroes it: `x_grad_mul`); False -> exact derivative of t.x = clamp(x/z) * z, i.e. lim * dz


# ---------------------------------------------------------------------------------------------
# per-Gaussian stage (E1-E3, E8, E12's SH part)
# ---------------------------------------------------------------------------------------------

def quat_to_rot(q: torch.Tensor) -> torch.Tensor:
    """reference utils/general.py:87-108 (without the normalisation, see Switches)."""
    r, x, y, z = q.unbind(-1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=-1)
    return R.reshape(q.shape[:-1] + (3, 3))


def sh_rest(deg: int, shs: Optional[torch.Tensor], dirs: torch.Tensor) -> torch.Tensor:
    """Bands l>=1 of reference utils/sh.py:57-112 applied to ``shs`` (N,M,3) holding the *rest*
    coefficients only (the DC term lives in the texture; models/texture_gaussian3d.py:97-98)."""
    out = torch.zeros(dirs.shape[0], 3, dtype=dirs.dtype, device=dirs.device)
    if shs is None or deg <= 0:
        return out
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    out = out - C1 * y * shs[:, 0] + C1 * z * shs[:, 1] - C1 * x * shs[:, 2]
    if deg > 1:
        xx, yy, zz = x * x, y * y, z * z
        xy, yz, xz = x * y, y * z, x * z
        out = (out + C2[0] * xy * shs[:, 3] + C2[1] * yz * shs[:, 4]
               + C2[2] * (2.0 * zz - xx - yy) * shs[:, 5] + C2[3] * xz * shs[:, 6]
               + C2[4] * (xx - yy) * shs[:, 7])
        if deg > 2:
            out = (out + C3[0] * y * (3 * xx - yy) * shs[:, 8] + C3[1] * xy * z * shs[:, 9]
                   + C3[2] * y * (4 * zz - xx - yy) * shs[:, 10]
                   + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * shs[:, 11]
                   + C3[4] * x * (4 * zz - xx - yy) * shs[:, 12]
                   + C3[5] * z * (xx - yy) * shs[:, 13] + C3[6] * x * (xx - 3 * yy) * shs[:, 14])
    return out


def preprocess(means3D, means2D, scales, rotations, opacities, shs, st: RasterSettings,
               sw: Switches = Switches(), cov3Ds_precomp=None):
    """Per-Gaussian projection. Returns a dict of per-Gaussian tensors (differentiable where it
    matters) plus integer radius / tile rect (E1-E3)."""
    dt, dev = means3D.dtype, means3D.device
    H, W = int(st.image_height), int(st.image_width)
    V = st.viewmatrix.to(dt)
    PM = st.projmatrix.to(dt)
    campos = st.campos.to(dt)
    N = means3D.shape[0]

    p_view = means3D @ V[:3, :3] + V[3, :3]
    p_hom = means3D @ PM[:3, :] + PM[3, :]
    p_w = 1.0 / (p_hom[:, 3] + 1e-7)
    p_proj = p_hom[:, :3] * p_w[:, None]
    depth = p_view[:, 2]
    in_front = depth > NEAR_CULL                                  # E1

    if cov3Ds_precomp is None:
        q = rotations
        if sw.normalize_quat:
            q = q / q.norm(dim=-1, keepdim=True)
        R = quat_to_rot(q)
        L = R * (scales * st.scale_modifier)[:, None, :]              # R @ diag(s*mod)
        Sigma = L @ L.transpose(1, 2)
    else:
        # render/render.py:52-53: gaussians.get_covariance(scaling_modifier) -> (N,6) in the order
        # xx,xy,xz,yy,yz,zz (utils/general.py:73-82), used as given
        c6 = cov3Ds_precomp
        Sigma = torch.stack([c6[:, 0], c6[:, 1], c6[:, 2], c6[:, 1], c6[:, 3], c6[:, 4],
                             c6[:, 2], c6[:, 4], c6[:, 5]], dim=-1).reshape(N, 3, 3)

    # E2: EWA projection
    limx, limy = 1.3 * st.tanfovx, 1.3 * st.tanfovy
    tz = p_view[:, 2]
    tz_safe = torch.where(in_front, tz, torch.ones_like(tz))
    tx = torch.clamp(p_view[:, 0] / tz_safe, -limx, limx) * tz_safe
    ty = torch.clamp(p_view[:, 1] / tz_safe, -limy, limy) * tz_safe
    if sw.upstream_clamp_grad:
        qx, qy = (p_view[:, 0] / tz_safe).detach(), (p_view[:, 1] / tz_safe).detach()
        tx = torch.where((qx < -limx) | (qx > limx), tx.detach(), tx)
        ty = torch.where((qy < -limy) | (qy > limy), ty.detach(), ty)
    fx = W / (2.0 * st.tanfovx)
    fy = H / (2.0 * st.tanfovy)
    zero = torch.zeros_like(tz)
    Jm = torch.stack([fx / tz_safe, zero, -fx * tx / (tz_safe * tz_safe),
                      zero, fy / tz_safe, -fy * ty / (tz_safe * tz_safe)], dim=-1).reshape(N, 2, 3)
    Wm = V[:3, :3].t()                                            # world -> view rotation
    T = Jm @ Wm
    cov = T @ Sigma @ T.transpose(1, 2)
    a = cov[:, 0, 0] + 0.3
    b = cov[:, 0, 1]
    c = cov[:, 1, 1] + 0.3
    det = a * c - b * b
    det_ok = det != 0
    det_safe = torch.where(det_ok, det, torch.ones_like(det))
    conic = torch.stack([c / det_safe, -b / det_safe, a / det_safe], dim=-1)

    # E3: radius + tile rect (integers, no gradient)
    with torch.no_grad():
        mid = 0.5 * (a + c)
        lam = mid + torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))
        radius = torch.ceil(3.0 * torch.sqrt(lam))
    ndc_xy = p_proj[:, :2]
    if means2D is not None:                                       # E13: screen-space probe in NDC units
        ndc_xy = ndc_xy + means2D[:, :2]
    xy = torch.stack([((ndc_xy[:, 0] + 1.0) * W - 1.0) * 0.5,
                      ((ndc_xy[:, 1] + 1.0) * H - 1.0) * 0.5], dim=-1)
    with torch.no_grad():
        gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
        xd, yd, rd = xy[:, 0].detach(), xy[:, 1].detach(), radius
        ok_fin = torch.isfinite(xd) & torch.isfinite(yd) & torch.isfinite(rd)
        xd = torch.where(ok_fin, xd, torch.zeros_like(xd))
        yd = torch.where(ok_fin, yd, torch.zeros_like(yd))
        rd = torch.where(ok_fin, rd, torch.zeros_like(rd))
        rx0 = torch.clamp(torch.floor((xd - rd) / TILE), 0, gx).to(torch.int64)
        rx1 = torch.clamp(torch.floor((xd + rd + (TILE - 1)) / TILE), 0, gx).to(torch.int64)
        ry0 = torch.clamp(torch.floor((yd - rd) / TILE), 0, gy).to(torch.int64)
        ry1 = torch.clamp(torch.floor((yd + rd + (TILE - 1)) / TILE), 0, gy).to(torch.int64)
        area = (rx1 - rx0) * (ry1 - ry0)
        visible = in_front & det_ok & ok_fin & (area > 0)
        radii = torch.where(visible, radius, torch.zeros_like(radius)).to(torch.int32)

    # E8: disc normal = rotation column of the smallest scale, facing the camera
    if cov3Ds_precomp is None:
        with torch.no_grad():
            kmin = torch.argmin(scales, dim=1)
        n_raw = torch.gather(R, 2, kmin[:, None, None].expand(N, 3, 1)).squeeze(-1)
    else:
        # E8 with a precomputed covariance [spec choice]: the same direction, obtained as the unit eigenvector of
        # the smallest eigenvalue (R[:, argmin s] IS that eigenvector when Sigma = R S^2 R^T); no gradient
        with torch.no_grad():
            n_raw = torch.linalg.eigh(Sigma.double()).eigenvectors[:, :, 0].to(dt)
    m = means3D - campos
    with torch.no_grad():
        flip = (n_raw * m).sum(-1) > 0
    normal = torch.where(flip[:, None], -n_raw, n_raw)

    # E12 (per-Gaussian part): view-dependent SH-rest colour
    dirs = m / m.norm(dim=1, keepdim=True)
    csh = sh_rest(int(st.sh_degree), shs, dirs)

    return dict(xy=xy, conic=conic, depth=depth, opacity=opacities.reshape(N), normal=normal, csh=csh,
                m=m, radii=radii, visible=visible, rect=(rx0, ry0, rx1, ry1), grid=(gx, gy))


# ---------------------------------------------------------------------------------------------
# binning (E4): (tile, depth, id) ascending
# ---------------------------------------------------------------------------------------------

def build_tile_lists(pre, depth_key_dtype=np.float32):
    """Returns (tile_ids[K], gauss_ids[K]) sorted by (tile, depth, gaussian index)."""
    rx0, ry0, rx1, ry1 = [t.cpu().numpy() for t in pre["rect"]]
    vis = pre["visible"].cpu().numpy()
    gx, gy = pre["grid"]
    ids = np.nonzero(vis)[0]
    w = (rx1 - rx0)[ids]
    h = (ry1 - ry0)[ids]
    cnt = w * h
    K = int(cnt.sum())
    if K == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    rep = np.repeat(np.arange(ids.shape[0]), cnt)
    start = np.cumsum(cnt) - cnt
    local = np.arange(K) - np.repeat(start, cnt)
    wrep = w[rep]
    tx = rx0[ids][rep] + local % wrep
    ty = ry0[ids][rep] + local // wrep
    tile = ty * gx + tx
    gid = ids[rep]
    depth = pre["depth"].detach().cpu().numpy().astype(depth_key_dtype)[gid]
    order = np.lexsort((gid, depth, tile))
    return tile[order].astype(np.int64), gid[order].astype(np.int64)


# ---------------------------------------------------------------------------------------------
# cube map lookup (E11)
# ---------------------------------------------------------------------------------------------

def cube_face_coords(u: torch.Tensor):
    """Direction (...,3) -> (face, sx, sy) with (sx,sy) in [-1,1]; inverse of reference
    NVDIFFREC/util.py:94-101 ``cube_to_dir`` (C twin ``dir_to_side`` cubemap.cu:49-61).
    Major axis ties resolve x > y > z."""
    x, y, z = u.unbind(-1)
    ax, ay, az = x.abs(), y.abs(), z.abs()
    is_x = (ax >= ay) & (ax >= az)
    is_y = (~is_x) & (ay >= az)
    m = torch.where(is_x, ax, torch.where(is_y, ay, az))
    inv = 1.0 / torch.clamp_min(m, 1e-20)
    face = torch.where(is_x, torch.where(x < 0, 1, 0),
                       torch.where(is_y, torch.where(y < 0, 3, 2), torch.where(z < 0, 5, 4)))
    sx_num = torch.where(is_x, torch.where(x < 0, z, -z),
                         torch.where(is_y, x, torch.where(z < 0, -x, x)))
    sy_num = torch.where(is_x, -y, torch.where(is_y, torch.where(y < 0, -z, z), -y))
    return face, sx_num * inv, sy_num * inv


# E11-alt (seamless): which texel lies one step beyond an edge of a face. CUBE_WRAP[f][side] = (f', xcode, ycode), side
# 0: x < 0, 1: x >= R, 2: y < 0, 3: y >= R; codes 0 -> 0, 1 -> R-1, 2 -> k, 3 -> R-1-k with k the position along the edge.
# Derived from the reference's face table (NVDIFFREC/util.py:94-101) by tools/gen_cube_wrap.py; re-derived in the tests.
CUBE_WRAP = [[(4, 1, 2), (5, 0, 2), (2, 1, 3), (3, 1, 2)], [(5, 1, 2), (4, 0, 2), (2, 0, 2), (3, 0, 3)],
             [(1, 2, 0), (0, 3, 0), (5, 3, 0), (4, 2, 0)], [(1, 3, 1), (0, 2, 1), (4, 2, 1), (5, 3, 1)],
             [(1, 1, 2), (0, 0, 2), (2, 2, 1), (3, 2, 0)], [(0, 1, 2), (1, 0, 2), (2, 3, 0), (3, 3, 1)]]


def cube_wrap_tap(face: torch.Tensor, x: torch.Tensor, y: torch.Tensor, R: int):
    """Texel (face, y, x) with x, y in [-1, R] -> the texel that holds it under seamless filtering: itself inside the
    face, else the texel of the adjacent face across that edge at the same place along the edge. A tap beyond a CORNER
    (x and y both outside) is first clamped in y, i.e. it takes the x-neighbour's corner texel."""
    tab = torch.tensor(CUBE_WRAP, dtype=torch.int64, device=face.device)          # (6,4,3)
    xo = (x < 0) | (x >= R)
    yo = (y < 0) | (y >= R)
    y = torch.where(xo & yo, y.clamp(0, R - 1), y)
    yo = yo & ~xo
    side = torch.where(xo, torch.where(x < 0, 0, 1), torch.where(y < 0, 2, 3))
    k = torch.where(xo, y, x).clamp(0, R - 1)
    ent = tab[face, side]                                                          # (...,3)
    vals = torch.stack([torch.zeros_like(k), torch.full_like(k, R - 1), k, R - 1 - k], dim=-1)
    xn = torch.gather(vals, -1, ent[..., 1:2]).squeeze(-1)
    yn = torch.gather(vals, -1, ent[..., 2:3]).squeeze(-1)
    out = xo | yo
    return torch.where(out, ent[..., 0], face), torch.where(out, yn, y), torch.where(out, xn, x)


def cube_sample(texture: torch.Tensor, u: torch.Tensor, seamless: bool = False) -> torch.Tensor:
    """Bilinear fetch, texture (6,R,R,C), u (...,3). ``seamless=False`` (E11): clamp-to-edge inside the face.
    ``seamless=True`` (E11-alt, what nvdiffrast's boundary_mode='cube' does for the reference's visuals,
    models/uv_map_gaussian3d.py:259): taps beyond an edge are fetched from the adjacent face (``cube_wrap_tap``)."""
    R = texture.shape[1]
    face, sx, sy = cube_face_coords(u)
    fx = (sx + 1.0) * (0.5 * R) - 0.5
    fy = (sy + 1.0) * (0.5 * R) - 0.5
    x0f = torch.floor(fx.detach())
    y0f = torch.floor(fy.detach())
    wx = fx - x0f
    wy = fy - y0f
    x0 = x0f.to(torch.int64).clamp(-1, R - 1)
    y0 = y0f.to(torch.int64).clamp(-1, R - 1)
    if seamless:
        taps = [texture[cube_wrap_tap(face, x0 + dx, y0 + dy, R)] for dy in (0, 1) for dx in (0, 1)]
        t00, t01, t10, t11 = taps
    else:
        x0c, x1c = x0.clamp(0, R - 1), (x0 + 1).clamp(0, R - 1)
        y0c, y1c = y0.clamp(0, R - 1), (y0 + 1).clamp(0, R - 1)
        t00 = texture[face, y0c, x0c]
        t01 = texture[face, y0c, x1c]
        t10 = texture[face, y1c, x0c]
        t11 = texture[face, y1c, x1c]
    wx, wy = wx[..., None], wy[..., None]
    top = t00 + wx * (t01 - t00)
    bot = t10 + wx * (t11 - t10)
    return top + wy * (bot - top)


# ---------------------------------------------------------------------------------------------
# per-pixel stage (E5-E12)
# ---------------------------------------------------------------------------------------------

def rasterize(means3D, means2D, shs, opacities, scales, rotations, uvs, gradient_uvs, texture,
              settings: RasterSettings, sw: Switches = Switches(), colors_precomp=None,
              extra_attrs=None, max_elems: int = 6_000_000, tile_subset=None, return_aux=False,
              dual_no_sh: bool = False, cov3Ds_precomp=None):
    """Returns (image(3,H,W), depth(1,H,W), norm(3,H,W), alpha(1,H,W), radii(N,), extra|None)
    [+ aux dict]. ``texture=None`` selects the plain-3DGS colour path (``diff_gauss``): colour is
    ``colors_precomp`` (N,3) if given, else max(0, SH_full(shs)+0.5) with shs (N,(deg+1)^2,3).

    ``tile_subset``: optional iterable of tile indices to render (CPU-baseline sampling); pixels of
    other tiles stay zero.

    ``dual_no_sh``: additionally blend the colour the SAME splats have with
    ``sh_degree = 0`` — max(0, C0*tex + 0.5) — into a second image, returned as
    ``aux["image_no_sh"]``; identical to a second call with ``sh_degree=0`` (what the reference does
    at models/texture_gaussian3d.py:375-389,505-511)."""
    st = settings
    dt, dev = means3D.dtype, means3D.device
    H, W = int(st.image_height), int(st.image_width)
    textured = texture is not None
    N = means3D.shape[0]

    if textured:
        assert cov3Ds_precomp is None, "cov3Ds_precomp is a diff_gauss (untextured) argument"
        pre = preprocess(means3D, means2D, scales, rotations, opacities, shs, st, sw)
    else:
        pre = preprocess(means3D, means2D, scales, rotations, opacities, None, st, sw, cov3Ds_precomp=cov3Ds_precomp)
        if colors_precomp is not None:
            pre["csh"] = colors_precomp - 0.5
        else:
            m = pre["m"]
            dirs = m / m.norm(dim=1, keepdim=True)
            pre["csh"] = C0 * shs[:, 0] + sh_rest(int(st.sh_degree), shs[:, 1:], dirs)

    tile_of, gid_of = build_tile_lists(pre, np.float64 if dt == torch.float64 else np.float32)
    gx, gy = pre["grid"]
    V = st.viewmatrix.to(dt)
    bg = st.bg.to(dt)
    E = 0 if extra_attrs is None else extra_attrs.shape[1]

    npix_pad = gx * gy * TILE * TILE
    acc_c = torch.zeros(npix_pad, 3, dtype=dt, device=dev)
    acc_c0 = torch.zeros(npix_pad, 3, dtype=dt, device=dev) if (dual_no_sh and textured) else None
    acc_d = torch.zeros(npix_pad, dtype=dt, device=dev)
    acc_n = torch.zeros(npix_pad, 3, dtype=dt, device=dev)
    acc_a = torch.zeros(npix_pad, dtype=dt, device=dev)
    acc_e = torch.zeros(npix_pad, E, dtype=dt, device=dev) if E else None
    final_T = torch.ones(npix_pad, dtype=dt, device=dev)
    n_contrib = torch.zeros(npix_pad, dtype=torch.int64, device=dev)
    ambiguous = torch.zeros(npix_pad, dtype=torch.bool, device=dev)
    grazing = torch.zeros(npix_pad, dtype=torch.bool, device=dev)
    texel_edge = torch.zeros(npix_pad, dtype=torch.bool, device=dev)
    n_blend = 0
    pair_contributes = np.zeros(tile_of.shape[0], dtype=bool)   # some pixel of the tile blends this (tile,Gaussian) pair

    # tile segments
    if tile_of.shape[0] > 0:
        tiles, starts, counts = np.unique(tile_of, return_index=True, return_counts=True)
    else:
        tiles, starts, counts = np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64)
    if tile_subset is not None:
        keep = np.isin(tiles, np.asarray(list(tile_subset), dtype=np.int64))
        tiles, starts, counts = tiles[keep], starts[keep], counts[keep]
    order = np.argsort(counts, kind="stable")
    tiles, starts, counts = tiles[order], starts[order], counts[order]

    lx = torch.arange(TILE, device=dev)
    loc_x = lx.repeat(TILE)                       # pixel p = row*16+col
    loc_y = lx.repeat_interleave(TILE)
    gid_all = torch.from_numpy(gid_of).to(dev)

    i = 0
    ntile = tiles.shape[0]
    while i < ntile:
        # batch of tiles with similar list length
        Lmax = int(counts[i])
        j = i
        while j < ntile and (j - i + 1) * int(counts[j]) * TILE * TILE <= max(max_elems, int(counts[i]) * 256):
            Lmax = int(counts[j])
            j += 1
        j = max(j, i + 1)
        Lmax = int(counts[j - 1])
        B = j - i
        t_ids = torch.from_numpy(tiles[i:j]).to(dev)
        cnt = torch.from_numpy(counts[i:j]).to(dev)
        stt = torch.from_numpy(starts[i:j]).to(dev)
        ar = torch.arange(Lmax, device=dev)
        valid = ar[None, :] < cnt[:, None]                                   # (B,L)
        src = torch.where(valid, stt[:, None] + ar[None, :], torch.zeros_like(ar)[None, :])
        ids = gid_all[src]                                                   # (B,L)
        ids = torch.where(valid, ids, torch.zeros_like(ids))

        tx = t_ids % gx
        ty = t_ids // gx
        px = tx[:, None] * TILE + loc_x[None, :]                             # (B,256)
        py = ty[:, None] * TILE + loc_y[None, :]
        inside = (px < W) & (py < H)
        pflat = t_ids[:, None] * (TILE * TILE) + (loc_y * TILE + loc_x)[None, :]   # padded-pixel index

        xy = pre["xy"][ids]                                                  # (B,L,2)
        con = pre["conic"][ids]
        op = pre["opacity"][ids]
        dx = xy[:, None, :, 0] - px[:, :, None].to(dt)                       # (B,256,L)
        dy = xy[:, None, :, 1] - py[:, :, None].to(dt)
        power = -0.5 * (con[:, None, :, 0] * dx * dx + con[:, None, :, 2] * dy * dy) - con[:, None, :, 1] * dx * dy
        alpha = torch.clamp_max(op[:, None, :] * torch.exp(power), ALPHA_MAX)
        ok = valid[:, None, :] & inside[:, :, None] & (power <= 0) & (alpha >= ALPHA_MIN)   # E5
        a_eff = torch.where(ok, alpha, torch.zeros_like(alpha))
        with torch.no_grad():
            T_test = torch.cumprod(1.0 - a_eff, dim=2)
            stop = ok & (T_test < T_STOP)
            stopped = torch.cumsum(stop.to(torch.int32), dim=2) > 0
            include = ok & ~stopped
            # threshold-proximity flags for the parity tests
            rel = 4e-6 if dt == torch.float32 else 1e-12
            near_a = valid[:, None, :] & inside[:, :, None] & ~stopped & \
                ((alpha - ALPHA_MIN).abs() < rel * THRESH_ULPS * ALPHA_MIN)
            near_t = ok & ((T_test - T_STOP).abs() < rel * THRESH_ULPS * T_STOP)
            amb = (near_a | near_t).any(dim=2)
            if Lmax > 1:
                # depth of the previous CONTRIBUTION of the pixel (not necessarily the previous list entry)
                zl = pre["depth"][ids].detach()                                  # (B,L), ascending inside a tile
                pos = torch.where(include, ar[None, None, :], torch.full_like(ar, -1)[None, None, :])
                prev = torch.cummax(pos, dim=2).values
                prev = torch.cat([torch.full_like(prev[:, :, :1], -1), prev[:, :, :-1]], dim=2)
                zprev = torch.gather(zl[:, None, :].expand(-1, include.shape[1], -1), 2, prev.clamp_min(0))
                tie = include & (prev >= 0) & ((zl[:, None, :] - zprev).abs() <= DEPTH_TIE_REL * zl[:, None, :].abs())
                amb = amb | tie.any(dim=2)
            last = torch.where(include, ar[None, None, :] + 1, torch.zeros_like(ar)[None, None, :]).amax(dim=2)
        a_inc = torch.where(include, alpha, torch.zeros_like(alpha))
        T_incl = torch.cumprod(1.0 - a_inc, dim=2)
        T_excl = torch.cat([torch.ones_like(T_incl[:, :, :1]), T_incl[:, :, :-1]], dim=2)
        wgt = a_inc * T_excl                                                 # (B,256,L)

        final_T = final_T.index_put((pflat.reshape(-1),), T_incl[:, :, -1].reshape(-1))
        n_contrib[pflat.reshape(-1)] = last.reshape(-1)
        ambiguous[pflat.reshape(-1)] = amb.reshape(-1)

        with torch.no_grad():
            hit = include.any(dim=1)                                            # (B,L)
            hb, hl = torch.nonzero(hit, as_tuple=True)
            pair_contributes[(stt[hb] + hl).cpu().numpy()] = True
        # sparse heavy part
        bi, pi, li = torch.nonzero(include, as_tuple=True)
        n_blend += int(bi.shape[0])
        if bi.shape[0] > 0:
            g = ids[bi, li]
            w_s = wgt[bi, pi, li]
            pix = pflat[bi, pi]
            nrm = pre["normal"][g]
            if textured:
                ndc_x = (2.0 * px[bi, pi].to(dt) + 1.0) / W - 1.0
                ndc_y = (2.0 * py[bi, pi].to(dt) + 1.0) / H - 1.0
                vray = torch.stack([ndc_x * st.tanfovx, ndc_y * st.tanfovy, torch.ones_like(ndc_x)], dim=-1)
                d_w = vray @ V[:3, :3].t()                                    # world direction of the pixel ray
                mvec = pre["m"][g]
                nd = (nrm * d_w).sum(-1)
                nm = (nrm * mvec).sum(-1)
                safe = nd.abs() >= ND_EPS                                     # E9
                with torch.no_grad():
                    # conditioning flag: t = n.m / n.d amplifies rounding by 1/cos(angle(n, d));
                    # below GRAZING_COS fp32 cannot hold 1e-4 / 1e-3 (the fp32 and fp64 oracles differ by more)
                    cosang = nd.abs() / (d_w.norm(dim=-1) * nrm.norm(dim=-1).clamp_min(1e-30))
                    grazing[pix[cosang < GRAZING_COS]] = True
                nd_s = torch.where(safe, nd, torch.ones_like(nd))
                tpar = torch.where(safe, nm / nd_s, torch.zeros_like(nd))
                delta = torch.where(safe[:, None], tpar[:, None] * d_w - mvec, torch.zeros_like(mvec))
                if sw.stopgrad_delta:
                    delta = delta.detach()
                Jg = gradient_uvs[g].reshape(-1, 3, 3)
                u = uvs[g] + (Jg @ delta[:, :, None]).squeeze(-1)            # E10
                with torch.no_grad():
                    # conditioning flag: clamp-to-edge cube sampling (E11) is discontinuous across face
                    # edges; a contribution whose two largest |u'| components tie within 1e-4 (relative)
                    # picks its face by rounding
                    ua = u.detach().abs()
                    top2 = torch.topk(ua, 2, dim=-1).values
                    ambiguous[pix[(top2[:, 0] - top2[:, 1]) < FACE_TIE_REL * top2[:, 0]]] = True
                    # conditioning flag: bilinear cell chosen by rounding (derivative jumps at texel boundaries)
                    Rt = texture.shape[1]
                    _, sxd, syd = cube_face_coords(u.detach())
                    fxd, fyd = (sxd + 1.0) * (0.5 * Rt) - 0.5, (syd + 1.0) * (0.5 * Rt) - 0.5
                    # the intersection amplifies rounding by 1 / cos(angle(n, d)): so does the margin
                    near = torch.minimum((fxd - fxd.round()).abs(), (fyd - fyd.round()).abs()) * cosang.clamp_min(GRAZING_COS) < TEXEL_TIE * Rt
                    texel_edge[pix[near]] = True
                tex = cube_sample(texture, u, sw.seamless_cube)               # E11
                col_pre = C0 * tex + pre["csh"][g] + 0.5
                with torch.no_grad():
                    texel_edge[pix[(col_pre.abs() < CLAMP_TIE).any(dim=-1)]] = True   # clamp mask decided by rounding
                col = torch.clamp_min(col_pre, 0.0)                          # E12
                if acc_c0 is not None:
                    acc_c0 = acc_c0.index_add(0, pix, w_s[:, None] * torch.clamp_min(C0 * tex + 0.5, 0.0))
                if sw.depth_of_intersection:
                    zc = (delta + mvec + st.campos.to(dt)) @ V[:3, 2] + V[3, 2]
                else:
                    zc = pre["depth"][g]
            else:
                col = torch.clamp_min(pre["csh"][g] + 0.5, 0.0) if colors_precomp is None else colors_precomp[g]
                zc = pre["depth"][g]
            acc_c = acc_c.index_add(0, pix, w_s[:, None] * col)
            acc_d = acc_d.index_add(0, pix, w_s * zc)                        # E7
            acc_n = acc_n.index_add(0, pix, w_s[:, None] * nrm)              # E8
            acc_a = acc_a.index_add(0, pix, w_s)
            if E:
                acc_e = acc_e.index_add(0, pix, w_s[:, None] * extra_attrs[g])
        i = j

    def unpad(t, ch):
        t = t.reshape(gy, gx, TILE, TILE, ch).permute(4, 0, 2, 1, 3).reshape(ch, gy * TILE, gx * TILE)
        return t[:, :H, :W]

    Tf = unpad(final_T.reshape(-1, 1), 1)
    image = unpad(acc_c, 3) + Tf * bg[:, None, None]                          # E6
    depth = unpad(acc_d.reshape(-1, 1), 1)
    norm = unpad(acc_n, 3)
    alpha_img = unpad(acc_a.reshape(-1, 1), 1)
    extra = unpad(acc_e, E) if E else None
    out = (image, depth, norm, alpha_img, pre["radii"], extra)
    image_no_sh = (unpad(acc_c0, 3) + Tf * bg[:, None, None]) if acc_c0 is not None else None
    if return_aux:
        aux = dict(image_no_sh=image_no_sh, final_T=Tf[0].detach(), n_contrib=unpad(n_contrib.reshape(-1, 1), 1)[0],
                   ambiguous=unpad(ambiguous.reshape(-1, 1), 1)[0] | unpad(grazing.reshape(-1, 1), 1)[0],
                   # for GRADIENT comparisons only (values are continuous across texel boundaries)
                   grad_ambiguous=unpad(ambiguous.reshape(-1, 1), 1)[0] | unpad(grazing.reshape(-1, 1), 1)[0] | unpad(texel_edge.reshape(-1, 1), 1)[0],
                   threshold_ambiguous=unpad(ambiguous.reshape(-1, 1), 1)[0],
                   grazing=unpad(grazing.reshape(-1, 1), 1)[0], texel_boundary=unpad(texel_edge.reshape(-1, 1), 1)[0], num_pairs=int(tile_of.shape[0]),
                   num_visible=int(pre["visible"].sum()), num_blend=n_blend, pre=pre,
                   tile_of=tile_of, gid_of=gid_of, pair_contributes=pair_contributes)
        return out + (aux,)
    return out
