"""ORACLE (test infrastructure, NOT product code) — scalar, loop-by-loop restatement of the same
spec as ``oracle/raster_ref.py`` (see that file's header: *** PARITY UNPINNED ***), written
independently with plain Python floats (float64) so that a vectorisation mistake in
``raster_ref`` cannot hide. Forward only, tiny inputs only (≤ a few hundred Gaussians, ≤ 64² px).

Spec items: E1 near cull, E2 EWA projection with +0.3 dilation, E3 radius/tile
rect, E4 (tile, depth, index) order, E5 blend rule, E6 background on colour only, E7 centre depth,
E8 facing disc normal, E9 ray–plane intersection, E10 first-order UV, E11 cube lookup with
clamp-to-edge, E12 colour = max(0, C0*tex + SH_rest + 0.5).
"""
from __future__ import annotations

import math

import numpy as np

TILE = 16
C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
      -1.0925484305920792, 0.5462742152960396)
C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
      -0.4570457994644658, 1.445305721320277, -0.5900435899266435)


def _rot(q):
    r, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)],
        [2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)],
        [2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]])


def _sh_rest(deg, sh, d):
    """sh (M,3) rest coefficients; reference utils/sh.py:57-112 bands >= 1."""
    out = np.zeros(3)
    if sh is None or deg <= 0:
        return out
    x, y, z = d
    out += -C1 * y * sh[0] + C1 * z * sh[1] - C1 * x * sh[2]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        out += (C2[0] * xy * sh[3] + C2[1] * yz * sh[4] + C2[2] * (2 * zz - xx - yy) * sh[5]
                + C2[3] * xz * sh[6] + C2[4] * (xx - yy) * sh[7])
        if deg > 2:
            out += (C3[0] * y * (3 * xx - yy) * sh[8] + C3[1] * xy * z * sh[9]
                    + C3[2] * y * (4 * zz - xx - yy) * sh[10] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[11]
                    + C3[4] * x * (4 * zz - xx - yy) * sh[12] + C3[5] * z * (xx - yy) * sh[13]
                    + C3[6] * x * (xx - 3 * yy) * sh[14])
    return out


def _cube(tex, u):
    R = tex.shape[1]
    x, y, z = u
    ax, ay, az = abs(x), abs(y), abs(z)
    if ax >= ay and ax >= az:
        m = ax
        face, sx, sy = (1, z, -y) if x < 0 else (0, -z, -y)
    elif ay >= az:
        m = ay
        face, sx, sy = (3, x, -z) if y < 0 else (2, x, z)
    else:
        m = az
        face, sx, sy = (5, -x, -y) if z < 0 else (4, x, -y)
    m = max(m, 1e-20)
    fx = (sx / m + 1) * 0.5 * R - 0.5
    fy = (sy / m + 1) * 0.5 * R - 0.5
    x0, y0 = math.floor(fx), math.floor(fy)
    wx, wy = fx - x0, fy - y0
    cl = lambda v: min(max(v, 0), R - 1)
    t00, t01 = tex[face, cl(y0), cl(x0)], tex[face, cl(y0), cl(x0 + 1)]
    t10, t11 = tex[face, cl(y0 + 1), cl(x0)], tex[face, cl(y0 + 1), cl(x0 + 1)]
    top = t00 + wx * (t01 - t00)
    bot = t10 + wx * (t11 - t10)
    return top + wy * (bot - top)


def rasterize_loop(means3D, shs, opacities, scales, rotations, uvs, gradient_uvs, texture, *,
                   H, W, tanfovx, tanfovy, bg, scale_modifier, viewmatrix, projmatrix, sh_degree, campos):
    """All array arguments are numpy float64. Returns image(3,H,W), depth(H,W), norm(3,H,W),
    alpha(H,W), radii(N)."""
    N = means3D.shape[0]
    V, PM = viewmatrix, projmatrix
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    recs = []
    radii = np.zeros(N, np.int32)
    for i in range(N):
        p = means3D[i]
        pv = p @ V[:3, :3] + V[3, :3]
        if pv[2] <= 0.2:
            continue
        ph = p @ PM[:3, :] + PM[3, :]
        pw = 1.0 / (ph[3] + 1e-7)
        ndc = ph[:3] * pw
        Rm = _rot(rotations[i])
        L = Rm * (scales[i] * scale_modifier)[None, :]
        Sig = L @ L.T
        limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
        tz = pv[2]
        tx = min(max(pv[0] / tz, -limx), limx) * tz
        ty = min(max(pv[1] / tz, -limy), limy) * tz
        fx, fy = W / (2 * tanfovx), H / (2 * tanfovy)
        J = np.array([[fx / tz, 0, -fx * tx / (tz * tz)], [0, fy / tz, -fy * ty / (tz * tz)]])
        T = J @ V[:3, :3].T
        cov = T @ Sig @ T.T
        a, b, c = cov[0, 0] + 0.3, cov[0, 1], cov[1, 1] + 0.3
        det = a * c - b * b
        if det == 0:
            continue
        conic = (c / det, -b / det, a / det)
        mid = 0.5 * (a + c)
        lam = mid + math.sqrt(max(0.1, mid * mid - det))
        rad = math.ceil(3 * math.sqrt(lam))
        x = ((ndc[0] + 1) * W - 1) * 0.5
        y = ((ndc[1] + 1) * H - 1) * 0.5
        cl = lambda v, hi: min(max(v, 0), hi)
        rx0, rx1 = cl(math.floor((x - rad) / TILE), gx), cl(math.floor((x + rad + TILE - 1) / TILE), gx)
        ry0, ry1 = cl(math.floor((y - rad) / TILE), gy), cl(math.floor((y + rad + TILE - 1) / TILE), gy)
        if (rx1 - rx0) * (ry1 - ry0) == 0:
            continue
        radii[i] = rad
        k = int(np.argmin(scales[i]))
        n = Rm[:, k].copy()
        m = p - campos
        if n @ m > 0:
            n = -n
        csh = _sh_rest(sh_degree, None if shs is None else shs[i], m / np.linalg.norm(m))
        recs.append(dict(i=i, x=x, y=y, conic=conic, z=pv[2], o=float(np.asarray(opacities).reshape(-1)[i]), n=n, m=m, csh=csh,
                         rect=(rx0, ry0, rx1, ry1)))
    img = np.zeros((3, H, W)); dep = np.zeros((H, W)); nrm = np.zeros((3, H, W)); alp = np.zeros((H, W))
    for ty_ in range(gy):
        for tx_ in range(gx):
            lst = [r for r in recs if r["rect"][0] <= tx_ < r["rect"][2] and r["rect"][1] <= ty_ < r["rect"][3]]
            lst.sort(key=lambda r: (r["z"], r["i"]))
            for py in range(ty_ * TILE, min(H, (ty_ + 1) * TILE)):
                for px in range(tx_ * TILE, min(W, (tx_ + 1) * TILE)):
                    T = 1.0
                    C = np.zeros(3); D = 0.0; Nn = np.zeros(3); A = 0.0
                    vx = ((2 * px + 1) / W - 1) * tanfovx
                    vy = ((2 * py + 1) / H - 1) * tanfovy
                    dw = np.array([vx, vy, 1.0]) @ V[:3, :3].T
                    for r in lst:
                        dx, dy = r["x"] - px, r["y"] - py
                        ca, cb, cc = r["conic"]
                        power = -0.5 * (ca * dx * dx + cc * dy * dy) - cb * dx * dy
                        if power > 0:
                            continue
                        al = min(0.99, r["o"] * math.exp(power))
                        if al < 1.0 / 255.0:
                            continue
                        Tt = T * (1 - al)
                        if Tt < 1e-4:
                            break
                        i = r["i"]
                        nd = r["n"] @ dw
                        if abs(nd) >= 1e-8:
                            t = (r["n"] @ r["m"]) / nd
                            delta = t * dw - r["m"]
                        else:
                            delta = np.zeros(3)
                        u = uvs[i] + gradient_uvs[i].reshape(3, 3) @ delta
                        col = np.maximum(C0 * _cube(texture, u) + r["csh"] + 0.5, 0.0)
                        w = al * T
                        C += w * col; D += w * r["z"]; Nn += w * r["n"]; A += w
                        T = Tt
                    img[:, py, px] = C + T * bg
                    dep[py, px] = D; nrm[:, py, px] = Nn; alp[py, px] = A
    return img, dep, nrm, alp, radii
