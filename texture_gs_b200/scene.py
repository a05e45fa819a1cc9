"""Seeded synthetic inputs for the rasterizer hot path.

Nothing here is on the measured path: it only *produces* the tensors the operator
``uv_tex_render`` (reference ``render/uv_tex_render.py:7-77``) reads from its two duck-typed
arguments:

* ``viewpoint_camera`` — ``FoVx FoVy image_height image_width world_view_transform
  full_proj_transform camera_center`` (reference ``utils/cameras.py:21-78``), built with the
  reference's matrix conventions (``utils/graphics.py:38-71``): row-vector convention,
  ``world_view_transform = W2C^T``, ``full_proj_transform = W2C^T @ P^T``, camera looks down +z,
  y down, znear 0.01, zfar 100.
* ``gaussians`` — ``get_xyz get_opacity get_scaling get_rotation get_shs get_texture get_uvs
  get_grad_uvs active_sh_degree`` (reference ``models/texture_gaussian3d.py:195-240``).

All randomness comes from CPU ``torch.Generator`` objects so a given seed yields the same scene
on every machine.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch

C0 = 0.28209479177387814  # SH DC constant; texture value t encodes rgb = C0*t + 0.5
                          # (reference models/texture_gaussian3d.py:16-21)


# --------------------------------------------------------------------------------------------
# camera
# --------------------------------------------------------------------------------------------

def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float) -> torch.Tensor:
    """Same matrix as reference ``utils/graphics.py:51-71`` (not transposed yet)."""
    ty = math.tan(fovy / 2)
    tx = math.tan(fovx / 2)
    top, right = ty * znear, tx * znear
    P = torch.zeros(4, 4, dtype=torch.float32)
    P[0, 0] = 2.0 * znear / (2 * right)
    P[1, 1] = 2.0 * znear / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


class SyntheticCamera:
    """Duck-type of reference ``utils/cameras.MiniCam`` (``utils/cameras.py:67-78``)."""

    def __init__(self, width: int, height: int, fovy: float, w2c: torch.Tensor,
                 znear: float = 0.01, zfar: float = 100.0, device="cpu"):
        self.image_width = int(width)
        self.image_height = int(height)
        self.FoVy = float(fovy)
        # FoVx from the aspect ratio (square pixels): tan(fx/2) = aspect * tan(fy/2)
        self.FoVx = 2.0 * math.atan(math.tan(fovy / 2) * width / height)
        self.znear, self.zfar = znear, zfar
        w2c = w2c.to(torch.float32)
        self.world_view_transform = w2c.t().contiguous().to(device)
        P = projection_matrix(znear, zfar, self.FoVx, self.FoVy).t().to(device)
        self.projection_matrix = P
        self.full_proj_transform = (self.world_view_transform @ P).contiguous()
        self.camera_center = torch.inverse(self.world_view_transform.cpu())[3, :3].contiguous().to(device)

    def to(self, device):
        for k in ("world_view_transform", "projection_matrix", "full_proj_transform", "camera_center"):
            setattr(self, k, getattr(self, k).to(device))
        return self


def look_at_w2c(center: torch.Tensor, target: torch.Tensor, world_up: torch.Tensor) -> torch.Tensor:
    """4x4 world-to-camera matrix for a camera at ``center`` looking at ``target``.

    Camera axes follow the reference's COLMAP convention (+z forward, +y down, +x right;
    ``dataset/dataset_readers.py:209``).
    """
    z = target - center
    z = z / z.norm()
    down = -world_up
    x = torch.linalg.cross(down, z)
    if float(x.norm()) < 1e-6:                      # looking straight along up: pick any x
        x = torch.linalg.cross(torch.tensor([1.0, 0.0, 0.0], dtype=z.dtype), z)
    x = x / x.norm()
    y = torch.linalg.cross(z, x)
    R = torch.stack([x, y, z], dim=0)               # rows = camera axes in world = W2C rotation
    w2c = torch.eye(4, dtype=torch.float64)
    w2c[:3, :3] = R
    w2c[:3, 3] = -R @ center
    return w2c


def orbit_cameras(n: int, width: int, height: int, radius: float = 2.5,
                  fovy_deg: float = 45.0, seed: int = 1, device="cpu"):
    """``n`` cameras with centres uniform on a sphere, looking at the origin."""
    g = torch.Generator().manual_seed(seed)
    dirs = torch.randn(n, 3, generator=g, dtype=torch.float64)
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    up = torch.tensor([0.0, -1.0, 0.0], dtype=torch.float64)   # COLMAP-style worlds: up = -y
    origin = torch.zeros(3, dtype=torch.float64)
    cams = []
    for i in range(n):
        w2c = look_at_w2c(dirs[i] * radius, origin, up)
        cams.append(SyntheticCamera(width, height, math.radians(fovy_deg), w2c, device=device))
    return cams


# --------------------------------------------------------------------------------------------
# Gaussians + texture
# --------------------------------------------------------------------------------------------

def _quat_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    ar, ax, ay, az = a.unbind(-1)
    br, bx, by, bz = b.unbind(-1)
    return torch.stack([
        ar * br - ax * bx - ay * by - az * bz,
        ar * bx + ax * br + ay * bz - az * by,
        ar * by - ax * bz + ay * br + az * bx,
        ar * bz + ax * by - ay * bx + az * br], dim=-1)


class SyntheticGaussians:
    """Duck-type of the attributes ``uv_tex_render`` reads from ``TextureGaussian3D``
    (reference ``render/uv_tex_render.py:15,34,42-53``). Leaf tensors carry ``requires_grad``
    exactly where the reference's do (xyz, opacity, scaling, rotation, shs, texture, uvs;
    *not* grad_uvs — ``models/texture_gaussian3d.py:227``)."""

    def __init__(self, xyz, opacity, scaling, rotation, shs, texture, uvs, grad_uvs,
                 active_sh_degree: int):
        self._t = dict(xyz=xyz, opacity=opacity, scaling=scaling, rotation=rotation, shs=shs,
                       texture=texture, uvs=uvs, grad_uvs=grad_uvs)
        self.active_sh_degree = int(active_sh_degree)

    get_xyz = property(lambda s: s._t["xyz"])
    get_opacity = property(lambda s: s._t["opacity"])
    get_scaling = property(lambda s: s._t["scaling"])
    get_rotation = property(lambda s: s._t["rotation"])
    get_shs = property(lambda s: s._t["shs"])
    get_texture = property(lambda s: s._t["texture"])
    get_uvs = property(lambda s: s._t["uvs"])
    get_grad_uvs = property(lambda s: s._t["grad_uvs"])

    GRAD_NAMES = ("xyz", "opacity", "scaling", "rotation", "shs", "texture", "uvs")

    def tensors(self):
        return dict(self._t)

    def to(self, device=None, dtype=None, requires_grad: Optional[bool] = None):
        out = {}
        for k, v in self._t.items():
            if v is None:
                out[k] = None
                continue
            w = v.detach().to(device=device, dtype=dtype).contiguous()
            rg = v.requires_grad if requires_grad is None else (requires_grad and k in self.GRAD_NAMES)
            out[k] = w.requires_grad_(bool(rg))
        return SyntheticGaussians(active_sh_degree=self.active_sh_degree, **out)

    def zero_grad(self):
        for v in self._t.values():
            if v is not None:
                v.grad = None


def band_limited_texture(R: int, seed: int = 2, cells: int = 0, device="cpu") -> torch.Tensor:
    """(6,R,R,3) cube texture in SH-DC encoding, rgb = smooth noise in [0.1,0.9].

    Low-resolution uniform noise (``cells`` per face edge, default R/16 but at least 4) upsampled
    bicubically, so neighbouring texels differ by O(1/16) of the dynamic range."""
    cells = cells or max(4, R // 16)
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(6, 3, cells, cells, generator=g, dtype=torch.float32).to(device)
    up = torch.nn.functional.interpolate(low, size=(R, R), mode="bicubic", align_corners=False)
    # keep rgb inside [0.1, 0.9]: a texel clamped to exactly 0 would put the colour clamp of
    # spec E12 (max(0, .)) on a knife-edge and make its gradient mask implementation-defined
    rgb = (0.1 + 0.8 * up.clamp_(0.0, 1.0)).permute(0, 2, 3, 1).contiguous()
    return (rgb - 0.5) / C0


def sphere_shell_scene(n: int, tex_res: int, sh_degree: int = 3, seed: int = 0, tex_seed: int = 2,
                       coverage: float = 4.0, device="cpu", requires_grad: bool = True,
                       max_sh_degree: int = 3) -> SyntheticGaussians:
    """Scene "sphere-shell" of SURVEY §8d: flat discs tangent to a radius≈1 shell, uv = direction,
    J = (I - uv uv^T)/|x| (the exact Jacobian of x -> x/|x|)."""
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(n, 3, generator=g, dtype=torch.float64)
    d = d / d.norm(dim=1, keepdim=True)
    r = 1.0 + 0.02 * torch.randn(n, 1, generator=g, dtype=torch.float64)
    xyz = r * d
    # rotation: local z -> d, then an in-plane spin about local z
    qa = torch.stack([1.0 + d[:, 2], -d[:, 1], d[:, 0], torch.zeros(n, dtype=torch.float64)], dim=-1)
    bad = qa.norm(dim=1) < 1e-6
    qa[bad] = torch.tensor([0.0, 1.0, 0.0, 0.0], dtype=torch.float64)
    qa = qa / qa.norm(dim=1, keepdim=True)
    th = 2 * math.pi * torch.rand(n, generator=g, dtype=torch.float64)
    qs = torch.stack([torch.cos(th / 2), torch.zeros_like(th), torch.zeros_like(th), torch.sin(th / 2)], -1)
    rot = _quat_mul(qa, qs)
    # scales: log-uniform in [a,b], N*pi*(2s)^2 ~= coverage * 4*pi
    s0 = math.sqrt(coverage / n)
    a, b = s0 / math.sqrt(2.0), s0 * math.sqrt(2.0)
    su = torch.exp(torch.rand(n, 2, generator=g, dtype=torch.float64) * (math.log(b) - math.log(a)) + math.log(a))
    scaling = torch.cat([su, torch.full((n, 1), math.exp(-20.0), dtype=torch.float64)], dim=1)
    opacity = 0.3 + 0.69 * torch.rand(n, 1, generator=g, dtype=torch.float64)
    uv = xyz / xyz.norm(dim=1, keepdim=True)
    J = (torch.eye(3, dtype=torch.float64)[None] - uv[:, :, None] * uv[:, None, :]) / xyz.norm(dim=1)[:, None, None]
    M = (max_sh_degree + 1) ** 2 - 1
    shs = 0.05 * torch.randn(n, M, 3, generator=g, dtype=torch.float64) if M > 0 else None
    tex = band_limited_texture(tex_res, seed=tex_seed, device=device)

    def f(t, rg):
        if t is None:
            return None
        return t.to(torch.float32).to(device).contiguous().requires_grad_(rg and requires_grad)

    return SyntheticGaussians(
        xyz=f(xyz, True), opacity=f(opacity, True), scaling=f(scaling, True), rotation=f(rot, True),
        shs=f(shs, True), texture=tex.requires_grad_(requires_grad), uvs=f(uv, True),
        grad_uvs=f(J.reshape(n, 9), False), active_sh_degree=sh_degree)


def output_cotangents(height: int, width: int, seed: int = 3, device="cpu"):
    """Fixed dense random cotangents (g_image, g_depth, g_norm, g_alpha) so that
    L = sum(image*g1)+sum(depth*g2)+sum(norm*g3)+sum(alpha*g4) exercises all four output grads."""
    g = torch.Generator().manual_seed(seed)
    mk = lambda c: torch.randn(c, height, width, generator=g, dtype=torch.float32).to(device)
    return mk(3), mk(1), mk(3), mk(1)


@dataclass
class Workload:
    name: str
    n_gaussians: int
    width: int
    height: int
    tex_res: int
    backward: bool
    renders_per_view: int = 1


# BASELINE.json configs (index -> workload); "dtu118" is replaced by the 300k synthetic
WORKLOADS = {
    "cfg0_10k_256": Workload("cfg0_10k_256", 10_000, 256, 256, 512, True),
    "cfg1_300k_800x600": Workload("cfg1_300k_800x600", 300_000, 800, 600, 1024, False, 2),
    "cfg2_500k_1080p": Workload("cfg2_500k_1080p", 500_000, 1920, 1080, 2048, True),
    "cfg4_1m_4k": Workload("cfg4_1m_4k", 1_000_000, 3840, 2160, 4096, False),
}
