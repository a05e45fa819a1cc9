"""Small forward+backward (+dual, + plain 3DGS mode) for compute-sanitizer runs."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from texture_gs_b200 import uv_tex_render, uv_tex_render_dual
from texture_gs_b200.scene import sphere_shell_scene, orbit_cameras, output_cotangents
g = sphere_shell_scene(3000, 64, device="cuda")
cam = orbit_cameras(1, 200, 120, device="cuda")[0]
cot = output_cotangents(120, 200, device="cuda")
bg = torch.tensor([0.1, 0.2, 0.3], device="cuda")
for fn in (uv_tex_render, uv_tex_render_dual):
    pkg = fn(cam, g, None, bg)
    outs = [pkg["render"], pkg["depth"], pkg["norm"], pkg["alpha"]]
    cots = list(cot)
    if "render_no_sh" in pkg:
        outs.append(pkg["render_no_sh"]); cots.append(cot[0])
    torch.autograd.backward(outs, cots)
    g.zero_grad()
torch.cuda.synchronize()
print("sanitize run ok")
# cold paths: extra_attrs channels (two channel groups) and cov3Ds_precomp in the plain 3DGS mode
import math
from texture_gs_b200 import GaussianRasterizationSettings, GaussianRasterizer
t = g.tensors()
st = GaussianRasterizationSettings(120, 200, math.tan(cam.FoVx / 2), math.tan(cam.FoVy / 2), bg, 1.0, cam.world_view_transform,
                                   cam.full_proj_transform, 3, cam.camera_center, False, False)
ex = torch.randn(t["xyz"].shape[0], 11, device="cuda", requires_grad=True)
out = GaussianRasterizer(st)(means3D=t["xyz"], means2D=torch.zeros_like(t["xyz"], requires_grad=True), opacities=t["opacity"], shs=t["shs"],
                             scales=t["scaling"], rotations=t["rotation"], uvs=t["uvs"], gradient_uvs=t["grad_uvs"], texture=t["texture"],
                             extra_attrs=ex)
(out[0].sum() + out[5].sum()).backward()
g.zero_grad()
q = t["rotation"].detach()
r, x, y, z = q.unbind(-1)
R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y), 2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                 2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=-1).reshape(-1, 3, 3)
Lm = R * t["scaling"].detach()[:, None, :]
S = Lm @ Lm.transpose(1, 2)
cov = torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], dim=-1).contiguous().requires_grad_(True)
out = GaussianRasterizer(st)(means3D=t["xyz"], means2D=torch.zeros_like(t["xyz"], requires_grad=True), opacities=t["opacity"],
                             colors_precomp=torch.rand_like(t["xyz"]), cov3Ds_precomp=cov)
(out[0].sum() + out[1].sum()).backward()
torch.cuda.synchronize()
print("sanitize cold paths ok", float(ex.grad.abs().sum()), float(cov.grad.abs().sum()))
