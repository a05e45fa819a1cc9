"""Where the training-shaped step spends its GPU time beyond the rasterizer (single stream, CUDA events, headline config):
render fwd+bwd with fixed cotangents / + uint8 decode / + photometric loss / + geometry losses. One JSON line."""
import json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from texture_gs_b200 import uv_tex_render
from texture_gs_b200.dist import GradBucket
from texture_gs_b200.losses import geometry_losses, photometric_loss
from texture_gs_b200.scene import orbit_cameras, output_cotangents, sphere_shell_scene
N, W, H, R = 500000, 1920, 1080, 2048
dev = torch.device("cuda")
g = sphere_shell_scene(N, R, device=dev)
cams = orbit_cameras(8, W, H, device=dev)
bg = torch.zeros(3, device=dev)
cot = output_cotangents(H, W, device=dev)
bucket = GradBucket(g.tensors())
u8 = (torch.rand(3, H, W, device=dev) * 255).to(torch.uint8)
m8 = (torch.rand(1, H, W, device=dev) > 0.2).to(torch.uint8) * 255
n8 = (torch.nn.functional.normalize(torch.randn(3, H, W, device=dev), dim=0) * 127).to(torch.int8)


def variant(level):
    def f(i):
        with bucket.fused():
            pkg = uv_tex_render(cams[i % 8], g, None, bg)
            if level == 0:
                torch.autograd.backward([pkg["render"], pkg["depth"], pkg["norm"], pkg["alpha"]], list(cot))
                return
            gt = u8.float().mul_(1 / 255.0); ga = m8.float().mul_(1 / 255.0); gn = n8.float().mul_(1 / 127.0)
            if level == 1:
                torch.autograd.backward([pkg["render"], pkg["depth"], pkg["norm"], pkg["alpha"]], list(cot))
                return
            loss = photometric_loss(pkg["render"], gt, 0.2)[0]
            if level >= 3:
                la, ln, ls = geometry_losses(pkg["alpha"], pkg["norm"], ga, gn, gt)
                loss = loss + la + 0.1 * ln + 0.5 * ls
            else:
                loss = loss + (pkg["alpha"] * cot[3]).sum() + (pkg["norm"] * cot[2]).sum()
            loss.backward()
    return f


out = {}
for name, lvl in (("render fwd+bwd, fixed cotangents", 0), ("+ uint8/int8 decode", 1), ("+ photometric loss (others via dot products)", 2),
                  ("+ geometry losses (= the e2e step)", 3)):
    f = variant(lvl)
    for i in range(4):
        f(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(16):
        f(i)
    e1.record(); torch.cuda.synchronize()
    out[name] = round(e0.elapsed_time(e1) / 16, 4)
print(json.dumps(out))
