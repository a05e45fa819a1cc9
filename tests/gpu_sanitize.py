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
