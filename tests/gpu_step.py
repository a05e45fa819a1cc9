"""Tiny driver for profiling: python tests/gpu_step.py N W H R iters [fwd_only]"""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from texture_gs_b200 import uv_tex_render, last_stats
from texture_gs_b200.scene import sphere_shell_scene, orbit_cameras, output_cotangents
N, W, H, R, iters = [int(x) for x in sys.argv[1:6]]
fwd_only = len(sys.argv) > 6
g = sphere_shell_scene(N, R, device="cuda")
cams = [c.to("cuda") for c in orbit_cameras(4, W, H)]
bg = torch.zeros(3, device="cuda")
cots = [c.cuda() for c in output_cotangents(H, W)]
for it in range(iters):
    torch.cuda.synchronize(); t0 = time.time()
    if fwd_only:
        with torch.no_grad():
            pkg = uv_tex_render(cams[it % 4], g, None, bg)
    else:
        pkg = uv_tex_render(cams[it % 4], g, None, bg)
        L = (pkg["render"] * cots[0]).sum() + (pkg["depth"] * cots[1]).sum() + (pkg["norm"] * cots[2]).sum() + (pkg["alpha"] * cots[3]).sum()
        L.backward()
        g.zero_grad()
    torch.cuda.synchronize()
    print(it, "ms", (time.time() - t0) * 1e3, last_stats())
