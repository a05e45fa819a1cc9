"""Per-stage timing of one library build (TEXGS_LIB env) on the headline config; prints one JSON line."""
import json, os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from texture_gs_b200 import uv_tex_render
from texture_gs_b200.profiling import StageTimer
from texture_gs_b200.scene import sphere_shell_scene, orbit_cameras, output_cotangents
N, W, H, R = 500000, 1920, 1080, 2048
g = sphere_shell_scene(N, R, device="cuda")
cams = [c for c in orbit_cameras(8, W, H, device="cuda")]
bg = torch.zeros(3, device="cuda")
cot = output_cotangents(H, W, device="cuda")
from texture_gs_b200.dist import GradBucket
fused = "fused" in sys.argv
bucket = GradBucket(g.tensors()) if fused else None
tm = StageTimer(8)
for it in range(12):
    ctx = tm.view() if it >= 4 else None
    if ctx: ctx.__enter__()
    if fused:
        with bucket.fused():
            pkg = uv_tex_render(cams[it % 8], g, None, bg)
            torch.autograd.backward([pkg["render"], pkg["depth"], pkg["norm"], pkg["alpha"]], list(cot))
    else:
        pkg = uv_tex_render(cams[it % 8], g, None, bg)
        torch.autograd.backward([pkg["render"], pkg["depth"], pkg["norm"], pkg["alpha"]], list(cot))
    if ctx: ctx.__exit__(None, None, None)
    if not fused:
        g.zero_grad()
torch.cuda.synchronize()
s = tm.summary()
print(json.dumps({"lib": os.environ.get("TEXGS_LIB", "default"), **{k: round(v, 4) for k, v in s.items()}}))
