"""VERDICT r1 item 2 / north-star "staging of texture tiles into shared memory": how compact is the texel footprint of a
16x16 pixel tile? Measured on the GPU with the product kernels themselves: the backward with a cotangent that is non-zero
on ONE tile only leaves a texture gradient that is non-zero exactly on the texels that tile's contributions fetched.
Per sampled tile: distinct texels, faces touched, and the bytes a per-face bounding box (what a 2-D TMA box would have
to bring in) covers, at 16 bytes per texel.   python tools/texel_footprint.py [workload] [n_tiles]  -> one JSON line"""
import json, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from texture_gs_b200 import uv_tex_render
from texture_gs_b200.scene import WORKLOADS, orbit_cameras, sphere_shell_scene

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2_500k_1080p"]
ntiles = int(sys.argv[2]) if len(sys.argv) > 2 else 96
dev = torch.device("cuda")
g = sphere_shell_scene(wl.n_gaussians, wl.tex_res, sh_degree=3, seed=0, device=dev)
cam = orbit_cameras(32, wl.width, wl.height, seed=1, device=dev)[5]
bg = torch.zeros(3, device=dev)
gx, gy = (wl.width + 15) // 16, (wl.height + 15) // 16
gen = torch.Generator().manual_seed(0)
rows = []
R = wl.tex_res
for t in torch.randperm(gx * gy, generator=gen)[:ntiles].tolist():
    tx, ty = t % gx, t // gx
    pkg = uv_tex_render(cam, g, None, bg)
    cot = torch.zeros_like(pkg["render"])
    cot[:, ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16] = 1.0
    g.zero_grad()
    pkg["render"].backward(cot)
    touched = (g.get_texture.grad.abs().sum(dim=-1) > 0)            # (6,R,R)
    n = int(touched.sum())
    if n == 0:
        continue
    box_texels, faces = 0, 0
    for f in range(6):
        idx = touched[f].nonzero()
        if idx.numel() == 0:
            continue
        faces += 1
        (y0, x0), (y1, x1) = idx.min(dim=0).values.tolist(), idx.max(dim=0).values.tolist()
        box_texels += (y1 - y0 + 1) * (x1 - x0 + 1)
    contrib = float(pkg["alpha"][:, ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16].gt(0).sum())
    rows.append((n, faces, box_texels, contrib))
tt = torch.tensor(rows, dtype=torch.float64)
q = lambda c, p: float(torch.quantile(tt[:, c], p))
smem_budget = 160 * 1024            # what a CTA could spare beside its 66 KB record ring
fits = float((tt[:, 2] * 16 <= smem_budget).double().mean())
print(json.dumps({"workload": wl.name, "tiles_sampled": len(rows), "texels_per_tile_median": q(0, 0.5), "texels_per_tile_p90": q(0, 0.9),
                  "texels_per_pixel_median": q(0, 0.5) / 256.0, "faces_per_tile_mean": float(tt[:, 1].mean()),
                  "bbox_kb_median": q(2, 0.5) * 16 / 1024, "bbox_kb_p10": q(2, 0.1) * 16 / 1024, "bbox_kb_p90": q(2, 0.9) * 16 / 1024,
                  "bbox_fill_median": float((tt[:, 0] / tt[:, 2]).median()),
                  "tiles_whose_bbox_fits_160KB": fits, "tile_bytes_touched_kb_median": q(0, 0.5) * 16 / 1024}))
