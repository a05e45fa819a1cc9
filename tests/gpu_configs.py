"""Forward-only timing of BASELINE configs[1] (300k, 800x600, R1024, 2 renders/view) and configs[4]
(1M, 4K, R4096) + fwd+bwd of configs[0]; prints one JSON line per config."""
import json, sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from texture_gs_b200 import uv_tex_render, last_stats
from texture_gs_b200.scene import WORKLOADS, sphere_shell_scene, orbit_cameras, output_cotangents

def run(name, iters=12, warm=3):
    wl = WORKLOADS[name]
    g = sphere_shell_scene(wl.n_gaussians, wl.tex_res, device="cuda", requires_grad=wl.backward)
    cams = orbit_cameras(8, wl.width, wl.height, device="cuda")
    bg = torch.zeros(3, device="cuda")
    cot = output_cotangents(wl.height, wl.width, device="cuda") if wl.backward else None
    def one(i):
        if wl.backward:
            pkg = uv_tex_render(cams[i % 8], g, None, bg)
            torch.autograd.backward([pkg["render"], pkg["depth"], pkg["norm"], pkg["alpha"]], list(cot))
            g.zero_grad()
        else:
            with torch.no_grad():
                for r in range(wl.renders_per_view):
                    g.active_sh_degree = 3 if r == 0 else 0     # retexture.py renders with SH, then sh_degree=0
                    uv_tex_render(cams[i % 8], g, None, bg)
                g.active_sh_degree = 3
    if wl.renders_per_view == 2 and "dual" in sys.argv:
        from texture_gs_b200 import uv_tex_render_dual
        def one(i):                                     # both images in one pass
            with torch.no_grad():
                uv_tex_render_dual(cams[i % 8], g, None, bg)
    for i in range(warm): one(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): one(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(json.dumps({"workload": name, "ms_per_view": ms, "views_per_s": 1e3 / ms, "backward": wl.backward,
                      "renders_per_view": wl.renders_per_view, "stats": last_stats()._asdict(),
                      "mem_gb": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
    del g; torch.cuda.empty_cache()

for n in ([a for a in sys.argv[1:] if a != "dual"] or ["cfg0_10k_256", "cfg1_300k_800x600", "cfg4_1m_4k"]):
    run(n)
