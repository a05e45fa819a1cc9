"""Stage-by-stage GPU diagnostic (not a pytest file): prints where CUDA and oracle first diverge."""
import ctypes as C, math, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from util import *
from texture_gs_b200 import _lib as L
from texture_gs_b200.rasterizer import last_stats

def main():
    n, w, h, r = 2000, 128, 96, 64
    g = sphere_shell_scene(n, r, sh_degree=3, seed=0, tex_seed=1)
    cam = orbit_cameras(1, w, h, seed=2)[0]
    cot = output_cotangents(h, w)
    ref, aux, gref = run_oracle(g, cam, bg=(0.2, 0.4, 0.6), cot=cot)
    torch.cuda.synchronize()
    got, stats, ggot = run_cuda(g, cam, bg=(0.2, 0.4, 0.6), cot=cot, debug=True)
    print("stats", stats, "oracle pairs", aux["num_pairs"], "visible", aux["num_visible"], "blend", aux["num_blend"])
    print("radii mismatches", int((got[4] != ref[4]).sum()))
    print(compare_images(got[:4], ref[:4], aux["ambiguous"]))
    for k, rg in gref.items():
        if rg is None: continue
        c = ggot[k]
        if k == "means2D": c, rg = c[:, :2], rg[:, :2]
        print("grad", k, "rel", rel_err(c.reshape(rg.shape), rg), "ref max", float(rg.abs().max()))
    # timing of a mid-size case
    from texture_gs_b200 import uv_tex_render
    for (N, W, H, R) in [(10_000, 256, 256, 512), (500_000, 1920, 1080, 2048)]:
        gg = sphere_shell_scene(N, R, device="cuda")
        cams = [c.to("cuda") for c in orbit_cameras(4, W, H)]
        bg = torch.zeros(3, device="cuda")
        cots = [c.cuda() for c in output_cotangents(H, W)]
        for it in range(3):
            torch.cuda.synchronize(); t0 = time.time()
            pkg = uv_tex_render(cams[it % 4], gg, None, bg)
            torch.cuda.synchronize(); t1 = time.time()
            Ls = (pkg["render"] * cots[0]).sum() + (pkg["depth"] * cots[1]).sum() + (pkg["norm"] * cots[2]).sum() + (pkg["alpha"] * cots[3]).sum()
            Ls.backward()
            torch.cuda.synchronize(); t2 = time.time()
            print(N, W, H, R, "fwd ms", (t1 - t0) * 1e3, "bwd ms", (t2 - t1) * 1e3, last_stats())
            gg.zero_grad()

if __name__ == "__main__":
    main()
