"""Warp-vote profile of the kernels on the CPU emulator (tests/simt): for every __ballot/__any/__all_sync call site of the
.cuh sources, how often it ran and how many lanes voted true — e.g. how many lanes of a pass really blend a splat, how
often a half-warp has run out of survivors while the other still works. Exact counts, no GPU:
    python tools/vote_profile.py N W H R          e.g. 500000 1920 1080 2048 (the headline config, ~3 min)"""
from pathlib import Path
import sys, math
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import torch
from simt import emu
from texture_gs_b200.scene import orbit_cameras, sphere_shell_scene, output_cotangents
n,w,h,r = (int(x) for x in sys.argv[1:5])
lib = emu.build()
g = sphere_shell_scene(n, r, sh_degree=3, seed=0); cam = orbit_cameras(32, w, h, seed=1)[5]; t = g.tensors()
kw = dict(means3D=t["xyz"], opacities=t["opacity"], scales=t["scaling"], rotations=t["rotation"], shs=t["shs"], uvs=t["uvs"], gradient_uvs=t["grad_uvs"], texture=t["texture"],
          H=h, W=w, tanfovx=math.tan(cam.FoVx/2), tanfovy=math.tan(cam.FoVy/2), bg=(0,0,0), viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, campos=cam.camera_center, sh_degree=3,
          cotangents=output_cotangents(h, w, seed=3))
lib.simt_profile_votes(1)
res = emu.rasterize(**kw)
prof = emu.vote_profile(lib)
lib.simt_profile_votes(0)
print("pairs", res.num_pairs, "visible", res.num_visible)
for k,(calls,hh) in sorted(prof.items(), key=lambda kv: kv[0]):
    tot = sum(i*c for i,c in enumerate(hh))
    print("%-46s calls %10d  mean true lanes %5.2f  none %5.1f%%  1-4 %5.1f%%  5-8 %5.1f%%  9-16 %5.1f%%  17-32 %5.1f%%" % (k, calls, tot/max(calls,1), 100*hh[0]/max(calls,1), 100*sum(hh[1:5])/max(calls,1), 100*sum(hh[5:9])/max(calls,1), 100*sum(hh[9:17])/max(calls,1), 100*sum(hh[17:])/max(calls,1)))
