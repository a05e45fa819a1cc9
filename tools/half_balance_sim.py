"""How many passes of the render kernels' blend loop would a better balance between the two half-warps save?
The shipped kernels walk the list chunk by chunk; per chunk the left and the right 4x4 block have n_L and n_R survivors and
the warp needs max(n_L, n_R) passes. This tool takes the exact (n_L, n_R) sequence of every warp from the CPU emulator
(tests/simt: every __ballot_sync of stream_issue is traced) and replays it under other policies — counts, no GPU:
    lockstep   what is shipped: sum over chunks of max(n_L, n_R)
    window-W   each half walks its own survivor queue, the halves may be at most W-1 chunks apart (a ring of W stages)
    free       no coupling at all: max(sum n_L, sum n_R)     (lower bound of the half-warp design)
    ideal      (sum n_L + sum n_R) / 2                        (both halves always busy)
Usage: python tools/half_balance_sim.py N W H R      e.g. 500000 1920 1080 2048 (headline config, ~3 min)"""
from pathlib import Path
import math
import sys
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import numpy as np
import torch
from simt import emu
from texture_gs_b200.scene import orbit_cameras, sphere_shell_scene, output_cotangents


def window_passes(nl, nr, w):
    """Passes needed when each half consumes its own queue and may run at most w-1 chunks ahead of the other."""
    n = len(nl)
    c = [0, 0]
    rem = [int(nl[0]), int(nr[0])]
    q = (nl, nr)
    passes = 0
    while True:
        for h in (0, 1):        # advance through finished chunks as far as the ring allows
            while rem[h] == 0 and c[h] + 1 < n and c[h] + 1 - min(c[0], c[1]) < w:
                c[h] += 1
                rem[h] = int(q[h][c[h]])
        if rem[0] == 0 and rem[1] == 0:
            o = 0 if c[0] <= c[1] else 1
            if c[o] + 1 >= n:
                return passes
            continue            # the slower half moves on next round (always possible: it is at the minimum)
        passes += 1
        for h in (0, 1):
            if rem[h]:
                rem[h] -= 1


def main():
    n, w, h, r = (int(x) for x in sys.argv[1:5])
    lib = emu.build()
    g = sphere_shell_scene(n, r, sh_degree=3, seed=0)
    cam = orbit_cameras(32, w, h, seed=1)[5]
    t = g.tensors()
    lib.simt_profile_votes(1)
    res = emu.rasterize(means3D=t["xyz"], opacities=t["opacity"], scales=t["scaling"], rotations=t["rotation"], shs=t["shs"], uvs=t["uvs"],
                        gradient_uvs=t["grad_uvs"], texture=t["texture"], H=h, W=w, tanfovx=math.tan(cam.FoVx / 2), tanfovy=math.tan(cam.FoVy / 2),
                        bg=(0, 0, 0), viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, campos=cam.camera_center,
                        sh_degree=3, cotangents=output_cotangents(h, w, seed=3))
    tr = emu.ballot_trace(lib)
    lib.simt_profile_votes(0)
    print("pairs", res.num_pairs, "visible", res.num_visible, "ballots traced", len(tr))
    launches = sorted(set(tr[:, 0].tolist()))
    for name, launch in zip(("render_fwd", "render_bwd"), launches):
        a = tr[tr[:, 0] == launch]
        key = a[:, 1] * 256 + a[:, 2]
        order = np.argsort(key, kind="stable")
        a, key = a[order], key[order]
        bounds = np.flatnonzero(np.diff(key)) + 1
        tot = dict(lockstep=0, window2=0, window3=0, free=0, ideal=0.0)
        chunks = 0
        for seq in np.split(a[:, 3], bounds):
            nl, nr = seq[0::2], seq[1::2]          # stream_issue ballots the left block, then the right one
            assert len(nl) == len(nr)
            chunks += len(nl)
            tot["lockstep"] += int(np.maximum(nl, nr).sum())
            tot["window2"] += window_passes(nl, nr, 2)
            tot["window3"] += window_passes(nl, nr, 3)
            tot["free"] += int(max(nl.sum(), nr.sum()))
            tot["ideal"] += (nl.sum() + nr.sum()) / 2
        base = tot["lockstep"]
        print(f"{name}: {chunks} chunks; passes " + "  ".join(f"{k} {v:.0f} ({100 * v / base:.1f} %)" for k, v in tot.items()))


if __name__ == "__main__":
    main()
