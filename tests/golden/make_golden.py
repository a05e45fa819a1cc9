"""Generates tests/golden/tiny_scene.npz.

The reference's own implementation of this path cannot be run (its rasterizer source is an
un-vendored pip-git dependency, see oracle/raster_ref.py header: PARITY UNPINNED), so the golden
vectors are produced by the independent scalar-loop restatement ``oracle/raster_loop.py`` in
float64 on a tiny seeded scene. They pin (a) the vectorised oracle and (b) the CUDA path against
silent drift.  Run from the repo root:  python tests/golden/make_golden.py
"""
import math
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.raster_loop import rasterize_loop  # noqa: E402
from texture_gs_b200.scene import orbit_cameras, sphere_shell_scene  # noqa: E402


def main():
    N, W, H, R = 400, 56, 40, 32
    g = sphere_shell_scene(N, R, sh_degree=3, seed=11, tex_seed=12, coverage=4.0)
    cam = orbit_cameras(1, W, H, seed=13)[0]
    t = {k: (None if v is None else v.detach().double().numpy()) for k, v in g.tensors().items()}
    bg = np.array([0.1, 0.2, 0.3])
    kw = dict(H=H, W=W, tanfovx=math.tan(cam.FoVx / 2), tanfovy=math.tan(cam.FoVy / 2), bg=bg, scale_modifier=1.0,
              viewmatrix=cam.world_view_transform.double().numpy(), projmatrix=cam.full_proj_transform.double().numpy(),
              sh_degree=3, campos=cam.camera_center.double().numpy())
    img, dep, nrm, alp, radii = rasterize_loop(t["xyz"], t["shs"], t["opacity"], t["scaling"], t["rotation"],
                                               t["uvs"], t["grad_uvs"], t["texture"], **kw)
    out = Path(__file__).resolve().parent / "tiny_scene.npz"
    np.savez_compressed(out, N=N, W=W, H=H, R=R, scene_seed=11, tex_seed=12, cam_seed=13, bg=bg,
                        image=img.astype(np.float32), depth=dep.astype(np.float32), norm=nrm.astype(np.float32),
                        alpha=alp.astype(np.float32), radii=radii)
    print("wrote", out, "alpha mean", alp.mean())


if __name__ == "__main__":
    main()
