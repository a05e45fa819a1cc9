"""what the fast-math intrinsics of the render kernels (ex2.approx behind __expf) cost in parity.
Renders BASELINE configs[2] (500 k, 1080p, R 2048, one view, forward) with the shipped build and with a -DTEXGS_FAST_EXP=0
build (libm expf), compares both with the float64 C oracle and prints, per output, how many pixels are over the 1e-4
tolerance in each build, how many of those the oracle flags as ill-conditioned, and the set differences.
    python tools/build_variants.py expf:-DTEXGS_FAST_EXP=0 && python tests/gpu_fastexp_diff.py build/variants/libtexgs_expf.so
Each build runs in its own subprocess (the library is loaded once per process)."""
import json, os, subprocess, sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
N, W, H, R, VIEW = 500_000, 1920, 1080, 2048, 5

if len(sys.argv) > 2 and sys.argv[1] == "--render":
    from util import run_cuda
    from texture_gs_b200.scene import orbit_cameras, sphere_shell_scene
    g = sphere_shell_scene(N, R, sh_degree=3, seed=0)
    cam = orbit_cameras(32, W, H, seed=1)[VIEW]
    got, stats, _ = run_cuda(g, cam, bg=(0.1, 0.2, 0.3))
    np.savez(sys.argv[2], image=got[0].numpy(), depth=got[1].numpy(), norm=got[2].numpy(), alpha=got[3].numpy())
    sys.exit(0)

variant = sys.argv[1]
outs = {}
for name, lib in (("shipped (__expf)", None), ("-DTEXGS_FAST_EXP=0 (expf)", variant)):
    env = dict(os.environ)
    if lib:
        env["TEXGS_LIB"] = lib
    f = f"/tmp/fastexp_{'a' if lib is None else 'b'}.npz"
    subprocess.run([sys.executable, __file__, "--render", f], check=True, env=env)
    outs[name] = np.load(f)
from util import run_c_oracle
from texture_gs_b200.scene import orbit_cameras, sphere_shell_scene
g = sphere_shell_scene(N, R, sh_degree=3, seed=0)
cam = orbit_cameras(32, W, H, seed=1)[VIEW]
ref, aux, _ = run_c_oracle(g, cam, bg=(0.1, 0.2, 0.3))
_, aux32, _ = run_c_oracle(g, cam, bg=(0.1, 0.2, 0.3), dtype=torch.float32)
amb = (aux["ambiguous"] | aux32["ambiguous"]).numpy()
rep = {"config": f"{N} Gaussians, {W}x{H}, R {R}, view {VIEW}", "pixels": H * W, "flagged_by_oracle": int(amb.sum()), "outputs": {}}
names = list(outs)
for k, r in zip(("image", "depth", "norm", "alpha"), ref[:4]):
    tol = 3e-4 if k == "depth" else 1e-4
    over = {n: (np.abs(outs[n][k].astype(np.float64) - r.numpy()).max(axis=0) > tol) for n in names}
    a, b = over[names[0]], over[names[1]]
    rep["outputs"][k] = {"tol": tol, "over_shipped": int(a.sum()), "over_expf": int(b.sum()),
                         "over_shipped_unflagged": int((a & ~amb).sum()), "over_expf_unflagged": int((b & ~amb).sum()),
                         "only_shipped": int((a & ~b).sum()), "only_expf": int((b & ~a).sum()),
                         "only_shipped_unflagged": int((a & ~b & ~amb).sum()), "only_expf_unflagged": int((b & ~a & ~amb).sum()),
                         "max_abs_between_builds": float(np.abs(outs[names[0]][k] - outs[names[1]][k]).max())}
print(json.dumps(rep))
