"""Run the REAL kernel source on the CPU (SIMT emulator, tests/simt) against the C oracle at a size of your choice:
    python tools/emu_check.py N W H R [fastmath_noise_ulps] [view] [fwd]
e.g. 10000 256 256 512 (BASELINE configs[0], 45 s) or 56000 640 360 1024 3 5 (a 1/9-scale copy of configs[2], 2-3 min).
Same comparison as tests/test_gpu_zz_fullsize.py makes on the GPU; the emulator runs ~1e5 warp collectives per second."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from simt import emu  # noqa: E402
from texture_gs_b200.scene import orbit_cameras, sphere_shell_scene  # noqa: E402
from util import check_against_c_oracle, run_emu  # noqa: E402

n, w, h, r = (int(x) for x in sys.argv[1:5])
noise = int(sys.argv[5]) if len(sys.argv) > 5 else 0
view = int(sys.argv[6]) if len(sys.argv) > 6 else 5
backward = not (len(sys.argv) > 7 and sys.argv[7] == "fwd")
emu.build().simt_set_fastmath_noise(noise)
g = sphere_shell_scene(n, r, sh_degree=3, seed=0)
cam = orbit_cameras(32, w, h, seed=1)[view]
t0 = time.time()
try:
    check_against_c_oracle(g, cam, bg=(0.1, 0.2, 0.3), runner=run_emu, max_flag=0.7, backward=backward)
    print(f"PASS ({n} splats, {w}x{h}, R={r}, fast-math noise {noise} ulp) in {time.time() - t0:.0f} s")
except AssertionError as e:
    print("FAIL", repr(e)[:1000])
    raise SystemExit(1)
