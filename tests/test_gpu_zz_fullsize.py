"""BASELINE.json configs[2] at FULL size (500 k Gaussians, 1920x1080, 6x2048^2 texture), forward and backward, compared
DIRECTLY with the oracle: the scalar C + OpenMP oracle (oracle/raster_c.c, equal to the torch oracle to 1e-15, see
tests/test_oracle_c.py) renders a whole 1080p view in seconds on the box's host cores. The property-based full-size test
(test_gpu_parity.py::test_full_size_properties) stays; this one is the bit the properties cannot see.
The comparison logic itself is exercised on CPU with the emulated kernels (tests/test_simt_kernels_cpu.py). Runs last."""
import pytest
import torch

from util import check_against_c_oracle
from texture_gs_b200.scene import orbit_cameras, sphere_shell_scene

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,w,h,r,view", [(500_000, 1920, 1080, 2048, 5)])
def test_full_size_forward_and_backward_against_the_c_oracle(n, w, h, r, view):
    if not torch.cuda.is_available():
        pytest.fail("GPU tests selected (-m gpu) but no CUDA device is visible")
    g = sphere_shell_scene(n, r, sh_degree=3, seed=0)
    cam = orbit_cameras(32, w, h, seed=1)[view]
    # R = 2048: a tap lands within TEXEL_TIE * R = 4e-3 texels of a texel boundary in a good part of the pixels: the oracle
    # flags 8.9 % of the pixels for values and 29.9 % for gradients (tests/golden/flag_fractions.json); caps = 1.2 x that
    check_against_c_oracle(g, cam, bg=(0.1, 0.2, 0.3), max_flag=0.36, max_amb=0.11)


@pytest.mark.parametrize("name,n,w,h,r", [("configs[1] stand-in", 300_000, 800, 600, 1024), ("configs[4]", 1_000_000, 3840, 2160, 4096)])
def test_forward_only_baseline_configs_against_the_c_oracle(name, n, w, h, r):
    """BASELINE.json configs[1] (300 k Gaussians, 800x600, R = 1024: the retexture.py shape; synthetic stand-in for the
    DTU checkpoint) and configs[4] (1 M Gaussians, 3840x2160, R = 4096, the bandwidth stress case), forward, every pixel
    of the four outputs against the C oracle. At 3840 pixels across, fp32 pixel coordinates resolve 2.4e-4 px: the float32
    build of the oracle itself is off by more than 1e-4 on ~0.1 % of the unflagged pixels there (none at 800x600), which is
    what the "no worse than 5x the float32 oracle" clause of the helper is for."""
    if not torch.cuda.is_available():
        pytest.fail("GPU tests selected (-m gpu) but no CUDA device is visible")
    g = sphere_shell_scene(n, r, sh_degree=3, seed=0)
    cam = orbit_cameras(32, w, h, seed=1)[3]
    # observed: 13.0 % / 31.1 % (configs[1]) and 8.0 % / 38.2 % (configs[4]) of the pixels flagged for values / gradients
    check_against_c_oracle(g, cam, bg=(0.0, 0.0, 0.0), max_flag=0.46, max_amb=0.16, backward=False)
