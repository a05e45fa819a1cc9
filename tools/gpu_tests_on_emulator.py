"""Run the ``-m gpu`` parity tests that reach the kernels through ``tests/util.run_cuda`` with the SIMT emulator
(tests/simt) standing in for CUDA — a pre-flight for changes to the oracle, its conditioning flags or the test helpers
on a box without a GPU:
    python tools/gpu_tests_on_emulator.py            # ~3 min: 12 tests
    python tools/gpu_tests_on_emulator.py config0    # + BASELINE configs[0] forward/backward (several minutes)
    python tools/gpu_tests_on_emulator.py -DTEXGS_CHUNK=16     # the same tests on a tuning-macro variant of the kernels
Tests that build CUDA tensors themselves (operator-level tests, fused buckets, dual render, ...) cannot be redirected
and are not selected. A pass here says nothing about PTX semantics or performance; it says the comparison code and the
oracle's flags still accept kernels whose arithmetic is the shipped source."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import pytest  # noqa: E402
import torch  # noqa: E402
import util  # noqa: E402

FLAGS = tuple(a for a in sys.argv[1:] if a.startswith("-D"))
if FLAGS:
    from simt import emu  # noqa: E402
    _variant = emu.build(extra_flags=FLAGS)

    def _run_variant(*a, **kw):
        return util.run_emu(*a, lib=_variant, **kw)
    util.run_cuda = _run_variant
else:
    util.run_cuda = util.run_emu
torch.cuda.is_available = lambda: True
SEL = ("test_golden_tiny_scene or test_forward_parity_small or test_backward_parity_small or test_clamp_paths or test_saturating "
       "or test_camera_inside or test_heterogeneous or test_long_tile_lists or test_forward_is_deterministic or test_spec_switch_inst "
       "or test_upstream_clamp")
if "config0" in sys.argv[1:]:
    SEL += " or test_config0_10k_256_forward_and_backward"
sys.exit(pytest.main([str(ROOT / "tests" / "test_gpu_parity.py"), "-m", "gpu", "-q", "--tb=short", "-k", SEL, "-p", "no:cacheprovider",
                      "--durations=8"]))
