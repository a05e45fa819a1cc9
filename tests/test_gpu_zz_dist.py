"""The fused multi-GPU texture step (texture_gs_b200.dist.DistTextureAdam: gradient pull over NVLink / NVSwitch multicast + Adam
on the owned shard + texel push, one kernel per rank) against NCCL all-reduce + the single-GPU TextureAdam. Needs two visible
GPUs (one process per GPU through torch.distributed.run); on a one-GPU box the test is skipped — tests/gpu_dist_adam.py is the
same check as a script (run on 2 and 8 GPUs this round: textures equal to 2.4e-7 / 9.5e-7)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_fused_data_parallel_texture_step_equals_allreduce_plus_adam():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests selected (-m gpu) but no CUDA device is visible")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (one process per GPU)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29617", str(ROOT / "tests" / "gpu_dist_adam.py"), "64", "2"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and lines, (r.stdout[-1500:], r.stderr[-1500:])
    rep = json.loads(lines[-1])
    assert rep["ok"] and "peer" in rep["paths"]
