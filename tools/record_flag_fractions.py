"""Record the fraction of pixels the oracle flags as fp32-ill-conditioned in every parity test
(tests/golden/flag_fractions.json; the budget tests/util.flag_budget checks against: 1.2 x recorded + 0.5 %).
The fractions depend on the oracle and the seeded scenes only, so no GPU is needed: every test runs up to the point
where its helper has computed the flags, stores them and skips.   python tools/record_flag_fractions.py"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
out = ROOT / "tests" / "golden" / "flag_fractions.json"
tmp = out.with_suffix(".tmp")
if tmp.exists():
    tmp.unlink()
env = dict(os.environ, TEXGS_RECORD_FLAGS=str(tmp), PYTHONPATH=str(ROOT) + os.pathsep + str(ROOT / "tests"))
code = ("import sys, torch, pytest; torch.cuda.is_available = lambda: True; "
        "sys.exit(pytest.main(['tests/test_gpu_parity.py', 'tests/test_gpu_zz_fullsize.py', 'tests/test_simt_kernels_cpu.py', "
        "'-q', '-p', 'no:cacheprovider', '--tb=no', '-x' if False else '-q']))")
subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env)
tmp.replace(out)
print(out.read_text()[:2000])
