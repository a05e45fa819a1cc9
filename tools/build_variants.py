"""Build alternative libtexgs_<name>.so files (same ABI, different -D tuning macros) for A/B timing
on the GPU box:  python tools/build_variants.py name1:-DX=1,-DY=2 name2:...   -> build/variants/"""
import subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from texture_gs_b200 import build as B
out = ROOT / "build" / "variants"
out.mkdir(parents=True, exist_ok=True)
for spec in sys.argv[1:]:
    name, _, flags = spec.partition(":")
    fl = [f for f in flags.split(",") if f]
    dst = out / f"libtexgs_{name}.so"
    cmd = [B.nvcc_path(), *B.NVCC_FLAGS, *fl, "-Xptxas=-v", str(B.SRC), "-o", str(dst)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        print(r.stderr); raise SystemExit(1)
    regs = [l for l in r.stderr.splitlines() if "Used" in l or "Compiling entry" in l or "spill" in l]
    keep = []
    for i, l in enumerate(regs):
        if "render_" in l and "ILi0ELb1" in l:
            keep += regs[i:i + 3]
    print(name, fl, "->", dst.name); print("\n".join(keep))
