"""Is the device code of the in-tree libtexgs.so the one that was measured? Builds the CUDA sources of a given commit in a
scratch directory with the same nvcc flags and compares the SASS of every kernel with the current in-tree build:
    python tools/sass_diff.py 4b8052d        (the last commit of round 1 that ran on a GPU)"""
import hashlib, re, subprocess, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from texture_gs_b200 import build as B


def kernels(lib):
    out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    d, name = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = m.group(1)
            d[name] = hashlib.sha1()
        elif name:
            d[name].update(ln.encode())
    return {k: v.hexdigest() for k, v in d.items()}


commit = sys.argv[1]
with tempfile.TemporaryDirectory() as tmp:
    tar = subprocess.run(["git", "-C", str(ROOT), "archive", commit, "texture_gs_b200/csrc", "include"], capture_output=True, check=True).stdout
    subprocess.run(["tar", "-x", "-C", tmp], input=tar, check=True)
    old = Path(tmp) / "libtexgs_old.so"
    flags = [f.replace(str(ROOT), tmp) for f in B.NVCC_FLAGS]
    subprocess.run([B.nvcc_path(), *flags, f"{tmp}/texture_gs_b200/csrc/texgs_api.cu", "-o", str(old)], check=True, capture_output=True)
    a, b = kernels(old), kernels(B.SRC.parent.parent / "libtexgs.so")
changed = sorted(k for k in a if k in b and a[k] != b[k])
print(f"{commit}: {len(a)} kernels, in-tree build: {len(b)}; identical SASS: {sum(a[k] == b.get(k) for k in a)}; changed: {len(changed)}; "
      f"removed: {len([k for k in a if k not in b])}; new: {len([k for k in b if k not in a])}")
for k in changed:
    print("  changed:", k)
for k in sorted(k for k in b if k not in a):
    print("  new:", k)
sys.exit(1 if changed else 0)
