"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.
usage: python tools/launch_summary.py gpurun_out/launches.csv "title" > profiles/xyz.md"""
import csv, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    us = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    name = re.sub(r"^void ", "", r[ik])
    name = re.sub(r"\(.*$", "", name)
    tot[name] += us
    cnt[name] += 1
total = sum(tot.values())
tex = sum(v for k, v in tot.items() if k.startswith("texgs_"))
print(f"# {sys.argv[2] if len(sys.argv) > 2 else 'ncu launch list'}\n")
print("Cold-cache, serialised per-launch times: compare SHARES, not absolutes. Includes the scene-construction kernels (torch).\n")
print(f"total {total / 1e3:.2f} ms over {sum(cnt.values())} launches; texgs_* kernels {tex / 1e3:.2f} ms = {100 * tex / total:.1f} %\n")
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:16]:
    print(f"| `{k[:80]}` | {cnt[k]} | {v:.1f} | {100 * v / total:.1f} % |")
