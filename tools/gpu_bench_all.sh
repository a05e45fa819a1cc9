#!/bin/bash
# bench.py on every workload + the CPU arm + the GPU suite; outputs in gpurun_out/$1_*  (bash tools/gpu_bench_all.sh r2b)
set -u
T=${1:-run}; O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q -s --tb=short -p no:cacheprovider ) > $O/${T}_pytest_gpu.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1
timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
for s in 1 2 4; do
  timeout 400 python bench.py --streams $s --no-cpu-baseline --no-stage-pass > $O/${T}_bench_streams$s.json 2> $O/${T}_bench_streams$s.err
done
timeout 600 python bench.py --workload cfg1_300k_800x600 > $O/${T}_bench_cfg1.json 2> $O/${T}_bench_cfg1.err
timeout 600 python bench.py --workload cfg4_1m_4k > $O/${T}_bench_cfg4.json 2> $O/${T}_bench_cfg4.err
timeout 600 python bench.py --workload cfg0_10k_256 > $O/${T}_bench_cfg0.json 2> $O/${T}_bench_cfg0.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
tail -3 $O/${T}_pytest_gpu.log; tail -1 $O/${T}_smoke.log
for f in $O/${T}_bench.json $O/${T}_bench_streams1.json $O/${T}_bench_streams2.json $O/${T}_bench_streams4.json $O/${T}_bench_cfg1.json $O/${T}_bench_cfg4.json $O/${T}_bench_cfg0.json $O/${T}_bench_ref.json; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print(sys.argv[1], d["metric"], "value %.2f" % d["value"], "e2e", e.get("value"), e.get("error"), "ms/step %.2f" % d["ms_per_step"],
          "clocks", (d.get("clocks") or {}).get("sm_mhz"), (d.get("clocks") or {}).get("reasons"), "roof", (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
done
