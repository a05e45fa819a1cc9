#!/bin/bash
# First GPU call of round 2 (one `gpurun --timeout 1500 -- bash tools/gpu_round2.sh`): everything written after the
# round-1 GPU budget was spent gets its first measurement. Outputs in gpurun_out/r2_*.
#   1. GPU test-suite + smoke (the multi-stream test runs last)
#   2. bench with the driver defaults (reference: 373 views/s in round 1)
#   3. bench with 2 and 3 streams per rank (experimental view pipeline), with the experimental
#      two-splats-per-iteration forward kernel (--fwd-ilp2), and with both
#   4. launch list of a 4-view step with 2 streams (do the small kernels really overlap the render kernels?)
#   5. A/B of the half-window build (-DTEXGS_HALF_WINDOW=1: the half-warps walk their queues independently, 7.8 % fewer
#      passes of the blend loop on the emulator): per-stage times of both builds, and the GPU parity tests on the variant
set -u
O=gpurun_out; mkdir -p $O
( time timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > $O/r2_pytest_gpu.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1
timeout 600 python bench.py > $O/r2_bench.json 2> $O/r2_bench.err
for s in 2 3; do
  timeout 400 python bench.py --streams $s --no-cpu-baseline > $O/r2_bench_streams$s.json 2> $O/r2_bench_streams$s.err
done
timeout 400 python bench.py --fwd-ilp2 --no-cpu-baseline > $O/r2_bench_ilp2.json 2> $O/r2_bench_ilp2.err
timeout 400 python bench.py --fwd-ilp2 --streams 2 --no-cpu-baseline > $O/r2_bench_ilp2_streams2.json 2> $O/r2_bench_ilp2_streams2.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_streams2.csv \
    python bench.py --streams 2 --views 4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2_launches_streams2.log 2>&1
timeout 600 python tools/build_variants.py default: halfwin:-DTEXGS_HALF_WINDOW=1 > $O/r2_variants_build.log 2>&1
for v in default halfwin; do
  TEXGS_LIB=build/variants/libtexgs_$v.so timeout 300 python tests/gpu_variants.py fused > $O/r2_variant_$v.json 2> $O/r2_variant_$v.err
done
( TEXGS_LIB=build/variants/libtexgs_halfwin.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider ) > $O/r2_pytest_halfwin.log 2>&1
tail -3 $O/r2_pytest_gpu.log; tail -1 $O/r2_smoke.log; tail -2 $O/r2_pytest_halfwin.log; cat $O/r2_variant_default.json $O/r2_variant_halfwin.json
for f in $O/r2_bench.json $O/r2_bench_streams2.json $O/r2_bench_streams3.json $O/r2_bench_ilp2.json $O/r2_bench_ilp2_streams2.json; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.1f" % d["value"], "e2e %.1f" % (d["e2e"] or {}).get("value", float("nan")), "ms/step %.2f" % d["ms_per_step"],
          "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
