#!/bin/bash
# Final evidence run of the round: GPU test-suite, smoke, bench (driver defaults), launch list and one --set full
# capture per texgs kernel at the headline config. Outputs in gpurun_out/.
set -u
O=gpurun_out; mkdir -p $O
( time timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > $O/final_pytest_gpu.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1
timeout 900 python bench.py > $O/final_bench.json 2> $O/final_bench.err
timeout 300 python tests/gpu_trainstep.py > $O/final_trainstep.log 2>&1
timeout 300 python tests/gpu_configs.py > $O/final_configs.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/final_launches.csv \
    python bench.py --views 4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/final_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:texgs_ -s 9 -c 8 -f -o $O/prof_all_r1_final2 \
    python tests/gpu_step.py 500000 1920 1080 2048 2 > $O/final_ncu_all.log 2>&1
tail -3 $O/final_pytest_gpu.log; tail -1 $O/final_smoke.log; cut -c1-300 $O/final_bench.json; tail -4 $O/final_trainstep.log; tail -3 $O/final_ncu_all.log
