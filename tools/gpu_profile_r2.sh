#!/bin/bash
# Round-2 evidence run: launch list + one --set full capture per texgs kernel at the headline config, the fast-exp parity
# diff and the texel-footprint measurement. Outputs in gpurun_out/r2p_*.
set -u
O=gpurun_out; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2p_launches.csv \
    python bench.py --views 4 --steps 1 --warmup 1 --streams 1 --no-e2e --no-cpu-baseline --no-stage-pass > $O/r2p_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:texgs_ -s 9 -c 8 -f -o $O/prof_all_r2 \
    python tests/gpu_step.py 500000 1920 1080 2048 2 > $O/r2p_ncu_all.log 2>&1
timeout 600 python tests/gpu_fastexp_diff.py build/variants/libtexgs_expf.so > $O/r2p_fastexp.json 2> $O/r2p_fastexp.err
( TEXGS_LIB=build/variants/libtexgs_expf.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_fullsize.py -m gpu -q --tb=short -p no:cacheprovider ) > $O/r2p_pytest_expf.log 2>&1
for w in cfg2_500k_1080p cfg1_300k_800x600 cfg4_1m_4k; do
  timeout 300 python tools/texel_footprint.py $w 96 > $O/r2p_footprint_$w.json 2> $O/r2p_footprint_$w.err
done
tail -2 $O/r2p_ncu_all.log; cat $O/r2p_fastexp.json; tail -2 $O/r2p_pytest_expf.log; cat $O/r2p_footprint_*.json; tail -3 $O/r2p_footprint_cfg2_500k_1080p.err
