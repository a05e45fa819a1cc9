#!/bin/bash
# One gpurun call of round 1 (re-entry session): full GPU test-suite, bench, the staged preprocess-backward A/B
# (timing + parity + memcheck), N1 timing and one ncu capture of its kernel. Everything lands in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
V=build/variants/libtexgs_stagesh.so
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=8 ) > $O/pytest_gpu.log 2>&1
timeout 400 python bench.py --steps 4 --warmup 3 > $O/bench.json 2> $O/bench.err
timeout 200 python tests/gpu_variants.py fused > $O/variant_default.json 2> $O/variant_default.err
TEXGS_LIB=$V timeout 200 python tests/gpu_variants.py fused > $O/variant_stagesh.json 2> $O/variant_stagesh.err
( TEXGS_LIB=$V timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "backward_parity_small or fused_bucket or plain_3dgs or extra_attrs" ) > $O/pytest_stagesh.log 2>&1
timeout 200 python tests/gpu_uvnet_time.py > $O/uvnet_time.json 2> $O/uvnet_time.err
( TEXGS_LIB=$V timeout 400 compute-sanitizer --tool memcheck python tests/gpu_sanitize.py ) > $O/sanitize_stagesh.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:texgs_uvmlp_fwd -c 1 -s 3 -f -o $O/uvmlp_r1 python tests/gpu_uvnet_time.py > $O/ncu_uvmlp.log 2>&1
tail -4 $O/pytest_gpu.log; tail -2 $O/pytest_stagesh.log; tail -1 $O/smoke.log; cut -c1-400 $O/bench.json; cat $O/variant_*.json; cat $O/uvnet_time.json; tail -3 $O/sanitize_stagesh.log
