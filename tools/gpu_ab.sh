#!/bin/bash
# A/B timing of build variants (tools/build_variants.py name:-DFLAG=1 ...) on the headline config: bash tools/gpu_ab.sh name1 name2 ...  (default build first)
O=gpurun_out; mkdir -p $O
timeout 200 python tests/gpu_variants.py fused > $O/ab_default.json 2> $O/ab_default.err
for n in "$@"; do
  TEXGS_LIB=build/variants/libtexgs_$n.so timeout 200 python tests/gpu_variants.py fused > $O/ab_$n.json 2> $O/ab_$n.err
done
cat $O/ab_default.json; for n in "$@"; do cat $O/ab_$n.json; done
