#!/bin/bash
# Scaling evidence on one 8-GPU box: bash tools/gpu_scale.sh <tag>   (gpurun --gpus 8). Outputs gpurun_out/<tag>_*.json
set -u
T=${1:-scale}; O=gpurun_out; mkdir -p $O
run() {  # n, name, extra flags...
  n=$1; name=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline "$@" > $O/${T}_${name}.json 2> $O/${T}_${name}.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline "$@" > $O/${T}_${name}.json 2> $O/${T}_${name}.err
  fi
}
run 1 n1
run 8 n8
# run 8 n8_peer --no-multicast   (r2s: 2900.7 vs 2949.2 views/s with multimem)
run 8 n8_allreduce --no-optimizer
run 4 n4
run 2 n2
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29977 \
   bench.py --gpus 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-stage-pass --no-optimizer 2>&1 | grep -iE "NVLS|algo|Connected|channels" | sort | uniq -c | sort -rn | head -12 > $O/${T}_nccl_info.txt
nvidia-smi topo -m > $O/${T}_topo.txt 2>&1
for f in n1 n8 n8_allreduce n4 n2; do
  python - "$O/${T}_$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print(sys.argv[1], "value %.1f" % d["value"], "e2e", e.get("value"), "ms/step %.2f" % d["ms_per_step"], d["impl_notes"]["parallelism"][:90], d["grad_checksum"])
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
done
cat $O/${T}_nccl_info.txt
