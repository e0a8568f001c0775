#!/bin/bash
# A/B of the counter all-reduce with and without the symmetric-window registration at N GPUs, host-side laps on.
N=${1:-2}; tag=${2:-r02h}
out=gpurun_out; mkdir -p $out
for v in win nowin; do
  if [ $v = nowin ]; then export GARLIC_NCCL_NO_WINDOW=1; else unset GARLIC_NCCL_NO_WINDOW; fi
  GARLIC_TIMING=1 NCCL_DEBUG=${NCCL_DEBUG_LEVEL:-WARN} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > $out/${tag}_n${N}_${v}.json 2> $out/${tag}_n${N}_${v}.err
  echo "$v rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$out/${tag}_n${N}_${v}.json").read().strip().splitlines()[-1])
    print("$v", d["ms_per_step"], d["value"], d["phases_ms_one_synchronised_step"])
except Exception as e: print("no json", e)
PY
  grep -E "r0\] (filter|call_roh:.*items)" $out/${tag}_n${N}_${v}.err | tail -4; grep -E "r5\] filter" $out/${tag}_n${N}_${v}.err | tail -2
done
