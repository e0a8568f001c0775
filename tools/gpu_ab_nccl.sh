#!/bin/bash
# A/B of the counter exchange: NVLink kernel fused with freq + keep (xchg.cu) against ncclAllReduce + freq_keep_kernel, at N GPUs, host-side laps on.
N=${1:-2}; tag=${2:-r02h}
out=gpurun_out; mkdir -p $out
for v in xchg nccl; do
  if [ $v = nccl ]; then export GARLIC_NO_XCHG=1; else unset GARLIC_NO_XCHG; fi
  GARLIC_TIMING=1 NCCL_DEBUG=${NCCL_DEBUG_LEVEL:-WARN} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 20 --warmup 5 --cpu-sample-multi 24 > $out/${tag}_n${N}_${v}.json 2> $out/${tag}_n${N}_${v}.err
  echo "$v rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$out/${tag}_n${N}_${v}.json").read().strip().splitlines()[-1])
    print("$v", d["ms_per_step"], d["value"], d["phases_ms_one_synchronised_step"], d.get("parity_vs_cpu_sample"), d.get("parity_freq_vs_host_counts"))
except Exception as e: print("no json", e)
PY
  grep -E "r0\] (filter|call_roh:.*items)" $out/${tag}_n${N}_${v}.err | tail -4; grep -E "r5\] filter" $out/${tag}_n${N}_${v}.err | tail -2
done
