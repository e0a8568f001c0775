#!/bin/bash
# Multi-GPU session (gpurun --gpus N): GPU tests incl. the 2-GPU ones, then the bench at N under torchrun.
# Usage: bash tools/gpu_multi.sh <tag> <N> [tests|notests]
tag=${1:-r02x}; N=${2:-2}; tests=${3:-tests}
out=gpurun_out; mkdir -p $out
if [ "$tests" = tests ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_gputests_${N}gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_gputests_${N}gpu.log
  tail -4 $out/${tag}_gputests_${N}gpu.log
fi
for n in $(echo $N | tr ',' ' '); do
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --steps 20 --warmup 5 --no-configs > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > $out/${tag}_bench_n${n}.json 2> $out/${tag}_bench_n${n}.err
  fi
  echo "bench n=$n rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$out/${tag}_bench_n${n}.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","parity_vs_cpu_sample","parity_freq_vs_host_counts","phases_ms_one_synchronised_step")}, d["e2e"])
except Exception as e: print("no json", e)
PY
done
