#!/bin/bash
# 8-GPU session: the headline bench at N = 8 (and optionally N = 4), then the C5 shape (100,000 individuals x 600 k SNPs).
tag=${1:-r02n}; out=gpurun_out; mkdir -p $out
run() {  # name, nproc, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $2 --steps 20 --warmup 5 $3 > $out/${tag}_$1.json 2> $out/${tag}_$1.err
  echo "$1 rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$out/${tag}_$1.json").read().strip().splitlines()[-1])
    print("$1", d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], d["clocks"], d.get("parity_vs_cpu_sample"), d.get("parity_freq_vs_host_counts"))
except Exception as e: print("no json", e)
PY
}
run bench_n8 8 "--cpu-sample-multi 24"
[ "$2" = n4 ] && run bench_n4 4 "--cpu-sample-multi 24"
run bench_c5_8gpu 8 "--n-ind 12500 --steps 10 --cpu-sample-multi 8"
