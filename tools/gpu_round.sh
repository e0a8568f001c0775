#!/bin/bash
# One GPU-box session: GPU tests, the bench (both arms), an ncu launch list of the bench step and one --set full capture
# of the dominant kernels.  Usage (through gpurun): bash tools/gpu_round.sh <tag> [tests|notests] [full|nofull]
tag=${1:-r02x}; tests=${2:-tests}; full=${3:-full}
out=gpurun_out; mkdir -p $out
if [ "$tests" = tests ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_gputests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_gputests.log
  tail -3 $out/${tag}_gputests.log
fi
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_own.json 2> $out/${tag}_bench_own.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-configs > $out/${tag}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
if [ "$full" = full ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"squeeze_bound_kernel|walk_units_kernel|select_kernel|count_packed_kernel|thin_windows_kernel|bound_tables_kernel" -s 12 -c 6 \
    -o $out/${tag}_pass2 -f python bench.py --steps 2 --warmup 1 --no-cpu --no-configs > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
head -c 3000 $out/${tag}_bench_own.json
