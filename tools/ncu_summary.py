#!/usr/bin/env python3
"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (tracked).

    python tools/ncu_summary.py launches gpurun_out/launches_r01.csv profiles/r01_launches.txt
    python tools/ncu_summary.py raw gpurun_out/prof_walk_r01.ncu-rep profiles/r01_walk_kernel_ncu.txt
"""
import collections
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput", "gpu__dram_throughput",
        "sm__throughput", "sm__warps_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit", "launch__shared_mem", "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate",
        "smsp__issue_active", "sm__inst_executed_pipe_fp64", "sm__pipe_fp64_cycles_active", "sm__inst_executed.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__pipe_tensor",
        "l1tex__data_pipe_lsu_wavefronts.sum", "smsp__average_warp", "smsp__warp_issue_stalled", "sm__inst_executed_pipe",
        "l1tex__t_bytes", "lts__t_bytes", "smsp__thread_inst_executed_per_inst_executed", "sm__sass_inst_executed_op"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for x in csv.DictReader(lines):
        k = x["Kernel Name"]
        v = float(x["Metric Value"].replace(",", ""))
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n")
        f.write("# source: %s\n%-90s %5s %12s %12s %7s\n" % (src, "kernel", "n", "total_ms", "avg_ms", "share"))
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-90s %5d %12.3f %12.3f %6.1f%%\n" % (k[:90], n, t / 1e6, t / 1e6 / n, 100 * t / tot))


def raw(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on; selected metrics per captured launch\n# source: %s\n" % src)
        ki = hdr.index("Kernel Name")
        for r in data:
            f.write("\n== %s\n" % r[ki])
            for i, h in enumerate(hdr):
                if any(h.startswith(k) for k in KEEP):
                    f.write("%-90s %18s %s\n" % (h, r[i], units[i]))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
