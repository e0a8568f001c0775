#!/usr/bin/env python3
"""Times the compaction (K3) / pruning-bound kernels of squeeze.cu in their three forms on the C2 shape:
fused (compaction + bound), compaction only (GARLIC_NO_PRUNE=1), bound only (a second window size on compacted rows).
CUDA events around each launch (garlic_gpu_last_stats[7]).  Usage: python tools/time_squeeze.py [n_ind] [n_loci]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from garlic_b200 import synth  # noqa: E402
from garlic_b200.api import GarlicGPU  # noqa: E402

n_ind = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
L0 = int(sys.argv[2]) if len(sys.argv) > 2 else 600_000
names, chr_off0, pos0, cens = synth.make_positions_genomewide(2, L0)
cen = np.array([cens["chr" + n] for n in names], np.int32)
row_bytes = ((L0 + 3) // 4 + 15) // 16 * 16
dev = torch.device("cuda", 0)
rows = bench.make_rows_torch(torch, dev, n_ind, L0, 2, 1000, row_bytes)


def run(label, env, W, W2=None):
    for k, v in env.items():
        os.environ[k] = v
    g = GarlicGPU(0)
    for k in env:
        os.environ.pop(k)
    g.set_shape(n_ind, L0, chr_off0, pos0)
    g.put_packed_dev(rows.data_ptr(), row_bytes)
    ts = []
    for _ in range(6):
        g.count_packed()
        g.filter()
        g.set_tables(0.001, 200000, cen)
        g.windows(W, W, individuals=np.arange(20, dtype=np.int32), exact=False)
        t = g.last_stats()["squeeze_ms"]
        if W2:
            g.piece_bounds(W2)
            t = g.last_stats()["squeeze_ms"]
        ts.append(t)
    print("%-40s %.4f ms (min of %s)" % (label, min(ts[1:]), ["%.3f" % x for x in ts[1:]]))
    g.close()


run("fused compaction + bound (W=50)", {}, 50)
run("compaction only", {"GARLIC_NO_PRUNE": "1"}, 50)
run("bound only (W=100 on compacted rows)", {}, 50, 100)
