#!/usr/bin/env python3
"""Debug aid: weighted pass 2 at C3-like scale, tensor-core pass vs exact kernel vs CPU assembly of dumped windows."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa

import bench  # noqa
from tools import run_configs as rc  # noqa
from garlic_b200.pipeline import interpolate_map  # noqa
from oracle import oracle as orc  # noqa

bench.C_void = ctypes.c_void_p
n_ind, L0, n_ld, W = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
g, rows, names, chr_off0, pos0, cens = rc.setup(n_ind, L0, 3)
cen_arr = np.array([cens["chr" + nm] for nm in names], np.int32)
C = len(names)
map_pos = [pos0[chr_off0[c]:chr_off0[c + 1]][2:-2:5].astype(np.int64) for c in range(C)]
map_cm = [np.round(p * 1.2e-6, 9) for p in map_pos]
chr_param = np.array([[map_pos[c][0], map_pos[c][-1], cen_arr[c][0], cen_arr[c][1]] for c in range(C)], np.int32)
freq, keep, L = g.filter(True, chr_param)
kept = g.get_kept_index()
pos = pos0[kept]
chr_off = np.searchsorted(kept, chr_off0)
gpos = np.empty(L)
for c in range(C):
    gpos[chr_off[c]:chr_off[c + 1]], _ = interpolate_map(pos[chr_off[c]:chr_off[c + 1]], map_pos[c], map_cm[c])
g.set_tables(0.001, 200000, cen_arr, gpos)
g.set_wlod(1e-9, 7)
ld_ind = np.sort(np.random.default_rng(3).choice(n_ind, n_ld, replace=False)).astype(np.int32)
for _ in range(int(os.environ.get("REPS_LD", "1"))):
    g.ld_band(W, ld_ind)
    print("after ld_band:", len(g.call_roh(W, 1.0, 0.25, weighted=True, exact=False)))
for _ in range(int(os.environ.get("REPS_ROH", "1"))):
    print("repeat call_roh:", len(g.call_roh(W, 1.0, 0.25, weighted=True, exact=False)))
a = g.call_roh(W, 1.0, 0.25, weighted=True, exact=False)
sa = g.last_stats()
b = g.call_roh(W, 1.0, 0.25, weighted=True, exact=True)
sb = g.last_stats()
print("mma", len(a), sa, "\nexact", len(b), sb)
print("equal", np.array_equal(a, b))
if not np.array_equal(a, b):
    sa_ = set(map(tuple, a.tolist())); sb_ = set(map(tuple, b.tolist()))
    da = sorted(sa_ - sb_); db = sorted(sb_ - sa_)
    print("only mma", len(da), da[:10]); print("only exact", len(db), db[:10])
    ind = (da + db)[0][0]
    win = g.windows(W, 1, weighted=True, individuals=np.array([ind], np.int32))[0]
    want = []
    for c in range(C):
        lo, hi = chr_off[c], chr_off[c + 1]
        s, e, _ = orc.assemble(win[lo:hi], pos[lo:hi], gpos[lo:hi], 1.0, W, 200000, 0.25, True, tuple(cen_arr[c]))
        # assemble returns positions; map back to indices
        for x, y in zip(s, e):
            want.append((ind, c, int(lo + np.searchsorted(pos[lo:hi], x)), int(lo + np.searchsorted(pos[lo:hi], y))))
    print("cpu   ", [t for t in want][:12])
    print("mma   ", [t for t in a.tolist() if t[0] == ind][:12])
    print("exact ", [t for t in b.tolist() if t[0] == ind][:12])
