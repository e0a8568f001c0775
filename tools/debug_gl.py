#!/usr/bin/env python3
"""Debug aid: GL pass 2, relay kernel (GARLIC_GL_WARPS=K) against the single-warp ring kernel."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from garlic_b200 import synth
from garlic_b200.pipeline import HotPath

W = int(sys.argv[1]); n_ind = int(sys.argv[2]); L0 = int(sys.argv[3])
names, offs, pos, cens = synth.make_positions_genomewide(31, L0, n_chr=3)
codes = synth.make_codes(31, n_ind, L0)
class DS: pass
ds = DS()
ds.chr_names, ds.chr_offsets, ds.pos, ds.centromeres = names, offs, pos, cens
rng = np.random.default_rng(5)
ds.gl = rng.choice(np.array([0.0004, 0.004, 0.04, 0.5, 3.0, 0.0, 150.0]), size=(L0, n_ind))
ds.gl_type = "PL"
hp = HotPath().load(ds, error=None, packed_rows=synth.pack_codes(codes))
g = hp.g
os.environ["GARLIC_GL_WARPS"] = "1"
ref = g.call_roh(W, 1.0, 0.25).copy()
for K in (2, 3, 4):
    os.environ["GARLIC_GL_WARPS"] = str(K)
    for rep in range(3):
        a = g.call_roh(W, 1.0, 0.25).copy()
        if np.array_equal(a, ref):
            print("K", K, "rep", rep, "equal", len(a))
        else:
            sa = set(map(tuple, a.tolist())); sr = set(map(tuple, ref.tolist()))
            print("K", K, "rep", rep, "DIFF only_relay", sorted(sa - sr)[:6], "only_ref", sorted(sr - sa)[:6], len(a), len(ref))
