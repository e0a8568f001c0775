#!/usr/bin/env python3
"""Repeatability stress: the same call_roh / filter sequence many times in one process must return identical
ROH (catches races between stream-ordered phases).  python tools/stress_repeat.py [reps]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, ctypes
import bench
from garlic_b200 import synth
from garlic_b200.api import GarlicGPU

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda", 0)

def run(n_ind, L0, W, use_gl, cutoff):
    names, chr_off0, pos0, cens = synth.make_positions_genomewide(7, L0)
    row_bytes = ((L0 + 3) // 4 + 15) // 16 * 16
    rows = bench.make_rows_torch(torch, dev, n_ind, L0, 7, 1000, row_bytes)
    torch.cuda.synchronize()                     # the library copies on its own stream
    g = GarlicGPU(0)
    g.set_shape(n_ind, L0, chr_off0, pos0)
    g.put_packed_dev(rows.data_ptr(), row_bytes)
    if use_gl:
        gen = torch.Generator(device=dev); gen.manual_seed(44)
        vals = torch.tensor([0.0004, 0.004, 0.04, 0.5, 3.0, 0.0, 150.0], dtype=torch.float64, device=dev)
        idx = torch.randint(0, 5, (n_ind, L0), generator=gen, device=dev)
        gl = vals[idx]
        print("gl checksum", float(gl.sum()))
        g._ck(g.lib.garlic_gpu_put_gl_dev(g.h, ctypes.c_void_p(gl.data_ptr()), 2))
    cen_arr = np.array([cens["chr" + nm] for nm in names], np.int32)
    outs = []
    for r in range(reps):
        g.count_packed()
        freq, keep, L = g.filter()
        g.set_tables(None if use_gl else 0.001, 200000, cen_arr)
        roh = g.call_roh(W, cutoff, 0.25)
        outs.append(roh.copy())
    ex = g.call_roh(W, cutoff, 0.25, exact=True)
    same = [np.array_equal(o, outs[0]) for o in outs]
    print("use_gl=%s n=%d L=%d W=%d: %d ROH, all %d repetitions identical: %s, equals exact chains: %s" % (
        use_gl, n_ind, L0, W, len(outs[0]), reps, all(same), np.array_equal(outs[0], ex)), [len(o) for o in outs][:8])
    g.close()
    return all(same) and np.array_equal(outs[0], ex)

ok = run(2000, 300_000, 50, False, 2.0)
ok &= run(500, 400_000, 200, True, 5.0)
ok &= run(100, 200_000, 64, False, 1.0)
sys.exit(0 if ok else 1)
