#!/usr/bin/env python3
"""Full-precision (raw fp64) window dumps of the REFERENCE's own calcLODWindows / calcwLODWindows for a few
golden cases → tests/golden/<case>/refwin.npz.  Needs oracle/_ref/ref_driver (make -C oracle ref, which
compiles /root/reference/src/*.cpp where they lie).  Run from the repo root:
    python tests/golden/make_refwin.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc, refdrv  # noqa: E402
from tests.common import arg, load_case, GOLDEN  # noqa: E402

for name in ["lod_small", "gl_pl", "wlod_cm"]:
    ds, args = load_case(name)
    W = arg(args, "--winsize", cast=int)
    err = arg(args, "--error", None, float)
    weighted = "--weighted" in args
    # the restatement is used only to code/filter the inputs the way the reference's loader does;
    # the windows stored are the reference functions' own
    res = orc.run_pipeline(ds, W, err, None, weighted=weighted, cm="--cm" in args)
    out = refdrv.run(res["chroms"], ds.n_ind, W, err, weighted=weighted, cm="--cm" in args, do_roh=False, threads=2)
    np.savez_compressed(os.path.join(GOLDEN, name, "refwin.npz"), **{"chr%d" % c: w for c, w in enumerate(out["win"])})
    print(name, [w.shape for w in out["win"]])
