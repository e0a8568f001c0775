#!/usr/bin/env python3
"""Time tped ingest of the C++ driver (garlic_b200 --freq-only: load → K0/K1/K2 → .freq.gz) with the GPU tokeniser (K0)
against host extraction of the allele characters (--host-tokenize), plain text and gz.

    python tools/ingest_bench.py [n_ind] [n_snps]
"""
import gzip
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "garlic_b200", "host", "garlic_b200")
n_ind = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
rng = np.random.default_rng(1)
with tempfile.TemporaryDirectory() as tmp:
    tped, tfam = os.path.join(tmp, "s.tped"), os.path.join(tmp, "s.tfam")
    p = rng.uniform(0.05, 0.95, L)
    with open(tped, "wb") as f:
        for l in range(L):
            a = np.where(rng.random(2 * n_ind) < p[l], ord("A"), ord("G")).astype(np.uint8)
            row = np.empty(4 * n_ind, np.uint8)
            row[0::2] = ord(" ")
            row[1::2] = a
            f.write(b"1 rs%d 0 %d" % (l, 1000 + 500 * l) + row.tobytes() + b"\n")
    with open(tfam, "w") as f:
        for i in range(n_ind):
            f.write("POP ind%d 0 0 0 0\n" % i)
    with open(tped, "rb") as fi, gzip.open(tped + ".gz", "wb", compresslevel=1) as fo:
        fo.write(fi.read())
    res = dict(n_ind=n_ind, n_snps=L, text_mb=os.path.getsize(tped) / 1e6)
    for name, path in (("plain", tped), ("gz", tped + ".gz")):
        for mode, extra in (("k0_gpu_tokeniser", []), ("host_tokenize", ["--host-tokenize"])):
            best = 1e9
            for _ in range(2):
                t0 = time.perf_counter()
                r = subprocess.run([BIN, "--tped", path, "--tfam", tfam, "--freq-only", "--build", "hg19", "--winsize", "50", "--error", "0.001", "--out", os.path.join(tmp, "o")] + extra,
                                   capture_output=True, text=True)
                best = min(best, time.perf_counter() - t0)
                assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
            res["%s_%s_s" % (name, mode)] = round(best, 3)
    print(json.dumps(res))
