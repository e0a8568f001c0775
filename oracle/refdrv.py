"""Runs oracle/_ref/ref_driver — the REFERENCE's own calcLODWindows / calcHR2LD / calcwLODWindows /
assembleROHWindows compiled from /root/reference/src (see ref_driver.cpp) — on arrays.

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's cpu_baseline and --impl reference legs).
"""
from __future__ import annotations

import os
import struct
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(_HERE, "_ref", "ref_driver")


def available():
    return os.path.exists(BIN) and os.access(BIN, os.X_OK)


def write_input(path, chroms, N, W, error, cutoff=0.0, overlap_frac=0.25, max_gap=200000, weighted=False,
                cm=False, mu=1e-9, M=7, threads=1, ld_individuals=None, dump_windows=True, dump_ld=False,
                do_roh=True):
    """chroms: list of dicts with name, cen (start,end), pos int32[L], gpos float64[L] or None,
    freq float64[L], geno int8[L,N] (codes 0/1/2/3), gl float64[L,N] or None."""
    use_gl = chroms[0].get("gl") is not None
    ld = np.zeros(0, np.int32) if ld_individuals is None else np.ascontiguousarray(ld_individuals, np.int32)
    with open(path, "wb") as f:
        f.write(b"GRLF")
        f.write(struct.pack("<13i", len(chroms), N, W, max_gap, int(use_gl), int(weighted), int(cm), M, threads,
                            len(ld), int(dump_windows), int(dump_ld), int(do_roh)))
        f.write(struct.pack("<4d", -1.0 if error is None else error, cutoff, overlap_frac, mu))
        for ch in chroms:
            L = len(ch["pos"])
            f.write(struct.pack("<i", L))
            f.write(ch["name"].encode()[:15].ljust(16, b"\0"))
            f.write(struct.pack("<2i", int(ch["cen"][0]), int(ch["cen"][1])))
            f.write(np.ascontiguousarray(ch["pos"], np.int32).tobytes())
            gp = ch.get("gpos")
            f.write(np.ascontiguousarray(np.zeros(L) if gp is None else gp, np.float64).tobytes())
            f.write(np.ascontiguousarray(ch["freq"], np.float64).tobytes())
            g = np.ascontiguousarray(ch["geno"], np.int8)
            assert g.shape == (L, N)
            f.write(g.tobytes())
            if use_gl:
                f.write(np.ascontiguousarray(ch["gl"], np.float64).tobytes())
        f.write(ld.tobytes())


def read_output(path, chroms, N, W, dump_windows=True, dump_ld=False):
    with open(path, "rb") as f:
        t_win, t_ld, t_roh = struct.unpack("<3d", f.read(24))
        (n_roh,) = struct.unpack("<q", f.read(8))
        rec = np.frombuffer(f.read(32 * n_roh), dtype=np.dtype([("ind", "<i4"), ("chr", "<i4"), ("start", "<f8"),
                                                                  ("stop", "<f8"), ("length", "<f8")]))
        out = dict(t_windows=t_win, t_ld=t_ld, t_roh=t_roh,
                   roh=[(int(r["ind"]), int(r["chr"]), int(r["start"]), int(r["stop"]), float(r["length"])) for r in rec])
        if dump_windows:
            out["win"] = []
            for ch in chroms:
                L = len(ch["pos"])
                out["win"].append(np.frombuffer(f.read(8 * N * L), np.float64).reshape(N, L))
        if dump_ld:
            out["LD"] = []
            for ch in chroms:
                L = len(ch["pos"])
                out["LD"].append(np.frombuffer(f.read(8 * L * W), np.float64).reshape(L, W))
    return out


def run(chroms, N, W, error, **kw):
    """One run of the reference functions; returns read_output()'s dict."""
    if not available():
        raise RuntimeError("oracle/_ref/ref_driver is not built (make -C oracle ref; needs /root/reference)")
    dump_windows = kw.get("dump_windows", True)
    dump_ld = kw.get("dump_ld", False)
    with tempfile.TemporaryDirectory() as tmp:
        pin, pout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
        write_input(pin, chroms, N, W, error, **kw)
        r = subprocess.run([BIN, pin, pout], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        if r.returncode != 0:
            raise RuntimeError("ref_driver failed: " + r.stderr[-500:])
        return read_output(pout, chroms, N, W, dump_windows, dump_ld)
