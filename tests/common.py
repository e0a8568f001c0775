"""Shared helpers for tests: golden-case loading and BED parsing."""
import json
import os

import numpy as np

from garlic_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    d = np.load(os.path.join(GOLDEN, name, "data.npz"), allow_pickle=False)
    names = [str(x) for x in d["chr_names"]]
    cents = {str(k): (int(v[0]), int(v[1])) for k, v in zip(d["cen_keys"], d["cen_vals"])}
    ds = synth.Dataset(chr_names=names, chr_offsets=d["chr_offsets"], pos=d["pos"],
                       snp_ids=["rs%d" % i for i in range(len(d["pos"]))], alleles=d["alleles"],
                       ind_ids=[str(x) for x in d["ind_ids"]], pop=str(d["pop"]), centromeres=cents)
    if "map_pos_0" in d:
        ds.map_pos = [d["map_pos_%d" % c] for c in range(len(names))]
        ds.map_cm = [d["map_cm_%d" % c] for c in range(len(names))]
    if "gl" in d:
        ds.gl = d["gl"]
        ds.gl_type = str(d["gl_type"])
    with open(os.path.join(GOLDEN, name, "cmd.txt")) as f:
        f.readline()
        args = json.loads(f.readline())
    return ds, args


def arg(args, flag, default=None, cast=str):
    if flag in args:
        return cast(args[args.index(flag) + 1])
    return default


def arg_list(args, flag, cast=float):
    if flag not in args:
        return None
    out = []
    for a in args[args.index(flag) + 1:]:
        if a.startswith("--"):
            break
        out.append(cast(a))
    return out


def golden_text(name, fn):
    with open(os.path.join(GOLDEN, name, fn)) as f:
        return f.read()


def log_value(name, key):
    for line in golden_text(name, "out.log").splitlines():
        if line.startswith(key):
            return line[len(key):].strip()
    return None


def parse_bed(text):
    """→ list per track of (chr, start, stop, cls, size_str)."""
    tracks = []
    for line in text.splitlines():
        if line.startswith("track"):
            tracks.append([])
        elif line.strip():
            t = line.split("\t")
            tracks[-1].append((t[0], int(t[1]), int(t[2]), t[3], t[4]))
    return tracks


# ---- flattening an oracle result into the concatenated, filtered arrays the product works on ----
PAD = 4096 + 64


def flatten(res, error, use_gl=False):
    """oracle.run_pipeline result → dict(codes[N,L] uint8, rows uint64[N,row_words], chr_off, pos, cen,
    freq, lut[L+PAD,4], gl[N,L+PAD] or None)."""
    from oracle import oracle as orc
    chroms = res["chroms"]
    codes = np.concatenate([ch["geno"] for ch in chroms], axis=0).T.copy().astype(np.uint8)   # [N, L]
    N, L = codes.shape
    row_words = (((L + PAD + 31) >> 5) + 3) & ~1
    rows8 = synth.pack_codes(codes, row_bytes=row_words * 8)
    rows = rows8.view(np.uint64).reshape(N, row_words)
    chr_off = np.zeros(len(chroms) + 1, np.int64)
    chr_off[1:] = np.cumsum([len(ch["pos"]) for ch in chroms])
    pos = np.concatenate([ch["pos"] for ch in chroms]).astype(np.int32)
    cen = np.array([ch["cen"] for ch in chroms], np.int32).reshape(-1)
    freq = np.zeros(L + PAD)
    freq[:L] = np.concatenate([ch["freq"] for ch in chroms])
    lut = np.zeros((L + PAD, 4))
    if error is not None:
        lut[:L] = orc.lod_lut(freq[:L], error)
    gl = None
    if use_gl:
        # GL mode: the product streams per-genotype LOD values (lod() of each genotype's own error rate)
        gl = np.zeros((N, L + PAD))
        gl[:, :L] = np.concatenate([orc.lod_matrix(ch["geno"], ch["freq"], ch["gl"]) for ch in chroms], axis=0).T
    return dict(codes=codes, rows=rows, row_words=row_words, chr_off=chr_off, pos=pos, cen=cen, freq=freq,
                lut=lut, gl=gl, N=N, L=L)


def oracle_roh_idx(res):
    """[(ind, chr, a_global, b_global)] from the oracle's per-chromosome indices."""
    offs = np.zeros(len(res["chroms"]) + 1, np.int64)
    offs[1:] = np.cumsum([len(ch["pos"]) for ch in res["chroms"]])
    return [(r[0], r[1], int(offs[r[1]] + r[5]), int(offs[r[1]] + r[6])) for r in res["roh"]]


def oracle_windows_matrix(res, step=1):
    """Dense [N, slots] matrix in the product's dump layout (MISSING where the reference has MISSING)."""
    mats = [ch["win"][:, ::step] for ch in res["chroms"]]
    return np.concatenate(mats, axis=1)
