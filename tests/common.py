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
