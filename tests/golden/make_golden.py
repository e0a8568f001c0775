#!/usr/bin/env python3
"""Generate tests/golden/*: synthetic inputs + outputs of the UNMODIFIED reference binary
(/root/reference/bin/linux/garlic, v1.1.6a) run in this container.

Run from the repo root:  python tests/golden/make_golden.py
Each case directory holds  data.npz (the synthetic dataset), cmd.txt (reference command line),
out.roh.bed, out.freq (decompressed), out.log, optionally out.kde and rawlod.npz (the --raw-lod
dump parsed to float64, NaN = "NA").  The reference is made deterministic with --lod-cutoff /
--size-bounds or --kde-subsample 0 / --ld-subsample 0.
"""
import glob
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from garlic_b200 import synth  # noqa: E402

REF = "/root/reference/bin/linux/garlic"
OUT = os.path.join(ROOT, "tests", "golden")


def save_dataset(ds, path):
    d = dict(chr_names=np.array(ds.chr_names), chr_offsets=ds.chr_offsets, pos=ds.pos,
             alleles=ds.alleles, ind_ids=np.array(ds.ind_ids), pop=np.array(ds.pop),
             cen_keys=np.array(list(ds.centromeres.keys())),
             cen_vals=np.array(list(ds.centromeres.values()), np.int64).reshape(-1, 2))
    if ds.map_pos is not None:
        for c in range(len(ds.chr_names)):
            d["map_pos_%d" % c] = ds.map_pos[c]
            d["map_cm_%d" % c] = ds.map_cm[c]
    if ds.gl is not None:
        d["gl"] = ds.gl
        d["gl_type"] = np.array(ds.gl_type)
    np.savez_compressed(path, **d)


ONLY = set(sys.argv[1:])     # optional: names of the cases to (re)generate


def run_case(name, ds, args, build=None, raw=False, keep_kde=False, freq_text=None):
    if ONLY and name not in ONLY:
        return
    cdir = os.path.join(OUT, name)
    if os.path.isdir(cdir):
        shutil.rmtree(cdir)
    os.makedirs(cdir)
    with tempfile.TemporaryDirectory() as tmp:
        p = ds.write(tmp)
        cmd = [REF, "--tped", p["tped"], "--tfam", p["tfam"], "--out", os.path.join(tmp, "out")]
        if build:
            cmd += ["--build", build]
        elif "centromere" in p:
            cmd += ["--centromere", p["centromere"]]
        if "map" in p and ("--weighted" in args or "--cm" in args):
            cmd += ["--map", p["map"]]
        if "tgls" in p:
            cmd += ["--tgls", p["tgls"], "--gl-type", ds.gl_type]
        if raw:
            cmd += ["--raw-lod"]
        if freq_text is not None:
            with open(os.path.join(tmp, "syn.freq"), "w") as f:
                f.write(freq_text)
            cmd += ["--freq-file", os.path.join(tmp, "syn.freq")]
        cmd += args
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stdout[-2000:], r.stderr[-2000:])
            raise SystemExit("reference failed for " + name)
        save_dataset(ds, os.path.join(cdir, "data.npz"))
        if freq_text is not None:
            with open(os.path.join(cdir, "in.freq"), "w") as f:
                f.write(freq_text)
        shown = [a.replace(tmp, "<tmp>") for a in cmd]
        with open(os.path.join(cdir, "cmd.txt"), "w") as f:
            f.write(" ".join(shown) + "\n")
            json.dump(args, f)
            f.write("\n")
        for fn in ("out.roh.bed", "out.log"):
            if os.path.exists(os.path.join(tmp, fn)):
                shutil.copy(os.path.join(tmp, fn), os.path.join(cdir, fn))
        # the log embeds tmp paths on its first lines; scrub them
        lg = open(os.path.join(cdir, "out.log")).read().replace(tmp, "<tmp>")
        open(os.path.join(cdir, "out.log"), "w").write(lg)
        if os.path.exists(os.path.join(tmp, "out.freq.gz")):
            with gzip.open(os.path.join(tmp, "out.freq.gz"), "rt") as f, open(os.path.join(cdir, "out.freq"), "w") as g:
                g.write(f.read())
        for k in glob.glob(os.path.join(tmp, "out.*SNPs.kde")):
            shutil.copy(k, os.path.join(cdir, os.path.basename(k)))
        if raw:
            arrs = {}
            for c, nm in enumerate(ds.chr_names):
                lab = nm if nm[0] == "c" else "chr" + nm
                fn = os.path.join(tmp, "out.%s.%s.raw.lod.windows.gz" % (ds.pop, lab))
                rows = []
                with gzip.open(fn, "rt") as f:
                    for line in f:
                        rows.append([np.nan if t == "NA" else float(t) for t in line.split()])
                arrs["chr%d" % c] = np.array(rows, np.float64)
            np.savez_compressed(os.path.join(cdir, "rawlod.npz"), **arrs)
    n = sum(1 for _ in open(os.path.join(cdir, "out.roh.bed")))
    print("%-28s ok  (%d bed lines)" % (name, n))


def main():
    SB = ["--size-bounds", "500000", "1500000"]
    # 1. unweighted LOD, custom centromeres (one strictly between SNPs, one swallowing SNPs),
    #    gaps > MAX_GAP, missing + half-missing calls, four (W, error, cutoff, overlap) settings
    ds = synth.make_dataset(seed=11, n_ind=30, chr_sizes=(2500, 2500))
    for k, (W, e, cut, ov) in enumerate([(25, "0.002", "1.5", "0.3"), (25, "0.002", "-5", "0"),
                                         (40, "0.002", "-60", "1"), (10, "0.01", "0.5", "0.25")]):
        run_case("lod_%d" % k, ds, ["--winsize", str(W), "--error", e, "--lod-cutoff", cut,
                                    "--overlap-frac", ov] + SB, raw=(k == 0))
    # 2. small, three chromosomes (one shorter than the window), full raw-lod, unknown-centromere chr
    ds = synth.make_dataset(seed=12, n_ind=12, chr_sizes=(900, 35, 700), chr_names=["1", "7", "Z"],
                            big_gap_frac=0.004)
    ds.centromeres.pop("chrZ", None)
    run_case("lod_small", ds, ["--winsize", "50", "--error", "0.001", "--lod-cutoff", "2.0"] + SB, raw=True)
    # 3. weighted wLOD with --cm and a map scaffold
    ds = synth.make_dataset(seed=13, n_ind=24, chr_sizes=(1400, 1200), with_map=True)
    run_case("wlod_cm", ds, ["--winsize", "25", "--error", "0.001", "--weighted", "--cm", "--threads", "3",
                             "--ld-subsample", "0", "--lod-cutoff", "0.5",
                             "--size-bounds", "0.5", "1.5"], raw=True)
    # 3b. the same data read as phased haplotypes: r2 LD from the first-copy bits instead of hr2 (--phased)
    run_case("wlod_phased", ds, ["--winsize", "25", "--error", "0.001", "--weighted", "--cm", "--phased", "--threads", "2",
                                 "--ld-subsample", "0", "--lod-cutoff", "0.5",
                                 "--size-bounds", "0.5", "1.5"], raw=True)
    run_case("lod_cm", ds, ["--winsize", "30", "--error", "0.001", "--cm", "--lod-cutoff", "1.0",
                            "--size-bounds", "0.5", "1.5"])
    # 4. genotype likelihoods, all three encodings
    for t in ("PL", "GL", "GQ"):
        ds = synth.make_dataset(seed=14, n_ind=16, chr_sizes=(1500, 1000), gl_type=t)
        run_case("gl_%s" % t.lower(), ds, ["--winsize", "30", "--lod-cutoff", "1.0"] + SB, raw=(t == "PL"))
    # 5. automatic cutoff (KDE over all individuals), automatic GMM bounds
    ds = synth.make_dataset(seed=15, n_ind=40, chr_sizes=(3000, 3000), n_roh=6)
    run_case("auto_cutoff", ds, ["--winsize", "30", "--error", "0.001", "--kde-subsample", "0"], keep_kde=True)
    # 6. --auto-overlap-frac with the hg19 table
    ds = synth.make_dataset(seed=16, n_ind=20, chr_sizes=(2500, 2500), centromere="hg19",
                            chr_names=["1", "2"])
    run_case("auto_overlap_hg19", ds, ["--winsize", "60", "--error", "0.001", "--auto-overlap-frac",
                                      "--lod-cutoff", "2.5"] + SB, build="hg19")
    # 7. --winsize-multi + --auto-winsize (selection from a list), KDE over all individuals
    ds = synth.make_dataset(seed=17, n_ind=24, chr_sizes=(3000, 2500), n_roh=6)
    run_case("winsize_multi", ds, ["--winsize-multi", "20", "30", "40", "--auto-winsize", "--error", "0.001",
                                  "--kde-subsample", "0"] + SB, keep_kde=True)
    # 7b. --auto-winsize alone: selectWinsize steps the window size up from --winsize until the KDE is smooth
    #     (garlic-roh.cpp:766-850); --auto-winsize --weighted: the size comes from the SNP density (:3-9);
    #     --no-kde-thinning: every window of every individual goes into the KDE (garlic-cli.cpp:171-174)
    ds = synth.make_dataset(seed=19, n_ind=24, chr_sizes=(3000, 2500), n_roh=6)
    run_case("auto_winsize", ds, ["--auto-winsize", "--winsize", "20", "--error", "0.001", "--kde-subsample", "0"] + SB,
             keep_kde=True)
    run_case("no_kde_thinning", ds, ["--winsize", "30", "--error", "0.001", "--kde-subsample", "0", "--no-kde-thinning"] + SB,
             keep_kde=True)
    ds = synth.make_dataset(seed=20, n_ind=24, chr_sizes=(1600, 1400), with_map=True)
    run_case("auto_winsize_weighted", ds, ["--auto-winsize", "--weighted", "--cm", "--error", "0.001", "--threads", "2",
                                           "--ld-subsample", "0", "--lod-cutoff", "0.5", "--size-bounds", "0.5", "1.5"])
    # 8. --freq-file: frequencies from a "panel" (differ from the sample's), 30 % of the rows name the other allele
    #    (the program must flip them, garlic-data.cpp:1422), a few rows are 0 / 1 (filtered although polymorphic here)
    ds = synth.make_dataset(seed=18, n_ind=20, chr_sizes=(2000, 1500))
    from oracle import oracle as orc
    geno, na, tot, one, freq = orc.code_tped(ds.alleles)
    rng = np.random.default_rng(18)
    panel = np.clip(freq + rng.normal(0, 0.05, len(freq)), 0.0, 1.0)
    panel[rng.random(len(freq)) < 0.01] = 0.0
    flip = rng.random(len(freq)) < 0.3
    lines = ["CHR\tSNP\tPOS\tALLELE\tFREQ"]
    c = 0
    for l in range(ds.n_loci):
        while l >= ds.chr_offsets[c + 1]:
            c += 1
        a = one[l]
        f = panel[l]
        if flip[l]:
            others = [x for x in np.unique(ds.alleles[l]) if x != a and x != ord("0")]
            a = others[0] if others else ord("N")
            f = 1 - f
        lines.append("chr%s\t%s\t%d\t%s\t%s" % (ds.chr_names[c], ds.snp_ids[l], ds.pos[l], chr(a), "%g" % f))
    run_case("freq_file", ds, ["--winsize", "40", "--error", "0.001", "--lod-cutoff", "2.0"] + SB,
             freq_text="\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
