"""Synthetic genotype data for tests and benchmarks.

Two flavours:
  * ``make_dataset`` – small text-level datasets (allele characters, positions, map scaffold, GL
    values) that can be written as tped/tfam/map/tgls files for the reference ``garlic`` binary
    and fed to the oracle / CUDA path as arrays.
  * ``make_packed`` – large pre-coded genotype matrices (2 bits per call, individual-major,
    the HBM layout of DESIGN.md §3) for benchmark shapes, generated without a text detour.

Genotype codes everywhere: 0/1/2 = copies of the "1" allele, 3 = missing (the reference's -9,
garlic-data.cpp:105-128).
"""
from __future__ import annotations

import gzip
import os
from dataclasses import dataclass, field

import numpy as np

# hg19 chromosome lengths (bp), autosomes 1..22 – used only to shape synthetic positions.
HG19_LEN = [249250621, 243199373, 198022430, 191154276, 180915260, 171115067, 159138663,
            146364022, 141213431, 135534747, 135006516, 133851895, 115169878, 107349540,
            102531392, 90354753, 81195210, 78077248, 59128983, 63025520, 48129895, 51304566]


def load_centromeres(build: str) -> dict:
    """Centromere table (reference garlic-centromeres.cpp:185,329,474) as {chr: (start, end)}."""
    here = os.path.dirname(os.path.abspath(__file__))
    out = {}
    with open(os.path.join(here, "centromeres.tsv")) as f:
        for line in f:
            if line.startswith("#"):
                continue
            b, c, s, e = line.split()
            if b == build:
                out[c] = (int(s), int(e))
    return out


@dataclass
class Dataset:
    chr_names: list            # e.g. ["1", "2"] as written in the tped
    chr_offsets: np.ndarray    # int64[C+1] into the concatenated SNP axis (pre-filter)
    pos: np.ndarray            # int32[L0] physical positions
    snp_ids: list              # L0 strings
    alleles: np.ndarray        # uint8[L0, N, 2] allele characters ('A','C','G','T', missing '0')
    ind_ids: list
    pop: str = "POP"
    centromeres: dict = field(default_factory=dict)   # {"chr1": (start, end)}
    map_pos: list = None       # per chr: int32 scaffold physical positions
    map_cm: list = None        # per chr: float64 scaffold genetic positions
    gl: np.ndarray = None      # float64[L0, N] raw tgls values
    gl_type: str = None
    tped_missing: str = "0"

    @property
    def n_ind(self):
        return self.alleles.shape[1]

    @property
    def n_loci(self):
        return self.alleles.shape[0]

    # ---------------------------------------------------------------- writers (reference inputs)
    def write(self, outdir: str, base: str = "syn", gz: bool = False) -> dict:
        os.makedirs(outdir, exist_ok=True)
        paths = {}
        opener = (lambda p: gzip.open(p, "wt")) if gz else (lambda p: open(p, "w"))
        ext = ".gz" if gz else ""
        tped = os.path.join(outdir, base + ".tped" + ext)
        with opener(tped) as f:
            c = 0
            for l in range(self.n_loci):
                while l >= self.chr_offsets[c + 1]:
                    c += 1
                a = self.alleles[l].reshape(-1)
                f.write("%s %s 0 %d %s\n" % (self.chr_names[c], self.snp_ids[l], self.pos[l],
                                             " ".join(map(chr, a))))
        paths["tped"] = tped
        tfam = os.path.join(outdir, base + ".tfam")
        with open(tfam, "w") as f:
            for i in self.ind_ids:
                f.write("%s %s 0 0 0 0\n" % (self.pop, i))
        paths["tfam"] = tfam
        if self.centromeres:
            cen = os.path.join(outdir, base + ".centromeres")
            with open(cen, "w") as f:
                for k, (s, e) in self.centromeres.items():
                    f.write("%s %d %d\n" % (k, s, e))
            paths["centromere"] = cen
        if self.map_pos is not None:
            mp = os.path.join(outdir, base + ".map")
            with open(mp, "w") as f:
                for c, name in enumerate(self.chr_names):
                    for p, g in zip(self.map_pos[c], self.map_cm[c]):
                        f.write("%s m%d_%d %s %d\n" % (name, c, p, repr(float(g)), p))
            paths["map"] = mp
        if self.gl is not None:
            tg = os.path.join(outdir, base + ".tgls" + ext)
            with opener(tg) as f:
                c = 0
                for l in range(self.n_loci):
                    while l >= self.chr_offsets[c + 1]:
                        c += 1
                    f.write("%s %s 0 %d %s\n" % (self.chr_names[c], self.snp_ids[l], self.pos[l],
                                                 " ".join(repr(float(v)) for v in self.gl[l])))
            paths["tgls"] = tg
        return paths


def _positions(rng, n, start, mean_gap, big_gap_frac, big_gap_bp, centro=None):
    gaps = rng.geometric(1.0 / mean_gap, size=n).astype(np.int64)
    big = rng.random(n) < big_gap_frac
    gaps[big] += big_gap_bp
    pos = start + np.cumsum(gaps)
    if centro is not None:
        cs, ce = centro
        # open a hole: shift everything at/after the centromere start past its end
        inside = pos >= cs
        pos[inside] += (ce - cs)
    return pos.astype(np.int32)


def make_dataset(seed=0, n_ind=30, chr_sizes=(2500, 2500), mean_gap=5000, big_gap_frac=0.0015,
                 big_gap_bp=250000, miss=0.01, half_miss=0.003, mono_frac=0.02, n_roh=4,
                 roh_snps=(80, 400), centromere="custom", with_map=False, map_every=7,
                 gl_type=None, chr_names=None, start_bp=1000000) -> Dataset:
    """Small text-level dataset. ``centromere``: "custom" places an interval inside each chromosome's
    data span (one strictly between two SNPs); "hg19" uses the hg19 table; None = none."""
    rng = np.random.default_rng(seed)
    C = len(chr_sizes)
    chr_names = chr_names or [str(c + 1) for c in range(C)]
    offs = np.zeros(C + 1, np.int64)
    offs[1:] = np.cumsum(chr_sizes)
    L0 = int(offs[-1])
    pos = np.zeros(L0, np.int32)
    cents = {}
    hg19 = load_centromeres("hg19") if centromere == "hg19" else None
    for c in range(C):
        n = chr_sizes[c]
        if centromere == "hg19":
            cen = hg19["chr" + chr_names[c]]
            st = max(1, cen[0] - (n // 2) * mean_gap)
            p = _positions(rng, n, st, mean_gap, big_gap_frac, big_gap_bp, cen)
        else:
            p = _positions(rng, n, start_bp, mean_gap, big_gap_frac, big_gap_bp)
            if centromere == "custom":
                k = n // 2 + int(rng.integers(-n // 8, n // 8))
                if c % 2 == 0 and p[k + 1] - p[k] > 2:
                    cents["chr" + chr_names[c]] = (int(p[k]) + 1, int(p[k + 1]) - 1)  # strictly between SNPs
                else:
                    cents["chr" + chr_names[c]] = (int(p[k]), int(p[k + 3]))          # swallows SNPs k..k+3
        pos[offs[c]:offs[c + 1]] = p
    if centromere == "hg19":
        cents = {"chr" + nm: hg19["chr" + nm] for nm in chr_names}

    # genotypes
    p1 = rng.uniform(0.05, 0.95, L0)
    mono = rng.random(L0) < mono_frac
    p1[mono] = rng.integers(0, 2, mono.sum()).astype(float)
    a = (rng.random((L0, n_ind, 2)) < p1[:, None, None])       # True = "1" allele
    # planted ROH: copy first haplotype over the second
    for i in range(n_ind):
        for _ in range(n_roh):
            c = int(rng.integers(0, C))
            ln = int(rng.integers(roh_snps[0], roh_snps[1]))
            ln = min(ln, chr_sizes[c] - 1)
            s = int(offs[c] + rng.integers(0, chr_sizes[c] - ln))
            a[s:s + ln, i, 1] = a[s:s + ln, i, 0]
    letters = np.frombuffer(b"ACGT", np.uint8)
    l1 = rng.integers(0, 4, L0)
    l0 = (l1 + rng.integers(1, 4, L0)) % 4
    alle = np.where(a, letters[l1][:, None, None], letters[l0][:, None, None]).astype(np.uint8)
    m = rng.random((L0, n_ind)) < miss
    alle[m] = ord("0")
    hm = rng.random((L0, n_ind)) < half_miss
    side = rng.integers(0, 2, (L0, n_ind))
    for s in (0, 1):
        sel = hm & (side == s)
        alle[sel, s] = ord("0")
    # a fully-missing SNP and an all-missing leading individual at one SNP (exercise oneAllele logic)
    if L0 > 50:
        alle[17, :, :] = ord("0")
        alle[23, 0, :] = ord("0")

    ds = Dataset(chr_names=list(chr_names), chr_offsets=offs, pos=pos,
                 snp_ids=["rs%d" % i for i in range(L0)], alleles=alle,
                 ind_ids=["ind%d" % i for i in range(n_ind)], centromeres=cents)
    if with_map:
        ds.map_pos, ds.map_cm = [], []
        for c in range(C):
            p = pos[offs[c]:offs[c + 1]]
            # scaffold starts a few SNPs in and ends a few SNPs early → out-of-bounds sites exist
            sel = p[3:-4:map_every].astype(np.int64)
            ds.map_pos.append(sel.astype(np.int32))
            ds.map_cm.append(np.round(sel * 1.2e-6 + 0.05 * np.sin(sel * 1e-6), 9))
    if gl_type:
        ds.gl_type = gl_type
        if gl_type == "PL":
            vals = np.array([0.0004, 0.004, 0.04, 0.5, 3.0, 0.0, 150.0])
        elif gl_type == "GL":
            vals = -np.array([0.0004, 0.004, 0.04, 0.5, 3.0, 0.0, 15.0]) / 10.0
        else:  # GQ
            vals = np.array([30.0, 20.0, 13.0, 50.0, 3.0, 0.0, 120.0])
        pr = np.array([0.3, 0.3, 0.2, 0.1, 0.08, 0.01, 0.01])
        ds.gl = vals[rng.choice(len(vals), size=(L0, n_ind), p=pr)]
    return ds


# ------------------------------------------------------------------------------------------------
# Host restatement of the coding step for *array-level* use in tests (not the oracle – the oracle's
# own C version lives in oracle/): characters → codes, used to build packed inputs for big cases.
# ------------------------------------------------------------------------------------------------
def pack_codes(codes_ind_major: np.ndarray, row_bytes: int | None = None) -> np.ndarray:
    """codes uint8[N, L] in {0,1,2,3} → uint8[N, row_bytes], 4 SNPs per byte, SNP s at bits
    2*(s%4) of byte s//4."""
    N, L = codes_ind_major.shape
    nb = (L + 3) // 4
    rb = row_bytes or ((nb + 15) // 16) * 16
    pad = np.full((N, nb * 4), 3, np.uint8)
    pad[:, :L] = codes_ind_major
    q = pad.reshape(N, nb, 4)
    by = (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).astype(np.uint8)
    out = np.full((N, rb), 0xFF, np.uint8)
    out[:, :nb] = by
    return out


def unpack_codes(packed: np.ndarray, L: int) -> np.ndarray:
    N = packed.shape[0]
    b = packed[:, :(L + 3) // 4]
    out = np.empty((N, b.shape[1], 4), np.uint8)
    for k in range(4):
        out[:, :, k] = (b >> (2 * k)) & 3
    return out.reshape(N, -1)[:, :L]


def make_positions_genomewide(seed, L, n_chr=22, big_gap_frac=0.001, build="hg19"):
    """Positions for benchmark shapes: n_chr autosomes with L_c ∝ hg19 length, geometric gaps,
    centromere hole per the build table, big_gap_frac of gaps forced > 200 kb."""
    rng = np.random.default_rng(seed)
    lens = np.array(HG19_LEN[:n_chr], np.float64)
    sizes = np.floor(lens / lens.sum() * L).astype(np.int64)
    sizes[0] += L - sizes.sum()
    cents = load_centromeres(build)
    offs = np.zeros(n_chr + 1, np.int64)
    offs[1:] = np.cumsum(sizes)
    pos = np.zeros(L, np.int32)
    for c in range(n_chr):
        n = int(sizes[c])
        cen = cents["chr%d" % (c + 1)]
        span = HG19_LEN[c] - (cen[1] - cen[0]) - 250000 * int(n * big_gap_frac + 3)
        mean_gap = max(2.0, span * 0.95 / n)
        p = _positions(rng, n, 10000, mean_gap, big_gap_frac, 250000, cen)
        pos[offs[c]:offs[c + 1]] = p
    names = [str(c + 1) for c in range(n_chr)]
    return names, offs, pos, {"chr" + nm: cents["chr" + nm] for nm in names}


def make_codes(seed, n_ind, L, miss=0.005, mono_frac=0.02, n_roh=5, roh_snps=(200, 2000),
               chunk=4096):
    """uint8[N, L] genotype codes, individual-major, Binomial(2,p) with planted homozygous tracts."""
    rng = np.random.default_rng(seed)
    p = rng.uniform(0.05, 0.95, L).astype(np.float32)
    mono = rng.random(L) < mono_frac
    p[mono] = 0.0
    out = np.empty((n_ind, L), np.uint8)
    for i0 in range(0, n_ind, chunk):
        n = min(chunk, n_ind - i0)
        h0 = rng.random((n, L), dtype=np.float32) < p
        h1 = rng.random((n, L), dtype=np.float32) < p
        for i in range(n):
            for _ in range(n_roh):
                ln = int(rng.integers(roh_snps[0], roh_snps[1]))
                s = int(rng.integers(0, L - ln))
                h1[i, s:s + ln] = h0[i, s:s + ln]
        g = (h0.astype(np.uint8) + h1.astype(np.uint8))
        m = rng.random((n, L), dtype=np.float32) < miss
        g[m] = 3
        out[i0:i0 + n] = g
    return out
