// coarse.cuh — the pruning pass in front of the fused window → ROH walker (K5 pass 2, unweighted, table mode).
//
// Most windows are nowhere near the cutoff: outside runs of homozygosity a window holds many heterozygous
// calls, each worth log10(error) < 0.  This pass proves that cheaply, per individual and per chunk (item),
// from bit operations on the packed genotypes alone — no table lookups, no floating point:
//
//   lod(s,g) = base[s] + {0 for the hom genotype with the smaller table value; D[s] >= 0 for the other hom
//              genotype ("rare" below); c_het - base[s] <= c_het for a heterozygote; -base[s] <= 0 if missing}
//   (base[s] = min(lut[s][0], lut[s][2]) >= 0, D[s] = |lut[s][0]-lut[s][2]|, c_het = max_s lut[s][1] < 0;
//    the table is lod() of the reference, src/garlic-roh.cpp:355-386)
//
//   => for every window t starting in the 16-SNP block k (t in [16k,16k+16)):
//        win(t) <= Bmax[k] + c_het * nhet_core(k) + Dlo[k] * nrare_lo(k) + Dhi[k] * nrare_hi(k)
//      nhet_core  = heterozygotes in the half-words every such window contains   (k+1 .. k+c1, c1=(W-16)>>4)
//      nrare_*    = "rare" homozygotes in the half-words any such window touches  (k .. k+c2,   c2=(W+14)>>4),
//                   split by whether D[s] <= kCoarseSplit (common SNPs, small D) or above (low-MAF SNPs)
//      Bmax[k]    = max_t sum_{s in [t,t+W)} base[s],   Dlo/Dhi[k] = max of D over the span within each class
//
// evaluated in integers (fixed point 2^-8, tables rounded up).  An (individual, item) pair none of whose
// blocks reaches cutoff - tol can hold no flagged window (src/garlic-roh.cpp:450), hence no coverage and no
// ROH, and is dropped; the others are gathered into dense per-item lists that the exact walker processes.
// The bound is conservative by construction; tests/test_host_emu.py checks on the CPU (same source) that every
// pair with a window >= cutoff survives.
#pragma once
#include "common.cuh"
#include "walk.cuh"

namespace garlic {

constexpr int kCoarseShift = 8;         // fixed-point scale of the bound: 1/256 (coefficients fit 16 bits)
constexpr int kCoarseMaxW = 208;        // ring of c2+2 prefix entries per lane; larger windows use the plain walker
constexpr int kCoarseMinW = 32;
constexpr double kCoarseSplit = 0.3;    // D above this is the "high" class of rare homozygotes

struct CoarseParams {
    const uint64_t* geno;
    int64_t row_words;       // even, so that a row is a whole number of 16-byte quads
    const uint2* tab;        // per half-word index j:
                             //   x = rare-allele mask of half-word j: bit 2i set <=> at SNP 16j+i the genotype-2
                             //       homozygote is the "rare" one; bit 2i+1 set <=> D[16j+i] > kCoarseSplit (high class)
                             //   y = bound coefficients of block j: Dlo | Dhi << 16 (unsigned 16-bit, fixed point)
    const int* bmax;         // per block k: Bmax in fixed point
    int chet_fixed;          // c_het in fixed point (negative), rounded toward zero
    int cut_fixed;           // (cutoff - tol) in fixed point, rounded down, minus slack
    int W, c1, c2;
    int n_lanes;
};

GHD int coarse_ring_len(int c2) { return c2 + 3; }

// Does individual `ind` have any block of item `it` whose bound reaches the cutoff?
// ring: per-lane ring (stride rstride) of the last c2+2 half-words' packed counts
//       p[q] = nhet | nrare_lo << 8 | nrare_hi << 16   (each <= 16).
// Sliding sums over the span (half-words k..k+c2) and the core (k+1..k+c1) are kept in the same packed form
// (every field stays < 256 for W <= kCoarseMaxW).  Rows are read one 16-byte quad (64 SNPs) at a time.
GHD bool coarse_item(const CoarseParams& P, const Item& it, int ind, uint32_t* ring, int rstride)
{
    const uint4* row = reinterpret_cast<const uint4*>(P.geno + (int64_t)ind * P.row_words);
    const int k_lo = it.w0 >> 4, k_hi = (it.own_hi - 1) >> 4;      // blocks holding the item's windows
    const int RL = coarse_ring_len(P.c2);
    for (int i = 0; i < RL; ++i) ring[i * rstride] = 0u;
    uint32_t s_span = 0, s_core = 0;
    bool cand = false;
    int wslot = 0;                                                  // slot of p[q]
    const int q_end = k_hi + P.c2;
    for (int q4 = k_lo >> 2; q4 * 4 <= q_end; ++q4) {
        const uint4 quad = row[q4];
        const uint32_t hw[4] = {quad.x, quad.y, quad.z, quad.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int q = q4 * 4 + j;
            const uint32_t h = hw[j];
            const uint2 tq = P.tab[q];
            const uint32_t lo = h & 0x55555555u, hi = (h >> 1) & 0x55555555u;
            const uint32_t het = lo & ~hi, hom2 = hi & ~lo, hom0 = ~(lo | hi) & 0x55555555u;
            const uint32_t rare = (hom2 & tq.x) | (hom0 & ~tq.x);
            const uint32_t high = (tq.x >> 1) & 0x55555555u;
            const uint32_t p = (uint32_t)popc32(het) | ((uint32_t)popc32(rare & ~high) << 8) | ((uint32_t)popc32(rare & high) << 16);
            ring[wslot * rstride] = p;
            // block k = q - c2: span = half-words k..q, core = k+1..k+c1
            int sl = wslot - (P.c2 + 1); if (sl < 0) sl += RL;      // p[k-1] leaves the span
            int sk = sl + 1; if (sk >= RL) sk -= RL;                // p[k] leaves the core ...
            int se = sk + P.c1; if (se >= RL) se -= RL;             // ... and p[k+c1] enters it
            s_span += p - ring[sl * rstride];
            s_core += ring[se * rstride] - ring[sk * rstride];
            const int k = q - P.c2;
            if (k >= k_lo && k <= k_hi) {
                const uint32_t co = P.tab[k].y;
                const int ub = P.bmax[k] + P.chet_fixed * (int)(s_core & 0xffu) + (int)(co & 0xffffu) * (int)((s_span >> 8) & 0xffu) +
                               (int)(co >> 16) * (int)((s_span >> 16) & 0xffu);
                cand |= (ub >= P.cut_fixed);
            }
            if (++wslot >= RL) wslot = 0;
        }
    }
    return cand;
}

}  // namespace garlic
