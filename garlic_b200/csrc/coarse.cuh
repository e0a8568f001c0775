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
    const uint32_t* mask;    // per half-word q: bit 2i set <=> at SNP 16q+i the genotype-2 homozygote is the "rare"
                             // one; bit 2i+1 set <=> D[16q+i] > kCoarseSplit (high class)
    const int2* cb;          // per block k (valid for k >= -16): x = Bmax, y = Dlo | Dhi << 16 (fixed point)
    int chet_fixed;          // c_het in fixed point (negative), rounded toward zero
    int cut_fixed;           // (cutoff - tol) in fixed point, rounded down, minus slack
    int W, c1, c2;
    int n_lanes;
};

constexpr int kCoarseC2Min = (kCoarseMinW + 14) >> 4, kCoarseC2Max = (kCoarseMaxW + 14) >> 4;   // 2 .. 13

// Does individual `ind` have any block of item `it` whose bound reaches the cutoff?
// The packed counts p[q] = nhet | nrare_lo << 8 | nrare_hi << 16 of the last 16 half-words live in registers
// (the scan is unrolled over 16 half-words = 4 sixteen-byte quads, so every ring slot is a compile-time
// register); sliding sums over the span (half-words k..k+C2) and the core (k+1..k+c1) are kept in the same
// packed form (every field stays < 256 for W <= kCoarseMaxW).
template <int C2>
GHD bool coarse_item(const CoarseParams& P, const Item& it, int ind)
{
    const uint4* row = reinterpret_cast<const uint4*>(P.geno + (int64_t)ind * P.row_words);
    const int k_lo = it.w0 >> 4, k_hi = (it.own_hi - 1) >> 4;      // blocks holding the item's windows
    const int q_end = k_hi + C2;
    const bool c1_near = (P.c1 == C2 - 1);                          // c1 is C2-1 or C2-2
    const uint32_t k_span = (uint32_t)(k_hi - k_lo);
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = 0u;
    uint32_t s_span = 0, s_core = 0;
    bool cand = false;
    for (int qb = k_lo & ~15; qb <= q_end; qb += 16) {
        const uint32_t* mq = P.mask + qb;
        const int2* cbq = P.cb + (qb - C2);
        const uint4* rq = row + (qb >> 2);
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
            const uint4 quad = rq[i4];
            const uint32_t hw[4] = {quad.x, quad.y, quad.z, quad.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i4 * 4 + j;                           // = q & 15, static after unrolling
                const uint32_t h = hw[j], m = mq[i];
                const uint32_t lo = h & 0x55555555u, hi = (h >> 1) & 0x55555555u;
                const uint32_t het = lo & ~hi, hom2 = hi & ~lo, hom0 = ~(lo | hi) & 0x55555555u;
                const uint32_t rare = (hom2 & m) | (hom0 & ~m);
                const uint32_t high = (m >> 1) & 0x55555555u;
                const uint32_t p = (uint32_t)popc32(het) | ((uint32_t)popc32(rare & ~high) << 8) | ((uint32_t)popc32(rare & high) << 16);
                r[i] = p;
                // block k = q - C2: span = half-words k..q, core = k+1..k+c1
                s_span += p - r[(i - C2 - 1) & 15];                 // p[k-1] leaves the span
                s_core += (c1_near ? r[(i - 1) & 15] : r[(i - 2) & 15]) - r[(i - C2) & 15];   // p[k+c1] enters, p[k] leaves
                const int2 c = cbq[i];
                const int ub = c.x + P.chet_fixed * (int)(s_core & 0xffu) + (int)((uint32_t)c.y & 0xffffu) * (int)((s_span >> 8) & 0xffu) +
                               (int)((uint32_t)c.y >> 16) * (int)((s_span >> 16) & 0xffu);
                cand |= (ub >= P.cut_fixed) && ((uint32_t)(qb + i - C2 - k_lo) <= k_span);
            }
        }
    }
    return cand;
}

// run-time window size → compile-time C2
GHD bool coarse_item_any(const CoarseParams& P, const Item& it, int ind)
{
    switch (P.c2) {
        case 2: return coarse_item<2>(P, it, ind);
        case 3: return coarse_item<3>(P, it, ind);
        case 4: return coarse_item<4>(P, it, ind);
        case 5: return coarse_item<5>(P, it, ind);
        case 6: return coarse_item<6>(P, it, ind);
        case 7: return coarse_item<7>(P, it, ind);
        case 8: return coarse_item<8>(P, it, ind);
        case 9: return coarse_item<9>(P, it, ind);
        case 10: return coarse_item<10>(P, it, ind);
        case 11: return coarse_item<11>(P, it, ind);
        case 12: return coarse_item<12>(P, it, ind);
        default: return coarse_item<13>(P, it, ind);
    }
}

}  // namespace garlic
