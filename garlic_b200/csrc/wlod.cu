// wlod.cu — the --weighted path: K6 LD band (hr², reference src/garlic-data.cpp:377-424,474-527,
// 558-583) and K5-W weighted windows fused with ROH assembly (src/garlic-roh.cpp:204-277,409-546).
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "common.cuh"
#include "walk.cuh"
#include "wlod.h"

namespace garlic {

// ------------------------------------------------------------------------------------------
// K6a: SNP-major bit-planes of the LD individuals.  planes[s][0..nw) = non-missing bits,
// planes[s][nw..2nw) = homozygous (g∈{0,2}) bits; individual j of the list is bit j&63 of word j>>6.
// Lanes = consecutive SNPs: a warp reads one packed word per individual (broadcast).
// ------------------------------------------------------------------------------------------
// ld_ind holds indices into the WHOLE sample; this GPU owns individuals [ind_lo, ind_lo + n_local) (its rows 0..) and
// fills only their bits — the other ranks' bits are zero here and arrive by the all-reduce in launch_ld_band.
// ph (--phased): four planes per SNP instead of two — non-missing, g==2, g==1 and the first-copy bit
// firstCopy = (first allele character == the "1" allele), garlic-data.cpp:129, read from the allele block K1 coded.
__global__ void ld_planes_kernel(const uint64_t* __restrict__ geno, int64_t row_words, const int* __restrict__ ld_ind,
                                 int n_ld, long long L, int nw, uint64_t* __restrict__ planes, int ind_lo, int n_local,
                                 LdPhase ph)
{
    const int np = ph.alleles ? 4 : 2;
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < L; s += (long long)gridDim.x * blockDim.x) {
        int one = 0;
        const uint8_t* arow = nullptr;
        if (ph.alleles) {
            const int s0 = ph.src[s];
            const unsigned long long k = ph.key[s0];
            one = (k == ~0ull) ? ph.missing : (int)(k & 0xff);
            arow = ph.alleles + (size_t)s0 * n_local * 2;
        }
        for (int w = 0; w < nw; ++w) {
            uint64_t nm = 0, hm = 0, g1 = 0, fc = 0;
            const int jmax = min(64, n_ld - w * 64);
            for (int j = 0; j < jmax; ++j) {
                const int ind = ld_ind[w * 64 + j] - ind_lo;
                if (ind < 0 || ind >= n_local) continue;
                const int g = (int)(geno[(int64_t)ind * row_words + (s >> 5)] >> (2 * (s & 31))) & 3;
                nm |= (uint64_t)(g != 3) << j;
                if (ph.alleles) {
                    hm |= (uint64_t)(g == 2) << j;
                    g1 |= (uint64_t)(g == 1) << j;
                    fc |= (uint64_t)(arow[2 * ind] == one) << j;
                } else {
                    hm |= (uint64_t)(g == 0 || g == 2) << j;
                }
            }
            planes[s * np * nw + w] = nm;
            planes[s * np * nw + nw + w] = hm;
            if (ph.alleles) {
                planes[s * np * nw + 2 * nw + w] = g1;
                planes[s * np * nw + 3 * nw + w] = fc;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// K6b: ordered pair matrix P[j][d] = hr2(i, j), i = j + d - (W-1), d ∈ [0, 2W-2]; P[j][W-1] = 1.
// hr2 exactly as garlic-data.cpp:558-583 (HA = homFreq[i], HB = homFreq[j]; counts over the LD
// individuals by popcount of the bit-planes).  Entries whose i falls outside j's chromosome are 0
// and never used.
// ------------------------------------------------------------------------------------------
__global__ void ld_pairs_kernel(const uint64_t* __restrict__ planes, int nw, const double* __restrict__ homf,
                                const int* __restrict__ chr_of, const int* __restrict__ chr_start, int n_chr,
                                long long L, int W, double* __restrict__ P)
{
    const int D = 2 * W - 1;
    const long long total = L * D;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long j = t / D;
        const int d = (int)(t % D);
        const long long i = j + d - (W - 1);
        double v = 0.0;
        const int c = chr_of[j];
        const long long lo = chr_start[c], hi = (c + 1 < n_chr) ? chr_start[c + 1] : L;
        if (i == j) v = 1.0;
        else if (i >= lo && i < hi) {
            const double HA = homf[i], HB = homf[j];
            if (HA > 0 && HA < 1 && HB > 0 && HB < 1) {
                const uint64_t* pi = planes + i * 2 * nw;
                const uint64_t* pj = planes + j * 2 * nw;
                int tot = 0, hab = 0;
                for (int w = 0; w < nw; ++w) {
                    tot += __popcll(pi[w] & pj[w]);
                    hab += __popcll(pi[nw + w] & pj[nw + w]);
                }
                double HAB = (double)hab, total_d = (double)tot;
                HAB /= total_d;
                const double H = HAB - HA * HB;
                const double HR2 = H * H / (HA * (1 - HA) * HB * (1 - HB));
                v = (HR2 > 1) ? 1.0 : HR2;
            }
        }
        P[t] = v;
    }
}

// --phased: the same ordered pair matrix with r2 between haplotypes (garlic-data.cpp:585-617):
//   x11 = 2·|G2i∧G2j| + |G1i∧G2j| + |G2i∧G1j| + |G1i∧G1j∧¬(Fi⊕Fj)|,  total = 2·|NMi∧NMj|  (popcounts over the LD
//   individuals), then the reference's divisions and multiplications in its order; p = the allele frequencies in use.
__global__ void ld_pairs_r2_kernel(const uint64_t* __restrict__ planes, int nw, const double* __restrict__ freq,
                                   const int* __restrict__ chr_of, const int* __restrict__ chr_start, int n_chr,
                                   long long L, int W, double* __restrict__ P)
{
    const int D = 2 * W - 1;
    const long long total = L * D;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long j = t / D;
        const int d = (int)(t % D);
        const long long i = j + d - (W - 1);
        double v = 0.0;
        const int c = chr_of[j];
        const long long lo = chr_start[c], hi = (c + 1 < n_chr) ? chr_start[c + 1] : L;
        if (i == j) v = 1.0;
        else if (i >= lo && i < hi) {
            const double pi = freq[i], pj = freq[j];
            if (pi > 0 && pi < 1 && pj > 0 && pj < 1) {
                const uint64_t* a = planes + i * 4 * nw;
                const uint64_t* b = planes + j * 4 * nw;
                int tot = 0, x = 0;
                for (int w = 0; w < nw; ++w) {
                    tot += __popcll(a[w] & b[w]);
                    const uint64_t a2 = a[nw + w], b2 = b[nw + w], a1 = a[2 * nw + w], b1 = b[2 * nw + w];
                    x += 2 * __popcll(a2 & b2) + __popcll(a1 & b2) + __popcll(a2 & b1) +
                         __popcll(a1 & b1 & ~(a[3 * nw + w] ^ b[3 * nw + w]));
                }
                double x11 = (double)x;
                x11 /= (double)(2 * tot);
                const double Dv = x11 - pi * pj;
                const double R2 = Dv * Dv / (pi * (1 - pi) * pj * (1 - pj));
                v = (R2 > 1) ? 1.0 : R2;
            }
        }
        P[t] = v;
    }
}

// ------------------------------------------------------------------------------------------
// K6c: LD[w][k] = Σ_{i=w}^{w+W-1} (i==w+k ? 1 : hr2(i, w+k)), ascending i from 0.0
// (garlic-data.cpp:489-494, 521-527) = sum of the contiguous slice P[w+k][W-1-k .. 2W-2-k].
// Thread (j, k), w = j-k; lanes = consecutive k of one j read overlapping slices of one P row.
// ------------------------------------------------------------------------------------------
__global__ void ld_sum_kernel(const double* __restrict__ P, const int* __restrict__ chr_of,
                              const int* __restrict__ chr_start, int n_chr, long long L, int W,
                              double* __restrict__ invld, double* __restrict__ ld_out)
{
    const int D = 2 * W - 1;
    const long long total = L * W;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long j = t / W;
        const int k = (int)(t % W);
        const long long w = j - k;
        const int c = chr_of[j];
        const long long lo = chr_start[c], hi = (c + 1 < n_chr) ? chr_start[c + 1] : L;
        if (w < lo || w >= hi - W + 1) continue;
        const double* row = P + j * D + (W - 1 - k);
        double acc = 0.0;
        for (int n = 0; n < W; ++n) acc += row[n];
        invld[w * (W + kInvFront + kInvBack) + kInvFront + k] = 1.0 / acc;
        if (ld_out) ld_out[w * W + k] = acc;
    }
}

// homFreq[d] = #(g in {0,2}) / #(non-missing) over ALL individuals, from the reduced per-SNP counters
// (calculateGenoFreq, garlic-data.cpp:656-676; 0/0 gives NaN exactly as there)
__global__ void hom_freq_kernel(const int* __restrict__ counts, long long L0, const int* __restrict__ src, long long L,
                                double* __restrict__ homf)
{
    for (long long d = blockIdx.x * (long long)blockDim.x + threadIdx.x; d < L; d += (long long)gridDim.x * blockDim.x) {
        const int s = src[d];
        double fh = (double)counts[2 * L0 + s];
        fh /= (double)counts[3 * L0 + s];
        homf[d] = fh;
    }
}
cudaError_t launch_hom_freq(const int* counts, long long L0, const int* src, long long L, double* homf, cudaStream_t st)
{
    if (!L) return cudaSuccess;
    long long b = (L + 255) / 256;
    hom_freq_kernel<<<(unsigned)(b > 148 * 16 ? 148 * 16 : b), 256, 0, st>>>(counts, L0, src, L, homf);
    return cudaGetLastError();
}

size_t ld_planes_words(long long L, int n_ld, bool phased) { return (size_t)L * (phased ? 4 : 2) * ((n_ld + 63) / 64); }
size_t ld_pairs_doubles(long long L, int W) { return (size_t)L * (2 * W - 1); }

cudaError_t launch_ld_band(const uint64_t* geno, int64_t row_words, const int* ld_ind, int n_ld,
                           const double* homf, const int* chr_of, const int* chr_start, int n_chr,
                           long long L, int W, double* invld, double* ld_out, cudaStream_t st, int* n_launches,
                           ncclComm_t comm, int ind_lo, int n_local, uint64_t* planes, double* P, const LdPhase& ph)
{
    *n_launches = 0;
    const int nw = (n_ld + 63) / 64;
    auto blocks = [](long long n) { long long b = (n + 255) / 256; return (unsigned)(b > 148 * 64 ? 148 * 64 : b); };
    const int np = ph.alleles ? 4 : 2;
    ld_planes_kernel<<<blocks(L), 256, 0, st>>>(geno, row_words, ld_ind, n_ld, L, nw, planes, ind_lo, n_local, ph);
    // individuals are sharded over GPUs: every rank sets the bits of the LD individuals it holds, the planes are
    // combined over NVLink (bits are disjoint, so SUM is OR) and each rank then builds the whole band itself
    if (comm) {
        const ncclResult_t nr = ncclAllReduce(planes, planes, (size_t)L * np * nw, ncclUint64, ncclSum, comm, st);
        if (nr != ncclSuccess) return cudaErrorUnknown;
    }
    if (ph.alleles) ld_pairs_r2_kernel<<<blocks(L * (2 * W - 1)), 256, 0, st>>>(planes, nw, ph.freq, chr_of, chr_start, n_chr, L, W, P);
    else ld_pairs_kernel<<<blocks(L * (2 * W - 1)), 256, 0, st>>>(planes, nw, homf, chr_of, chr_start, n_chr, L, W, P);
    if (ld_out) cudaMemsetAsync(ld_out, 0, (size_t)L * W * sizeof(double), st);
    ld_sum_kernel<<<blocks(L * W), 256, 0, st>>>(P, chr_of, chr_start, n_chr, L, W, invld, ld_out);
    *n_launches = 3;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K5-W: weighted windows.  Every window is a fresh, ascending sum
//   wLOD(t) = Σ_k score[t+k] · (1.0 / LD[t][k])        (garlic-roh.cpp:253-273)
// with separate multiply and add (the reference is built without FMA).  Coverage count and
// run-length assembly as in the unweighted walker, in per-step form (their cost is negligible
// next to the W multiply-adds).
// ------------------------------------------------------------------------------------------
template <int SRC, bool ROH, bool DUMP>
__device__ __forceinline__ void wlod_walk_item(const WlodParams& Q, const Item& it, int k_slot, bool active,
                                               uint32_t* ring, int rstride)
{
    const WalkParams& P = Q.base;
    const int W = P.W;
    const int ind = P.ind_list ? P.ind_list[k_slot] : k_slot;
    const uint64_t* row = P.geno + (int64_t)ind * P.row_words;
    const double* glrow = (SRC == 1) ? gl_lane(P, ind) : nullptr;
    const int NW = ((W + 31) >> 5) + 1;
    for (int w = 0; w < NW; ++w) ring[w * rstride] = 0;
    int cov = 0, run_start = -1;
    const int t_end = it.own_hi;
    for (int t = it.w0, q = 0; t < t_end; ++t, ++q) {
        bool f = false;
        if (t < it.we) {
            const double* inv = Q.invld + (int64_t)t * (W + kInvFront + kInvBack) + kInvFront;
            double acc = 0.0;
            for (int k = 0; k < W; ++k) {
                const int s = t + k;
                const int g = (int)(row[s >> 5] >> (2 * (s & 31))) & 3;
                double sc;
                if (SRC == 0) sc = Q.wlut[(int64_t)s * 4 + g];
                else sc = glrow[(int64_t)s * kGlLanes] * Q.nomut[s] * Q.norec[s];
                acc += sc * inv[k];
            }
            f = acc >= P.cutoff;
            dump_window<DUMP>(P, it, k_slot, active, t, acc);
        }
        // flag history ring, bit-addressed by step counter q
        const int wq = (q >> 5) % NW;
        uint32_t cur = (q & 31) ? ring[wq * rstride] : 0u;
        cur |= (uint32_t)f << (q & 31);
        ring[wq * rstride] = cur;
        uint32_t o = 0;
        if (q >= W) o = (ring[(((q - W) >> 5) % NW) * rstride] >> ((q - W) & 31)) & 1u;
        cov += (int)f - (int)o;
        if (ROH && t >= it.own_lo) {
            const bool c = cov >= P.thr;
            if (c && run_start < 0) run_start = t;
            else if (!c && run_start >= 0) { emit_run(P, it, ind, active, run_start, t - 1); run_start = -1; }
        }
    }
    if (ROH && run_start >= 0) emit_run(P, it, ind, active, run_start, it.own_hi - 1);
}

// ------------------------------------------------------------------------------------------
// K5-W fast pass: weighted windows of pass 2 as a banded matrix product on the FP64 tensor cores.
//   Win[i][t] = Σ_m Sc[i][m] · Wt[m][t],   m = s - t0 over the SNPs a tile of 8 windows t0..t0+7 touches,
//   Sc[i][m] = score of individual i at SNP t0+m (wlut[s][g] or gl·nomut·norec), Wt[m][j] = 1/LD[t0+j][m-j] (0 outside).
// One warp owns 32 individuals of one item and works on blocks of 32 windows = 4 tiles of 8.  It walks the block's
// SNPs in quads (the k = 4 of mma.sync.m8n8k4.f64): per quad ONE score fragment per 8-individual row group (A: one
// double per lane, a table lookup through the packed genotype word) is shared by every window tile whose band holds
// the quad, each tile adding one weight fragment (B: one double per lane, read without predicates from zero-padded
// weight rows) — 16 DMMA = 4096 multiply-adds for 8 loads in the steady state, operands loaded one quad ahead.  Which
// tiles hold which quads is a compile-time schedule (ramp-up, steady, ramp-down), so no DMMA is predicated.
// The 8x8 accumulator tiles become window flags (cutoff ± tol), the flag bits are routed by shuffles to the lane that
// owns each individual, and four tiles make the 32-bit flag word that cover_block (walk.cuh) turns into coverage and
// run records — exactly as the unweighted walker does.
// DMMA fuses and reorders the sum, so values differ from the reference's mul-then-add chain in the last bits:
// windows within tol of the cutoff mark their (individual, segment) pair ambiguous, and those pairs are re-walked
// by the exact kernel above (garlic_gpu_call_roh).  Window dumps (KDE, --raw-lod) always use the exact kernel.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int SRC>
__global__ void __launch_bounds__(128)
wlod_mma_kernel(const WlodParams Q, const Item* __restrict__ items, int n_items, int n_groups)
{
    extern __shared__ uint32_t ring_smem[];   // [NW][blockDim.x] flag-word history (W > 32)
    const WalkParams& P = Q.base;
    const int W = P.W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gpb = blockDim.x >> 5;
    const int gblocks = (n_groups + gpb - 1) / gpb;
    const long long total = (long long)n_items * gblocks;
    const int NW = ((W + 31) >> 5) + 1, r = (32 - (W & 31)) & 31;
    uint32_t* ring = ring_smem + threadIdx.x;
    const int rstride = blockDim.x;
    const double cut_hi = P.cutoff + P.tol, cut_lo = P.cutoff - P.tol;
    const int j = lane >> 2, mq = lane & 3;    // window column / SNP-within-k-step of this lane's B element; row j of A
    for (long long u = blockIdx.x; u < total; u += gridDim.x) {
        const int item = (int)(u / gblocks);
        const int group = (int)(u % gblocks) * gpb + warp;
        if (group >= n_groups) continue;
        const Item it = items[item];
        const int k_own = group * 32 + lane;
        const bool active = k_own < P.n_lanes;
        const int ind = P.ind_list ? P.ind_list[active ? k_own : P.n_lanes - 1] : (active ? k_own : P.n_lanes - 1);
        // the four individuals whose scores this lane supplies (row j of each 8-row group)
        const uint64_t* rowA[4];
        const double* glA[4];
#pragma unroll
        for (int rg = 0; rg < 4; ++rg) {
            int k = group * 32 + 8 * rg + j;
            if (k >= P.n_lanes) k = P.n_lanes - 1;
            const int ia = P.ind_list ? P.ind_list[k] : k;
            rowA[rg] = P.geno + (int64_t)ia * P.row_words;
            glA[rg] = (SRC == 1) ? gl_lane(P, ia) : nullptr;
        }
        LaneState S;
        S.win = 0; S.cov = 0; S.run_start = -1; S.fw = 0; S.hist = 0; S.ambig = false;
        if (W > 32) for (int w = 0; w < NW; ++w) ring[w * rstride] = 0;
        int wr = 1 % NW;
        // Quads (4 SNPs) of a block: tile q (windows tb+8q..+7) holds quads 2q .. E+2q, E = (W+6)/4.  Rounded to pairs of
        // quads (E2 odd) the tile sets are compile-time constants: ramp-up {0},{0,1},{0,1,2}, steady {0..3}, ramp-down
        // {1,2,3},{2,3},{3} — no predicated tensor-core instructions; the operand loads need no predicates either
        // because weight rows are zero-padded (wlod.h) and windows past the segment end are masked afterwards.
        const int E2 = ((W + 6) / 4) | 1;
        const int ldw = W + kInvFront + kInvBack;
        // blocks start on a multiple of 4 SNPs (windows before w0 are masked off below): a quad then never straddles a
        // packed 64-bit genotype word and the word reload is a warp-uniform branch taken every 8th quad
        for (int tb = it.w0 & ~3; tb < it.own_hi; tb += 32) {
            uint32_t fhi = 0, flo = 0;
            // c[q][rg]: 8 windows tb+8q.. x 8 individuals 8rg.. (two accumulator columns per lane)
            double c[4][4][2];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int rg = 0; rg < 4; ++rg) { c[q][rg][0] = 0.0; c[q][rg][1] = 0.0; }
            // this lane's weight rows: window tb+8q+j of each tile; element for block-relative SNP m is invq[q][m]
            const double* invq[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) invq[q] = Q.invld + (int64_t)(tb + 8 * q + j) * ldw + kInvFront - (8 * q + j);
            uint64_t gw[4], gwn[4];                                    // packed genotype word in use / the next one, in flight
            {
                const int s0 = tb + mq;
#pragma unroll
                for (int rg = 0; rg < 4; ++rg) { gw[rg] = rowA[rg][s0 >> 5]; gwn[rg] = rowA[rg][(s0 >> 5) + 1]; }
            }
            // operands of quad kq: ONE A fragment (scores) per row group, shared by every tile of the mask, one B
            // fragment (weights) per tile
            auto load_quad = [&](auto mask, int kq, double (&a)[4], double (&b)[4]) {
                constexpr int MB = decltype(mask)::value;
                const int m = 4 * kq + mq, s = tb + m;
                const int sh = 2 * (s & 31);
                if (kq > 0 && ((tb + 4 * kq) & 31) == 0) {
#pragma unroll
                    for (int rg = 0; rg < 4; ++rg) { gw[rg] = gwn[rg]; gwn[rg] = rowA[rg][(s >> 5) + 1]; }
                }
#pragma unroll
                for (int rg = 0; rg < 4; ++rg) {
                    const uint32_t g8 = ((uint32_t)(gw[rg] >> sh) & 3u) * 8u;
                    if (SRC == 0) a[rg] = __ldg(reinterpret_cast<const double*>(reinterpret_cast<const char*>(Q.wlut + (int64_t)s * 4) + g8));
                    else a[rg] = glA[rg][(int64_t)s * kGlLanes] * Q.nomut[s] * Q.norec[s];
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) if ((MB >> q) & 1) b[q] = __ldg(invq[q] + m);
            };
            auto mma_quad = [&](auto mask, const double (&a)[4], const double (&b)[4]) {
                constexpr int MB = decltype(mask)::value;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if ((MB >> q) & 1) {
#pragma unroll
                        for (int rg = 0; rg < 4; ++rg) dmma_m8n8k4(c[q][rg][0], c[q][rg][1], a[rg], b[q]);
                    }
            };
            using M1 = std::integral_constant<int, 0x1>; using M3 = std::integral_constant<int, 0x3>;
            using M7 = std::integral_constant<int, 0x7>; using MF = std::integral_constant<int, 0xF>;
            using ME = std::integral_constant<int, 0xE>; using MC = std::integral_constant<int, 0xC>;
            using M8 = std::integral_constant<int, 0x8>;
            double a0[4], b0[4], a1[4], b1[4];
            // loads run one quad ahead of the tensor-core work
            load_quad(M1(), 0, a0, b0);
            load_quad(M1(), 1, a1, b1); mma_quad(M1(), a0, b0);
            load_quad(M3(), 2, a0, b0); mma_quad(M1(), a1, b1);
            load_quad(M3(), 3, a1, b1); mma_quad(M3(), a0, b0);
            load_quad(M7(), 4, a0, b0); mma_quad(M3(), a1, b1);
            load_quad(M7(), 5, a1, b1); mma_quad(M7(), a0, b0);
            load_quad(MF(), 6, a0, b0); mma_quad(M7(), a1, b1);
            int kq = 6;
#pragma unroll 1
            for (; kq < E2; kq += 2) {
                load_quad(MF(), kq + 1, a1, b1); mma_quad(MF(), a0, b0);
                load_quad(MF(), kq + 2, a0, b0); mma_quad(MF(), a1, b1);
            }
            load_quad(ME(), kq + 1, a1, b1); mma_quad(ME(), a0, b0);
            load_quad(MC(), kq + 2, a0, b0); mma_quad(ME(), a1, b1);
            load_quad(MC(), kq + 3, a1, b1); mma_quad(MC(), a0, b0);
            load_quad(M8(), kq + 4, a0, b0); mma_quad(MC(), a1, b1);
            load_quad(M8(), kq + 5, a1, b1); mma_quad(M8(), a0, b0);
            mma_quad(M8(), a1, b1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                // accumulator (row j, columns 2mq, 2mq+1) → flag bits of windows t0+2mq, t0+2mq+1 of individual 8rg+j
                uint32_t myhi = 0, mylo = 0;
#pragma unroll
                for (int rg = 0; rg < 4; ++rg) {
                    uint32_t bh = ((uint32_t)(c[q][rg][0] >= cut_hi) | ((uint32_t)(c[q][rg][1] >= cut_hi) << 1)) << (2 * mq);
                    uint32_t bl = ((uint32_t)(c[q][rg][0] >= cut_lo) | ((uint32_t)(c[q][rg][1] >= cut_lo) << 1)) << (2 * mq);
                    bh |= __shfl_xor_sync(0xffffffffu, bh, 1); bh |= __shfl_xor_sync(0xffffffffu, bh, 2);
                    bl |= __shfl_xor_sync(0xffffffffu, bl, 1); bl |= __shfl_xor_sync(0xffffffffu, bl, 2);
                    // owner lane o holds individual o = 8·(o/8) + (o%8): its byte sits in lanes 4·(o%8)..+3 of row group o/8
                    const uint32_t gh = __shfl_sync(0xffffffffu, bh, 4 * (lane & 7));
                    const uint32_t gl_ = __shfl_sync(0xffffffffu, bl, 4 * (lane & 7));
                    if (rg == (lane >> 3)) { myhi = gh; mylo = gl_; }
                }
                fhi |= myhi << (8 * q);
                flo |= mylo << (8 * q);
            }
            const int nv = it.we - tb;                                 // valid windows of the block
            uint32_t vm = nv >= 32 ? 0xffffffffu : (nv <= 0 ? 0u : ((1u << nv) - 1u));
            if (tb < it.w0) vm &= ~((1u << (it.w0 - tb)) - 1u);
            fhi &= vm; flo &= vm;
            S.ambig |= (fhi != flo);
            uint32_t ow = 0;
            if (W > 32) {
                int r0 = wr + 1; if (r0 >= NW) r0 -= NW;
                int r1 = r0 + 1; if (r1 >= NW) r1 -= NW;
                const uint32_t w0_ = ring[r0 * rstride], w1_ = ring[r1 * rstride];
                ow = r ? ((w0_ >> r) | (w1_ << (32 - r))) : w0_;
            }
            const bool full = (tb + 31 < it.we) && (tb >= it.own_lo) && (tb + 31 < it.own_hi);
            if (full) cover_block<true>(P, it, S, ind, active, fhi, ow, tb);
            else cover_block<false>(P, it, S, ind, active, fhi, ow, tb);
            if (W > 32) {
                ring[wr * rstride] = fhi;
                if (++wr >= NW) wr = 0;
            }
        }
        if (S.run_start >= 0) emit_run(P, it, ind, active, S.run_start, it.own_hi - 1);
        if (S.ambig && active) {
            const unsigned p = atomicAdd(P.out_count + 1, 1u);
            if (p < P.amb_cap) {
                RohRec rr;
                rr.ind = ind; rr.a = 0; rr.b = 0; rr.tag = it.seg;
                P.amb[p] = rr;
            }
        }
    }
}

cudaError_t launch_wlod_mma(const WlodParams& Q, const Item* items, int n_items, bool gl_mode, cudaStream_t st)
{
    if (n_items == 0 || Q.base.n_lanes == 0) return cudaSuccess;
    const int threads = 128;
    const int n_groups = (Q.base.n_lanes + 31) / 32;
    const int gpb = threads / 32;
    const long long total = (long long)n_items * ((n_groups + gpb - 1) / gpb);
    const int NW = ((Q.base.W + 31) >> 5) + 1;
    const size_t smem = (size_t)NW * threads * sizeof(uint32_t);
    long long grid = total;
    const long long cap = 148ll * 16 * 8;
    if (grid > cap) grid = cap;
    if (smem > 48 * 1024) {                                    // flag history of very large windows
        cudaError_t e = gl_mode ? cudaFuncSetAttribute(wlod_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                : cudaFuncSetAttribute(wlod_mma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if (gl_mode) wlod_mma_kernel<1><<<(unsigned)grid, threads, smem, st>>>(Q, items, n_items, n_groups);
    else wlod_mma_kernel<0><<<(unsigned)grid, threads, smem, st>>>(Q, items, n_items, n_groups);
    return cudaGetLastError();
}

template <int SRC, bool ROH, bool DUMP>
__global__ void __launch_bounds__(128)
wlod_walk_kernel(const WlodParams Q, const Item* __restrict__ items, int n_items, int n_groups)
{
    extern __shared__ uint32_t ring_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gpb = blockDim.x >> 5;
    const int gblocks = (n_groups + gpb - 1) / gpb;
    const long long total = (long long)n_items * gblocks;
    for (long long u = blockIdx.x; u < total; u += gridDim.x) {
        const int item = (int)(u / gblocks);
        const int group = (int)(u % gblocks) * gpb + warp;
        if (group >= n_groups) continue;
        const int k = group * 32 + lane;
        const bool active = k < Q.base.n_lanes;
        const Item it = items[item];
        wlod_walk_item<SRC, ROH, DUMP>(Q, it, active ? k : Q.base.n_lanes - 1, active, ring_smem + threadIdx.x, blockDim.x);
    }
}

template <int SRC, bool ROH, bool DUMP>
static cudaError_t launch_wlod_t(const WlodParams& Q, const Item* items, int n_items, cudaStream_t st)
{
    if (n_items == 0 || Q.base.n_lanes == 0) return cudaSuccess;
    const int threads = 128;
    const int n_groups = (Q.base.n_lanes + 31) / 32;
    const int gpb = threads / 32;
    const long long total = (long long)n_items * ((n_groups + gpb - 1) / gpb);
    const int NW = ((Q.base.W + 31) >> 5) + 1;
    const size_t smem = (size_t)NW * threads * sizeof(uint32_t);
    long long grid = total;
    const long long cap = 148ll * 16 * 8;
    if (grid > cap) grid = cap;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(wlod_walk_kernel<SRC, ROH, DUMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    wlod_walk_kernel<SRC, ROH, DUMP><<<(unsigned)grid, threads, smem, st>>>(Q, items, n_items, n_groups);
    return cudaGetLastError();
}

cudaError_t launch_wlod_walk(const WlodParams& Q, const Item* items, int n_items, bool gl_mode, bool roh,
                             bool dump, cudaStream_t st)
{
    if (gl_mode) {
        if (roh && !dump) return launch_wlod_t<1, true, false>(Q, items, n_items, st);
        if (!roh && dump) return launch_wlod_t<1, false, true>(Q, items, n_items, st);
        return launch_wlod_t<1, true, true>(Q, items, n_items, st);
    }
    if (roh && !dump) return launch_wlod_t<0, true, false>(Q, items, n_items, st);
    if (!roh && dump) return launch_wlod_t<0, false, true>(Q, items, n_items, st);
    return launch_wlod_t<0, true, true>(Q, items, n_items, st);
}

}  // namespace garlic
