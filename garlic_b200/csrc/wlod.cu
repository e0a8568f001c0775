// wlod.cu — the --weighted path: K6 LD band (hr², reference src/garlic-data.cpp:377-424,474-527,
// 558-583) and K5-W weighted windows fused with ROH assembly (src/garlic-roh.cpp:204-277,409-546).
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include <algorithm>
#include <cstdlib>
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
        invld[w * inv_stride(W) + kInvFront + k] = 1.0 / acc;
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

// ------------------------------------------------------------------------------------------
// K6, fused: pair values and window sums of a tile of 32 SNPs j in one CTA, nothing but the weight rows written.
// LD[w][k] for w + k = j only needs row j of the ordered pair matrix (P[j][d] = pair(i = j + d - (W-1), j)), so the row
// never has to leave the SM:
//   stage   the bit-planes of SNPs [j0 - (W-1), j0 + 31 + (W-1)] (contiguous in global memory) and their homFreq /
//           frequencies go to shared memory, plane rows padded by one word (consecutive SNPs = consecutive lanes:
//           conflict-free 8-byte reads);
//   phase 1 a warp takes a row j, keeps j's plane words in registers, lanes take d = lane, lane + 32, …: AND + POPC over
//           the words of SNP i, then hr2 (garlic-data.cpp:558-583) or, --phased, r2 (:585-617) in the reference's operation
//           order → P tile in shared memory (row stride 2W-1 doubles, odd);
//   phase 2 lanes = rows, a thread forms four neighbouring window sums LD[j-k][k], k = 4g … 4g+3, from one pass over the
//           W+3 entries they share — each sum ascending from 0.0 as the reference adds (:489-494, 521-527);
//   write   the reciprocals leave through shared memory so that every weight row receives its 32 consecutive k at once.
// The popcounts are the only part a tensor-core contraction could replace: they are about a quarter
// of this kernel's instructions, the rest is fp64 division and addition in a fixed order.
// ------------------------------------------------------------------------------------------
constexpr int kLdRows = 32;            // SNPs j per CTA
constexpr int kLdMaxWords = 8;         // plane words per SNP held in registers: n_ld <= 512

struct LdFusedParams {
    const uint64_t* planes; int nw;    // [L][np * nw]
    const double* hf;                  // homFreq[L] (hr2) or freq[L] (r2)
    const int* chr_of; const int* chr_start; int n_chr;
    long long L; int W;
    double* invld; double* ld_out;
};

// shared-memory layout: [planes | homFreq] and, aliasing them once phase 1 is over, the result tile [32][W + 1]; then P
__host__ __device__ inline size_t ld_fused_p_offset(int W, int nw, int np)
{
    const size_t span = kLdRows + 2 * (size_t)(W - 1);
    const size_t a = span * (size_t)(np * nw + 1) * 8 + (span + (span & 1)) * 8;
    const size_t b = (size_t)kLdRows * (W + 1) * 8;
    return ((a > b ? a : b) + 15) & ~(size_t)15;
}

template <bool PHASED>
__global__ void __launch_bounds__(256)
ld_band_fused_kernel(const LdFusedParams Q)
{
    extern __shared__ __align__(16) unsigned char ld_smem[];
    constexpr int NP = PHASED ? 4 : 2;
    const int W = Q.W, D = 2 * W - 1, nw = Q.nw;
    const int span = kLdRows + 2 * (W - 1);                 // SNPs whose planes the tile touches
    const int rw = NP * nw + 1;                             // padded plane row, in words
    uint64_t* s_pl = reinterpret_cast<uint64_t*>(ld_smem);                       // [span][rw]
    double* s_hf = reinterpret_cast<double*>(s_pl + (size_t)span * rw);           // [span]
    double* s_P = reinterpret_cast<double*>(ld_smem + ld_fused_p_offset(W, nw, NP));  // [32][D]
    double* s_R = reinterpret_cast<double*>(ld_smem);                             // [32][W + 1], aliases the planes after phase 1
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long n_tiles = (Q.L + kLdRows - 1) / kLdRows;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long j0 = tile * kLdRows, s_lo = j0 - (W - 1);
        __syncthreads();                                    // the previous tile's results have been written out
        // ---- stage
        for (long long e = threadIdx.x; e < (long long)span * NP * nw; e += 256) {
            const int r = (int)(e / (NP * nw)), c = (int)(e % (NP * nw));
            const long long s = s_lo + r;
            s_pl[(size_t)r * rw + c] = (s >= 0 && s < Q.L) ? Q.planes[s * NP * nw + c] : 0ull;
        }
        for (int r = threadIdx.x; r < span; r += 256) {
            const long long s = s_lo + r;
            s_hf[r] = (s >= 0 && s < Q.L) ? Q.hf[s] : 0.0;
        }
        __syncthreads();
        // ---- phase 1: P[jj][d]
        for (int jj = warp; jj < kLdRows; jj += 8) {
            const long long j = j0 + jj;
            if (j >= Q.L) break;                            // (warp-uniform)
            const int rj = jj + (W - 1);
            uint64_t bj[NP * kLdMaxWords];
#pragma unroll
            for (int c = 0; c < NP * kLdMaxWords; ++c) bj[c] = (c % kLdMaxWords) < nw ? s_pl[(size_t)rj * rw + (c / kLdMaxWords) * nw + (c % kLdMaxWords)] : 0ull;
            const int cj = Q.chr_of[j];
            const long long lo = Q.chr_start[cj], hi = (cj + 1 < Q.n_chr) ? Q.chr_start[cj + 1] : Q.L;
            const double HB = s_hf[rj];
            for (int d = lane; d < D; d += 32) {
                const long long i = j + d - (W - 1);
                const int ri = jj + d;
                double v = 0.0;
                if (i == j) v = 1.0;
                else if (i >= lo && i < hi) {
                    const double HA = s_hf[ri];
                    if (HA > 0 && HA < 1 && HB > 0 && HB < 1) {
                        const uint64_t* a = s_pl + (size_t)ri * rw;
                        int tot = 0, x = 0;
#pragma unroll
                        for (int w = 0; w < kLdMaxWords; ++w) {
                            if (w < nw) {
                                tot += __popcll(a[w] & bj[w]);
                                if (PHASED) {
                                    const uint64_t a2 = a[nw + w], b2 = bj[kLdMaxWords + w], a1 = a[2 * nw + w], b1 = bj[2 * kLdMaxWords + w];
                                    x += 2 * __popcll(a2 & b2) + __popcll(a1 & b2) + __popcll(a2 & b1) +
                                         __popcll(a1 & b1 & ~(a[3 * nw + w] ^ bj[3 * kLdMaxWords + w]));
                                } else {
                                    x += __popcll(a[nw + w] & bj[kLdMaxWords + w]);
                                }
                            }
                        }
                        if (PHASED) {                       // r2(), garlic-data.cpp:585-617 (A = SNP i, B = SNP j)
                            double x11 = (double)x;
                            x11 /= (double)(2 * tot);
                            const double Dv = x11 - HA * HB;
                            const double R2 = Dv * Dv / (HA * (1 - HA) * HB * (1 - HB));
                            v = (R2 > 1) ? 1.0 : R2;
                        } else {                            // hr2(), garlic-data.cpp:558-583
                            double HAB = (double)x, total_d = (double)tot;
                            HAB /= total_d;
                            const double H = HAB - HA * HB;
                            const double HR2 = H * H / (HA * (1 - HA) * HB * (1 - HB));
                            v = (HR2 > 1) ? 1.0 : HR2;
                        }
                    }
                }
                s_P[(size_t)jj * D + d] = v;
            }
        }
        __syncthreads();                                    // P complete; the planes are dead from here on
        // ---- phase 2: window sums, lanes = rows
        {
            const int jj = lane;
            const long long j = j0 + jj;
            const bool row_on = j < Q.L;
            long long lo = 0, hi = 0;
            if (row_on) { const int cj = Q.chr_of[j]; lo = Q.chr_start[cj]; hi = (cj + 1 < Q.n_chr) ? Q.chr_start[cj + 1] : Q.L; }
            const double* prow = s_P + (size_t)jj * D;
            const int n_groups = (W + 3) >> 2;
            for (int g = warp; g < n_groups; g += 8) {
                const int k0 = 4 * g;                       // k0 … k0+3; the slice of k starts at W-1-k
                const int base = W - 1 - (k0 + 3);          // first entry any of the four slices reads (may be < 0 for k > W-1)
                double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
                if (row_on) {
                    for (int n = 0; n < W + 3; ++n) {
                        const int idx = base + n;
                        const double v = (idx >= 0 && idx < D) ? prow[idx] : 0.0;
                        if (n >= 3 && n < 3 + W) acc0 += v;     // k0:     entries base+3 … base+3+W-1
                        if (n >= 2 && n < 2 + W) acc1 += v;     // k0 + 1
                        if (n >= 1 && n < 1 + W) acc2 += v;     // k0 + 2
                        if (n < W) acc3 += v;                   // k0 + 3
                    }
                }
                double* r = s_R + (size_t)jj * (W + 1) + k0;
                r[0] = acc0;
                if (k0 + 1 < W) r[1] = acc1;
                if (k0 + 2 < W) r[2] = acc2;
                if (k0 + 3 < W) r[3] = acc3;
            }
            (void)lo; (void)hi;
        }
        __syncthreads();
        // ---- write: weight row w receives k = j - w for the tile's j: consecutive k, consecutive lanes
        {
            const int ldw = inv_stride(W);
            const int n_w = kLdRows + W - 1;                // rows w = j0 - (W-1) … j0 + 31
            for (int wi = warp; wi < n_w; wi += 8) {
                const long long w = j0 - (W - 1) + wi;
                if (w < 0 || w >= Q.L) continue;
                const int cw = Q.chr_of[w];
                const long long lo = Q.chr_start[cw], hi = (cw + 1 < Q.n_chr) ? Q.chr_start[cw + 1] : Q.L;
                if (w < lo || w >= hi - W + 1) continue;    // no window starts here (the row stays zero)
                const int jj = lane;                        // j = j0 + jj, k = j - w
                const long long k = j0 + jj - w;
                if (k >= 0 && k < W && j0 + jj < Q.L) {
                    const double acc = s_R[(size_t)jj * (W + 1) + k];
                    Q.invld[w * ldw + kInvFront + k] = 1.0 / acc;
                    if (Q.ld_out) Q.ld_out[w * W + k] = acc;
                }
            }
        }
    }
}

static size_t ld_fused_smem(int W, int nw, bool phased)
{
    return ld_fused_p_offset(W, nw, phased ? 4 : 2) + (size_t)kLdRows * (2 * W - 1) * 8;
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
    // fused pair values + window sums while a tile's rows fit in shared memory (n_ld <= 512 and moderate W); the ordered pair
    // matrix in global memory (P) is the general form
    const size_t fused_smem = ld_fused_smem(W, nw, ph.alleles != nullptr);
    if (nw <= kLdMaxWords && fused_smem <= 200 * 1024 && getenv("GARLIC_LD_UNFUSED") == nullptr) {
        LdFusedParams Q;
        Q.planes = planes; Q.nw = nw; Q.hf = ph.alleles ? ph.freq : homf; Q.chr_of = chr_of; Q.chr_start = chr_start; Q.n_chr = n_chr;
        Q.L = L; Q.W = W; Q.invld = invld; Q.ld_out = ld_out;
        if (ld_out) cudaMemsetAsync(ld_out, 0, (size_t)L * W * sizeof(double), st);
        const long long n_tiles = (L + kLdRows - 1) / kLdRows;
        const unsigned grid = (unsigned)std::min<long long>(n_tiles, 148ll * 16);
        cudaError_t e;
        if (ph.alleles) {
            e = cudaFuncSetAttribute(ld_band_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem);
            if (e != cudaSuccess) return e;
            ld_band_fused_kernel<true><<<grid, 256, fused_smem, st>>>(Q);
        } else {
            e = cudaFuncSetAttribute(ld_band_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem);
            if (e != cudaSuccess) return e;
            ld_band_fused_kernel<false><<<grid, 256, fused_smem, st>>>(Q);
        }
        *n_launches = 2;
        return cudaGetLastError();
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
            const double* inv = Q.invld + (int64_t)t * inv_stride(W) + kInvFront;
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

template <int SRC, int MINB>
__global__ void __launch_bounds__(128, MINB)
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
        const int ldw = inv_stride(W);
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

// ------------------------------------------------------------------------------------------
// K5-W fast pass, operands through shared memory (table mode, moderate window sizes).
// Same banded product, same tile schedule, same flag / coverage logic as wlod_mma_kernel above; what changes is where
// the DMMA operands come from.  There every quad cost eight loads through L1 whose addresses the compiler rebuilt from
// scratch (≈ 125 instructions per 16 DMMA, 4.4 long-scoreboard stall cycles per issue: the tensor pipe sat at 60 %).
// Here the four warps of a CTA — four individual groups of the same item — walk the item's blocks of 32 windows
// together, and per block ONE pair of bulk copies (cp.async.bulk, completing on an mbarrier) stages
//   * the block's 32 weight rows (contiguous in global memory: 32 · inv_stride(W) doubles) and
//   * the score-table entries of the SNPs the block touches ((W + 40) · 32 bytes)
// into one of two buffers, the next block's copies in flight while this one is multiplied.  A fragments are then one
// LDS.64 at  table + 32·m + 8·g  (conflict-free: a quad's four SNPs own eight banks each, equal genotypes broadcast), B
// fragments one LDS.64 at  row·(stride − 1) + m, both one quad ahead of the DMMAs that consume them.
// ------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ uint32_t wm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wm_mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wm_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void wm_mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wm_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wm_bulk_load(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(wm_smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(wm_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wm_mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WM_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WM_WAIT_DONE;\n"
        "bra WM_WAIT_LOOP;\n"
        "WM_WAIT_DONE:\n"
        "}\n" ::"r"(wm_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ double wm_lds(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
}  // namespace

__global__ void __launch_bounds__(128, 3)
wlod_mma_smem_kernel(const WlodParams Q, const Item* __restrict__ items, int n_items, int n_groups, int wt_bytes, int tab_bytes)
{
    extern __shared__ __align__(128) unsigned char wm_smem[];
    const WalkParams& P = Q.base;
    const int W = P.W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gblocks = (n_groups + 3) >> 2;
    const long long total = (long long)n_items * gblocks;
    const int NW = ((W + 31) >> 5) + 1, r = (32 - (W & 31)) & 31;
    const int buf_bytes = wt_bytes + tab_bytes;
    uint32_t* ring = reinterpret_cast<uint32_t*>(wm_smem + 2 * buf_bytes) + threadIdx.x;
    const int rstride = 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(wm_smem + 2 * buf_bytes + (size_t)NW * 128 * sizeof(uint32_t));
    const uint32_t buf_u32 = wm_smem_u32(wm_smem);
    const double cut_hi = P.cutoff + P.tol, cut_lo = P.cutoff - P.tol;
    const int j = lane >> 2, mq = lane & 3;    // window column / SNP-within-k-step of this lane's B element; row j of A
    const int ldw = inv_stride(W);
    if (threadIdx.x == 0) { wm_mbar_init(bar, 1); wm_mbar_init(bar + 1, 1); }
    __syncthreads();
    uint32_t ph0 = 0, ph1 = 0;
    // byte offsets inside a buffer: this lane's weight element of tile q at block-relative SNP m = 4 kq + mq sits at
    // boff[q] + 32 kq; its table entry at wt_bytes + 32 mq + 128 kq (+ 8 g)
    uint32_t boff[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) boff[q] = (uint32_t)(((8 * q + j) * (ldw - 1) + kInvFront + mq) * 8);
    const uint32_t toff = (uint32_t)(wt_bytes + 32 * mq);
    for (long long u = blockIdx.x; u < total; u += gridDim.x) {
        const int item = (int)(u / gblocks);
        const int group_raw = (int)(u % gblocks) * 4 + warp;
        const bool warp_on = group_raw < n_groups;                 // a warp without a group runs along (barriers) and emits nothing
        const int group = warp_on ? group_raw : n_groups - 1;
        const Item it = items[item];
        const int k_own = group * 32 + lane;
        const bool active = warp_on && k_own < P.n_lanes;
        const int ind = P.ind_list ? P.ind_list[k_own < P.n_lanes ? k_own : P.n_lanes - 1] : (k_own < P.n_lanes ? k_own : P.n_lanes - 1);
        const uint64_t* rowA[4];
#pragma unroll
        for (int rg = 0; rg < 4; ++rg) {
            int k = group * 32 + 8 * rg + j;
            if (k >= P.n_lanes) k = P.n_lanes - 1;
            const int ia = P.ind_list ? P.ind_list[k] : k;
            rowA[rg] = P.geno + (int64_t)ia * P.row_words;
        }
        LaneState S;
        S.win = 0; S.cov = 0; S.run_start = -1; S.fw = 0; S.hist = 0; S.ambig = false;
        if (W > 32) for (int w = 0; w < NW; ++w) ring[w * rstride] = 0;
        int wr = 1 % NW;
        const int E2 = ((W + 6) / 4) | 1;
        const int tb0 = it.w0 & ~3;
        const int n_blk = (it.own_hi - tb0 + 31) >> 5;
        auto issue = [&](int b) {                                  // thread 0: both copies of block b into buffer b & 1
            const int tbb = tb0 + 32 * b;
            unsigned char* d = wm_smem + (size_t)(b & 1) * buf_bytes;
            wm_mbar_expect_tx(bar + (b & 1), (uint32_t)buf_bytes);
            wm_bulk_load(d, Q.invld + (int64_t)tbb * ldw, (uint32_t)wt_bytes, bar + (b & 1));
            wm_bulk_load(d + wt_bytes, Q.wlut + (int64_t)tbb * 4, (uint32_t)tab_bytes, bar + (b & 1));
        };
        if (threadIdx.x == 0) { issue(0); if (n_blk > 1) issue(1); }
#pragma unroll 1
        for (int b = 0; b < n_blk; ++b) {
            const int tb = tb0 + 32 * b;
            if (b & 1) { wm_mbar_wait(bar + 1, ph1); ph1 ^= 1u; } else { wm_mbar_wait(bar, ph0); ph0 ^= 1u; }
            const uint32_t base = buf_u32 + (uint32_t)((b & 1) * buf_bytes);
            uint32_t fhi = 0, flo = 0;
            double c[4][4][2];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int rg = 0; rg < 4; ++rg) { c[q][rg][0] = 0.0; c[q][rg][1] = 0.0; }
            uint64_t gw[4], gwn[4];                                    // packed genotype word in use / the next one, in flight
            {
                const int s0 = tb + mq;
#pragma unroll
                for (int rg = 0; rg < 4; ++rg) { gw[rg] = rowA[rg][s0 >> 5]; gwn[rg] = rowA[rg][(s0 >> 5) + 1]; }
            }
            const uint32_t tbase = base + toff;
            auto load_quad = [&](auto mask, int kq, double (&a)[4], double (&bf)[4]) {
                constexpr int MB = decltype(mask)::value;
                const int s = tb + 4 * kq + mq;
                const int sh = 2 * (s & 31);
                if (kq > 0 && ((tb + 4 * kq) & 31) == 0) {
#pragma unroll
                    for (int rg = 0; rg < 4; ++rg) { gw[rg] = gwn[rg]; gwn[rg] = rowA[rg][(s >> 5) + 1]; }
                }
                const uint32_t ta = tbase + (uint32_t)(128 * kq);
#pragma unroll
                for (int rg = 0; rg < 4; ++rg) a[rg] = wm_lds(ta + (((uint32_t)(gw[rg] >> sh) & 3u) << 3));
#pragma unroll
                for (int q = 0; q < 4; ++q) if ((MB >> q) & 1) bf[q] = wm_lds(base + boff[q] + (uint32_t)(32 * kq));
            };
            auto mma_quad = [&](auto mask, const double (&a)[4], const double (&bf)[4]) {
                constexpr int MB = decltype(mask)::value;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if ((MB >> q) & 1) {
#pragma unroll
                        for (int rg = 0; rg < 4; ++rg) dmma_m8n8k4(c[q][rg][0], c[q][rg][1], a[rg], bf[q]);
                    }
            };
            using M1 = std::integral_constant<int, 0x1>; using M3 = std::integral_constant<int, 0x3>;
            using M7 = std::integral_constant<int, 0x7>; using MF = std::integral_constant<int, 0xF>;
            using ME = std::integral_constant<int, 0xE>; using MC = std::integral_constant<int, 0xC>;
            using M8 = std::integral_constant<int, 0x8>;
            double a0[4], b0[4], a1[4], b1[4];
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
                uint32_t myhi = 0, mylo = 0;
#pragma unroll
                for (int rg = 0; rg < 4; ++rg) {
                    uint32_t bh = ((uint32_t)(c[q][rg][0] >= cut_hi) | ((uint32_t)(c[q][rg][1] >= cut_hi) << 1)) << (2 * mq);
                    uint32_t bl = ((uint32_t)(c[q][rg][0] >= cut_lo) | ((uint32_t)(c[q][rg][1] >= cut_lo) << 1)) << (2 * mq);
                    bh |= __shfl_xor_sync(0xffffffffu, bh, 1); bh |= __shfl_xor_sync(0xffffffffu, bh, 2);
                    bl |= __shfl_xor_sync(0xffffffffu, bl, 1); bl |= __shfl_xor_sync(0xffffffffu, bl, 2);
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
            __syncthreads();                                       // every warp is done with this buffer …
            if (threadIdx.x == 0 && b + 2 < n_blk) issue(b + 2);  // … so the block after next may land in it
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
    // table mode and moderate windows: operands staged through shared memory (two buffers of 32 weight rows + the block's
    // table entries per CTA; three CTAs per SM need <= 75 KB each)
    if (!gl_mode && getenv("GARLIC_MMA_GLOBAL") == nullptr) {
        const int ldw = inv_stride(Q.base.W);
        const int wt_bytes = 32 * ldw * 8, tab_bytes = ((Q.base.W + 40 + 3) & ~3) * 32;
        const size_t smem_s = 2 * (size_t)(wt_bytes + tab_bytes) + (size_t)NW * 128 * sizeof(uint32_t) + 16;
        if (smem_s <= 75 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(wlod_mma_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s);
            if (e != cudaSuccess) return e;
            wlod_mma_smem_kernel<<<(unsigned)grid, 128, smem_s, st>>>(Q, items, n_items, n_groups, wt_bytes, tab_bytes);
            return cudaGetLastError();
        }
    }
    // MINB = 4 caps the kernel at 128 registers (four CTAs = sixteen warps per SM instead of twelve) at the price of a
    // few spilled values outside the quad loop; GARLIC_MMA_MINB=3 selects the uncapped build for A/B timing
    static const bool lb4 = []() { const char* e = getenv("GARLIC_MMA_MINB"); return !(e && atoi(e) == 3); }();
    auto launch = [&](auto kern) -> cudaError_t {
        if (smem > 48 * 1024) {                                // flag history of very large windows
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kern<<<(unsigned)grid, threads, smem, st>>>(Q, items, n_items, n_groups);
        return cudaGetLastError();
    };
    if (gl_mode) return lb4 ? launch(wlod_mma_kernel<1, 4>) : launch(wlod_mma_kernel<1, 3>);
    return lb4 ? launch(wlod_mma_kernel<0, 4>) : launch(wlod_mma_kernel<0, 3>);
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
