// kde.cu — computeKDE (garlic-kde.cpp:14-101) on the device, for the windows pass 1 left in HBM:
// nrd0 bandwidth (gsl_stats_sd, two gsl quantiles: garlic-kde.cpp:130-140), the 512 equally spaced targets from
// min - 3h to max + 3h (:44-66), and the Gauss transform  y[m] = sum_i (1/n) exp(-(t_m - x_i)^2 / h^2)  evaluated
// exactly (FIGTree's kernel convention; the reference asks FIGTree for an eps = 1e-2 approximation of this sum, :82).
// Nothing is sorted or compacted: MISSING / NaN slots are skipped in place, the two quantiles are four order
// statistics found together by an 8-pass radix selection (256-bin histograms of the order-preserving 64-bit keys),
// every floating-point reduction runs in a fixed order (block partials, then one block), so the result is
// reproducible.  One stream-ordered chain, no host round trip; the host only normalises the 512 values and runs the
// reference's mode heuristic on them.
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"
#include "kernels.h"

namespace garlic {

namespace {

constexpr int kKdeBlocks = 148 * 4;       // fixed grid of the reductions (partials are combined in block order)
constexpr int kKdeThreads = 256;
constexpr int kKdeChunk = 1024;           // sources staged per step of the transform

__device__ __forceinline__ bool kde_valid(double v) { return v != kMissing && v == v; }
__device__ __forceinline__ unsigned long long kde_key(double v)      // order-preserving: a < b <=> key(a) < key(b)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double kde_unkey(unsigned long long k)
{
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ double block_sum(double v, double* sh)     // fixed order: lanes, then warps
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    return t;                                                         // valid in thread 0
}

// pass A: count, sum, min, max of the valid values — per-block partials
__global__ void __launch_bounds__(kKdeThreads)
kde_moments_kernel(const double* __restrict__ v, long long n, double* __restrict__ part)
{
    __shared__ double sh[kKdeThreads / 32];
    double s = 0.0, c = 0.0, mn = 1e300, mx = -1e300;
    for (long long i = blockIdx.x * (long long)kKdeThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kKdeThreads) {
        const double x = v[i];
        if (kde_valid(x)) { s += x; c += 1.0; mn = x < mn ? x : mn; mx = x > mx ? x : mx; }
    }
    const double ts = block_sum(s, sh), tc = block_sum(c, sh);
#pragma unroll
    for (int o = 16; o; o >>= 1) { const double a = __shfl_down_sync(0xffffffffu, mn, o), b = __shfl_down_sync(0xffffffffu, mx, o); mn = a < mn ? a : mn; mx = b > mx ? b : mx; }
    __shared__ double smn[kKdeThreads / 32], smx[kKdeThreads / 32];
    if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kKdeThreads / 32; ++w) { mn = smn[w] < mn ? smn[w] : mn; mx = smx[w] > mx ? smx[w] : mx; }
        part[4 * blockIdx.x] = tc; part[4 * blockIdx.x + 1] = ts; part[4 * blockIdx.x + 2] = mn; part[4 * blockIdx.x + 3] = mx;
    }
}

// state[]: 0 n, 1 mean, 2 min, 3 max, 4 sum of squared deviations, 5 h, 6 grid lo, 7 grid hi; sel[]: ranks / prefixes
struct KdeSel { unsigned long long rank[4], prefix[4]; unsigned hist[4][256]; };

__global__ void kde_prepare_kernel(const double* __restrict__ part, int n_part, double* __restrict__ state, KdeSel* __restrict__ sel)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double c = 0.0, s = 0.0, mn = 1e300, mx = -1e300;
        for (int b = 0; b < n_part; ++b) { c += part[4 * b]; s += part[4 * b + 1]; mn = part[4 * b + 2] < mn ? part[4 * b + 2] : mn; mx = part[4 * b + 3] > mx ? part[4 * b + 3] : mx; }
        state[0] = c; state[1] = c > 0 ? s / c : 0.0; state[2] = mn; state[3] = mx;
        // gsl_stats_quantile_from_sorted_data: index = f (n - 1), lhs = floor, the pair (lhs, lhs + 1)
        const unsigned long long n = (unsigned long long)c;
        const double fs[2] = {0.25, 0.75};
        for (int k = 0; k < 2; ++k) {
            const double index = n ? fs[k] * (double)(n - 1) : 0.0;
            const unsigned long long lhs = (unsigned long long)index;
            sel->rank[2 * k] = lhs;
            sel->rank[2 * k + 1] = (n && lhs + 1 < n) ? lhs + 1 : lhs;
        }
        for (int k = 0; k < 4; ++k) sel->prefix[k] = 0ull;
    }
    for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) (&sel->hist[0][0])[i] = 0u;
}

// pass B: sum of squared deviations from the mean (gsl_stats_sd: variance with the N - 1 denominator)
__global__ void __launch_bounds__(kKdeThreads)
kde_ssq_kernel(const double* __restrict__ v, long long n, const double* __restrict__ state, double* __restrict__ part)
{
    __shared__ double sh[kKdeThreads / 32];
    const double mean = state[1];
    double s = 0.0;
    for (long long i = blockIdx.x * (long long)kKdeThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kKdeThreads) {
        const double x = v[i];
        if (kde_valid(x)) { const double d = x - mean; s += d * d; }
    }
    const double t = block_sum(s, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = t;
}

// radix selection, one digit (8 bits from `shift`) of the four order statistics at once
__global__ void __launch_bounds__(kKdeThreads)
kde_hist_kernel(const double* __restrict__ v, long long n, int shift, KdeSel* __restrict__ sel)
{
    __shared__ unsigned sh[4][256];
    for (int i = threadIdx.x; i < 4 * 256; i += kKdeThreads) (&sh[0][0])[i] = 0u;
    unsigned long long pre[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) pre[k] = sel->prefix[k];
    __syncthreads();
    const bool first = shift == 56;
    for (long long i = blockIdx.x * (long long)kKdeThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kKdeThreads) {
        const double x = v[i];
        if (!kde_valid(x)) continue;
        const unsigned long long key = kde_key(x);
        const unsigned d = (unsigned)(key >> shift) & 255u;
        const unsigned long long hi = first ? 0ull : (key >> (shift + 8));
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (first || hi == (pre[k] >> (shift + 8))) atomicAdd(&sh[k][d], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * 256; i += kKdeThreads) {
        const unsigned c = (&sh[0][0])[i];
        if (c) atomicAdd(&sel->hist[0][0] + i, c);
    }
}

__global__ void kde_pick_kernel(int shift, KdeSel* __restrict__ sel)
{
    const int k = threadIdx.x;
    if (k < 4) {
        unsigned long long r = sel->rank[k], acc = 0ull;
        int b = 0;
        for (; b < 255; ++b) { const unsigned long long c = sel->hist[k][b]; if (acc + c > r) break; acc += c; }
        sel->rank[k] = r - acc;
        sel->prefix[k] |= (unsigned long long)b << shift;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) (&sel->hist[0][0])[i] = 0u;
}

// nrd0 (garlic-kde.cpp:130-140) and the target grid (:44-66)
__global__ void kde_bandwidth_kernel(const double* __restrict__ part, int n_part, const KdeSel* __restrict__ sel, double* __restrict__ state,
                                     double* __restrict__ x, int m)
{
    __shared__ double s_lo, s_hi;
    if (threadIdx.x == 0) {
        double ssq = 0.0;
        for (int b = 0; b < n_part; ++b) ssq += part[b];
        const double n = state[0];
        const double sd = sqrt(ssq / n * (n / (n - 1.0)));              // gsl_stats_sd
        double q[2];
        for (int k = 0; k < 2; ++k) {
            const double index = (k ? 0.75 : 0.25) * (n - 1.0);
            const double lhs = floor(index), delta = index - lhs;
            const double a = kde_unkey(sel->prefix[2 * k]), b = kde_unkey(sel->prefix[2 * k + 1]);
            q[k] = (lhs == n - 1.0) ? a : (1.0 - delta) * a + delta * b;    // gsl_stats_quantile_from_sorted_data
        }
        const double iqr = q[1] - q[0];
        const double lo = sd < iqr / 1.34 ? sd : iqr / 1.34;
        const double h = 0.9 * lo * pow(n, -0.2);
        state[4] = ssq; state[5] = h;
        s_lo = state[2] - 3.0 * h; s_hi = state[3] + 3.0 * h;
        state[6] = s_lo; state[7] = s_hi;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) x[i] = ((double)(i + 1) / (double)m) * (s_hi - s_lo) + s_lo;
}

// the transform: a thread owns one target, the block streams its share of the sources through shared memory
__global__ void __launch_bounds__(512)
kde_gauss_kernel(const double* __restrict__ v, long long n, const double* __restrict__ state, const double* __restrict__ x, int m,
                 double* __restrict__ part)
{
    __shared__ double src[kKdeChunk];
    const double hsq = state[5] * state[5];
    const int t0 = threadIdx.x;
    double acc[2] = {0.0, 0.0};                                          // m <= 1024: at most two targets per thread
    const double tx0 = t0 < m ? x[t0] : 0.0, tx1 = t0 + 512 < m ? x[t0 + 512] : 0.0;
    const long long n_chunks = (n + kKdeChunk - 1) / kKdeChunk;
    for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        __syncthreads();
        for (int j = threadIdx.x; j < kKdeChunk; j += 512) {
            const long long i = c * kKdeChunk + j;
            const double s = i < n ? v[i] : kMissing;
            src[j] = kde_valid(s) ? s : __longlong_as_double(0x7ff8000000000000ll);
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < kKdeChunk; ++j) {
            const double s = src[j];
            if (s != s) continue;                                        // warp-uniform: every thread reads the same source
            const double d0 = tx0 - s, d1 = tx1 - s;
            acc[0] += exp(-(d0 * d0) / hsq);
            if (m > 512) acc[1] += exp(-(d1 * d1) / hsq);
        }
    }
    if (t0 < m) part[(long long)blockIdx.x * m + t0] = acc[0];
    if (t0 + 512 < m) part[(long long)blockIdx.x * m + t0 + 512] = acc[1];
}

__global__ void kde_finish_kernel(const double* __restrict__ part, int n_part, int m, const double* __restrict__ state, double* __restrict__ y)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    double s = 0.0;
    for (int b = 0; b < n_part; ++b) s += part[(long long)b * m + t];
    y[t] = s * (1.0 / state[0]);                                         // q_i = 1/n (garlic-kde.cpp:73-77)
}

}  // namespace

size_t kde_scratch_doubles(int m)
{
    // partials of the moments (4 per block) | partials of the squared deviations | state[8] | x[m] | y[m] | transform partials | selection state
    return (size_t)4 * kKdeBlocks + kKdeBlocks + 8 + 2 * (size_t)m + (size_t)kKdeBlocks * m + (sizeof(KdeSel) + 7) / 8;
}

// v[n] on the device (MISSING / NaN entries are skipped).  out (device, 8 + 2 m doubles): state[8] (n, mean, min, max,
// ssq, h, grid lo, grid hi), x[m], y[m] (weights 1/n applied, not normalised).  launches: kernels enqueued.
cudaError_t launch_kde(const double* v, long long n, int m, double* scratch, double** out, int* launches, cudaStream_t st)
{
    if (m < 2 || m > 1024) return cudaErrorInvalidValue;
    double* part4 = scratch;
    double* part1 = part4 + 4 * kKdeBlocks;
    double* state = part1 + kKdeBlocks;
    double* x = state + 8;
    double* y = x + m;
    double* gpart = y + m;
    KdeSel* sel = reinterpret_cast<KdeSel*>(gpart + (size_t)kKdeBlocks * m);
    *out = state;
    int nl = 0;
    kde_moments_kernel<<<kKdeBlocks, kKdeThreads, 0, st>>>(v, n, part4); ++nl;
    kde_prepare_kernel<<<1, 256, 0, st>>>(part4, kKdeBlocks, state, sel); ++nl;
    kde_ssq_kernel<<<kKdeBlocks, kKdeThreads, 0, st>>>(v, n, state, part1); ++nl;
    for (int shift = 56; shift >= 0; shift -= 8) {
        kde_hist_kernel<<<kKdeBlocks, kKdeThreads, 0, st>>>(v, n, shift, sel); ++nl;
        kde_pick_kernel<<<1, 256, 0, st>>>(shift, sel); ++nl;
    }
    kde_bandwidth_kernel<<<1, 256, 0, st>>>(part1, kKdeBlocks, sel, state, x, m); ++nl;
    kde_gauss_kernel<<<kKdeBlocks, 512, 0, st>>>(v, n, state, x, m, gpart); ++nl;
    kde_finish_kernel<<<(m + 255) / 256, 256, 0, st>>>(gpart, kKdeBlocks, m, state, y); ++nl;
    if (launches) *launches = nl;
    return cudaGetLastError();
}

}  // namespace garlic
