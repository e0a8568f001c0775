// xchg.cu — the path's one exchange on the critical path (the per-SNP allele / total counters of all shards,
// summed before freq and the monomorphic filter, garlic-data.cpp:141) as ONE kernel over NVLink peer memory, fused with
// the freq + keep evaluation that consumes the sums (freq_keep_kernel), instead of ncclAllReduce + a second launch.
//
// Every rank keeps [flags | counters | freq | keep] in one device allocation whose IPC handle the ranks exchange once
// (capi.cu:ensure_xchg), so each rank holds device pointers into every peer's block.  Rank r owns the SNP slice
// [r L0 / N, (r+1) L0 / N):
//   A  (all CTAs) rank r tells every peer "my counters of call #seq are final" — the kernel is stream-ordered behind the
//      count kernel — and waits until every peer has said so;
//   B  for its slice it loads the two counters of all N ranks through the peer pointers (system-scope loads: L1 may hold
//      the previous call's values), sums them, evaluates freq = nalleles / total and the keep predicate exactly as
//      freq_keep_kernel does, and stores sums, freq and keep into EVERY rank's block (the sums in place: a counter is
//      read and overwritten by its slice's owner only);
//   C  the last CTA to finish (its stores fenced at system scope) tells every peer "my slice has landed" and holds the
//      kernel until all peers have said the same: what follows on this rank's stream sees complete tables.
// 4.8 MB of counters at C2: each rank reads (N-1)/N of 4.8 MB / N ... about 9 MB cross NVLink per rank, two flag
// round trips — against an all-reduce whose latency (not bandwidth) was the largest multi-GPU loss of the step.
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"
#include "kernels.h"

namespace garlic {

namespace {

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed_sys(const int* p)
{
    int v;
    asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256)
xchg_freq_keep_kernel(const XchgParams P)
{
    __shared__ bool s_last;
    // ---- A: counters of every rank are final
    if (blockIdx.x == 0 && threadIdx.x < P.n) st_release_sys(P.flags[threadIdx.x] + kXchgReady + P.rank, P.seq);
    if (threadIdx.x < P.n) {
        const unsigned* f = P.flags[P.rank] + kXchgReady + threadIdx.x;
        while ((int)(ld_acquire_sys(f) - P.seq) < 0) { }
    }
    __syncthreads();
    // ---- B: this rank's slice
    const long long lo = P.L0 * P.rank / P.n, hi = P.L0 * (P.rank + 1) / P.n;
    for (long long s = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; s < hi; s += (long long)gridDim.x * blockDim.x) {
        int na = 0, tot = 0;
#pragma unroll 4
        for (int p = 0; p < P.n; ++p) {
            na += ld_relaxed_sys(P.counts[p] + s);
            tot += ld_relaxed_sys(P.counts[p] + P.L0 + s);
        }
        const double f = (tot == 0) ? 0.0 : ((double)na / (double)tot);      // garlic-data.cpp:141
        bool k = (f > 0 && f < 1);
        if (P.oob) {
            const int* cp = P.chr_param + 4 * P.chr_of[s];   // scaffold first, last, centromere start, end
            const int ps = P.pos[s];
            k = k && !(ps < cp[0]) && !(ps > cp[1]) && !(ps > cp[2] && ps < cp[3]);
        }
#pragma unroll 4
        for (int p = 0; p < P.n; ++p) {
            P.counts[p][s] = na;
            P.counts[p][P.L0 + s] = tot;
            P.freq[p][s] = f;
            P.keep[p][s] = (uint8_t)k;
        }
    }
    // ---- C: every rank's slice has landed everywhere
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(P.done_counter, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence_system();                                // the other CTAs' stores (fenced before their increment)
        if (threadIdx.x == 0) *P.done_counter = 0u;
        if (threadIdx.x < P.n) {
            st_release_sys(P.flags[threadIdx.x] + kXchgDone + P.rank, P.seq);
            const unsigned* f = P.flags[P.rank] + kXchgDone + threadIdx.x;
            while ((int)(ld_acquire_sys(f) - P.seq) < 0) { }
        }
    }
}

}  // namespace

cudaError_t launch_xchg_freq_keep(const XchgParams& P, cudaStream_t st)
{
    if (!P.L0 || P.n < 2 || P.n > kXchgMaxRanks) return cudaErrorInvalidValue;
    const long long per = (P.L0 + P.n - 1) / P.n;
    long long blocks = (per + 255) / 256;
    if (blocks > 148 * 2) blocks = 148 * 2;                    // all CTAs resident: they spin in phase A
    if (blocks < 1) blocks = 1;
    xchg_freq_keep_kernel<<<(unsigned)blocks, 256, 0, st>>>(P);
    return cudaGetLastError();
}

}  // namespace garlic
