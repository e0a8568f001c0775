// squeeze.cu — K3 (column compaction of the packed genotype matrix, filterMonomorphic[AndOOB]Sites,
// garlic-data.cpp:871-1195, as a bit-level gather) fused with the pruning bound of K5 pass 2 (bound.cuh), and the
// candidate selection that follows once the cutoff is known.
//
// squeeze_bound_kernel: one lane = one individual, one warp = 32 individuals marching along the SNP axis over a range of
// 256-SNP pieces.  The rows' input half-words stream through a per-warp shared-memory ring filled by cp.async (each
// copy instruction moves 128 contiguous bytes of two rows, so global reads are whole lines instead of one sector per
// lane); the compaction plan (which input half-words make up an output half-word) is the same for every individual and
// is read with warp-uniform loads; the compacted half-word goes (a) through a shared staging tile to coalesced
// 128-byte row stores and (b), still in its register, into the bound (three POPC and a dozen integer operations).
// SQUEEZE = false: bound only, over rows that are already compacted (another window size on the same data).
// BOUND = false: compaction only (GL / weighted modes, window sizes outside the bound's range).
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include "common.cuh"
#include "bound.cuh"
#include "kernels.h"

namespace garlic {

namespace {
constexpr int kSqWarps = 2;                               // 43 KB of shared memory per CTA: five CTAs = ten warps per SM
constexpr int kSqStages = 4;                              // ring = 4 stages of 32 input half-words (128 B) per row
constexpr int kSqRingHw = kSqStages * 32;                 // input half-word a of a row sits at ring position a & 127
constexpr int kSqRowBytes = kSqRingHw * 4 + 8;            // + 8 B pad: rows start 2 banks apart
constexpr int kSqRingBytes = 32 * kSqRowBytes;            // 32 rows
constexpr int kSqOutRowBytes = 136;                       // output staging: 16 words (128 B) per row + pad
constexpr int kSqOutBytes = 32 * kSqOutRowBytes;
constexpr int kSqStashBytes = 640;                        // the current piece's plan heads, bound tables, block maxima
constexpr int kSqWarpBytes = kSqRingBytes + kSqOutBytes + kSqStashBytes;

__device__ __forceinline__ uint32_t sq_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async8(uint32_t dst_smem, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_dyn(int pending)     // at most `pending` (0..3) newest groups still in flight
{
    switch (pending) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    }
}
}  // namespace

// ------------------------------------------------------------------------------------------
// tables of the bound for window size W (bound.cuh): one thread per half-word / block
// ------------------------------------------------------------------------------------------
__global__ void bound_tables_kernel(const double* __restrict__ lut, long long n_hw, long long L, int W,
                                    uint4* __restrict__ hw, int2* __restrict__ bc, int* __restrict__ invalid)
{
    // bc[q] = {Bmax of block q - C2, chet of half-word q}: what the step at half-word q needs next to hw[q]
    const int c2 = bound_c2(W);
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n_hw; k += (long long)gridDim.x * blockDim.x) {
        int bad = 0, chet = 0;
        hw[k] = bound_hw_entry(lut, k, L, &chet, &bad);
        bc[k].y = chet;
        bc[k].x = k >= c2 ? bound_block_max(lut, k - c2, W) : 0;
        if (bad) atomicOr(invalid, 1);
    }
}

cudaError_t launch_bound_tables(const double* lut, long long n_hw, long long L, int W, uint4* hw, int2* bc, int* invalid,
                                cudaStream_t st)
{
    if (!n_hw) return cudaSuccess;
    long long blocks = (n_hw + 127) / 128;
    if (blocks > 148 * 32) blocks = 148 * 32;
    bound_tables_kernel<<<(unsigned)blocks, 128, 0, st>>>(lut, n_hw, L, W, hw, bc, invalid);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// compaction plan: one thread per output half-word (bound.cuh:plan_half) from the gather list src[]
// ------------------------------------------------------------------------------------------
__global__ void plan_kernel(const int* __restrict__ src, const int* __restrict__ n_kept, long long n_q,
                            uint4* __restrict__ head, uint4* __restrict__ segs, int4* __restrict__ piece_rng)
{
    const long long L = *n_kept;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n_q; q += (long long)gridDim.x * blockDim.x) {
        uint4 hd;
        plan_half(src, L, q, &hd, segs + q * kPlanSegMax);
        head[q] = hd;
        // per piece: first / last input half-word it reads, and which of its 16 half-words need the slow path
        const unsigned slow = __ballot_sync(__activemask(), !plan_is_fast(hd));
        if ((q & 15) == 0) {
            const long long d0 = q * 16;
            int4 r = make_int4(0, -1, 0, 0);               // nothing to read
            if (d0 < L) { const long long d1 = d0 + kPiece - 1 < L ? d0 + kPiece - 1 : L - 1; r.x = src[d0] >> 4; r.y = src[d1] >> 4; }
            r.z = (int)((slow >> (threadIdx.x & 16)) & 0xffffu);
            piece_rng[q >> 4] = r;
        }
    }
}

cudaError_t launch_plan(const int* src, const int* n_kept, long long n_q, uint4* head, uint4* segs, int4* piece_rng,
                        cudaStream_t st)
{
    if (!n_q) return cudaSuccess;
    long long blocks = (n_q + 127) / 128;
    if (blocks > 148 * 16) blocks = 148 * 16;
    plan_kernel<<<(unsigned)blocks, 128, 0, st>>>(src, n_kept, n_q, head, segs, piece_rng);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// the fused pass
// ------------------------------------------------------------------------------------------
template <bool SQUEEZE, bool BOUND, int C2, int LAG>
__global__ void __launch_bounds__(kSqWarps * 32)
squeeze_bound_kernel(const SqueezeParams P)
{
    extern __shared__ __align__(16) unsigned char sq_smem[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    unsigned char* ring_s = sq_smem + (size_t)warp * kSqWarpBytes;
    unsigned char* out_s = ring_s + kSqRingBytes;
    const uint32_t* my_ring = reinterpret_cast<const uint32_t*>(ring_s + lane * kSqRowBytes);
    const uint2* my_ring64 = reinterpret_cast<const uint2*>(ring_s + lane * kSqRowBytes);
    uint32_t* my_out = reinterpret_cast<uint32_t*>(out_s + lane * kSqOutRowBytes);   // this lane's 32 output half-words
    uint4* sh_head = reinterpret_cast<uint4*>(out_s + kSqOutBytes);
    uint4* sh_hw = sh_head + 16;
    int2* sh_bc = reinterpret_cast<int2*>(sh_hw + 16);
    const int c8 = lane & 15, rsel = lane >> 4;           // copies / flushes: a half-warp covers 128 B of one row
    const uint32_t ring_u32 = sq_smem_u32(ring_s);
    const int n_rg = (P.n_ind + 31) >> 5;
    const long long n_kept = SQUEEZE ? (long long)*P.n_kept : 0;
    const long long n_out_words = SQUEEZE ? (n_kept + 31) >> 5 : 0;
    const long long n_tasks = (long long)n_rg * P.n_col_tasks;
    for (long long task = (long long)blockIdx.x * kSqWarps + warp; task < n_tasks; task += (long long)gridDim.x * kSqWarps) {
        const int rg = (int)(task % n_rg), ct = (int)(task / n_rg);
        const int pc_lo = ct * P.pieces_per_task;
        const int pc_hi = min(P.n_pieces, pc_lo + P.pieces_per_task);
        const int row = rg * 32 + lane;
        const bool active = row < P.n_ind;
        const uint32_t* grow32 = reinterpret_cast<const uint32_t*>(P.gin + (int64_t)(active ? row : P.n_ind - 1) * P.in_words);
        // ---- input ring: stage t = input half-words [32 t, 32 t + 32) of the 32 rows, in ring positions (32 t) & 127 …
        int issued = 0;                                    // stages below this one have been issued (or skipped)
        auto issue = [&](int t) {
            const long long w0 = (long long)t * 16;        // first 64-bit word of the stage
            if (w0 + 16 <= P.in_words) {
                const uint32_t dst = ring_u32 + (uint32_t)((t & (kSqStages - 1)) * 128 + c8 * 8);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int r = 2 * i + rsel;
                    int gr = rg * 32 + r;
                    gr = gr < P.n_ind ? gr : P.n_ind - 1;
                    cp_async8(dst + r * kSqRowBytes, P.gin + (int64_t)gr * P.in_words + w0 + c8);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // Input half-words [a_lo, a_hi] must be readable: at most two stages (the usual case: a piece reads about 17
        // consecutive input half-words), the other two ring stages stay in flight ahead.  Returns false when the range is
        // wider (a long run of dropped SNPs inside the piece): that piece reads global memory directly.
        auto stage_in = [&](int a_lo, int a_hi) -> bool {
            if (a_hi < a_lo) return true;                  // nothing to read
            const int lo_st = a_lo >> 5, hi_st = a_hi >> 5;
            if (hi_st - lo_st > 1) return false;
            if (issued < lo_st) {                          // stages nobody reads are never copied; what is still in
                asm volatile("cp.async.wait_group 0;" ::: "memory");   // flight lands before its slot gets a new owner
                issued = lo_st;
            }
            __syncwarp();                                  // every lane is done with the stages about to be overwritten
#pragma unroll 1
            while (issued <= lo_st + kSqStages - 1) issue(issued++);
            cp_async_wait_dyn(issued - 1 - hi_st);
            __syncwarp();
            return true;
        };

        BoundState S;
        bound_reset(S);
        // The plan heads and bound tables of a piece are the same for every individual: 16 lanes fetch them with one
        // coalesced load each, one piece ahead (the L2 latency hides behind a whole piece of work), and park them in
        // shared memory, from where every half-word step takes them with broadcast reads.
        uint4 r_head = make_uint4(0u, 0u, 0u, 0u), r_hw = make_uint4(0u, 0u, 0u, 0u);
        int2 r_bc = make_int2(0, 0);
        int4 r_rng = make_int4(0, -1, 0, 0);
        auto fetch = [&](int piece) {
            if (lane < 16) {
                const long long q = (long long)piece * 16 + lane;
                if (SQUEEZE) r_head = P.plan_head[q];
                if (BOUND) { r_hw = P.hw[q]; r_bc = P.bc[q]; }
            }
            r_rng = SQUEEZE ? P.piece_rng[piece] : make_int4(piece * 16, piece * 16 + 15, 0, 0);
        };
        fetch(pc_lo);
        const int pi_end = BOUND ? pc_hi + 1 : pc_hi;     // one more piece drains the bound's C2 half-words of lag
#pragma unroll 1
        for (int pi = pc_lo; pi < pi_end; ++pi) {
            const long long qb = (long long)pi * 16;
            __syncwarp();                                  // the previous piece's stash has been read
            const int4 rng = r_rng;
            const bool ring_ok = stage_in(rng.x, rng.y);
            if (lane < 16) {
                if (SQUEEZE) sh_head[lane] = r_head;
                if (BOUND) { sh_hw[lane] = r_hw; sh_bc[lane] = r_bc; }
            }
            __syncwarp();
            fetch(pi + 1);                                 // the tables have two pieces of slack
            const uint32_t* hsrc;                          // where the bound reads the piece's 16 half-words
            if (SQUEEZE) {
                // ---- phase A: the piece's 16 output half-words, by plan.  Branch-free for the usual half-word — 16
                // consecutive sources with at most one dropped SNP in between: a funnel shift of the 64-bit window at
                // input half-word a, then the fields above the dropped one move down by one; the two 8-byte reads are
                // conflict-free (rows start two banks apart) — the others (slow bits of the piece) are redone after.
                uint32_t* dst = my_out + ((pi & 1) << 4);
                unsigned slow = ring_ok ? (unsigned)rng.z : 0xffffu;
                if (ring_ok) {
#pragma unroll
                    for (int I = 0; I < 16; ++I) {
                        const uint4 hd = sh_head[I];
                        const uint32_t u = hd.y >> 1;
                        const uint2 p0 = my_ring64[u & (kSqRingHw / 2 - 1)], p1 = my_ring64[(u + 1) & (kSqRingHw / 2 - 1)];
                        const bool odd = hd.y & 1u;
                        const uint32_t x0 = odd ? p0.y : p0.x, x1 = odd ? p1.x : p0.y, x2 = odd ? p1.y : p1.x;
                        const uint32_t sh = hd.z & 255u;
                        const uint32_t lo = __funnelshift_r(x0, x1, sh), hi = __funnelshift_r(x1, x2, sh);
                        const uint32_t slo = __funnelshift_r(lo, hi, 2u);
                        dst[I] = (lo & hd.w) | (slo & ~hd.w);
                    }
                }
#pragma unroll 1
                for (; slow; slow &= slow - 1u) {          // several dropped SNPs, long dropped runs, the end of the data
                    const int I = __ffs((int)slow) - 1;
                    const uint4 hd = sh_head[I];
                    uint32_t h;
                    if (hd.x & 0x100u) {
                        const uint32_t a = hd.y;
                        if (ring_ok) h = plan_window(hd, my_ring[a & (kSqRingHw - 1)], my_ring[(a + 1) & (kSqRingHw - 1)], my_ring[(a + 2) & (kSqRingHw - 1)]);
                        else h = plan_window(hd, grow32[a], plan_need1(hd) ? grow32[a + 1] : 0u, plan_need2(hd) ? grow32[a + 2] : 0u);
                    } else {
                        h = hd.w;
                        const uint4* sg = P.plan_seg + (qb + I) * kPlanSegMax;
                        const int ns = (int)hd.x;
#pragma unroll 1
                        for (int k = 0; k < ns; ++k) {
                            const uint4 g = sg[k];
                            const uint32_t x = ring_ok ? my_ring[g.x & (kSqRingHw - 1)] : grow32[g.x];
                            h |= ((x >> g.y) & g.z) << g.w;
                        }
                    }
                    dst[I] = h;
                }
                hsrc = dst;
            } else {
                hsrc = my_ring + (qb & (kSqRingHw - 1));   // a piece is half a stage: never wraps
            }
            if (BOUND) {
                // ---- phase B: the bound over the 16 half-words, ring indices compile-time
#pragma unroll
                for (int I = 0; I < 16; ++I) {
                    const long long q = qb + I;
                    bound_step<C2, LAG>(S, hsrc[I], sh_hw[I], sh_bc[I], I);
                    if (((I - C2) & 15) == 15) {           // block k = q - C2 closes its piece
                        const int piece = (int)((q - C2) >> 4);
                        if (piece >= pc_lo && piece < pc_hi && active)
                            P.pmax[(int64_t)piece * P.pmax_stride + row] = bound_pack(S.pm_all, S.pm_tail);
                        S.pm_all = -0x40000000; S.pm_tail = -0x40000000;
                    }
                }
            }
            // ---- two pieces = 16 output words per row: flush through shared memory, 128 B per row per store
            if (SQUEEZE && pi < pc_hi && ((pi & 1) || pi == pc_hi - 1)) {
                __syncwarp();
                const long long wbase = 16ll * (pi >> 1);
                const int nw = (pi & 1) ? 16 : 8;
#pragma unroll 4
                for (int i = 0; i < 16; ++i) {
                    const int r = 2 * i + rsel;
                    const uint64_t v = *reinterpret_cast<const uint64_t*>(out_s + r * kSqOutRowBytes + c8 * 8);
                    const int gr = rg * 32 + r;
                    const long long w = wbase + c8;
                    if (gr < P.n_ind && c8 < nw && w < n_out_words) P.gout[(int64_t)gr * P.out_words + w] = v;
                }
                __syncwarp();
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");   // nothing of this task is still landing
        __syncwarp();
    }
}

template <bool SQUEEZE, bool BOUND, int C2, int LAG>
static cudaError_t launch_sq_t(const SqueezeParams& P, unsigned grid, size_t smem, cudaStream_t st)
{
    auto kern = squeeze_bound_kernel<SQUEEZE, BOUND, C2, LAG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kSqWarps * 32, smem, st>>>(P);
    return cudaGetLastError();
}

template <bool SQUEEZE, int LAG>
static cudaError_t launch_sq_c2(const SqueezeParams& P, int c2, unsigned grid, size_t smem, cudaStream_t st)
{
    switch (c2) {
        case 2: return launch_sq_t<SQUEEZE, true, 2, LAG>(P, grid, smem, st);
        case 3: return launch_sq_t<SQUEEZE, true, 3, LAG>(P, grid, smem, st);
        case 4: return launch_sq_t<SQUEEZE, true, 4, LAG>(P, grid, smem, st);
        case 5: return launch_sq_t<SQUEEZE, true, 5, LAG>(P, grid, smem, st);
        case 6: return launch_sq_t<SQUEEZE, true, 6, LAG>(P, grid, smem, st);
        case 7: return launch_sq_t<SQUEEZE, true, 7, LAG>(P, grid, smem, st);
        case 8: return launch_sq_t<SQUEEZE, true, 8, LAG>(P, grid, smem, st);
        case 9: return launch_sq_t<SQUEEZE, true, 9, LAG>(P, grid, smem, st);
        case 10: return launch_sq_t<SQUEEZE, true, 10, LAG>(P, grid, smem, st);
        case 11: return launch_sq_t<SQUEEZE, true, 11, LAG>(P, grid, smem, st);
        case 12: return launch_sq_t<SQUEEZE, true, 12, LAG>(P, grid, smem, st);
        case 13: return launch_sq_t<SQUEEZE, true, 13, LAG>(P, grid, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

// pieces per warp task: enough tasks for a few waves of 12 warps per SM, an even number of pieces (the output
// staging tile holds two), and long enough that the extra piece the bound drains costs a few percent
int squeeze_pieces_per_task(int n_ind, int n_pieces)
{
    const int n_rg = (n_ind + 31) / 32;
    const int want_tasks = 148 * 10 * 4;
    int n_col = std::max(1, (want_tasks + n_rg - 1) / n_rg);
    int ppt = (n_pieces + n_col - 1) / n_col;
    ppt = std::max(ppt, 16);
    ppt += ppt & 1;
    return ppt;
}

cudaError_t launch_squeeze_bound(SqueezeParams P, bool squeeze, int c2, cudaStream_t st)
{
    if (!P.n_ind || !P.n_pieces) return cudaSuccess;
    P.pieces_per_task = squeeze_pieces_per_task(P.n_ind, P.n_pieces);
    P.n_col_tasks = (P.n_pieces + P.pieces_per_task - 1) / P.pieces_per_task;
    const int n_rg = (P.n_ind + 31) / 32;
    const long long n_tasks = (long long)n_rg * P.n_col_tasks;
    long long grid = (n_tasks + kSqWarps - 1) / kSqWarps;
    if (grid > 148ll * 5 * 8) grid = 148ll * 5 * 8;
    const size_t smem = (size_t)kSqWarps * kSqWarpBytes;
    if (c2 <= 0) {
        if (!squeeze) return cudaErrorInvalidValue;
        return launch_sq_t<true, false, 2, 1>(P, (unsigned)grid, smem, st);
    }
    if (P.lag == 1) return squeeze ? launch_sq_c2<true, 1>(P, c2, (unsigned)grid, smem, st) : launch_sq_c2<false, 1>(P, c2, (unsigned)grid, smem, st);
    if (P.lag == 2) return squeeze ? launch_sq_c2<true, 2>(P, c2, (unsigned)grid, smem, st) : launch_sq_c2<false, 2>(P, c2, (unsigned)grid, smem, st);
    return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------
// candidate selection: one CTA per item thresholds the piece maxima of every individual (bound.cuh:
// bound_item_candidate), writes the item's dense candidate list in individual order and appends one work unit per
// `lanes_per_unit` candidates to the walker's queue.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
select_kernel(const Item* __restrict__ items, int n_items, const uint32_t* __restrict__ pmax, int64_t stride, int n_ind,
              int cut_store, const int* __restrict__ invalid, int* __restrict__ cand_list, int cand_stride,
              unsigned* __restrict__ cand_cnt, int2* __restrict__ units, unsigned* __restrict__ n_units, unsigned unit_cap,
              int lanes_per_unit)
{
    __shared__ int s_w[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (*invalid) cut_store = -32768;                      // the table breaks the bound's assumptions: keep everybody
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const Item it = items[item];
        // a thread takes individuals t, t + 256, …: all of its loads are independent, one block scan per item; the
        // list comes out ordered by (thread, individual) — any order will do, the runs are sorted afterwards
        int mine = 0;
        for (int ind = threadIdx.x; ind < n_ind; ind += 256) mine += bound_item_candidate(pmax, stride, ind, it, cut_store) ? 1 : 0;
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        __syncthreads();                                   // s_w of the previous item has been read
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const int v = s_w[w]; total += v; if (w < warp) before += v; }
        int pos = before + incl - mine;
        if (mine) {
            int* list = cand_list + (int64_t)item * cand_stride;
            for (int ind = threadIdx.x; ind < n_ind; ind += 256)
                if (bound_item_candidate(pmax, stride, ind, it, cut_store)) list[pos++] = ind;
        }
        if (threadIdx.x == 0) {
            cand_cnt[item] = (unsigned)total;
            const int nu = (total + lanes_per_unit - 1) / lanes_per_unit;
            if (nu) {
                const unsigned b = atomicAdd(n_units, (unsigned)nu);
                for (int u = 0; u < nu; ++u)
                    if (b + u < unit_cap) units[b + u] = make_int2(item, u * lanes_per_unit);
                atomicAdd(n_units + 1, (unsigned)total);  // total candidate pairs (statistics)
            }
        }
    }
}

cudaError_t launch_select(const Item* items, int n_items, const uint32_t* pmax, int64_t stride, int n_ind, int cut_store,
                          const int* invalid, int* cand_list, int cand_stride, unsigned* cand_cnt, int2* units, unsigned* n_units,
                          unsigned unit_cap, int lanes_per_unit, cudaStream_t st)
{
    if (!n_items || !n_ind) return cudaSuccess;
    const int grid = n_items < 148 * 8 ? n_items : 148 * 8;
    select_kernel<<<grid, 256, 0, st>>>(items, n_items, pmax, stride, n_ind, cut_store, invalid, cand_list, cand_stride, cand_cnt, units,
                                        n_units, unit_cap, lanes_per_unit);
    return cudaGetLastError();
}

}  // namespace garlic
