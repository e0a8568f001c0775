// host_emu.cpp — compiles the product's __host__ __device__ walker (garlic_b200/csrc/walk.cuh) and
// segment/stitch logic (segments.h) for the CPU so the closed-form / chunked / bit-parallel logic can
// be checked against the literal oracle in a container without a GPU.  TEST HARNESS ONLY: the
// product never runs this; on the GPU box the same code is exercised through the CUDA kernels.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../garlic_b200/csrc/common.cuh"
#include "../garlic_b200/csrc/walk.cuh"
#include "../garlic_b200/csrc/segments.h"
#include "../garlic_b200/csrc/bound.cuh"
#include <cmath>

using namespace garlic;

// the walker reads GL-mode values lane-interleaved (common.cuh:gl_lane); tests hand a row-major [n_ind][gl_stride] matrix
static std::vector<double> interleave_gl(const double* gl, int64_t gl_stride, int n_ind)
{
    std::vector<double> t;
    if (!gl) return t;
    t.assign((size_t)((n_ind + 31) / 32) * gl_stride * kGlLanes, 0.0);
    for (int i = 0; i < n_ind; ++i)
        for (int64_t s = 0; s < gl_stride; ++s) t[((size_t)(i >> 5) * gl_stride + s) * kGlLanes + (i & 31)] = gl[(size_t)i * gl_stride + s];
    return t;
}

extern "C" {

// returns number of ROH written (≤ cap); out4 = (ind, chr, a, b) quadruples; *n_amb = ambiguous pairs
int emu_call_roh(const uint64_t* geno, int64_t row_words, const double* lut, const double* gl, int64_t gl_stride,
                 const double* freq, int n_ind, int n_chr, const int64_t* chr_off_, const int32_t* pos_,
                 const int32_t* cen_, int max_gap, int W, double cutoff, int thr, int chunk, double tol,
                 int32_t* out4, int cap, int* n_amb, int* n_items)
{
    std::vector<int64_t> chr_off(chr_off_, chr_off_ + n_chr + 1);
    const int64_t L = chr_off[n_chr];
    std::vector<int32_t> pos(pos_, pos_ + L), cen(cen_, cen_ + 2 * n_chr);
    std::vector<Segment> segs;
    std::vector<Item> items;
    build_segments(chr_off, pos, cen, max_gap, W, segs);
    build_items(chr_off, W, segs, chunk, 0, items);
    *n_items = (int)items.size();
    std::vector<RohRec> recs(1 << 20), amb(1 << 16);
    unsigned cnt[4] = {0, 0, 0, 0};
    WalkParams P;
    memset(&P, 0, sizeof(P));
    const std::vector<double> glt = interleave_gl(gl, gl_stride, n_ind);
    P.geno = geno; P.row_words = row_words; P.lut = lut; P.gl = gl ? glt.data() : nullptr; P.gl_stride = gl_stride; P.freq = freq;
    P.n_lanes = n_ind; P.W = W; P.thr = thr; P.cutoff = cutoff; P.tol = tol;
    P.out = recs.data(); P.out_count = cnt; P.out_cap = (unsigned)recs.size(); P.amb = amb.data(); P.amb_cap = (unsigned)amb.size();
    const int NW = ((W + 31) >> 5) + 1;
    std::vector<uint32_t> ring(NW);
    for (const Item& it : items)
        for (int k = 0; k < n_ind; ++k) {
            if (gl) walk_item<1, true, false>(P, it, k, true, ring.data(), 1, (const char*)lut, 0);
            else walk_item<0, true, false>(P, it, k, true, ring.data(), 1, (const char*)lut, 0);
        }
    recs.resize(cnt[0]);
    *n_amb = (int)cnt[1];
    std::vector<RohRec> merged;
    stitch_runs(recs, thr, merged);
    int n = 0;
    for (const RohRec& r : merged) {
        if (n < cap) { out4[4 * n] = r.ind; out4[4 * n + 1] = segs[r.tag >> 2].chr; out4[4 * n + 2] = r.a; out4[4 * n + 3] = r.b; }
        ++n;
    }
    return n;
}

// window dump: out[n_ind][slots], pre-filled by the caller with MISSING
int emu_windows(const uint64_t* geno, int64_t row_words, const double* lut, const double* gl, int64_t gl_stride,
                const double* freq, int n_ind, int n_chr, const int64_t* chr_off_, const int32_t* pos_,
                const int32_t* cen_, int max_gap, int W, int chunk, int step, double* out, int64_t slots)
{
    std::vector<int64_t> chr_off(chr_off_, chr_off_ + n_chr + 1);
    const int64_t L = chr_off[n_chr];
    std::vector<int32_t> pos(pos_, pos_ + L), cen(cen_, cen_ + 2 * n_chr);
    std::vector<Segment> segs;
    std::vector<Item> items;
    build_segments(chr_off, pos, cen, max_gap, W, segs);
    build_items(chr_off, W, segs, chunk, step, items);
    unsigned cnt[4] = {0, 0, 0, 0};
    WalkParams P;
    memset(&P, 0, sizeof(P));
    const std::vector<double> glt = interleave_gl(gl, gl_stride, n_ind);
    P.geno = geno; P.row_words = row_words; P.lut = lut; P.gl = gl ? glt.data() : nullptr; P.gl_stride = gl_stride; P.freq = freq;
    P.n_lanes = n_ind; P.W = W; P.thr = 1; P.cutoff = 0; P.tol = 0; P.out_count = cnt;
    P.dump = out; P.dump_stride = slots; P.dump_step = step;
    const int NW = ((W + 31) >> 5) + 1;
    std::vector<uint32_t> ring(NW);
    for (const Item& it : items)
        for (int k = 0; k < n_ind; ++k) {
            if (gl) walk_item<1, false, true>(P, it, k, true, ring.data(), 1, (const char*)lut, 0);
            else walk_item<0, false, true>(P, it, k, true, ring.data(), 1, (const char*)lut, 0);
        }
    return (int)items.size();
}
}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// pruning bound (bound.cuh) on the CPU: the same table construction, per-half-word step and candidate predicate the
// kernels of squeeze.cu run.  pmax[n_pieces][n_ind] as the fused kernel stores it.
// ---------------------------------------------------------------------------------------------------------------
template <int C2, int LAG, bool PARTIAL = false>
static void emu_bound_row(const uint32_t* hw_row, long long n_hw, const std::vector<uint4>& hw,
                          int n_pieces, uint32_t* pmax, int64_t stride, int ind, uint32_t low_mask = 0u)
{
    BoundState S;
    bound_reset(S);
    for (int pi = 0; pi <= n_pieces; ++pi) {
        for (int I = 0; I < 16; ++I) {
            const long long q = (long long)pi * 16 + I;
            const uint32_t h = q < n_hw ? hw_row[q] : 0xffffffffu;
            bound_step<C2, LAG, PARTIAL>(S, h, hw[q], I, low_mask);
            if (((I - C2) & 15) == 15) {
                const long long piece = (q - C2) >> 4;
                if (piece >= 0 && piece < n_pieces) pmax[piece * stride + ind] = bound_pack(S.pm_all, S.pm_tail);
                S.pm_all = -0x40000000; S.pm_tail = -0x40000000;
            }
        }
    }
}

extern "C" {

// returns 0, or 1 if the table breaks the bound's assumptions
int emu_bound(const uint64_t* geno, int64_t row_words, const double* lut, long long L, int n_ind, int W, uint32_t* pmax,
              int n_pieces)
{
    const int c2 = bound_c2(W), lag = bound_lag(W);
    if (W < kBoundMinW || (lag != 1 && lag != 2)) return -1;
    const long long n_hw = 16ll * (n_pieces + 2);
    std::vector<uint4> hw(n_hw);
    int invalid = 0;
    for (long long k = 0; k < n_hw; ++k) {
        hw[k] = bound_hw_entry(lut, k, L, &invalid);
        hw[k].w = k >= c2 ? (uint32_t)bound_block_max(lut, k - c2, W) : 0u;
    }
    for (int i = 0; i < n_ind; ++i) {
        const uint32_t* row = reinterpret_cast<const uint32_t*>(geno + (int64_t)i * row_words);
        const long long row_hw = row_words * 2;
        if (bound_partial(W)) {
            if (c2 == 1) emu_bound_row<1, 1, true>(row, row_hw, hw, n_pieces, pmax, n_ind, i, bound_low_mask(W));
            else emu_bound_row<2, 2, true>(row, row_hw, hw, n_pieces, pmax, n_ind, i, bound_low_mask(W));
            continue;
        }
        switch (c2) {
#define CASE(C) case C: if (lag == 1) emu_bound_row<C, 1>(row, row_hw, hw, n_pieces, pmax, n_ind, i); \
                        else emu_bound_row<C, 2>(row, row_hw, hw, n_pieces, pmax, n_ind, i); break;
            CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
            default: return -1;
        }
    }
    return invalid;
}

// piece-aligned items and their candidate flags: out[n_items][n_ind]; item_bounds[n_items][3] = w0, own_hi, we
int emu_select(const uint32_t* pmax, int n_ind, int n_chr, const int64_t* chr_off_, const int32_t* pos_, const int32_t* cen_,
               int max_gap, int W, double cutoff, double tol, int item_pieces, uint8_t* out, int32_t* item_bounds, int cap_items)
{
    std::vector<int64_t> chr_off(chr_off_, chr_off_ + n_chr + 1);
    const int64_t L = chr_off[n_chr];
    std::vector<int32_t> pos(pos_, pos_ + L), cen(cen_, cen_ + 2 * n_chr);
    std::vector<Segment> segs;
    std::vector<Item> items;
    build_segments(chr_off, pos, cen, max_gap, W, segs);
    build_items_aligned(chr_off, W, segs, kPiece * item_pieces, items);
    if ((int)items.size() > cap_items) return -1;
    int ok = 0;
    const int cut = bound_cut_store(cutoff, tol, &ok);
    if (!ok) return -2;
    for (size_t i = 0; i < items.size(); ++i) {
        item_bounds[3 * i] = items[i].w0; item_bounds[3 * i + 1] = items[i].own_hi; item_bounds[3 * i + 2] = items[i].we;
        for (int k = 0; k < n_ind; ++k) out[i * n_ind + k] = bound_item_candidate(pmax, n_ind, k, items[i], cut, bound_c2(W)) ? 1 : 0;
    }
    return (int)items.size();
}

// K3 by plan (bound.cuh:plan_half): rows_in [n_ind][in_words] → rows_out [n_ind][out_words] (pre-filled by the caller)
void emu_squeeze(const uint64_t* rows_in, int64_t in_words, const int32_t* src, long long L, int n_ind, uint64_t* rows_out,
                 int64_t out_words)
{
    const long long n_q = (L + 15) / 16;
    uint4 head, segs[kPlanSegMax];
    for (long long q = 0; q < n_q; ++q) {
        plan_half(src, L, q, &head, segs);
        const int a_base = (src[(q & ~15ll) * 16] >> 4) & ~1, I = (int)(q & 15);
        const uint32_t code = plan_fast_code(head, a_base, I);
        for (int i = 0; i < n_ind; ++i) {
            const uint32_t* hin = reinterpret_cast<const uint32_t*>(rows_in + (int64_t)i * in_words);
            uint32_t h;
            if (code != 0xffffu) {                         // the kernel's branch-free path
                const uint32_t* w = hin + a_base + I;
                h = plan_fast_apply(code, w[0], w[1], w[2], w[3], w[4]);
            } else if (head.x & 0x100u) {
                h = plan_window(head, hin[head.y], plan_need1(head) ? hin[head.y + 1] : 0u, plan_need2(head) ? hin[head.y + 2] : 0u);
            } else {
                h = head.w;
                for (uint32_t s = 0; s < head.x; ++s) h |= ((hin[segs[s].x] >> segs[s].y) & segs[s].z) << segs[s].w;
            }
            reinterpret_cast<uint32_t*>(rows_out + (int64_t)i * out_words)[q] = h;
        }
    }
}

}
