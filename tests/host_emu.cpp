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

using namespace garlic;

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
    P.geno = geno; P.row_words = row_words; P.lut = lut; P.gl = gl; P.gl_stride = gl_stride; P.freq = freq;
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
    P.geno = geno; P.row_words = row_words; P.lut = lut; P.gl = gl; P.gl_stride = gl_stride; P.freq = freq;
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

}
