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
#include "../garlic_b200/csrc/coarse.cuh"
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


// pruning pass (coarse.cuh) on the CPU: out[n_items][n_ind] = candidate flag; items as the chunked fast pass
// builds them.  The table construction restates coarse_tables_kernel (kernels.cu).
int emu_coarse(const uint64_t* geno, int64_t row_words, const double* lut, int n_ind, int n_chr, const int64_t* chr_off_,
               const int32_t* pos_, const int32_t* cen_, int max_gap, int W, double cutoff, double tol, double error,
               int chunk, uint8_t* out, int32_t* item_bounds /* [n_items][3] = w0, own_hi, we */, int cap_items)
{
    std::vector<int64_t> chr_off(chr_off_, chr_off_ + n_chr + 1);
    const int64_t L = chr_off[n_chr];
    std::vector<int32_t> pos(pos_, pos_ + L), cen(cen_, cen_ + 2 * n_chr);
    std::vector<Segment> segs;
    std::vector<Item> items;
    build_segments(chr_off, pos, cen, max_gap, W, segs);
    build_items(chr_off, W, segs, chunk, 0, items);
    if ((int)items.size() > cap_items) return -1;
    const int c1 = (W - 16) >> 4, c2 = (W + 14) >> 4;
    const int64_t n_hw = (L + 4160 - 512) >> 4;
    const double scale = (double)(1 << kCoarseShift);
    std::vector<uint32_t> mask(n_hw);
    std::vector<int2> cbv(n_hw + 16);
    for (int i = 0; i < 16; ++i) { cbv[i].x = 0; cbv[i].y = 0; }
    for (int64_t k = 0; k < n_hw; ++k) {
        const int64_t s0 = k * 16;
        double b = 0.0;
        for (int i = 0; i < W; ++i) { const double* e = lut + (s0 + i) * 4; b += std::fmin(e[0], e[2]); }
        double bmax = b;
        for (int j = 1; j < 16; ++j) {
            const double* eo = lut + (s0 + j - 1) * 4;
            const double* ei = lut + (s0 + j - 1 + W) * 4;
            b = b - std::fmin(eo[0], eo[2]) + std::fmin(ei[0], ei[2]);
            bmax = std::fmax(bmax, b);
        }
        double dlo = 0.0, dhi = 0.0;
        for (int64_t s = s0; s < s0 + 16 * (c2 + 1); ++s) {
            const double* e = lut + s * 4;
            const double d = std::fabs(e[0] - e[2]);
            if (d > kCoarseSplit) dhi = std::fmax(dhi, d); else dlo = std::fmax(dlo, d);
        }
        uint32_t m = 0;
        for (int j = 0; j < 16; ++j) {
            const double* e = lut + (s0 + j) * 4;
            if (e[2] > e[0]) m |= 1u << (2 * j);
            if (std::fabs(e[0] - e[2]) > kCoarseSplit) m |= 2u << (2 * j);
        }
        const double qlo = std::fmin(std::ceil(dlo * scale) + 1.0, 65535.0), qhi = std::fmin(std::ceil(dhi * scale) + 1.0, 65535.0);
        mask[k] = m;
        cbv[16 + k].x = (int)std::ceil(bmax * scale) + 2;
        cbv[16 + k].y = (int)((uint32_t)qlo | ((uint32_t)qhi << 16));
    }
    CoarseParams P;
    P.geno = geno; P.row_words = row_words; P.mask = mask.data(); P.cb = cbv.data() + 16;
    P.W = W; P.c1 = c1; P.c2 = c2; P.n_lanes = n_ind;
    P.chet_fixed = (int)std::ceil((std::log10(error) + 1e-9) * scale);
    P.cut_fixed = (int)std::floor((cutoff - tol) * scale - 2.0);
    for (size_t i = 0; i < items.size(); ++i) {
        item_bounds[3 * i] = items[i].w0; item_bounds[3 * i + 1] = items[i].own_hi; item_bounds[3 * i + 2] = items[i].we;
        for (int k = 0; k < n_ind; ++k) out[i * n_ind + k] = coarse_item_any(P, items[i], k) ? 1 : 0;
    }
    return (int)items.size();
}

}
