// kernels.h — launch wrappers of the hand-written kernels (internal C++ interface; the public
// boundary is the C ABI in include/garlic_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

namespace garlic {

constexpr int kWalkThreads = 256;     // 8 individual groups (warps) of one item per CTA
constexpr int kTileSnpsMax = 1536;    // table tile staged per CTA: 1536 SNPs × 32 B = 48 KB

// dense per-item individual lists produced by the pruning pass (list == nullptr: every individual)
struct CandList {
    const int* list = nullptr;       // [n_items][stride]
    const unsigned* cnt = nullptr;   // [n_items]
    int stride = 0;
};

// tile_snps > 0: every item touches at most tile_snps SNPs → stage its table slice in shared memory
cudaError_t launch_walk(const WalkParams& P, const Item* items, int n_items, bool gl_mode, bool roh,
                        bool dump, int tile_snps, const CandList& cl, cudaStream_t st);
struct CoarseParams;
cudaError_t launch_coarse_tables(const double* lut, long long n_hw, int W, uint32_t* mask, int2* cb, cudaStream_t st);
cudaError_t launch_coarse(const CoarseParams& P, const Item* items, int n_items, int* cand_list, unsigned* cand_cnt,
                          int cand_stride, cudaStream_t st);
cudaError_t launch_thin_windows(const uint64_t* geno, int64_t row_words, const double* lut, const int* ind_list, int n_lanes,
                                const int3* segs, int n_segs, const int2* meta, int n_chr, long long n_slots, int step, int W,
                                double* dump, int64_t dump_stride, const double* gl, int64_t gl_stride, cudaStream_t st);
cudaError_t launch_fill_f64(double* p, size_t n, double v, cudaStream_t st);
cudaError_t launch_bucket_by_individual(const RohRec* in, const unsigned* count, unsigned cap, unsigned* hist, int n_ind,
                                        RohRec* out, int thr, cudaStream_t st);
cudaError_t launch_tokenize_tped(const char* text, const long long* off, int n_snp, int n_ind, int ind_lo, uint8_t* alleles,
                                 int* nonblank, cudaStream_t st);
cudaError_t launch_first_allele(const uint8_t* alleles, int n_snp, int n_ind, int ind_offset, int missing,
                                unsigned long long* key, cudaStream_t st);
cudaError_t launch_code_alleles(const uint8_t* alleles, int n_snp, int n_ind, int missing,
                                const unsigned long long* key, long long snp0, uint64_t* geno,
                                int64_t row_words, int* counts, long long L0, cudaStream_t st);
cudaError_t launch_count_packed(const uint64_t* geno, int64_t row_words, int n_ind, long long L0,
                                int* counts, cudaStream_t st);
cudaError_t launch_freq_keep(const int* counts, long long L0, const int* pos, const int* chr_of,
                             const int* chr_param, int oob, double* freq, uint8_t* keep, cudaStream_t st);
cudaError_t launch_compact_geno(const uint64_t* gin, int64_t in_words, long long n_in_words, const uint32_t* keepw,
                                const int* first_word, const uint8_t* first_skip, long long L, uint64_t* gout,
                                int64_t out_words, int n_ind, cudaStream_t st);
cudaError_t launch_keep_scan(const uint8_t* keep, long long L0, const int* chr_of0, int n_chr, int* block_counts,
                             int* total, int* src, uint32_t* keepw, int* first_word, uint8_t* first_skip,
                             int* chr_off_kept, cudaStream_t st);
cudaError_t launch_bad_pairs(const int* pos, const int* chr_of, const int* cen, int max_gap, long long L, int* list,
                             unsigned* count, unsigned cap, cudaStream_t st);
cudaError_t launch_gather_i32(const int* in, const int* src, long long L, int* out, cudaStream_t st);
cudaError_t launch_compact_gl(const double* in, int64_t in_stride, const int* src, long long L, const uint64_t* geno0,
                              int64_t row_words0, const double* freq0, double* out, int64_t out_stride, int n_ind, int type,
                              cudaStream_t st);
cudaError_t launch_gather_f64(const double* in, const int* src, long long L, double* out, cudaStream_t st);
cudaError_t launch_build_lut(const double* freq, long long L, double error, double* lut, cudaStream_t st);
cudaError_t launch_wlod_weights(const int* pos, const double* gpos, const int* chr_of, const int* chr_start,
                                long long L, double mu, int M, double* nomut, double* norec, cudaStream_t st);
cudaError_t launch_build_wlut(const double* lut, const double* nomut, const double* norec, long long L,
                              double* wlut, cudaStream_t st);

}  // namespace garlic
