// wlod.h — K6 (LD band) and K5-W (weighted windows → ROH) launch wrappers.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include "common.cuh"

namespace garlic {

// A weight row is kInvFront zeros, the W reciprocals, then zeros up to the stride: the tensor-core pass reads a few
// elements either side of the band and must find 0 there (no predicates on its operand loads).
constexpr int kInvFront = 8, kInvBack = 24;
constexpr int kWlodMmaMinW = 10;   // the tensor-core pass's tile schedule needs (W+6)/4 >= 4
// (even: a block of 32 weight rows then starts and ends on 16 bytes, what a bulk copy into shared memory needs)
GHD int inv_stride(int W) { return W + kInvFront + kInvBack + (W & 1); }

struct WlodParams {
    WalkParams base;
    const double* wlut;    // [L+pad][4]  lod(g)*nomut*norec for a global error
    const double* invld;   // weight rows, one per window: invld[w * kInvStride(W) + kInvFront + k] = 1.0 / LD[w][k]; zeros around
    const double* nomut;   // [L+pad]     (GL mode: score evaluated per genotype)
    const double* norec;   // [L+pad]
};

// --phased LD (r2 between haplotypes, garlic-data.cpp:585-617): where the first-copy bits and frequencies come from.
// alleles == nullptr: unphased (hr2).
struct LdPhase {
    const uint8_t* alleles = nullptr;          // [L0][n_local][2] allele characters as K0/K1 left them
    const unsigned long long* key = nullptr;   // [L0] first-allele keys (low byte = the "1" allele, ~0 = none)
    const int* src = nullptr;                  // [L] kept SNP -> source SNP
    int missing = '0';
    const double* freq = nullptr;              // [L] allele frequencies in use
};

cudaError_t launch_wlod_walk(const WlodParams& Q, const Item* items, int n_items, bool gl_mode, bool roh,
                             bool dump, cudaStream_t st);

// fast pass of pass 2 on the FP64 tensor cores (tolerance-checked; needs base.tol > 0)
cudaError_t launch_wlod_mma(const WlodParams& Q, const Item* items, int n_items, bool gl_mode, cudaStream_t st);

// LD band: hr² pair matrix over the listed individuals → window sums → reciprocal.  ld_ind indexes the whole sample;
// this GPU holds individuals [ind_lo, ind_lo+n_local); with comm the bit-planes are all-reduced across ranks.
// invld: padded weight rows (see above), zero-filled by the caller; ld_out (optional): [L][W] the sums themselves
// (reference LDData layout).
cudaError_t launch_ld_band(const uint64_t* geno, int64_t row_words, const int* ld_ind, int n_ld,
                           const double* homf, const int* chr_of, const int* chr_start, int n_chr,
                           long long L, int W, double* invld, double* ld_out, cudaStream_t st, int* n_launches,
                           ncclComm_t comm, int ind_lo, int n_local, uint64_t* planes, double* P, const LdPhase& ph);
// scratch the caller provides: bit-planes [L][2][ceil(n_ld/64)] words, ordered pair matrix [L][2W-1] doubles
size_t ld_planes_words(long long L, int n_ld, bool phased);
size_t ld_pairs_doubles(long long L, int W);
cudaError_t launch_hom_freq(const int* counts, long long L0, const int* src, long long L, double* homf, cudaStream_t st);

}  // namespace garlic
