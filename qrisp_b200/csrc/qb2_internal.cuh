// Internal helpers shared by the translation units of libqb2.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "qrisp_b200.h"

namespace qb2 {

template <typename real>
struct Traits;
template <>
struct Traits<float> {
    using vec = float2;
    using frac = uint32_t;
    using frac2 = uint2;
    static constexpr int dtype = QB2_C64;
};
template <>
struct Traits<double> {
    using vec = double2;
    using frac = uint64_t;
    using frac2 = ulonglong2;
    static constexpr int dtype = QB2_C128;
};

// error plumbing (qb2_api.cu)
int set_error(int code, const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);
void count_launch();
int sm_count();

#define QB2_CUDA(call)                                   \
    do {                                                 \
        int _rc = ::qb2::check_cuda((call), #call);      \
        if (_rc != QB2_OK) return _rc;                   \
    } while (0)

#define QB2_MAX_DEVICES 64
#define QB2_MAX_PEERS 16

// sparse input of the first pass after a reset (QB2_FEATURE_ZERO_INPUT): the listed non-zero amplitudes,
// sorted by tile number; `loc` = (thread << r) | register under the FIRST phase's distribution
struct SparseInput {
    const uint32_t *tile;  // [count] ascending tile numbers
    const uint32_t *loc;   // [count]
    const void *amp;       // [count] amplitudes in the pass dtype
    uint32_t count;
    // (rides along: the per-launch extras of a pass travel in this one by-value struct)
    double *tile_norms;    // [tiles] sum of |amp|**2 of every tile as it is stored, or null
};

struct PassLaunch {
    void *state;
    uint64_t hi_base;
    const qb2_pass_header *h;
    const void *blob_dev;
    uint32_t ntiles;
    cudaStream_t s;
    SparseInput sparse;
    const void *jit_kernel;  // a specialised kernel of this very pass (qb2_jit_load), or null
};

// one translation unit per kernel variant (qb2_pass_inst.cu compiled with -DQB2_INST=k, see Makefile):
// k = 2 * geometry + full; geometries in the order of qb2_pass.cu's table; 12, 13: the TMA kernel
#define QB2_N_PASS_VARIANTS 14
#define QB2_DECLARE_VARIANT(k) int launch_pass_variant_##k(const PassLaunch &L);
QB2_DECLARE_VARIANT(0) QB2_DECLARE_VARIANT(1) QB2_DECLARE_VARIANT(2) QB2_DECLARE_VARIANT(3) QB2_DECLARE_VARIANT(4)
QB2_DECLARE_VARIANT(5) QB2_DECLARE_VARIANT(6) QB2_DECLARE_VARIANT(7) QB2_DECLARE_VARIANT(8) QB2_DECLARE_VARIANT(9)
QB2_DECLARE_VARIANT(10) QB2_DECLARE_VARIANT(11) QB2_DECLARE_VARIANT(12) QB2_DECLARE_VARIANT(13)
#undef QB2_DECLARE_VARIANT

// launchers implemented per translation unit
int launch_pass(void *state, int n, uint64_t hi_base, const qb2_pass_header *h, const void *blob_dev, int dtype,
                cudaStream_t s, const void *sparse_dev, uint32_t n_sparse, const void *jit_kernel = nullptr,
                double *tile_norms_dev = nullptr);
int launch_chunk_sums(const void *state, int n, int cb, double *sums_dev, int dtype, cudaStream_t s);
int launch_scan_sums(double *sums_dev, uint64_t count, cudaStream_t s);
int launch_sample_draw_cb(const void *state, int n, int cb, const double *chunk_sums_dev, const double *uniforms_dev, int64_t shots,
                          uint64_t *indices_dev, int dtype, cudaStream_t s);
int launch_init_sparse(void *state, int n, const uint64_t *idx_dev, const void *amp_dev, uint32_t count, int dtype,
                       cudaStream_t s);
int launch_init_basis(void *state, int n, uint64_t basis, int dtype, cudaStream_t s);
int launch_matrix_big(const void *src, void *dst, int n, uint64_t hi_base, const void *matrix_dev, const int *targets,
                      int k, uint64_t cmask, uint64_t cval, int dtype, cudaStream_t s);
int launch_block_tc(void *state, int n, uint64_t hi_base, const void *gate_dev, const int *targets, uint64_t cmask, uint64_t cval,
                    cudaStream_t s);
int pack_block_tc(const double *u, float *out);
int launch_mux_ry(void *state, int n, uint64_t hi_base, int target, const int *ctrl, int nc, const void *cs_dev, uint64_t cmask,
                  uint64_t cval, int dtype, cudaStream_t s);
int launch_swap_pull(void *dst, const void *const *peers, int world, int rank, int n, const int *perm, int dtype, cudaStream_t s);
int launch_permute_bits(const void *src, void *dst, int n, const int *perm, int dtype, cudaStream_t s);
int launch_norm2(const void *state, int n, double *out_dev, int dtype, cudaStream_t s);
int launch_probabilities(const void *state, int n, uint64_t hi_base, const int *mes_pos, int m, double *probs_dev,
                         int dtype, cudaStream_t s);
int launch_sample_prepare(const void *state, int n, double *chunk_sums_dev, int dtype, cudaStream_t s);
int launch_sample_draw(const void *state, int n, const double *chunk_sums_dev, const double *uniforms_dev,
                       int64_t shots, uint64_t *indices_dev, int dtype, cudaStream_t s);
int sample_chunk_bits(int n);
int launch_prune(const double *probs_dev, int m, double ratio, unsigned long long *scratch_dev, uint64_t *idx_dev, double *val_dev,
                 uint64_t capacity, cudaStream_t s);

}  // namespace qb2
