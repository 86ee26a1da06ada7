// qb2_api.cu -- the extern "C" boundary of libqb2.so (include/qrisp_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstring>

#include <cuda.h>  // CUmodule / CUfunction types; entry points come from cudaGetDriverEntryPoint

#include "qb2_internal.cuh"

namespace qb2 {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return QB2_OK;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
        return set_error(QB2_ENODEV, "%s: %s", what, cudaGetErrorString(e));
    return set_error(QB2_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {  // of the CURRENT device (cached per device ordinal)
    static int cached[QB2_MAX_DEVICES] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= QB2_MAX_DEVICES) return 148;
    if (cached[dev] == 0) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
            cached[dev] = sms;
        else
            return 148;
    }
    return cached[dev];
}

static int check_dtype(int dtype, const char *who) {
    if (dtype != QB2_C64 && dtype != QB2_C128) return set_error(QB2_EINVAL, "%s: bad dtype %d", who, dtype);
    return QB2_OK;
}

}  // namespace qb2

using namespace qb2;

extern "C" {

int qb2_abi_version(void) { return QB2_ABI_VERSION; }

const char *qb2_last_error(void) { return g_err; }

int qb2_device_info(int *sm_count_out, int *smem_optin_bytes, size_t *total_mem_bytes) {
    int dev = 0;
    QB2_CUDA(cudaGetDevice(&dev));
    int sms = 0, smem = 0;
    QB2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    QB2_CUDA(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    size_t free_b = 0, total_b = 0;
    QB2_CUDA(cudaMemGetInfo(&free_b, &total_b));
    if (sm_count_out) *sm_count_out = sms;
    if (smem_optin_bytes) *smem_optin_bytes = smem;
    if (total_mem_bytes) *total_mem_bytes = total_b;
    return QB2_OK;
}

uint64_t qb2_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void qb2_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }

int qb2_malloc(void **dev_ptr, size_t bytes) {
    if (!dev_ptr) return set_error(QB2_EINVAL, "qb2_malloc: null pointer");
    QB2_CUDA(cudaMalloc(dev_ptr, bytes));
    return QB2_OK;
}
int qb2_free(void *dev_ptr) {
    QB2_CUDA(cudaFree(dev_ptr));
    return QB2_OK;
}
int qb2_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes, void *stream) {
    QB2_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return QB2_OK;
}
int qb2_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes, void *stream) {
    QB2_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return QB2_OK;
}
int qb2_stream_sync(void *stream) {
    QB2_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return QB2_OK;
}

int qb2_init_basis(void *state, int n, uint64_t basis, int dtype, void *stream) {
    if (int rc = check_dtype(dtype, "qb2_init_basis")) return rc;
    return launch_init_basis(state, n, basis, dtype, (cudaStream_t)stream);
}

int qb2_init_sparse(void *state, int n, const uint64_t *idx_dev, const void *amp_dev, uint32_t count, int dtype,
                    void *stream) {
    if (int rc = check_dtype(dtype, "qb2_init_sparse")) return rc;
    return launch_init_sparse(state, n, idx_dev, amp_dev, count, dtype, (cudaStream_t)stream);
}

int qb2_run_pass(void *state, int n, uint64_t hi_base, const void *header, const void *blob_dev, int dtype,
                 void *stream, const void *sparse_dev, uint32_t n_sparse) {
    if (int rc = check_dtype(dtype, "qb2_run_pass")) return rc;
    return launch_pass(state, n, hi_base, reinterpret_cast<const qb2_pass_header *>(header), blob_dev, dtype,
                       (cudaStream_t)stream, sparse_dev, n_sparse);
}

int qb2_run_pass_ex(void *state, int n, uint64_t hi_base, const void *header, const void *blob_dev, int dtype, void *stream,
                    const void *sparse_dev, uint32_t n_sparse, void *jit_kernel, double *tile_norms_dev) {
    if (int rc = check_dtype(dtype, "qb2_run_pass_ex")) return rc;
    return launch_pass(state, n, hi_base, reinterpret_cast<const qb2_pass_header *>(header), blob_dev, dtype,
                       (cudaStream_t)stream, sparse_dev, n_sparse, jit_kernel, tile_norms_dev);
}

int qb2_sample_scan(double *sums_dev, uint64_t count, void *stream) { return launch_scan_sums(sums_dev, count, (cudaStream_t)stream); }

int qb2_sample_draw_chunks(const void *state, int n, int chunk_bits, const double *chunk_sums_dev, const double *uniforms_dev,
                           int64_t shots, uint64_t *indices_dev, int dtype, void *stream) {
    if (int rc = check_dtype(dtype, "qb2_sample_draw_chunks")) return rc;
    return launch_sample_draw_cb(state, n, chunk_bits, chunk_sums_dev, uniforms_dev, shots, indices_dev, dtype, (cudaStream_t)stream);
}

int qb2_jit_load(const char *cubin_path, void **kernel_out) {
    if (!cubin_path || !kernel_out) return set_error(QB2_EINVAL, "qb2_jit_load: null pointer");
    // driver API through the runtime's entry-point lookup (no link against libcuda): the module lives in the
    // primary context of the current device, which a runtime call makes current
    typedef CUresult (*load_fn)(CUmodule *, const char *);
    typedef CUresult (*getf_fn)(CUfunction *, CUmodule, const char *);
    static load_fn mod_load = nullptr;
    static getf_fn mod_get = nullptr;
    if (!mod_load || !mod_get) {
        cudaDriverEntryPointQueryResult q;
        void *f1 = nullptr, *f2 = nullptr;
        if (cudaGetDriverEntryPoint("cuModuleLoad", &f1, cudaEnableDefault, &q) != cudaSuccess || !f1 ||
            cudaGetDriverEntryPoint("cuModuleGetFunction", &f2, cudaEnableDefault, &q) != cudaSuccess || !f2)
            return set_error(QB2_ECUDA, "qb2_jit_load: the driver has no module loader");
        mod_load = reinterpret_cast<load_fn>(f1);
        mod_get = reinterpret_cast<getf_fn>(f2);
    }
    QB2_CUDA(cudaFree(nullptr));
    CUmodule mod;
    CUresult r = mod_load(&mod, cubin_path);
    if (r != CUDA_SUCCESS) return set_error(QB2_ECUDA, "qb2_jit_load: cuModuleLoad(%s) failed with %d", cubin_path, (int)r);
    CUfunction fn;
    r = mod_get(&fn, mod, "qb2_pass_jit");
    if (r != CUDA_SUCCESS) return set_error(QB2_ECUDA, "qb2_jit_load: no kernel qb2_pass_jit in %s (%d)", cubin_path, (int)r);
    *kernel_out = (void *)fn;
    return QB2_OK;
}

int qb2_run_pass_jit(void *kernel, void *state, int n, uint64_t hi_base, const void *header, const void *blob_dev,
                     int dtype, void *stream, const void *sparse_dev, uint32_t n_sparse) {
    if (!kernel) return set_error(QB2_EINVAL, "qb2_run_pass_jit: null kernel");
    if (int rc = check_dtype(dtype, "qb2_run_pass_jit")) return rc;
    return launch_pass(state, n, hi_base, reinterpret_cast<const qb2_pass_header *>(header), blob_dev, dtype,
                       (cudaStream_t)stream, sparse_dev, n_sparse, kernel);
}

int qb2_apply_matrix_big(const void *src, void *dst, int n, uint64_t hi_base, const void *matrix_dev,
                         const int *targets, int k, uint64_t ctrl_mask, uint64_t ctrl_value, int dtype,
                         void *stream) {
    if (int rc = check_dtype(dtype, "qb2_apply_matrix_big")) return rc;
    return launch_matrix_big(src, dst, n, hi_base, matrix_dev, targets, k, ctrl_mask, ctrl_value, dtype,
                             (cudaStream_t)stream);
}

int qb2_apply_block_tc(void *state, int n, uint64_t hi_base, const void *block_dev, const int *targets,
                       uint64_t ctrl_mask, uint64_t ctrl_value, void *stream) {
    if (!state || !block_dev || !targets) return set_error(QB2_EINVAL, "qb2_apply_block_tc: null pointer");
    return launch_block_tc(state, n, hi_base, block_dev, targets, ctrl_mask, ctrl_value, (cudaStream_t)stream);
}

int qb2_apply_mux_ry(void *state, int n, uint64_t hi_base, int target, const int *ctrl, int n_ctrl, const void *cs_dev,
                     uint64_t ctrl_mask, uint64_t ctrl_value, int dtype, void *stream) {
    if (int rc = check_dtype(dtype, "qb2_apply_mux_ry")) return rc;
    return launch_mux_ry(state, n, hi_base, target, ctrl, n_ctrl, cs_dev, ctrl_mask, ctrl_value, dtype, (cudaStream_t)stream);
}

int qb2_pack_block_tc(const void *matrix_c128_host, void *out_host) {
    if (!matrix_c128_host || !out_host) return set_error(QB2_EINVAL, "qb2_pack_block_tc: null pointer");
    return pack_block_tc(reinterpret_cast<const double *>(matrix_c128_host), reinterpret_cast<float *>(out_host));
}

int qb2_swap_pull(void *dst, const void *const *peer_states, int world, int rank, int n, const int *perm, int dtype, void *stream) {
    if (int rc = check_dtype(dtype, "qb2_swap_pull")) return rc;
    return launch_swap_pull(dst, peer_states, world, rank, n, perm, dtype, (cudaStream_t)stream);
}

int qb2_permute_bits(const void *src, void *dst, int n, const int *perm, int dtype, void *stream) {
    if (int rc = check_dtype(dtype, "qb2_permute_bits")) return rc;
    return launch_permute_bits(src, dst, n, perm, dtype, (cudaStream_t)stream);
}

int qb2_norm2(const void *state, int n, double *out_dev, int dtype, void *stream) {
    if (int rc = check_dtype(dtype, "qb2_norm2")) return rc;
    return launch_norm2(state, n, out_dev, dtype, (cudaStream_t)stream);
}

int qb2_probabilities(const void *state, int n, uint64_t hi_base, const int *mes_pos, int m, double *probs_dev,
                      int dtype, void *stream) {
    if (int rc = check_dtype(dtype, "qb2_probabilities")) return rc;
    return launch_probabilities(state, n, hi_base, mes_pos, m, probs_dev, dtype, (cudaStream_t)stream);
}

int qb2_prune_probabilities(const double *probs_dev, int m, double cutoff_ratio, uint64_t *scratch_dev, uint64_t *idx_dev,
                            double *val_dev, uint64_t capacity, void *stream) {
    return launch_prune(probs_dev, m, cutoff_ratio, reinterpret_cast<unsigned long long *>(scratch_dev), idx_dev, val_dev, capacity,
                        (cudaStream_t)stream);
}

int qb2_sample_chunk_bits(int n) { return sample_chunk_bits(n); }

int qb2_sample_prepare(const void *state, int n, double *chunk_sums_dev, int dtype, void *stream) {
    if (int rc = check_dtype(dtype, "qb2_sample_prepare")) return rc;
    return launch_sample_prepare(state, n, chunk_sums_dev, dtype, (cudaStream_t)stream);
}

int qb2_sample_draw(const void *state, int n, const double *chunk_sums_dev, const double *uniforms_dev,
                    int64_t shots, uint64_t *indices_dev, int dtype, void *stream) {
    if (int rc = check_dtype(dtype, "qb2_sample_draw")) return rc;
    return launch_sample_draw(state, n, chunk_sums_dev, uniforms_dev, shots, indices_dev, dtype, (cudaStream_t)stream);
}

}  // extern "C"
