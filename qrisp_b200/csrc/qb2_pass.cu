// qb2_pass.cu -- validation and dispatch of the fused multi-gate pass (see include/qrisp_b200.h), and
// the state preparation kernels.  The kernels themselves: qb2_pass_kernel.cuh (register-pipelined,
// every geometry), qb2_pass_tma.cuh (bulk tensor copies, complex64 T=13 r=5), instantiated one per
// translation unit by qb2_pass_inst.cu.
#include <cstdlib>
#include <cstring>

#include "qb2_internal.cuh"

namespace qb2 {

namespace {

// every offset / count of the small section is checked here, on the host: the kernel reads the
// descriptors through the constant bank without bounds checks
int validate_pass(const qb2_pass_header *h, int n, int dtype) {
    const uint32_t small = h->small_bytes;
    auto inside = [&](uint64_t off, uint64_t bytes) { return off >= sizeof(qb2_pass_header) && off + bytes <= small; };
    const int NR = 1 << h->r, TB = h->T - h->r;
    const size_t csize = dtype == QB2_C64 ? 8 : 16;
    if (!inside(QB2_IO_OFFSET, sizeof(qb2_io))) return set_error(QB2_EINVAL, "qb2_run_pass: io block outside the small section");
    if (!inside(h->phases_off, (uint64_t)h->n_phases * sizeof(qb2_phase)))
        return set_error(QB2_EINVAL, "qb2_run_pass: phases outside the small section");
    if (!inside(h->ops_off, (uint64_t)h->n_ops * sizeof(qb2_op)) && h->n_ops)
        return set_error(QB2_EINVAL, "qb2_run_pass: ops outside the small section");
    if (h->coef_off > small || h->u16_off > small || h->u64_off > small || h->tabdesc_off > small ||
        (h->coef_off & 15u) || (h->u64_off & 7u) || (h->u16_off & 1u) || (h->tabdesc_off & 7u))
        return set_error(QB2_EINVAL, "qb2_run_pass: descriptor sections outside the small section or misaligned");
    if (h->tab_off < small || h->tab_off > h->total_bytes) return set_error(QB2_EINVAL, "qb2_run_pass: table section misplaced");
    const uint8_t *pb = reinterpret_cast<const uint8_t *>(h);
    const qb2_phase *phases = reinterpret_cast<const qb2_phase *>(pb + h->phases_off);
    const qb2_op *ops = reinterpret_cast<const qb2_op *>(pb + h->ops_off);
    const uint32_t n16 = (small - h->u16_off) / 2u, n64 = (h->tabdesc_off > h->u64_off ? h->tabdesc_off - h->u64_off : small - h->u64_off) / 8u;
    const uint32_t ncoef = (uint32_t)((h->u64_off > h->coef_off ? h->u64_off - h->coef_off : 0u) / csize);
    const uint32_t ntab = (h->u16_off > h->tabdesc_off ? h->u16_off - h->tabdesc_off : 0u) / (uint32_t)sizeof(qb2_tab);
    for (uint32_t p = 0; p < h->n_phases; ++p) {
        const qb2_phase &ph = phases[p];
        if ((uint64_t)ph.first_op + ph.n_ops > h->n_ops) return set_error(QB2_EINVAL, "qb2_run_pass: phase %u names ops outside the pass", p);
        if ((uint64_t)ph.pcol + TB > n64 || (uint64_t)ph.rphys + NR > n64)
            return set_error(QB2_EINVAL, "qb2_run_pass: phase %u index arrays outside the u64 area", p);
        const bool maps = p > 0 || (h->features & QB2_FEATURE_TMA);
        if (maps && ((uint64_t)ph.xw + TB + NR > n16 || (uint64_t)ph.xr + TB + NR > n16))
            return set_error(QB2_EINVAL, "qb2_run_pass: phase %u slot maps outside the u16 area", p);
    }
    bool need_full = false;
    for (uint32_t o = 0; o < h->n_ops; ++o) {
        const qb2_op &op = ops[o];
        switch (op.kind) {
            case QB2_OP_MAT1: case QB2_OP_MAT2: case QB2_OP_MAT3: case QB2_OP_MAT4: {
                const uint32_t d = 1u << op.kind;
                if ((uint64_t)op.data + d * d > ncoef) return set_error(QB2_EINVAL, "qb2_run_pass: op %u coefficients outside the pass", o);
                if (op.kind != QB2_OP_MAT1) need_full = true;
                if (op.b0 >= h->r || (op.kind == QB2_OP_MAT2 && (op.b1 >= h->r || op.b1 <= op.b0)) || (int)op.kind > h->r)
                    return set_error(QB2_EINVAL, "qb2_run_pass: op %u register bits outside the phase", o);
                break;
            }
            case QB2_OP_DIAG: need_full = true;  // fall through
            case QB2_OP_DIAG_THR: case QB2_OP_DIAG_BIT:
                if ((uint64_t)op.data + op.ntab_thr + op.ntab_reg > ntab) return set_error(QB2_EINVAL, "qb2_run_pass: op %u tables outside the pass", o);
                if (op.kind == QB2_OP_DIAG_BIT && op.b0 >= h->r) return set_error(QB2_EINVAL, "qb2_run_pass: op %u register bit outside the phase", o);
                break;
            case QB2_OP_DIAG_VEC:
                if ((uint64_t)op.data + NR > ncoef) return set_error(QB2_EINVAL, "qb2_run_pass: op %u factors outside the pass", o);
                break;
            case QB2_OP_BFLY:
                if ((uint64_t)op.data + 4 + NR / 2 > ncoef || (uint64_t)op.tab + op.ntab_reg > ntab || op.b0 >= h->r)
                    return set_error(QB2_EINVAL, "qb2_run_pass: butterfly %u outside the pass", o);
                break;
            default:
                return set_error(QB2_EINVAL, "qb2_run_pass: op %u has unknown kind %u", o, op.kind);
        }
        if (op.code >= QB2_CODE_END) return set_error(QB2_EINVAL, "qb2_run_pass: op %u has unknown dispatch code %u", o, op.code);
        if ((op.kind == QB2_OP_BFLY || op.kind == QB2_OP_DIAG_THR || op.kind == QB2_OP_DIAG_BIT) && op.b1 > h->n_slots)
            return set_error(QB2_EINVAL, "qb2_run_pass: op %u names phase slot %u of %u", o, op.b1, h->n_slots);
    }
    if (need_full && !(h->features & QB2_FEATURE_FULL))
        return set_error(QB2_EINVAL, "qb2_run_pass: the pass holds MAT2/3/4 or general DIAG ops but does not ask for the full kernel");
    return QB2_OK;
}

typedef int (*variant_fn)(const PassLaunch &);

}  // namespace

int launch_pass(void *state, int n, uint64_t hi_base, const qb2_pass_header *h, const void *blob_dev, int dtype, cudaStream_t s,
                const void *sparse_dev, uint32_t n_sparse, const void *jit_kernel, double *tile_norms_dev) {
    if (!state || !h || !blob_dev) return set_error(QB2_EINVAL, "qb2_run_pass: null pointer");
    if (h->magic != QB2_PASS_MAGIC) return set_error(QB2_EINVAL, "qb2_run_pass: bad magic 0x%08x", h->magic);
    if ((int)h->dtype != dtype) return set_error(QB2_EINVAL, "qb2_run_pass: pass dtype %d != state dtype %d", h->dtype, dtype);
    if ((int)h->n != n) return set_error(QB2_EINVAL, "qb2_run_pass: pass built for n=%u, state has n=%d", h->n, n);
    if (h->T > QB2_MAX_TILE_BITS || h->T > n || h->r > h->T || h->r > QB2_MAX_REG_BITS)
        return set_error(QB2_ELIMIT, "qb2_run_pass: T=%u r=%u n=%d outside limits", h->T, h->r, n);
    if (h->n_seg > QB2_MAX_SEGS) return set_error(QB2_ELIMIT, "qb2_run_pass: too many segments");
    if (h->small_bytes > QB2_MAX_SMALL_BYTES || (h->small_bytes & 15u) || h->small_bytes < sizeof(qb2_pass_header))
        return set_error(QB2_ELIMIT, "qb2_run_pass: small section of %u bytes", h->small_bytes);
    if (h->n_phases < 1 || h->n_phases > QB2_MAX_PHASES)
        return set_error(QB2_EINVAL, "qb2_run_pass: %u phases (1..%d allowed)", h->n_phases, QB2_MAX_PHASES);
    if (((uintptr_t)blob_dev & 15u) != 0) return set_error(QB2_EINVAL, "qb2_run_pass: blob not 16-byte aligned");
    if (n - (int)h->T > 31) return set_error(QB2_ELIMIT, "qb2_run_pass: too many tiles");
    if (h->n_slots > QB2_MAX_SLOTS) return set_error(QB2_ELIMIT, "qb2_run_pass: %u phase slots (at most %d)", h->n_slots, QB2_MAX_SLOTS);
    if (int rc = validate_pass(h, n, dtype)) return rc;
    PassLaunch L;
    L.state = state;
    L.hi_base = hi_base;
    L.h = h;
    L.blob_dev = blob_dev;
    L.ntiles = 1u << (n - h->T);
    L.s = s;
    L.sparse.tile = nullptr;
    L.sparse.loc = nullptr;
    L.sparse.amp = nullptr;
    L.sparse.count = 0;
    L.sparse.tile_norms = tile_norms_dev;
    L.jit_kernel = jit_kernel;
    if (h->features & QB2_FEATURE_ZERO_INPUT) {
        if (n_sparse > QB2_MAX_SPARSE) return set_error(QB2_ELIMIT, "qb2_run_pass: %u sparse input entries (at most %d)", n_sparse, QB2_MAX_SPARSE);
        if (n_sparse && !sparse_dev) return set_error(QB2_EINVAL, "qb2_run_pass: sparse input list missing");
        if (((uintptr_t)sparse_dev & 15u) != 0) return set_error(QB2_EINVAL, "qb2_run_pass: sparse input list not 16-byte aligned");
        // layout of the list: uint32 tile[count], uint32 loc[count] (each padded to 16 bytes), then the amplitudes
        const size_t pad = ((size_t)n_sparse * 4 + 15) & ~(size_t)15;
        const uint8_t *sp = reinterpret_cast<const uint8_t *>(sparse_dev);
        L.sparse.tile = reinterpret_cast<const uint32_t *>(sp);
        L.sparse.loc = reinterpret_cast<const uint32_t *>(sp + pad);
        L.sparse.amp = sp + 2 * pad;
        L.sparse.count = n_sparse;
    }
    const int tb = h->T - h->r;
    const int full = (h->features & QB2_FEATURE_FULL) ? 1 : 0;
    static const variant_fn table[QB2_N_PASS_VARIANTS] = {
        launch_pass_variant_0, launch_pass_variant_1, launch_pass_variant_2,  launch_pass_variant_3,  launch_pass_variant_4,
        launch_pass_variant_5, launch_pass_variant_6, launch_pass_variant_7,  launch_pass_variant_8,  launch_pass_variant_9,
        launch_pass_variant_10, launch_pass_variant_11, launch_pass_variant_12, launch_pass_variant_13};
    static const bool no_tma = getenv("QB2_NO_TMA") != nullptr;
    if ((h->features & QB2_FEATURE_TMA) && !no_tma) {
        if (dtype != QB2_C64 || h->T != 13 || h->r != 5)
            return set_error(QB2_EINVAL, "qb2_run_pass: QB2_FEATURE_TMA needs complex64, T=13, r=5");
        const int rc = table[12 + full](L);
        if (rc != QB2_ELIMIT || jit_kernel) return rc;
        // the tile did not fit a tensor map (or the driver lacks the encoder): the other kernel runs the
        // same pass from the same descriptors
    }
    int geom = -1;
    if (dtype == QB2_C64) {
        if (h->r == 4 && tb <= 8) geom = 0;
        else if (h->r == 4 && tb == 9) geom = 1;
        else if (h->r == 5 && tb <= 8) geom = 2;
        else if (h->r == 3 && tb <= 8) geom = 3;
    } else if (dtype == QB2_C128) {
        if (h->r == 3 && tb <= 8) geom = 4;
        else if (h->r == 4 && tb <= 8) geom = 5;
    }
    if (geom < 0) return set_error(QB2_ELIMIT, "qb2_run_pass: no kernel variant for dtype=%d r=%u T=%u", dtype, h->r, h->T);
    // the register-pipelined kernel does not write tile norms: a separate read of the state fills them (the caller
    // asked for them under the contract "tile t = amplitudes [t << T, (t + 1) << T)")
    L.sparse.tile_norms = nullptr;
    if (int rc = table[2 * geom + full](L)) return rc;
    if (tile_norms_dev) return launch_chunk_sums(state, n, (int)h->T, tile_norms_dev, dtype, s);
    return QB2_OK;
}

namespace {

template <typename real>
__global__ void init_basis_kernel(typename Traits<real>::vec *__restrict__ state, const uint64_t size, const uint64_t basis) {
    using vec = typename Traits<real>::vec;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += stride) {
        vec v;
        v.x = (i == basis) ? (real)1 : (real)0;
        v.y = 0;
        state[i] = v;
    }
}

template <typename real>
__global__ void scatter_kernel(typename Traits<real>::vec *__restrict__ state, const uint64_t *__restrict__ idx,
                               const typename Traits<real>::vec *__restrict__ amp, const uint32_t count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) state[idx[i]] = amp[i];
}

}  // namespace

int launch_init_basis(void *state, int n, uint64_t basis, int dtype, cudaStream_t s) {
    if (!state || n < 0 || n > 40) return set_error(QB2_EINVAL, "qb2_init_basis: bad arguments");
    const uint64_t size = 1ull << n;
    if (basis != QB2_NO_BASIS && basis >= size) return set_error(QB2_EINVAL, "qb2_init_basis: basis out of range");
    uint64_t blocks = (size + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (dtype == QB2_C64)
        init_basis_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<float2 *>(state), size, basis);
    else if (dtype == QB2_C128)
        init_basis_kernel<double><<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<double2 *>(state), size, basis);
    else
        return set_error(QB2_EINVAL, "qb2_init_basis: bad dtype %d", dtype);
    count_launch();
    return check_cuda(cudaGetLastError(), "init_basis launch");
}

int launch_init_sparse(void *state, int n, const uint64_t *idx_dev, const void *amp_dev, uint32_t count, int dtype,
                       cudaStream_t s) {
    if (int rc = launch_init_basis(state, n, QB2_NO_BASIS, dtype, s)) return rc;
    if (count == 0) return QB2_OK;
    if (!idx_dev || !amp_dev) return set_error(QB2_EINVAL, "qb2_init_sparse: null list");
    const unsigned blocks = (count + 255u) / 256u;
    if (dtype == QB2_C64)
        scatter_kernel<float><<<blocks, 256, 0, s>>>(reinterpret_cast<float2 *>(state), idx_dev,
                                                     reinterpret_cast<const float2 *>(amp_dev), count);
    else
        scatter_kernel<double><<<blocks, 256, 0, s>>>(reinterpret_cast<double2 *>(state), idx_dev,
                                                      reinterpret_cast<const double2 *>(amp_dev), count);
    count_launch();
    return check_cuda(cudaGetLastError(), "init_sparse launch");
}

}  // namespace qb2
