// qb2_block_tc.cu -- dense 5-qubit blocks on the 5th-generation tensor cores (tcgen05.mma + TMEM).
//
// Replaces, for k = 5 dense targets, the reference's grouped-unitary contraction (circuit_preprocessing.py:119-157
// `group_qc` -> tensor_factor.py:67-89 `apply_matrix` -> bi_arrays.py:1175-1299 `tensordot`).  A 32 x 32
// complex block costs 32 complex multiply-adds per amplitude: 64 packed FP32 instructions on the SIMT pipe,
// which is 4 ms of pure FP32 issue per sweep at 30 qubits -- a dense contraction, the one place where the
// profile says tensor cores (north star).  Here it is a real GEMM
//
//     D[m][n] = sum_k A[m][k] * B[n][k]        M = 128 state columns per tile, N = K = 64
//
//   A[m][2c + (re|im)]  = the 32 target amplitudes of state column m (interleaved re, im: the layout HBM has)
//   B[2r + re][2c + re] = Re U[r][c]     B[2r + re][2c + im] = -Im U[r][c]
//   B[2r + im][2c + re] = Im U[r][c]     B[2r + im][2c + im] =  Re U[r][c]
//
// so that row m of D is the column's 32 new amplitudes, interleaved again: a thread owns one column from the
// global load to the global store and the update is in place.  complex64 accuracy comes from the 3 x TF32
// split: x = hi + lo with hi = the top 19 bits (exactly a TF32 number) and lo = x - hi (exact in FP32);
// D = A_hi B_hi + A_lo B_hi + A_hi B_lo leaves an error of 2**-21 relative per product (tested at 1e-5
// absolute against the oracle, measured ~1e-7).  B's halves are split once on the host (qb2_pack_block_tc).
//
// One CTA per SM, three independent groups of 128 threads (one thread per TMEM lane; named barriers): a
// group's threads load their columns with coalesced 8-byte loads, split, and write both halves into the
// group's shared-memory tiles in the canonical K-major no-swizzle UMMA layout (8-row x 16-byte core
// matrices; LBO = 128 B between K chunks, SBO = 2 KB between row groups); the group's first thread issues
// the 24 tcgen05.mma (8 K-steps x 3 products) into the group's 64 TMEM columns and commits to the group's
// mbarrier; every warp reads its 32 lanes of the accumulator back with tcgen05.ld and stores the column.
// The groups share the block's operand (32 KB) and run out of step, so loads, tensor work and stores of
// three tiles overlap on every SM (224 KB of shared memory, 192 of the 512 TMEM columns).
#include <cstring>

#include "qb2_internal.cuh"

namespace qb2 {
namespace {

constexpr int TC_K = 5;                              // qubits of a block
constexpr int TC_D = 1 << TC_K;                      // 32 complex rows
constexpr int TC_R = 2 * TC_D;                       // 64: N and K of the real GEMM
constexpr int TC_M = 128;                            // state columns per tile = threads per CTA = TMEM lanes
constexpr uint32_t TC_LBO = 128;                     // bytes between the 16-byte K chunks of one 8-row group
constexpr uint32_t TC_SBO = (TC_R / 4) * TC_LBO;     // 2048: bytes between 8-row groups
constexpr uint32_t TC_A_BYTES = TC_M * TC_R * 4;     // 32 KB per half
constexpr uint32_t TC_B_BYTES = TC_R * TC_R * 4;     // 16 KB per half
constexpr int TC_G = 3;                              // groups per CTA
constexpr uint32_t TC_SMEM = TC_G * 2 * TC_A_BYTES + 2 * TC_B_BYTES + 64;
constexpr uint32_t TC_TMEM_COLS = 256;               // allocation: a power of two >= TC_G * 64

// element (row, k) of an operand tile in the canonical layout
__host__ __device__ constexpr uint32_t tc_off(uint32_t row, uint32_t k) {
    return (row >> 3) * TC_SBO + (k >> 2) * TC_LBO + (row & 7u) * 16u + (k & 3u) * 4u;
}

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start address, leading and
// stride byte offsets in 16-byte units, version 1 (Blackwell), no swizzle, K-major
__device__ __forceinline__ uint64_t umma_desc(const uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(TC_LBO >> 4) << 16) | ((uint64_t)(TC_SBO >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N = 64, M = 128
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_R >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ void umma_tf32(const uint32_t d_tmem, const uint64_t a, const uint64_t b, const uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a), "l"(b), "r"(TC_IDESC), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tc_mbar_wait(const uint32_t bar, const uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ float tf32_rna(const float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

struct BlockTargets {
    int t[TC_K];  // ascending physical positions; matrix index bit i <-> t[i]
};

__global__ void __launch_bounds__(TC_G *TC_M, 1)
    block5_tc_kernel(float2 *__restrict__ state, const uint64_t hi_base, const float *__restrict__ gate, const BlockTargets T,
                     const uint64_t cmask, const uint64_t cval, const uint32_t ntiles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t g = threadIdx.x / TC_M, tid = threadIdx.x % TC_M, warp = tid >> 5;
    uint8_t *const b_hi = smem;
    uint8_t *const b_lo = smem + TC_B_BYTES;
    uint8_t *const a_hi = smem + 2 * TC_B_BYTES + g * 2 * TC_A_BYTES;
    uint8_t *const a_lo = a_hi + TC_A_BYTES;
    uint64_t *const bars = reinterpret_cast<uint64_t *>(smem + 2 * TC_B_BYTES + TC_G * 2 * TC_A_BYTES);
    uint32_t *const tmem_slot = reinterpret_cast<uint32_t *>(bars + TC_G);
    const uint32_t bar = smem_addr(bars + g);

    // the block's two halves (already in the canonical layout: qb2_pack_block_tc), once per CTA
    for (uint32_t i = threadIdx.x; i < 2 * TC_B_BYTES / 16; i += TC_G * TC_M)
        reinterpret_cast<uint4 *>(b_hi)[i] = __ldg(reinterpret_cast<const uint4 *>(gate) + i);
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "n"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot + g * (uint32_t)TC_R;  // this group's 64 accumulator columns
    auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(g + 1u), "n"(TC_M) : "memory"); };

    uint64_t st[TC_K];
#pragma unroll
    for (int i = 0; i < TC_K; ++i) st[i] = 1ull << T.t[i];
    const uint32_t row_off = (tid >> 3) * TC_SBO + (tid & 7u) * 16u;
    const uint64_t da_hi = umma_desc(smem_addr(a_hi)), da_lo = umma_desc(smem_addr(a_lo));
    const uint64_t db_hi = umma_desc(smem_addr(b_hi)), db_lo = umma_desc(smem_addr(b_lo));
    uint32_t parity = 0;

    for (uint32_t tile = blockIdx.x * TC_G + g; tile < ntiles; tile += gridDim.x * TC_G) {
        // this thread's column: the tile's columns are consecutive in the non-target bits (coalesced)
        uint64_t idx = (uint64_t)tile * TC_M + tid;
#pragma unroll
        for (int i = 0; i < TC_K; ++i) {
            const int p = T.t[i];
            idx = ((idx >> p) << (p + 1)) | (idx & ((1ull << p) - 1ull));
        }
        // a column the controls switch off is neither loaded nor stored (its accumulator row is garbage); a tile
        // in which no column is active (controls above the tile's column bits: the usual case) is skipped by the
        // whole group -- one barrier with an OR reduction
        const bool active = (((idx | hi_base) ^ cval) & cmask) == 0ull;
        if (cmask != 0ull) {
            uint32_t any;
            asm volatile(
                "{\n"
                ".reg .pred p, q;\n"
                "setp.ne.u32 q, %1, 0;\n"
                "bar.red.or.pred p, %2, %3, q;\n"
                "selp.u32 %0, 1, 0, p;\n"
                "}\n"
                : "=r"(any)
                : "r"((uint32_t)active), "r"(g + 1u), "n"(TC_M)
                : "memory");
            if (!any) continue;
        }
        float2 *const base = state + idx;
        if (active) {
#pragma unroll
            for (int c = 0; c < TC_D; c += 2) {
                uint64_t off = 0;
#pragma unroll
                for (int i = 1; i < TC_K; ++i)
                    if ((c >> i) & 1) off += st[i];
                const float2 x0 = base[off], x1 = base[off + st[0]];
                float4 hi, lo;
                hi.x = __uint_as_float(__float_as_uint(x0.x) & 0xffffe000u);
                hi.y = __uint_as_float(__float_as_uint(x0.y) & 0xffffe000u);
                hi.z = __uint_as_float(__float_as_uint(x1.x) & 0xffffe000u);
                hi.w = __uint_as_float(__float_as_uint(x1.y) & 0xffffe000u);
                // lo is rounded to the nearest TF32 here: the tensor core would truncate it, a bias towards zero
                // that shrinks the norm by ~5e-7 per block
                lo.x = tf32_rna(x0.x - hi.x), lo.y = tf32_rna(x0.y - hi.y), lo.z = tf32_rna(x1.x - hi.z), lo.w = tf32_rna(x1.y - hi.w);
                // k = 2c .. 2c+3: one 16-byte chunk of the row
                *reinterpret_cast<float4 *>(a_hi + row_off + (c >> 1) * TC_LBO) = hi;
                *reinterpret_cast<float4 *>(a_lo + row_off + (c >> 1) * TC_LBO) = lo;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        group_sync();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (uint32_t s = 0; s < TC_R / 8; ++s) {  // one instruction covers K = 8 = two 16-byte chunks
                const uint64_t adv = (uint64_t)((s * 2u * TC_LBO) >> 4);
                umma_tf32(tmem, da_hi + adv, db_hi + adv, s > 0 ? 1u : 0u);
                umma_tf32(tmem, da_lo + adv, db_hi + adv, 1u);
                umma_tf32(tmem, da_hi + adv, db_lo + adv, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        tc_mbar_wait(bar, parity);
        parity ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- the column's new amplitudes: lane `tid` of the accumulator, 64 columns
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t r[32];
            const uint32_t taddr = tmem + ((warp * 32u) << 16) + (uint32_t)h * 32u;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (active) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int c = h * 16 + j;
                    uint64_t off = 0;
#pragma unroll
                    for (int i = 0; i < TC_K; ++i)
                        if ((c >> i) & 1) off += st[i];
                    base[off] = make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                }
            }
        }
        // the accumulator and both A halves are free again once every warp of the group has read its lanes
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        group_sync();
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "n"(TC_TMEM_COLS) : "memory");
}

}  // namespace

int launch_block_tc(void *state, int n, uint64_t hi_base, const void *gate_dev, const int *targets, uint64_t cmask, uint64_t cval,
                    cudaStream_t s) {
    if (n < TC_K + 7) return set_error(QB2_EINVAL, "qb2_apply_block_tc: needs at least %d local qubits", TC_K + 7);
    BlockTargets T;
    for (int i = 0; i < TC_K; ++i) {
        T.t[i] = targets[i];
        if (targets[i] < 0 || targets[i] >= n || (i > 0 && targets[i] <= targets[i - 1]))
            return set_error(QB2_EINVAL, "qb2_apply_block_tc: targets must be ascending local positions");
        if ((cmask >> targets[i]) & 1ull) return set_error(QB2_EINVAL, "qb2_apply_block_tc: a target is also a control");
    }
    static bool attr_set[QB2_MAX_DEVICES] = {false};
    int dev = 0;
    QB2_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= QB2_MAX_DEVICES) return set_error(QB2_ELIMIT, "device ordinal %d outside the engine's table", dev);
    if (!attr_set[dev]) {
        QB2_CUDA(cudaFuncSetAttribute(block5_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        attr_set[dev] = true;
    }
    const uint32_t ntiles = 1u << (n - TC_K - 7);
    uint32_t grid = (uint32_t)sm_count();
    if (grid > (ntiles + TC_G - 1) / TC_G) grid = (ntiles + TC_G - 1) / TC_G;
    block5_tc_kernel<<<grid, TC_G * TC_M, TC_SMEM, s>>>(reinterpret_cast<float2 *>(state), hi_base, reinterpret_cast<const float *>(gate_dev), T,
                                                 cmask, cval, ntiles);
    count_launch();
    return check_cuda(cudaGetLastError(), "block5_tc_kernel launch");
}

// ---------------------------------------------------------------- multiplexed real rotation
// The middle factor of a cosine-sine decomposition: on `target`, the rotation [[c_j, -s_j], [s_j, c_j]] where j
// is spelled by the `nc` control positions (a uniformly controlled Ry).  Streaming, HBM bound: 16 B per
// amplitude.  With it a dense block on k = 6..8 qubits becomes controlled 5-qubit blocks on the tensor cores
// and k - 5 of these sweeps (engine.py: _run_block_csd).
namespace {

struct MuxDesc {
    int target;
    int nc;
    int ctrl[QB2_MUX_MAX_CONTROLS];
};

template <typename vec, typename real>
__global__ void __launch_bounds__(256) mux_ry_kernel(vec *__restrict__ state, const uint64_t pairs, const uint64_t hi_base,
                                                     const MuxDesc D, const vec *__restrict__ cs, const uint64_t cmask,
                                                     const uint64_t cval) {
    extern __shared__ __align__(16) uint8_t mux_smem[];
    vec *const tab = reinterpret_cast<vec *>(mux_smem);
    for (uint32_t i = threadIdx.x; i < (1u << D.nc); i += blockDim.x) tab[i] = cs[i];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t tbit = 1ull << D.target;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < pairs; p += stride) {
        const uint64_t i0 = ((p >> D.target) << (D.target + 1)) | (p & (tbit - 1ull));
        const uint64_t x = i0 | hi_base;
        if (((x ^ cval) & cmask) != 0ull) continue;
        uint32_t j = 0;
#pragma unroll
        for (int c = 0; c < QB2_MUX_MAX_CONTROLS; ++c)
            if (c < D.nc) j |= (uint32_t)((x >> D.ctrl[c]) & 1ull) << c;
        const vec w = tab[j];  // (cos, sin)
        const vec a0 = state[i0], a1 = state[i0 | tbit];
        vec b0, b1;
        b0.x = w.x * a0.x - w.y * a1.x;
        b0.y = w.x * a0.y - w.y * a1.y;
        b1.x = w.y * a0.x + w.x * a1.x;
        b1.y = w.y * a0.y + w.x * a1.y;
        state[i0] = b0;
        state[i0 | tbit] = b1;
    }
}

}  // namespace

int launch_mux_ry(void *state, int n, uint64_t hi_base, int target, const int *ctrl, int nc, const void *cs_dev, uint64_t cmask,
                  uint64_t cval, int dtype, cudaStream_t s) {
    if (!state || !cs_dev || n < 1 || n > 40 || target < 0 || target >= n || nc < 0 || nc > QB2_MUX_MAX_CONTROLS || (nc && !ctrl))
        return set_error(QB2_EINVAL, "qb2_apply_mux_ry: bad arguments");
    MuxDesc D;
    D.target = target;
    D.nc = nc;
    for (int c = 0; c < QB2_MUX_MAX_CONTROLS; ++c) D.ctrl[c] = 0;
    for (int c = 0; c < nc; ++c) {
        if (ctrl[c] < 0 || ctrl[c] >= 64 || ctrl[c] == target) return set_error(QB2_EINVAL, "qb2_apply_mux_ry: bad control position");
        D.ctrl[c] = ctrl[c];
    }
    if ((cmask >> target) & 1ull) return set_error(QB2_EINVAL, "qb2_apply_mux_ry: the target is also a control");
    const uint64_t pairs = 1ull << (n - 1);
    uint64_t blocks = (pairs + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (dtype == QB2_C64)
        mux_ry_kernel<float2, float><<<(unsigned)blocks, 256, sizeof(float2) << nc, s>>>(reinterpret_cast<float2 *>(state), pairs, hi_base, D,
                                                                                   reinterpret_cast<const float2 *>(cs_dev), cmask, cval);
    else if (dtype == QB2_C128)
        mux_ry_kernel<double2, double><<<(unsigned)blocks, 256, sizeof(double2) << nc, s>>>(reinterpret_cast<double2 *>(state), pairs, hi_base, D,
                                                                                      reinterpret_cast<const double2 *>(cs_dev), cmask, cval);
    else
        return set_error(QB2_EINVAL, "qb2_apply_mux_ry: bad dtype %d", dtype);
    count_launch();
    return check_cuda(cudaGetLastError(), "mux_ry launch");
}

// host: U (32 x 32 complex128, row-major, index bit i <-> targets[i]) -> B_hi, B_lo in the kernel's layout
int pack_block_tc(const double *u, float *out) {
    float *hi = out, *lo = out + TC_R * TC_R;
    for (uint32_t nn = 0; nn < (uint32_t)TC_R; ++nn)
        for (uint32_t kk = 0; kk < (uint32_t)TC_R; ++kk) {
            const uint32_t r = nn >> 1, c = kk >> 1;
            const double re = u[2 * (r * TC_D + c)], im = u[2 * (r * TC_D + c) + 1];
            const double v = (nn & 1u) == 0u ? ((kk & 1u) == 0u ? re : -im) : ((kk & 1u) == 0u ? im : re);
            const float f = (float)v;
            uint32_t bits;
            memcpy(&bits, &f, 4);
            bits &= 0xffffe000u;
            float h;
            memcpy(&h, &bits, 4);
            const uint32_t o = tc_off(nn, kk) / 4u;
            hi[o] = h;
            float l = (float)(v - (double)h);  // rounded to the nearest TF32 (the tensor core would truncate)
            memcpy(&bits, &l, 4);
            bits = (bits + 0x1000u) & 0xffffe000u;
            memcpy(&l, &bits, 4);
            lo[o] = l;
        }
    return QB2_OK;
}

}  // namespace qb2
