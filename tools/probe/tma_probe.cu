// tma_probe.cu -- development probe (not part of libqb2.so): does the tile traffic of a pass
// run at HBM speed when the tiles move with ONE bulk tensor copy (TMA) each?
//
// Skeleton of the TMA pass kernel: one CTA per SM, two compute groups of 256 threads; a 64 KB
// landing buffer (cp.async.bulk.tensor load, mbarrier), one 64 KB buffer per group that is the
// source of the bulk tensor store.  A tile is any set of T=13 index bits that contains bits 0..3
// (whole 128-byte lines); the tensor map describes it as dim0 = bits 0..3, one "base" dimension
// of box size 1 whose coordinate carries all non-tile bits, and one dimension per run of tile bits.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tma_probe tma_probe.cu
//   ./tma_probe n fma_per_amp swizzle(0/1) overlap(0/1) tilebits...   e.g.  ./tma_probe 30 16 1 1 12 13 14 15 16 17 18 19 20
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

constexpr int T = 13, NR = 32, GROUP = 256, TILE_BYTES = 8 << T;

struct Seg { uint8_t src, dst, len, rsv; };
struct Params {
    Seg seg[16];
    int n_seg;
    int ndim;            // tensor rank
    int base_dim;        // which coordinate carries the non-tile bits (>> base_shift)
    int base_shift;
    uint32_t ntiles;
    int swz;
    int rlow;            // 1: registers = tile bits 0..4 at the store (worst case for the banks)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load5(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_store5(const CUtensorMap *m, const void *src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(GROUP) : "memory"); }

// byte offset of tile index t in a buffer written by the TMA (128-byte rows, optional 128B swizzle)
__device__ __forceinline__ uint32_t tile_off(uint32_t t, int swz) {
    const uint32_t row = t >> 4, chunk = (t >> 1) & 7u;
    return row * 128u + (((swz ? (chunk ^ (row & 7u)) : chunk)) << 4) + (t & 1u) * 8u;
}

template <int NFMA>
__global__ void __launch_bounds__(2 * GROUP, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Params P, float wr, float wi) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *landing = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);  // 128B swizzle atoms are 1024-byte aligned
    uint8_t *gbuf0 = landing + TILE_BYTES;
    __shared__ __align__(8) uint64_t full_bar[2];
    const int tid = threadIdx.x & (GROUP - 1);
    const int g = threadIdx.x >> 8;
    uint8_t *gbuf = gbuf0 + g * TILE_BYTES;
    if (threadIdx.x == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();
    auto coords = [&](uint32_t t, int (&c)[5]) {
        uint64_t tb = 0;
        for (int s = 0; s < P.n_seg; ++s) tb |= (uint64_t)((t >> P.seg[s].src) & ((1u << P.seg[s].len) - 1u)) << P.seg[s].dst;
        for (int i = 0; i < 5; ++i) c[i] = 0;
        c[P.base_dim] = (int)(tb >> P.base_shift);
        return tb;
    };
    auto issue_load = [&](uint32_t k) {  // k-th tile of this CTA
        const uint32_t t = blockIdx.x + k * gridDim.x;
        if (t >= P.ntiles) return;
        int c[5];
        coords(t, c);
        mbar_expect_tx(&full_bar[k & 1], TILE_BYTES);
        tma_load5(landing, &tmap, &full_bar[k & 1], c[0], c[1], c[2], c[3], c[4]);
    };
    if (threadIdx.x == 0) issue_load(0);
    // phase-0 distribution: registers = tile bits 8..12, lanes = tile bits 0..3 (+4..7)
    uint32_t ld_off[1];
    ld_off[0] = tile_off((uint32_t)tid, P.swz);  // + j * 256 amps = j * 16 rows * 128 B (row & 7 unchanged)
    float2 a[NR];
    uint32_t kpar = 0;  // parity of this group's full barrier
    for (uint32_t k = g; ; k += 2) {
        const uint32_t t = blockIdx.x + k * gridDim.x;
        if (t >= P.ntiles) break;
        mbar_wait(&full_bar[g], kpar);
        kpar ^= 1u;
#pragma unroll
        for (int j = 0; j < NR; ++j) a[j] = *reinterpret_cast<const float2 *>(landing + ld_off[0] + j * 2048);
        group_bar(g);  // everybody has read the landing buffer
        if (tid == 0) issue_load(k + 1);
        // ---- stand-in for the ops of the pass
#pragma unroll
        for (int it = 0; it < NFMA / 2; ++it) {
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                const float2 x = a[j];
                a[j] = __ffma2_rn(make_float2(x.y, x.x), make_float2(-wi, wi), __fmul2_rn(x, make_float2(wr, wr)));
            }
        }
        // ---- the previous store of this group has to be done reading the buffer
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        group_bar(g);
        if (P.rlow) {
            // registers = tile bits 0..4: thread bits = tile bits 5..12
#pragma unroll
            for (int j = 0; j < NR; ++j) *reinterpret_cast<float2 *>(gbuf + tile_off(((uint32_t)tid << 5) | j, P.swz)) = a[j];
        } else {
#pragma unroll
            for (int j = 0; j < NR; ++j) *reinterpret_cast<float2 *>(gbuf + ld_off[0] + j * 2048) = a[j];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        group_bar(g);
        if (tid == 0) {
            int c[5];
            coords(t, c);
            tma_store5(&tmap, gbuf, c[0], c[1], c[2], c[3], c[4]);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void fill_kernel(float2 *s, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        s[i] = make_float2((float)(i & 0xffff) * (1.0f / 65536.0f), (float)((i >> 16) & 0xffff) * (1.0f / 65536.0f));
}
__global__ void check_kernel(const float2 *s, uint64_t n, float sr, float si, unsigned long long *bad) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float x = (float)(i & 0xffff) * (1.0f / 65536.0f), y = (float)((i >> 16) & 0xffff) * (1.0f / 65536.0f);
        const float er = x * sr - y * si, ei = x * si + y * sr;
        if (fabsf(s[i].x - er) > 1e-4f || fabsf(s[i].y - ei) > 1e-4f) atomicAdd(bad, 1ull);
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int NFMA>
float run(const CUtensorMap &tm, const Params &P, int grid, int reps, float wr, float wi) {
    auto k = probe_kernel<NFMA>;
    const size_t smem = 3 * TILE_BYTES + 1024;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k<<<grid, 2 * GROUP, smem>>>(tm, P, wr, wi);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) k<<<grid, 2 * GROUP, smem>>>(tm, P, wr, wi);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main(int argc, char **argv) {
    int n = argc > 1 ? atoi(argv[1]) : 28;
    int nfma = argc > 2 ? atoi(argv[2]) : 0;
    int swz = argc > 3 ? atoi(argv[3]) : 1;
    int overlap = argc > 4 ? atoi(argv[4]) : 1;
    int rlow = argc > 5 ? atoi(argv[5]) : 0;
    std::vector<int> tile = {0, 1, 2, 3};
    for (int i = 6; i < argc; ++i) tile.push_back(atoi(argv[i]));
    for (int p = 4; (int)tile.size() < T; ++p) {
        bool in = false;
        for (int x : tile) in |= x == p;
        if (!in) tile.push_back(p);
    }
    std::vector<int> tp(tile);
    for (size_t i = 0; i < tp.size(); ++i)
        for (size_t j = i + 1; j < tp.size(); ++j)
            if (tp[j] < tp[i]) std::swap(tp[i], tp[j]);
    printf("n=%d fma/amp=%d swizzle=%d overlap=%d rlow=%d tile:", n, nfma, swz, overlap, rlow);
    for (int x : tp) printf(" %d", x);
    printf("\n");
    const uint64_t N = 1ull << n;
    float2 *state;
    CK(cudaMalloc(&state, N * 8));
    fill_kernel<<<148 * 8, 256>>>(state, N);
    CK(cudaDeviceSynchronize());

    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }

    // ---- tensor map
    Params P;
    memset(&P, 0, sizeof(P));
    cuuint64_t gdim[5] = {1, 1, 1, 1, 1}, gstr[4] = {0, 0, 0, 0};
    cuuint32_t box[5] = {1, 1, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
    int nd = 0;
    gdim[0] = 32; box[0] = 32;  // 16 amplitudes = 32 floats = 128 bytes
    nd = 1;
    bool is_tile[64] = {false};
    for (int x : tp) is_tile[x] = true;
    if (overlap) {
        // base dimension: stride 128 B, box 1, coordinate = all non-tile bits >> 4
        gdim[nd] = 1ull << (n - 4); gstr[nd - 1] = 128; box[nd] = 1;
        P.base_dim = nd; P.base_shift = 4; ++nd;
        int p = 4;
        while (p < n) {
            if (!is_tile[p]) { ++p; continue; }
            int q = p;
            while (q < n && is_tile[q] && q - p < 8) ++q;
            if (nd >= 5) { printf("too many tile runs for a rank-5 tensor map\n"); return 2; }
            gdim[nd] = 1ull << (n - p); gstr[nd - 1] = 8ull << p; box[nd] = 1u << (q - p); ++nd;
            p = q;
        }
    } else {
        // disjoint bit ranges [4, n): alternating tile / non-tile; the non-tile bits need ONE coordinate, so
        // only tiles whose non-tile bits form one run above the tile bits are expressible... use the
        // last range as base
        int p = 4;
        P.base_dim = -1;
        while (p < n) {
            int q = p;
            const bool tl = is_tile[p];
            while (q < n && is_tile[q] == tl && (!tl || q - p < 8)) ++q;
            if (nd >= 5) { printf("too many ranges for a rank-5 tensor map\n"); return 2; }
            gdim[nd] = 1ull << (q - p); gstr[nd - 1] = 8ull << p; box[nd] = tl ? 1u << (q - p) : 1u;
            if (!tl) {
                if (P.base_dim >= 0) { printf("non-overlapping form needs a single non-tile run\n"); return 2; }
                P.base_dim = nd; P.base_shift = p;
            }
            ++nd;
            p = q;
        }
        if (P.base_dim < 0) P.base_dim = 0;
    }
    P.ndim = nd;
    while (nd < 5) { gdim[nd] = 1; gstr[nd - 1] = gstr[nd - 2] * gdim[nd - 1]; box[nd] = 1; ++nd; }
    CUtensorMap tm;
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, state, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d; dims:", (int)r);
    for (int i = 0; i < 5; ++i) printf(" (size %llu box %u stride %llu)", (unsigned long long)gdim[i], box[i], i ? (unsigned long long)gstr[i - 1] : 4ull);
    printf("\n");
    if (r != CUDA_SUCCESS) return 3;
    // segments: tile id bits -> non-tile positions
    int src = 0;
    for (int p = 4; p < n;) {
        if (is_tile[p]) { ++p; continue; }
        int q = p;
        while (q < n && !is_tile[q]) ++q;
        P.seg[P.n_seg++] = Seg{(uint8_t)src, (uint8_t)p, (uint8_t)(q - p), 0};
        src += q - p;
        p = q;
    }
    P.ntiles = 1u << (n - T);
    P.swz = swz;
    P.rlow = rlow;
    int sm = 0;
    CK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0));
    const float th = 0.37f;
    const float wr = cosf(th), wi = sinf(th);
    float ms = 0;
    const int reps = 5;
    int applied = 0;
    switch (nfma) {
        case 0: ms = run<0>(tm, P, sm, reps, wr, wi); applied = 0; break;
        case 8: ms = run<8>(tm, P, sm, reps, wr, wi); applied = 4; break;
        case 16: ms = run<16>(tm, P, sm, reps, wr, wi); applied = 8; break;
        case 24: ms = run<24>(tm, P, sm, reps, wr, wi); applied = 12; break;
        case 32: ms = run<32>(tm, P, sm, reps, wr, wi); applied = 16; break;
        case 48: ms = run<48>(tm, P, sm, reps, wr, wi); applied = 24; break;
        default: printf("fma/amp must be one of 0 8 16 24 32 48\n"); return 4;
    }
    const double gb = 2.0 * N * 8 / 1e9;
    printf("RESULT n=%d fma=%d swz=%d rlow=%d: %.3f ms per sweep, %.1f GB/s (%.3f of 6541.5)\n", n, nfma, swz, rlow, ms, gb / (ms * 1e-3),
           gb / (ms * 1e-3) / 6541.5);
    // check: the data went through (reps + 1) sweeps of `applied` multiplications by w
    const int total = applied * (reps + 1);
    float sr = 1, si = 0;
    for (int i = 0; i < total; ++i) {
        const float nr = sr * wr - si * wi, ni = sr * wi + si * wr;
        sr = nr; si = ni;
    }
    unsigned long long *bad, hbad = 0;
    CK(cudaMalloc(&bad, 8));
    CK(cudaMemset(bad, 0, 8));
    check_kernel<<<148 * 8, 256>>>(state, N, sr, si, bad);
    CK(cudaMemcpy(&hbad, bad, 8, cudaMemcpyDeviceToHost));
    printf("CHECK %s (%llu bad of %llu)\n", hbad ? "FAILED" : "ok", hbad, (unsigned long long)N);
    return hbad ? 5 : 0;
}
