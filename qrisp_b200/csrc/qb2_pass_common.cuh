// qb2_pass_common.cuh -- arithmetic and helpers shared by the pass kernels (see qb2_pass_kernel.cuh,
// qb2_pass_tma.cuh).  complex64 arithmetic runs on Blackwell's packed FP32 pipe: an amplitude is one
// 64-bit register pair (re, im) from the load to the store, and every operation is an FADD2 / FMUL2 /
// FFMA2 on the pair.  A complex multiplication is two instructions (the operand swizzles .LO_HI /
// scalar splat / per-half negation are free), a real butterfly is two.
#pragma once
#include <cuda.h>  // CUfunction / CUtensorMap types; entry points come from cudaGetDriverEntryPoint

#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "qb2_internal.cuh"

#ifndef QB2_MINB4
#define QB2_MINB4 4
#endif

namespace qb2 {

namespace {

constexpr int ROOT_BITS = 8;
constexpr int NROOT = 1 << ROOT_BITS;

// ---- complex arithmetic on one amplitude
template <typename real>
struct Cx;
template <>
struct Cx<float> {
    using amp = float2;
    static __device__ __forceinline__ amp add(amp a, amp b) { return __fadd2_rn(a, b); }
    static __device__ __forceinline__ amp sub(amp a, amp b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
    static __device__ __forceinline__ amp rmul(amp a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
    static __device__ __forceinline__ amp rfma(amp a, float s, amp c) { return __ffma2_rn(a, make_float2(s, s), c); }
    // a * (cr + i ci): FMUL2 a, cr.splat ; FFMA2 a.LO_HI, (-ci, ci), t
    static __device__ __forceinline__ amp cmul(amp a, float cr, float ci) {
        return __ffma2_rn(make_float2(a.y, a.x), make_float2(-ci, ci), __fmul2_rn(a, make_float2(cr, cr)));
    }
    static __device__ __forceinline__ amp cfma(amp a, float cr, float ci, amp acc) {
        return __ffma2_rn(make_float2(a.y, a.x), make_float2(-ci, ci), __ffma2_rn(a, make_float2(cr, cr), acc));
    }
};
template <>
struct Cx<double> {
    using amp = double2;
    static __device__ __forceinline__ amp add(amp a, amp b) { return make_double2(a.x + b.x, a.y + b.y); }
    static __device__ __forceinline__ amp sub(amp a, amp b) { return make_double2(a.x - b.x, a.y - b.y); }
    static __device__ __forceinline__ amp rmul(amp a, double s) { return make_double2(a.x * s, a.y * s); }
    static __device__ __forceinline__ amp rfma(amp a, double s, amp c) { return make_double2(fma(a.x, s, c.x), fma(a.y, s, c.y)); }
    static __device__ __forceinline__ amp cmul(amp a, double cr, double ci) {
        return make_double2(fma(a.x, cr, -a.y * ci), fma(a.x, ci, a.y * cr));
    }
    static __device__ __forceinline__ amp cfma(amp a, double cr, double ci, amp acc) {
        return make_double2(fma(a.x, cr, fma(-a.y, ci, acc.x)), fma(a.x, ci, fma(a.y, cr, acc.y)));
    }
};

template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// run f(integral_constant<int, v>) for a runtime v in [0, N): a switch, so that the compiler sees
// mutually exclusive arms (a chain of ifs gets if-converted into register copies)
template <int N, typename F>
__device__ __forceinline__ void dispatch(const int v, F &&f) {
    static_assert(N <= 8, "dispatch: at most 8 arms");
    switch (v) {
        case 0: if constexpr (N > 0) f(std::integral_constant<int, 0>{}); break;
        case 1: if constexpr (N > 1) f(std::integral_constant<int, 1>{}); break;
        case 2: if constexpr (N > 2) f(std::integral_constant<int, 2>{}); break;
        case 3: if constexpr (N > 3) f(std::integral_constant<int, 3>{}); break;
        case 4: if constexpr (N > 4) f(std::integral_constant<int, 4>{}); break;
        case 5: if constexpr (N > 5) f(std::integral_constant<int, 5>{}); break;
        case 6: if constexpr (N > 6) f(std::integral_constant<int, 6>{}); break;
        case 7: if constexpr (N > 7) f(std::integral_constant<int, 7>{}); break;
        default: break;
    }
}

// compile-time bookkeeping for a gate on register bits P0 < P1 < ... (K of them)
template <int RB, int K, int P0, int P1, int P2, int P3>
struct BitMap {
    __host__ __device__ static constexpr int pos(int i) { return i == 0 ? P0 : i == 1 ? P1 : i == 2 ? P2 : P3; }
    __host__ __device__ static constexpr int tmask() {
        int m = 0;
        for (int i = 0; i < K; ++i) m |= 1 << pos(i);
        return m;
    }
    // matrix index t -> register index bits (matrix bit i <-> register bit pos(i))
    __host__ __device__ static constexpr int dep(int t) {
        int v = 0;
        for (int i = 0; i < K; ++i) v |= ((t >> i) & 1) << pos(i);
        return v;
    }
    // group number g -> register index with the target bits clear
    __host__ __device__ static constexpr int base(int g) {
        int b = 0, gi = 0;
        for (int p = 0; p < RB; ++p)
            if (!((tmask() >> p) & 1)) {
                b |= ((g >> gi) & 1) << p;
                ++gi;
            }
        return b;
    }
};

template <typename real, int RB, int K, int P0, int P1 = -1, int P2 = -1, int P3 = -1>
__device__ __forceinline__ void apply_mat(typename Cx<real>::amp (&a)[1 << RB], const real *__restrict__ m,
                                          const uint32_t jmask) {
    using BM = BitMap<RB, K, P0, P1, P2, P3>;
    using C = Cx<real>;
    using amp = typename C::amp;
    constexpr int D = 1 << K;
    constexpr int NG = (1 << RB) / D;
    static_for<0, NG>([&](auto G) {
        constexpr int base = BM::base(decltype(G)::value);
        if ((jmask >> base) & 1u) {
            amp x[D];
            static_for<0, D>([&](auto Cc) { x[decltype(Cc)::value] = a[base | BM::dep(decltype(Cc)::value)]; });
            static_for<0, D>([&](auto Rw) {
                constexpr int row = decltype(Rw)::value;
                amp acc = C::cmul(x[0], m[2 * (row * D)], m[2 * (row * D) + 1]);
                static_for<1, D>([&](auto Cc) {
                    constexpr int col = decltype(Cc)::value;
                    acc = C::cfma(x[col], m[2 * (row * D + col)], m[2 * (row * D + col) + 1], acc);
                });
                a[base | BM::dep(row)] = acc;
            });
        }
    });
}

// exp(2 pi i * turns / 2**32): 8 high bits from a table of exact roots, the centred low
// 24 bits through a short Taylor series (|theta| <= pi/256: the dropped terms are < 1e-9).
__device__ __forceinline__ float2 unit_phase(const uint32_t turns, const float2 *__restrict__ roots) {
    const uint32_t s = turns + (1u << 23);
    const uint32_t k = s >> 24;
    const int low = (int)(s & 0xFFFFFFu) - (1 << 23);
    const float th = (float)low * 1.4629180792671596e-9f;  // 2 pi / 2**32
    const float t2 = th * th;
    const float c = fmaf(-0.5f, t2, 1.0f);
    const float sn = fmaf(th * t2, -0.16666667f, th);
    const float2 w = roots[k];
    return make_float2(fmaf(w.x, c, -w.y * sn), fmaf(w.x, sn, w.y * c));
}

__device__ __forceinline__ double2 unit_phase(const uint64_t turns, const double2 *) {
    double s, c;
    sincospi((double)(int64_t)turns * 0x1p-63, &s, &c);
    return make_double2(c, s);
}

__device__ __forceinline__ uint32_t tab_index(const qb2_tab &t, const uint64_t x) {
    const int np = t.p[0].rsv;  // number of pieces
    uint32_t idx = 0;
    for (int p = 0; p < np; ++p) {
        const qb2_piece pc = t.p[p];
        idx |= ((uint32_t)(x >> pc.shift) & ((1u << pc.len) - 1u)) << pc.out;
    }
    return idx;
}

// A table whose entries are a multiple of the index needs no memory: the phase of a ladder of
// controlled phases with geometric angles (QFT, QPE, Fourier arithmetic) is
// mult * (bits of the physical index) mod 1 turn.
__device__ __forceinline__ bool tab_is_mul(const qb2_tab &t) { return t.p[1].rsv != 0; }
template <typename frac>
__device__ __forceinline__ frac tab_mul(const qb2_tab &t, const uint64_t x) {
    const qb2_mul &m = reinterpret_cast<const qb2_mul &>(t);
    uint32_t v = (uint32_t)(x >> m.shift) & m.mask;
    if (m.brev) v = __brev(v) >> m.rsh;  // angles that halve with increasing position
    return (frac)m.mult * (frac)v;
}



__device__ __forceinline__ float2 ld_na(const float2 *p) {
    float2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ld_na(const double2 *p) {
    double2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// zero the registers, then drop in the entries of tile `tl`; returns true when the tile has none.
// `sp_tile` is the shared-memory copy of SP.tile (binary search: the same for every thread of the CTA)
template <typename real, int RB>
__device__ __forceinline__ bool sparse_synth(typename Cx<real>::amp (&a)[1 << RB], const uint32_t *sp_tile, const SparseInput &SP,
                                             const uint32_t tl, const uint32_t tid) {
    using amp = typename Cx<real>::amp;
    static_for<0, (1 << RB)>([&](auto J) {
        a[decltype(J)::value].x = 0;
        a[decltype(J)::value].y = 0;
    });
    uint32_t lo = 0, hi = SP.count;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sp_tile[mid] < tl)
            lo = mid + 1;
        else
            hi = mid;
    }
    bool any = false;
    for (uint32_t e = lo; e < SP.count && sp_tile[e] == tl; ++e) {
        any = true;
        const uint32_t loc = __ldg(SP.loc + e);
        if ((loc >> RB) == tid) {
            const amp v = reinterpret_cast<const amp *>(SP.amp)[e];
            const uint32_t j = loc & ((1u << RB) - 1u);
            static_for<0, (1 << RB)>([&](auto J) {
                if (j == (uint32_t)decltype(J)::value) a[decltype(J)::value] = v;
            });
        }
    }
    return !any;
}

// ---- shared-memory tile access by XOR slot maps.  The slot of register j is base ^ x[j], x[j] = XOR of
// the columns of j's set bits.  Walking the registers in Gray-code order turns that into ONE XOR per
// access (the column of the bit that flips), with the five columns in uniform registers: no per-register
// table look-up, whatever the map.  `cols16` points at the register entries of a slot map (uint16[2**r]
// slots in amplitudes): entry 1 << i is the column of register bit i.
//
// Addresses are byte offsets from the start of the kernel's dynamic shared memory.  WIN is the address
// of that start in the CTA's shared window (the driver reserves the first 1 KB; the launcher checks
// cudaDevAttrReservedSharedMemoryPerBlock and the kernel checks its own window): it rides as the
// immediate of every LDS / STS, so an access costs no address add.
template <int WIN>
__device__ __forceinline__ void sts_amp(const uint32_t off, const float2 v) {
    asm volatile("st.shared.v2.f32 [%0+%3], {%1, %2};" ::"r"(off), "f"(v.x), "f"(v.y), "n"(WIN) : "memory");
}
template <int WIN>
__device__ __forceinline__ float2 lds_amp(const uint32_t off) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(off), "n"(WIN) : "memory");
    return v;
}
template <int G>
struct Gray {
    static constexpr int value = G ^ (G >> 1);
    // bit that differs between Gray<G> and Gray<G+1>: the lowest zero bit of G
    static constexpr int flip = (G & 1) == 0 ? 0 : ((G & 2) == 0 ? 1 : ((G & 4) == 0 ? 2 : ((G & 8) == 0 ? 3 : ((G & 16) == 0 ? 4 : 5))));
};
// the columns of the register bits of one side of an exchange, as byte deltas (entry 1 << i of the slot
// map's register part): read from the descriptors by the interpreter, constants in a specialised kernel
template <int RB>
struct GrayCols {
    uint32_t d[RB];
};
template <int RB>
__device__ __forceinline__ GrayCols<RB> gray_cols(const uint16_t *cols16) {
    GrayCols<RB> g;
#pragma unroll
    for (int i = 0; i < RB; ++i) g.d[i] = (uint32_t)cols16[1 << i] * (uint32_t)sizeof(float2);
    return g;
}
template <int RB, int WIN>
__device__ __forceinline__ void smem_store_gray(uint32_t off, const GrayCols<RB> &gc, const float2 (&a)[1 << RB]) {
    const uint32_t (&d)[RB] = gc.d;
    static_for<0, (1 << RB)>([&](auto Gi) {
        constexpr int g = decltype(Gi)::value;
        sts_amp<WIN>(off, a[Gray<g>::value]);
        if constexpr (g + 1 < (1 << RB)) off ^= d[Gray<g>::flip];
    });
}
template <int RB, int WIN>
__device__ __forceinline__ void smem_load_gray(uint32_t off, const GrayCols<RB> &gc, float2 (&a)[1 << RB]) {
    const uint32_t (&d)[RB] = gc.d;
    static_for<0, (1 << RB)>([&](auto Gi) {
        constexpr int g = decltype(Gi)::value;
        a[Gray<g>::value] = lds_amp<WIN>(off);
        if constexpr (g + 1 < (1 << RB)) off ^= d[Gray<g>::flip];
    });
}
// register j at base + (j << K) amplitudes: immediate offsets
template <int RB, int WIN, int K>
__device__ __forceinline__ void smem_store_reg(const uint32_t off, const float2 (&a)[1 << RB]) {
    static_for<0, (1 << RB)>([&](auto J) {
        constexpr int j = decltype(J)::value;
        sts_amp<WIN + j * (int)(sizeof(float2) << K)>(off, a[j]);
    });
}
template <int RB, int WIN, int K>
__device__ __forceinline__ void smem_load_reg(const uint32_t off, float2 (&a)[1 << RB]) {
    static_for<0, (1 << RB)>([&](auto J) {
        constexpr int j = decltype(J)::value;
        a[j] = lds_amp<WIN + j * (int)(sizeof(float2) << K)>(off);
    });
}
// one side of an exchange: `regular` = registers at base + (j << k) (k from the phase's xflags), else any map
template <int RB, int WIN>
__device__ __forceinline__ void smem_store_any(const uint32_t off, const bool regular, const uint32_t k, const GrayCols<RB> &cols16,
                                               const float2 (&a)[1 << RB]) {
    if (regular && k == 8u)
        smem_store_reg<RB, WIN, 8>(off, a);
    else if (regular && k == 4u)
        smem_store_reg<RB, WIN, 4>(off, a);
    else if (regular && k == 7u)
        smem_store_reg<RB, WIN, 7>(off, a);
    else
        smem_store_gray<RB, WIN>(off, cols16, a);
}
template <int RB, int WIN>
__device__ __forceinline__ void smem_load_any(const uint32_t off, const bool regular, const uint32_t k, const GrayCols<RB> &cols16,
                                              float2 (&a)[1 << RB]) {
    if (regular && k == 8u)
        smem_load_reg<RB, WIN, 8>(off, a);
    else if (regular && k == 4u)
        smem_load_reg<RB, WIN, 4>(off, a);
    else if (regular && k == 7u)
        smem_load_reg<RB, WIN, 7>(off, a);
    else
        smem_load_gray<RB, WIN>(off, cols16, a);
}

// ---- development aid (make trace): time stamps of CTA 0 / thread 0 along its first tiles
#ifdef QB2_TRACE
#define QB2_TRACE_TILES 24
#define QB2_TRACE_STAMPS 64
__device__ unsigned long long qb2_trace_buf[QB2_TRACE_TILES * QB2_TRACE_STAMPS];
#define TR(k)                                                                                   \
    do {                                                                                        \
        if (blockIdx.x == 0 && threadIdx.x == 0 && iter < QB2_TRACE_TILES && (k) < QB2_TRACE_STAMPS) \
            qb2_trace_buf[iter * QB2_TRACE_STAMPS + (k)] = clock64();                           \
    } while (0)
#else
#define TR(k) \
    do {      \
    } while (0)
#endif

// launch of a specialised kernel (a CUfunction from qb2_jit_load) through the driver entry points
inline int launch_jit_kernel(const void *kernel, unsigned grid, unsigned block, size_t smem, cudaStream_t s, void **args) {
    typedef CUresult (*attr_fn)(CUfunction, CUfunction_attribute, int);
    typedef CUresult (*launch_fn)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **, void **);
    static attr_fn set_attr = nullptr;
    static launch_fn launch = nullptr;
    if (!set_attr || !launch) {
        cudaDriverEntryPointQueryResult q;
        void *f1 = nullptr, *f2 = nullptr;
        if (cudaGetDriverEntryPoint("cuFuncSetAttribute", &f1, cudaEnableDefault, &q) != cudaSuccess || !f1 ||
            cudaGetDriverEntryPoint("cuLaunchKernel", &f2, cudaEnableDefault, &q) != cudaSuccess || !f2)
            return set_error(QB2_ECUDA, "qb2_run_pass_jit: the driver has no cuLaunchKernel");
        set_attr = reinterpret_cast<attr_fn>(f1);
        launch = reinterpret_cast<launch_fn>(f2);
    }
    CUfunction fn = reinterpret_cast<CUfunction>(const_cast<void *>(kernel));
    CUresult r = set_attr(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, smem > 48 * 1024 ? (int)smem : 48 * 1024);
    if (r != CUDA_SUCCESS) return set_error(QB2_ECUDA, "qb2_run_pass_jit: cuFuncSetAttribute failed with %d", (int)r);
    r = launch(fn, grid, 1, 1, block, 1, 1, (unsigned)smem, (CUstream)s, args, nullptr);
    if (r != CUDA_SUCCESS) return set_error(QB2_ECUDA, "qb2_run_pass_jit: cuLaunchKernel failed with %d", (int)r);
    count_launch();
    return QB2_OK;
}

struct PassParams {  // the small section of a pass, passed by value: lives in the constant bank
    uint4 q[QB2_MAX_SMALL_BYTES / 16];
};

// pair g of a gate on register bit B: j0 has bit B clear, j1 = j0 | 1 << B
template <int B>
__host__ __device__ constexpr int pair_j0(int g) {
    return ((g >> B) << (B + 1)) | (g & ((1 << B) - 1));
}

// ---- 2x2 gates on register bit B.  ALLJ: every pair takes part (no register-bit controls).
template <typename real, int RB, int B, bool ALLJ>
__device__ __forceinline__ void mat1_general(typename Cx<real>::amp (&a)[1 << RB], const real (&m)[8], const uint32_t jmask) {
    using C = Cx<real>;
    static_for<0, (1 << RB) / 2>([&](auto G) {
        constexpr int j0 = pair_j0<B>(decltype(G)::value), j1 = j0 | (1 << B);
        if (ALLJ || ((jmask >> j0) & 1u)) {
            const auto x0 = a[j0], x1 = a[j1];
            a[j0] = C::cfma(x1, m[2], m[3], C::cmul(x0, m[0], m[1]));
            a[j1] = C::cfma(x1, m[6], m[7], C::cmul(x0, m[4], m[5]));
        }
    });
}

template <typename real, int RB, int B, bool ALLJ>
__device__ __forceinline__ void mat1_real(typename Cx<real>::amp (&a)[1 << RB], const real (&m)[8], const uint32_t jmask) {
    using C = Cx<real>;
    static_for<0, (1 << RB) / 2>([&](auto G) {
        constexpr int j0 = pair_j0<B>(decltype(G)::value), j1 = j0 | (1 << B);
        if (ALLJ || ((jmask >> j0) & 1u)) {
            const auto x0 = a[j0], x1 = a[j1];
            a[j0] = C::rfma(x1, m[2], C::rmul(x0, m[0]));
            a[j1] = C::rfma(x1, m[6], C::rmul(x0, m[4]));
        }
    });
}

// anti-diagonal 2x2: new0 = p * x1, new1 = q * x0
template <typename real, int RB, int B, bool ALLJ>
__device__ __forceinline__ void mat1_anti(typename Cx<real>::amp (&a)[1 << RB], const real (&m)[8], const uint32_t jmask) {
    using C = Cx<real>;
    static_for<0, (1 << RB) / 2>([&](auto G) {
        constexpr int j0 = pair_j0<B>(decltype(G)::value), j1 = j0 | (1 << B);
        if (ALLJ || ((jmask >> j0) & 1u)) {
            const auto x0 = a[j0], x1 = a[j1];
            a[j0] = C::cmul(x1, m[2], m[3]);
            a[j1] = C::cmul(x0, m[4], m[5]);
        }
    });
}

// Register swaps are written as XOR swaps on the raw words: with plain copies the compiler
// renames registers inside every arm of the bit dispatch and then copies the WHOLE register
// array at the join (about a thousand MOVs plus spills for the 20 CX arms).
__device__ __forceinline__ void xor_swap(float2 &p, float2 &q, const uint32_t mask) {
    uint32_t px = __float_as_uint(p.x), py = __float_as_uint(p.y), qx = __float_as_uint(q.x), qy = __float_as_uint(q.y);
    const uint32_t tx = (px ^ qx) & mask, ty = (py ^ qy) & mask;
    px ^= tx, qx ^= tx, py ^= ty, qy ^= ty;
    p = make_float2(__uint_as_float(px), __uint_as_float(py));
    q = make_float2(__uint_as_float(qx), __uint_as_float(qy));
}
__device__ __forceinline__ void xor_swap(double2 &p, double2 &q, const uint32_t mask) {
    const long long m = (long long)(int)mask;  // sign-extend the all-ones / zero mask
    long long px = __double_as_longlong(p.x), py = __double_as_longlong(p.y), qx = __double_as_longlong(q.x),
              qy = __double_as_longlong(q.y);
    const long long tx = (px ^ qx) & m, ty = (py ^ qy) & m;
    px ^= tx, qx ^= tx, py ^= ty, qy ^= ty;
    p = make_double2(__longlong_as_double(px), __longlong_as_double(py));
    q = make_double2(__longlong_as_double(qx), __longlong_as_double(qy));
}

template <typename real, int RB, int B, bool ALLJ>
__device__ __forceinline__ void mat1_swap(typename Cx<real>::amp (&a)[1 << RB], const uint32_t jmask) {
    uint32_t ones = 0xffffffffu;
    asm volatile("" : "+r"(ones));  // opaque: a known mask would be folded back into copies
    static_for<0, (1 << RB) / 2>([&](auto G) {
        constexpr int j0 = pair_j0<B>(decltype(G)::value), j1 = j0 | (1 << B);
        // branch free: the mask is all ones for a pair that takes part
        const uint32_t mask = ALLJ ? ones : 0u - ((jmask >> j0) & 1u);
        xor_swap(a[j0], a[j1], mask);
    });
}

// X on register bit B controlled by register bit CB (= 1): a fixed permutation of registers
template <typename real, int RB, int B, int CB>
__device__ __forceinline__ void mat1_swap_ctrl(typename Cx<real>::amp (&a)[1 << RB]) {
    uint32_t ones = 0xffffffffu;
    asm volatile("" : "+r"(ones));
    static_for<0, (1 << RB) / 2>([&](auto G) {
        constexpr int j0 = pair_j0<B>(decltype(G)::value), j1 = j0 | (1 << B);
        if constexpr ((j0 >> CB) & 1) xor_swap(a[j0], a[j1], ones);
    });
}

// ---- butterfly on register bit B: real 2x2 block, then out1 *= c[pair].
// MODE 0: c = w (thread phase only)            MODE 1: c[g] = v[g] (host factors only)
// MODE 2: c[g] = w * v[g]                      MODE 3: as 2, v[g] depends on g's low B bits only
// HAD: the block is the unnormalised Hadamard (add / subtract)
template <typename real, int RB, int B, int MODE, bool HAD>
__device__ __forceinline__ void bfly(typename Cx<real>::amp (&a)[1 << RB], const real ma, const real mb, const real mc,
                                     const real md, const typename Cx<real>::amp w, const real *__restrict__ v) {
    using C = Cx<real>;
    using amp = typename C::amp;
    constexpr int NP = (1 << RB) / 2;
    // distinct factors: one live at a time (registers are the scarce resource here)
    constexpr int NC = MODE == 2 ? NP : (MODE == 3 ? (1 << B) : 1);
    static_for<0, NC>([&](auto I) {
        constexpr int i = decltype(I)::value;
        [[maybe_unused]] amp c = w;
        if constexpr (MODE == 2 || MODE == 3) {
            const amp f = reinterpret_cast<const amp *>(v)[i];  // one 64-bit constant load per factor
            c = C::cmul(w, f.x, f.y);
        }
        static_for<0, NP / NC>([&](auto Hh) {
            constexpr int g = decltype(Hh)::value * NC + i;
            constexpr int j0 = pair_j0<B>(g), j1 = j0 | (1 << B);
            const amp x0 = a[j0], x1 = a[j1];
            amp t;
            if constexpr (HAD) {  // s * [[1, 1], [1, -1]] with s left to the host (QB2_BFLY_HAD)
                a[j0] = C::add(x0, x1);
                t = C::sub(x0, x1);
            } else {
                a[j0] = C::rfma(x1, mb, C::rmul(x0, ma));
                t = C::rfma(x1, md, C::rmul(x0, mc));
            }
            if constexpr (MODE == 1)
            {
                const amp f = reinterpret_cast<const amp *>(v)[g];
                a[j1] = C::cmul(t, f.x, f.y);
            }
            else
                a[j1] = C::cmul(t, c.x, c.y);
        });
    });
}

}  // namespace
}  // namespace qb2
