// qb2_pass.cu -- the fused multi-gate pass kernel (see include/qrisp_b200.h).
//
// Replaces the reference's per-gate tensor contraction (tensor_factor.py:67-89 ->
// bi_arrays.py:1175-1299 `tensordot`, which reshapes, transposes and multiplies the whole
// state once per gate) by ONE streaming sweep per group of gates:
//
//   global --(coalesced 128 B lines)--> registers --ops--> [smem exchange --ops-->]* --> global
//
// Every thread owns 2**r amplitudes in registers.  Gates whose dense targets are register
// bits are pure register arithmetic (fully unrolled, so all register indices are compile
// time constants).  A smem "exchange" re-distributes the tile so that other tile positions
// become register bits; its slot map is an XOR-linear function of the thread id chosen by
// the host so that both the store and the load side are bank-conflict free.
//
// complex64 arithmetic runs on Blackwell's packed FP32 pipe: an amplitude is one 64-bit
// register pair (re, im) from the load to the store, and every operation is an FADD2 / FMUL2 /
// FFMA2 on the pair.  A complex multiplication is two instructions (the operand swizzles
// .LO_HI / scalar splat / per-half negation are free), a real butterfly is two.
//
// HBM-bound by construction: algorithmic bytes per launch = 2 * 2**n * sizeof(amplitude).
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
        amp c = w;
        if constexpr (MODE == 2 || MODE == 3) c = C::cmul(w, v[2 * i], v[2 * i + 1]);
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
                a[j1] = C::cmul(t, v[2 * g], v[2 * g + 1]);
            else
                a[j1] = C::cmul(t, c.x, c.y);
        });
    });
}

// FULL = false leaves out the rarely used op kinds (general per-amplitude DIAG, 4x4 / 8x8 /
// 16x16 blocks): the common passes then run a kernel a fraction of the size, which matters
// because the interpreter's working set of code competes for the instruction cache.
// Experiments that did not pay and are gone from the code: an L2 prefetch of
// the next tile (bulk or plain prefetch instructions); staggering the start of the CTAs that share
// an SM; register permutations written as plain copies (the compiler then copies the whole array at
// every join); folding leading X / CX gates of a phase into the load table / the exchange's read
// map (the irregular exchange path it forces costs what the gates cost); fetching the next tile into
// the idle exchange buffer with cp.async during the last phase (issuing 32 eight-byte async copies
// per thread costs more than the latency it hides: 25.9 vs 24.3 ms).
template <typename real, int RB, int MAXT, int MINB, bool FULL>
__global__ void __launch_bounds__(MAXT, MINB)
    pass_kernel(typename Traits<real>::vec *__restrict__ state, const typename Traits<real>::frac *__restrict__ tables,
                const uint64_t hi_base, const uint32_t ntiles, const uint32_t ctx_off, const uint32_t slot_off,
                const uint32_t tile_off, const __grid_constant__ PassParams P) {
    using vec = typename Traits<real>::vec;
    using frac = typename Traits<real>::frac;
    using C = Cx<real>;
    using amp = typename C::amp;
    static_assert(sizeof(amp) == sizeof(vec), "an amplitude is one vector register group");
    constexpr int NR = 1 << RB;
    extern __shared__ __align__(16) uint8_t smem[];

    const uint32_t tid = threadIdx.x;
    vec *const roots = reinterpret_cast<vec *>(smem);
    if constexpr (std::is_same<real, float>::value) {
        for (uint32_t i = tid; i < NROOT; i += blockDim.x) {
            float s, c;
            sincospif((float)i * (2.0f / NROOT), &s, &c);
            roots[i] = make_float2(c, s);
        }
        __syncthreads();
    }

    // every descriptor read below is a constant-bank load with a warp-uniform offset
    const uint8_t *const pb = reinterpret_cast<const uint8_t *>(&P);
    const qb2_pass_header *const H = reinterpret_cast<const qb2_pass_header *>(pb);
    const qb2_io *const IO = reinterpret_cast<const qb2_io *>(pb + QB2_IO_OFFSET);  // fixed offsets: immediate operands
    const qb2_phase *const phases = reinterpret_cast<const qb2_phase *>(pb + H->phases_off);
    const qb2_op *const ops = reinterpret_cast<const qb2_op *>(pb + H->ops_off);
    const real *const coefs = reinterpret_cast<const real *>(pb + H->coef_off);
    const uint16_t *const a16 = reinterpret_cast<const uint16_t *>(pb + H->u16_off);
    const uint64_t *const a64 = reinterpret_cast<const uint64_t *>(pb + H->u64_off);
    const qb2_tab *const tabs = reinterpret_cast<const qb2_tab *>(pb + H->tabdesc_off);
    const int TB = (int)H->T - RB;  // thread bits
    const uint32_t n_phases = H->n_phases;
    const int n_seg = H->n_seg;

    // ---- per-thread context of every phase, computed once per CTA: the thread's physical
    // offset and its shared-memory byte offsets on the store / load side of the exchange that
    // enters the phase.  Each thread reads back only its own entry.
    uint4 *const ctx = reinterpret_cast<uint4 *>(smem + ctx_off);
    const uint32_t nthreads = blockDim.x;
    for (uint32_t p = 0; p < n_phases; ++p) {
        const qb2_phase &ph = phases[p];
        uint64_t xt = 0;
        uint32_t wb = 0, rb = 0;
        for (int b = 0; b < TB; ++b)
            if ((tid >> b) & 1u) {
                xt |= a64[ph.pcol + b];
                if (p > 0) {
                    wb ^= a16[ph.xw + b];
                    rb ^= a16[ph.xr + b];
                }
            }
        ctx[p * nthreads + tid] =
            make_uint4((uint32_t)xt, (uint32_t)(xt >> 32), wb * (uint32_t)sizeof(vec), rb * (uint32_t)sizeof(vec));
    }
    {  // one more entry: the thread's physical offset at the store (qb2_io.store_col)
        uint64_t xt = 0;
        for (int b = 0; b < TB; ++b)
            if ((tid >> b) & 1u) xt |= IO->store_col[b];
        ctx[n_phases * nthreads + tid] = make_uint4((uint32_t)xt, (uint32_t)(xt >> 32), 0u, 0u);
    }
    // ---- phase slots (qb2_op.b1): a phase made of multiplier ladders is linear in the index bits,
    // so it is the sum of a thread part (here, once per CTA) and a tile part (once per tile, below)
    const uint32_t n_slots = H->n_slots;
    auto two_slots = [&](const qb2_op &op) { return op.kind == QB2_OP_DIAG_BIT && !(op.flags & 1u); };
    frac *const stile = reinterpret_cast<frac *>(smem + slot_off);                         // [2][QB2_MAX_SLOTS]
    uint16_t *const slot_op = reinterpret_cast<uint16_t *>(stile + 2 * QB2_MAX_SLOTS);     // [QB2_MAX_SLOTS]
    frac *const sthr = reinterpret_cast<frac *>(slot_op + QB2_MAX_SLOTS);                  // [n_slots][threads]
    // value of slot `which` of an op on index x: an op with thread-only ladders AND register-bit
    // ladders (DIAG_BIT with both phases live) owns two slots: the thread-only sum, then the sum of all
    auto ladder_sum = [&](const qb2_op &op, const uint32_t which, const uint64_t x) {
        const qb2_tab *tb = tabs + (op.kind == QB2_OP_BFLY ? (uint32_t)op.tab : (uint32_t)op.data);
        const int nthr = op.ntab_thr, nall = nthr + op.ntab_reg;
        const int n = (two_slots(op) && which == 0u) ? nthr : nall;
        frac r = 0;
        for (int i = 0; i < n; ++i) r += tab_mul<frac>(tb[i], x);
        return r;
    };
    auto has_slot = [&](const qb2_op &op) {
        return op.b1 != 0 && (op.kind == QB2_OP_BFLY || op.kind == QB2_OP_DIAG_THR || op.kind == QB2_OP_DIAG_BIT);
    };
    if (n_slots) {
        for (uint32_t p = 0; p < n_phases; ++p) {
            const uint4 c = ctx[p * nthreads + tid];
            const uint64_t x = hi_base | ((uint64_t)c.y << 32 | c.x);
            const uint32_t op_end = phases[p].first_op + phases[p].n_ops;
            for (uint32_t o = phases[p].first_op; o < op_end; ++o) {
                const qb2_op &op = ops[o];
                if (has_slot(op)) {
                    const uint32_t ns = two_slots(op) ? 2u : 1u;
                    for (uint32_t w = 0; w < ns; ++w) {
                        sthr[(op.b1 - 1u + w) * nthreads + tid] = ladder_sum(op, w, x);
                        if (tid == 0) slot_op[op.b1 - 1u + w] = (uint16_t)(o | (w << 15));
                    }
                }
            }
        }
        __syncthreads();
    }
    uint8_t *const tile_b = smem + tile_off;
    const uint8_t *const state_b = reinterpret_cast<const uint8_t *>(state);

    auto tile_base = [&](uint32_t t) {
        uint64_t tb = 0;
        for (int s = 0; s < n_seg; ++s) {
            const qb2_seg sg = H->seg[s];
            tb |= (uint64_t)((t >> sg.src) & ((1u << sg.len) - 1u)) << sg.dst;
        }
        return tb;
    };

    amp a[NR];

    // ---- software pipeline across tiles: the loads of the CTA's next tile are issued in the same
    // loop that stores the current one (register by register), so they are in flight while the
    // stores drain and the slots are set up; a tile's life is: ops, [exchange, ops]*, store + load.
    const uint4 c0 = ctx[tid];
    const uint64_t xthr0 = (uint64_t)c0.y << 32 | c0.x;  // this thread's offset under the first phase
    auto load_base = [&](const uint64_t tb) {
        const uint8_t *gb = state_b + (tb | xthr0) * sizeof(vec);
#ifndef QB2_NOPIN
        asm volatile("" : "+l"(gb));  // keep base + uniform offset: two adds per address
#endif
        return gb;
    };
    // QB2_FEATURE_ZERO_INPUT: the state is a fresh |0...0> that nobody has written: nothing is read
    const bool zero_in = (H->features & QB2_FEATURE_ZERO_INPUT) != 0u;
    auto synth_tile = [&](const uint64_t tb) {
        static_for<0, NR>([&](auto J) {
            a[decltype(J)::value].x = 0;
            a[decltype(J)::value].y = 0;
        });
        if ((hi_base | tb | xthr0) == 0) a[0].x = 1;  // register 0 of this thread is global index 0
    };
    uint32_t iter = 0;
    uint32_t t = blockIdx.x;
    uint64_t tbase = 0;
    if (t < ntiles) {
        tbase = tile_base(t);
        if (zero_in) {
            synth_tile(tbase);
        } else {
            const uint8_t *gb = load_base(tbase);
            static_for<0, NR>([&](auto J) {
                constexpr int j = decltype(J)::value;
                a[j] = ld_na(reinterpret_cast<const amp *>(gb + IO->load_off[j]));
            });
        }
    }
    for (; t < ntiles; t += gridDim.x, ++iter) {
        TR(0);
        uint64_t xloc = tbase | xthr0;  // local physical index of this thread's register 0
        TR(1);
        // ---- tile part of the slotted phases: one thread per slot, while the loads are in flight.
        // Two buffers: a warp may run one tile ahead of the others, never two (barrier below).
        const uint32_t spar = (iter & 1u) * QB2_MAX_SLOTS;
        if (n_slots) {
            const uint32_t nwarps = nthreads >= 32u ? nthreads >> 5 : 1u;  // tiny registers run less than a warp
            for (uint32_t sl = tid >> 5; sl < n_slots; sl += nwarps)
                if ((tid & 31u) == 0u) {
                    const uint32_t so = slot_op[sl];
                    stile[spar + sl] = ladder_sum(ops[so & 0x7fffu], so >> 15, tbase);
                }
            __syncthreads();
        }
        TR(2);
        for (uint32_t p = 0; p < n_phases; ++p) {
            const qb2_phase &ph = phases[p];
            if (p > 0) {
                TR(4 + 2 * p);
                // ---- exchange through shared memory
                const uint4 c = ctx[p * nthreads + tid];
                const uint32_t xf = ph.xflags;
                {
                    uint8_t *wp = tile_b + c.z;
                    if (xf & 1u) {  // register j sits at slot j << kw: immediate offsets
                        const uint32_t stride = (uint32_t)sizeof(vec) << (xf >> 8 & 0xffu);
                        static_for<0, NR>([&](auto J) {
                            constexpr int j = decltype(J)::value;
                            *reinterpret_cast<amp *>(wp + j * stride) = a[j];
                        });
                    } else {
                        const uint16_t *xw = a16 + ph.xw + TB;
                        static_for<0, NR>([&](auto J) {
                            constexpr int j = decltype(J)::value;
                            *reinterpret_cast<amp *>(tile_b + (c.z ^ (xw[j] * (uint32_t)sizeof(vec)))) = a[j];
                        });
                    }
                }
                __syncthreads();
                {
                    xloc = tbase | ((uint64_t)c.y << 32 | c.x);
                    const uint8_t *rp_ = tile_b + c.w;
                    if (xf & 2u) {
                        const uint32_t stride = (uint32_t)sizeof(vec) << (xf >> 16 & 0xffu);
                        static_for<0, NR>([&](auto J) {
                            constexpr int j = decltype(J)::value;
                            a[j] = *reinterpret_cast<const amp *>(rp_ + j * stride);
                        });
                    } else {
                        const uint16_t *xr = a16 + ph.xr + TB;
                        static_for<0, NR>([&](auto J) {
                            constexpr int j = decltype(J)::value;
                            a[j] = *reinterpret_cast<const amp *>(tile_b + (c.w ^ (xr[j] * (uint32_t)sizeof(vec))));
                        });
                    }
                }
                __syncthreads();
                TR(5 + 2 * p);
            }
            const uint64_t X = hi_base | xloc;
            // ---- ops of this phase, on registers
            const uint32_t op_end = ph.first_op + ph.n_ops;
            for (uint32_t o = ph.first_op; o < op_end; ++o) {
                const qb2_op &op = ops[o];
                // ---- thread phases of the op, as turn fractions, in front of the dispatch (the arms below
                // are register arithmetic): s0 over the descriptors that depend on thread/tile bits only,
                // s1 = s0 + those that also depend on register bit b0 (their b0=1 half; multiplier
                // ladders vanish for b0=0).  Slotted ops: two shared-memory reads.
                frac s0 = 0, s1 = 0;
                if (op.kind >= QB2_OP_DIAG_THR && op.kind != QB2_OP_DIAG_VEC) {  // DIAG_THR, DIAG_BIT, BFLY
                    if (op.b1 != 0) {
                        const uint32_t sl = op.b1 - 1u;
                        s1 = sthr[sl * nthreads + tid] + stile[spar + sl];
                        s0 = s1;  // one slot: thread-only ladders (DIAG_THR) or register-bit ladders only
                        if (two_slots(op)) s1 = sthr[(sl + 1u) * nthreads + tid] + stile[spar + sl + 1u];
                    } else {
                        const int nthr = op.ntab_thr, nreg = op.ntab_reg;
                        const qb2_tab *tb = tabs + (op.kind == QB2_OP_BFLY ? (uint32_t)op.tab : (uint32_t)op.data);
                        for (int i = 0; i < nthr; ++i)
                            s0 += tab_is_mul(tb[i]) ? tab_mul<frac>(tb[i], X) : __ldg(tables + tb[i].off + tab_index(tb[i], X));
                        s1 = s0;
                        tb += nthr;
                        for (int i = 0; i < nreg; ++i) {
                            if (tab_is_mul(tb[i])) {
                                s1 += tab_mul<frac>(tb[i], X);
                            } else {
                                const frac *e = tables + tb[i].off + tab_index(tb[i], X);
                                s0 += __ldg(e);
                                s1 += __ldg(e + tb[i].regoff);
                            }
                        }
                    }
                }
                // real 2x2 gate on register bit B, then a phase on its B=1 output that depends on the
                // other register bits (host-computed v[pair]) and on the thread (w1)
                auto do_bfly = [&](auto B, auto M, auto Hd) {
                    constexpr int bb = decltype(B)::value, mode = decltype(M)::value;
                    constexpr bool had = decltype(Hd)::value != 0;
                    if constexpr (bb < RB) {
                        const real *m = coefs + 2 * (size_t)op.data;
                        amp w1;
                        w1.x = 1, w1.y = 0;
                        if constexpr (mode != 1) w1 = unit_phase(s1, roots);
                        bfly<real, RB, bb, mode, had>(a, m[0], m[2], m[4], m[6], w1, m + 8);
                    }
                };
                // phase that depends on the thread and on ONE register bit: two phases per thread
                auto do_diag_bit = [&](auto B, auto S) {
                    constexpr int bb = decltype(B)::value;
                    constexpr bool skip0 = decltype(S)::value != 0;
                    if constexpr (bb < RB) {
                        const amp w1 = unit_phase(s1, roots);
                        amp w0 = w1;
                        if constexpr (!skip0) w0 = unit_phase(s0, roots);
                        static_for<0, NR>([&](auto J) {
                            constexpr int j = decltype(J)::value;
                            if constexpr ((j >> bb) & 1)
                                a[j] = C::cmul(a[j], w1.x, w1.y);
                            else if constexpr (!skip0)
                                a[j] = C::cmul(a[j], w0.x, w0.y);
                        });
                    }
                };
                // 2x2 gate on register bit B; thread-level controls may diverge, and divergent code
                // cannot use the uniform datapath: the coefficients are read before the test
                auto do_mat1 = [&](auto B, auto S, auto A) {
                    constexpr int bb = decltype(B)::value, shape = decltype(S)::value;
                    constexpr bool allj = decltype(A)::value != 0;
                    if constexpr (bb < RB) {
                        const real *m = coefs + 2 * (size_t)op.data;
                        real c8[8];
                        if constexpr (shape != QB2_MAT1_SWAP) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) c8[k] = m[k];
                        }
                        const uint32_t jmask = op.jmask;
                        if (((X ^ op.cval) & op.cmask) == 0) {
                            if constexpr (shape == QB2_MAT1_SWAP)
                                mat1_swap<real, RB, bb, allj>(a, jmask);
                            else if constexpr (shape == QB2_MAT1_REAL)
                                mat1_real<real, RB, bb, allj>(a, c8, jmask);
                            else if constexpr (shape == QB2_MAT1_ANTI)
                                mat1_anti<real, RB, bb, allj>(a, c8, jmask);
                            else
                                mat1_general<real, RB, bb, allj>(a, c8, jmask);
                        }
                    }
                };
                auto do_swapc = [&](auto B, auto CB) {
                    constexpr int bb = decltype(B)::value, cb = decltype(CB)::value;
                    if constexpr (bb < RB && cb < RB && bb != cb) {
                        if (((X ^ op.cval) & op.cmask) == 0) mat1_swap_ctrl<real, RB, bb, cb>(a);
                    }
                };
                // run of unnormalised butterflies on register bits BS, BS-1, ... (QB2_CODE_CHAIN): one
                // decode, straight-line code: the shared-memory reads and the root lookup of one
                // member overlap the arithmetic of the member before
                auto do_chain = [&](auto BS, auto N) {
                    constexpr int bs = decltype(BS)::value, n = decltype(N)::value;
                    if constexpr (bs < RB) {
                        static_for<0, n>([&](auto K) {
                            constexpr int k = decltype(K)::value;
                            const qb2_op &mo = ops[o + k];
                            amp w;
                            w.x = 1, w.y = 0;
                            if (mo.b1 != 0) {
                                const uint32_t sl = mo.b1 - 1u;
                                w = unit_phase((frac)(sthr[sl * nthreads + tid] + stile[spar + sl]), roots);
                            }
                            const real *m = coefs + 2 * (size_t)mo.data;
                            bfly<real, RB, bs - k, 3, true>(a, m[0], m[2], m[4], m[6], w, m + 8);
                        });
                        o += n - 1;
                    }
                };
                // run of DIAG_BIT ops, one per register bit in the mask (QB2_CODE_DBITS): one decode
                auto do_dbits = [&]() {
                    const uint32_t mask = op.tab;
                    uint32_t idx = 0;
                    static_for<0, RB>([&](auto Bi) {
                        constexpr int bb = RB - 1 - decltype(Bi)::value;
                        if ((mask >> bb) & 1u) {
                            const uint32_t sl = ops[o + idx].b1 - 1u;
                            ++idx;
                            const amp w = unit_phase((frac)(sthr[sl * nthreads + tid] + stile[spar + sl]), roots);
                            static_for<0, NR>([&](auto J) {
                                constexpr int j = decltype(J)::value;
                                if constexpr ((j >> bb) & 1) a[j] = C::cmul(a[j], w.x, w.y);
                            });
                        }
                    });
                    o += idx - 1u;
                };
                // CX chain inside the registers (QB2_CODE_CXL): bit TOP controls TOP-1, TOP-1 controls TOP-2...
                auto do_cxl = [&](auto TOP, auto N) {
                    constexpr int top = decltype(TOP)::value, n = decltype(N)::value;
                    if constexpr (top < RB) {
                        static_for<0, n>([&](auto K) {
                            constexpr int k = decltype(K)::value;
                            mat1_swap_ctrl<real, RB, top - k - 1, top - k>(a);
                        });
                        o += n - 1;
                    }
                };
                auto do_mat2 = [&](auto B0, auto B1) {
                    if constexpr (FULL && decltype(B1)::value < RB) {
                        if (((X ^ op.cval) & op.cmask) == 0)
                            apply_mat<real, RB, 2, decltype(B0)::value, decltype(B1)::value>(a, coefs + 2 * (size_t)op.data, op.jmask);
                    }
                };
#define IC(v) std::integral_constant<int, v>{}
#define BFLY_CASES(b)                                                      \
    case QB2_CODE_BFLY + 8 * b + 0: do_bfly(IC(b), IC(0), IC(0)); break;     \
    case QB2_CODE_BFLY + 8 * b + 1: do_bfly(IC(b), IC(0), IC(1)); break;     \
    case QB2_CODE_BFLY + 8 * b + 2: do_bfly(IC(b), IC(1), IC(0)); break;     \
    case QB2_CODE_BFLY + 8 * b + 3: do_bfly(IC(b), IC(1), IC(1)); break;     \
    case QB2_CODE_BFLY + 8 * b + 4: do_bfly(IC(b), IC(2), IC(0)); break;     \
    case QB2_CODE_BFLY + 8 * b + 5: do_bfly(IC(b), IC(2), IC(1)); break;     \
    case QB2_CODE_BFLY + 8 * b + 6: do_bfly(IC(b), IC(3), IC(0)); break;     \
    case QB2_CODE_BFLY + 8 * b + 7: do_bfly(IC(b), IC(3), IC(1)); break;
#define DBIT_CASES(b)                                                    \
    case QB2_CODE_DIAG_BIT + 2 * b + 0: do_diag_bit(IC(b), IC(0)); break; \
    case QB2_CODE_DIAG_BIT + 2 * b + 1: do_diag_bit(IC(b), IC(1)); break;
#define MAT1_CASES(b)                                                      \
    case QB2_CODE_MAT1 + 8 * b + 0: do_mat1(IC(b), IC(0), IC(0)); break;     \
    case QB2_CODE_MAT1 + 8 * b + 1: do_mat1(IC(b), IC(0), IC(1)); break;     \
    case QB2_CODE_MAT1 + 8 * b + 2: do_mat1(IC(b), IC(1), IC(0)); break;     \
    case QB2_CODE_MAT1 + 8 * b + 3: do_mat1(IC(b), IC(1), IC(1)); break;     \
    case QB2_CODE_MAT1 + 8 * b + 4: do_mat1(IC(b), IC(2), IC(0)); break;     \
    case QB2_CODE_MAT1 + 8 * b + 5: do_mat1(IC(b), IC(2), IC(1)); break;     \
    case QB2_CODE_MAT1 + 8 * b + 6: do_mat1(IC(b), IC(3), IC(0)); break;     \
    case QB2_CODE_MAT1 + 8 * b + 7: do_mat1(IC(b), IC(3), IC(1)); break;
#define SWAPC_CASES(b)                                                 \
    case QB2_CODE_SWAPC + 5 * b + 0: do_swapc(IC(b), IC(0)); break;      \
    case QB2_CODE_SWAPC + 5 * b + 1: do_swapc(IC(b), IC(1)); break;      \
    case QB2_CODE_SWAPC + 5 * b + 2: do_swapc(IC(b), IC(2)); break;      \
    case QB2_CODE_SWAPC + 5 * b + 3: do_swapc(IC(b), IC(3)); break;      \
    case QB2_CODE_SWAPC + 5 * b + 4: do_swapc(IC(b), IC(4)); break;
                switch (op.code) {
                    BFLY_CASES(0) BFLY_CASES(1) BFLY_CASES(2) BFLY_CASES(3) BFLY_CASES(4)
                    DBIT_CASES(0) DBIT_CASES(1) DBIT_CASES(2) DBIT_CASES(3) DBIT_CASES(4)
                    MAT1_CASES(0) MAT1_CASES(1) MAT1_CASES(2) MAT1_CASES(3) MAT1_CASES(4)
                    SWAPC_CASES(0) SWAPC_CASES(1) SWAPC_CASES(2) SWAPC_CASES(3) SWAPC_CASES(4)
                    case QB2_CODE_DIAG_THR: {
                        const amp w0 = unit_phase(s0, roots);
                        static_for<0, NR>([&](auto J) {
                            constexpr int j = decltype(J)::value;
                            a[j] = C::cmul(a[j], w0.x, w0.y);
                        });
                        break;
                    }
                    case QB2_CODE_DIAG_VEC: {
                        // phases that depend on register bits only: host-computed, uniform
                        const real *cv = coefs + 2 * (size_t)op.data;
                        static_for<0, NR>([&](auto J) {
                            constexpr int j = decltype(J)::value;
                            a[j] = C::cmul(a[j], cv[2 * j], cv[2 * j + 1]);
                        });
                        break;
                    }
                    case QB2_CODE_DIAG_GEN: {
                        if constexpr (FULL) {  // general: one table sum per amplitude
                            const int nthr = op.ntab_thr, nreg = op.ntab_reg;
                            const qb2_tab *tb = tabs + op.data;
                            frac sum = 0;  // thread tables again as a fraction (w0 above is already a phase)
                            for (int i = 0; i < nthr; ++i)
                                sum += tab_is_mul(tb[i]) ? tab_mul<frac>(tb[i], X) : __ldg(tables + tb[i].off + tab_index(tb[i], X));
                            tb += nthr;
                            const frac *t0 = tables, *t1 = tables, *t2 = tables, *t3 = tables;
                            const uint16_t *g0 = a16, *g1 = a16, *g2 = a16, *g3 = a16;
                            t0 = tables + tb[0].off + tab_index(tb[0], X);
                            g0 = a16 + tb[0].regoff;
                            if (nreg > 1) {
                                t1 = tables + tb[1].off + tab_index(tb[1], X);
                                g1 = a16 + tb[1].regoff;
                            }
                            if (nreg > 2) {
                                t2 = tables + tb[2].off + tab_index(tb[2], X);
                                g2 = a16 + tb[2].regoff;
                            }
                            if (nreg > 3) {
                                t3 = tables + tb[3].off + tab_index(tb[3], X);
                                g3 = a16 + tb[3].regoff;
                            }
                            static_for<0, NR>([&](auto J) {
                                constexpr int j = decltype(J)::value;
                                frac sj = sum + __ldg(t0 + g0[j]);
                                if (nreg > 1) sj += __ldg(t1 + g1[j]);
                                if (nreg > 2) sj += __ldg(t2 + g2[j]);
                                if (nreg > 3) sj += __ldg(t3 + g3[j]);
                                const amp w = unit_phase(sj, roots);
                                a[j] = C::cmul(a[j], w.x, w.y);
                            });
                        }
                        break;
                    }
                    case QB2_CODE_MAT2 + 0: do_mat2(IC(0), IC(1)); break;
                    case QB2_CODE_MAT2 + 1: do_mat2(IC(0), IC(2)); break;
                    case QB2_CODE_MAT2 + 2: do_mat2(IC(0), IC(3)); break;
                    case QB2_CODE_MAT2 + 3: do_mat2(IC(0), IC(4)); break;
                    case QB2_CODE_MAT2 + 4: do_mat2(IC(1), IC(2)); break;
                    case QB2_CODE_MAT2 + 5: do_mat2(IC(1), IC(3)); break;
                    case QB2_CODE_MAT2 + 6: do_mat2(IC(1), IC(4)); break;
                    case QB2_CODE_MAT2 + 7: do_mat2(IC(2), IC(3)); break;
                    case QB2_CODE_MAT2 + 8: do_mat2(IC(2), IC(4)); break;
                    case QB2_CODE_MAT2 + 9: do_mat2(IC(3), IC(4)); break;
                    case QB2_CODE_MAT3:
                        if constexpr (FULL && RB >= 3) {
                            if (((X ^ op.cval) & op.cmask) == 0) apply_mat<real, RB, 3, 0, 1, 2>(a, coefs + 2 * (size_t)op.data, op.jmask);
                        }
                        break;
                    case QB2_CODE_MAT4:
                        if constexpr (FULL && RB >= 4) {
                            if (((X ^ op.cval) & op.cmask) == 0) apply_mat<real, RB, 4, 0, 1, 2, 3>(a, coefs + 2 * (size_t)op.data, op.jmask);
                        }
                        break;
                    case QB2_CODE_DBITS: do_dbits(); break;
                    case QB2_CODE_CXL + 0: do_cxl(IC(2), IC(2)); break;
                    case QB2_CODE_CXL + 1: do_cxl(IC(3), IC(2)); break;
                    case QB2_CODE_CXL + 2: do_cxl(IC(3), IC(3)); break;
                    case QB2_CODE_CXL + 3: do_cxl(IC(4), IC(2)); break;
                    case QB2_CODE_CXL + 4: do_cxl(IC(4), IC(3)); break;
                    case QB2_CODE_CXL + 5: do_cxl(IC(4), IC(4)); break;
                    case QB2_CODE_CHAIN + 0: do_chain(IC(1), IC(2)); break;
                    case QB2_CODE_CHAIN + 1: do_chain(IC(2), IC(2)); break;
                    case QB2_CODE_CHAIN + 2: do_chain(IC(2), IC(3)); break;
                    case QB2_CODE_CHAIN + 3: do_chain(IC(3), IC(2)); break;
                    case QB2_CODE_CHAIN + 4: do_chain(IC(3), IC(3)); break;
                    case QB2_CODE_CHAIN + 5: do_chain(IC(3), IC(4)); break;
                    case QB2_CODE_CHAIN + 6: do_chain(IC(4), IC(2)); break;
                    case QB2_CODE_CHAIN + 7: do_chain(IC(4), IC(3)); break;
                    case QB2_CODE_CHAIN + 8: do_chain(IC(4), IC(4)); break;
                    case QB2_CODE_CHAIN + 9: do_chain(IC(4), IC(5)); break;
                    default: break;
                }
                TR(24 + (o < 36u ? o : 36u));
#undef IC
#undef BFLY_CASES
#undef DBIT_CASES
#undef MAT1_CASES
#undef SWAPC_CASES
            }
        }
        // ---- store with the last phase's distribution; load the next tile behind every store
        {
            TR(62);
            const uint4 c = ctx[n_phases * nthreads + tid];
            uint8_t *gb = reinterpret_cast<uint8_t *>(state) + (tbase | ((uint64_t)c.y << 32 | c.x)) * sizeof(vec);
#ifndef QB2_NOPIN
            asm volatile("" : "+l"(gb));
#endif
            const uint32_t tn = t + gridDim.x;
            if (tn < ntiles && !zero_in) {
                tbase = tile_base(tn);
                const uint8_t *gn = load_base(tbase);
                static_for<0, NR>([&](auto J) {
                    constexpr int j = decltype(J)::value;
                    __stcs(reinterpret_cast<amp *>(gb + IO->store_off[j]), a[j]);
                    a[j] = ld_na(reinterpret_cast<const amp *>(gn + IO->load_off[j]));
                });
            } else if (tn < ntiles) {
                static_for<0, NR>([&](auto J) {
                    constexpr int j = decltype(J)::value;
                    __stcs(reinterpret_cast<amp *>(gb + IO->store_off[j]), a[j]);
                });
                tbase = tile_base(tn);
                synth_tile(tbase);
            } else {
                static_for<0, NR>([&](auto J) {
                    constexpr int j = decltype(J)::value;
                    __stcs(reinterpret_cast<amp *>(gb + IO->store_off[j]), a[j]);
                });
            }
            TR(63);
        }
    }
}

template <typename real>
__global__ void init_basis_kernel(typename Traits<real>::vec *__restrict__ state, const uint64_t size,
                                  const uint64_t basis) {
    using vec = typename Traits<real>::vec;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += stride) {
        vec v;
        v.x = (i == basis) ? (real)1 : (real)0;
        v.y = 0;
        state[i] = v;
    }
}

template <typename real, int RB, int MAXT, int MINB, bool FULL>
int launch_kernel(void *state, uint64_t hi_base, const qb2_pass_header *h, const void *blob_dev, uint32_t ntiles,
                  cudaStream_t s) {
    using vec = typename Traits<real>::vec;
    using frac = typename Traits<real>::frac;
    auto kern = pass_kernel<real, RB, MAXT, MINB, FULL>;
    static bool attr_set = false;
    static int occ_cache_threads = -1, occ_cache_smem = -1, occ_cache = 0;
    const uint32_t threads = 1u << (h->T - RB);
    const uint32_t ctx_off = std::is_same<real, float>::value ? NROOT * sizeof(float2) : 0;
    if (h->n_slots > QB2_MAX_SLOTS) return set_error(QB2_ELIMIT, "pass uses %u phase slots (at most %d)", h->n_slots, QB2_MAX_SLOTS);
    const uint32_t slot_off = ctx_off + (h->n_phases + 1) * threads * 16u;
    const uint32_t slot_bytes =
        h->n_slots ? (uint32_t)(2 * QB2_MAX_SLOTS * sizeof(frac) + QB2_MAX_SLOTS * 2 + (size_t)h->n_slots * threads * sizeof(frac)) : 0u;
    const uint32_t tile_off = (slot_off + slot_bytes + 15u) & ~15u;
    // one exchange buffer when there is an exchange
    const size_t tile_bytes = h->n_phases > 1 ? ((size_t)sizeof(vec) << h->T) : 0;
    static const size_t pad_smem = getenv("QB2_PAD_SMEM") ? (size_t)atoi(getenv("QB2_PAD_SMEM")) : 0;  // experiment
    const size_t smem = tile_off + tile_bytes + pad_smem;
    if (!attr_set) {
        QB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    if (smem > 200 * 1024) return set_error(QB2_ELIMIT, "pass needs %zu bytes of shared memory", smem);
    if (occ_cache_threads != (int)threads || occ_cache_smem != (int)smem) {
        int occ = 0;
        QB2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, (int)threads, smem));
        if (occ < 1) return set_error(QB2_ELIMIT, "pass kernel does not fit on an SM (threads=%u smem=%zu)", threads, smem);
        occ_cache = occ;
        occ_cache_threads = (int)threads;
        occ_cache_smem = (int)smem;
    }
    uint32_t grid = (uint32_t)(occ_cache * sm_count());
    if (grid > ntiles) grid = ntiles;
    // the small section travels as a kernel parameter (constant bank); the tables stay in HBM
    static PassParams params;
    memcpy(&params, h, h->small_bytes);
    const frac *tables = reinterpret_cast<const frac *>(reinterpret_cast<const uint8_t *>(blob_dev) + h->tab_off);
    kern<<<grid, threads, smem, s>>>(reinterpret_cast<vec *>(state), tables, hi_base, ntiles, ctx_off, slot_off, tile_off, params);
    count_launch();
    return check_cuda(cudaGetLastError(), "pass_kernel launch");
}

template <typename real, int RB, int MAXT, int MINB>
int launch_variant(void *state, uint64_t hi_base, const qb2_pass_header *h, const void *blob_dev, uint32_t ntiles,
                   cudaStream_t s) {
    // the host says which op kinds the pass contains (header.features)
    if (h->features & QB2_FEATURE_FULL)
        return launch_kernel<real, RB, MAXT, MINB, true>(state, hi_base, h, blob_dev, ntiles, s);
    return launch_kernel<real, RB, MAXT, MINB, false>(state, hi_base, h, blob_dev, ntiles, s);
}

}  // namespace

int launch_pass(void *state, int n, uint64_t hi_base, const qb2_pass_header *h, const void *blob_dev, int dtype,
                cudaStream_t s) {
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
    const uint32_t ntiles = 1u << (n - h->T);
    const int tb = h->T - h->r;
#ifdef QB2_ONLY_F5  // development builds: one variant, for SASS inspection
    if (dtype == QB2_C64 && h->r == 5 && tb <= 8) return launch_kernel<float, 5, 256, 2, false>(state, hi_base, h, blob_dev, ntiles, s);
#else
    if (dtype == QB2_C64) {
        if (h->r == 4 && tb <= 8) return launch_variant<float, 4, 256, QB2_MINB4>(state, hi_base, h, blob_dev, ntiles, s);
        if (h->r == 4 && tb == 9) return launch_variant<float, 4, 512, 2>(state, hi_base, h, blob_dev, ntiles, s);
        if (h->r == 5 && tb <= 8) return launch_variant<float, 5, 256, 2>(state, hi_base, h, blob_dev, ntiles, s);
        if (h->r == 3 && tb <= 8) return launch_variant<float, 3, 256, 6>(state, hi_base, h, blob_dev, ntiles, s);
    } else if (dtype == QB2_C128) {
        if (h->r == 3 && tb <= 8) return launch_variant<double, 3, 256, 4>(state, hi_base, h, blob_dev, ntiles, s);
        if (h->r == 4 && tb <= 8) return launch_variant<double, 4, 256, 2>(state, hi_base, h, blob_dev, ntiles, s);
    }
#endif
    return set_error(QB2_ELIMIT, "qb2_run_pass: no kernel variant for dtype=%d r=%u T=%u", dtype, h->r, h->T);
}

#ifdef QB2_TRACE
extern "C" int qb2_debug_read_trace(unsigned long long *out, int count) {
    if (count > QB2_TRACE_TILES * QB2_TRACE_STAMPS) count = QB2_TRACE_TILES * QB2_TRACE_STAMPS;
    cudaDeviceSynchronize();
    const cudaError_t e = cudaMemcpyFromSymbol(out, qb2_trace_buf, sizeof(unsigned long long) * count);
    static unsigned long long zero[QB2_TRACE_TILES * QB2_TRACE_STAMPS];
    cudaMemcpyToSymbol(qb2_trace_buf, zero, sizeof(zero));
    return e == cudaSuccess ? count : -1;
}
#endif

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

}  // namespace qb2
