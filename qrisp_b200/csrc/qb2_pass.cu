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

template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
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
__device__ __forceinline__ void apply_mat(real (&re)[1 << RB], real (&im)[1 << RB], const real *__restrict__ m,
                                          const uint32_t jmask) {
    using BM = BitMap<RB, K, P0, P1, P2, P3>;
    constexpr int D = 1 << K;
    constexpr int NG = (1 << RB) / D;
    static_for<0, NG>([&](auto G) {
        constexpr int base = BM::base(decltype(G)::value);
        if ((jmask >> base) & 1u) {
            real xr[D], xi[D];
            static_for<0, D>([&](auto C) {
                constexpr int j = base | BM::dep(decltype(C)::value);
                xr[decltype(C)::value] = re[j];
                xi[decltype(C)::value] = im[j];
            });
            static_for<0, D>([&](auto Rw) {
                constexpr int row = decltype(Rw)::value;
                real ar = 0, ai = 0;
                static_for<0, D>([&](auto C) {
                    constexpr int col = decltype(C)::value;
                    const real mr = m[2 * (row * D + col)];
                    const real mi = m[2 * (row * D + col) + 1];
                    ar = fma(mr, xr[col], ar);
                    ar = fma(-mi, xi[col], ar);
                    ai = fma(mr, xi[col], ai);
                    ai = fma(mi, xr[col], ai);
                });
                constexpr int j = base | BM::dep(row);
                re[j] = ar;
                im[j] = ai;
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

__device__ __forceinline__ void opaque(float &x) { asm volatile("" : "+f"(x)); }
__device__ __forceinline__ void opaque(double &x) { asm volatile("" : "+d"(x)); }

template <typename V>
__device__ __forceinline__ V ld_stream(const V *p) {
    return __ldcs(p);
}
template <typename V>
__device__ __forceinline__ void st_stream(V *p, const V v) {
    __stcs(p, v);
}

// Coefficients of the op being executed: N reals from the device copy of the pass, fetched
// with 16-byte loads from one address for the whole warp (a single sector per load) into
// registers ONCE per op -- constant-bank reads indexed by a non-uniform register would be
// re-issued for every register pair.
template <typename real, int N>
__device__ __forceinline__ void load_coefs(const real *__restrict__ g, real (&c)[N]) {
    if constexpr (sizeof(real) == 4) {
        static_for<0, N / 4>([&](auto I) {
            constexpr int i = decltype(I)::value;
            const float4 v = __ldg(reinterpret_cast<const float4 *>(g) + i);
            c[4 * i] = v.x;
            c[4 * i + 1] = v.y;
            c[4 * i + 2] = v.z;
            c[4 * i + 3] = v.w;
        });
    } else {
        static_for<0, N / 2>([&](auto I) {
            constexpr int i = decltype(I)::value;
            const double2 v = __ldg(reinterpret_cast<const double2 *>(g) + i);
            c[2 * i] = v.x;
            c[2 * i + 1] = v.y;
        });
    }
}

template <typename real, int RB, int B>
__device__ __forceinline__ void mat1_general(real (&re)[1 << RB], real (&im)[1 << RB], const real (&m)[8],
                                             const uint32_t jmask) {
    static_for<0, (1 << RB) / 2>([&](auto G) {
        constexpr int g = decltype(G)::value;
        constexpr int j0 = ((g >> B) << (B + 1)) | (g & ((1 << B) - 1));
        constexpr int j1 = j0 | (1 << B);
        if ((jmask >> j0) & 1u) {
            const real x0r = re[j0], x0i = im[j0], x1r = re[j1], x1i = im[j1];
            re[j0] = fma(m[0], x0r, fma(-m[1], x0i, fma(m[2], x1r, -m[3] * x1i)));
            im[j0] = fma(m[0], x0i, fma(m[1], x0r, fma(m[2], x1i, m[3] * x1r)));
            re[j1] = fma(m[4], x0r, fma(-m[5], x0i, fma(m[6], x1r, -m[7] * x1i)));
            im[j1] = fma(m[4], x0i, fma(m[5], x0r, fma(m[6], x1i, m[7] * x1r)));
        }
    });
}

// ---- mbarrier + 1-D bulk async copy (TMA) wrappers for the tile prefetch pipeline
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct PassParams {  // the small section of a pass, passed by value: lives in the constant bank
    uint4 q[QB2_MAX_SMALL_BYTES / 16];
};

template <typename real, int RB, int B>
__device__ __forceinline__ void mat1_real(real (&re)[1 << RB], real (&im)[1 << RB], const real a, const real b,
                                          const real c, const real d, const uint32_t jmask) {
    static_for<0, (1 << RB) / 2>([&](auto G) {
        constexpr int g = decltype(G)::value;
        constexpr int j0 = ((g >> B) << (B + 1)) | (g & ((1 << B) - 1));
        constexpr int j1 = j0 | (1 << B);
        if ((jmask >> j0) & 1u) {
            const real x0r = re[j0], x0i = im[j0], x1r = re[j1], x1i = im[j1];
            re[j0] = fma(a, x0r, b * x1r);
            im[j0] = fma(a, x0i, b * x1i);
            re[j1] = fma(c, x0r, d * x1r);
            im[j1] = fma(c, x0i, d * x1i);
        }
    });
}

// anti-diagonal 2x2: new0 = p * x1, new1 = q * x0
template <typename real, int RB, int B>
__device__ __forceinline__ void mat1_anti(real (&re)[1 << RB], real (&im)[1 << RB], const real pr, const real pi,
                                          const real qr, const real qi, const uint32_t jmask) {
    static_for<0, (1 << RB) / 2>([&](auto G) {
        constexpr int g = decltype(G)::value;
        constexpr int j0 = ((g >> B) << (B + 1)) | (g & ((1 << B) - 1));
        constexpr int j1 = j0 | (1 << B);
        if ((jmask >> j0) & 1u) {
            const real x0r = re[j0], x0i = im[j0], x1r = re[j1], x1i = im[j1];
            re[j0] = fma(pr, x1r, -pi * x1i);
            im[j0] = fma(pr, x1i, pi * x1r);
            re[j1] = fma(qr, x0r, -qi * x0i);
            im[j1] = fma(qr, x0i, qi * x0r);
        }
    });
}

template <typename real, int RB, int B>
__device__ __forceinline__ void mat1_swap(real (&re)[1 << RB], real (&im)[1 << RB], const uint32_t jmask) {
    static_for<0, (1 << RB) / 2>([&](auto G) {
        constexpr int g = decltype(G)::value;
        constexpr int j0 = ((g >> B) << (B + 1)) | (g & ((1 << B) - 1));
        constexpr int j1 = j0 | (1 << B);
        if ((jmask >> j0) & 1u) {
            const real tr = re[j0], ti = im[j0];
            re[j0] = re[j1];
            im[j0] = im[j1];
            re[j1] = tr;
            im[j1] = ti;
        }
    });
}

// FULL = false leaves out the rarely used op kinds (general per-amplitude DIAG, 4x4 / 8x8 /
// 16x16 blocks): the common passes then run a kernel a third of the size, which matters
// because the interpreter's working set of code competes for the instruction cache.
// PF = true: the next tile is fetched into shared memory by bulk async copies (TMA) while the
// current one is processed, so that HBM latency overlaps the register work instead of adding
// to it.  Two tile buffers; the buffer the current tile came from doubles as its exchange
// scratch.  Needs the phase-0 distribution the planner marks with QB2_FEATURE_PREFETCH.
template <typename real, int RB, int MAXT, int MINB, bool FULL, bool PF>
__global__ void __launch_bounds__(MAXT, MINB)
    pass_kernel(typename Traits<real>::vec *__restrict__ state, const typename Traits<real>::frac *__restrict__ tables,
                const real *__restrict__ gcoefs, const uint64_t hi_base, const uint32_t ntiles, const uint32_t roots_off, const uint32_t ctx_off,
                const uint32_t tile_off, const __grid_constant__ PassParams P) {
    using vec = typename Traits<real>::vec;
    using frac = typename Traits<real>::frac;
    constexpr int NR = 1 << RB;
    extern __shared__ __align__(16) uint8_t smem[];

    const uint32_t tid = threadIdx.x;
    vec *const roots = reinterpret_cast<vec *>(smem + roots_off);
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
    uint8_t *tile_b = smem + tile_off;
    const uint8_t *const state_b = reinterpret_cast<const uint8_t *>(state);

    auto tile_base = [&](uint32_t t) {
        uint64_t tb = 0;
        for (int s = 0; s < n_seg; ++s) {
            const qb2_seg sg = H->seg[s];
            tb |= (uint64_t)((t >> sg.src) & ((1u << sg.len) - 1u)) << sg.dst;
        }
        return tb;
    };

    // ---- prefetch pipeline state: thread c moves chunk c (16 amplitudes, one 128-byte line)
    const uint32_t tile_bytes = (uint32_t)sizeof(vec) << H->T;
    uint64_t *const mbar = reinterpret_cast<uint64_t *>(smem + tile_off + 2 * tile_bytes);
    uint64_t ch_goff = 0;  // physical offset of this thread's chunk inside a tile
    uint32_t ch_soff = 0;  // byte offset of the chunk inside a tile buffer
    auto prefetch = [&](uint32_t t, uint32_t which) {
        if (tid == 0) mbar_expect_tx(&mbar[which], tile_bytes);
        bulk_g2s(smem + tile_off + which * tile_bytes + ch_soff, state_b + (tile_base(t) | ch_goff) * sizeof(vec),
                 16 * sizeof(vec), &mbar[which]);
    };
    if constexpr (PF) {
        const uint16_t *pf = a16 + H->pf;
        for (int i = 0; i < (int)H->T - 4; ++i)
            if ((tid >> i) & 1u) {
                ch_goff |= 1ull << (pf[i] & 0xffu);
                ch_soff |= (uint32_t)sizeof(vec) << (pf[i] >> 8);
            }
        if (tid == 0) {
            mbar_init(&mbar[0], 1);
            mbar_init(&mbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (blockIdx.x < ntiles) prefetch(blockIdx.x, 0);
    }

    real re[NR], im[NR];

    uint32_t iter = 0;
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++iter) {
        // ---- physical offset of the tile (tile id bits deposited into the non-tile positions)
        const uint64_t tbase = tile_base(t);
        // ---- phase 0: load
        uint64_t xloc;  // local physical index of this thread's register 0
        if constexpr (PF) {
            const uint32_t cur = iter & 1u;
            const uint4 c = ctx[tid];
            xloc = tbase | ((uint64_t)c.y << 32 | c.x);
            tile_b = smem + tile_off + cur * tile_bytes;
            mbar_wait(&mbar[cur], (iter >> 1) & 1u);
            // phase-0 layout of the buffer: thread bit b = slot bit b, register j = slot j << TB
            const uint8_t *rb0 = tile_b + tid * (uint32_t)sizeof(vec);
            const uint32_t stride = (uint32_t)sizeof(vec) << TB;
            static_for<0, NR>([&](auto J) {
                constexpr int j = decltype(J)::value;
                const vec v = *reinterpret_cast<const vec *>(rb0 + j * stride);
                re[j] = v.x;
                im[j] = v.y;
            });
            // everyone is done with both buffers (this one: just read; the other: last tile's
            // exchange scratch): hand the other one to the async proxy for the next tile
            fence_proxy_async();
            __syncthreads();
            if (t + gridDim.x < ntiles) prefetch(t + gridDim.x, cur ^ 1u);
        } else {
            const qb2_phase &ph = phases[0];
            const uint4 c = ctx[tid];
            xloc = tbase | ((uint64_t)c.y << 32 | c.x);
            const uint8_t *gb = state_b + xloc * sizeof(vec);
#ifndef QB2_NOPIN
            asm volatile("" : "+l"(gb));  // keep base + uniform offset: two adds per address
#endif
            const uint64_t *rp = a64 + ph.rphys;
            static_for<0, NR>([&](auto J) {
                constexpr int j = decltype(J)::value;
                const vec v = ld_stream(reinterpret_cast<const vec *>(gb + rp[j] * sizeof(vec)));
                re[j] = v.x;
                im[j] = v.y;
            });
        }
        for (uint32_t p = 0; p < n_phases; ++p) {
            const qb2_phase &ph = phases[p];
            if (p > 0) {
                // ---- exchange through shared memory
                const uint4 c = ctx[p * nthreads + tid];
                const uint32_t xf = ph.xflags;
                {
                    uint8_t *wp = tile_b + c.z;
                    if (xf & 1u) {  // register j sits at slot j << kw: immediate offsets
                        const uint32_t stride = (uint32_t)sizeof(vec) << (xf >> 8 & 0xffu);
                        static_for<0, NR>([&](auto J) {
                            constexpr int j = decltype(J)::value;
                            vec v;
                            v.x = re[j];
                            v.y = im[j];
                            *reinterpret_cast<vec *>(wp + j * stride) = v;
                        });
                    } else {
                        const uint16_t *xw = a16 + ph.xw + TB;
                        static_for<0, NR>([&](auto J) {
                            constexpr int j = decltype(J)::value;
                            vec v;
                            v.x = re[j];
                            v.y = im[j];
                            *reinterpret_cast<vec *>(tile_b + (c.z ^ (xw[j] * (uint32_t)sizeof(vec)))) = v;
                        });
                    }
                }
#ifndef QB2_EXP_NOSYNC
                __syncthreads();
#endif
                {
                    xloc = tbase | ((uint64_t)c.y << 32 | c.x);
                    const uint8_t *rp_ = tile_b + c.w;
                    if (xf & 2u) {
                        const uint32_t stride = (uint32_t)sizeof(vec) << (xf >> 16 & 0xffu);
                        static_for<0, NR>([&](auto J) {
                            constexpr int j = decltype(J)::value;
                            const vec v = *reinterpret_cast<const vec *>(rp_ + j * stride);
                            re[j] = v.x;
                            im[j] = v.y;
                        });
                    } else {
                        const uint16_t *xr = a16 + ph.xr + TB;
                        static_for<0, NR>([&](auto J) {
                            constexpr int j = decltype(J)::value;
                            const vec v =
                                *reinterpret_cast<const vec *>(tile_b + (c.w ^ (xr[j] * (uint32_t)sizeof(vec))));
                            re[j] = v.x;
                            im[j] = v.y;
                        });
                    }
                }
#ifndef QB2_EXP_NOSYNC
                __syncthreads();
#endif
            }
            const uint64_t X = hi_base | xloc;
            // ---- ops of this phase, on registers
            const uint32_t op_end = ph.first_op + ph.n_ops;
            for (uint32_t o = ph.first_op; o < op_end; ++o) {
                const qb2_op &op = ops[o];
                const uint32_t kind = op.kind;
                if (kind == QB2_OP_BFLY) {
                    // 2x2 gate on register bit b0, then a phase on its b0=1 output that depends on
                    // the other register bits (host-computed v[pair]) and on the thread (w1)
                    const uint32_t fl = op.flags;
                    vec w1;
                    w1.x = 1;
                    w1.y = 0;
                    if (fl & QB2_BFLY_W) {
                        const qb2_tab *tb = tabs + op.tab;
                        const int nreg = op.ntab_reg;
                        frac s1 = 0;
                        for (int i = 0; i < nreg; ++i)
                            s1 += tab_is_mul(tb[i])
                                      ? tab_mul<frac>(tb[i], X)
                                      : __ldg(tables + tb[i].off + tab_index(tb[i], X) + tb[i].regoff);
                        w1 = unit_phase(s1, roots);
                    }
                    // REAL 2x2 block only (the planner fuses nothing else); the four entries are
                    // read once, the pair factors v[] straight from the constant bank
                    const real *m = coefs + 2 * (size_t)op.data;
                    const real *v = m + 8;
                    real ma = m[0], mb = m[2], mc = m[4], md = m[6];
                    opaque(ma), opaque(mb), opaque(mc), opaque(md);  // no re-reads inside the pair loop
                    const int b0 = op.b0;
                    static_for<0, RB>([&](auto B) {
                        constexpr int bb = decltype(B)::value;
                        if (b0 == bb) {
                            static_for<0, NR / 2>([&](auto G) {
                                constexpr int g = decltype(G)::value;
                                constexpr int j0 = ((g >> bb) << (bb + 1)) | (g & ((1 << bb) - 1));
                                constexpr int j1 = j0 | (1 << bb);
                                const real x0r = re[j0], x0i = im[j0], x1r = re[j1], x1i = im[j1];
                                re[j0] = fma(ma, x0r, mb * x1r);
                                im[j0] = fma(ma, x0i, mb * x1i);
                                const real tr = fma(mc, x0r, md * x1r);
                                const real ti = fma(mc, x0i, md * x1i);
                                const real vr = v[2 * g], vi = v[2 * g + 1];
                                const real cr = fma(w1.x, vr, -w1.y * vi);
                                const real ci = fma(w1.x, vi, w1.y * vr);
                                re[j1] = fma(tr, cr, -ti * ci);
                                im[j1] = fma(tr, ci, ti * cr);
                            });
                        }
                    });
                } else if (kind >= QB2_OP_DIAG) {
                    const qb2_tab *tb = tabs + op.data;
                    frac sum = 0;
                    const int nthr = op.ntab_thr;
                    for (int i = 0; i < nthr; ++i)
                        sum += tab_is_mul(tb[i]) ? tab_mul<frac>(tb[i], X)
                                                 : __ldg(tables + tb[i].off + tab_index(tb[i], X));
                    tb += nthr;
                    const int nreg = op.ntab_reg;
                    if (kind == QB2_OP_DIAG_THR) {
                        const vec w = unit_phase(sum, roots);
                        static_for<0, NR>([&](auto J) {
                            constexpr int j = decltype(J)::value;
                            const real xr_ = re[j], xi_ = im[j];
                            re[j] = fma(xr_, w.x, -xi_ * w.y);
                            im[j] = fma(xr_, w.y, xi_ * w.x);
                        });
                    } else if (kind == QB2_OP_DIAG_BIT) {
                        // phase depends on the thread and on ONE register bit: two phases per thread
                        frac s0 = sum, s1 = sum;
                        for (int i = 0; i < nreg; ++i) {
                            if (tab_is_mul(tb[i])) {  // b0 * (mult * index bits): nothing for b0 = 0
                                s1 += tab_mul<frac>(tb[i], X);
                            } else {
                                const frac *e = tables + tb[i].off + tab_index(tb[i], X);
                                s0 += __ldg(e);
                                s1 += __ldg(e + tb[i].regoff);
                            }
                        }
                        const bool skip0 = (op.flags & 1u) != 0;
                        vec w0;
                        w0.x = 1;
                        w0.y = 0;
                        if (!skip0) w0 = unit_phase(s0, roots);
                        const vec w1 = unit_phase(s1, roots);
                        const int b0 = op.b0;
                        static_for<0, RB>([&](auto B) {
                            constexpr int bb = decltype(B)::value;
                            if (b0 == bb) {
                                static_for<0, NR>([&](auto J) {
                                    constexpr int j = decltype(J)::value;
                                    if constexpr ((j >> bb) & 1) {
                                        const real xr_ = re[j], xi_ = im[j];
                                        re[j] = fma(xr_, w1.x, -xi_ * w1.y);
                                        im[j] = fma(xr_, w1.y, xi_ * w1.x);
                                    } else {
                                        if (!skip0) {
                                            const real xr_ = re[j], xi_ = im[j];
                                            re[j] = fma(xr_, w0.x, -xi_ * w0.y);
                                            im[j] = fma(xr_, w0.y, xi_ * w0.x);
                                        }
                                    }
                                });
                            }
                        });
                    } else if (kind == QB2_OP_DIAG_VEC) {
                        // phases that depend on register bits only: host-computed, uniform
                        const real *cv = coefs + 2 * (size_t)op.data;
                        static_for<0, NR>([&](auto J) {
                            constexpr int j = decltype(J)::value;
                            const real wr = cv[2 * j], wi = cv[2 * j + 1];
                            const real xr_ = re[j], xi_ = im[j];
                            re[j] = fma(xr_, wr, -xi_ * wi);
                            im[j] = fma(xr_, wi, xi_ * wr);
                        });
                    } else if constexpr (FULL) {  // QB2_OP_DIAG: general, one table sum per amplitude
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
                            frac s = sum + __ldg(t0 + g0[j]);
                            if (nreg > 1) s += __ldg(t1 + g1[j]);
                            if (nreg > 2) s += __ldg(t2 + g2[j]);
                            if (nreg > 3) s += __ldg(t3 + g3[j]);
                            const vec w = unit_phase(s, roots);
                            const real xr_ = re[j], xi_ = im[j];
                            re[j] = fma(xr_, w.x, -xi_ * w.y);
                            im[j] = fma(xr_, w.y, xi_ * w.x);
                        });
                    }
                } else {
                  // The thread-level control below may diverge, and divergent code cannot use the
                  // uniform datapath: read the 2x2 coefficients before it.
                  const real *m = coefs + 2 * (size_t)op.data;
                  real c8[8];
                  if (kind == QB2_OP_MAT1) {
#pragma unroll
                      for (int k = 0; k < 8; ++k) c8[k] = m[k];
                  }
                  if (((X ^ op.cval) & op.cmask) == 0) {
                    const uint32_t jmask = op.jmask;
                    const int b0 = op.b0, b1 = op.b1;
                    if (kind == QB2_OP_MAT1) {
                        const uint32_t shape = op.flags;

                        if (shape == QB2_MAT1_SWAP) {
                            static_for<0, RB>([&](auto B) {
                                if (b0 == decltype(B)::value) mat1_swap<real, RB, decltype(B)::value>(re, im, jmask);
                            });
                        } else if (shape == QB2_MAT1_REAL) {
                            const real a = c8[0], b = c8[2], c = c8[4], d = c8[6];
                            static_for<0, RB>([&](auto B) {
                                if (b0 == decltype(B)::value)
                                    mat1_real<real, RB, decltype(B)::value>(re, im, a, b, c, d, jmask);
                            });
                        } else if (shape == QB2_MAT1_ANTI) {
                            const real pr = c8[2], pi = c8[3], qr = c8[4], qi = c8[5];
                            static_for<0, RB>([&](auto B) {
                                if (b0 == decltype(B)::value)
                                    mat1_anti<real, RB, decltype(B)::value>(re, im, pr, pi, qr, qi, jmask);
                            });
                        } else {
                            static_for<0, RB>([&](auto B) {
                                if (b0 == decltype(B)::value)
                                    mat1_general<real, RB, decltype(B)::value>(re, im, c8, jmask);
                            });
                        }
                    } else if constexpr (FULL) {
                        if (kind == QB2_OP_MAT2) {
                            static_for<0, RB>([&](auto B0) {
                                static_for<decltype(B0)::value + 1, RB>([&](auto B1) {
                                    if (b0 == decltype(B0)::value && b1 == decltype(B1)::value)
                                        apply_mat<real, RB, 2, decltype(B0)::value, decltype(B1)::value>(re, im, m, jmask);
                                });
                            });
                        } else if (kind == QB2_OP_MAT3) {
                            if constexpr (RB >= 3) apply_mat<real, RB, 3, 0, 1, 2>(re, im, m, jmask);
                        } else if (kind == QB2_OP_MAT4) {
                            if constexpr (RB >= 4) apply_mat<real, RB, 4, 0, 1, 2, 3>(re, im, m, jmask);
                        }
                    }
                  }
                }
            }
        }
        // ---- store with the last phase's distribution
        {
            const uint64_t *rp = a64 + phases[n_phases - 1].rphys;
            uint8_t *gb = reinterpret_cast<uint8_t *>(state) + xloc * sizeof(vec);
#ifndef QB2_NOPIN
            asm volatile("" : "+l"(gb));
#endif
            static_for<0, NR>([&](auto J) {
                constexpr int j = decltype(J)::value;
                vec v;
                v.x = re[j];
                v.y = im[j];
                st_stream(reinterpret_cast<vec *>(gb + rp[j] * sizeof(vec)), v);
            });
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

struct Variant {
    const void *fn;
    int rb, maxt;
    bool attr_set;
};

template <typename real, int RB, int MAXT, int MINB, bool FULL, bool PF>
int launch_kernel(void *state, uint64_t hi_base, const qb2_pass_header *h, const void *blob_dev, uint32_t ntiles,
                  cudaStream_t s) {
    using vec = typename Traits<real>::vec;
    using frac = typename Traits<real>::frac;
    auto kern = pass_kernel<real, RB, MAXT, MINB, FULL, PF>;
    static bool attr_set = false;
    static int occ_cache_threads = -1, occ_cache_smem = -1, occ_cache = 0;
    const uint32_t threads = 1u << (h->T - RB);
    const uint32_t roots_off = 0;
    const uint32_t ctx_off = std::is_same<real, float>::value ? NROOT * sizeof(float2) : 0;
    const uint32_t tile_off = ctx_off + h->n_phases * threads * 16u;
    // PF: two tile buffers + two mbarriers; else one exchange buffer when there is an exchange
    const size_t tile_bytes = PF ? 2 * ((size_t)sizeof(vec) << h->T) + 16
                                 : (h->n_phases > 1 ? ((size_t)sizeof(vec) << h->T) : 0);
    const size_t smem = tile_off + tile_bytes;
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
    const real *gcoefs = reinterpret_cast<const real *>(reinterpret_cast<const uint8_t *>(blob_dev) + h->coef_off);
    kern<<<grid, threads, smem, s>>>(reinterpret_cast<vec *>(state), tables, gcoefs, hi_base, ntiles, roots_off, ctx_off,
                                     tile_off, params);
    count_launch();
    return check_cuda(cudaGetLastError(), "pass_kernel launch");
}

template <typename real, int RB, int MAXT, int MINB>
int launch_variant(void *state, uint64_t hi_base, const qb2_pass_header *h, const void *blob_dev, uint32_t ntiles,
                   cudaStream_t s) {
    // the host says which op kinds the pass contains and whether its phase-0 distribution
    // suits the prefetch pipeline (header.features)
    static const bool no_pf = getenv("QB2_PREFETCH") == nullptr;  // experimental: slower than direct loads on B200
    if constexpr (std::is_same<real, float>::value && RB == 4 && MAXT == 256) {
        if ((h->features & QB2_FEATURE_PREFETCH) && h->T == 12 && !no_pf && ntiles >= 2) {
            if (h->features & QB2_FEATURE_FULL)
                return launch_kernel<real, RB, MAXT, 2, true, true>(state, hi_base, h, blob_dev, ntiles, s);
            return launch_kernel<real, RB, MAXT, 3, false, true>(state, hi_base, h, blob_dev, ntiles, s);
        }
    }
    if (h->features & QB2_FEATURE_FULL)
        return launch_kernel<real, RB, MAXT, MINB, true, false>(state, hi_base, h, blob_dev, ntiles, s);
    return launch_kernel<real, RB, MAXT, MINB, false, false>(state, hi_base, h, blob_dev, ntiles, s);
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
    if (dtype == QB2_C64) {
        if (h->r == 4 && tb <= 8) return launch_variant<float, 4, 256, QB2_MINB4>(state, hi_base, h, blob_dev, ntiles, s);
        if (h->r == 4 && tb == 9) return launch_variant<float, 4, 512, 2>(state, hi_base, h, blob_dev, ntiles, s);
        if (h->r == 5 && tb <= 8) return launch_variant<float, 5, 256, 2>(state, hi_base, h, blob_dev, ntiles, s);
        if (h->r == 3 && tb <= 8) return launch_variant<float, 3, 256, 6>(state, hi_base, h, blob_dev, ntiles, s);
    } else if (dtype == QB2_C128) {
        if (h->r == 3 && tb <= 8) return launch_variant<double, 3, 256, 4>(state, hi_base, h, blob_dev, ntiles, s);
        if (h->r == 4 && tb <= 8) return launch_variant<double, 4, 256, 2>(state, hi_base, h, blob_dev, ntiles, s);
    }
    return set_error(QB2_ELIMIT, "qb2_run_pass: no kernel variant for dtype=%d r=%u T=%u", dtype, h->r, h->T);
}

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
