// qb2_pass_kernel.cuh -- the register-pipelined pass kernel: every geometry and both dtypes.
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
#pragma once
#include "qb2_pass_common.cuh"

namespace qb2 {
namespace {

// The tables of a specialised kernel (generated translation unit: qrisp_b200/jit.py); dummies otherwise.
#ifndef QB2_JIT_NPHASES
#define QB2_JIT_NPHASES 1
#define QB2_JIT_PHASES_FILE "qb2_jit_none.inc"
#define QB2_JIT_TB 8
__host__ __device__ constexpr uint32_t jit_ndec(uint32_t) { return 0; }
__host__ __device__ constexpr uint32_t jit_sched(uint32_t, uint32_t) { return 0; }
__host__ __device__ constexpr qb2_op jit_op(uint32_t) { return qb2_op{}; }
__host__ __device__ constexpr qb2_phase jit_phase(uint32_t) { return qb2_phase{}; }
__host__ __device__ constexpr uint32_t jit_a16(uint32_t) { return 0; }
#define QB2_JIT_DUMMIES 1
#endif

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
// (`pb`: the small section of the pass, the by-value kernel parameter; JIT: a specialised kernel whose phases, op
// records, decode schedule and slot maps are compile-time constants of the generated translation unit, see
// qb2_pass_tma.cuh and qrisp_b200/jit.py)
template <typename real, int RB, bool FULL, bool JIT>
__device__ __forceinline__ void pass_body(typename Traits<real>::vec *__restrict__ state,
                                          const typename Traits<real>::frac *__restrict__ tables, const uint64_t hi_base,
                                          const uint32_t ntiles, const uint32_t ctx_off, const uint32_t slot_off,
                                          const uint32_t tile_off, const SparseInput &SP, const uint32_t sp_off,
                                          const uint8_t *const pb) {
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
    const qb2_pass_header *const H = reinterpret_cast<const qb2_pass_header *>(pb);
    const qb2_io *const IO = reinterpret_cast<const qb2_io *>(pb + QB2_IO_OFFSET);  // fixed offsets: immediate operands
    const qb2_phase *const phases = reinterpret_cast<const qb2_phase *>(pb + H->phases_off);
    const qb2_op *const ops = reinterpret_cast<const qb2_op *>(pb + H->ops_off);
    const real *const coefs = reinterpret_cast<const real *>(pb + H->coef_off);
    const uint16_t *const a16 = reinterpret_cast<const uint16_t *>(pb + H->u16_off);
    const uint64_t *const a64 = reinterpret_cast<const uint64_t *>(pb + H->u64_off);
    const qb2_tab *const tabs = reinterpret_cast<const qb2_tab *>(pb + H->tabdesc_off);
    const int TB = JIT ? (int)QB2_JIT_TB : (int)H->T - RB;  // thread bits
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
    // QB2_FEATURE_ZERO_INPUT: the state has not been written yet: nothing is read, the tile is built from
    // the listed non-zero amplitudes (|0...0>: one entry).  A tile without an entry stays zero under
    // every op of the pass (the ops are linear), so its phases are skipped and zeros are stored.
    const bool zero_in = (H->features & QB2_FEATURE_ZERO_INPUT) != 0u;
    uint32_t *const sp_tile = reinterpret_cast<uint32_t *>(smem + sp_off);
    if (zero_in) {
        for (uint32_t i = tid; i < SP.count; i += blockDim.x) sp_tile[i] = SP.tile[i];
        __syncthreads();
    }
    bool tile_zero = false;
    auto synth_tile = [&](const uint32_t tl) {
        tile_zero = sparse_synth<real, RB>(a, sp_tile, SP, tl, tid);
    };
    uint32_t iter = 0;
    uint32_t t = blockIdx.x;
    uint64_t tbase = 0;
    if (t < ntiles) {
        tbase = tile_base(t);
        if (zero_in) {
            synth_tile(t);
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
#define QB2_SLOT_W(sl) unit_phase((frac)(sthr[(sl) * nthreads + tid] + stile[spar + (sl)]), roots)
        if constexpr (JIT) {
#define QB2_OP_DECL(name, idx) constexpr qb2_op name = jit_op(idx)
#define QB2_PHASE_DECL(name, idx) constexpr qb2_phase name = jit_phase(idx)
#define QB2_A16(i) jit_a16(i)
#define QB2_SWITCH(x) { constexpr uint32_t qb2_code_ = (x);
#define QB2_CASE(c) if constexpr (qb2_code_ == (uint32_t)(c))
#define QB2_BREAK
#define QB2_DEFAULT
#define QB2_SWITCH_END }
#define QB2_ADVANCE(k)
            if (!tile_zero) {
#define QB2_PHASE_FILE "qb2_pass_kernel_phase.inc"
#define QB2_DEP(x) ((x) + (JIT ? 0u : 0u))  /* value-dependent: the arms not taken are never instantiated */
#include QB2_JIT_PHASES_FILE
#undef QB2_DEP
#undef QB2_PHASE_FILE
            }
#undef QB2_OP_DECL
#undef QB2_PHASE_DECL
#undef QB2_A16
#undef QB2_SWITCH
#undef QB2_CASE
#undef QB2_BREAK
#undef QB2_DEFAULT
#undef QB2_SWITCH_END
#undef QB2_ADVANCE
#undef QB2_OPS_FILE
        } else {
#define QB2_OP_DECL(name, idx) const qb2_op &name = ops[idx]
#define QB2_PHASE_DECL(name, idx) const qb2_phase &name = phases[idx]
#define QB2_A16(i) a16[i]
#define QB2_SWITCH(x) switch (x) {
#define QB2_CASE(c) case (c):
#define QB2_BREAK break;
#define QB2_DEFAULT default: break;
#define QB2_SWITCH_END }
#define QB2_ADVANCE(k) o += (k)
#define QB2_OPS_FILE "qb2_pass_ops_loop.inc"
            for (uint32_t p = 0; p < n_phases && !tile_zero; ++p) {
#include "qb2_pass_kernel_phase.inc"
            }
#undef QB2_OP_DECL
#undef QB2_PHASE_DECL
#undef QB2_A16
#undef QB2_SWITCH
#undef QB2_CASE
#undef QB2_BREAK
#undef QB2_DEFAULT
#undef QB2_SWITCH_END
#undef QB2_ADVANCE
#undef QB2_OPS_FILE
        }
#undef QB2_SLOT_W
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
                synth_tile(tn);
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

template <typename real, int RB, int MAXT, int MINB, bool FULL>
__global__ void __launch_bounds__(MAXT, MINB)
    pass_kernel(typename Traits<real>::vec *__restrict__ state, const typename Traits<real>::frac *__restrict__ tables,
                const uint64_t hi_base, const uint32_t ntiles, const uint32_t ctx_off, const uint32_t slot_off,
                const uint32_t tile_off, const SparseInput SP, const uint32_t sp_off, const __grid_constant__ PassParams P) {
    pass_body<real, RB, FULL, false>(state, tables, hi_base, ntiles, ctx_off, slot_off, tile_off, SP, sp_off,
                                     reinterpret_cast<const uint8_t *>(&P));
}

template <typename real, int RB, int MAXT, int MINB, bool FULL>
int launch_kernel(const PassLaunch &L) {
    using vec = typename Traits<real>::vec;
    using frac = typename Traits<real>::frac;
    auto kern = pass_kernel<real, RB, MAXT, MINB, FULL>;
    const qb2_pass_header *h = L.h;
    // per-device caches: the shared-memory opt-in and the occupancy belong to a (function, device) pair
    static bool attr_set[QB2_MAX_DEVICES] = {false};
    static int occ_cache_threads[QB2_MAX_DEVICES], occ_cache_smem[QB2_MAX_DEVICES], occ_cache[QB2_MAX_DEVICES];
    int dev = 0;
    QB2_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= QB2_MAX_DEVICES) return set_error(QB2_ELIMIT, "device ordinal %d outside the engine's table", dev);
    const uint32_t threads = 1u << (h->T - RB);
    const uint32_t ctx_off = std::is_same<real, float>::value ? NROOT * sizeof(float2) : 0;
    if (h->n_slots > QB2_MAX_SLOTS) return set_error(QB2_ELIMIT, "pass uses %u phase slots (at most %d)", h->n_slots, QB2_MAX_SLOTS);
    const uint32_t slot_off = ctx_off + (h->n_phases + 1) * threads * 16u;
    const uint32_t slot_bytes =
        h->n_slots ? (uint32_t)(2 * QB2_MAX_SLOTS * sizeof(frac) + QB2_MAX_SLOTS * 2 + (size_t)h->n_slots * threads * sizeof(frac)) : 0u;
    const uint32_t sp_off = (slot_off + slot_bytes + 15u) & ~15u;
    const uint32_t sp_bytes = (h->features & QB2_FEATURE_ZERO_INPUT) ? ((L.sparse.count * 4u + 15u) & ~15u) : 0u;
    const uint32_t tile_off = sp_off + sp_bytes;
    // one exchange buffer when there is an exchange
    const size_t tile_bytes = h->n_phases > 1 ? ((size_t)sizeof(vec) << h->T) : 0;
    const size_t smem = tile_off + tile_bytes;
    if (!attr_set[dev]) {
        QB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[dev] = true;
        occ_cache_threads[dev] = -1;
    }
    if (smem > 200 * 1024) return set_error(QB2_ELIMIT, "pass needs %zu bytes of shared memory", smem);
    if (occ_cache_threads[dev] != (int)threads || occ_cache_smem[dev] != (int)smem) {
        int occ = 0;
        QB2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, (int)threads, smem));
        if (occ < 1) return set_error(QB2_ELIMIT, "pass kernel does not fit on an SM (threads=%u smem=%zu)", threads, smem);
        occ_cache[dev] = occ;
        occ_cache_threads[dev] = (int)threads;
        occ_cache_smem[dev] = (int)smem;
    }
    uint32_t grid = (uint32_t)(occ_cache[dev] * sm_count());
    if (grid > L.ntiles) grid = L.ntiles;
    // the small section travels as a kernel parameter (constant bank); the tables stay in HBM.  The copy
    // is per call (stack): two host threads may launch at the same time
    PassParams params;
    memcpy(&params, h, h->small_bytes);
    const frac *tables = reinterpret_cast<const frac *>(reinterpret_cast<const uint8_t *>(L.blob_dev) + h->tab_off);
    if (L.jit_kernel) {
        void *st = L.state;
        uint64_t hi = L.hi_base;
        uint32_t nt = L.ntiles, co = ctx_off, so = slot_off, to = tile_off, spo = sp_off;
        SparseInput sp = L.sparse;
        void *args[] = {&st, &tables, &hi, &nt, &co, &so, &to, &sp, &spo, &params};
        return launch_jit_kernel(L.jit_kernel, grid, threads, smem, L.s, args);
    }
    kern<<<grid, threads, smem, L.s>>>(reinterpret_cast<vec *>(L.state), tables, L.hi_base, L.ntiles, ctx_off, slot_off, tile_off,
                                       L.sparse, sp_off, params);
    count_launch();
    return check_cuda(cudaGetLastError(), "pass_kernel launch");
}

}  // namespace
}  // namespace qb2
