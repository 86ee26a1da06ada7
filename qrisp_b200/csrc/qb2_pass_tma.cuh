// qb2_pass_tma.cuh -- the TMA pass kernel (complex64, T = 13 tile bits, r = 5 register bits).
//
// Same pass format and the same op interpreter (qb2_pass_ops.inc) as the register-pipelined kernel of
// qb2_pass_kernel.cuh; what differs is how a tile travels:
//
//   HBM --one cp.async.bulk.tensor (64 KB box, mbarrier)--> landing buffer --LDS--> registers
//       --ops--> [group buffer exchange --ops-->]* --STS--> group buffer --one bulk tensor store--> HBM
//
// One CTA per SM, two compute groups of 256 threads (named barriers), three 64 KB tile buffers with the
// 128-byte TMA swizzle: the landing buffer is being filled with the next tile while both groups compute,
// and each group's buffer (exchanges, then source of the store) drains to HBM while the group works on
// its next tile.  Loads and stores cost the compute warps no issue slots and no address arithmetic, and
// no warp ever waits on a global load.
//
// A tile is any set of 13 index bits that contains bits 0..3 (whole 128-byte lines).  The tensor map
// (host: make_tile_map) has dim0 = bits 0..3, a "base" dimension of box size 1 whose coordinate carries
// all non-tile bits of the tile's address, and one dimension (stride 2**p amplitudes) per run of tile bits
// p.. -- up to three runs of at most 8 bits.  In shared memory the box is 512 rows of 128 bytes, 16-byte
// chunk c of row r stored at chunk c ^ (r & 7); the host expresses that as XOR columns (phase 0's xr / xw
// maps of the pass), the same form every exchange uses.
#pragma once
#include <cuda.h>  // CUtensorMap and the encoder's prototype; the entry point comes from cudaGetDriverEntryPoint

#include "qb2_pass_common.cuh"

namespace qb2 {
namespace {

constexpr int TMA_T = 13;
constexpr int TMA_GROUP = 256;
constexpr uint32_t TMA_TILE_BYTES = 8u << TMA_T;
constexpr int TMA_WIN = 1024;  // shared-window address of the dynamic shared memory: the driver's reserved 1 KB comes first
constexpr uint32_t TMA_AUX_BYTES = 232448u - 3u * TMA_TILE_BYTES - 1024u;  // what is left of 227 KB (1 KB kept free)

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
// rank-5 tile copies; only coordinate 1 (the base dimension) is ever non-zero
__device__ __forceinline__ void tma_load_tile(void *dst, const CUtensorMap *m, uint64_t *bar, int base) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %3, %3, %3}], [%2];"
        ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(0), "r"(base) : "memory");
}
__device__ __forceinline__ void tma_store_tile(const CUtensorMap *m, const void *src, int base) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %2, %2, %2}], [%1];"
                 ::"l"(m), "r"(smem_u32(src)), "r"(0), "r"(base) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void group_bar(uint32_t g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1u), "n"(TMA_GROUP) : "memory"); }

// The decode schedule of a specialised kernel (generated translation unit: qrisp_b200/jit.py); dummies here.
#ifndef QB2_JIT_NPHASES
#define QB2_JIT_NPHASES 1
#define QB2_JIT_DUMMIES 1
#define QB2_JIT_PHASES_FILE "qb2_jit_none.inc"
__host__ __device__ constexpr uint32_t jit_ndec(uint32_t) { return 0; }
__host__ __device__ constexpr uint32_t jit_sched(uint32_t, uint32_t) { return 0; }
__host__ __device__ constexpr qb2_op jit_op(uint32_t) { return qb2_op{}; }
__host__ __device__ constexpr qb2_phase jit_phase(uint32_t) { return qb2_phase{}; }
__host__ __device__ constexpr uint32_t jit_a16(uint32_t) { return 0; }
#endif

template <int RB>
__device__ __forceinline__ GrayCols<RB> jit_gray_cols(const uint32_t first) {
    GrayCols<RB> g;
#pragma unroll
    for (int i = 0; i < RB; ++i) g.d[i] = jit_a16(first + (1u << i)) * (uint32_t)sizeof(float2);
    return g;
}

// `pb`: the small section of the pass -- the by-value kernel parameter (constant bank) in the ahead-of-time
// kernels, a const array with a known initialiser in a specialised one.  `features`: header.features as the
// launcher wants them (QB2_FEATURE_ZERO_INPUT is decided per launch).
template <bool FULL, bool JIT>
__device__ __forceinline__ void pass_tma_body(const CUtensorMap &tmap, const uint32_t *__restrict__ tables, const uint64_t hi_base,
                                              const uint32_t ntiles, const uint32_t ctx_off, const uint32_t slot_off,
                                              const uint32_t sp_off, const uint32_t inv_mask, const uint32_t features,
                                              const SparseInput &SP, const uint8_t *const pb) {
    using real = float;
    using vec = float2;
    using frac = uint32_t;
    using C = Cx<real>;
    using amp = typename C::amp;
    constexpr int RB = 5;
    constexpr int NR = 1 << RB;
    constexpr int TB = TMA_T - RB;
    // dynamic shared memory: the three tile buffers first (the 128B swizzle pattern repeats every 1 KB, so they
    // sit at the 1 KB-aligned start of the window), the descriptors' working data behind them.  Tile
    // accesses are smem[offset] with the buffer's offset folded into the per-thread slot offset: the
    // base of the array is an immediate of the LDS / STS, not an add per access.
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *const aux = smem + 3u * TMA_TILE_BYTES;

    const uint32_t g = threadIdx.x >> 8;     // compute group
    const uint32_t tid = threadIdx.x & 255u;  // thread of the group
    constexpr uint32_t nthreads = TMA_GROUP;
    if (smem_u32(smem) != (uint32_t)TMA_WIN) __trap();  // (the launcher has checked the reserved size: cannot happen)
    vec *const roots = reinterpret_cast<vec *>(aux);
    for (uint32_t i = threadIdx.x; i < NROOT; i += blockDim.x) {
        float s, c;
        sincospif((float)i * (2.0f / NROOT), &s, &c);
        roots[i] = make_float2(c, s);
    }
    uint64_t *const full_bar = reinterpret_cast<uint64_t *>(aux + NROOT * sizeof(vec));  // [2]: one per group
    float *const norm_part = reinterpret_cast<float *>(full_bar + 2);                    // [2 groups][8 warps]

    const qb2_pass_header *const H = reinterpret_cast<const qb2_pass_header *>(pb);
    const qb2_phase *const phases = reinterpret_cast<const qb2_phase *>(pb + H->phases_off);
    const qb2_op *const ops = reinterpret_cast<const qb2_op *>(pb + H->ops_off);
    const real *const coefs = reinterpret_cast<const real *>(pb + H->coef_off);
    const uint16_t *const a16 = reinterpret_cast<const uint16_t *>(pb + H->u16_off);
    const uint64_t *const a64 = reinterpret_cast<const uint64_t *>(pb + H->u64_off);
    const qb2_tab *const tabs = reinterpret_cast<const qb2_tab *>(pb + H->tabdesc_off);
    const uint32_t n_phases = H->n_phases;
    const int n_seg = H->n_seg;

    uint8_t *const landing = smem;                              // buffer 0
    const uint32_t gofs = (1u + g) * TMA_TILE_BYTES;            // this group's buffer: 1 or 2
    uint8_t *const gbuf = smem + gofs;

    // ---- per-thread context of every phase (the two groups share it): low 32 bits of the thread's physical
    // offset | slot under the store side of the entering exchange (13 bits), slot under its load side
    // (13 bits), bits 32.. of the offset (6 bits).  Phase 0's slots are those of the landing buffer (load
    // side) and of the final store (store side, under the LAST phase's distribution).
    uint2 *const ctx = reinterpret_cast<uint2 *>(aux + ctx_off);
    if (g == 0) {
        for (uint32_t p = 0; p < n_phases; ++p) {
            const qb2_phase &ph = phases[p];
            uint64_t xt = 0;
            uint32_t wb = 0, rb = 0;
            for (int b = 0; b < TB; ++b)
                if ((tid >> b) & 1u) {
                    xt |= a64[ph.pcol + b];
                    wb ^= a16[ph.xw + b];
                    rb ^= a16[ph.xr + b];
                }
            ctx[p * nthreads + tid] = make_uint2((uint32_t)xt, wb | (rb << 13) | ((uint32_t)(xt >> 32) << 26));
        }
    }
    auto ctx_x = [&](const uint2 c) { return (uint64_t)(c.y >> 26) << 32 | c.x; };
    // ---- phase slots: thread parts once per CTA (shared by the groups), tile parts once per tile and group
    const uint32_t n_slots = H->n_slots;
    auto two_slots = [&](const qb2_op &op) { return op.kind == QB2_OP_DIAG_BIT && !(op.flags & 1u); };
    frac *const stile_all = reinterpret_cast<frac *>(aux + slot_off);                       // [2 groups][2][QB2_MAX_SLOTS]
    frac *const stile = stile_all + g * 2 * QB2_MAX_SLOTS;
    uint16_t *const slot_op = reinterpret_cast<uint16_t *>(stile_all + 4 * QB2_MAX_SLOTS);   // [QB2_MAX_SLOTS]
    uint16_t *const slot_ofs = slot_op + QB2_MAX_SLOTS;                                      // [QB2_MAX_SLOTS]
    // Thread parts: slot s at slot_area + slot_ofs[s] KB.  A slot whose ladders touch no bit outside the tile
    // (bit s of inv_mask, worked out by the launcher) is the same for every tile of the pass: its unit phase is
    // computed ONCE per CTA and kept as a float2 per thread (2 KB per slot); the other slots keep their
    // turn fraction (1 KB per slot) and get the tile part and the root-table evaluation per tile.
    uint8_t *const slot_area = reinterpret_cast<uint8_t *>(slot_ofs + QB2_MAX_SLOTS);
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
        if (threadIdx.x == 0) {
            uint32_t kb = 0;
            for (uint32_t sl = 0; sl < n_slots; ++sl) {
                slot_ofs[sl] = (uint16_t)kb;
                kb += ((inv_mask >> sl) & 1u) ? 2u : 1u;
            }
        }
        __syncthreads();  // the root table and the slot offsets
        if (g == 0) {
            for (uint32_t p = 0; p < n_phases; ++p) {
                uint64_t xt = 0;
                for (int b = 0; b < TB; ++b)
                    if ((tid >> b) & 1u) xt |= a64[phases[p].pcol + b];
                const uint64_t x = hi_base | xt;
                const uint32_t op_end = phases[p].first_op + phases[p].n_ops;
                for (uint32_t o = phases[p].first_op; o < op_end; ++o) {
                    const qb2_op &op = ops[o];
                    if (has_slot(op)) {
                        const uint32_t ns = two_slots(op) ? 2u : 1u;
                        for (uint32_t w = 0; w < ns; ++w) {
                            const uint32_t sl = op.b1 - 1u + w;
                            const frac thr = ladder_sum(op, w, x);
                            uint8_t *const at = slot_area + (uint32_t)slot_ofs[sl] * 1024u;
                            if ((inv_mask >> sl) & 1u)
                                reinterpret_cast<float2 *>(at)[tid] = unit_phase(thr, roots);
                            else
                                reinterpret_cast<frac *>(at)[tid] = thr;
                            if (tid == 0) slot_op[sl] = (uint16_t)(o | (w << 15));
                        }
                    }
                }
            }
        }
    }
    const uint32_t var_mask = (n_slots >= 32u ? 0xffffffffu : ((1u << n_slots) - 1u)) & ~inv_mask;  // slots with a tile part
    // ---- sparse input (first pass after a reset): the launcher has filled the whole shard with zeros (a
    // plain streaming fill runs at 7.5 TB/s, tile-shaped bulk stores at 4.7); a tile without an entry stays
    // zero under every op of the pass (the ops are linear), so the kernel only visits the tiles that have
    // entries: the DISTINCT tile numbers of the (sorted) list, worked out here once per CTA.
    const bool zero_in = (features & QB2_FEATURE_ZERO_INPUT) != 0u;
    uint32_t *const sp_tile = reinterpret_cast<uint32_t *>(aux + sp_off);
    uint32_t *const sp_uniq = sp_tile + ((SP.count + 3u) & ~3u);
    uint32_t n_uniq = 0;
    if (zero_in) {
        for (uint32_t i = threadIdx.x; i < SP.count; i += blockDim.x) sp_tile[i] = SP.tile[i];
        __syncthreads();
        // distinct tiles by warp ballots: chunk c = entries [32c, 32c+32); flag = first entry of its tile
        uint32_t *const sp_cnt = sp_uniq + ((SP.count + 3u) & ~3u);  // [chunks] distinct tiles per chunk
        const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
        const uint32_t chunks = (SP.count + 31u) >> 5;
        for (uint32_t c = wid; c < chunks; c += nw) {
            const uint32_t i = c * 32u + lane;
            const bool first = i < SP.count && (i == 0u || sp_tile[i] != sp_tile[i - 1u]);
            const uint32_t m = __ballot_sync(0xffffffffu, first);
            if (lane == 0u) sp_cnt[c] = (uint32_t)__popc(m);
        }
        __syncthreads();
        for (uint32_t c = wid; c < chunks; c += nw) {
            uint32_t base = 0;
            for (uint32_t d = 0; d < c; ++d) base += sp_cnt[d];
            const uint32_t i = c * 32u + lane;
            const bool first = i < SP.count && (i == 0u || sp_tile[i] != sp_tile[i - 1u]);
            const uint32_t m = __ballot_sync(0xffffffffu, first);
            if (first) sp_uniq[base + (uint32_t)__popc(m & ((1u << lane) - 1u))] = sp_tile[i];
        }
        for (uint32_t d = 0; d < chunks; ++d) n_uniq += sp_cnt[d];
    }
    if (threadIdx.x == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();

    auto tile_base = [&](uint32_t t) {
        uint64_t tb = 0;
        for (int s = 0; s < n_seg; ++s) {
            const qb2_seg sg = H->seg[s];
            tb |= (uint64_t)((t >> sg.src) & ((1u << sg.len) - 1u)) << sg.dst;
        }
        return tb;
    };
    // k-th tile of this CTA: loaded into the landing buffer for group k & 1
    auto issue_load = [&](const uint32_t k) {
        const uint32_t t = blockIdx.x + k * gridDim.x;
        if (t >= ntiles) return;
        mbar_expect_tx(&full_bar[k & 1u], TMA_TILE_BYTES);
        tma_load_tile(landing, &tmap, &full_bar[k & 1u], (int)(tile_base(t) >> 4));
    };
    if (threadIdx.x == 0 && !zero_in) issue_load(0);

    const uint2 c0 = ctx[tid];
    const uint64_t xthr0 = ctx_x(c0);
    const uint32_t land_rb = ((c0.y >> 13) & 0x1fffu) * (uint32_t)sizeof(vec);
    const uint32_t store_wb = (c0.y & 0x1fffu) * (uint32_t)sizeof(vec);
    // phase 0's maps = the landing buffer (load side) and the final store (store side)
    uint32_t io_xf;
    GrayCols<RB> io_rcols, io_wcols;
    if constexpr (JIT) {
        constexpr qb2_phase ph0 = jit_phase(0);
        io_xf = ph0.xflags;
        io_rcols = jit_gray_cols<RB>(ph0.xr + TB);
        io_wcols = jit_gray_cols<RB>(ph0.xw + TB);
    } else {
        io_xf = phases[0].xflags;
        io_rcols = gray_cols<RB>(a16 + phases[0].xr + TB);
        io_wcols = gray_cols<RB>(a16 + phases[0].xw + TB);
    }

    amp a[NR];
    uint32_t kpar = 0;  // parity of this group's full barrier
    uint32_t iter = 0;
    uint32_t spar = 0;  // which of the two tile-part buffers this tile uses
    auto slot_w = [&](const uint32_t sl) -> amp {
        const uint8_t *const at = slot_area + (uint32_t)slot_ofs[sl] * 1024u;
        if ((inv_mask >> sl) & 1u) return reinterpret_cast<const float2 *>(at)[tid];
        return unit_phase((frac)(reinterpret_cast<const frac *>(at)[tid] + stile[spar + sl]), roots);
    };
    for (uint32_t k = g;; k += 2u, ++iter) {
        uint32_t t = blockIdx.x + k * gridDim.x;
        if (zero_in) {
            if (t >= n_uniq) break;
            t = sp_uniq[t];  // only tiles with an entry
        } else if (t >= ntiles) {
            break;
        }
        const uint64_t tbase = tile_base(t);
        uint64_t xloc = tbase | xthr0;
        // ---- tile part of the slotted phases: one thread per slot
        spar = (iter & 1u) * QB2_MAX_SLOTS;
        if (var_mask) {
            for (uint32_t sl = tid >> 5; sl < n_slots; sl += nthreads >> 5)
                if ((tid & 31u) == 0u && ((var_mask >> sl) & 1u)) {
                    const uint32_t so = slot_op[sl];
                    stile[spar + sl] = ladder_sum(ops[so & 0x7fffu], so >> 15, tbase);
                }
        }
        bool tile_zero = false;
        if (zero_in) {
            tile_zero = sparse_synth<real, RB>(a, sp_tile, SP, t, tid);
            if (var_mask) group_bar(g);
        } else {
            mbar_wait(&full_bar[g], kpar);
            kpar ^= 1u;
            // (the offsets are the same for every tile: an opaque copy keeps the compiler from hoisting 32
            // addresses out of the tile loop and spilling them)
            uint32_t rb_t = land_rb;
            asm volatile("" : "+r"(rb_t));
            smem_load_any<RB, TMA_WIN>(rb_t, (io_xf & 2u) != 0u, io_xf >> 16 & 0xffu, io_rcols, a);
            group_bar(g);  // the landing buffer is free again (and the slots' tile parts are visible)
            if (tid == 0) issue_load(k + 1u);
        }
        // the group's buffer may still be the source of the previous tile's store
        bool gbuf_free = false;
        auto acquire_gbuf = [&]() {
            if (!gbuf_free) {
                if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                group_bar(g);
                gbuf_free = true;
            }
        };
        if (tile_zero) continue;  // (cannot happen: only listed tiles are visited)
#define QB2_SLOT_W(sl) slot_w(sl)
        if constexpr (JIT) {
            // specialised kernel: phases and decode points are compile-time constants (jit_* tables of the
            // generated translation unit), so every descriptor read folds and only the arms in use remain
#define QB2_OP_DECL(name, idx) constexpr qb2_op name = jit_op(idx)
#define QB2_PHASE_DECL(name, idx) constexpr qb2_phase name = jit_phase(idx)
#define QB2_GRAY_COLS(first) jit_gray_cols<RB>(first)
#define QB2_SWITCH(x) { constexpr uint32_t qb2_code_ = (x);
#define QB2_CASE(c) if constexpr (qb2_code_ == (uint32_t)(c))
#define QB2_BREAK
#define QB2_DEFAULT
#define QB2_SWITCH_END }
#define QB2_ADVANCE(k)
            // (textual unrolling, no lambdas around the phases: the registers must stay registers)
#define QB2_PHASE_FILE "qb2_pass_tma_phase.inc"
#define QB2_DEP(x) ((x) + (JIT ? 0u : 0u))  /* value-dependent: the arms not taken are never instantiated */
#include QB2_JIT_PHASES_FILE
#undef QB2_DEP
#undef QB2_PHASE_FILE
#undef QB2_OP_DECL
#undef QB2_PHASE_DECL
#undef QB2_GRAY_COLS
#undef QB2_SWITCH
#undef QB2_CASE
#undef QB2_BREAK
#undef QB2_DEFAULT
#undef QB2_SWITCH_END
#undef QB2_OPS_FILE
#undef QB2_ADVANCE
        } else {
#define QB2_OP_DECL(name, idx) const qb2_op &name = ops[idx]
#define QB2_PHASE_DECL(name, idx) const qb2_phase &name = phases[idx]
#define QB2_GRAY_COLS(first) gray_cols<RB>(a16 + (first))
#define QB2_SWITCH(x) switch (x) {
#define QB2_CASE(c) case (c):
#define QB2_BREAK break;
#define QB2_DEFAULT default: break;
#define QB2_SWITCH_END }
#define QB2_ADVANCE(k) o += (k)
#define QB2_OPS_FILE "qb2_pass_ops_loop.inc"
            for (uint32_t p = 0; p < n_phases; ++p) {
#include "qb2_pass_tma_phase.inc"
            }
#undef QB2_OP_DECL
#undef QB2_PHASE_DECL
#undef QB2_GRAY_COLS
#undef QB2_SWITCH
#undef QB2_CASE
#undef QB2_BREAK
#undef QB2_DEFAULT
#undef QB2_SWITCH_END
#undef QB2_OPS_FILE
#undef QB2_ADVANCE
        }
#undef QB2_SLOT_W
        // ---- the tile goes to the group's buffer in the layout of the tensor map, then ONE bulk store
        acquire_gbuf();
        {
            uint32_t wb_t = gofs + store_wb;
            asm volatile("" : "+r"(wb_t));
            smem_store_any<RB, TMA_WIN>(wb_t, (io_xf & 1u) != 0u, io_xf >> 8 & 0xffu, io_wcols, a);
        }
        if (SP.tile_norms) {
            // |amp|**2 of the tile as it leaves (the sampler's first level: saves its read of the whole state)
            float part = 0.f;
            static_for<0, NR>([&](auto J) {
                constexpr int j = decltype(J)::value;
                part = fmaf(a[j].x, a[j].x, fmaf(a[j].y, a[j].y, part));
            });
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if ((tid & 31u) == 0u) norm_part[g * 8u + (tid >> 5)] = part;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        group_bar(g);
        if (tid == 0) {
            tma_store_tile(&tmap, gbuf, (int)(tbase >> 4));
            if (SP.tile_norms) {
                double sum = 0.0;
#pragma unroll
                for (int w = 0; w < 8; ++w) sum += (double)norm_part[g * 8u + w];
                SP.tile_norms[t] = sum;
            }
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <bool FULL>
__global__ void __launch_bounds__(2 * TMA_GROUP, 1)
    pass_kernel_tma(const __grid_constant__ CUtensorMap tmap, const uint32_t *__restrict__ tables, const uint64_t hi_base,
                    const uint32_t ntiles, const uint32_t ctx_off, const uint32_t slot_off, const uint32_t sp_off,
                    const uint32_t inv_mask, const uint32_t features, const SparseInput SP, const __grid_constant__ PassParams P) {
    pass_tma_body<FULL, false>(tmap, tables, hi_base, ntiles, ctx_off, slot_off, sp_off, inv_mask, features, SP,
                               reinterpret_cast<const uint8_t *>(&P));
}

// streaming fill of a shard with zeros: 16-byte stores, four per thread and iteration
__global__ void __launch_bounds__(256) zero_fill_kernel(uint4 *__restrict__ p, const size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        __stcs(p + i, z);
        __stcs(p + i + stride, z);
        __stcs(p + i + 2 * stride, z);
        __stcs(p + i + 3 * stride, z);
    }
    for (; i < n16; i += stride) __stcs(p + i, z);
}

inline int launch_zero_fill(void *state, size_t bytes, cudaStream_t s) {
    const size_t n16 = bytes / 16;
    size_t blocks = (n16 + 1023) / 1024;
    const size_t cap = (size_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    zero_fill_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<uint4 *>(state), n16);
    count_launch();
    return check_cuda(cudaGetLastError(), "zero_fill launch");
}

typedef CUresult (*tensor_map_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                         const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Tensor map of a pass's tiles over the local state (see the top of this file).  Returns false when the
// tile does not fit a rank-5 map (more than three runs of tile bits above bit 3).
inline bool make_tile_map(CUtensorMap *tm, void *state, int n, const qb2_pass_header *h, tensor_map_encode_fn encode) {
    bool nontile[64] = {false};
    for (int s = 0; s < h->n_seg; ++s)
        for (int b = 0; b < h->seg[s].len; ++b) nontile[h->seg[s].dst + b] = true;
    for (int p = 0; p < 4; ++p)
        if (nontile[p]) return false;
    cuuint64_t gdim[5] = {32, 1, 1, 1, 1}, gstr[4] = {128, 128, 128, 128};
    cuuint32_t box[5] = {32, 1, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
    gdim[1] = 1ull << (n - 4);  // base dimension: 128-byte lines, box 1
    int nd = 2;
    for (int p = 4; p < n;) {
        if (nontile[p]) {
            ++p;
            continue;
        }
        int q = p;
        while (q < n && !nontile[q] && q - p < 8) ++q;
        if (nd >= 5) return false;
        gdim[nd] = 1ull << (n - p);
        gstr[nd - 1] = 8ull << p;
        box[nd] = 1u << (q - p);
        ++nd;
        p = q;
    }
    for (; nd < 5; ++nd) gstr[nd - 1] = 8ull << n;  // unused dimensions: size 1
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, state, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// QB2_ELIMIT + "tma:" message = the caller falls back to the register-pipelined kernel
template <bool FULL>
int launch_kernel_tma(const PassLaunch &L) {
    auto kern = pass_kernel_tma<FULL>;
    const qb2_pass_header *h = L.h;
    static tensor_map_encode_fn encode = nullptr;
    static bool attr_set[QB2_MAX_DEVICES] = {false};
    int dev = 0;
    QB2_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= QB2_MAX_DEVICES) return set_error(QB2_ELIMIT, "device ordinal %d outside the engine's table", dev);
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
            return set_error(QB2_ELIMIT, "tma: the driver has no cuTensorMapEncodeTiled");
        encode = reinterpret_cast<tensor_map_encode_fn>(fn);
    }
    const uint32_t ctx_off = NROOT * sizeof(float2) + 128u;  // roots, then 2 barriers + 16 partial norms
    const uint32_t slot_off = ctx_off + h->n_phases * TMA_GROUP * 8u;
    if (h->n_slots > QB2_MAX_SLOTS) return set_error(QB2_ELIMIT, "pass uses %u phase slots (at most %d)", h->n_slots, QB2_MAX_SLOTS);
    // slots whose ladders read no bit outside the tile: their phase does not change from tile to tile, the kernel
    // evaluates it once per CTA (2 KB of shared memory per such slot instead of 1 KB: as many as fit)
    uint32_t inv_mask = 0;
    {
        const uint8_t *hb = reinterpret_cast<const uint8_t *>(h);
        uint64_t nontile = 0;
        for (int sgi = 0; sgi < h->n_seg; ++sgi) nontile |= ((1ull << h->seg[sgi].len) - 1ull) << h->seg[sgi].dst;
        const qb2_op *ops = reinterpret_cast<const qb2_op *>(hb + h->ops_off);
        const qb2_tab *tabs = reinterpret_cast<const qb2_tab *>(hb + h->tabdesc_off);
        for (uint32_t o = 0; o < h->n_ops; ++o) {
            const qb2_op &op = ops[o];
            if (op.b1 == 0 || !(op.kind == QB2_OP_BFLY || op.kind == QB2_OP_DIAG_THR || op.kind == QB2_OP_DIAG_BIT)) continue;
            const bool two = op.kind == QB2_OP_DIAG_BIT && !(op.flags & 1u);
            const qb2_tab *tb = tabs + (op.kind == QB2_OP_BFLY ? (uint32_t)op.tab : (uint32_t)op.data);
            for (uint32_t w = 0; w < (two ? 2u : 1u); ++w) {
                const int cnt = (two && w == 0u) ? op.ntab_thr : op.ntab_thr + op.ntab_reg;
                bool inv = true;
                for (int i = 0; i < cnt && inv; ++i) {
                    const qb2_mul &m = reinterpret_cast<const qb2_mul &>(tb[i]);
                    if (!m.is_mul || (((uint64_t)m.mask << m.shift) & nontile) != 0ull) inv = false;
                }
                const uint32_t sl = op.b1 - 1u + w;
                if (inv && sl < 32u) inv_mask |= 1u << sl;
            }
        }
    }
    static const bool no_inv = getenv("QB2_NO_SLOT_INV") != nullptr;  // (A/B switch for measurements)
    if (no_inv) inv_mask = 0;
    const uint32_t slot_fixed = 4u * QB2_MAX_SLOTS * 4u + 2u * QB2_MAX_SLOTS * 2u;
    // sparse input: the tile numbers and the list of distinct ones
    const uint32_t sp_bytes = (h->features & QB2_FEATURE_ZERO_INPUT) ? 2u * ((L.sparse.count * 4u + 15u) & ~15u) + 256u : 0u;
    auto slot_bytes_for = [&](uint32_t mask) { return slot_fixed + (h->n_slots + (uint32_t)__builtin_popcount(mask)) * TMA_GROUP * 4u; };
    while (inv_mask && ((slot_off + slot_bytes_for(inv_mask) + 15u) & ~15u) + sp_bytes > TMA_AUX_BYTES)
        inv_mask &= ~(1u << (31 - __builtin_clz(inv_mask)));  // drop the highest slot until the rest fits
    const uint32_t slot_bytes = slot_bytes_for(inv_mask);
    const uint32_t sp_off = (slot_off + slot_bytes + 15u) & ~15u;
    const uint32_t aux_bytes = sp_off + sp_bytes;
    if (aux_bytes > TMA_AUX_BYTES) return set_error(QB2_ELIMIT, "tma: descriptors need %u bytes of shared memory", aux_bytes);
    const size_t smem = (size_t)aux_bytes + 3u * TMA_TILE_BYTES;
    CUtensorMap tm;
    if (!make_tile_map(&tm, L.state, (int)h->n, h, encode)) return set_error(QB2_ELIMIT, "tma: tile shape not expressible");
    if (!attr_set[dev]) {
        int reserved = 0;
        QB2_CUDA(cudaDeviceGetAttribute(&reserved, cudaDevAttrReservedSharedMemoryPerBlock, dev));
        if (reserved != TMA_WIN) return set_error(QB2_ELIMIT, "tma: %d bytes of reserved shared memory (the kernel assumes %d)", reserved, TMA_WIN);
        QB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        attr_set[dev] = true;
    }
    uint32_t grid = (uint32_t)sm_count();
    if (grid > L.ntiles) grid = L.ntiles;
    if (h->features & QB2_FEATURE_ZERO_INPUT) {
        // the state has not been written yet: zeros everywhere first (one streaming fill), then the pass visits
        // the tiles that hold an entry of the sparse input
        if (int rc = launch_zero_fill(L.state, (size_t)8 << h->n, L.s)) return rc;
        if (L.sparse.count == 0) return QB2_OK;  // (a rank without entries: its shard is all zeros)
        if (grid > L.sparse.count) grid = L.sparse.count;
    }
    PassParams params;
    memcpy(&params, h, h->small_bytes);
    const uint32_t *tables = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(L.blob_dev) + h->tab_off);
    if (L.jit_kernel) {
        // a kernel specialised for this pass's structure: same body, same arguments
        uint64_t hi = L.hi_base;
        uint32_t nt = L.ntiles, co = ctx_off, so = slot_off, spo = sp_off, im = inv_mask, ft = h->features;
        SparseInput sp = L.sparse;
        void *args[] = {&tm, &tables, &hi, &nt, &co, &so, &spo, &im, &ft, &sp, &params};
        return launch_jit_kernel(L.jit_kernel, grid, 2 * TMA_GROUP, smem, L.s, args);
    }
    kern<<<grid, 2 * TMA_GROUP, smem, L.s>>>(tm, tables, L.hi_base, L.ntiles, ctx_off, slot_off, sp_off, inv_mask, h->features, L.sparse, params);
    count_launch();
    return check_cuda(cudaGetLastError(), "pass_kernel_tma launch");
}

}  // namespace
}  // namespace qb2
