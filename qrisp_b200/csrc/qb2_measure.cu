// qb2_measure.cu -- measurement, sampling, layout and the big-matrix correctness path.
//
//   norm2 / probabilities : BiArray.squared_norm, DenseBiArray.multi_measure
//                           (reference: bi_arrays.py:1030-1059, bi_array_helper.py:315-387)
//   sample_prepare / draw : numpy.random.choice over cl_prob (simulator.py:190-201)
//   permute_bits          : TensorFactor.swap / unravel (tensor_factor.py:56-64, :163-169)
//   matrix_big            : TensorFactor.apply_matrix for unitaries wider than a pass holds
#include "qb2_internal.cuh"

namespace qb2 {

namespace {

struct BitRuns {  // index transform: out |= ((in >> src) & mask(len)) << dst
    int n;
    uint8_t src[48], dst[48], len[48];
};

__device__ __forceinline__ uint64_t apply_runs(const BitRuns &r, const uint64_t in) {
    uint64_t out = 0;
    for (int i = 0; i < r.n; ++i) out |= ((in >> r.src[i]) & ((1ull << r.len[i]) - 1ull)) << r.dst[i];
    return out;
}

template <typename vec>
__device__ __forceinline__ double abs2(const vec v) {
    return (double)v.x * (double)v.x + (double)v.y * (double)v.y;
}

// ------------------------------------------------------------------ permute_bits
template <typename vec>
__global__ void permute_bits_kernel(const vec *__restrict__ src, vec *__restrict__ dst, const uint64_t size,
                                    const BitRuns runs) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < size; j += stride)
        dst[j] = src[apply_runs(runs, j)];
}

// ------------------------------------------------------------------ qubit swap over peer memory
// One kernel = the local bit permutation AND the all-to-all of a qubit swap: every rank PULLS its new shard
// straight out of its peers' HBM (NVLink / NVSwitch peer mappings).  Destination index j = (s, low) with s the
// top g bits: the 16-byte unit comes from rank s, index runs((rank, low)) -- what
// permute_bits followed by all_to_all_single would have delivered, without the extra HBM pass on the sender
// and without a staging copy.  Units of 16 bytes: the lowest positions are never evicted, so they stay put.
struct PeerPtrs {
    const uint4 *p[QB2_MAX_PEERS];
};

__global__ void __launch_bounds__(256)
    swap_pull_kernel(const PeerPtrs peers, uint4 *__restrict__ dst, const uint64_t units, const int low_bits,
                     const uint64_t rank_hi, const BitRuns runs, const int first_peer) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t low_mask = (1ull << low_bits) - 1ull;
    const uint64_t peer_mask = (units >> low_bits) - 1ull;
    constexpr int U = 4;  // loads in flight per thread: NVLink round trips are long
    // (peers are visited starting with the one after this rank, so the ranks do not all hit the same source at once)
    for (uint64_t k0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k0 < units; k0 += U * stride) {
        uint4 v[U];
        uint64_t d[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t k = k0 + (uint64_t)u * stride;
            if (k < units) {
                const uint64_t s = ((k >> low_bits) + (uint64_t)first_peer) & peer_mask;
                const uint64_t low = k & low_mask;
                const uint4 *src = peers.p[s] + apply_runs(runs, rank_hi | low);
                asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                             : "l"(src));
                d[u] = (s << low_bits) | low;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (k0 + (uint64_t)u * stride < units) dst[d[u]] = v[u];
    }
}

// ------------------------------------------------------------------ norm2
__device__ __forceinline__ double block_sum(double v, double *sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double t = (threadIdx.x < (unsigned)nw) ? sh[threadIdx.x] : 0.0;
    if (w == 0)
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    __syncthreads();
    return t;  // valid in warp 0
}

template <typename vec>
__global__ void norm2_kernel(const vec *__restrict__ state, const uint64_t size, double *__restrict__ out) {
    __shared__ double sh[32];
    double acc = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += stride) acc += abs2(state[i]);
    const double t = block_sum(acc, sh);
    if (threadIdx.x == 0) atomicAdd(out, t);
}

// ------------------------------------------------------------------ probabilities
struct MesDesc {
    int m;
    int8_t pos[64];  // pos[k]: physical position of outcome bit (m-1-k)
};

constexpr int PROB_CHUNK_BITS = 12;
constexpr int PROB_THREADS = 256;

template <typename vec>
__global__ void __launch_bounds__(PROB_THREADS)
    probabilities_kernel(const vec *__restrict__ state, const int n, const int C, const uint64_t hi_base,
                         const MesDesc md, double *__restrict__ probs) {
    extern __shared__ double vals[];
    __shared__ int8_t lo_out[PROB_CHUNK_BITS];  // outcome bit of the measured local positions, ascending position
    __shared__ int n_lo;
    const uint32_t tid = threadIdx.x;
    const uint32_t csize = 1u << C;
    const uint64_t nchunks = 1ull << (n - C);
    if (tid == 0) {
        int a = 0;
        for (int p = 0; p < C; ++p)
            for (int k = 0; k < md.m; ++k)
                if (md.pos[k] == p) lo_out[a++] = (int8_t)(md.m - 1 - k);
        n_lo = a;
    }
    __syncthreads();
    const int a = n_lo;
    for (uint64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const vec *src = state + (chunk << C);
        for (uint32_t i = tid; i < csize; i += PROB_THREADS) vals[i] = abs2(__ldcs(src + i));
        __syncthreads();
        // fold away the non-measured local positions, highest first
        uint32_t S = csize;
        for (int p = C - 1; p >= 0; --p) {
            bool measured = false;
            for (int k = 0; k < md.m; ++k) measured |= (md.pos[k] == p);
            if (measured) continue;
            // current index of position p = number of positions below it (none of them folded yet)
            const int q = p;
            const uint32_t half = S >> 1;
            double tmp[(1 << PROB_CHUNK_BITS) / 2 / PROB_THREADS];
#pragma unroll
            for (int it = 0; it < (1 << PROB_CHUNK_BITS) / 2 / PROB_THREADS; ++it) {
                const uint32_t k = tid + it * PROB_THREADS;
                if (k < half) {
                    const uint32_t i0 = ((k >> q) << (q + 1)) | (k & ((1u << q) - 1u));
                    tmp[it] = vals[i0] + vals[i0 | (1u << q)];
                }
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < (1 << PROB_CHUNK_BITS) / 2 / PROB_THREADS; ++it) {
                const uint32_t k = tid + it * PROB_THREADS;
                if (k < half) vals[k] = tmp[it];
            }
            __syncthreads();
            S = half;
        }
        // outcome bits fixed by the chunk number and the rank bits
        const uint64_t X = hi_base | (chunk << C);
        uint64_t o_hi = 0;
        for (int k = 0; k < md.m; ++k)
            if (md.pos[k] >= C) o_hi |= ((X >> md.pos[k]) & 1ull) << (md.m - 1 - k);
        for (uint32_t k = tid; k < S; k += PROB_THREADS) {
            uint64_t o = o_hi;
            for (int b = 0; b < a; ++b) o |= (uint64_t)((k >> b) & 1u) << lo_out[b];
            const double v = vals[k];
            if (v != 0.0) atomicAdd(probs + o, v);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ sampling
constexpr int SAMPLE_THREADS = 256;

template <typename vec>
__global__ void __launch_bounds__(SAMPLE_THREADS)
    chunk_sums_kernel(const vec *__restrict__ state, const int cb, const uint64_t nchunks, double *__restrict__ sums) {
    __shared__ double sh[32];
    const uint32_t csize = 1u << cb;
    for (uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const vec *src = state + (c << cb);
        double acc = 0;
        if constexpr (sizeof(vec) == 8) {
            // complex64: two amplitudes per 16-byte load, four loads in flight per thread (same double
            // arithmetic per amplitude as abs2(): the draw kernel re-adds the same numbers)
            auto two = [](const float4 v) {
                return (double)v.x * (double)v.x + (double)v.y * (double)v.y + ((double)v.z * (double)v.z + (double)v.w * (double)v.w);
            };
            const float4 *src4 = reinterpret_cast<const float4 *>(src);
            const uint32_t n4 = csize >> 1;
            uint32_t i = threadIdx.x;
            for (; i + 3 * SAMPLE_THREADS < n4; i += 4 * SAMPLE_THREADS) {
                const float4 a = __ldcs(src4 + i), b = __ldcs(src4 + i + SAMPLE_THREADS),
                             c4 = __ldcs(src4 + i + 2 * SAMPLE_THREADS), d = __ldcs(src4 + i + 3 * SAMPLE_THREADS);
                acc += (two(a) + two(b)) + (two(c4) + two(d));
            }
            for (; i < n4; i += SAMPLE_THREADS) acc += two(__ldcs(src4 + i));
            if (csize == 1u && threadIdx.x == 0) acc = abs2(__ldcs(src));
        } else {
            for (uint32_t i = threadIdx.x; i < csize; i += SAMPLE_THREADS) acc += abs2(__ldcs(src + i));
        }
        const double t = block_sum(acc, sh);
        if (threadIdx.x == 0) sums[c] = t;
    }
}

// in-place inclusive scan of the chunk sums over many CTAs: block totals, a scan of the totals by one CTA,
// then every block scans its own 2048 entries on top of its offset (coalesced, double accumulation)
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_PER_THREAD = 8;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_PER_THREAD;
constexpr int SCAN_MAX_BLOCKS = 1 << 16;  // scanned by one CTA in the middle step

__device__ __forceinline__ double block_scan_exclusive(const double v, double *sh, double &total) {
    // exclusive scan of one value per thread over a CTA of SCAN_THREADS threads; `total` on every thread
    double x = v;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) sh[w] = x;
    __syncthreads();
    if (w == 0) {
        double t = lane < SCAN_THREADS / 32 ? sh[lane] : 0.0;
        for (int o = 1; o < 32; o <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += y;
        }
        if (lane < SCAN_THREADS / 32) sh[lane] = t;
    }
    __syncthreads();
    const double before = w > 0 ? sh[w - 1] : 0.0;
    total = sh[SCAN_THREADS / 32 - 1];
    __syncthreads();
    return before + x - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_totals_kernel(const double *__restrict__ v, const uint64_t n, double *__restrict__ totals) {
    __shared__ double sh[32];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK;
    double acc = 0;
    for (int k = 0; k < SCAN_PER_THREAD; ++k) {
        const uint64_t i = base + (uint64_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) acc += v[i];
    }
    double total;
    block_scan_exclusive(acc, sh, total);
    if (threadIdx.x == 0) totals[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_offsets_kernel(double *__restrict__ totals, const uint32_t nblocks) {
    // exclusive scan of the block totals, in place, by one CTA (at most SCAN_MAX_BLOCKS entries)
    __shared__ double sh[32];
    double run = 0;
    for (uint32_t base = 0; base < nblocks; base += SCAN_THREADS) {
        const uint32_t i = base + threadIdx.x;
        const double x = i < nblocks ? totals[i] : 0.0;
        double total;
        const double ex = block_scan_exclusive(x, sh, total);
        if (i < nblocks) totals[i] = run + ex;
        run += total;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(double *__restrict__ v, const uint64_t n, const double *__restrict__ offsets) {
    __shared__ double sh[32];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_PER_THREAD;
    double x[SCAN_PER_THREAD];
    double acc = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; ++k) {
        x[k] = base + k < n ? v[base + k] : 0.0;
        acc += x[k];
    }
    double total;
    double run = offsets[blockIdx.x] + block_scan_exclusive(acc, sh, total);
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; ++k) {
        run += x[k];
        if (base + k < n) v[base + k] = run;
    }
}

template <typename vec>
__global__ void __launch_bounds__(SAMPLE_THREADS)
    sample_draw_kernel(const vec *__restrict__ state, const int cb, const uint64_t nchunks,
                       const double *__restrict__ cums, const double *__restrict__ uniforms, const int64_t shots,
                       uint64_t *__restrict__ indices) {
    __shared__ double incl[SAMPLE_THREADS];
    __shared__ uint64_t s_chunk;
    __shared__ double s_resid;
    __shared__ int s_pick, s_last;
    const uint32_t tid = threadIdx.x;
    const uint32_t csize = 1u << cb;
    const uint32_t seg = csize >= SAMPLE_THREADS ? csize / SAMPLE_THREADS : 1;
    const double total = cums[nchunks - 1];
    for (int64_t shot = blockIdx.x; shot < shots; shot += gridDim.x) {
        if (tid == 0) {
            double target = uniforms[shot] * total;
            const double below = nextafter(total, 0.0);
            if (target > below) target = below;
            if (target < 0) target = 0;
            // first chunk whose inclusive sum exceeds the target
            uint64_t lo = 0, hi = nchunks - 1;
            while (lo < hi) {
                const uint64_t mid = (lo + hi) >> 1;
                if (cums[mid] > target)
                    hi = mid;
                else
                    lo = mid + 1;
            }
            s_chunk = lo;
            s_resid = target - (lo > 0 ? cums[lo - 1] : 0.0);
            s_pick = -1;
            s_last = -1;
        }
        __syncthreads();
        const uint64_t c = s_chunk;
        const double resid = s_resid;
        const vec *src = state + (c << cb);
        const uint32_t start = tid * seg;
        double mine = 0;
        if (start < csize)
            for (uint32_t i = 0; i < seg; ++i) mine += abs2(src[start + i]);
        incl[tid] = mine;
        __syncthreads();
        for (uint32_t o = 1; o < SAMPLE_THREADS; o <<= 1) {
            const double add = (tid >= o) ? incl[tid - o] : 0.0;
            __syncthreads();
            incl[tid] += add;
            __syncthreads();
        }
        const double my_incl = incl[tid];
        const double my_excl = my_incl - mine;
        if (mine > 0.0) {
            atomicMax(&s_last, (int)tid);
            if (my_incl > resid && my_excl <= resid) atomicMax(&s_pick, (int)tid);
        }
        __syncthreads();
        const int pick = s_pick >= 0 ? s_pick : s_last;
        if ((int)tid == pick) {
            double run = my_excl;
            uint32_t chosen = start, last_nz = start;
            bool found = false;
            for (uint32_t i = 0; i < seg; ++i) {
                const double p = abs2(src[start + i]);
                if (p > 0.0) last_nz = start + i;
                run += p;
                if (!found && run > resid && p > 0.0) {
                    chosen = start + i;
                    found = true;
                }
            }
            indices[shot] = (c << cb) | (uint64_t)(found ? chosen : last_nz);
        }
        if (pick < 0 && tid == 0) indices[shot] = (c << cb);  // all-zero chunk: cannot happen for a normalised state
        __syncthreads();
    }
}

// ------------------------------------------------------------------ big dense matrix (correctness path)
struct BigDesc {
    int k;
    int8_t tgt[16];
};

template <typename real>
__global__ void __launch_bounds__(256)
    matrix_big_kernel(const typename Traits<real>::vec *__restrict__ src, typename Traits<real>::vec *__restrict__ dst,
                      const uint64_t size, const uint64_t hi_base, const typename Traits<real>::vec *__restrict__ M,
                      const BigDesc bd, const uint64_t cmask, const uint64_t cval) {
    using vec = typename Traits<real>::vec;
    extern __shared__ uint64_t dep[];  // column -> physical offset
    const uint32_t D = 1u << bd.k;
    uint64_t tmask = 0;
    for (int t = 0; t < bd.k; ++t) tmask |= 1ull << bd.tgt[t];
    for (uint32_t c = threadIdx.x; c < D; c += blockDim.x) {
        uint64_t v = 0;
        for (int t = 0; t < bd.k; ++t) v |= (uint64_t)((c >> t) & 1u) << bd.tgt[t];
        dep[c] = v;
    }
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += stride) {
        if ((((hi_base | i) ^ cval) & cmask) != 0) {
            dst[i] = src[i];
            continue;
        }
        uint32_t row = 0;
        for (int t = 0; t < bd.k; ++t) row |= (uint32_t)((i >> bd.tgt[t]) & 1ull) << t;
        const uint64_t base = i & ~tmask;
        const vec *mrow = M + (size_t)row * D;
        real ar = 0, ai = 0;
        for (uint32_t c = 0; c < D; ++c) {
            const vec m = mrow[c];
            const vec x = src[base | dep[c]];
            ar = fma(m.x, x.x, ar);
            ar = fma(-m.y, x.y, ar);
            ai = fma(m.x, x.y, ai);
            ai = fma(m.y, x.x, ai);
        }
        vec o;
        o.x = ar;
        o.y = ai;
        dst[i] = o;
    }
}

unsigned grid_for(uint64_t items, unsigned threads, unsigned per_sm) {
    uint64_t blocks = (items + threads - 1) / threads;
    const uint64_t cap = (uint64_t)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace

int launch_swap_pull(void *dst, const void *const *peers, int world, int rank, int n, const int *perm, int dtype, cudaStream_t s) {
    int g = 0;
    while ((1 << g) < world) ++g;
    const int unit_bits = dtype == QB2_C64 ? 1 : 0;  // amplitudes per 16-byte unit, as a bit count
    if (!dst || !peers || world < 2 || world > QB2_MAX_PEERS || (1 << g) != world || rank < 0 || rank >= world || n < g + 2 || n > 40)
        return set_error(QB2_EINVAL, "qb2_swap_pull: bad arguments");
    // index transform in amplitudes: bit p of the source index = bit perm[p] of (rank, low); identity when perm is NULL
    int pm[40];
    uint64_t seen = 0;
    for (int p = 0; p < n; ++p) {
        pm[p] = perm ? perm[p] : p;
        if (pm[p] < 0 || pm[p] >= n || ((seen >> pm[p]) & 1)) return set_error(QB2_EINVAL, "qb2_swap_pull: not a permutation");
        seen |= 1ull << pm[p];
    }
    if (unit_bits && pm[0] != 0) return set_error(QB2_EINVAL, "qb2_swap_pull: position 0 must stay in place (16-byte units)");
    BitRuns r;  // in 16-byte units: drop the lowest amplitude bit of a complex64 state
    r.n = 0;
    int p = unit_bits;
    while (p < n) {
        int len = 1;
        while (p + len < n && pm[p + len] == pm[p] + len) ++len;
        r.src[r.n] = (uint8_t)(pm[p] - unit_bits);
        r.dst[r.n] = (uint8_t)(p - unit_bits);
        r.len[r.n] = (uint8_t)len;
        ++r.n;
        p += len;
    }
    PeerPtrs pp;
    for (int k = 0; k < world; ++k) {
        if (!peers[k]) return set_error(QB2_EINVAL, "qb2_swap_pull: null peer pointer");
        pp.p[k] = reinterpret_cast<const uint4 *>(peers[k]);
    }
    const int nu = n - unit_bits;  // bits of a unit index
    const uint64_t units = 1ull << nu;
    const unsigned grid = grid_for(units, 256, 16);
    swap_pull_kernel<<<grid, 256, 0, s>>>(pp, reinterpret_cast<uint4 *>(dst), units, nu - g, (uint64_t)rank << (nu - g), r, rank + 1);
    count_launch();
    return check_cuda(cudaGetLastError(), "swap_pull launch");
}

int launch_permute_bits(const void *src, void *dst, int n, const int *perm, int dtype, cudaStream_t s) {
    if (!src || !dst || !perm || n < 1 || n > 40 || src == dst)
        return set_error(QB2_EINVAL, "qb2_permute_bits: bad arguments");
    uint64_t seen = 0;
    for (int p = 0; p < n; ++p) {
        if (perm[p] < 0 || perm[p] >= n || ((seen >> perm[p]) & 1)) return set_error(QB2_EINVAL, "qb2_permute_bits: not a permutation");
        seen |= 1ull << perm[p];
    }
    // bit p of i (source) = bit perm[p] of j (destination): i = sum ((j >> perm[p]) & 1) << p, merged into runs
    BitRuns r;
    r.n = 0;
    int p = 0;
    while (p < n) {
        int len = 1;
        while (p + len < n && perm[p + len] == perm[p] + len) ++len;
        r.src[r.n] = (uint8_t)perm[p];
        r.dst[r.n] = (uint8_t)p;
        r.len[r.n] = (uint8_t)len;
        ++r.n;
        p += len;
    }
    const uint64_t size = 1ull << n;
    const unsigned grid = grid_for(size, 256, 16);
    if (dtype == QB2_C64)
        permute_bits_kernel<float2><<<grid, 256, 0, s>>>((const float2 *)src, (float2 *)dst, size, r);
    else
        permute_bits_kernel<double2><<<grid, 256, 0, s>>>((const double2 *)src, (double2 *)dst, size, r);
    count_launch();
    return check_cuda(cudaGetLastError(), "permute_bits launch");
}

int launch_norm2(const void *state, int n, double *out_dev, int dtype, cudaStream_t s) {
    if (!state || !out_dev || n < 0 || n > 40) return set_error(QB2_EINVAL, "qb2_norm2: bad arguments");
    QB2_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double), s));
    const uint64_t size = 1ull << n;
    const unsigned grid = grid_for(size, 256, 8);
    if (dtype == QB2_C64)
        norm2_kernel<float2><<<grid, 256, 0, s>>>((const float2 *)state, size, out_dev);
    else
        norm2_kernel<double2><<<grid, 256, 0, s>>>((const double2 *)state, size, out_dev);
    count_launch();
    return check_cuda(cudaGetLastError(), "norm2 launch");
}

int launch_probabilities(const void *state, int n, uint64_t hi_base, const int *mes_pos, int m, double *probs_dev,
                         int dtype, cudaStream_t s) {
    if (!state || !probs_dev || n < 1 || n > 40 || m < 0 || m > 62 || (m > 0 && !mes_pos))
        return set_error(QB2_EINVAL, "qb2_probabilities: bad arguments");
    MesDesc md;
    md.m = m;
    uint64_t seen = 0;
    for (int k = 0; k < m; ++k) {
        if (mes_pos[k] < 0 || mes_pos[k] > 62 || ((seen >> mes_pos[k]) & 1))
            return set_error(QB2_EINVAL, "qb2_probabilities: bad or duplicate position %d", mes_pos[k]);
        seen |= 1ull << mes_pos[k];
        md.pos[k] = (int8_t)mes_pos[k];
    }
    const int C = n < PROB_CHUNK_BITS ? n : PROB_CHUNK_BITS;
    const uint64_t nchunks = 1ull << (n - C);
    uint64_t blocks = nchunks;
    const uint64_t cap = (uint64_t)sm_count() * 6;
    if (blocks > cap) blocks = cap;
    const size_t smem = sizeof(double) << C;
    if (dtype == QB2_C64)
        probabilities_kernel<float2><<<(unsigned)blocks, PROB_THREADS, smem, s>>>((const float2 *)state, n, C, hi_base, md, probs_dev);
    else
        probabilities_kernel<double2><<<(unsigned)blocks, PROB_THREADS, smem, s>>>((const double2 *)state, n, C, hi_base, md, probs_dev);
    count_launch();
    return check_cuda(cudaGetLastError(), "probabilities launch");
}

int sample_chunk_bits(int n) {
    int cb = n < 12 ? n : 12;
    if (n - 18 > cb) cb = n - 18;
    return cb;
}

// in-place inclusive scan of `n` doubles; the block offsets live in a small per-device scratch buffer
static int launch_scan(double *v, uint64_t n, cudaStream_t s) {
    static double *scratch[QB2_MAX_DEVICES] = {nullptr};
    int dev = 0;
    QB2_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= QB2_MAX_DEVICES) return set_error(QB2_ELIMIT, "device ordinal %d outside the engine's table", dev);
    const uint64_t nblocks = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    if (nblocks > (uint64_t)SCAN_MAX_BLOCKS) return set_error(QB2_ELIMIT, "scan of %llu entries", (unsigned long long)n);
    if (!scratch[dev]) QB2_CUDA(cudaMalloc(&scratch[dev], SCAN_MAX_BLOCKS * sizeof(double)));
    scan_totals_kernel<<<(unsigned)nblocks, SCAN_THREADS, 0, s>>>(v, n, scratch[dev]);
    scan_offsets_kernel<<<1, SCAN_THREADS, 0, s>>>(scratch[dev], (uint32_t)nblocks);
    scan_apply_kernel<<<(unsigned)nblocks, SCAN_THREADS, 0, s>>>(v, n, scratch[dev]);
    count_launch();
    count_launch();
    count_launch();
    return check_cuda(cudaGetLastError(), "scan launch");
}

int launch_sample_prepare(const void *state, int n, double *chunk_sums_dev, int dtype, cudaStream_t s) {
    if (!state || !chunk_sums_dev || n < 0 || n > 40) return set_error(QB2_EINVAL, "qb2_sample_prepare: bad arguments");
    const int cb = sample_chunk_bits(n);
    const uint64_t nchunks = 1ull << (n - cb);
    uint64_t blocks = nchunks;
    const uint64_t cap = (uint64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (dtype == QB2_C64)
        chunk_sums_kernel<float2><<<(unsigned)blocks, SAMPLE_THREADS, 0, s>>>((const float2 *)state, cb, nchunks, chunk_sums_dev);
    else
        chunk_sums_kernel<double2><<<(unsigned)blocks, SAMPLE_THREADS, 0, s>>>((const double2 *)state, cb, nchunks, chunk_sums_dev);
    count_launch();
    QB2_CUDA(cudaGetLastError());
    return launch_scan(chunk_sums_dev, nchunks, s);
}

int launch_chunk_sums(const void *state, int n, int cb, double *sums_dev, int dtype, cudaStream_t s) {
    if (!state || !sums_dev || cb < 0 || cb > n) return set_error(QB2_EINVAL, "chunk sums: bad arguments");
    const uint64_t nchunks = 1ull << (n - cb);
    uint64_t blocks = nchunks;
    const uint64_t cap = (uint64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (dtype == QB2_C64)
        chunk_sums_kernel<float2><<<(unsigned)blocks, SAMPLE_THREADS, 0, s>>>((const float2 *)state, cb, nchunks, sums_dev);
    else
        chunk_sums_kernel<double2><<<(unsigned)blocks, SAMPLE_THREADS, 0, s>>>((const double2 *)state, cb, nchunks, sums_dev);
    count_launch();
    return check_cuda(cudaGetLastError(), "chunk_sums launch");
}

int launch_scan_sums(double *sums_dev, uint64_t count, cudaStream_t s) {
    if (!sums_dev || count < 1) return set_error(QB2_EINVAL, "qb2_sample_scan: bad arguments");
    return launch_scan(sums_dev, count, s);
}

int launch_sample_draw(const void *state, int n, const double *chunk_sums_dev, const double *uniforms_dev,
                       int64_t shots, uint64_t *indices_dev, int dtype, cudaStream_t s) {
    return launch_sample_draw_cb(state, n, sample_chunk_bits(n), chunk_sums_dev, uniforms_dev, shots, indices_dev, dtype, s);
}

int launch_sample_draw_cb(const void *state, int n, int cb, const double *chunk_sums_dev, const double *uniforms_dev,
                          int64_t shots, uint64_t *indices_dev, int dtype, cudaStream_t s) {
    if (!state || !chunk_sums_dev || !uniforms_dev || !indices_dev || n < 0 || n > 40 || shots < 0 || cb < 0 || cb > n || cb > 20)
        return set_error(QB2_EINVAL, "qb2_sample_draw: bad arguments");
    if (shots == 0) return QB2_OK;
    const uint64_t nchunks = 1ull << (n - cb);
    uint64_t blocks = (uint64_t)shots;
    const uint64_t cap = (uint64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (dtype == QB2_C64)
        sample_draw_kernel<float2><<<(unsigned)blocks, SAMPLE_THREADS, 0, s>>>((const float2 *)state, cb, nchunks, chunk_sums_dev, uniforms_dev, shots, indices_dev);
    else
        sample_draw_kernel<double2><<<(unsigned)blocks, SAMPLE_THREADS, 0, s>>>((const double2 *)state, cb, nchunks, chunk_sums_dev, uniforms_dev, shots, indices_dev);
    count_launch();
    return check_cuda(cudaGetLastError(), "sample_draw launch");
}

// ---- outcome pruning on the device: the reference keeps, for more than 10 measured qubits, the outcomes
// with p > max(p) * cutoff_ratio (bi_array_helper.py:315-331) -- a max and a compare.  Three small kernels
// over the 2**m probabilities: max, count, compact.  Only the survivors travel to the host.
namespace {
__global__ void prob_max_kernel(const double *__restrict__ p, const uint64_t size, unsigned long long *__restrict__ out) {
    __shared__ double sh[32];
    double m = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += (uint64_t)gridDim.x * blockDim.x)
        m = fmax(m, p[i]);
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        // non-negative doubles order like their bit patterns
        if (threadIdx.x == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
    }
}
__global__ void prob_compact_kernel(const double *__restrict__ p, const uint64_t size, const unsigned long long *__restrict__ maxbits,
                                    const double ratio, unsigned long long *__restrict__ count, uint64_t *__restrict__ idx,
                                    double *__restrict__ val, const uint64_t capacity) {
    const double thr = __longlong_as_double((long long)*maxbits) * ratio;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += (uint64_t)gridDim.x * blockDim.x) {
        const double v = p[i];
        if (v > thr) {
            const unsigned long long at = atomicAdd(count, 1ull);
            if (at < capacity) {  // (capacity 0: the counting run)
                idx[at] = i;
                val[at] = v;
            }
        }
    }
}
}  // namespace

int launch_prune(const double *probs_dev, int m, double ratio, unsigned long long *scratch_dev, uint64_t *idx_dev, double *val_dev,
                 uint64_t capacity, cudaStream_t s) {
    if (!probs_dev || !scratch_dev || m < 0 || m > 34) return set_error(QB2_EINVAL, "qb2_prune_probabilities: bad arguments");
    if (capacity && (!idx_dev || !val_dev)) return set_error(QB2_EINVAL, "qb2_prune_probabilities: null output");
    const uint64_t size = 1ull << m;
    uint64_t blocks = (size + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (capacity == 0) {  // first call: maximum and survivor count (scratch[0] = max bits, scratch[1] = count)
        QB2_CUDA(cudaMemsetAsync(scratch_dev, 0, 16, s));
        prob_max_kernel<<<(unsigned)blocks, 256, 0, s>>>(probs_dev, size, scratch_dev);
        count_launch();
    } else {  // second call: compact (the maximum is still in scratch[0])
        QB2_CUDA(cudaMemsetAsync(scratch_dev + 1, 0, 8, s));
    }
    prob_compact_kernel<<<(unsigned)blocks, 256, 0, s>>>(probs_dev, size, scratch_dev, ratio, scratch_dev + 1, idx_dev, val_dev, capacity);
    count_launch();
    return check_cuda(cudaGetLastError(), "prune launch");
}

int launch_matrix_big(const void *src, void *dst, int n, uint64_t hi_base, const void *matrix_dev, const int *targets,
                      int k, uint64_t cmask, uint64_t cval, int dtype, cudaStream_t s) {
    if (!src || !dst || src == dst || !matrix_dev || !targets || n < 1 || n > 40)
        return set_error(QB2_EINVAL, "qb2_apply_matrix_big: bad arguments");
    if (k < 1 || k > 10 || k > n) return set_error(QB2_ELIMIT, "qb2_apply_matrix_big: k=%d outside 1..10", k);
    BigDesc bd;
    bd.k = k;
    uint64_t seen = 0;
    for (int t = 0; t < k; ++t) {
        if (targets[t] < 0 || targets[t] >= n || ((seen >> targets[t]) & 1))
            return set_error(QB2_EINVAL, "qb2_apply_matrix_big: bad target %d", targets[t]);
        seen |= 1ull << targets[t];
        bd.tgt[t] = (int8_t)targets[t];
    }
    if (cmask & seen) return set_error(QB2_EINVAL, "qb2_apply_matrix_big: control on a target");
    const uint64_t size = 1ull << n;
    const unsigned grid = grid_for(size, 256, 8);
    const size_t smem = sizeof(uint64_t) << k;
    if (dtype == QB2_C64)
        matrix_big_kernel<float><<<grid, 256, smem, s>>>((const float2 *)src, (float2 *)dst, size, hi_base, (const float2 *)matrix_dev, bd, cmask, cval);
    else
        matrix_big_kernel<double><<<grid, 256, smem, s>>>((const double2 *)src, (double2 *)dst, size, hi_base, (const double2 *)matrix_dev, bd, cmask, cval);
    count_launch();
    return check_cuda(cudaGetLastError(), "matrix_big launch");
}

}  // namespace qb2
