// qb2_prefix.cu -- sparse start of a simulation: HOST code, no device work (include/qrisp_b200.h
// qb2_sparse_prefix).
//
// Counterpart of the reference's sparse start: a fresh TensorFactor is a sparse BiArray and gates are
// contracted sparsely while the state has few non-zero amplitudes (tensor_factor.py:32-50, :86-89;
// bi_arrays.py:1197-1213).  Here the state is a plain list of (index, amplitude) pairs; a gate maps every
// entry to at most 2**k entries (one per non-zero of the matrix column it selects), accumulated in an
// open-addressing table.  The walk stops in front of the first gate that would outgrow `max_support`.
#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

#include "qb2_internal.cuh"

namespace qb2 {
namespace {

typedef std::complex<double> cplx;

struct Table {  // open addressing, linear probing, keys are state indices
    std::vector<uint64_t> key;
    std::vector<cplx> val;
    std::vector<uint8_t> used;
    std::vector<uint32_t> order;  // slots in insertion order
    uint32_t mask;
    explicit Table(uint32_t cap) {
        uint32_t sz = 16;
        while (sz < 4 * cap) sz <<= 1;
        key.resize(sz);
        val.resize(sz);
        used.assign(sz, 0);
        mask = sz - 1;
    }
    void clear() {
        for (uint32_t s : order) used[s] = 0;
        order.clear();
    }
    // returns false when a new key would exceed `limit` entries
    bool add(uint64_t k, cplx v, uint32_t limit) {
        uint32_t s = (uint32_t)((k * 0x9E3779B97F4A7C15ull) >> 40) & mask;
        while (used[s]) {
            if (key[s] == k) {
                val[s] += v;
                return true;
            }
            s = (s + 1) & mask;
        }
        if (order.size() >= limit) return false;
        used[s] = 1;
        key[s] = k;
        val[s] = v;
        order.push_back(s);
        return true;
    }
};

}  // namespace
}  // namespace qb2

using namespace qb2;

extern "C" int qb2_sparse_prefix(int n_instr, const int32_t *qoff, const int32_t *qidx, const int64_t *moff, const double *mat,
                                 uint32_t max_support, double drop_tol, uint64_t *idx, double *amp, uint32_t *count,
                                 int32_t *consumed) {
    if (n_instr < 0 || !qoff || !qidx || !moff || !mat || !idx || !amp || !count || !consumed)
        return set_error(QB2_EINVAL, "qb2_sparse_prefix: null pointer");
    if (max_support < 1 || *count > max_support) return set_error(QB2_EINVAL, "qb2_sparse_prefix: bad list size");
    std::vector<uint64_t> cur_i(idx, idx + *count);
    std::vector<cplx> cur_a(*count);
    for (uint32_t e = 0; e < *count; ++e) cur_a[e] = cplx(amp[2 * e], amp[2 * e + 1]);
    Table tab(max_support);
    // a gate may create entries that cancel: while it runs the table may hold up to 2x the limit
    const uint32_t hard = 2 * max_support;
    int done = 0;
    for (; done < n_instr; ++done) {
        const int k = qoff[done + 1] - qoff[done];
        if (k < 1 || k > 12) return set_error(QB2_ELIMIT, "qb2_sparse_prefix: instruction %d acts on %d qubits", done, k);
        const int32_t *q = qidx + qoff[done];
        const uint32_t D = 1u << k;
        const cplx *M = reinterpret_cast<const cplx *>(mat) + moff[done];
        uint64_t bit[12], qmask = 0;
        for (int i = 0; i < k; ++i) {
            if (q[i] < 0 || q[i] > 62) return set_error(QB2_EINVAL, "qb2_sparse_prefix: bad qubit %d", q[i]);
            bit[i] = 1ull << q[i];  // matrix index bit k-1-i <-> qubits[i]
            qmask |= bit[i];
        }
        // monomial matrix (one non-zero per column: diagonal phases, CX / X ladders, controlled phases -- most gates of
        // a state preparation or a QFT): every entry maps to exactly one entry, in place, no table
        {
            uint32_t rowof[4096];
            bool mono = D <= 64;
            for (uint32_t col = 0; col < D && mono; ++col) {
                int nz = 0;
                for (uint32_t row = 0; row < D; ++row) {
                    const cplx m = M[(size_t)row * D + col];
                    if (m.real() != 0.0 || m.imag() != 0.0) {
                        ++nz;
                        rowof[col] = row;
                    }
                }
                mono = nz == 1;
            }
            if (mono) {
                for (size_t e = 0; e < cur_i.size(); ++e) {
                    const uint64_t x = cur_i[e];
                    uint32_t col = 0;
                    for (int i = 0; i < k; ++i)
                        if (x & bit[i]) col |= 1u << (k - 1 - i);
                    const uint32_t row = rowof[col];
                    uint64_t y = x & ~qmask;
                    for (int i = 0; i < k; ++i)
                        if (row & (1u << (k - 1 - i))) y |= bit[i];
                    cur_i[e] = y;
                    cur_a[e] *= M[(size_t)row * D + col];
                }
                continue;
            }
        }
        tab.clear();
        bool fits = true;
        for (size_t e = 0; e < cur_i.size() && fits; ++e) {
            const uint64_t x = cur_i[e], base = x & ~qmask;
            uint32_t col = 0;
            for (int i = 0; i < k; ++i)
                if (x & bit[i]) col |= 1u << (k - 1 - i);
            for (uint32_t row = 0; row < D; ++row) {
                const cplx m = M[(size_t)row * D + col];
                if (m.real() == 0.0 && m.imag() == 0.0) continue;
                uint64_t y = base;
                for (int i = 0; i < k; ++i)
                    if (row & (1u << (k - 1 - i))) y |= bit[i];
                if (!tab.add(y, m * cur_a[e], hard)) {
                    fits = false;
                    break;
                }
            }
        }
        if (!fits) break;
        uint32_t alive = 0;
        for (uint32_t s : tab.order)
            if (std::abs(tab.val[s]) > drop_tol) ++alive;
        if (alive > max_support) break;
        cur_i.clear();
        cur_a.clear();
        for (uint32_t s : tab.order)
            if (std::abs(tab.val[s]) > drop_tol) {
                cur_i.push_back(tab.key[s]);
                cur_a.push_back(tab.val[s]);
            }
    }
    std::vector<uint32_t> ord(cur_i.size());
    for (uint32_t e = 0; e < ord.size(); ++e) ord[e] = e;
    std::sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return cur_i[a] < cur_i[b]; });
    for (uint32_t e = 0; e < ord.size(); ++e) {
        idx[e] = cur_i[ord[e]];
        amp[2 * e] = cur_a[ord[e]].real();
        amp[2 * e + 1] = cur_a[ord[e]].imag();
    }
    *count = (uint32_t)ord.size();
    *consumed = done;
    return QB2_OK;
}
