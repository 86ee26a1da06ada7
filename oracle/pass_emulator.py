"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A numpy interpreter of the packed pass format of ``include/qrisp_b200.h`` (qb2_pass_header,
qb2_phase, qb2_op, qb2_tab).  It executes a blob the way ``pass_kernel`` in
``qrisp_b200/csrc/qb2_pass.cu`` does -- same thread/register distribution, same
shared-memory slot maps, same table lookups -- so that the host planner
(``qrisp_b200/planner.py``) can be checked on a machine without a GPU, and so that the GPU
tests have a second, independent reading of the format.  Nothing under ``qrisp_b200/``
imports this module.

It also reports the bank-conflict degree of every exchange so that the planner's claim
("both sides conflict free") is tested, not assumed.
"""

from __future__ import annotations

import struct

import numpy as np

MAGIC = 0x32425151


class ParsedPass:
    pass


def parse(blob):
    b = bytes(blob)
    (magic, total, small, n, T, r, dtype, n_seg, n_phases, phases_off, ops_off, coef_off, u16_off, u64_off,
     tabdesc_off, tab_off, n_ops, _features, _pf, n_slots) = struct.unpack_from("<IIIIBBBBIIIIIIIIIIHH", b, 0)
    assert magic == MAGIC and total == len(b), (hex(magic), total, len(b))
    P = ParsedPass()
    P.n, P.T, P.r, P.dtype, P.small, P.n_slots = n, T, r, dtype, small, n_slots
    P.features = _features
    assert _pf == 0, "header.rsv16"
    P.segs = [struct.unpack_from("<BBBB", b, 64 + 4 * i)[:3] for i in range(n_seg)]
    P.phases = [struct.unpack_from("<8I", b, phases_off + 32 * i)[:7] for i in range(n_phases)]
    P.ops = []
    for i in range(n_ops):
        cmask, cval, jmask, data, code, kind, b0, b1, nthr, nreg, flags, tab = struct.unpack_from(
            "<QQIHHBBBBBBH", b, ops_off + 32 * i)
        P.ops.append(dict(cmask=cmask, cval=cval, jmask=jmask, data=data, code=code, kind=kind, b0=b0, b1=b1, nthr=nthr,
                          nreg=nreg, flags=flags, tab=tab))
        assert code == _expected_code(P.ops[-1], 1 << r) or 131 <= code < 148, (code, P.ops[-1])
    P.io = np.frombuffer(b, dtype=np.uint64, count=80, offset=128)
    # butterfly chains (QB2_CODE_CHAIN): the first record's code names (first bit, length); the
    # members are consecutive unnormalised butterflies on descending bits inside one phase, each with
    # v[] that depends on lower bits only, and a slot if it has a thread phase
    shapes = [(bs, k) for bs in range(1, 5) for k in range(2, bs + 2)]
    for i, op in enumerate(P.ops):
        if 131 <= op["code"] < 141:
            bs, k = shapes[op["code"] - 131]
            ph = next(p for p in P.phases if p[0] <= i < p[0] + p[1])
            assert i + k <= ph[0] + ph[1], "chain leaves its phase"
            for j, m in enumerate(P.ops[i : i + k]):
                assert m["kind"] == 9 and m["b0"] == bs - j and m["flags"] & 16 and m["flags"] & 2 and m["flags"] & 8, m
                assert m["b1"] or not m["flags"] & 4, m
                if j:
                    assert m["code"] == _expected_code(m, 1 << r)
    # fused DIAG_BIT runs (QB2_CODE_DBITS): `tab` of the first record is the mask of the members'
    # register bits; the members follow in descending bit order, each with flags & 1 and a slot
    for i, op in enumerate(P.ops):
        if op["code"] == 141:
            bits = [bb for bb in range(r - 1, -1, -1) if (op["tab"] >> bb) & 1]
            ph = next(p for p in P.phases if p[0] <= i < p[0] + p[1])
            assert len(bits) >= 2 and i + len(bits) <= ph[0] + ph[1]
            for bb, m in zip(bits, P.ops[i : i + len(bits)]):
                assert m["kind"] == 7 and m["b0"] == bb and m["flags"] & 1 and m["b1"] and m["nthr"] == 0, m
    # CX ladders (QB2_CODE_CXL): members are X on bit top-k-1 controlled by register bit top-k, no
    # controls outside the registers
    cxl = [(bs, k) for bs in range(2, 5) for k in range(2, bs + 1)]
    for i, op in enumerate(P.ops):
        if 142 <= op["code"] < 148:
            top, k = cxl[op["code"] - 142]
            ph = next(p for p in P.phases if p[0] <= i < p[0] + p[1])
            assert i + k <= ph[0] + ph[1]
            for j, m in enumerate(P.ops[i : i + k]):
                assert m["kind"] == 1 and m["flags"] == 3 and m["cmask"] == 0 and m["b1"] - 1 == top - j and m["b0"] == top - j - 1, m
    # phase slots: every slotted op owns its own slot(s) inside [0, n_slots)
    used = []
    for op in P.ops:
        if op["kind"] in (6, 7, 9) and op["b1"]:
            need = 2 if op["kind"] == 7 and not op["flags"] & 1 else 1
            used.extend(range(op["b1"] - 1, op["b1"] - 1 + need))
    assert sorted(used) == list(range(n_slots)), (used, n_slots)
    assert n_slots <= 32
    real_t = np.float32 if dtype == 0 else np.float64
    frac_t = np.uint32 if dtype == 0 else np.uint64
    P.real_t, P.frac_t = real_t, frac_t
    P.coefs = np.frombuffer(b, dtype=real_t, count=(u64_off - coef_off) // np.dtype(real_t).itemsize, offset=coef_off)
    P.a64 = np.frombuffer(b, dtype=np.uint64, count=(tabdesc_off - u64_off) // 8, offset=u64_off)
    ntd = (u16_off - tabdesc_off) // 32
    P.tabs = []
    for i in range(ntd):
        raw = struct.unpack_from("<24BII", b, tabdesc_off + 32 * i)
        if raw[7]:  # qb2_mul
            _z, shift, brev, rsh, is_mul, mask, _r, mult, _r3 = struct.unpack_from("<3xBBBBBIIQQ", b, tabdesc_off + 32 * i)
            P.tabs.append(dict(mul=(shift, mask, mult, brev, rsh)))
            continue
        pieces = [(raw[4 * k], raw[4 * k + 1], raw[4 * k + 2]) for k in range(raw[3])]
        P.tabs.append(dict(pieces=pieces, off=raw[24], regoff=raw[25]))
    P.a16 = np.frombuffer(b, dtype=np.uint16, count=(small - u16_off) // 2, offset=u16_off)
    P.tables = np.frombuffer(b, dtype=frac_t, count=(total - tab_off) // np.dtype(frac_t).itemsize, offset=tab_off)
    return P


def _expected_code(op, NR):
    """qb2_op.code as the header defines it (QB2_CODE_*), recomputed from the other fields."""
    kind, b0, b1, fl = op["kind"], op["b0"], op["b1"], op["flags"]
    if kind == 9:
        mode = 0 if not fl & 2 else (1 if not fl & 4 else (3 if fl & 8 else 2))
        return 8 * b0 + 2 * mode + (1 if fl & 16 else 0)
    if kind in (6, 8, 5):
        return {6: 40, 8: 41, 5: 42}[kind]
    if kind == 7:
        return 44 + 2 * b0 + (fl & 1)
    if kind == 1:
        if fl == 3 and b1:
            return 94 + 5 * b0 + (b1 - 1)
        return 54 + 8 * b0 + 2 * fl + (1 if op["jmask"] == (1 << NR) - 1 else 0)
    if kind == 2:
        pairs = [(a, c) for a in range(5) for c in range(a + 1, 5)]
        return 119 + pairs.index((b0, b1))
    return 129 if kind == 3 else 130


def _tab_index(tab, X):
    idx = np.zeros(X.shape, dtype=np.uint64)
    for shift, ln, out in tab["pieces"]:
        if ln:
            idx |= ((X >> np.uint64(shift)) & np.uint64((1 << ln) - 1)) << np.uint64(out)
    return idx


def _mul_phase(tdesc, X, frac_t):
    shift, mask, mult, brev, rsh = tdesc["mul"]
    v = ((X >> np.uint64(shift)) & np.uint64(mask)).astype(np.uint64)
    if brev:
        rv = np.zeros_like(v)
        for bit in range(32):
            rv |= ((v >> np.uint64(bit)) & np.uint64(1)) << np.uint64(31 - bit)
        v = rv >> np.uint64(rsh)
    bits = 32 if frac_t == np.uint32 else 64
    with np.errstate(over="ignore"):
        prod = v * np.uint64(mult % (1 << 64))  # wraps modulo 2**64
    return (prod & np.uint64((1 << bits) - 1)).astype(frac_t)


def _phase(turns, dtype):
    """exp(2 pi i turns / 2**bits) in the pass precision (the kernel's own approximation is
    good to ~1e-7; the emulator uses the exact value)."""
    if dtype == 0:
        ang = turns.astype(np.float64) * (2 * np.pi / 4294967296.0)
    else:
        ang = turns.astype(np.int64).astype(np.float64) * (2 * np.pi / 18446744073709551616.0)
    return np.exp(1j * ang)


def tma_slot(t):
    """Slot (amplitudes) of tile index ``t`` in a buffer written by a bulk tensor copy with the 128-byte
    swizzle (CU_TENSOR_MAP_SWIZZLE_128B): 128-byte rows, 16-byte chunk c of row r sits at chunk c ^ (r & 7)."""
    t = np.asarray(t, dtype=np.int64)
    row, chunk, half = t >> 4, (t >> 1) & 7, t & 1
    return row * 16 + ((chunk ^ (row & 7)) << 1) + half


def check_tma_maps(P, report=None):
    """QB2_FEATURE_TMA: phase 0's xr / xw maps must send every (thread, register) of the first / last
    phase to the slot the tensor map's box order + swizzle gives its tile index; the tile must be
    expressible as a rank-5 box (positions 0..3, at most three runs of at most 8 higher positions)."""
    T, r = P.T, P.r
    NR, TB = 1 << r, T - r
    assert P.dtype == 0 and T == 13 and r == 5
    nontile = set()
    for src, dst, ln in P.segs:
        nontile.update(range(dst, dst + ln))
    tile_pos = [x for x in range(P.n) if x not in nontile]
    assert len(tile_pos) == T and tile_pos[:4] == [0, 1, 2, 3]
    dims, i, high = 0, 0, tile_pos[4:]
    while i < len(high):
        j = i
        while j + 1 < len(high) and high[j + 1] == high[j] + 1 and j + 1 - i < 8:
            j += 1
        dims += 1
        i = j + 1
    assert dims <= 3, "tile needs more than three box dimensions"
    rank = {x: i for i, x in enumerate(tile_pos)}
    tid = np.arange(1 << TB, dtype=np.int64)
    first_op, nops, pcol0, rphys0, xw, xr, xflags = P.phases[0]
    for side, ph, m, flag, sh in (("load", P.phases[0], xr, 2, 16), ("store", P.phases[-1], xw, 1, 8)):
        pcol, rphys = ph[2], ph[3]
        tix = np.zeros((1 << TB, NR), dtype=np.int64)  # tile index of (thread, register)
        for b in range(TB):
            x = int(P.a64[pcol + b]).bit_length() - 1
            tix += (((tid >> b) & 1) << rank[x])[:, None]
        for j in range(NR):
            v = int(P.a64[rphys + j])
            tix[:, j] += sum(1 << rank[x] for x in range(64) if (v >> x) & 1)
        want = tma_slot(tix)
        tb = np.zeros(1 << TB, dtype=np.int64)
        for b in range(TB):
            tb[((tid >> b) & 1).astype(bool)] ^= int(P.a16[m + b])
        rreg = P.a16[m + TB : m + TB + NR].astype(np.int64)
        got = tb[:, None] ^ rreg[None, :]
        assert np.array_equal(got, want), f"TMA {side} map does not match the swizzled box layout"
        if xflags & flag:
            k = (xflags >> sh) & 0xFF
            assert np.array_equal(rreg, np.arange(NR) << k) and not np.any(tb & ((NR - 1) << k))
            assert k in (7, 8)
        _check_conflicts(got, 16, report, "tma-" + side, 0)


def run(blob, state, hi_base=0, report=None, sparse=None):
    """Apply the pass to ``state`` (complex numpy array of 2**n, modified copy returned).  ``sparse``:
    the (tile, loc, amp) list of QB2_FEATURE_ZERO_INPUT -- the state is then not read."""
    P = parse(blob)
    if P.features & 2:
        check_tma_maps(P, report)
    n, T, r = P.n, P.T, P.r
    NR, TB = 1 << r, T - r
    nthreads = 1 << TB
    cdt = np.complex64 if P.dtype == 0 else np.complex128
    state = np.array(state, dtype=cdt)
    assert state.size == 1 << n
    tid = np.arange(nthreads, dtype=np.uint64)
    ntiles = 1 << (n - T)
    conflict_lanes = 16 if P.dtype == 0 else 8

    def thread_cols(base_idx, arr, xor):
        acc = np.zeros(nthreads, dtype=np.uint64)
        for b in range(TB):
            v = np.uint64(int(arr[base_idx + b]))
            sel = ((tid >> np.uint64(b)) & np.uint64(1)).astype(bool)
            acc[sel] = (acc[sel] ^ v) if xor else (acc[sel] | v)
        return acc

    for t in range(ntiles):
        tbase = 0
        for src, dst, ln in P.segs:
            tbase |= ((t >> src) & ((1 << ln) - 1)) << dst
        ph0 = P.phases[0]
        assert len(P.phases) <= 8
        xloc = np.uint64(tbase) | thread_cols(ph0[2], P.a64, False)
        rp = P.a64[ph0[3] : ph0[3] + NR]
        csize = 8 if P.dtype == 0 else 16
        assert np.array_equal(P.io[:NR], rp * np.uint64(csize)), "qb2_io.load_off"
        addr = xloc[:, None] | (P.io[:NR] // np.uint64(csize))[None, :]
        load_addr = addr.astype(np.int64)
        regs = state[addr.astype(np.int64)]  # [thread, j]
        if sparse is not None:
            sp_tile, sp_loc, sp_amp = sparse
            assert np.all(np.diff(sp_tile.astype(np.int64)) >= 0), "sparse list must be sorted by tile"
            regs = np.zeros_like(regs)
            hit = np.flatnonzero(sp_tile == t)
            for e in hit:
                regs[int(sp_loc[e]) >> r, int(sp_loc[e]) & (NR - 1)] = sp_amp[e]
            if len(hit) == 0:
                state[addr.astype(np.int64)] = 0  # the kernel stores zeros without running the ops
                continue
        seen = np.zeros(1 << T, dtype=bool)
        for pi, (first_op, nops, pcol, rphys, xw, xr, xflags) in enumerate(P.phases):
            if pi > 0:
                tile = np.zeros(1 << T, dtype=cdt)
                filled = np.zeros(1 << T, dtype=bool)
                wb = thread_cols(xw, P.a16, True)
                wreg = P.a16[xw + TB : xw + TB + NR].astype(np.uint64)
                if xflags & 1:  # the kernel uses j << kw (added, not XORed) instead of the table
                    kw = (xflags >> 8) & 0xFF
                    assert np.array_equal(wreg, np.arange(NR, dtype=np.uint64) << np.uint64(kw))
                    assert not np.any(wb & np.uint64((NR - 1) << kw))
                slots = (wb[:, None] ^ wreg[None, :]).astype(np.int64)
                assert slots.max() < (1 << T)
                assert not filled[slots.reshape(-1)].any() and len(np.unique(slots)) == slots.size, "store slots collide"
                tile[slots] = regs
                _check_conflicts(slots, conflict_lanes, report, "store", pi)
                rb = thread_cols(xr, P.a16, True)
                rreg = P.a16[xr + TB : xr + TB + NR].astype(np.uint64)
                if xflags & 2:
                    kr = (xflags >> 16) & 0xFF
                    assert np.array_equal(rreg, np.arange(NR, dtype=np.uint64) << np.uint64(kr))
                    assert not np.any(rb & np.uint64((NR - 1) << kr))
                slots = (rb[:, None] ^ rreg[None, :]).astype(np.int64)
                assert len(np.unique(slots)) == slots.size, "load slots collide"
                regs = tile[slots]
                _check_conflicts(slots, conflict_lanes, report, "load", pi)
                xloc = np.uint64(tbase) | thread_cols(pcol, P.a64, False)
            X = np.uint64(hi_base) | xloc
            for o in range(first_op, first_op + nops):
                op = P.ops[o]
                if op["kind"] in (6, 7, 9) and op["b1"]:
                    # phase slot: the kernel evaluates the op's ladders separately on the thread's
                    # part of the index (once per CTA) and on the tile's part (once per tile) and adds
                    # the two -- check that this is what the ladders evaluate to on the whole index
                    assert op["b1"] <= P.n_slots
                    first = op["tab"] if op["kind"] == 9 else op["data"]
                    descs = P.tabs[first : first + op["nthr"] + op["nreg"]]
                    assert descs and all("mul" in d for d in descs), "a slot needs multiplier ladders only"
                    x_thr = np.uint64(hi_base) | (xloc & ~np.uint64(tbase))
                    x_tile = np.full(nthreads, np.uint64(tbase), dtype=np.uint64)
                    for d in descs:
                        whole = _mul_phase(d, X, P.frac_t)
                        with np.errstate(over="ignore"):
                            split = (_mul_phase(d, x_thr, P.frac_t) + _mul_phase(d, x_tile, P.frac_t)).astype(P.frac_t)
                        assert np.array_equal(whole, split), "ladder is not additive over thread / tile bits"
                if op["kind"] == 9:  # BFLY
                    b0, fl = op["b0"], op["flags"]
                    cf = P.coefs[2 * op["data"] : 2 * op["data"] + 8 + NR].astype(np.float64)
                    m = (cf[0:8:2] + 1j * cf[1:8:2]).reshape(2, 2)
                    if fl & 1:
                        m = m.real.astype(np.complex128)
                    if fl & 16:  # QB2_BFLY_HAD: the kernel adds and subtracts, the coefficients are not read
                        m = np.array([[1, 1], [1, -1]], dtype=np.complex128)
                    v = cf[8::2] + 1j * cf[9::2]
                    s1 = np.zeros(nthreads, dtype=P.frac_t)
                    if fl & 4:
                        for tdesc in P.tabs[op["tab"] : op["tab"] + op["nreg"]]:
                            if "mul" in tdesc:
                                s1 = s1 + _mul_phase(tdesc, X, P.frac_t)
                            else:
                                e = (np.uint64(tdesc["off"]) + _tab_index(tdesc, X)).astype(np.int64)
                                s1 = s1 + P.tables[e + tdesc["regoff"]]
                    w1 = _phase(s1, P.dtype) if fl & 4 else np.ones(nthreads, dtype=np.complex128)
                    new = regs.copy()
                    for g in range(NR // 2):
                        j0 = ((g >> b0) << (b0 + 1)) | (g & ((1 << b0) - 1))
                        j1 = j0 | (1 << b0)
                        x0, x1 = regs[:, j0].astype(np.complex128), regs[:, j1].astype(np.complex128)
                        # QB2_BFLY_VLOW (8): the kernel reads v[] for the low b0 bits of the pair only
                        gv = g & ((1 << b0) - 1) if fl & 8 else g
                        c = w1 * (v[gv] if fl & 2 else 1.0)
                        new[:, j0] = (m[0, 0] * x0 + m[0, 1] * x1).astype(cdt)
                        new[:, j1] = ((m[1, 0] * x0 + m[1, 1] * x1) * c).astype(cdt)
                    regs = new
                elif op["kind"] >= 5:
                    kind = op["kind"]
                    tb = P.tabs[op["data"] : op["data"] + op["nthr"] + op["nreg"]] if kind != 8 else []
                    s = np.zeros(nthreads, dtype=P.frac_t)
                    for tdesc in tb[: op["nthr"]]:
                        if "mul" in tdesc:
                            s = s + _mul_phase(tdesc, X, P.frac_t)
                        else:
                            s = s + P.tables[(np.uint64(tdesc["off"]) + _tab_index(tdesc, X)).astype(np.int64)]
                    if kind == 6:  # DIAG_THR
                        regs = (regs * _phase(s, P.dtype)[:, None]).astype(cdt)
                    elif kind == 7:  # DIAG_BIT
                        s0, s1 = s.copy(), s.copy()
                        for tdesc in tb[op["nthr"] :]:
                            if "mul" in tdesc:
                                s1 = s1 + _mul_phase(tdesc, X, P.frac_t)
                                continue
                            e = (np.uint64(tdesc["off"]) + _tab_index(tdesc, X)).astype(np.int64)
                            s0 = s0 + P.tables[e]
                            s1 = s1 + P.tables[e + tdesc["regoff"]]
                        w0, w1 = _phase(s0, P.dtype), _phase(s1, P.dtype)
                        new = regs.copy()
                        for j in range(NR):
                            if (j >> op["b0"]) & 1:
                                new[:, j] = (regs[:, j] * w1).astype(cdt)
                            elif not (op["flags"] & 1):
                                new[:, j] = (regs[:, j] * w0).astype(cdt)
                        regs = new
                    elif kind == 8:  # DIAG_VEC
                        cv = P.coefs[2 * op["data"] : 2 * op["data"] + 2 * NR].astype(np.float64)
                        regs = (regs * (cv[0::2] + 1j * cv[1::2])[None, :]).astype(cdt)
                    else:
                        sj = np.repeat(s[:, None], NR, axis=1)
                        for tdesc in tb[op["nthr"] :]:
                            base = np.uint64(tdesc["off"]) + _tab_index(tdesc, X)
                            go = P.a16[tdesc["regoff"] : tdesc["regoff"] + NR].astype(np.uint64)
                            sj = sj + P.tables[(base[:, None] + go[None, :]).astype(np.int64)]
                        regs = (regs * _phase(sj, P.dtype)).astype(cdt)
                else:
                    k = op["kind"]
                    D = 1 << k
                    bits = [op["b0"], op["b1"]][:k] if k <= 2 else list(range(k))
                    m = P.coefs[2 * op["data"] : 2 * op["data"] + 2 * D * D].astype(np.float64)
                    m = (m[0::2] + 1j * m[1::2]).reshape(D, D)
                    if k == 1:  # the specialised 2x2 paths read only part of the coefficients
                        if op["flags"] == 1:
                            m = m.real.astype(np.complex128)
                        elif op["flags"] == 2:
                            m = np.array([[0, m[0, 1]], [m[1, 0], 0]])
                        elif op["flags"] == 3:
                            m = np.array([[0, 1], [1, 0]], dtype=np.complex128)
                    on = ((X ^ np.uint64(op["cval"])) & np.uint64(op["cmask"])) == 0
                    tmask = sum(1 << bb for bb in bits)
                    jmask = op["jmask"]
                    if k == 1 and op["flags"] == 3 and op["b1"]:
                        # CX inside the registers: the kernel ignores jmask and swaps the pairs whose
                        # register bit b1-1 is set
                        cb = op["b1"] - 1
                        assert cb != op["b0"] and cb < r
                        jmask = sum(1 << j for j in range(NR) if (j >> cb) & 1)
                    new = regs.copy()
                    for base in range(NR):
                        if base & tmask or not (jmask >> base) & 1:
                            continue
                        js = [base | sum(((c >> i) & 1) << bits[i] for i in range(k)) for c in range(D)]
                        x = regs[:, js].astype(np.complex128)
                        y = x @ m.T
                        new[np.ix_(on, js)] = y[on].astype(cdt)
                    regs = new
        # store: qb2_io.store_col / store_off (the tile may be stored relabelled; then the columns
        # differ from the last phase's own, but they must still cover the tile exactly once)
        sloc = np.full(nthreads, np.uint64(tbase), dtype=np.uint64)
        for bb in range(TB):
            sel = ((tid >> np.uint64(bb)) & np.uint64(1)).astype(bool)
            sloc[sel] |= P.io[64 + bb]
        addr = (sloc[:, None] | (P.io[32 : 32 + NR] // np.uint64(csize))[None, :]).astype(np.int64)
        assert len(np.unique(addr)) == addr.size, "store addresses collide"
        assert np.array_equal(np.sort(addr.reshape(-1)), np.sort(load_addr.reshape(-1))), "store leaves the tile"
        state[addr] = regs
    return state


def _check_conflicts(slots, lanes, report, side, phase):
    """Worst bank-conflict degree over wavefronts of ``lanes`` consecutive threads."""
    if report is None:
        return
    nthreads = slots.shape[0]
    lanes = min(lanes, nthreads)
    worst = 1
    for j in range(slots.shape[1]):
        g = slots[:, j].reshape(-1, lanes) % lanes  # bank group of each lane
        for row in g:
            worst = max(worst, int(np.bincount(row, minlength=lanes).max()))
    report.append((phase, side, worst))
