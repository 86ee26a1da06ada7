"""Pass planner: normalised ops -> packed passes for ``qb2_run_pass`` (include/qrisp_b200.h).

This is the B200 replacement of the reference's circuit preprocessing for simulation
(reference: circuit_preprocessing.py:119-157 ``group_qc`` and :840-852
``circuit_preprocessor``).  The reference groups gates so that fewer, larger dense matrices
are multiplied into the state; here gates are grouped so that the state makes as few trips
through HBM as possible:

* a PASS is a set of ops whose dense targets fit into one tile of ``T`` physical positions
  (the lowest ``L`` positions are always part of the tile so that every global access is a
  full 128-byte line);
* inside a pass, a PHASE is a set of ops whose dense targets fit into the ``r`` register
  positions of the phase; between phases the tile is re-distributed through shared memory;
* diagonal ops and controls never constrain anything: they are evaluated from the physical
  index, wherever the qubit lives (tile, other tile, other rank).

The planner also owns the qubit-permutation schedule: ``pos[q]`` is the physical position of
logical qubit ``q``; an exact SWAP only edits that map.

Everything here is host logic on small arrays; nothing touches amplitudes.
"""

from __future__ import annotations

import os
import math
import struct

import numpy as np

QB2_PASS_MAGIC = 0x32425151
QB2_C64, QB2_C128 = 0, 1
OP_MAT = {1: 1, 2: 2, 3: 3, 4: 4}
OP_DIAG, OP_DIAG_THR, OP_DIAG_BIT, OP_DIAG_VEC = 5, 6, 7, 8
MAT1_GENERAL, MAT1_REAL, MAT1_ANTI, MAT1_SWAP = 0, 1, 2, 3
OP_BFLY = 9
BFLY_REAL, BFLY_V, BFLY_W, BFLY_VLOW, BFLY_HAD = 1, 2, 4, 8, 16
MAX_SEGS = 16
MAX_PHASES = 8
MAX_SMALL_BYTES = 8192
MAX_SLOTS = 32
SMEM_PER_CTA = 113 * 1024  # two CTAs per SM
# the TMA kernel (qb2_pass_tma.cuh): one CTA per SM, three 64 KB tile buffers; what is left of 227 KB
TMA_AUX_BYTES = 232448 - 3 * 65536 - 1024
MAX_SPARSE = 1024  # QB2_MAX_SPARSE: entries of a sparse input list
FEATURE_FULL, FEATURE_TMA, FEATURE_ZERO_INPUT = 1, 2, 4
MAX_TAB_BITS = 10
MAX_PIECES = 6
MAX_REG_TABS = 4
HEADER_BYTES = 128
IO_BYTES = 640  # qb2_io: register offsets at the load and at the store, thread columns of the store


# dispatch codes of the pass kernel (include/qrisp_b200.h QB2_CODE_*)
CODE_BFLY, CODE_DIAG_THR, CODE_DIAG_VEC, CODE_DIAG_GEN, CODE_DIAG_BIT = 0, 40, 41, 42, 44
CODE_MAT1, CODE_SWAPC, CODE_MAT2, CODE_MAT3, CODE_MAT4 = 54, 94, 119, 129, 130
CODE_CHAIN = 131  # + index of (first bit, length) in CHAIN_SHAPES
CHAIN_SHAPES = [(bs, n) for bs in range(1, 5) for n in range(2, bs + 2)]


def op_code(kind, b0, b1, flags, jmask, NR):
    """The single number the kernel switches on: op kind x register bit x variant."""
    if kind == OP_BFLY:
        mode = 0 if not flags & BFLY_V else (1 if not flags & BFLY_W else (3 if flags & BFLY_VLOW else 2))
        return CODE_BFLY + 8 * b0 + 2 * mode + (1 if flags & BFLY_HAD else 0)
    if kind == OP_DIAG_THR:
        return CODE_DIAG_THR
    if kind == OP_DIAG_VEC:
        return CODE_DIAG_VEC
    if kind == OP_DIAG:
        return CODE_DIAG_GEN
    if kind == OP_DIAG_BIT:
        return CODE_DIAG_BIT + 2 * b0 + (flags & 1)
    if kind == 1:
        if flags == MAT1_SWAP and b1:
            return CODE_SWAPC + 5 * b0 + (b1 - 1)
        allj = 1 if jmask == (1 << NR) - 1 else 0
        return CODE_MAT1 + 8 * b0 + 2 * flags + allj
    if kind == 2:
        pairs = [(a, b) for a in range(5) for b in range(a + 1, 5)]
        return CODE_MAT2 + pairs.index((b0, b1))
    return CODE_MAT3 if kind == 3 else CODE_MAT4


def pack_op(cmask, cval, jmask, data, kind, b0, b1, nthr, nreg, flags, tab, NR, code=None):
    assert data < 65536
    if code is None:
        code = op_code(kind, b0, b1, flags, jmask, NR)
    return struct.pack("<QQIHHBBBBBBH", cmask, cval, jmask, data, code, kind, b0, b1, nthr, nreg, flags, tab)


CODE_DBITS = 141  # run of DIAG_BIT ops, one per register bit (mask of the bits in qb2_op.tab)


def fuse_diag_bit_runs(recs, first):
    """Consecutive DIAG_BIT ops of a phase whose b0=0 phase vanishes, each on its own register bit
    and with a phase slot, become one kernel op (QB2_CODE_DBITS): one decode, members in descending
    bit order (diagonal ops commute), straight-line code."""
    i = first
    while i < len(recs):
        n = 0
        bits = set()
        while i + n < len(recs):
            m = recs[i + n]
            if not (m["kind"] == OP_DIAG_BIT and m["flags"] & 1 and m["b1"] != 0 and m["nthr"] == 0 and m["b0"] not in bits):
                break
            bits.add(m["b0"])
            n += 1
        if n >= 2:
            recs[i : i + n] = sorted(recs[i : i + n], key=lambda m: -m["b0"])
            recs[i]["code"] = CODE_DBITS
            recs[i]["tab"] = sum(1 << b for b in bits)
            i += n
        else:
            i += max(n, 1)


CODE_CXL = 142  # + index of (top control bit, length) in CXL_SHAPES
CXL_SHAPES = [(bs, n) for bs in range(2, 5) for n in range(2, bs + 1)]


def fuse_cx_ladders(recs, first):
    """X on register bit b-1 controlled by register bit b, then b-2 by b-1, ...: the shape a chain of
    CX gates takes inside the registers (GHZ preparation, ripple carries).  One kernel op
    (QB2_CODE_CXL) for the run."""
    i = first
    while i < len(recs):
        r = recs[i]
        n = 0
        if r["kind"] == 1 and r["flags"] == MAT1_SWAP and r["b1"]:
            top = r["b1"] - 1
            while i + n < len(recs):
                m = recs[i + n]
                if not (m["kind"] == 1 and m["flags"] == MAT1_SWAP and m["cmask"] == 0 and m["b1"] - 1 == top - n
                        and m["b0"] == top - n - 1):
                    break
                n += 1
            while n >= 2 and (top, n) not in CXL_SHAPES:
                n -= 1
        if n >= 2:
            r["code"] = CODE_CXL + CXL_SHAPES.index((top, n))
            i += n
        else:
            i += 1


def fuse_butterfly_chains(recs, first, NR):
    """``recs[first:]`` are the argument tuples of one phase's ops.  Runs of unnormalised butterflies
    on register bits b, b-1, b-2... become one kernel op (QB2_CODE_CHAIN): the kernel decodes once
    and runs the members back to back in straight-line code, so the latency of fetching one
    member's phase hides behind the arithmetic of the one before.  Members keep their own
    records; a member without v[] gets the (all ones) v[] it carries anyway."""
    need = BFLY_HAD
    i = first
    while i < len(recs):
        r = recs[i]
        n = 0
        if r["kind"] == OP_BFLY:
            while i + n < len(recs):
                m = recs[i + n]
                ok = (m["kind"] == OP_BFLY and m["flags"] & need == need and m["b0"] == r["b0"] - n
                      and (m["flags"] & BFLY_VLOW or not m["flags"] & BFLY_V)  # v[] must depend on lower bits only
                      and (m["b1"] != 0 or not m["flags"] & BFLY_W))            # a thread phase needs its slot
                if not ok:
                    break
                n += 1
        if n >= 2 and (r["b0"], n) in CHAIN_SHAPES:
            for m in recs[i : i + n]:
                m["flags"] |= BFLY_V | BFLY_VLOW
            r["code"] = CODE_CHAIN + CHAIN_SHAPES.index((r["b0"], n))
            i += n
        else:
            i += 1



class PlanConfig:
    """Tile geometry.  ``T`` tile bits, ``r`` register bits, ``L`` forced low positions."""

    def __init__(self, n, dtype=QB2_C64, T=None, r=None, L=4, tma=None):
        self.dtype = dtype
        # tuning knobs (bench sweeps): QB2_REG_BITS / QB2_TILE_BITS override the defaults
        if r is None and os.environ.get("QB2_REG_BITS"):
            r = int(os.environ["QB2_REG_BITS"])
        if T is None and os.environ.get("QB2_TILE_BITS"):
            T = int(os.environ["QB2_TILE_BITS"])
        if r is None:
            # complex64: 32 amplitudes per thread (64 data registers, 2 CTAs of 256 threads per
            # SM) measured faster than 16 on every workload
            r = 5 if dtype == QB2_C64 else 3
        if T is None:
            T = r + 8
        T = min(T, n)
        r = min(r, T)
        self.n, self.T, self.r = n, T, r
        self.L = min(L, T - r) if T - r >= 1 else 0
        # lanes that one shared-memory wavefront covers: 16 for 8-byte, 8 for 16-byte amplitudes
        self.conflict_bits = 4 if dtype == QB2_C64 else 3
        self.max_mat = min(self.r, 4)  # widest dense op the pass kernel applies in registers
        self.fuse_bfly = os.environ.get("QB2_NO_BFLY") is None
        self.store_relabel = os.environ.get("QB2_NO_RELABEL") is None
        self.phase_slots = os.environ.get("QB2_NO_SLOTS") is None
        self.had_bfly = os.environ.get("QB2_NO_HAD") is None
        self.fuse_chains = os.environ.get("QB2_NO_CHAIN") is None
        # passes whose tile is one box of a tensor map run on the TMA kernel (complex64, T=13, r=5)
        want_tma = os.environ.get("QB2_NO_TMA") is None if tma is None else bool(tma)
        self.tma = dtype == QB2_C64 and self.T == 13 and self.r == 5 and self.L == 4 and want_tma

    @property
    def TB(self):
        return self.T - self.r


class BigMatStep:
    """A dense unitary too wide for a pass: runs through ``qb2_apply_matrix_big``."""

    kind = "big"

    def __init__(self, positions, matrix, cmask, cval):
        self.positions, self.matrix, self.cmask, self.cval = positions, matrix, cmask, cval


class DescriptorOverflow(RuntimeError):
    pass


class PassStep:
    kind = "pass"

    def __init__(self, blob, n_ops, n_phases, tile_pos, n_gate_ops, segs=(), R0=(), TP0=(), tma=False):
        self.blob, self.n_ops, self.n_phases, self.tile_pos = blob, n_ops, n_phases, tile_pos
        self.n_gate_ops = n_gate_ops
        # what a sparse input list needs (engine.sparse_input_list): tile-number segments and the first
        # phase's distribution (register positions, thread positions)
        self.segs, self.R0, self.TP0 = list(segs), list(R0), list(TP0)
        self.tma = tma
        self.last = False


def tma_tile_ok(tile_pos, n):
    """Can the tile be ONE box of a rank-5 tensor map?  Positions 0..3 are dim 0 (a 128-byte line), one
    dimension of box size 1 carries the non-tile bits, which leaves three dimensions for runs of higher tile
    positions of at most 8 bits each (qb2_pass_tma.cuh make_tile_map)."""
    if list(tile_pos[:4]) != [0, 1, 2, 3]:
        return False
    high = tile_pos[4:]
    dims = 0
    i = 0
    while i < len(high):
        j = i
        while j + 1 < len(high) and high[j + 1] == high[j] + 1 and j + 1 - i < 8:
            j += 1
        dims += 1
        i = j + 1
    return dims <= 3


def tma_columns(T):
    """Slot (in amplitudes) of tile-index bit i in a tile buffer the TMA fills with the 128-byte swizzle:
    row = index >> 4, the 16-byte chunk (index >> 1) & 7 is stored at chunk ^ (row & 7)."""
    return [(1 << i) | ((1 << (i - 3)) if 4 <= i <= 6 else 0) for i in range(T)]


# ----------------------------------------------------------------------------------------
# scheduling
# ----------------------------------------------------------------------------------------


def _commuting_pick(ops, accept):
    """Generic list scheduler: walk ``ops`` in order, take every op that ``accept`` takes and
    that commutes with everything skipped so far; return (taken, rest)."""
    taken, rest = [], []
    blocked_nd, blocked_d = set(), set()
    for op in ops:
        ok = True
        for q in op.nd:
            if q in blocked_nd or q in blocked_d:
                ok = False
                break
        if ok:
            for q in op.d:
                if q in blocked_nd:
                    ok = False
                    break
        if ok and accept(op):
            taken.append(op)
        else:
            rest.append(op)
            blocked_nd.update(op.nd)
            blocked_d.update(op.d)
    return taken, rest


class Planner:
    """Turns a list of normalised ops into steps.  ``pos`` is updated in place."""

    def __init__(self, cfg: PlanConfig, pos, n_total=None, window=4096, max_pass_ops=256):
        self.max_pass_ops = max_pass_ops
        self.cfg = cfg
        self.pos = pos  # list: logical qubit -> physical position
        self.n_total = cfg.n if n_total is None else n_total
        self.window = window
        # Hadamard-type butterflies run unnormalised (x0 + x1, x0 - x1): the factor they leave out
        # is the same for every amplitude, so it is carried here and multiplied in by a later op
        # that multiplies every amplitude anyway, at the latest by the last pass of plan()
        self.pending_scale = 1.0

    # -------------------------------------------------------------- passes
    def plan(self, ops, on_step=None):
        """Plan ``ops``.  Returns ``(steps, need)``: ``need`` is None when everything was
        planned, else the sorted non-local positions the next op needs as dense targets
        (multi-GPU: the engine swaps them in and calls :meth:`plan` again with
        ``self.pending``).  ``on_step(step)`` is called for every step as soon as it is built:
        the engine launches it right away, so that the device sweeps pass k while the host
        plans pass k+1."""
        cfg = self.cfg
        pos = self.pos
        lowset = set(range(cfg.L))
        steps = []
        rest = list(ops)
        self.pending = []
        while rest:
            head, tail = rest[: self.window], rest[self.window :]
            first = head[0]
            if first.kind == "mat" and len(first.targets) > cfg.max_mat:
                tp = [pos[q] for q in first.targets]
                if any(p >= cfg.n for p in tp):
                    self.pending = rest
                    return steps, sorted(p for p in tp if p >= cfg.n)
                steps.append(self._big_step(first))
                if on_step is not None:
                    on_step(steps[-1])
                rest = rest[1:]
                continue
            tile_need = set()
            nonlocal_hit = []
            resolved = []

            def accept(op):
                if len(resolved) >= self.max_pass_ops:
                    return False
                if op.kind == "relabel":
                    pos[op.a], pos[op.b] = pos[op.b], pos[op.a]
                    return True
                if op.kind == "mat":
                    if len(op.targets) > cfg.max_mat:
                        return False
                    need = {pos[q] for q in op.targets}
                    if any(p >= cfg.n for p in need):
                        if not nonlocal_hit:
                            nonlocal_hit.append(sorted(p for p in need if p >= cfg.n))
                        return False
                    need -= lowset
                    if len(tile_need | need) > cfg.T - cfg.L:
                        return False
                    tile_need.update(need)
                # ops are resolved to positions when accepted: a Relabel accepted later
                # must not change them
                resolved.append(self._resolve(op))
                return True

            taken, rest2 = _commuting_pick(head, accept)
            if not taken:
                if nonlocal_hit:
                    self.pending = rest
                    return steps, nonlocal_hit[0]
                raise RuntimeError(f"planner made no progress on {head[0]}")
            rest = rest2 + tail
            if resolved:
                # the carried scale is settled by the last pass of the plan and in front of anything
                # that is not a pass (relabels move no data and may follow)
                nxt = next((op for op in rest if op.kind != "relabel"), None)
                settle = nxt is None or (nxt.kind == "mat" and len(nxt.targets) > cfg.max_mat)
                built = self._build_passes(resolved, settle)[0]
                if nxt is None and built and built[-1].kind == "pass":
                    built[-1].last = True  # nothing follows: the engine may ask this pass for its per-tile norms
                steps.extend(built)
                if on_step is not None:
                    for st in built:
                        on_step(st)
        return steps, None

    def _build_passes(self, rops, settle=True):
        """One pass for ``rops``; if its descriptor outgrows the shared-memory staging
        area the list is halved (small states put whole circuits into one tile).  A pass may
        relabel tile positions at its store (see :meth:`_build_pass`): ``self.pos`` and the ops
        still waiting are moved along."""
        lowset = set(range(self.cfg.L))
        need = set()
        for op in rops:
            if op[0] == "mat":
                need.update(op[1])
        need -= lowset
        if len(need) <= self.cfg.T - self.cfg.L:
            try:
                step, sigma = self._build_pass(rops, need, settle)
                for q, x in enumerate(self.pos):
                    self.pos[q] = sigma.get(x, x)
                return [step], sigma
            except DescriptorOverflow:
                pass
        if len(rops) == 1:
            raise RuntimeError("a single op does not fit into a pass descriptor")
        half = len(rops) // 2
        first, sig1 = self._build_passes(rops[:half], False)
        second, sig2 = self._build_passes([_remap_op(op, sig1) for op in rops[half:]], settle)
        both = {x: sig2.get(y, y) for x, y in sig1.items()}
        for x, y in sig2.items():
            both.setdefault(x, y)
        return first + second, both

    def _resolve(self, op):
        """Replace logical qubits by physical positions (the map may change later)."""
        pos = self.pos
        if op.kind == "diag":
            return ("diag", tuple(pos[q] for q in op.qubits), op.turns)
        return ("mat", tuple(pos[q] for q in op.targets), op.matrix, {pos[q]: v for q, v in op.controls.items()})

    def _big_step(self, op):
        pos = self.pos
        tp = [pos[q] for q in op.targets]
        cmask = cval = 0
        for q, v in op.controls.items():
            cmask |= 1 << pos[q]
            cval |= int(v) << pos[q]
        return BigMatStep(tp, op.matrix, cmask, cval)

    # -------------------------------------------------------------- one pass
    def _build_pass(self, rops, tile_need, settle=True):
        cfg = self.cfg
        tile = set(range(cfg.L)) | set(tile_need)
        p = 0
        while len(tile) < cfg.T:
            if p not in tile:
                tile.add(p)
            p += 1
        tile_pos = sorted(tile)
        assert len(tile_pos) == cfg.T and tile_pos[-1] < cfg.n
        tma = cfg.tma and tma_tile_ok(tile_pos, cfg.n)

        # ---- phases: greedy, commutation aware
        phases = []  # (R list, ops)
        rest = list(rops)
        while rest:
            R = []
            fixed = [None]  # target list of a 3/4-qubit dense op: must sit on register bits 0..k-1

            def accept(op):
                if op[0] == "diag":
                    return True
                tg = op[1]
                if len(tg) >= 3:
                    if fixed[0] is not None and fixed[0] != tuple(sorted(tg)):
                        return False
                    new = [t for t in tg if t not in R]
                    if len(R) + len(new) > cfg.r:
                        return False
                    fixed[0] = tuple(sorted(tg))
                    R.extend(new)
                    return True
                new = [t for t in tg if t not in R]
                if len(R) + len(new) > cfg.r:
                    return False
                R.extend(new)
                return True

            taken, rest = _commuting_pick([_RView(o) for o in rest], lambda v: accept(v.op))
            taken = [v.op for v in taken]
            rest = [v.op for v in rest]
            assert taken
            if fixed[0] is not None:
                R = list(fixed[0]) + [x for x in R if x not in fixed[0]]
            phases.append([R, taken, fixed[0] is not None])

        # fill register sets, keep the low positions on the lanes for the first / last phase
        low_lane = set(tile_pos[: min(cfg.conflict_bits, cfg.TB)])
        for ph in phases:
            R = ph[0]
            n_targets = len(R)
            for cand in reversed(tile_pos):
                if len(R) >= cfg.r:
                    break
                if cand not in R and cand not in low_lane:
                    R.append(cand)
            for cand in reversed(tile_pos):
                if len(R) >= cfg.r:
                    break
                if cand not in R:
                    R.append(cand)
            if not ph[2]:
                # register bit order: the first gate's target on the HIGHEST bit, fillers lowest.
                # A butterfly's phases then depend only on register bits below its own
                # (QB2_BFLY_VLOW): later butterflies and idle registers sit below it.
                ph[0] = R = R[n_targets:][::-1] + R[:n_targets][::-1]
        default_R = [x for x in reversed(tile_pos) if x not in low_lane][: cfg.r]
        if len(default_R) < cfg.r:
            default_R = list(reversed(tile_pos))[: cfg.r]
        # (the TMA kernel moves tiles with bulk copies: any first / last distribution reads / writes the tile
        # buffers directly, at worst with a two-way bank conflict)
        if not tma and set(phases[0][0]) & low_lane and len(default_R) == cfg.r and not (set(default_R) & low_lane):
            phases.insert(0, [list(default_R), [], False])
        sigma = {}
        if not tma and set(phases[-1][0]) & low_lane and len(default_R) == cfg.r and not (set(default_R) & low_lane):
            if cfg.store_relabel:
                # The last phase holds some of the lowest positions in registers: a store under that
                # distribution would scatter 8-byte words.  Instead of one more exchange the tile is
                # stored RELABELLED: the positions on the lane bits trade places with the low
                # positions, so every 16-lane group still writes one 128-byte line; the qubits
                # concerned simply live somewhere else from now on (``pos`` follows).
                tp_last = [x for x in tile_pos if x not in phases[-1][0]]
                on_lanes = tp_last[: len(low_lane)]
                movers = [x for x in on_lanes if x not in low_lane]
                free_low = [x for x in sorted(low_lane) if x not in on_lanes]
                assert len(movers) == len(free_low)
                for x, y in zip(movers, free_low):
                    sigma[x], sigma[y] = y, x
            else:
                phases.append([list(default_R), [], False])
        if len(phases) > MAX_PHASES:
            raise DescriptorOverflow(f"{len(phases)} phases in one pass")
        scale0 = self.pending_scale
        try:
            return self._pack(tile_pos, phases, sigma, settle, tma), sigma
        except DescriptorOverflow:
            self.pending_scale = scale0
            raise

    # -------------------------------------------------------------- packing
    def _pack(self, tile_pos, phases, sigma=None, settle=True, tma=False):
        cfg = self.cfg
        T, r, TB, n = cfg.T, cfg.r, cfg.TB, cfg.n
        NR = 1 << r
        csize = 8 if cfg.dtype == QB2_C64 else 16
        real_t = np.float32 if cfg.dtype == QB2_C64 else np.float64
        frac_t = np.uint32 if cfg.dtype == QB2_C64 else np.uint64

        a16, a64, coefs, tabdescs, tables, recs, phases_b = [], [], [], [], [], [], []
        n_tab_entries = 0
        n_gate_ops = 0
        n_slots = 0
        # shared memory left for phase slots when two CTAs share an SM (launch_kernel's layout:
        # root table, per-phase thread contexts, slot area, exchange buffer)
        fsize = 4 if cfg.dtype == QB2_C64 else 8
        threads = 1 << TB
        fixed = (2048 if cfg.dtype == QB2_C64 else 0) + (len(phases) + 1) * threads * 16
        fixed += (csize << T) if len(phases) > 1 else 0
        fixed += 2 * MAX_SLOTS * fsize + 2 * MAX_SLOTS + 16
        slot_cap = max(0, min(MAX_SLOTS, (SMEM_PER_CTA - fixed) // (threads * fsize))) if cfg.phase_slots else 0
        if tma:
            # launch_kernel_tma's layout: roots, barriers, 8-byte thread contexts, slot area, sparse list
            fixed = 2048 + 64 + len(phases) * threads * 8 + 4 * MAX_SLOTS * fsize + 2 * MAX_SLOTS + 16 + 4 * MAX_SPARSE + 16
            slot_cap = max(0, min(MAX_SLOTS, (TMA_AUX_BYTES - fixed) // (threads * fsize))) if cfg.phase_slots else 0

        def add16(vals):
            off = len(a16)
            a16.extend(int(v) & 0xFFFF for v in vals)
            return off

        def add64(vals):
            off = len(a64)
            a64.extend(int(v) for v in vals)
            return off

        prev = None
        op_count = 0
        for R, pops, _fixed in phases:
            assert len(R) == r and len(set(R)) == r
            TP = [x for x in tile_pos if x not in R]
            pcol = add64(1 << x for x in TP)
            rphys = add64(sum(((j >> i) & 1) << R[i] for i in range(r)) for j in range(NR))
            xw = xr = xflags = 0
            if prev is None and tma:
                # the tile buffers of the TMA kernel: this phase reads the buffer the load fills, the
                # LAST phase writes the buffer the store drains
                ccol = dict(zip(tile_pos, tma_columns(T)))
                lR = phases[-1][0]
                lTP = [x for x in tile_pos if x not in lR]
                xw = add16([ccol[x] for x in lTP] + [_xor_cols(ccol, lR, j) for j in range(NR)])
                xr = add16([ccol[x] for x in TP] + [_xor_cols(ccol, R, j) for j in range(NR)])
                kw, kr = _regular_shift(ccol, lR, lTP), _regular_shift(ccol, R, TP)
                if kw is not None:
                    xflags |= 1 | (kw << 8)
                if kr is not None:
                    xflags |= 2 | (kr << 16)
            if prev is not None:
                pR, pTP = prev
                col = _slot_columns(tile_pos, pR, pTP, R, TP, cfg.conflict_bits)
                xw = add16([col[x] for x in pTP] + [_xor_cols(col, pR, j) for j in range(NR)])
                xr = add16([col[x] for x in TP] + [_xor_cols(col, R, j) for j in range(NR)])
                kw, kr = _regular_shift(col, pR, pTP), _regular_shift(col, R, TP)
                if kw is not None:
                    xflags |= 1 | (kw << 8)
                if kr is not None:
                    xflags |= 2 | (kr << 16)
            first_op = op_count
            # ---- ops: merge runs of diagonal terms; fuse "2x2 gate + phases on its 1-output"
            def emit_tables(dop):
                nonlocal n_tab_entries
                for tab in dop["thr"] + dop["reg"]:
                    if tab[0] == "mul":
                        tabdescs.append(tab)
                        continue
                    _, positions, vals = tab
                    thr_positions = [x for x in positions if x not in R]
                    reg_positions = [x for x in positions if x in R]
                    nthr = len(thr_positions)
                    regoff = 0
                    if reg_positions and dop["kind"] == OP_DIAG_BIT:
                        regoff = 1 << nthr  # the b0=1 half follows the b0=0 half
                    elif reg_positions:
                        regoff = add16(
                            sum(((j >> R.index(x)) & 1) << (nthr + k) for k, x in enumerate(reg_positions))
                            for j in range(NR)
                        )
                    tabdescs.append((_pieces(thr_positions), n_tab_entries, regoff))
                    tables.append(_to_frac(vals, frac_t))
                    n_tab_entries += len(vals)

            def take_slot(dop):
                """1 + phase slot for an op whose descriptors are all multiplier ladders, else 0."""
                nonlocal n_slots
                need = 2 if dop["kind"] == OP_DIAG_BIT and not dop.get("flags", 0) & 1 else 1
                if dop["kind"] not in (OP_DIAG_THR, OP_DIAG_BIT) or n_slots + need > slot_cap:
                    return 0
                tabs_ = dop["thr"] + dop["reg"]
                if not tabs_ or any(t[0] != "mul" for t in tabs_):
                    return 0
                first = n_slots + 1
                n_slots += need
                return first

            def emit_diag(dop):
                nonlocal op_count
                data = len(tabdescs)
                if dop["kind"] == OP_DIAG_VEC:
                    data = len(coefs)
                    coefs.extend(np.exp(2j * np.pi * dop["vec"]) * self.pending_scale)
                    self.pending_scale = 1.0
                emit_tables(dop)
                recs.append(dict(cmask=0, cval=0, jmask=0, data=data, kind=dop["kind"], b0=dop.get("b0", 0), b1=take_slot(dop),
                                 nthr=len(dop["thr"]), nreg=len(dop["reg"]), flags=dop.get("flags", 0), tab=0))
                op_count += 1

            def diag_run(start):
                j = start
                while j < len(pops) and pops[j][0] == "diag":
                    j += 1
                return pops[start:j], j

            i = 0
            while i < len(pops):
                op = pops[i]
                if op[0] == "diag":
                    terms, i = diag_run(i)
                    n_gate_ops += len(terms)
                    for dop in _build_diag_ops(terms, R, cfg):
                        emit_diag(dop)
                    continue
                _, tg, mat, ctrls = op
                n_gate_ops += 1
                k = len(tg)
                bits = [R.index(t) for t in tg]
                order = sorted(range(k), key=lambda a: bits[a])
                if k >= 3:
                    assert sorted(bits) == list(range(k)), "3/4-qubit dense ops sit on register bits 0..k-1"
                m = _permute_matrix(mat, order)
                sbits = sorted(bits)
                i += 1
                if k == 1 and not ctrls and cfg.fuse_bfly and np.max(np.abs(m.imag)) < 1e-16:
                    # butterfly: the phases that follow and vanish for b0=0 ride on the gate
                    terms, nxt = diag_run(i)
                    if terms:
                        b0 = sbits[0]
                        dops = _build_diag_ops(terms, R, cfg)
                        vec_op = next((d for d in dops if d["kind"] == OP_DIAG_VEC), None)
                        if vec_op is not None:
                            t = vec_op["vec"]
                            zero_half = [j for j in range(NR) if not (j >> b0) & 1]
                            if np.any(np.abs(t[zero_half] - np.rint(t[zero_half])) > 0):
                                vec_op = None
                        bit_op = next(
                            (d for d in dops if d["kind"] == OP_DIAG_BIT and d["b0"] == b0 and d.get("flags", 0) & 1
                             and len(d["reg"]) <= 255), None)
                        if vec_op is not None or bit_op is not None:
                            n_gate_ops += len(terms)
                            i = nxt
                            flags = BFLY_REAL if np.max(np.abs(m.imag)) < 1e-16 else 0
                            mr = m.real
                            if cfg.had_bfly and mr[0, 0] != 0 and mr[0, 0] == mr[0, 1] == mr[1, 0] == -mr[1, 1]:
                                # s * [[1, 1], [1, -1]]: run as add / subtract, s joins the pending scale
                                flags |= BFLY_HAD
                                self.pending_scale *= float(mr[0, 0])
                            data = len(coefs)
                            coefs.extend(m.reshape(-1))
                            if vec_op is not None:
                                flags |= BFLY_V
                                pairs = [((g >> b0) << (b0 + 1)) | (g & ((1 << b0) - 1)) | (1 << b0) for g in range(NR // 2)]
                                vturns = vec_op["vec"][pairs]
                                low = np.arange(NR // 2) & ((1 << b0) - 1)
                                if np.array_equal(vturns, vturns[low]):
                                    flags |= BFLY_VLOW
                                coefs.extend(np.exp(2j * np.pi * vturns))
                            else:
                                coefs.extend(np.ones(NR // 2, dtype=np.complex128))
                            tab0 = len(tabdescs)
                            ntab = 0
                            slot = 0
                            if bit_op is not None:
                                flags |= BFLY_W
                                emit_tables(bit_op)
                                ntab = len(bit_op["reg"])
                                slot = take_slot(bit_op)
                            assert tab0 < 65536
                            recs.append(dict(cmask=0, cval=0, jmask=0, data=data, kind=OP_BFLY, b0=b0, b1=slot, nthr=0, nreg=ntab,
                                             flags=flags, tab=tab0))
                            op_count += 1
                            for d in dops:
                                if d is not vec_op and d is not bit_op:
                                    emit_diag(d)
                            continue
                cmask = cval = 0
                jmask = 0
                regc = {R.index(p): v for p, v in ctrls.items() if p in R}
                for p, v in ctrls.items():
                    if p not in R:
                        cmask |= 1 << p
                        cval |= int(v) << p
                for j in range(NR):
                    if all(((j >> b) & 1) == v for b, v in regc.items()):
                        jmask |= 1 << j
                shape = _mat1_shape(m) if k == 1 else 0
                if self.pending_scale != 1.0 and not ctrls and not (k == 1 and shape == MAT1_SWAP):
                    m = m * self.pending_scale  # every amplitude passes through this matrix
                    self.pending_scale = 1.0
                data = len(coefs)
                coefs.extend(m.reshape(-1))
                b1 = sbits[1] if k > 1 else 0
                if k == 1 and shape == MAT1_SWAP and len(regc) == 1 and list(regc.values())[0] == 1:
                    b1 = 1 + list(regc.keys())[0]  # CX inside the registers: a fixed register permutation
                recs.append(dict(cmask=cmask, cval=cval, jmask=jmask, data=data, kind=OP_MAT[k], b0=sbits[0], b1=b1, nthr=0, nreg=0,
                                 flags=shape, tab=0))
                op_count += 1
            if settle and R is phases[-1][0] and self.pending_scale != 1.0:
                emit_diag({"kind": OP_DIAG_VEC, "vec": np.zeros(NR), "thr": [], "reg": []})
            if cfg.fuse_chains:
                fuse_butterfly_chains(recs, first_op, NR)
                fuse_diag_bit_runs(recs, first_op)
                fuse_cx_ladders(recs, first_op)
            phases_b.append(struct.pack("<8I", first_op, op_count - first_op, pcol, rphys, xw, xr, xflags, 0))
            prev = (R, TP)

        # ---- segments: tile id bits -> non-tile local positions
        segs = []
        src = 0
        p = 0
        nontile = [x for x in range(n) if x not in set(tile_pos)]
        i = 0
        while i < len(nontile):
            j = i
            while j + 1 < len(nontile) and nontile[j + 1] == nontile[j] + 1:
                j += 1
            segs.append((src, nontile[i], j - i + 1))
            src += j - i + 1
            i = j + 1
        if len(segs) > MAX_SEGS:
            raise RuntimeError("too many tile segments")

        ops_b = [pack_op(NR=NR, **r) for r in recs]

        # ---- layout of the small section
        def align(x, a):
            return (x + a - 1) // a * a

        off = HEADER_BYTES + IO_BYTES
        phases_off = off
        off += 32 * len(phases_b)
        ops_off = off
        off += 32 * len(ops_b)
        off = align(off, 16)
        coef_off = off
        off += csize * len(coefs)
        off = align(off, 8)
        u64_off = off
        off += 8 * len(a64)
        tabdesc_off = off
        off += 32 * len(tabdescs)
        u16_off = off
        off += 2 * len(a16)
        small = align(off, 16)
        if small > MAX_SMALL_BYTES:
            raise DescriptorOverflow(f"pass descriptor of {small} bytes exceeds the staging limit")
        tab_off = align(small, 256)
        total = align(tab_off + n_tab_entries * np.dtype(frac_t).itemsize, 256)

        # ops outside the lean kernel's repertoire (include/qrisp_b200.h QB2_FEATURE_FULL)
        features = 1 if any(rec[24] in (2, 3, 4, OP_DIAG) for rec in ops_b) else 0  # rec[24] = qb2_op.kind
        features |= FEATURE_TMA if tma else 0
        blob = bytearray(total)
        hdr = struct.pack(
            "<IIIIBBBBIIIIIIIIIIHH",
            QB2_PASS_MAGIC,
            total,
            small,
            n,
            T,
            r,
            cfg.dtype,
            len(segs),
            len(phases_b),
            phases_off,
            ops_off,
            coef_off,
            u16_off,
            u64_off,
            tabdesc_off,
            tab_off,
            len(ops_b),
            features,
            0,
            n_slots,
        )
        assert len(hdr) == 64
        blob[0:64] = hdr
        for i, (s, d, ln) in enumerate(segs):
            blob[64 + 4 * i : 68 + 4 * i] = struct.pack("<BBBB", s, d, ln, 0)
        # qb2_io: where register j of a thread goes at the load (first phase) and at the store (last)
        sigma = sigma or {}
        io = np.zeros(IO_BYTES // 8, dtype=np.uint64)
        R0, Rl = phases[0][0], phases[-1][0]
        for j in range(NR):
            io[j] = sum(((j >> i) & 1) << R0[i] for i in range(r)) * csize
            io[32 + j] = sum(((j >> i) & 1) << sigma.get(Rl[i], Rl[i]) for i in range(r)) * csize
        for b, x in enumerate(x for x in tile_pos if x not in Rl):
            io[64 + b] = 1 << sigma.get(x, x)  # store_col: physical offset of thread bit b at the store
        blob[HEADER_BYTES : HEADER_BYTES + IO_BYTES] = io.tobytes()
        blob[phases_off : phases_off + 32 * len(phases_b)] = b"".join(phases_b)
        blob[ops_off : ops_off + 32 * len(ops_b)] = b"".join(ops_b)
        if coefs:
            c = np.asarray(coefs, dtype=np.complex128)
            inter = np.empty(2 * len(c), dtype=real_t)
            inter[0::2] = c.real
            inter[1::2] = c.imag
            blob[coef_off : coef_off + inter.nbytes] = inter.tobytes()
        if a64:
            blob[u64_off : u64_off + 8 * len(a64)] = np.asarray(a64, dtype=np.uint64).tobytes()
        for i, td in enumerate(tabdescs):
            if td[0] == "mul":
                _, shift, mask, mult, brev, rsh = td
                rec = struct.pack("<3xBBBBBIIQQ", 0, shift, brev, rsh, 1, mask, 0, mult, 0)
            else:
                pieces, toff, regoff = td
                pb = b"".join(
                    struct.pack("<BBBB", s, ln, o, len(pieces) if k == 0 else 0) for k, (s, ln, o) in enumerate(pieces)
                )
                pb += b"\0" * (24 - len(pb))
                rec = pb + struct.pack("<II", toff, regoff)
            assert len(rec) == 32
            blob[tabdesc_off + 32 * i : tabdesc_off + 32 * i + 32] = rec
        if a16:
            blob[u16_off : u16_off + 2 * len(a16)] = np.asarray(a16, dtype=np.uint16).tobytes()
        if tables:
            tb = np.concatenate(tables)
            blob[tab_off : tab_off + tb.nbytes] = tb.tobytes()
        R0 = phases[0][0]
        return PassStep(bytes(blob), op_count, len(phases_b), tile_pos, n_gate_ops, segs=segs, R0=R0,
                        TP0=[x for x in tile_pos if x not in R0], tma=tma)


class _RView:
    """Adapter giving a resolved op the ``nd`` / ``d`` interface of the list scheduler
    (in terms of physical positions)."""

    __slots__ = ("op",)

    def __init__(self, op):
        self.op = op

    @property
    def nd(self):
        return self.op[1] if self.op[0] == "mat" else ()

    @property
    def d(self):
        return self.op[1] if self.op[0] == "diag" else tuple(self.op[3].keys())

    @property
    def kind(self):
        return self.op[0]


# ----------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------


def _remap_op(op, sigma):
    """A resolved op after the tile positions moved by ``sigma`` (a store relabel)."""
    if not sigma:
        return op
    if op[0] == "diag":
        return ("diag", tuple(sigma.get(x, x) for x in op[1]), op[2])
    return ("mat", tuple(sigma.get(x, x) for x in op[1]), op[2], {sigma.get(x, x): v for x, v in op[3].items()})


def _permute_matrix(mat, order):
    """New matrix whose index bit ``a`` is the old index bit ``order[a]``."""
    k = len(order)
    if order == list(range(k)):
        return np.asarray(mat, dtype=np.complex128)
    d = 1 << k
    idx = np.zeros(d, dtype=np.int64)
    for a in range(k):
        idx |= ((np.arange(d) >> a) & 1) << order[a]
    return np.asarray(mat, dtype=np.complex128)[np.ix_(idx, idx)]


def _slot_columns(tile_pos, write_R, write_tp, read_R, read_tp, conflict_bits):
    """Shared-memory slot map of an exchange: XOR-linear in the tile-local index.

    Returns ``col[position] -> uint16``; the slot of a tile-local index is the XOR of the
    columns of its set bits.  The map is built so that the ``2**c`` lanes of a wavefront hit
    ``2**c`` different bank groups both when the tile is stored under the old distribution
    (lane bits = ``write_tp[:c]``) and when it is loaded under the new one (``read_tp[:c]``).
    Above the lane rows come the old register positions, then the new ones, each in register
    order: when they do not collide with a lane row, register ``j`` simply sits at slot
    ``j << k`` and the kernel needs no per-register table.
    """
    c = min(conflict_bits, len(write_tp), len(read_tp))
    W = list(write_tp[:c])
    Rr = list(read_tp[:c])
    col = {}
    for i, x in enumerate(W):
        col[x] = 1 << i
    free_rows = [i for i, x in enumerate(W) if x not in Rr]
    read_only = [x for x in Rr if x not in W]
    assert len(free_rows) == len(read_only)
    nxt = c
    for x in list(write_R) + list(read_R) + list(tile_pos):
        if x in col:
            continue
        col[x] = 1 << nxt
        nxt += 1
    for row, x in zip(free_rows, read_only):
        col[x] ^= 1 << row
    return col


def _regular_shift(col, R, TP):
    """k such that register j sits at slot ``j << k`` (columns of R are 1<<k, 1<<(k+1)...) and no
    thread column touches those slot bits (so that XOR and + agree), else None."""
    c0 = col[R[0]]
    if c0 & (c0 - 1):
        return None
    k = c0.bit_length() - 1
    for i, x in enumerate(R):
        if col[x] != 1 << (k + i):
            return None
    span = ((1 << len(R)) - 1) << k
    if any(col[x] & span for x in TP):
        return None
    return k


def _xor_cols(col, R, j):
    v = 0
    for i, x in enumerate(R):
        if (j >> i) & 1:
            v ^= col[x]
    return v


def _pieces(positions):
    """Runs of consecutive positions -> (shift, len, out) pieces; index bit k <-> positions[k]."""
    pieces = []
    i = 0
    while i < len(positions):
        j = i
        while j + 1 < len(positions) and positions[j + 1] == positions[j] + 1:
            j += 1
        pieces.append((positions[i], j - i + 1, i))
        i = j + 1
    assert len(pieces) <= MAX_PIECES
    return pieces


def _n_runs(positions):
    return sum(1 for i, x in enumerate(positions) if i == 0 or positions[i - 1] != x - 1)


def _to_frac(turns, frac_t):
    t = np.asarray(turns, dtype=np.float64)
    t = t - np.floor(t)
    if frac_t == np.uint32:
        v = np.rint(t * 4294967296.0)
        v = np.where(v >= 4294967296.0, 0.0, v)
        return v.astype(np.uint64).astype(np.uint32)
    hi = np.floor(t * 4294967296.0)
    lo = np.rint((t * 4294967296.0 - hi) * 4294967296.0)
    carry = lo >= 4294967296.0
    lo = np.where(carry, 0.0, lo)
    hi = np.where(carry, hi + 1, hi)
    hi = np.where(hi >= 4294967296.0, 0.0, hi)
    return (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)


def _mat1_shape(m):
    """Which specialised 2x2 kernel path applies (include/qrisp_b200.h QB2_MAT1_*)."""
    tiny = 1e-16
    if abs(m[0, 0]) < tiny and abs(m[1, 1]) < tiny:
        if abs(m[0, 1] - 1) < tiny and abs(m[1, 0] - 1) < tiny:
            return MAT1_SWAP
        return MAT1_ANTI
    if np.max(np.abs(m.imag)) < tiny:
        return MAT1_REAL
    return MAT1_GENERAL


def _merge_tables(terms, Rset, max_bits):
    """Greedy merge of diagonal terms into tables.  Returns [(positions, [terms])] where
    ``positions`` is the union of the terms' positions (a set)."""
    tabs = []
    for positions, turns in terms:
        pset = set(positions)
        for tab in tabs:
            u = tab[0] | pset
            if len(u) <= max_bits and _n_runs(sorted(x for x in u if x not in Rset)) <= MAX_PIECES:
                tab[0] = u
                tab[1].append((positions, turns))
                break
        else:
            if len(pset) > max_bits or _n_runs(sorted(x for x in pset if x not in Rset)) > MAX_PIECES:
                raise RuntimeError("diagonal term too wide for a phase table")
            tabs.append([pset, [(positions, turns)]])
    return tabs


def _table_values(order, tterms):
    """Sum of the terms over the index space whose bit b is position ``order[b]``."""
    nb = len(order)
    idx = np.arange(1 << nb)
    vals = np.zeros(1 << nb, dtype=np.float64)
    for positions, turns in tterms:
        sub = np.zeros(1 << nb, dtype=np.int64)
        for b, x in enumerate(positions):
            sub |= ((idx >> order.index(x)) & 1) << b
        vals += turns[sub]
    return vals


def _build_diag_ops(terms, R, cfg):
    """Turn a run of (commuting) diagonal terms into DIAG ops.

    The terms are split by how their phase depends on the register bits of the phase,
    because that decides how much work a thread has to do per amplitude:

    * no register bit       -> thread table (one lookup per thread);
    * register bits only    -> one host-computed vector of unit factors (DIAG_VEC);
    * exactly one register bit (+ thread bits) -> DIAG_BIT: two phases per thread;
    * anything else         -> general DIAG: a table sum per amplitude.

    Every table is ``(positions, turns)``: thread/tile positions ascending as the low index
    bits (neighbouring lanes read neighbouring entries), register positions above them.
    """
    Rset = set(R)
    NR = 1 << len(R)
    vec = np.zeros(NR, dtype=np.float64)
    have_vec = False
    thr_terms, gen_terms = [], []
    bit_terms = {}
    for _, positions, turns in terms:
        regp = [x for x in positions if x in Rset]
        if len(regp) == len(positions):
            j = np.arange(NR)
            sub = np.zeros(NR, dtype=np.int64)
            for b, x in enumerate(positions):
                sub |= ((j >> R.index(x)) & 1) << b
            vec += turns[sub]
            have_vec = True
        elif not regp:
            thr_terms.append((positions, turns))
        elif len(regp) == 1:
            bit_terms.setdefault(R.index(regp[0]), []).append((positions, turns))
        else:
            gen_terms.append((positions, turns))

    frac_bits = 32 if cfg.dtype == QB2_C64 else 64
    thr_tabs = []
    for pset, tterms in _merge_tables(thr_terms, Rset, MAX_TAB_BITS):
        order = sorted(pset)
        thr_tabs.append(("tab", order, _table_values(order, tterms)))

    ops = []
    if have_vec and np.any(np.abs(vec - np.rint(vec)) > 0):
        ops.append({"kind": OP_DIAG_VEC, "vec": vec, "thr": [], "reg": []})
    for b in sorted(bit_terms):
        muls, rest = _fit_multipliers(bit_terms[b], R[b], frac_bits)
        reg_tabs = list(muls)
        for pset, tterms in _merge_tables(rest, Rset, MAX_TAB_BITS):
            order = sorted(x for x in pset if x not in Rset) + [R[b]]
            reg_tabs.append(("tab", order, _table_values(order, tterms)))
        for s in range(0, len(reg_tabs), 200):
            ops.append({"kind": OP_DIAG_BIT, "b0": b, "thr": [], "reg": reg_tabs[s : s + 200]})
    gen_tabs = []
    for pset, tterms in _merge_tables(gen_terms, Rset, MAX_TAB_BITS):
        order = sorted(x for x in pset if x not in Rset) + sorted((x for x in pset if x in Rset), key=R.index)
        gen_tabs.append(("tab", order, _table_values(order, tterms)))
    for s in range(0, len(gen_tabs), MAX_REG_TABS):
        ops.append({"kind": OP_DIAG, "thr": [], "reg": gen_tabs[s : s + MAX_REG_TABS]})
    # the thread tables ride on the first op that evaluates tables anyway
    while thr_tabs:
        chunk, thr_tabs = thr_tabs[:200], thr_tabs[200:]
        host = next((o for o in ops if o["kind"] in (OP_DIAG_BIT, OP_DIAG) and not o["thr"]), None)
        if host is None:
            ops.append({"kind": OP_DIAG_THR, "thr": chunk, "reg": []})
        else:
            host["thr"] = chunk
    frac_t = np.uint32 if cfg.dtype == QB2_C64 else np.uint64
    for o in ops:
        if o["kind"] == OP_DIAG_BIT and not o["thr"]:
            if all(t[0] == "mul" or not np.any(_to_frac(t[2][: len(t[2]) // 2], frac_t)) for t in o["reg"]):
                o["flags"] = 1  # the b0=0 phase is exactly zero: skip those amplitudes
    return ops


def _fit_multipliers(terms, regpos, frac_bits):
    """Controlled-phase ladders: terms ``exp(2 pi i theta_a x_b x_a)`` whose angles are
    geometric in the partner position, ``theta_a = c * 2**(p_a - shift)`` mod 1, need no
    table: phase = c * (index bits).  Returns (multiplier entries, remaining terms)."""
    mod = 1 << frac_bits
    cands, rest = [], []
    scale = 4294967296.0 * (4294967296.0 if frac_bits == 64 else 1.0)
    for positions, turns in terms:
        if len(positions) == 2:
            # (plain floats: this runs once per controlled phase of the circuit)
            ib = 0 if positions[0] == regpos else 1
            t00, t01, t10, t11 = float(turns[0]), float(turns[1]), float(turns[2]), float(turns[3])  # [bit1, bit0]
            # entries other than x_a = x_b = 1 must be whole turns (phase 1 to 1e-15)
            if max(abs(t00 - round(t00)), abs(t01 - round(t01)), abs(t10 - round(t10))) < 1e-16:
                a = positions[1 - ib]
                f = _frac64(t11) if frac_bits == 64 else int(round((t11 - math.floor(t11)) * scale))
                cands.append((a, f % mod, [(positions, turns)]))
                continue
        rest.append((positions, turns))
    # several terms on the same pair of positions (CZ twice, CP after CP) are ONE rung of a ladder:
    # their angles add (a ladder's mask can hold a partner position only once)
    merged = {}
    for a, f, orig in cands:
        if a in merged:
            merged[a] = (a, (merged[a][1] + f) % mod, merged[a][2] + orig)
        else:
            merged[a] = (a, f, orig)
    cands = list(merged.values())
    muls = []
    cands.sort(key=lambda c: c[0])

    def close(x, y):
        return abs((x - y + mod // 2) % mod - mod // 2) <= 4

    while cands:
        shift = cands[0][0]
        window = [c for c in cands if c[0] - shift < 32]
        # ascending ladder: theta_a = c0 * 2**(a - shift)
        c_up = window[0][1]
        up = [c for c in window if close(c_up << (c[0] - shift), c[1])]
        # descending ladder: theta_a = c1 * 2**(top - a)
        top = window[-1][0]
        c_dn = window[-1][1]
        dn = [c for c in window if close(c_dn << (top - c[0]), c[1])]
        group = up if len(up) >= len(dn) else dn
        if len(group) >= 1:
            mask = 0
            for a, _, _ in group:
                mask |= 1 << (a - shift)
            if group is up:
                muls.append(("mul", shift, mask, c_up, 0, 0))
            else:
                muls.append(("mul", shift, mask, c_dn, 1, 31 - (top - shift)))
            ids = {id(c) for c in group}
            cands = [c for c in cands if id(c) not in ids]
        else:
            rest.extend(cands[0][2])
            cands = cands[1:]
    return muls, rest


def _frac64(turn):
    t = float(turn) - np.floor(float(turn))
    hi = int(np.floor(t * 4294967296.0))
    lo = int(np.rint((t * 4294967296.0 - hi) * 4294967296.0))
    return ((hi << 32) + lo) % (1 << 64)
