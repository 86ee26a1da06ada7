"""Specialised pass kernels: one kernel per pass, its descriptors compiled in.

The pass kernel is an interpreter: every tile it re-reads the pass's descriptors (constant bank), dispatches on
op codes and computes exchange addresses from tables.  For a program that is replayed -- a benchmark loop, a
variational circuit, any circuit whose structure repeats -- that decode work can be paid ONCE at compile time:
this module writes a translation unit that holds the pass's small section as a ``const`` array and its decode
schedule as ``constexpr`` tables, includes the very same kernel source (``csrc/qb2_pass_tma.cuh``,
``pass_tma_body<FULL, JIT=true>``) and compiles it with ``nvcc -cubin -arch=sm_100a``.  With the descriptors
constant the optimiser folds the reads, resolves every ``switch`` and keeps only the arms the pass uses; what
is left is straight-line arithmetic plus the tile movement.  Results are those of the interpreter (same source,
same arithmetic).

Cubins are cached by the SHA-256 of (small section, kernel sources) under ``qrisp_b200/jit_cache/`` (override:
``QB2_JIT_CACHE``) and loaded through ``qb2_jit_load``.  Compilation is opt-in
(``StateVector(..., specialize=True)`` / ``QB2_SPECIALIZE=1`` / ``Program.specialize()``): it costs seconds per
pass, which only a replayed program earns back.  By default an engine uses the kernels the cache already holds
and compiles nothing (``specialize=False`` / ``QB2_SPECIALIZE=0``: interpreter only).  Without ``nvcc`` the
interpreter simply stays in charge.  Both pass kernels (TMA and register-pipelined), every geometry.
"""

from __future__ import annotations

import ctypes
import hashlib
import os
import shutil
import struct
import subprocess
import tempfile
import threading

from . import _lib
from .planner import CHAIN_SHAPES, CODE_CHAIN, CODE_CXL, CODE_DBITS, CXL_SHAPES, FEATURE_FULL

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
SOURCES = ("qb2_pass_ops_loop.inc", "qb2_pass_tma.cuh", "qb2_pass_tma_phase.inc", "qb2_pass_kernel.cuh", "qb2_pass_kernel_phase.inc", "qb2_pass_ops.inc",
           "qb2_pass_common.cuh", "qb2_internal.cuh")

_kernels = {}  # key -> kernel handle (int) or None when the build failed
_lock = threading.Lock()
_source_digest = None


def cache_dir():
    d = os.environ.get("QB2_JIT_CACHE") or os.path.join(HERE, "jit_cache")
    os.makedirs(d, exist_ok=True)
    return d


def find_nvcc():
    for cand in (os.environ.get("QB2_NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def _sources_digest():
    global _source_digest
    if _source_digest is None:
        h = hashlib.sha256()
        for name in SOURCES:
            with open(os.path.join(CSRC, name), "rb") as f:
                h.update(f.read())
        with open(os.path.join(INCLUDE, "qrisp_b200.h"), "rb") as f:
            h.update(f.read())
        _source_digest = h.digest()
    return _source_digest


def small_section(blob):
    small = int.from_bytes(blob[8:12], "little")
    return bytes(blob[:small])


def decode_schedule(small):
    """Per phase, the op indices at which the interpreter decodes (a fused run is ONE decode): what the loop
    ``for (o = first; o < end; ++o) { ...; o += consumed - 1; }`` of the kernel visits."""
    (n_phases, phases_off, ops_off) = struct.unpack_from("<III", small, 20)
    sched = []
    for p in range(n_phases):
        first, count = struct.unpack_from("<II", small, phases_off + 32 * p)
        o, points = first, []
        while o < first + count:
            points.append(o)
            code, = struct.unpack_from("<H", small, ops_off + 32 * o + 22)
            tab, = struct.unpack_from("<H", small, ops_off + 32 * o + 30)
            if CODE_CHAIN <= code < CODE_CHAIN + len(CHAIN_SHAPES):
                o += CHAIN_SHAPES[code - CODE_CHAIN][1]
            elif code == CODE_DBITS:
                o += bin(tab).count("1")
            elif CODE_CXL <= code < CODE_CXL + len(CXL_SHAPES):
                o += CXL_SHAPES[code - CODE_CXL][1]
            else:
                o += 1
        sched.append(points)
    return sched


# kernel geometry of the register-pipelined kernel by (dtype, r, thread bits): real, RB, MAXT, MINB -- the table
# of launch_pass (csrc/qb2_pass.cu)
def _reg_geometry(dtype, r, tb):
    if dtype == 0:
        if r == 4 and tb <= 8:
            return "float", 4, 256, 4
        if r == 4 and tb == 9:
            return "float", 4, 512, 2
        if r == 5 and tb <= 8:
            return "float", 5, 256, 2
        if r == 3 and tb <= 8:
            return "float", 3, 256, 6
    elif dtype == 1:
        if r == 3 and tb <= 8:
            return "double", 3, 256, 4
        if r == 4 and tb <= 8:
            return "double", 4, 256, 2
    return None


def _structure(small):
    """What a specialised kernel is compiled from: header geometry, phase and op records, the u16 area (slot
    maps) and the decode schedule."""
    T, r, dtype, _nseg = struct.unpack_from("<BBBB", small, 16)
    (n_phases, phases_off, ops_off, _coef_off, u16_off) = struct.unpack_from("<IIIII", small, 20)
    n_ops, features = struct.unpack_from("<II", small, 52)
    phases = [struct.unpack_from("<8I", small, phases_off + 32 * i) for i in range(n_phases)]
    ops = [struct.unpack_from("<QQIHHBBBBBBH", small, ops_off + 32 * i) for i in range(n_ops)]
    a16 = struct.unpack_from(f"<{(len(small) - u16_off) // 2}H", small, u16_off)
    return dict(T=T, r=r, dtype=dtype, features=features, phases=phases, ops=ops, a16=a16, sched=decode_schedule(small))


def kernel_source(small, tma):
    """The translation unit of one pass: its STRUCTURE (phases, op records, decode schedule, slot maps) as
    ``constexpr`` tables -- front-end constants, so that only the arms in use are ever instantiated --, then the
    kernel source itself.  Coefficients and tables stay what they are for the interpreter: read from the
    by-value kernel parameter (constant bank), at offsets that are now immediates."""
    st = _structure(small)
    sched = st["sched"]
    flat = [o for pts in sched for o in pts] or [0]
    offs, acc = [], 0
    for pts in sched:
        offs.append(acc)
        acc += len(pts)
    full = "true" if st["features"] & FEATURE_FULL else "false"
    lines = [
        "// generated by qrisp_b200/jit.py: one pass, structure compiled in",
        '#include "qb2_internal.cuh"',
        f"#define QB2_JIT_NPHASES {len(st['phases'])}",
        f"#define QB2_JIT_TB {st['T'] - st['r']}",
        '#define QB2_JIT_PHASES_FILE "jit_phases.inc"',
        "namespace qb2 {",
        "namespace {",
        "__host__ __device__ constexpr uint32_t jit_ndec(uint32_t p) {",
        f"    constexpr uint32_t t[] = {{{', '.join(str(len(pts)) for pts in sched)}}};",
        "    return t[p];",
        "}",
        "__host__ __device__ constexpr uint32_t jit_sched(uint32_t p, uint32_t i) {",
        f"    constexpr uint32_t off[] = {{{', '.join(map(str, offs))}}};",
        f"    constexpr uint32_t s[] = {{{', '.join(map(str, flat))}}};",
        "    return s[off[p] + i];",
        "}",
        "__host__ __device__ constexpr qb2_phase jit_phase(uint32_t p) {",
        "    constexpr qb2_phase t[] = {",
    ]
    for ph in st["phases"]:
        lines.append("        {" + ", ".join(f"{v}u" for v in ph) + "},")
    lines += ["    };", "    return t[p];", "}", "__host__ __device__ constexpr qb2_op jit_op(uint32_t o) {", "    constexpr qb2_op t[] = {"]
    for op in st["ops"] or [(0,) * 12]:
        cmask, cval, jmask, data, code, kind, b0, b1, nthr, nreg, flags, tab = op
        lines.append(f"        {{0x{cmask:x}ull, 0x{cval:x}ull, 0x{jmask:x}u, {data}, {code}, {kind}, {b0}, {b1}, {nthr}, {nreg}, {flags}, {tab}}},")
    lines += ["    };", "    return t[o];", "}", "__host__ __device__ constexpr uint32_t jit_a16(uint32_t i) {", "    constexpr uint16_t t[] = {"]
    a16 = list(st["a16"]) or [0]
    for i in range(0, len(a16), 24):
        lines.append("        " + ", ".join(map(str, a16[i : i + 24])) + ",")
    lines += ["    };", f"    return i < {len(a16)}u ? t[i] : 0u;", "}", "}  // namespace", "}  // namespace qb2"]
    if tma:
        lines += [
            '#include "qb2_pass_tma.cuh"',
            'extern "C" __global__ void __launch_bounds__(2 * qb2::TMA_GROUP, 1)',
            "    qb2_pass_jit(const __grid_constant__ CUtensorMap tmap, const uint32_t *__restrict__ tables, const uint64_t hi_base,",
            "                 const uint32_t ntiles, const uint32_t ctx_off, const uint32_t slot_off, const uint32_t sp_off,",
            "                 const uint32_t inv_mask, const uint32_t features, const qb2::SparseInput SP,",
            "                 const __grid_constant__ qb2::PassParams P) {",
            f"    qb2::pass_tma_body<{full}, true>(tmap, tables, hi_base, ntiles, ctx_off, slot_off, sp_off, inv_mask, features, SP,",
            "                                   reinterpret_cast<const uint8_t *>(&P));",
            "}",
            "",
        ]
    else:
        real, rb, maxt, minb = _reg_geometry(st["dtype"], st["r"], st["T"] - st["r"])
        lines += [
            '#include "qb2_pass_kernel.cuh"',
            f'extern "C" __global__ void __launch_bounds__({maxt}, {minb})',
            f"    qb2_pass_jit(qb2::Traits<{real}>::vec *__restrict__ state, const qb2::Traits<{real}>::frac *__restrict__ tables,",
            "                 const uint64_t hi_base, const uint32_t ntiles, const uint32_t ctx_off, const uint32_t slot_off,",
            "                 const uint32_t tile_off, const qb2::SparseInput SP, const uint32_t sp_off,",
            "                 const __grid_constant__ qb2::PassParams P) {",
            f"    qb2::pass_body<{real}, {rb}, {full}, true>(state, tables, hi_base, ntiles, ctx_off, slot_off, tile_off, SP, sp_off,",
            "                                         reinterpret_cast<const uint8_t *>(&P));",
            "}",
            "",
        ]
    return "\n".join(lines)


def schedule_files(small):
    """The unrolled schedule as text: ``jit_phases.inc`` includes the kernel's phase text once per phase, with
    ``jit_ops_p<k>.inc`` -- one inclusion of qb2_pass_ops.inc per decode point -- as that phase's ops."""
    sched = decode_schedule(small)
    files = {}
    top = []
    for p, pts in enumerate(sched):
        top += ["{", f"    constexpr uint32_t p = QB2_DEP({p}u);", "#undef QB2_OPS_FILE", f'#define QB2_OPS_FILE "jit_ops_p{p}.inc"',
                "#include QB2_PHASE_FILE", "}"]
        body = []
        for o in pts:
            body += ["{", f"    constexpr uint32_t o = QB2_DEP({o}u);", '#include "qb2_pass_ops.inc"', "}"]
        files[f"jit_ops_p{p}.inc"] = "\n".join(body) + "\n"
    files["jit_phases.inc"] = "\n".join(top) + "\n"
    return files


def pass_key(small, tma):
    """Cache key: the kernel depends on the pass's STRUCTURE only (geometry, phase and op records, slot maps,
    decode schedule, kernel kind); coefficients and tables are read from the pass at run time, so passes that
    differ in angles share a kernel."""
    T, r, dtype, _nseg = struct.unpack_from("<BBBB", small, 16)
    (n_phases, phases_off, ops_off, _coef_off, u16_off) = struct.unpack_from("<IIIII", small, 20)
    n_ops, features = struct.unpack_from("<II", small, 52)
    h = hashlib.sha256(_sources_digest())
    h.update(struct.pack("<IIIIIII", 1 if tma else 0, T, r, dtype, n_phases, n_ops, features & FEATURE_FULL))
    h.update(small[phases_off : phases_off + 32 * n_phases])
    h.update(small[ops_off : ops_off + 32 * n_ops])
    h.update(small[u16_off:])
    return h.hexdigest()[:32]


def build_cubin(small, tma, key=None, keep_source=False):
    """Compile the specialised kernel of one pass; returns the cubin path (cached) or None without nvcc."""
    key = pass_key(small, tma) if key is None else key
    path = os.path.join(cache_dir(), key + ".cubin")
    if os.path.exists(path):
        return path
    nvcc = find_nvcc()
    if nvcc is None:
        return None
    with tempfile.TemporaryDirectory(dir=cache_dir()) as tmp:
        src = os.path.join(tmp, "pass.cu")
        with open(src, "w") as f:
            f.write(kernel_source(small, tma))
        for name, text in schedule_files(small).items():
            with open(os.path.join(tmp, name), "w") as f:
                f.write(text)
        out = os.path.join(tmp, "pass.cubin")
        cmd = [nvcc, "-cubin", "-O3", "-std=c++17", "-arch=sm_100a", "-lineinfo", "-I", tmp, "-I", INCLUDE, "-I", CSRC, "-o", out, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise _lib.Qb2Error("specialised pass kernel failed to compile:\n" + res.stderr[-4000:])
        if keep_source:
            shutil.copy(src, os.path.join(cache_dir(), key + ".cu"))
        os.replace(out, path)
    return path


def kernel_for(step, build=True):
    """The specialised kernel of a planned pass (``PassStep``) as an integer handle, or None (no kernel geometry
    for it, no nvcc, or ``build`` False and nothing cached in this process or on disk)."""
    small = small_section(step.blob)
    tma = bool(getattr(step, "tma", False)) and os.environ.get("QB2_NO_TMA") is None
    if not tma:
        T, r, dtype, _nseg = struct.unpack_from("<BBBB", small, 16)
        if _reg_geometry(dtype, r, T - r) is None:
            return None
    key = pass_key(small, tma)
    with _lock:
        if key in _kernels:
            return _kernels[key]
    path = os.path.join(cache_dir(), key + ".cubin")
    handle = None
    try:
        if not os.path.exists(path):
            if not build:
                return None
            path = build_cubin(small, tma, key)
        if path is not None:
            k = ctypes.c_void_p()
            _lib.check(_lib.load().qb2_jit_load(path.encode(), ctypes.byref(k)), "qb2_jit_load")
            handle = k.value
    except Exception as exc:  # noqa: BLE001 -- a kernel that cannot be built or loaded is not an error: the interpreter runs
        import warnings

        warnings.warn(f"qrisp_b200.jit: no specialised kernel for this pass ({str(exc)[:300]}); the interpreter runs it")
        handle = None
    with _lock:
        _kernels[key] = handle
    return handle


def kernels_for(steps, workers=None):
    """Specialised kernels for a list of passes, compiling what is missing in parallel (one nvcc process per
    distinct structure, ``workers`` at a time: default = the host's cores, at most 16).  Returns the handles in
    the order of ``steps`` (None where a pass has none)."""
    from concurrent.futures import ThreadPoolExecutor

    todo = {}
    for st in steps:
        small = small_section(st.blob)
        tma = bool(getattr(st, "tma", False)) and os.environ.get("QB2_NO_TMA") is None
        if not tma:
            T, r, dtype, _nseg = struct.unpack_from("<BBBB", small, 16)
            if _reg_geometry(dtype, r, T - r) is None:
                continue
        key = pass_key(small, tma)
        with _lock:
            known = key in _kernels
        if not known and not os.path.exists(os.path.join(cache_dir(), key + ".cubin")):
            todo.setdefault(key, (small, tma))
    if todo and find_nvcc() is not None:
        workers = workers or max(1, min(16, os.cpu_count() or 1))
        def build(kv):
            try:
                build_cubin(kv[1][0], kv[1][1], kv[0])
            except Exception:  # noqa: BLE001 -- reported by kernel_for below, once per pass
                pass

        with ThreadPoolExecutor(max_workers=workers) as pool:
            list(pool.map(build, todo.items()))
    return [kernel_for(st) for st in steps]


def prebuild(instructions, num_qubits, world=1, dtype="complex64", workers=None):
    """Compile, WITHOUT a device, the specialised kernels of the passes ``world`` ranks would plan for the
    unitary ``instructions`` on ``num_qubits`` qubits, into the cache (nvcc cross-compiles).  Returns
    ``(passes, cubins present afterwards)``.  ``__graft_entry__.build()`` does this for the bench workload."""
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np

    from .distributed import PlanOnlyShard

    steps, _entry, _pos = PlanOnlyShard(num_qubits, world, dtype=np.dtype(dtype)).plan(list(instructions))
    todo = {}
    for st in steps:
        if st.kind != "pass":
            continue
        small = small_section(st.blob)
        tma = bool(getattr(st, "tma", False)) and os.environ.get("QB2_NO_TMA") is None
        if not tma:
            T, r, dt, _nseg = struct.unpack_from("<BBBB", small, 16)
            if _reg_geometry(dt, r, T - r) is None:
                continue
        todo.setdefault(pass_key(small, tma), (small, tma))
    workers = workers or max(1, min(16, os.cpu_count() or 1))
    with ThreadPoolExecutor(max_workers=workers) as pool:
        paths = list(pool.map(lambda kv: build_cubin(kv[1][0], kv[1][1], kv[0]), todo.items()))
    return len(todo), sum(1 for p in paths if p is not None)

