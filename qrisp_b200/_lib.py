"""ctypes binding of libqb2.so (include/qrisp_b200.h).

The product path has no CPU fallback: if the library is missing, or a call fails, this
module raises.  Build the library with ``make`` (or ``python -c "import __graft_entry__ as g;
g.build()"``) at the repository root.
"""

from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libqb2.so")

QB2_ABI_VERSION = 10
QB2_BLOCK_TC_BYTES = 32768  # include/qrisp_b200.h
QB2_C64, QB2_C128 = 0, 1
QB2_NO_BASIS = (1 << 64) - 1

# every exported symbol of include/qrisp_b200.h: name -> (restype, argtypes)
_c = ctypes
SYMBOLS = {
    "qb2_abi_version": (_c.c_int, []),
    "qb2_last_error": (_c.c_char_p, []),
    "qb2_device_info": (_c.c_int, [_c.POINTER(_c.c_int), _c.POINTER(_c.c_int), _c.POINTER(_c.c_size_t)]),
    "qb2_launch_count": (_c.c_uint64, []),
    "qb2_reset_launch_count": (None, []),
    "qb2_malloc": (_c.c_int, [_c.POINTER(_c.c_void_p), _c.c_size_t]),
    "qb2_free": (_c.c_int, [_c.c_void_p]),
    "qb2_memcpy_h2d": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "qb2_memcpy_d2h": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "qb2_stream_sync": (_c.c_int, [_c.c_void_p]),
    "qb2_init_basis": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_uint64, _c.c_int, _c.c_void_p]),
    "qb2_init_sparse": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_uint32, _c.c_int, _c.c_void_p]),
    "qb2_sparse_prefix": (
        _c.c_int,
        [_c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_uint32, _c.c_double, _c.c_void_p, _c.c_void_p,
         _c.POINTER(_c.c_uint32), _c.POINTER(_c.c_int32)],
    ),
    "qb2_run_pass": (
        _c.c_int,
        [_c.c_void_p, _c.c_int, _c.c_uint64, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_uint32],
    ),
    "qb2_run_pass_ex": (
        _c.c_int,
        [_c.c_void_p, _c.c_int, _c.c_uint64, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_uint32,
         _c.c_void_p, _c.c_void_p],
    ),
    "qb2_sample_scan": (_c.c_int, [_c.c_void_p, _c.c_uint64, _c.c_void_p]),
    "qb2_sample_draw_chunks": (
        _c.c_int,
        [_c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_int, _c.c_void_p],
    ),
    "qb2_jit_load": (_c.c_int, [_c.c_char_p, _c.POINTER(_c.c_void_p)]),
    "qb2_run_pass_jit": (
        _c.c_int,
        [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_uint64, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_uint32],
    ),
    "qb2_apply_matrix_big": (
        _c.c_int,
        [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_uint64, _c.c_void_p, _c.POINTER(_c.c_int), _c.c_int, _c.c_uint64,
         _c.c_uint64, _c.c_int, _c.c_void_p],
    ),
    "qb2_apply_block_tc": (
        _c.c_int,
        [_c.c_void_p, _c.c_int, _c.c_uint64, _c.c_void_p, _c.POINTER(_c.c_int), _c.c_uint64, _c.c_uint64, _c.c_void_p],
    ),
    "qb2_pack_block_tc": (_c.c_int, [_c.c_void_p, _c.c_void_p]),
    "qb2_apply_mux_ry": (
        _c.c_int,
        [_c.c_void_p, _c.c_int, _c.c_uint64, _c.c_int, _c.POINTER(_c.c_int), _c.c_int, _c.c_void_p, _c.c_uint64, _c.c_uint64,
         _c.c_int, _c.c_void_p],
    ),
    "qb2_swap_pull": (
        _c.c_int,
        [_c.c_void_p, _c.POINTER(_c.c_void_p), _c.c_int, _c.c_int, _c.c_int, _c.POINTER(_c.c_int), _c.c_int, _c.c_void_p],
    ),
    "qb2_permute_bits": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.POINTER(_c.c_int), _c.c_int, _c.c_void_p]),
    "qb2_norm2": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_void_p, _c.c_int, _c.c_void_p]),
    "qb2_probabilities": (
        _c.c_int,
        [_c.c_void_p, _c.c_int, _c.c_uint64, _c.POINTER(_c.c_int), _c.c_int, _c.c_void_p, _c.c_int, _c.c_void_p],
    ),
    "qb2_prune_probabilities": (
        _c.c_int,
        [_c.c_void_p, _c.c_int, _c.c_double, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_uint64, _c.c_void_p],
    ),
    "qb2_sample_chunk_bits": (_c.c_int, [_c.c_int]),
    "qb2_sample_prepare": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_void_p, _c.c_int, _c.c_void_p]),
    "qb2_sample_draw": (
        _c.c_int,
        [_c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_int, _c.c_void_p],
    ),
}


class Qb2Error(RuntimeError):
    pass


_lib = None


def load():
    """Load libqb2.so and declare every prototype.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Qb2Error(
            f"{LIB_PATH} is missing: the B200 engine is not built (run `make` at the repo root). "
            "There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.qb2_abi_version() != QB2_ABI_VERSION:
        raise Qb2Error(f"libqb2.so has ABI {lib.qb2_abi_version()}, the host expects {QB2_ABI_VERSION}")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().qb2_last_error().decode(errors="replace")
        raise Qb2Error(f"{what} failed with code {rc}: {msg}")
