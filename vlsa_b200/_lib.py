"""ctypes binding of libvlsa_b200.so (include/vlsa_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` / ``vlsa_b200.build``.
There is NO fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VLSA_B200_LIB") or os.path.join(_HERE, "lib", "libvlsa_b200.so")

_lib = None

c_f32p = C.c_void_p
c_i64p = C.c_void_p
c_i32p = C.c_void_p

_SIGNATURES = {
    "vlsa_version": (C.c_int, []),
    "vlsa_error_string": (C.c_char_p, [C.c_int]),
    "vlsa_agg_plan": (C.c_int, [C.POINTER(C.c_int64), C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int32)]),
    "vlsa_agg_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "vlsa_split16_row_bytes": (C.c_size_t, []),
    "vlsa_split16_pack": (C.c_int, [c_f32p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "vlsa_agg_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_i64p, c_i32p, C.c_int, C.c_int, C.c_int, c_f32p, C.c_int,
                               C.c_int, C.c_float, c_f32p, c_f32p, c_f32p, C.c_int, c_f32p, C.c_void_p, C.c_size_t,
                               c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
    "vlsa_agg_partial_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_i64p, c_i32p, C.c_int, C.c_int, C.c_int, c_f32p, C.c_int,
                                       C.c_float, C.c_void_p, C.c_size_t, C.c_void_p]),
    "vlsa_agg_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_i64p, c_i32p, C.c_int, C.c_int, C.c_int, c_f32p, C.c_int,
                               C.c_int, C.c_float, c_f32p, c_f32p, C.c_int, c_f32p,
                               c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p,
                               c_f32p, c_f32p, c_f32p,
                               C.c_void_p, C.c_size_t,
                               c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
    "vlsa_agg_pooled_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_i64p, c_i32p, C.c_int, C.c_int, C.c_int, c_f32p, C.c_int,
                                      C.c_int, C.c_float, C.c_void_p, C.c_size_t, c_f32p, c_f32p, C.c_void_p]),
    "vlsa_agg_pooled_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_i64p, c_i32p, C.c_int, C.c_int, C.c_int, c_f32p, C.c_int,
                                      C.c_int, C.c_float, c_f32p, c_f32p, c_f32p, C.c_void_p, C.c_size_t, c_f32p,
                                      C.c_void_p]),
    "vlsa_agg_pooled_bwd_dx": (C.c_int, [c_f32p, c_i64p, C.c_int, C.c_int64, c_f32p, C.c_int, C.c_int, C.c_float, c_f32p,
                                         c_f32p, c_f32p, c_f32p, C.c_void_p]),
    "vlsa_attn_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_f32p, C.c_int, C.c_int, C.c_float, c_f32p, c_f32p,
                                C.c_void_p]),
    "vlsa_interp_fwd": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, C.c_int, c_f32p, c_f32p, C.c_int, C.c_int, c_f32p,
                                  c_f32p, c_f32p, C.c_void_p]),
    "vlsa_surv_loss_fwd_bwd": (C.c_int, [c_f32p, c_i64p, c_i64p, C.c_int, C.c_int, c_f32p, C.c_float, C.c_float,
                                         C.c_float, C.c_float, C.c_float, C.c_int, c_f32p, c_f32p, c_f32p, c_f32p,
                                         C.c_void_p]),
    "vlsa_logit_pool_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int, C.c_int]),
    "vlsa_logit_pool_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_f32p, C.c_int, c_f32p, C.c_int, C.c_int,
                                      C.c_void_p, C.c_size_t, c_f32p, c_i64p, C.c_void_p]),
    "vlsa_adam_segment_bytes": (C.c_size_t, []),
    "vlsa_adam_step": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_float,
                                 C.c_float, C.c_float, C.c_void_p]),
    "vlsa_forward_host_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int, C.c_int, C.c_int]),
    "vlsa_feat_pool_workspace_bytes": (C.c_size_t, []),
    "vlsa_feat_pool_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, c_f32p, C.c_int, c_f32p, C.c_void_p,
                                     C.c_size_t, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
    "vlsa_row_normalize": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_f32p, C.c_void_p]),
    "vlsa_forward_host": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_int, c_f32p, C.c_int, C.c_int, C.c_float,
                                    c_f32p, c_f32p, c_f32p, C.c_int, c_f32p, C.c_void_p, C.c_size_t,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}


class VlsaLibraryError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the C-ABI library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VlsaLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  vlsa_b200 has no CPU / PyTorch fallback.")
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError:
            continue          # symbol check is a test (tests/test_cabi_symbols.py), not an import-time failure
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().vlsa_error_string(code)
        raise VlsaLibraryError(f"{what} failed with code {code}: {msg.decode() if msg else '?'}")
