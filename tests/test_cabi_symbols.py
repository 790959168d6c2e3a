"""CPU: the C-ABI library builds, loads and exports every symbol include/vlsa_b200.h declares; the host-only
planner behaves.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from vlsa_b200 import _lib, build
    build.build_library()
    return _lib.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vlsa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vlsa_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_something():
    syms = declared_symbols()
    assert "vlsa_agg_fwd" in syms and "vlsa_agg_bwd" in syms and "vlsa_surv_loss_fwd_bwd" in syms
    assert len(syms) >= 10


def test_every_declared_symbol_is_exported(lib):
    from vlsa_b200 import _lib
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/vlsa_b200.h but not exported"
        assert name in _lib._SIGNATURES, f"{name} has no ctypes signature in vlsa_b200/_lib.py"
    for name in _lib._SIGNATURES:
        assert name in declared_symbols(), f"{name} bound in _lib.py but not declared in the header"


def test_version_and_error_strings(lib):
    assert lib.vlsa_version() >= 100
    assert lib.vlsa_error_string(0) == b"success"
    assert b"invalid" in lib.vlsa_error_string(-1)
    assert b"workspace" in lib.vlsa_error_string(-2)


def _plan(lib, sizes, sms=148):
    cu = np.zeros(len(sizes) + 1, dtype=np.int64)
    np.cumsum(np.asarray(sizes, dtype=np.int64), out=cu[1:])
    cs = np.zeros(len(sizes) + 1, dtype=np.int32)
    rows = C.c_int(0)
    rc = lib.vlsa_agg_plan(cu.ctypes.data_as(C.POINTER(C.c_int64)), len(sizes), sms, C.byref(rows),
                           cs.ctypes.data_as(C.POINTER(C.c_int32)))
    return rc, rows.value, cs


@pytest.mark.parametrize("sizes", [[50000] * 32, [1], [0], [0, 5, 0], [2798], [1000, 37, 2798, 1, 513, 4096, 255, 1500],
                                   [100000], [10] * 1000, []])
def test_plan_covers_every_row_once(lib, sizes):
    rc, rows, cs = _plan(lib, sizes)
    assert rc == 0
    assert rows > 0 and rows % 16 == 0
    assert cs[0] == 0 and (np.diff(cs) >= 0).all()
    for n, c in zip(sizes, np.diff(cs)):
        assert c == -(-n // rows)                    # ceil(n / chunk_rows) chunks, none for an empty bag
    if sum(sizes) >= 148 * 16 * 8:
        assert cs[-1] >= 148                          # every SM gets work


def test_plan_rejects_bad_input(lib):
    cu = np.array([0, 5, 3], dtype=np.int64)          # negative bag size
    cs = np.zeros(3, dtype=np.int32)
    rows = C.c_int(0)
    rc = lib.vlsa_agg_plan(cu.ctypes.data_as(C.POINTER(C.c_int64)), 2, 148, C.byref(rows),
                           cs.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == -1
    assert lib.vlsa_agg_plan(None, 2, 148, C.byref(rows), cs.ctypes.data_as(C.POINTER(C.c_int32))) == -1


def test_workspace_query_grows_with_chunks(lib):
    a = lib.vlsa_agg_workspace_bytes(100, 4, 4)
    b = lib.vlsa_agg_workspace_bytes(200, 4, 4)
    c = lib.vlsa_agg_workspace_bytes(100, 4, 16)
    assert 0 < a < b and a < c
    assert lib.vlsa_agg_workspace_bytes(100, 4, 17) == 0
    assert lib.vlsa_logit_pool_workspace_bytes(1000, 4, 10) >= 1000 * 4 * 4


def test_null_arguments_are_rejected_without_a_gpu(lib):
    # argument validation happens before any CUDA call
    assert lib.vlsa_agg_fwd(None, 0, 16, None, None, 1, 16, 1, None, 4, 0, 100.0, None, None, None, 4, None, None, 0,
                            None, None, None, None, None, None, None, None, None) == -1
    assert lib.vlsa_agg_pooled_fwd(None, 0, 32, None, None, 1, 32, 1, None, 4, 0, 100.0, None, 0, None, None, None) == -1
    assert lib.vlsa_agg_pooled_bwd(None, 0, 32, None, None, 1, 32, 1, None, 4, 0, 100.0, None, None, None, None, 0, None,
                                   None) == -1
    assert lib.vlsa_agg_pooled_bwd_dx(None, None, 1, 10, None, 4, 0, 100.0, None, None, None, None, None) == -1
    assert lib.vlsa_surv_loss_fwd_bwd(None, None, None, 1, 4, None, 1.0, 1.0, 0.0, 1e-7, 1.0, 0, None, None, None,
                                      None, None) == -1
    assert lib.vlsa_logit_pool_fwd(None, 0, 10, None, 4, None, 1, 10, None, 0, None, None, None) == -1
    assert lib.vlsa_feat_pool_fwd(None, 0, 10, 0, None, 4, None, None, 0, None, None, None, None, None) == -1
    assert lib.vlsa_row_normalize(None, 0, 10, None, None) == -1
    assert lib.vlsa_feat_pool_workspace_bytes() >= 296 * 512 * 4
    # kernel-selection bits other than the documented ones are refused
    assert lib.vlsa_agg_pooled_fwd(None, 0x800, 32, None, None, 1, 32, 1, None, 4, 0, 100.0, None, 0, None, None, None) in (-1, -3)
