"""TEST INFRASTRUCTURE — ctypes binding of oracle/_build/libvlsa_oracle.so (the plain-C fp64 restatement)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libvlsa_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(HERE, "vlsa_oracle.c")):
            subprocess.run(["make", "-C", HERE, "-s"], check=True)
        _lib = C.CDLL(SO)
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def forward(X, Q, W, b, T, scale, logit_scale, want_attn=False):
    X, Q, W, b, T = _f(X), _f(Q), _f(W), _f(b), _f(T)
    N, P, R = X.shape[0], Q.shape[0], T.shape[0]
    f, g = np.zeros(512), np.zeros(512)
    logits, inc = np.zeros(R), np.zeros(R)
    A = np.zeros((P, max(N, 1))) if want_attn else None
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib().vlsa_oracle_forward(p(X), C.c_int64(N), p(Q), P, p(W), p(b), p(T), R, C.c_double(scale),
                                   C.c_double(logit_scale), p(f), p(g), p(logits), p(inc), p(A) if want_attn else None)
    assert rc == 0
    return f, g, logits, inc, A


def losses(p, t, e, ls_exp, alpha=0.0, eps=1e-7):
    p = np.ascontiguousarray(p, dtype=np.float64)
    t = np.ascontiguousarray(t, dtype=np.int64)
    e = np.ascontiguousarray(e, dtype=np.int64)
    out = np.zeros(2)
    q = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib().vlsa_oracle_losses(q(p), q(t), q(e), p.shape[0], p.shape[1], C.c_double(ls_exp), C.c_double(alpha),
                                  C.c_double(eps), q(out))
    assert rc == 0
    return out


def logit_pool(X, T, logit_scale, mode, k):
    X, T = _f(X), _f(T)
    pooled = np.zeros(T.shape[0])
    pred = np.zeros(1, dtype=np.int64)
    q = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib().vlsa_oracle_logit_pool(q(X), C.c_int64(X.shape[0]), q(T), T.shape[0], C.c_double(logit_scale), mode, k,
                                      q(pooled), q(pred))
    assert rc == 0
    return pooled, int(pred[0])
