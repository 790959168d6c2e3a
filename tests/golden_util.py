"""Rebuild the inputs of a golden case from its seed (same code path as make_golden.py)."""
import os

import numpy as np
import torch

from vlsa_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def _ckpt():
    if "ck" not in _cache:
        z = np.load(os.path.join(GOLDEN, "blca_ckpt_params.npz"))
        _cache["ck"] = {k: torch.from_numpy(z[k].copy()) for k in z.files}
    return _cache["ck"]


def _real():
    if "real" not in _cache:
        _cache["real"] = torch.from_numpy(np.load(os.path.join(GOLDEN, "blca_bag_A9ST.npy")))
    return _cache["real"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def rebuild_inputs(name, case):
    """-> (bags: list[Tensor [N,D]], params dict, t, e).  Asserts the regenerated X matches the
    checksum stored by make_golden.py so a generator drift is reported as such."""
    ck = _ckpt()
    W, b = ck["W"], ck["b"]
    seed = int(case["seed"])
    ns = [int(v) for v in np.atleast_1d(case["n"])]
    kind = str(case["kind"])
    if name.startswith("single_"):
        P, R = int(case["P"]), int(case["R"])
        bags = [synth.make_bag(kind, ns[0], seed)]
        params = synth.make_params(P, R, seed + 100000, w=W, b=b)
    elif name.startswith("real_"):
        P, R = int(case["P"]), int(case["R"])
        bags = [_real()]
        params = synth.make_params(P, R, seed, w=W, b=b)
        if P == 12:
            params["residual_features"] = ck["residual_features"].clone()
        params["logit_scale"] = ck["logit_scale"].clone()
    elif name.startswith("batch_"):
        P, R = int(case["P"]), int(case["R"])
        bags = [synth.make_bag(kind, n, seed + i) for i, n in enumerate(ns)]
        params = synth.make_params(P, R, seed + 100000, w=W, b=b)
    elif name.startswith("zeroshot_"):
        R = int(case["R"])
        bags = [synth.make_bag(kind, ns[0], seed)]
        params = synth.make_params(1, R, seed + 100000, w=W, b=b)
    else:
        raise ValueError(name)
    xs = np.atleast_1d(case["x_sum"])
    for x, s in zip(bags, xs):
        got = x.double().sum().item()
        assert abs(got - float(s)) <= 1e-9 * max(1.0, abs(float(s))), \
            f"{name}: synthetic generator drifted (sum {got} vs golden {float(s)})"
    t = torch.from_numpy(case["t"].copy()) if "t" in case else None
    e = torch.from_numpy(case["e"].copy()) if "e" in case else None
    return bags, params, t, e
