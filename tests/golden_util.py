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


def rebuild_interp(case):
    """Inputs of an `interp_*` golden (tests/golden/make_golden_r02.py::interp_inputs): (X [N,512], params)."""
    ck = _ckpt()
    P, R, seed, kind = int(case["P"]), int(case["R"]), int(case["seed"]), str(case["kind"])
    params = synth.make_params(P, R, seed, w=ck["W"], b=ck["b"])
    if kind == "real":
        X = _real()
        if P == 12:
            params["residual_features"] = ck["residual_features"].clone()
        params["logit_scale"] = ck["logit_scale"].clone()
    else:
        X = synth.make_bag(kind, int(np.atleast_1d(case["n"])[0]), seed + 7)
    got = X.double().sum().item()
    assert abs(got - float(case["x_sum"])) <= 1e-9 * max(1.0, abs(float(case["x_sum"]))), "synthetic generator drifted"
    return X, params


def rebuild_zeroshot2(case):
    """Inputs of a `zeroshot2_*` golden: (X [N,512], params with text_features / logit_scale)."""
    ck = _ckpt()
    R, seed, n = int(case["R"]), int(case["seed"]), int(np.atleast_1d(case["n"])[0])
    X = synth.make_bag(str(case["kind"]), n, seed)
    got = X.double().sum().item()
    assert abs(got - float(case["x_sum"])) <= 1e-9 * max(1.0, abs(float(case["x_sum"]))), "synthetic generator drifted"
    return X, synth.make_params(1, R, seed + 100000, w=ck["W"], b=ck["b"])


# ---- VLFAN variants: cases and seeded inputs shared by the generator and the tests -----------------
VARIANT_CASES = [
    dict(P=4, gated=True, pooling="mean", pred_head="default", hid=32, kind="g1", ns=(1000, 37), seed=31001),
    dict(P=12, gated=True, pooling="mean", pred_head="default", hid=32, kind="g0", ns=(2798,), seed=31002),
    dict(P=4, gated=False, pooling="max", pred_head="default", hid=32, kind="g1", ns=(1000, 513), seed=31003),
    dict(P=12, gated=False, pooling="weight", pred_head="default", hid=32, kind="g1", ns=(300, 2000, 64), seed=31004),
    dict(P=8, gated=False, pooling="attention", pred_head="default", hid=32, kind="g1", ns=(1000, 129), seed=31005),
    dict(P=6, gated=False, pooling="gated_attention", pred_head="default", hid=32, kind="g0", ns=(700,), seed=31006),
    dict(P=4, gated=False, pooling="mean", pred_head="Identity", hid=32, kind="g1", ns=(1000,), seed=31007),
    dict(P=3, gated=True, pooling="attention", pred_head="Identity", hid=32, kind="g1", ns=(257, 1), seed=31008),
    dict(P=16, gated=True, pooling="max", pred_head="default", hid=32, kind="g1", ns=(1500, 33), seed=31009),
    dict(P=4, gated=False, pooling="mean", pred_head="default", hid=32, kind="g1", ns=(1000, 37), seed=31010, feat_proj=True),
    dict(P=12, gated=True, pooling="mean", pred_head="default", hid=32, kind="g0", ns=(700, 1), seed=31011, feat_proj=True),
    dict(P=6, gated=False, pooling="attention", pred_head="default", hid=32, kind="g1", ns=(513,), seed=31012, feat_proj=True),
]


def variant_name(case):
    name = "variant_P{P}_{g}_{pooling}_{h}_{kind}".format(g="gated" if case["gated"] else "plain",
                                                          h="id" if case["pred_head"] == "Identity" else "lin", **case)
    return name + ("_proj" if case.get("feat_proj") else "")


def variant_inputs(case):
    """Seeded inputs of a variant case: bags, the queries Q [P (+1), 512] (prototype directions + 0.5 * randn, like the
    TaskRes query), the pooling parameters, the gradient seeds G [1,512] per bag; W, b from the shipped checkpoint."""
    ck = _ckpt()
    seed, P, hid = case["seed"], case["P"], case["hid"]
    g = torch.Generator().manual_seed(seed)
    nq = P + 1 if case["gated"] else P
    proto = torch.nn.functional.normalize(torch.randn(nq, 512, generator=g), dim=-1)
    Q = (0.5 * torch.randn(nq, 512, generator=g) + proto).contiguous()
    pool = {}
    if case["pooling"] == "weight":
        pool["weight"] = torch.randn(1, P, generator=g)
    elif case["pooling"] == "attention":
        pool = {"attention.0.weight": torch.randn(hid, 512, generator=g) / 512 ** 0.5,
                "attention.0.bias": 0.1 * torch.randn(hid, generator=g),
                "attention.2.weight": torch.randn(1, hid, generator=g) / hid ** 0.5,
                "attention.2.bias": 0.1 * torch.randn(1, generator=g)}
    elif case["pooling"] == "gated_attention":
        pool = {"fc1.0.weight": torch.randn(hid, 512, generator=g) / 512 ** 0.5, "fc1.0.bias": 0.1 * torch.randn(hid, generator=g),
                "score.0.weight": torch.randn(hid, 512, generator=g) / 512 ** 0.5, "score.0.bias": 0.1 * torch.randn(hid, generator=g),
                "fc2.weight": torch.randn(1, hid, generator=g) / hid ** 0.5, "fc2.bias": 0.1 * torch.randn(1, generator=g)}
    bags = [synth.make_bag(case["kind"], n, seed + 10 + i) for i, n in enumerate(case["ns"])]
    G = [torch.randn(1, 512, generator=g) for _ in bags]
    proj = None
    if case.get("feat_proj"):               # Feat_Projecter (model/layers.py:65-82): Linear near the identity + LayerNorm
        proj = {"projecter.0.weight": torch.eye(512) + 0.3 * torch.randn(512, 512, generator=g) / 512 ** 0.5,
                "projecter.0.bias": 0.1 * torch.randn(512, generator=g),
                "projecter.1.weight": 1.0 + 0.1 * torch.randn(512, generator=g),
                "projecter.1.bias": 0.1 * torch.randn(512, generator=g)}
    return {"Q": Q, "pool": pool, "bags": bags, "G": G, "W": ck["W"], "b": ck["b"], "proj": proj}
