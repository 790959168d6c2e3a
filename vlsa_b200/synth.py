"""Seeded synthetic bags and parameters.

Everything here runs on the CPU with an explicit ``torch.Generator`` so that the
golden-vector script (run where /root/reference exists), the CPU tests and the
GPU tests all regenerate bit-identical inputs from a seed instead of shipping
multi-megabyte fixtures.

G1 "CONCH-like" reproduces the statistics of the reference's shipped bag
(assert/blca-test-WSI-TCGA-XF-A9ST.pt: row norm ~25, pairwise cosine ~0.7, low
rank); G0 "randn stress" is the adversarial case for parity (pooled vector almost
cancels).
"""
from __future__ import annotations

import math

import torch

BASE_SEED = 1234
D_FEAT = 512


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def bag_g0(n: int, seed: int, d: int = D_FEAT) -> torch.Tensor:
    """G0: X = 1.1 * randn(N, D) (element std of the real bag)."""
    g = _gen(seed)
    return 1.1 * torch.randn(n, d, generator=g, dtype=torch.float32)


def bag_g1(n: int, seed: int, d: int = D_FEAT, rank: int = 64, target_norm: float = 25.0) -> torch.Tensor:
    """G1: X = m + c * randn(N, rank) @ Bm + 0.15 * randn(N, D), mean row norm = 25."""
    g = _gen(seed)
    m = torch.randn(d, generator=g, dtype=torch.float32)
    m = 21.0 * m / m.norm()
    decay = torch.tensor([0.85 ** k for k in range(rank)], dtype=torch.float32)
    bm = torch.randn(rank, d, generator=g, dtype=torch.float32) * decay[:, None]
    z = torch.randn(n, rank, generator=g, dtype=torch.float32)
    noise = 0.15 * torch.randn(n, d, generator=g, dtype=torch.float32)
    low = z @ bm
    # bisection on c so that the mean row norm hits target_norm (uses <=4096 rows)
    probe = slice(0, min(n, 4096))
    lo, hi = 0.0, 4.0
    for _ in range(40):
        c = 0.5 * (lo + hi)
        rn = (m + c * low[probe] + noise[probe]).norm(dim=-1).mean().item()
        if rn < target_norm:
            lo = c
        else:
            hi = c
    c = 0.5 * (lo + hi)
    return (m + c * low + noise).contiguous()


def make_bag(kind: str, n: int, seed: int, d: int = D_FEAT) -> torch.Tensor:
    if kind == "g0":
        return bag_g0(n, seed, d)
    if kind == "g1":
        return bag_g1(n, seed, d)
    raise ValueError(f"unknown bag kind {kind!r}")


def make_params(p: int, r: int, seed: int, d: int = D_FEAT, w: torch.Tensor | None = None,
                b: torch.Tensor | None = None) -> dict:
    """Parameters of one VLSA instance.

    prompt_features = unit(randn(P, D)); residual = randn(P, D) (prompt_adapter.py:94);
    T = randn(R, D); W, b = given (shipped checkpoint) or nn.Linear-style uniform init;
    logit_scale = 4.0309 (checkpoint value).
    """
    g = _gen(seed)
    pf = torch.randn(p, d, generator=g, dtype=torch.float32)
    pf = pf / pf.norm(dim=-1, keepdim=True)
    res = torch.randn(p, d, generator=g, dtype=torch.float32)
    t = torch.randn(r, d, generator=g, dtype=torch.float32)
    if w is None:
        bound = 1.0 / math.sqrt(d)
        w = (torch.rand(d, d, generator=g, dtype=torch.float32) * 2 - 1) * bound
        b = (torch.rand(d, generator=g, dtype=torch.float32) * 2 - 1) * bound
    return {
        "prompt_features": pf,
        "residual_features": res,
        "res_ratio": 0.5,
        "text_features": t,
        "W": w.clone(),
        "b": b.clone(),
        "logit_scale": torch.tensor(4.0309, dtype=torch.float32),
    }


def make_labels(bsz: int, r: int, seed: int) -> tuple[torch.Tensor, torch.Tensor]:
    """t ~ U{0..R-1}, e ~ Bernoulli(0.45) (BLCA event ratio)."""
    g = _gen(seed)
    t = torch.randint(0, r, (bsz,), generator=g, dtype=torch.int64)
    e = (torch.rand(bsz, generator=g) < 0.45).to(torch.int64)
    return t, e
