"""Time the variant path (ops.pooled forward + per-prototype-gradient backward) next to the shipped path at the
headline shape (32 bags x 50k patches, fp32).  Writes gpurun_out/variant_time.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vlsa_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, N = 32, 50000
X = torch.randn(B * N, 512, device=dev) * 1.1 + 0.9
plan = ops.make_plan([N] * B, dev)
gbytes = B * N * 2048 / 1e9
out = {}


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


for P in (4, 8, 12):
    g = torch.Generator().manual_seed(P)
    Q = torch.randn(P, 512, generator=g).to(dev).requires_grad_(True)
    W = (torch.randn(512, 512, generator=g) / 22).to(dev).requires_grad_(True)
    bias = torch.zeros(512, device=dev, requires_grad=True)
    dO = torch.randn(B, P, 512, generator=g).to(dev)
    df = torch.randn(B, 512, generator=g).to(dev)

    def fwd_pooled():
        return ops.pooled(X, plan, Q, False)

    def step_pooled():
        O, _ = ops.pooled(X, plan, Q, False)
        O.backward(dO)

    def step_mean():
        f, _ = ops.encode(X, plan, Q, W, bias)
        f.backward(df)

    def step_pooled_dx():                       # the rows themselves need a gradient (feat_proj in front)
        O, _ = ops.pooled(X, plan, Q, False)
        torch.autograd.grad(O, (X, Q), dO)      # no accumulation into .grad: the kernels alone

    tf, ts, tm = timed(fwd_pooled), timed(step_pooled), timed(step_mean)
    X.requires_grad_(True)
    tx = timed(step_pooled_dx, iters=4)
    X.requires_grad_(False)
    out[f"P{P}"] = {"pooled_fwd_ms": tf, "pooled_fwd_bwd_ms": ts, "pooled_bwd_ms": ts - tf,
                    "pooled_bwd_GBps": gbytes / ((ts - tf) * 1e-3), "mean_fwd_bwd_ms": tm,
                    "dx_ms": tx - ts, "dx_GBps_read_plus_write": 2 * gbytes / ((tx - ts) * 1e-3)}
    print(P, out[f"P{P}"])
group = os.environ.get("VLSA_GEN_GROUP", "default")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/variant_time_group_{group}.json", "w"), indent=1)
