import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from vlsa_b200 import ops, synth
dev = torch.device("cuda:0")
P = R = 12
sizes = [2798, 1000, 37]
bags = [synth.make_bag("g1", n, 100 + i) for i, n in enumerate(sizes)]
pr = synth.make_params(P, R, 7)
X = torch.cat(bags, 0).to(dev); plan = ops.make_plan(sizes, dev)
Q = (pr["res_ratio"] * pr["residual_features"] + pr["prompt_features"]).to(dev)
W, b, T, ls = (pr[k].to(dev) for k in ("W", "b", "text_features", "logit_scale"))
# fp64 truth on the GPU
Qn = torch.nn.functional.normalize(Q.double(), dim=-1)
O64, A64 = [], []
for bi, n in enumerate(sizes):
    Xd = bags[bi].to(dev).double()
    S = ops.coattn_scale() * Qn @ torch.nn.functional.normalize(Xd, dim=-1).T
    A = torch.softmax(S, -1); A64.append(A); O64.append(A @ Xd)
print("chunk_rows", plan.chunk_rows, "chunks", plan.total_chunks)
for var in ("simt", "tc"):
    ops.set_agg_variant(var)
    o = ops.aggregate_forward_raw(X, plan, Q, W, b, T, ls, need_bwd=True)
    torch.cuda.synchronize()
    for bi in range(3):
        Oe = (o["O"][bi].double() - O64[bi]).abs().max().item() / O64[bi].abs().max().item()
        # per-prototype relative error in the direction that matters: |O - O64| / |O64| per row
        row = ((o["O"][bi].double() - O64[bi]).norm(dim=-1) / O64[bi].norm(dim=-1)).max().item()
        # ml consistency: A from (m, l) vs truth at the arg-max row
        m, l = o["ml"][bi, :, 0].double(), o["ml"][bi, :, 1].double()
        S = ops.coattn_scale() * Qn @ torch.nn.functional.normalize(bags[bi].to(dev).double(), dim=-1).T
        Arec = torch.exp(S - m[:, None]) / l[:, None]
        Ae = ((Arec - A64[bi]).abs().max(dim=1).values / A64[bi].max(dim=1).values).max().item()
        print(f"{var:4s} bag {bi}: O max rel err {Oe:.2e}  per-prototype |dO|/|O| {row:.2e}   A from (m,l) rel err {Ae:.2e}")
ops.set_agg_variant(None)
print("log-sum-exp per prototype: variant vs fp64")
for var in ("simt", "tc"):
    ops.set_agg_variant(var)
    o = ops.aggregate_forward_raw(X, plan, Q, W, b, T, ls, need_bwd=True)
    torch.cuda.synchronize()
    for bi in range(3):
        S = ops.coattn_scale() * Qn @ torch.nn.functional.normalize(bags[bi].to(dev).double(), dim=-1).T
        lse64 = torch.logsumexp(S, dim=-1)
        lse = o["ml"][bi, :, 0].double() + o["ml"][bi, :, 1].double().log()
        print(f"  {var:4s} bag {bi}: max |LSE - LSE64| {float((lse - lse64).abs().max()):.3e}  signed mean {float((lse - lse64).mean()):+.3e}   max score {float(S.max()):.2f}  m[0] {float(o['ml'][bi,0,0]):.4f} l[0] {float(o['ml'][bi,0,1]):.4e}")
ops.set_agg_variant(None)
