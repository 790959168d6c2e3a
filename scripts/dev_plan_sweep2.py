"""Dev helper (GPU box): streaming-kernel time against the chunk size of the plan, for ragged steps (TCGA-sized and 1k-100k bags):
how much of a small step is lost to the quantisation of chunks over the persistent CTAs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from vlsa_b200 import ops, synth
dev = torch.device("cuda:0")

def plan_with(sizes, t):
    sizes = np.asarray(sizes, dtype=np.int64)
    cu = np.zeros(len(sizes) + 1, dtype=np.int64); np.cumsum(sizes, out=cu[1:])
    rows = 32 * t
    cs = np.zeros(len(sizes) + 1, dtype=np.int32); cs[1:] = np.cumsum((sizes + rows - 1) // rows)
    return ops.BagPlan(cu, cs, rows, torch.from_numpy(cu).to(dev), torch.from_numpy(cs).to(dev))

def timeit(fn, iters=30):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

for name, lo, hi, seed in (("tcga 1k-20k", 1e3, 2e4, 0), ("tcga 1k-20k (b)", 1e3, 2e4, 1), ("ragged 1k-100k", 1e3, 1e5, 2), ("equal 10k", 1e4, 1e4, 3)):
    sizes = [int(v) for v in np.exp(np.random.default_rng(seed).uniform(np.log(lo), np.log(hi), 32))]
    tot = sum(sizes)
    Xs = [torch.randn(tot, 512, device=dev) * 1.1 + 0.7 for _ in range(3)]
    Xb = [x.to(torch.bfloat16) for x in Xs]
    default = ops.make_plan(sizes, dev)
    print(f"== {name}: {tot} rows, default chunk_rows {default.chunk_rows} ({default.total_chunks} chunks)")
    for P, data, tag in ((4, Xs, "fp32 P=4 simt"), (12, Xs, "fp32 P=12 tc"), (12, Xb, "bf16 P=12")):
        pr = synth.make_params(P, P, 1)
        Q = (0.5 * pr["residual_features"] + pr["prompt_features"]).to(dev)
        ws = ops._workspace(default, P, dev)
        base = timeit(lambda i: ops.aggregate_partial_only(data[i % 3], default, Q, ws))
        res = []
        for t in (8, 9, 10, 11, 12, 14, 16, 20, 24, 32, 48):
            pl = plan_with(sizes, t)
            ws2 = ops._workspace(pl, P, dev)
            us = timeit(lambda i: ops.aggregate_partial_only(data[i % 3], pl, Q, ws2))
            res.append(f"{t}:{pl.total_chunks}c {us:.0f}")
        print(f"   {tag:14s} default {base:.0f} us | " + "  ".join(res), flush=True)
    del Xs, Xb
