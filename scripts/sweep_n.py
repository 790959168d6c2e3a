"""BASELINE configs[1] and configs[3]: variable-N sweep (1k / 10k / 50k / 100k patches, batch 1 and 32) of the full
forward (VLSA.forward: aggregation + merge + adapter + head + incidence) and of the streaming kernel alone, fp32,
device-resident inputs rotating over >= 2 distinct batches (never re-timing the same bag out of L2 for B=1:
8 distinct bags).  Prints a markdown table.  usage: python scripts/sweep_n.py [P ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vlsa_b200 import ops, synth
dev = torch.device("cuda:0")
PEAK = 6544.0
DT = torch.bfloat16 if os.environ.get("DEV_DTYPE") == "bf16" else torch.float32
ES = 2 if DT == torch.bfloat16 else 4
Ps = [int(a) for a in sys.argv[1:]] or [4, 12]

def timeit(fn, iters):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

print("| P | N | bags | kernel us | GB/s | of 6544 | forward us | WSI/s |")
print("|---|---|---|---|---|---|---|---|")
for P in Ps:
    pr = synth.make_params(P, P, 1)
    Q = (0.5 * pr["residual_features"] + pr["prompt_features"]).to(dev)
    W, b, T, ls = (pr[k].to(dev) for k in ("W", "b", "text_features", "logit_scale"))
    for N in (1000, 10000, 50000, 100000):
        for B in (1, 32):
            nset = 2 if B * N * 512 * ES > 3e8 else 8
            Xs = [(torch.randn(N * B, 512, device=dev) * 1.1 + 0.7).to(DT) for _ in range(nset)]
            plan = ops.make_plan([N] * B, dev)
            ws = ops._workspace(plan, P, dev)
            iters = 20 if B * N >= 1e5 else 100
            ms_k = timeit(lambda i: ops.aggregate_partial_only(Xs[i % nset], plan, Q, ws), iters)
            def fwd(i):
                ops.aggregate_infer(Xs[i % nset], plan, Q, W, b, T, ls)
            ms_f = timeit(fwd, iters)
            gb = N * B * 512 * ES / 1e9
            print(f"| {P} | {N} | {B} | {ms_k*1e3:.1f} | {gb/ms_k*1e3:.0f} | {gb/ms_k*1e3/PEAK:.2f} | {ms_f*1e3:.1f} | {B/ms_f*1e3:.0f} |", flush=True)
            del Xs
