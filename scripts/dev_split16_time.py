"""Dev helper (GPU box): fp32 rows (agg_tc_kernel) against the same bags as a split16 cohort (agg_split_kernel): kernel and
training-step times at 32 x 50k rows, and the one-off packing time.  argv: [P ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vlsa_b200 import ops, synth
from vlsa_b200.dataset import DeviceCohort
dev = torch.device("cuda:0")
N, B = 50000, 32
Ps = [int(v) for v in sys.argv[1:]] or [12]


def timeit(fn, iters=20, warm=3):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


Xs = [torch.randn(N * B, 512, device=dev) * 1.1 + 0.7 for _ in range(2)]
cohorts = []
for X in Xs:
    c = DeviceCohort(dev, B * ((N + 15) // 16 * 16), layout="split16")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for b in range(B):
        c.add(b, X[b * N:(b + 1) * N])
    e1.record(); torch.cuda.synchronize()
    cohorts.append(c)
print(f"packing 32 x 50k rows: {e0.elapsed_time(e1):.2f} ms (once per cohort upload)")
gb = N * B * 512 * 4 / 1e9
for P in Ps:
    pr = synth.make_params(P, P, 1)
    leaf = lambda z: z.detach().clone().to(dev).requires_grad_(True)
    res, W, b, T, ls = (leaf(pr[k]) for k in ("residual_features", "W", "b", "text_features", "logit_scale"))
    pf = pr["prompt_features"].to(dev)
    t, e = synth.make_labels(B, P, 9); t, e = t.to(dev), e.to(dev)
    plan_rows = ops.make_plan([N] * B, dev)
    plans_c = [c.plan(list(range(B))) for c in cohorts]
    ws = ops._workspace(plan_rows, P, dev)
    Qd = (0.5 * res + pf).detach()
    ops.set_agg_variant("tc")
    for name, data in (("fp32 rows (agg_tc)", [(X, plan_rows) for X in Xs]), ("split16 cohort (agg_split)", [(c.X, p) for c, p in zip(cohorts, plans_c)])):
        ms_k = timeit(lambda i: ops.aggregate_partial_only(data[i % 2][0], data[i % 2][1], Qd, ws))

        def step(i):
            for z in (res, W, b, T, ls): z.grad = None
            logits, g, Tn, inc, ml = ops.aggregate(data[i % 2][0], data[i % 2][1], 0.5 * res + pf, W, b, T, ls)
            total, *_ = ops.surv_loss(logits, t, e, ls)
            total.backward()
        ms_s = timeit(step, iters=10)
        print(f"P={P:2d} {name:28s}: fwd kernel {ms_k*1e3:7.1f} us = {gb/ms_k*1e3:6.0f} GB/s ({gb/ms_k*1e3/6544*100:5.1f}% of 6544) | "
              f"train step {ms_s*1e3:7.1f} us = {2*gb/ms_s*1e3:6.0f} GB/s over 2 reads ({2*gb/ms_s*1e3/6544*100:5.1f}%)", flush=True)
    ops.set_agg_variant(None)
