"""Dev helper (GPU box): run-to-run bit stability of the tcgen05 kernels under repetition (the generic->async proxy fence of
agg_tc_kernel's tiles and of every kernel's weight operand is executed by the MMA issuer, not by the writers: a visibility
race would show as a sporadic mismatch).  argv: [iterations]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vlsa_b200 import ops, synth
from vlsa_b200.dataset import DeviceCohort
dev = torch.device("cuda:0")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
sizes = [50000, 20000, 3000, 17, 1, 9000, 33000, 12345]
bags = [synth.make_bag("g1", n, 50 + i) for i, n in enumerate(sizes)]
P = 12
pr = synth.make_params(P, P, 3)
t, e = synth.make_labels(len(sizes), P, 9)
Xf = torch.cat(bags, 0).to(dev)
cohort = DeviceCohort(dev, sum((n + 15) // 16 * 16 for n in sizes), layout="split16")
for i, b in enumerate(bags):
    cohort.add(i, b)
only = os.environ.get("DEV_ONLY")
cases = {"fp32 rows (agg_tc)": (Xf, ops.make_plan(sizes, dev)), "bf16 rows (agg_bf16)": (Xf.to(torch.bfloat16), ops.make_plan(sizes, dev)),
         "split16 cohort (agg_split)": (cohort.X, cohort.plan(list(range(len(sizes)))))}


def once(X, plan):
    leaf = lambda z: z.detach().clone().to(dev).requires_grad_(True)
    res, W, b, T, ls = (leaf(pr[k]) for k in ("residual_features", "W", "b", "text_features", "logit_scale"))
    Q = pr["res_ratio"] * res + pr["prompt_features"].to(dev)
    logits, g, Tn, inc, ml = ops.aggregate(X, plan, Q, W, b, T, ls)
    total, *_ = ops.surv_loss(logits, t.to(dev), e.to(dev), ls)
    total.backward()
    return [logits.detach(), ml.detach(), res.grad, W.grad, T.grad]


ops.set_agg_variant("tc")
for name, (X, plan) in cases.items():
    if only and only not in name:
        continue
    ref = once(X, plan)
    bad = 0
    for _ in range(iters):
        out = once(X, plan)
        bad += int(not all(torch.equal(a, b) for a, b in zip(ref, out)))
    torch.cuda.synchronize()
    print(f"{name}: {iters} repetitions of forward + backward, {bad} differed from the first", flush=True)
ops.set_agg_variant(None)
