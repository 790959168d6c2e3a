"""Dev helper (GPU box): small forward + loss + backward through the three tcgen05 kernels (fp32 rows, bf16 rows, split16 cohort)
and the packing kernel, for compute-sanitizer --tool memcheck."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vlsa_b200 import ops, synth
from vlsa_b200.dataset import DeviceCohort
dev = torch.device("cuda:0")
ops.set_agg_variant("tc")
for P, sizes in ((12, [2798, 1000, 37, 1, 16, 17]), (4, [513, 64])):
    bags = [synth.make_bag("g1", n, 100 + i) for i, n in enumerate(sizes)]
    pr = synth.make_params(P, P, 7)
    t, e = synth.make_labels(len(sizes), P, 9)
    cohort = DeviceCohort(dev, sum((n + 15) // 16 * 16 for n in sizes), layout="split16")
    for i, b in enumerate(bags):
        cohort.add(i, b)
    Xf = torch.cat(bags, 0).to(dev)
    for name, X, plan in (("fp32", Xf, ops.make_plan(sizes, dev)), ("bf16", Xf.to(torch.bfloat16), ops.make_plan(sizes, dev)),
                          ("split16", cohort.X, cohort.plan(list(range(len(sizes)))))):
        leaf = lambda z: z.detach().clone().to(dev).requires_grad_(True)
        res, W, b, T, ls = (leaf(pr[k]) for k in ("residual_features", "W", "b", "text_features", "logit_scale"))
        Q = pr["res_ratio"] * res + pr["prompt_features"].to(dev)
        logits, g, Tn, inc, ml = ops.aggregate(X, plan, Q, W, b, T, ls)
        total, *_ = ops.surv_loss(logits, t.to(dev), e.to(dev), ls)
        total.backward()
        torch.cuda.synchronize()
        print(name, P, sizes, float(total), float(res.grad.abs().sum()), flush=True)
ops.set_agg_variant(None)
print("done")
