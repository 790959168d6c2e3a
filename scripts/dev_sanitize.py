"""Dev helper (GPU box): one small forward + loss + backward through both streaming kernels, for compute-sanitizer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vlsa_b200 import ops, synth
dev = torch.device("cuda:0")
for variant in (os.environ.get("DEV_VARIANTS", "tc,simt").split(",")):
    ops.set_agg_variant(variant)
    for P, sizes in ((12, [2798, 1000, 37, 1]), (4, [513, 64])):
        bags = [synth.make_bag("g1", n, 100 + i) for i, n in enumerate(sizes)]
        pr = synth.make_params(P, P, 7)
        t, e = synth.make_labels(len(sizes), P, 9)
        X = torch.cat(bags, 0).to(dev)
        plan = ops.make_plan(sizes, dev)
        leaf = lambda z: z.detach().clone().to(dev).requires_grad_(True)
        res, W, b, T, ls = (leaf(pr[k]) for k in ("residual_features", "W", "b", "text_features", "logit_scale"))
        Q = pr["res_ratio"] * res + pr["prompt_features"].to(dev)
        logits, g, Tn, inc, ml = ops.aggregate(X, plan, Q, W, b, T, ls)
        total, *_ = ops.surv_loss(logits, t.to(dev), e.to(dev), ls)
        total.backward()
        torch.cuda.synchronize()
        print(variant, P, sizes, float(total), float(res.grad.abs().sum()), flush=True)
print("done")
