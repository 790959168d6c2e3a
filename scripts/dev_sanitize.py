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
# variant path: pooled forward, per-prototype-gradient backward (one launch and grouped), dX kernel
ops.set_agg_variant(None)
for P, sizes, prenorm in ((3, [513, 64, 1], False), (12, [700, 37], True), (16, [300, 0, 129], False)):
    g = torch.Generator().manual_seed(P)
    bags = [synth.make_bag("g1", n, 300 + i) for i, n in enumerate(sizes)]
    X = torch.cat(bags, 0).to(dev).requires_grad_(True)
    plan = ops.make_plan(sizes, dev)
    Q = torch.randn(P, 512, generator=g).to(dev).requires_grad_(True)
    dO = torch.randn(len(sizes), P, 512, generator=g).to(dev)
    O, ml = ops.pooled(X, plan, Q, prenorm)
    dX, dQ = torch.autograd.grad(O, (X, Q), dO)
    W = (torch.randn(512, 512, generator=g) / 22).to(dev).requires_grad_(True)
    f, _ = ops.encode(X, plan, Q, W, torch.zeros(512, device=dev), None, prenorm)
    dX2, = torch.autograd.grad(f.sum(), X)
    torch.cuda.synchronize()
    print("variant", P, sizes, float(O.abs().sum()), float(dX.abs().sum()), float(dQ.abs().sum()), float(dX2.abs().sum()), flush=True)
print("done")
