import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oracle import vlsa_oracle as O
from vlsa_b200 import ops, synth
dev = torch.device("cuda:0")
for P, sizes, kind in ((12, [2798, 1000, 37], "g1"), (12, [5000, 3001], "g0"), (4, [10000, 33], "g1"), (16, [5000, 33], "g0")):
    bags = [synth.make_bag(kind, n, 100 + i).to(torch.bfloat16).float() for i, n in enumerate(sizes)]
    pr = synth.make_params(P, P, 7)
    t, e = synth.make_labels(len(sizes), P, 9)
    ref = O.forward_with_grads(bags, pr["prompt_features"], pr["residual_features"], pr["W"], pr["b"], pr["text_features"], pr["logit_scale"], t, e, dtype=torch.float64)
    inc_ref = torch.softmax(ref["logits"], -1).numpy(); gref = ref["d_residual"].numpy()
    for variant in ("simt", "tc"):
        ops.set_agg_variant(variant)
        X = torch.cat(bags, 0).to(dev).to(torch.bfloat16)
        plan = ops.make_plan(sizes, dev)
        leaf = lambda z: z.detach().clone().to(dev).requires_grad_(True)
        res, W, b, T, ls = (leaf(pr[k]) for k in ("residual_features", "W", "b", "text_features", "logit_scale"))
        Q = pr["res_ratio"] * res + pr["prompt_features"].to(dev)
        logits, g, Tn, inc, ml = ops.aggregate(X, plan, Q, W, b, T, ls)
        total, *_ = ops.surv_loss(logits, t.to(dev), e.to(dev), ls)
        total.backward(); torch.cuda.synchronize()
        print(f"P={P} {sizes} {variant}: IF err {np.abs(inc.detach().cpu().numpy() - inc_ref).max():.2e} logits err {np.abs(logits.detach().cpu().numpy() - ref['logits'].numpy()).max():.2e} "
              f"d_res rel {np.abs(res.grad.cpu().numpy() - gref).max() / np.abs(gref).max():.2e}")
    ops.set_agg_variant(None)
