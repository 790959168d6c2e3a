"""Dev helper: which pass (forward / backward) of which kernel costs gradient accuracy on the ill-conditioned case."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import vlsa_oracle as O
from vlsa_b200 import ops, synth
dev = torch.device("cuda:0")
P = R = 12
sizes = [2798, 1000, 37]
bags = [synth.make_bag("g1", n, 100 + i) for i, n in enumerate(sizes)]
pr = synth.make_params(P, R, 7)
t, e = synth.make_labels(len(sizes), R, 9)
ref = O.forward_with_grads(bags, pr["prompt_features"], pr["residual_features"], pr["W"], pr["b"], pr["text_features"], pr["logit_scale"], t, e, dtype=torch.float64)
gref = ref["d_residual"].numpy()
X = torch.cat(bags, 0).to(dev); plan = ops.make_plan(sizes, dev)
for fv in ("tc", "simt"):
    for bv in ("tc", "simt"):
        ops.set_agg_variant(fv)
        leaf = lambda z: z.detach().clone().to(dev).requires_grad_(True)
        res, W, b, T, ls = (leaf(pr[k]) for k in ("residual_features", "W", "b", "text_features", "logit_scale"))
        Q = pr["res_ratio"] * res + pr["prompt_features"].to(dev)
        logits, g, Tn, inc, ml = ops.aggregate(X, plan, Q, W, b, T, ls)
        total, *_ = ops.surv_loss(logits, t.to(dev), e.to(dev), ls)
        ops.set_agg_variant(bv)
        total.backward(); torch.cuda.synchronize()
        err = np.abs(res.grad.cpu().numpy() - gref).max() / np.abs(gref).max()
        print(f"forward {fv:4s} backward {bv:4s}: d_residual rel err {err:.2e}   loss err {abs(total.item()-ref['loss'].item()):.2e}")
ops.set_agg_variant(None)
