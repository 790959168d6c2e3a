import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import vlsa_oracle as O
from vlsa_b200 import ops, synth
dev = torch.device("cuda:0")
P = R = 12
pr = synth.make_params(P, R, 7)
for sizes in ([2798], [1000], [37], [32], [33], [16], [17], [48], [64], [2798, 1000, 37]):
    bags = [synth.make_bag("g1", n, 100 + i) for i, n in enumerate(sizes)]
    t, e = synth.make_labels(len(sizes), R, 9)
    ref = O.forward_with_grads(bags, pr["prompt_features"], pr["residual_features"], pr["W"], pr["b"], pr["text_features"], pr["logit_scale"], t, e, dtype=torch.float64)
    gref = ref["d_residual"].numpy()
    X = torch.cat(bags, 0).to(dev); plan = ops.make_plan(sizes, dev)
    line = f"sizes {str(sizes):22s} chunk_rows {plan.chunk_rows:4d}:"
    for name, flag in (("simt", 0x100), ("tma", 0x200), ("reg", 0x400)):
        ops._agg_variant_flag = flag
        leaf = lambda z: z.detach().clone().to(dev).requires_grad_(True)
        res, W, b, T, ls = (leaf(pr[k]) for k in ("residual_features", "W", "b", "text_features", "logit_scale"))
        Q = pr["res_ratio"] * res + pr["prompt_features"].to(dev)
        logits, g, Tn, inc, ml = ops.aggregate(X, plan, Q, W, b, T, ls)
        total, *_ = ops.surv_loss(logits, t.to(dev), e.to(dev), ls)
        total.backward(); torch.cuda.synchronize()
        err = np.abs(res.grad.cpu().numpy() - gref).max() / max(np.abs(gref).max(), 1e-30)
        line += f"  {name} {err:.1e}"
    print(line)
ops._agg_variant_flag = 0
