import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vlsa_b200 import ops, synth
dev = torch.device("cuda:0")
P, N, B = 12, int(os.environ.get("DEV_N", 5000)), int(os.environ.get("DEV_B", 32))
pr = synth.make_params(P, P, 1)
X = torch.randn(N * B, 512, device=dev) * 1.1 + 0.7
leaf = lambda z: z.detach().clone().to(dev).requires_grad_(True)
res, W, b, T, ls = (leaf(pr[k]) for k in ("residual_features", "W", "b", "text_features", "logit_scale"))
pf = pr["prompt_features"].to(dev)
t, e = synth.make_labels(B, P, 9); t, e = t.to(dev), e.to(dev)
plan = ops.make_plan([N] * B, dev)
ops.set_agg_variant("tc")
for it in range(3):
    for z in (res, W, b, T, ls): z.grad = None
    logits, g, Tn, inc, ml = ops.aggregate(X, plan, 0.5 * res + pf, W, b, T, ls)
    total, *_ = ops.surv_loss(logits, t, e, ls)
    total.backward()
    torch.cuda.synchronize()
    print("iter", it, float(total), float(res.grad.abs().max()), flush=True)
