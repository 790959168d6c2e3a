import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import vlsa_oracle as O
from vlsa_b200 import ops, synth
dev = torch.device("cuda:0")
P = R = 12
pr = synth.make_params(P, R, 7)
allb = [synth.make_bag("g1", n, 100 + i) for i, n in enumerate([2798, 1000, 37])]
for idx in ([0, 1], [1, 2], [0, 2], [1, 0], [2, 1, 0]):
    bags = [allb[i] for i in idx]; sizes = [b.shape[0] for b in bags]
    t, e = synth.make_labels(len(sizes), R, 9)
    ref = O.forward_with_grads(bags, pr["prompt_features"], pr["residual_features"], pr["W"], pr["b"], pr["text_features"], pr["logit_scale"], t, e, dtype=torch.float64)
    gref = ref["d_residual"].numpy()
    X = torch.cat(bags, 0).to(dev); plan = ops.make_plan(sizes, dev)
    line = f"sizes {str(sizes):18s}:"
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
# forward consistency: packed vs single-bag outputs of the TMA kernel
sizes = [2798, 1000, 37]
X = torch.cat(allb, 0).to(dev); plan = ops.make_plan(sizes, dev)
Qd = (pr["res_ratio"] * pr["residual_features"] + pr["prompt_features"]).to(dev)
args = [pr[k].to(dev) for k in ("W", "b", "text_features", "logit_scale")]
ops._agg_variant_flag = 0x200
pk = ops.aggregate_forward_raw(X, plan, Qd, *args, need_bwd=True)
for bi, n in enumerate(sizes):
    sg = ops.aggregate_forward_raw(allb[bi].to(dev), ops.make_plan([n], dev), Qd, *args, need_bwd=True)
    dO = ((pk["O"][bi] - sg["O"][0]).norm(dim=-1) / sg["O"][0].norm(dim=-1)).max().item()
    lse_p = pk["ml"][bi, :, 0].double() + pk["ml"][bi, :, 1].double().log(); lse_s = sg["ml"][0, :, 0].double() + sg["ml"][0, :, 1].double().log()
    print(f"bag {bi}: packed vs single  O rel {dO:.2e}   LSE diff {float((lse_p - lse_s).abs().max()):.2e}   IF diff {float((pk['incidence'][bi]-sg['incidence'][0]).abs().max()):.2e}")
ops._agg_variant_flag = 0
