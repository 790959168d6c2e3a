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
Qd = (pr["res_ratio"] * pr["residual_features"] + pr["prompt_features"]).to(dev)
Qn = torch.nn.functional.normalize(Qd.double(), dim=-1)
O64 = []
for bi in range(3):
    Xd = bags[bi].to(dev).double()
    A = torch.softmax(ops.coattn_scale() * Qn @ torch.nn.functional.normalize(Xd, dim=-1).T, -1)
    O64.append(A @ Xd)
for name, flag in (("simt", 0x100), ("tc (TMA)", 0x200), ("tc_reg (round 1)", 0x400)):
    ops._agg_variant_flag = flag
    leaf = lambda z: z.detach().clone().to(dev).requires_grad_(True)
    res, W, b, T, ls = (leaf(pr[k]) for k in ("residual_features", "W", "b", "text_features", "logit_scale"))
    Q = pr["res_ratio"] * res + pr["prompt_features"].to(dev)
    logits, g, Tn, inc, ml = ops.aggregate(X, plan, Q, W, b, T, ls)
    total, *_ = ops.surv_loss(logits, t.to(dev), e.to(dev), ls)
    total.backward(); torch.cuda.synchronize()
    err = np.abs(res.grad.cpu().numpy() - gref).max() / np.abs(gref).max()
    o = ops.aggregate_forward_raw(X, plan, Qd, W.detach(), b.detach(), T.detach(), ls.detach(), need_bwd=True)
    oe = max(((o["O"][bi].double() - O64[bi]).norm(dim=-1) / O64[bi].norm(dim=-1)).max().item() for bi in range(3))
    print(f"{name:18s}: d_residual rel err {err:.2e}   O per-prototype rel err {oe:.2e}")
ops._agg_variant_flag = 0
outs = {}
for name, flag in (("simt", 0x100), ("tma", 0x200), ("reg", 0x400)):
    ops._agg_variant_flag = flag
    o = ops.aggregate_forward_raw(X, plan, Qd, pr["W"].to(dev), pr["b"].to(dev), pr["text_features"].to(dev), pr["logit_scale"].to(dev), need_bwd=True)
    torch.cuda.synchronize()
    outs[name] = {k: v.double().clone() for k, v in o.items() if isinstance(v, torch.Tensor) and k != "_workspace"}
ops._agg_variant_flag = 0
lse = {k: v["ml"][..., 0] + v["ml"][..., 1].log() for k, v in outs.items()}
lse64 = torch.stack([torch.logsumexp(ops.coattn_scale() * Qn @ torch.nn.functional.normalize(bags[bi].to(dev).double(), dim=-1).T, -1) for bi in range(3)])
for k in lse:
    print(f"LSE err {k:5s}: max {float((lse[k] - lse64).abs().max()):.2e}   per bag {[f'{float(x):.1e}' for x in (lse[k] - lse64).abs().max(dim=1).values]}")
print("m   tma - reg:", float((outs["tma"]["ml"][..., 0] - outs["reg"]["ml"][..., 0]).abs().max()))
print("l   tma / reg - 1:", float((outs["tma"]["ml"][..., 1] / outs["reg"]["ml"][..., 1] - 1).abs().max()))
print("O   tma vs reg rel:", float(((outs["tma"]["O"] - outs["reg"]["O"]).norm(dim=-1) / outs["reg"]["O"].norm(dim=-1)).max()))
print("v   tma vs reg rel:", float(((outs["tma"]["v"] - outs["reg"]["v"]).norm(dim=-1) / outs["reg"]["v"].norm(dim=-1)).max()))
