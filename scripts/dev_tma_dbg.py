import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vlsa_b200 import ops, synth
dev = torch.device("cuda:0")
P = 12
for sizes in ([1, 7, 16, 17, 33], [2798, 1000, 37]):
    bags = [synth.make_bag("g1", n, 100 + i) for i, n in enumerate(sizes)]
    bags[1] = bags[1] * 1e-3; bags[2] = bags[2] * 3e3
    pr = synth.make_params(P, P, 7)
    t, e = synth.make_labels(len(sizes), P, 9)
    X = torch.cat(bags, 0).to(dev)
    plan = ops.make_plan(sizes, dev)
    res = {}
    for var in ("simt", "tc", "ref64"):
        dt = torch.float64 if var == "ref64" else torch.float32
        leaf = lambda z: z.detach().clone().to(dev).to(dt).requires_grad_(True)
        r, W, b, T, ls = (leaf(pr[k]) for k in ("residual_features", "W", "b", "text_features", "logit_scale"))
        if var == "ref64":
            Q = pr["res_ratio"] * r + pr["prompt_features"].to(dev).double()
            Qn = Q / Q.norm(dim=-1, keepdim=True)
            lg = []
            for bi in range(len(sizes)):
                Xd = bags[bi].to(dev).double()
                Xn = Xd / Xd.norm(dim=-1, keepdim=True)
                A = torch.softmax(ops.coattn_scale() * Qn @ Xn.T, -1)
                v = (A @ Xd).mean(0, keepdim=True)
                f = v @ W.T + b
                gg = f / f.norm(dim=-1, keepdim=True); Tn = T / T.norm(dim=-1, keepdim=True)
                lg.append(ls.exp() * gg @ Tn.T)
            logits = torch.cat(lg, 0)
            # same upstream gradient as the fp32 loss kernel gives on the simt logits
            (logits * res["simt"]["dlogits"].double()).sum().backward()
        else:
            ops.set_agg_variant(var)
            Q = pr["res_ratio"] * r + pr["prompt_features"].to(dev)
            logits, g, Tn, inc, ml = ops.aggregate(X, plan, Q, W, b, T, ls)
            logits.retain_grad()
            total, *_ = ops.surv_loss(logits, t.to(dev), e.to(dev), ls.detach())
            total.backward()
            torch.cuda.synchronize()
            ops.set_agg_variant(None)
        res[var] = dict(dres=r.grad.double().cpu(), dW=W.grad.double().cpu(), db=b.grad.double().cpu(), dT=T.grad.double().cpu(),
                        dlogits=(logits.grad if var != "ref64" else res["simt"]["dlogits"]).detach().clone())
    print("sizes", sizes)
    for k in ("dres", "dW", "db", "dT"):
        ref = res["ref64"][k]
        for var in ("simt", "tc"):
            d = (res[var][k] - ref).abs().max().item() / ref.abs().max().item()
            print(f"   {k:5s} {var:5s} rel err vs fp64 autograd {d:.3e}   (max |ref| {ref.abs().max().item():.3e})")
    print("   dlogits simt vs tc", (res["simt"]["dlogits"] - res["tc"]["dlogits"]).abs().max().item(), res["simt"]["dlogits"].abs().max().item())
