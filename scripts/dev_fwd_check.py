"""Dev helper (GPU box): forward parity against golden/oracle + quick kernel timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from vlsa_b200 import ops, synth
from golden_util import load_case, rebuild_inputs
from conftest import golden_cases

dev = torch.device("cuda:0")
names = golden_cases("single_") + golden_cases("real_")
worst = 0.0
for name in names:
    case = load_case(name)
    bags, pr, t, e = rebuild_inputs(name, case)
    X = bags[0].to(dev)
    Q = (pr["res_ratio"] * pr["residual_features"] + pr["prompt_features"]).to(dev)
    plan = ops.make_plan([X.shape[0]], dev)
    out = ops.aggregate_forward_raw(X, plan, Q, pr["W"].to(dev), pr["b"].to(dev), pr["text_features"].to(dev),
                                    pr["logit_scale"].to(dev))
    torch.cuda.synchronize()
    err = np.abs(out["incidence"].cpu().numpy() - case["if_f64"]).max()
    errf = np.abs(out["f"].cpu().numpy() - case["f_f64"]).max()
    ref_err = np.abs(case["if_f32"] - case["if_f64"]).max()
    worst = max(worst, err)
    print(f"{name:40s} IF err vs f64 {err:.2e} (ref f32: {ref_err:.2e})  f err {errf:.2e} chunks {plan.total_chunks} x {plan.chunk_rows}")
print("WORST IF err", worst)

# timing
for (P, N, B) in ((4, 50000, 32), (12, 50000, 32), (4, 50000, 1), (4, 10000, 32), (16, 50000, 8)):
    pr = synth.make_params(P, P, 1)
    X = torch.randn(N * B, 512, device=dev) * 1.1
    Q = (0.5 * pr["residual_features"] + pr["prompt_features"]).to(dev)
    W, b, T, ls = (pr[k].to(dev) for k in ("W", "b", "text_features", "logit_scale"))
    plan = ops.make_plan([N] * B, dev)
    ws = ops._workspace(plan, P, dev)
    for _ in range(3):
        ops.aggregate_forward_raw(X, plan, Q, W, b, T, ls, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        ops.aggregate_forward_raw(X, plan, Q, W, b, T, ls, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    gb = N * B * 512 * 4 / 1e9
    print(f"P={P} N={N} B={B}: {ms*1e3:.1f} us/step  {gb/ms*1e3:.0f} GB/s  ({gb/ms*1e3/6538.9*100:.1f}% of 6538.9)  {B/ms*1e3:.0f} WSI/s  chunks={plan.total_chunks}x{plan.chunk_rows}")
