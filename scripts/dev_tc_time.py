"""Dev helper (GPU box): kernel-only timing of the streaming kernel for both variants + quick parity spot check."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from vlsa_b200 import ops, synth
    dev = torch.device("cuda:0")
    cfgs = [(4, 50000, 32), (8, 50000, 32), (12, 50000, 32), (16, 50000, 32), (12, 10000, 32), (12, 50000, 1)]
    for (P, N, B) in cfgs:
        pr = synth.make_params(P, P, 1)
        Xs = [torch.randn(N * B, 512, device=dev) * 1.1 for _ in range(2 if B > 1 else 8)]
        Q = (0.5 * pr["residual_features"] + pr["prompt_features"]).to(dev)
        plan = ops.make_plan([N] * B, dev)
        ws = ops._workspace(plan, P, dev)
        for i in range(3):
            ops.aggregate_partial_only(Xs[i % len(Xs)], plan, Q, ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for i in range(iters):
            ops.aggregate_partial_only(Xs[i % len(Xs)], plan, Q, ws)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        gb = N * B * 512 * 4 / 1e9
        print(f"   P={P:2d} N={N} B={B:2d}: {ms*1e3:8.1f} us  {gb/ms*1e3:6.0f} GB/s ({gb/ms*1e3/6538.9*100:5.1f}%)  chunks={plan.total_chunks}x{plan.chunk_rows}", flush=True)
    # parity spot check vs fp64 torch on GPU
    P, N = 12, 20000
    pr = synth.make_params(P, P, 3)
    X = synth.make_bag("g0", N, 5).to(dev)
    Q = (0.5 * pr["residual_features"] + pr["prompt_features"]).to(dev)
    plan = ops.make_plan([N], dev)
    out = ops.aggregate_forward_raw(X, plan, Q, pr["W"].to(dev), pr["b"].to(dev), pr["text_features"].to(dev), pr["logit_scale"].to(dev))
    Xd, Qd = X.double(), Q.double()
    Qn = Qd / Qd.norm(dim=-1, keepdim=True); Xn = Xd / Xd.norm(dim=-1, keepdim=True)
    A = torch.softmax(ops.coattn_scale() * Qn @ Xn.T, -1)
    v = (A @ Xd).mean(0)
    print("   v rel err vs fp64:", ((out["v"][0].double() - v).norm() / v.norm()).item(), flush=True)
else:
    for variant in (sys.argv[1:] or ["simt", "tc"]):
        print("==", variant, flush=True)
        env = dict(os.environ, VLSA_AGG_VARIANT=variant)
        subprocess.run(["timeout", "300", sys.executable, __file__, "child"], env=env)
