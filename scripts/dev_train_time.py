"""Dev helper (GPU box): timing of the streaming kernel alone, of the full forward and of one training step
(forward + fused loss + backward) per kernel variant (subprocess per variant), plus a gradient spot check
against fp64 torch autograd on the GPU."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from vlsa_b200 import ops, synth
    dev = torch.device("cuda:0")
    PEAK = 6650.0

    def timeit(fn, iters=10, warm=3):
        for i in range(warm): fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters): fn(i)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    cfgs = [(4, 50000, 32), (12, 50000, 32), (16, 50000, 32), (12, 10000, 32), (12, 2798, 32)]
    if os.environ.get("DEV_QUICK"): cfgs = cfgs[:2]
    if os.environ.get("DEV_PSWEEP"): cfgs = [(p_, 50000, 32) for p_ in [int(x) for x in os.environ.get('DEV_PLIST', '2,4,5,6,8,10,12').split(',')]]
    for (P, N, B) in cfgs:
        pr = synth.make_params(P, P, 1)
        Xs = [torch.randn(N * B, 512, device=dev) * 1.1 + 0.7 for _ in range(2)]
        leaf = lambda z: z.detach().clone().to(dev).requires_grad_(True)
        res, W, b, T, ls = (leaf(pr[k]) for k in ("residual_features", "W", "b", "text_features", "logit_scale"))
        pf = pr["prompt_features"].to(dev)
        t, e = synth.make_labels(B, P, 9)
        t, e = t.to(dev), e.to(dev)
        plan = ops.make_plan([N] * B, dev)
        ws = ops._workspace(plan, P, dev)
        Qd = (0.5 * res + pf).detach()
        gb = N * B * 512 * 4 / 1e9
        ms_k = timeit(lambda i: ops.aggregate_partial_only(Xs[i % 2], plan, Qd, ws), iters=20)

        def fwd(i):
            with torch.no_grad():
                ops.aggregate_forward_raw(Xs[i % 2], plan, Qd, W, b, T, ls, need_bwd=False)
        ms_f = timeit(fwd)

        def step(i):
            for z in (res, W, b, T, ls): z.grad = None
            Q = 0.5 * res + pf
            logits, g, Tn, inc, ml = ops.aggregate(Xs[i % 2], plan, Q, W, b, T, ls)
            total, *_ = ops.surv_loss(logits, t, e, ls)
            total.backward()
        ms_s = timeit(step)
        print(f"   P={P:2d} N={N} B={B}: kernel {ms_k*1e3:7.1f} us {gb/ms_k*1e3:6.0f} GB/s ({gb/ms_k*1e3/PEAK*100:4.1f}%) | "
              f"fwd {ms_f*1e3:7.1f} us | train step {ms_s*1e3:7.1f} us = {2*gb/ms_s*1e3:6.0f} GB/s over 2 reads "
              f"({2*gb/ms_s*1e3/PEAK*100:4.1f}%)  bwd-ish {1e3*(ms_s-ms_f):7.1f} us", flush=True)
    # gradient spot check vs fp64 autograd
    for (P, N, kind) in [(12, 20000, "g1"), (7, 3001, "g0"), (4, 5000, "g1")]:
        pr = synth.make_params(P, P, 3)
        X = synth.make_bag(kind, N, 5).to(dev)
        leaf = lambda z, dt=torch.float32: z.detach().clone().to(dev).to(dt).requires_grad_(True)
        res, W, b, T, ls = (leaf(pr[k]) for k in ("residual_features", "W", "b", "text_features", "logit_scale"))
        pf = pr["prompt_features"].to(dev)
        t, e = synth.make_labels(1, P, 9); t, e = t.to(dev), e.to(dev)
        plan = ops.make_plan([N], dev)
        logits, g, Tn, inc, ml = ops.aggregate(X, plan, 0.5 * res + pf, W, b, T, ls)
        total, *_ = ops.surv_loss(logits, t, e, ls)
        total.backward()
        # fp64 reference of the aggregation part: d total / d residual through logits (use our d_logits via autograd on fp64 graph)
        res64, W64, b64, T64, ls64 = (leaf(pr[k], torch.float64) for k in ("residual_features", "W", "b", "text_features", "logit_scale"))
        Xd = X.double()
        Q = 0.5 * res64 + pf.double()
        Qn = Q / Q.norm(dim=-1, keepdim=True); Xn = Xd / Xd.norm(dim=-1, keepdim=True)
        A = torch.softmax(ops.coattn_scale() * Qn @ Xn.T, -1)
        v = (A @ Xd).mean(0, keepdim=True)
        f = v @ W64.T + b64
        gg = f / f.norm(dim=-1, keepdim=True); Tn64 = T64 / T64.norm(dim=-1, keepdim=True)
        lg = ls64.exp() * gg @ Tn64.T
        # same upstream gradient as the CUDA path saw
        dl = torch.autograd.grad(ops.surv_loss(logits.detach().requires_grad_(True), t, e, ls.detach())[0], [], allow_unused=True) if False else None
        lg32 = logits.detach().clone().requires_grad_(True)
        tot2, *_ = ops.surv_loss(lg32, t, e, ls.detach()); tot2.backward()
        (lg * lg32.grad.double()).sum().backward()
        rel = lambda a, r: ((a.double() - r).abs().max() / r.abs().max()).item()
        print(f"   grad check P={P} N={N} {kind}: logits err {(logits.double()-lg).abs().max().item():.2e}  d_residual rel {rel(res.grad, res64.grad):.2e}  d_W rel {rel(W.grad, W64.grad):.2e}", flush=True)
else:
    for variant in (sys.argv[1:] or ["simt", "tc"]):
        print("==", variant, flush=True)
        env = dict(os.environ, VLSA_AGG_VARIANT=variant)
        subprocess.run(["timeout", "600", sys.executable, __file__, "child"], env=env)
