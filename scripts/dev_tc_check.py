"""Dev helper (GPU box): the tcgen05 kernels (variant `tc`: register-staged for fp32 rows, TMA-fed for bf16 rows) against the CUDA-core kernel
(parity of forward outputs and gradients on the same inputs), run-to-run bit stability, and kernel / train-step times.
    VLSA_B200_LIB=/path/to/variant.so python scripts/dev_tc_check.py [quick|timeonly] [variants ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vlsa_b200 import ops, synth

dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "full"
variants = sys.argv[2:] or ["tc"]
FEWP = bool(os.environ.get("VLSA_DEV_FEWP"))
DT = torch.bfloat16 if os.environ.get("DEV_DTYPE") == "bf16" else torch.float32
PS = [int(v) for v in os.environ.get("DEV_PS", "").split(",") if v]


def run(variant, bags, pr, dtype=None):
    dtype = dtype or DT
    ops.set_agg_variant(variant)
    X = torch.cat(bags, 0).to(dev).to(dtype)
    plan = ops.make_plan([b.shape[0] for b in bags], dev)
    leaf = lambda z: z.detach().clone().to(dev).requires_grad_(True)
    res, W, b, T, ls = (leaf(pr[k]) for k in ("residual_features", "W", "b", "text_features", "logit_scale"))
    Q = pr["res_ratio"] * res + pr["prompt_features"].to(dev)
    logits, g, Tn, inc, ml = ops.aggregate(X, plan, Q, W, b, T, ls)
    gen = torch.Generator().manual_seed(3)
    dl = torch.randn(logits.shape, generator=gen).to(dev)
    total = (logits * dl).sum()
    total.backward()
    torch.cuda.synchronize()
    ops.set_agg_variant(None)
    return dict(inc=inc.detach().cpu(), logits=logits.detach().cpu(), ml=ml.detach().cpu(), dres=res.grad.cpu(), dW=W.grad.cpu(),
                loss=float(total))


if mode != "timeonly":
    cases = [(12, [37], "g1"), (12, [1, 7, 16, 17, 33], "g1"), (12, [2798, 1000, 37], "g1"), (12, [5000, 3001], "g0"),
             (7, [2000, 999], "g0"), (16, [4097], "g1")]
    if mode != "quick":
        cases += [(12, [50000, 20000], "g1"), (12, [50000], "g0"), (8, [100000], "g1"), (12, [3000] * 40 + [17, 1, 250], "g1")]
    if FEWP:
        cases = [c for c in cases if c[0] in (4, 12)]
    for v in variants:
        worst = 0.0
        for P, sizes, kind in cases:
            bags = [synth.make_bag(kind, n, 100 + i) for i, n in enumerate(sizes)]
            if kind == "g1" and len(sizes) > 2:
                bags[1] = bags[1] * 1e-3          # a bag of tiny rows and one of huge rows: the power-of-two row scale
                bags[2] = bags[2] * 3e3
            pr = synth.make_params(P, P, 7)
            a = run("simt", bags, pr)
            b = run(v, bags, pr)
            rel = lambda x, y: float((x - y).abs().max() / max(float(y.abs().max()), 1e-30))
            err_if = float((a["inc"] - b["inc"]).abs().max())
            print(f"{v:6s} P={P:2d} sizes={sizes[:6]}{'...' if len(sizes) > 6 else ''} {kind}: IF |d| {err_if:.2e}  logits |d| "
                  f"{float((a['logits'] - b['logits']).abs().max()):.2e}  d_res rel {rel(b['dres'], a['dres']):.2e}  dW rel "
                  f"{rel(b['dW'], a['dW']):.2e}  loss {a['loss']:.6f} / {b['loss']:.6f}", flush=True)
            worst = max(worst, err_if)
        print(f"{v}: worst IF diff vs simt: {worst:.3e}")
        bags = [synth.make_bag("g1", n, 5 + i) for i, n in enumerate([20000, 3000])]
        pr = synth.make_params(12, 12, 7)
        r = [run(v, bags, pr) for _ in range(4)]
        print(f"{v}: bit-stable over 4 runs:", all(bool((r[0]["inc"] == x["inc"]).all() and (r[0]["dres"] == x["dres"]).all()) for x in r[1:]))


def timeit(fn, iters=20, warm=3):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


N, B = 50000, 32
for P in (PS or ([12] if (mode == "quick" or FEWP) else [12, 8, 16])):
    pr = synth.make_params(P, P, 1)
    Xs = [(torch.randn(N * B, 512, device=dev) * 1.1 + 0.7).to(DT) for _ in range(2)]
    leaf = lambda z: z.detach().clone().to(dev).requires_grad_(True)
    res, W, b, T, ls = (leaf(pr[k]) for k in ("residual_features", "W", "b", "text_features", "logit_scale"))
    pf = pr["prompt_features"].to(dev)
    t, e = synth.make_labels(B, P, 9); t, e = t.to(dev), e.to(dev)
    plan = ops.make_plan([N] * B, dev)
    ws = ops._workspace(plan, P, dev)
    Qd = (0.5 * res + pf).detach()
    gb = N * B * 512 * Xs[0].element_size() / 1e9
    for variant in variants + ["simt"]:
        ops.set_agg_variant(variant)
        ms_k = timeit(lambda i: ops.aggregate_partial_only(Xs[i % 2], plan, Qd, ws))

        def step(i):
            for z in (res, W, b, T, ls): z.grad = None
            logits, g, Tn, inc, ml = ops.aggregate(Xs[i % 2], plan, 0.5 * res + pf, W, b, T, ls)
            total, *_ = ops.surv_loss(logits, t, e, ls)
            total.backward()
        ms_s = timeit(step, iters=10)
        print(f"P={P:2d} {variant:7s}: fwd kernel {ms_k*1e3:7.1f} us = {gb/ms_k*1e3:6.0f} GB/s ({gb/ms_k*1e3/6544*100:5.1f}% of 6544) | "
              f"train step {ms_s*1e3:7.1f} us = {2*gb/ms_s*1e3:6.0f} GB/s over 2 reads ({2*gb/ms_s*1e3/6544*100:5.1f}%)", flush=True)
    ops.set_agg_variant(None)
