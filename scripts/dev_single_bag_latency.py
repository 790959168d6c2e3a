"""BASELINE configs[1]: one synthetic bag N=10k, D=512, P=R=4, batch 1, fp32 — wall-clock latency per call of the
reference-facing entry points, host side included (200 calls, one synchronize at the end)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vlsa_b200 import ops, synth  # noqa: E402
from vlsa_b200.model import VLSA  # noqa: E402

dev = torch.device("cuda:0")
out = {}
for P, N in ((4, 10000), (12, 10000), (4, 50000)):
    pr = synth.make_params(P, P, 3)
    img = dict(name="VLFAN", dim_in=512, use_feat_proj=False, query="Text", num_query=P, query_text_method="TaskRes")
    net = VLSA({"name": "mahmoodlab/conch"}, img, {"name": "CoOp"}, text_features=pr["text_features"],
               query_prompt_features=pr["prompt_features"], logit_scale_init=float(pr["logit_scale"])).to(dev)
    X = synth.make_bag("g1", N, 11).to(dev).unsqueeze(0)
    Xp = X[0]
    enc = net.mil_encoder
    T = net.forward_text_only()

    def wall(fn, iters=200):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / iters * 1e6

    plan = ops.make_plan([N], dev)
    ws = ops._workspace(plan, P, dev)
    with torch.no_grad():
        Q = enc.get_query().contiguous()
        rec = {
            "make_plan_us": wall(lambda: ops.make_plan([N], dev)),
            "raw_forward_given_plan_us": wall(lambda: ops.aggregate_forward_raw(
                Xp, plan, Q, enc.visual_adapter.weight, enc.visual_adapter.bias, T, net.logit_scale, need_bwd=False,
                workspace=ws)),
            "partial_kernel_only_us": wall(lambda: ops.aggregate_partial_only(Xp, plan, Q, ws)),
            "module_forward_no_grad_us": wall(lambda: net(X)),
        }
    rec["module_forward_with_grad_us"] = wall(lambda: net(X))

    def step():
        logits, _, _ = net(X)
        logits.sum().backward()
    rec["module_forward_backward_us"] = wall(step, iters=100)
    out[f"P{P}_N{N}"] = rec
    print(P, N, {k: round(v, 1) for k, v in rec.items()}, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/single_bag_latency.json", "w"), indent=1)
