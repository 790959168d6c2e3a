"""Dev helper (GPU box): streaming-kernel time vs chunk size (plan built for a pretended SM count)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vlsa_b200 import ops, synth
dev = torch.device("cuda:0")
for P in (4, 12):
    pr = synth.make_params(P, P, 1)
    N, B = 50000, 32
    Xs = [torch.randn(N * B, 512, device=dev) * 1.1 + 0.7 for _ in range(2)]
    Q = (0.5 * pr["residual_features"] + pr["prompt_features"]).to(dev)
    for sms in (148, 111, 74, 49, 37):
        plan = ops.make_plan([N] * B, dev, sms=sms)
        ws = ops._workspace(plan, P, dev)
        for i in range(3): ops.aggregate_partial_only(Xs[i % 2], plan, Q, ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20): ops.aggregate_partial_only(Xs[i % 2], plan, Q, ws)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"P={P} plan-sms={sms}: chunks={plan.total_chunks} x {plan.chunk_rows} rows  {ms*1e3:.1f} us  {N*B*2048/ms/1e6:.0f} GB/s", flush=True)
