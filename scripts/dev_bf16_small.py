"""Dev helper (GPU box): where the fixed cost of one small bf16 call goes (host wall per call vs kernel duration under ncu)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vlsa_b200 import ops, synth
dev = torch.device("cuda:0")
P = 12
pr = synth.make_params(P, P, 1)
Q = (0.5 * pr["residual_features"] + pr["prompt_features"]).to(dev)
for N in (1000, 10000):
    for dt in (torch.float32, torch.bfloat16, "split16"):
        if dt == "split16":
            from vlsa_b200.dataset import DeviceCohort
            coh = DeviceCohort(dev, (N + 15) // 16 * 16, layout="split16")
            coh.add(0, torch.randn(N, 512, device=dev) * 1.1 + 0.7)
            X, plan = coh.X, coh.plan([0])
        else:
            X = (torch.randn(N, 512, device=dev) * 1.1 + 0.7).to(dt)
            plan = ops.make_plan([N], dev)
        ws = ops._workspace(plan, P, dev)
        for _ in range(20): ops.aggregate_partial_only(X, plan, Q, ws)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(300): ops.aggregate_partial_only(X, plan, Q, ws)
        t_issue = (time.perf_counter() - t0) / 300 * 1e6
        torch.cuda.synchronize(); t_all = (time.perf_counter() - t0) / 300 * 1e6
        print(f"N={N} {str(dt)[-8:]:8s} chunks={plan.total_chunks} x {plan.chunk_rows}: host issue {t_issue:.1f} us/call, incl. drain {t_all:.1f} us/call", flush=True)
