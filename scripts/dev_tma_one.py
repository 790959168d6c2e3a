"""Dev helper: a few launches of the TMA-fed tcgen05 forward kernel alone (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vlsa_b200 import ops, synth
dev = torch.device("cuda:0")
P, N, B = int(os.environ.get("DEV_P", 12)), 50000, 32
pr = synth.make_params(P, P, 1)
X = torch.randn(N * B, 512, device=dev) * 1.1 + 0.7
Q = (0.5 * pr["residual_features"] + pr["prompt_features"]).to(dev)
plan = ops.make_plan([N] * B, dev)
ws = ops._workspace(plan, P, dev)
ops.set_agg_variant(os.environ.get("DEV_VARIANT", "tc"))
for _ in range(3):
    ops.aggregate_partial_only(X, plan, Q, ws)
torch.cuda.synchronize()
