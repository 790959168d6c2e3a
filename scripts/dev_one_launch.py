"""Dev helper (GPU box): a few launches of the aggregation kernel alone (for ncu).  argv: variant [P] [bwd]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vlsa_b200 import ops, synth
dev = torch.device("cuda:0")
variant = sys.argv[1] if len(sys.argv) > 1 else "tc"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 12
N, B = 50000, 32
pr = synth.make_params(P, P, 1)
X = torch.randn(N * B, 512, device=dev) * 1.1 + 0.7
if os.environ.get('DEV_DTYPE') == 'bf16':
    X = X.to(torch.bfloat16)
Q = (0.5 * pr["residual_features"] + pr["prompt_features"]).to(dev)
plan = ops.make_plan([N] * B, dev)
if os.environ.get('DEV_LAYOUT') == 'split16':
    from vlsa_b200.dataset import DeviceCohort
    cohort = DeviceCohort(dev, B * ((N + 15) // 16 * 16), layout="split16")
    for b in range(B):
        cohort.add(b, X[b * N:(b + 1) * N])
    X, plan = cohort.X, cohort.plan(list(range(B)))
ws = ops._workspace(plan, P, dev)
ops.set_agg_variant(variant)
for _ in range(4):
    ops.aggregate_partial_only(X, plan, Q, ws)
torch.cuda.synchronize()
