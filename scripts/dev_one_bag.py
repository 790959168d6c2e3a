"""Dev helper (GPU box, under ncu): a few one-bag forward calls (BASELINE configs[1]) for a per-kernel launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vlsa_b200 import synth
from vlsa_b200.model import VLSA
dev = torch.device("cuda:0")
P, N = int(sys.argv[1]), int(sys.argv[2])
pr = synth.make_params(P, P, 3)
img = dict(name="VLFAN", dim_in=512, use_feat_proj=False, query="Text", num_query=P, query_text_method="TaskRes")
net = VLSA({"name": "mahmoodlab/conch"}, img, {"name": "CoOp"}, text_features=pr["text_features"],
           query_prompt_features=pr["prompt_features"], logit_scale_init=float(pr["logit_scale"])).to(dev).eval()
X = synth.make_bag("g1", N, 11).to(dev).unsqueeze(0)
with torch.no_grad():
    for _ in range(4):
        net(X)
torch.cuda.synchronize()
