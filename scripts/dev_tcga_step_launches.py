"""Dev helper (GPU box, under ncu): a few optimizer steps on TCGA-sized bags from a split16 cohort, for a per-kernel launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from vlsa_b200 import synth
from vlsa_b200.dataset import DeviceCohort
from vlsa_b200.runner import VLSAHandler
dev = torch.device("cuda:0")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 12
layout = sys.argv[2] if len(sys.argv) > 2 else "split16"
n_pat, bs = 64, 32
sizes = np.exp(np.random.default_rng(0).uniform(np.log(1000), np.log(20000), n_pat)).astype(int)
cohort = DeviceCohort(dev, int(sum((n + 15) // 16 * 16 for n in sizes)), layout=layout)
for i, n in enumerate(sizes):
    cohort.add(i, torch.randn(int(n), 512, device=dev) * 1.1 + 0.7)
net = bench.build_net(P, P, dev).train()
handler = VLSAHandler({"task": "vlsa", "arch": "VLSA", "loss_type": "SurvIFMLE-SurvEMD", "opt_name": "adam", "opt_lr": 2e-4}, net=net, device=dev)
t_lab, e_lab = synth.make_labels(n_pat, P, 7)
lab = torch.stack([t_lab, e_lab], 1)
for k in range(4):
    ids = list(range(32 * (k % 2), 32 * (k % 2) + 32))
    handler.step_packed(cohort.X, cohort.plan(ids), lab[ids])
torch.cuda.synchronize()
