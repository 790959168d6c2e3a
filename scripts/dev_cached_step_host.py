"""Dev helper (GPU box): host-side profile of VLSAHandler.update_network_cached on realistic TCGA-sized bags (the GPU work of a
step is a few hundred microseconds: the wall clock is what Python / launches / synchronisation cost)."""
import cProfile, os, pstats, sys, time
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
sync = (sys.argv[3] == "sync") if len(sys.argv) > 3 else False
n_pat, bs = 128, 32
sizes = np.exp(np.random.default_rng(0).uniform(np.log(1000), np.log(20000), n_pat)).astype(int)
cohort = DeviceCohort(dev, int(sum((n + 15) // 16 * 16 for n in sizes)), layout=layout)
for i, n in enumerate(sizes):
    cohort.add(i, torch.randn(int(n), 512, device=dev) * 1.1 + 0.7)
net = bench.build_net(P, P, dev).train()
handler = VLSAHandler({"task": "vlsa", "arch": "VLSA", "loss_type": "SurvIFMLE-SurvEMD", "opt_name": "adam", "opt_lr": 2e-4}, net=net, device=dev)
t_lab, e_lab = synth.make_labels(n_pat, P, 7)
ys_all = [torch.stack([t_lab[i], e_lab[i]]).float().reshape(1, 2) for i in range(n_pat)]
rng = np.random.default_rng(1)


def epoch():
    order = rng.permutation(n_pat).tolist()
    for s0 in range(0, n_pat, bs):
        ids = order[s0:s0 + bs]
        handler.update_network_cached(cohort, ids, [ys_all[i] for i in ids], sync=sync)


for _ in range(3): epoch()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): epoch()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"P={P} {layout}: {1e3 * dt / (10 * n_pat / bs):.3f} ms per optimizer step (32 bags, mean {sizes.mean():.0f} rows), {10 * n_pat / dt:.0f} bags/s")
# GPU time of the same steps: everything queued behind a long-running spin so that the host is far ahead
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(int(2e9 * 0.05))
e0.record()
for _ in range(2): epoch()
e1.record(); torch.cuda.synchronize()
print(f"GPU time per step when the host is ahead: {e0.elapsed_time(e1) / (2 * n_pat / bs):.3f} ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(10): epoch()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
