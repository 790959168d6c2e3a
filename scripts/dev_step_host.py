"""Dev helper (GPU box): host-side cost of one VLSAHandler optimizer step (tiny bags: the GPU work is negligible, the
wall clock is what Python / the launches cost).  argv: [P]"""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from vlsa_b200 import ops, synth
from vlsa_b200.runner import VLSAHandler
dev = torch.device("cuda:0")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 4
nb = 32
sizes = [1000 + 37 * i for i in range(nb)]
X = torch.randn(sum(sizes), 512, device=dev)
plan = ops.make_plan(sizes, dev)
net = bench.build_net(P, P, dev).train()
handler = VLSAHandler({"task": "vlsa", "arch": "VLSA", "loss_type": "SurvIFMLE-SurvEMD", "opt_name": "adam", "opt_lr": 2e-4}, net=net, device=dev)
t_lab, e_lab = synth.make_labels(nb, P, 7)
label = torch.stack([t_lab, e_lab], 1).to(dev)


def step():
    handler.bucket.zero()
    logits, _, _, _ = handler.net.forward_packed(X, plan)
    loss = handler.calc_objective_loss(logits, label, norm=nb)
    loss.backward()
    handler.bucket.pack(loss.detach().reshape(1))
    handler.bucket.all_reduce()
    handler.bucket.unpack()
    handler.optimizer.step()


for _ in range(20): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): step()
torch.cuda.synchronize()
print(f"P={P}: {1e3 * (time.perf_counter() - t0) / 200:.3f} ms per step (wall, tiny bags)")
phases = {"zero": lambda: handler.bucket.zero(), }
pr = cProfile.Profile()
pr.enable()
for _ in range(200): step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr).sort_stats("cumulative")
st.print_stats(28)
