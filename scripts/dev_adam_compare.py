"""Dev helper (GPU box): VLSAHandler with BucketAdam vs with torch.optim.Adam, same bags, parameter drift per step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from vlsa_b200 import ops, synth
from vlsa_b200.runner import VLSAHandler
dev = torch.device("cuda:0")
P = 12
rng = np.random.default_rng(0)
groups = []
for gidx in range(4):
    sizes = [int(v) for v in np.exp(rng.uniform(np.log(1000), np.log(20000), 32))]
    bags = [synth.make_bag("g1", n, 40 + 100 * gidx + i) for i, n in enumerate(sizes)]          # HOST bags
    t, e = synth.make_labels(len(sizes), P, 5 + gidx)
    groups.append(([b.unsqueeze(0) for b in bags], [torch.stack([t[i], e[i]]).float().reshape(1, 2) for i in range(len(sizes))]))
hs = []
for adam in (True, False):
    net = bench.build_net(P, P, dev).train()
    hs.append(VLSAHandler({"task": "vlsa", "arch": "VLSA", "loss_type": "SurvIFMLE-SurvEMD", "opt_name": "adam", "opt_lr": 2e-4,
                           "opt_weight_decay": 1e-5, "vlsa_bucket_adam": adam}, net=net, device=dev))
for step in range(10):
    xs, ys = groups[step % 4]
    la, _ = hs[0]._update_network(xs, ys, sync=False)
    lb, _ = hs[1]._update_network(xs, ys, sync=False)
    line = [f"step {step}: loss {float(la):.6f} {float(lb):.6f}"]
    for (k, va), (_, vb) in zip(hs[0].net.state_dict().items(), hs[1].net.state_dict().items()):
        d = (va - vb).abs().max().item()
        line.append(f"{k.split('.')[-1]} {d:.2e} (|v| {vb.abs().max().item():.2e})")
    print("  ".join(line))
    if step == 4:
        sa, sb = hs[0].optimizer.state_dict(), hs[1].optimizer.state_dict()
        for i in sa["state"]:
            print("   state", i, float(sa["state"][i]["step"]), float(sb["state"][i]["step"]),
                  (sa["state"][i]["exp_avg"] - sb["state"][i]["exp_avg"]).abs().max().item(),
                  (sa["state"][i]["exp_avg_sq"] - sb["state"][i]["exp_avg_sq"]).abs().max().item(), sb["state"][i]["exp_avg_sq"].abs().max().item())
        print("   groups", [(g["lr"], g["weight_decay"], g["params"]) for g in sa["param_groups"]], [(g["lr"], g["weight_decay"], g["params"]) for g in sb["param_groups"]])
