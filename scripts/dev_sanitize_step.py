"""Dev helper (GPU box): two handler steps (fused C-call step + bucket Adam) and a lean inference call on ragged bags, for
compute-sanitizer --tool memcheck / racecheck (covers merge_fwd_kernel, adam_step_kernel, the FusedTrainStep buffers)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from vlsa_b200 import ops, synth
from vlsa_b200.runner import VLSAHandler
dev = torch.device("cuda:0")
for P, dt in ((12, torch.float32), (4, torch.float32), (12, torch.bfloat16)):
    sizes = [2798, 1000, 37, 1, 16, 17, 513]
    bags = [synth.make_bag("g1", n, 100 + i).to(dev).to(dt) for i, n in enumerate(sizes)]
    t, e = synth.make_labels(len(sizes), P, 9)
    ys = [torch.stack([t[i], e[i]]).float().reshape(1, 2) for i in range(len(sizes))]
    net = bench.build_net(P, P, dev).train()
    h = VLSAHandler({"task": "vlsa", "arch": "VLSA", "loss_type": "SurvIFMLE-SurvEMD", "opt_name": "adam", "opt_lr": 2e-4}, net=net, device=dev)
    for _ in range(2):
        loss, preds = h._update_network([b.unsqueeze(0) for b in bags], ys)
    net.eval()
    with torch.no_grad():
        out = net.forward_packed(torch.cat(bags, 0), ops.make_plan(sizes, dev))
        one = net(bags[0].unsqueeze(0))
    torch.cuda.synchronize()
    print(P, dt, loss, float(out[0].abs().sum()), float(one[0].abs().sum()), flush=True)
print("done")
