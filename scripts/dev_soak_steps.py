"""Dev helper (GPU box): soak of the handler's two step paths — random ragged steps (bag counts 1-40, sizes 0-30k rows incl. empty
bags, fp32 / bf16 rows, P in {4, 7, 12}) through the fused C-call step + bucket Adam + no-sync entry on one handler and through
the autograd path on a twin; losses, predictions and weights must stay bit-identical step after step (persistent buffers, the pinned
upload ring and the plan cache are all exercised with changing shapes)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from vlsa_b200 import synth
from vlsa_b200.runner import VLSAHandler
dev = torch.device("cuda:0")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 150
rng = np.random.default_rng(123)
pool = torch.randn(400000, 512, device=dev) * 1.1 + 0.7
bad = 0
t0 = time.time()
for P, dt in ((12, torch.float32), (4, torch.float32), (7, torch.bfloat16)):
    cfg = {"task": "vlsa", "arch": "VLSA", "loss_type": "SurvIFMLE-SurvEMD", "opt_name": "adam", "opt_lr": 2e-4}
    ha = VLSAHandler(dict(cfg), net=bench.build_net(P, P, dev).train(), device=dev)
    hb = VLSAHandler(dict(cfg, vlsa_fused_step=False), net=bench.build_net(P, P, dev).train(), device=dev)
    src = pool.to(dt)
    pending = []
    for s in range(steps):
        nb = int(rng.integers(1, 41))
        sizes = [int(v) for v in np.exp(rng.uniform(np.log(1), np.log(30000), nb))]
        if rng.random() < 0.3:
            sizes[int(rng.integers(0, nb))] = 0
        if sum(sizes) == 0:
            sizes[0] = 5
        xs, at = [], 0
        for n in sizes:
            if at + n > src.shape[0]:
                at = 0
            xs.append(src[at:at + n].unsqueeze(0)); at += n
        t, e = synth.make_labels(nb, P, 1000 + s)
        ys = [torch.stack([t[i], e[i]]).float().reshape(1, 2) for i in range(nb)]
        la, pa = ha._update_network(xs, ys, sync=False)
        lb, pb = hb._update_network(xs, ys, sync=bool(s % 7 == 0))
        pending.append((s, la, pa, lb, pb))
        if len(pending) == 10 or s == steps - 1:
            torch.cuda.synchronize()
            for (k, la, pa, lb, pb) in pending:
                ok = float(la) == float(lb) and torch.equal(pa.cpu(), pb.cpu())
                bad += int(not ok)
                if not ok:
                    print(f"P={P} {dt}: step {k} differs: {float(la)} vs {float(lb)}", flush=True)
            pending = []
    same = all(torch.equal(va, vb) for (_, va), (_, vb) in zip(ha.net.state_dict().items(), hb.net.state_dict().items()))
    print(f"P={P} {str(dt)[6:]}: {steps} random steps, weights identical: {same}", flush=True)
    bad += int(not same)
print(f"done in {time.time() - t0:.1f} s, {bad} mismatches")
