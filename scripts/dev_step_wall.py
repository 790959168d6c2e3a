"""Dev helper (GPU box): wall clock of one optimizer step of VLSAHandler on TCGA-sized bags (32 bags of 1k-20k rows per step, drawn
from a device cohort): fused C-call step vs autograd step, with and without the per-step host synchronisation; plus the wall clock
and the GPU-only time (CUDA-graph replay of the same launches) of one-bag forward calls (BASELINE configs[1])."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from vlsa_b200 import ops, synth
from vlsa_b200.dataset import DeviceCohort
from vlsa_b200.runner import VLSAHandler
dev = torch.device("cuda:0")
out = {}
n_pat, bs = 128, 32
sizes = np.exp(np.random.default_rng(0).uniform(np.log(1000), np.log(20000), n_pat)).astype(int)
for P, layout in (() if "single" in sys.argv else ((12, "split16"), (4, "rows"), (12, "rows"))):
    cohort = DeviceCohort(dev, int(sum((n + 15) // 16 * 16 for n in sizes)), layout=layout)
    for i, n in enumerate(sizes):
        cohort.add(i, torch.randn(int(n), 512, device=dev) * 1.1 + 0.7)
    t_lab, e_lab = synth.make_labels(n_pat, P, 7)
    ys_all = [torch.stack([t_lab[i], e_lab[i]]).float().reshape(1, 2) for i in range(n_pat)]
    for fused in (True, False):
        for sync in (False, True):
            net = bench.build_net(P, P, dev).train()
            handler = VLSAHandler({"task": "vlsa", "arch": "VLSA", "loss_type": "SurvIFMLE-SurvEMD", "opt_name": "adam", "opt_lr": 2e-4,
                                   "vlsa_fused_step": fused}, net=net, device=dev)
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
            ms = 1e3 * dt / (10 * n_pat / bs)
            out[f"P{P}_{layout}_{'fused' if fused else 'autograd'}_{'sync' if sync else 'nosync'}"] = ms
            print(f"P={P:2d} {layout:8s} {'fused   ' if fused else 'autograd'} {'sync  ' if sync else 'nosync'}: {ms:.3f} ms per optimizer step "
                  f"(32 bags, mean {sizes.mean():.0f} rows), {10 * n_pat / dt:.0f} bags/s", flush=True)
    del cohort

# one bag per call
from vlsa_b200.model import VLSA
for P, N in ((4, 10000), (12, 10000), (12, 2798)):
    pr = synth.make_params(P, P, 3)
    img = dict(name="VLFAN", dim_in=512, use_feat_proj=False, query="Text", num_query=P, query_text_method="TaskRes")
    net = VLSA({"name": "mahmoodlab/conch"}, img, {"name": "CoOp"}, text_features=pr["text_features"],
               query_prompt_features=pr["prompt_features"], logit_scale_init=float(pr["logit_scale"])).to(dev).eval()
    X = synth.make_bag("g1", N, 11).to(dev).unsqueeze(0)

    def wall(fn, iters=300):
        for _ in range(30): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(iters): fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / iters * 1e6
    with torch.no_grad():
        us_call = wall(lambda: net(X))
        rec = {"module_forward_no_grad_us": us_call}
        try:
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3): net(X)
            torch.cuda.current_stream().wait_stream(s)
            with torch.cuda.graph(g):
                o = net(X)
            rec["graph_replay_us"] = wall(lambda: g.replay())
        except Exception as ex:
            rec["graph_replay_us"] = f"capture failed: {type(ex).__name__}: {str(ex)[:200]}"
    out[f"single_P{P}_N{N}"] = rec
    print(P, N, rec, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/step_wall.json", "w"), indent=1)
