"""Disk -> GPU throughput of the flat feature store: an on-disk PatchFeatureStore
(memory-mapped, one flat file) -> WSIPatchSurvStore.steps() (threaded copies into pinned staging) -> AsyncBagLoader (H2D on a copy
stream) -> VLSA.forward_packed -> incidence back on the host, against the reference's way of feeding a step (one
torch.save'd tensor per slide -> torch.load + cat + .float() -> .cuda() per bag -> forward per bag: dataset/PatchWSI.py:197-215,
runner/vlsa_handler.py:205,322-337 — timed here with OUR forward so that only the feeding differs).
    python scripts/dev_store_e2e.py [--bags 64] [--rows 20000] [--P 4]"""
import argparse, json, os, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from vlsa_b200 import ops, synth
from vlsa_b200.dataset import AsyncBagLoader, PatchFeatureStore, WSIPatchSurvStore, build_store

ap = argparse.ArgumentParser()
ap.add_argument("--bags", type=int, default=64)
ap.add_argument("--rows", type=int, default=20000)
ap.add_argument("--P", type=int, default=4)
ap.add_argument("--threads", type=int, default=8)
args = ap.parse_args()
dev = torch.device("cuda:0")
P = R = args.P
net = bench.build_net(P, R, dev)
T = net.forward_text_only().contiguous()
root = tempfile.mkdtemp(prefix="vlsa_store_e2e_")
try:
    rs = np.random.RandomState(0)
    sizes = [int(v) for v in rs.randint(args.rows // 2, args.rows * 3 // 2, args.bags)]
    # the same slides in both formats
    def slides():
        for i, n in enumerate(sizes):
            yield f"s{i:04d}", synth.make_bag("g1", n, 9000 + i)
    t0 = time.time()
    meta = build_store(os.path.join(root, "flat"), slides())
    os.makedirs(os.path.join(root, "pt"))
    for sid, x in slides():
        torch.save(x, os.path.join(root, "pt", sid + ".pt"))
    os.sync()
    gb = meta["rows"] * 2048 / 1e9
    print(f"[store] {args.bags} slides, {meta['rows']} rows, {gb:.2f} GB per format, written in {time.time() - t0:.1f} s", flush=True)
    pids = [f"p{i:04d}" for i in range(args.bags)]
    pid2sids = {p: [f"s{i:04d}"] for i, p in enumerate(pids)}
    pid2label = {p: (1.0, 1.0) for p in pids}

    def run_flat():
        store = PatchFeatureStore(os.path.join(root, "flat"))
        ds = WSIPatchSurvStore(store, pids, pid2sids, pid2label)
        out = []
        with torch.no_grad():
            for batch in AsyncBagLoader(ds.steps(batch_size=32, threads=args.threads), dev, depth=2):
                batch.wait()
                out.append(torch.softmax(net.forward_packed(batch.X, batch.plan, T)[0], -1).cpu())
        return torch.cat(out)

    def run_reference_feeding():
        out = []
        with torch.no_grad():
            for i in range(args.bags):
                x = torch.load(os.path.join(root, "pt", f"s{i:04d}.pt")).float()          # PatchWSI.py:205-212
                x = x.unsqueeze(0).cuda()                                                  # vlsa_handler.py:205 (pageable, synchronous)
                out.append(torch.softmax(net(x)[0], -1).cpu())                                              # one forward per bag, result back per bag
        return torch.cat(out)

    res = {}
    for name, fn in (("flat store -> steps() -> AsyncBagLoader -> forward_packed", run_flat), ("per-slide .pt -> torch.load -> .cuda() -> forward per bag", run_reference_feeding)):
        fn()                                           # warm: page cache, allocator, plans
        torch.cuda.synchronize(); t0 = time.time()
        inc = fn()
        torch.cuda.synchronize(); dt = time.time() - t0
        res[name] = {"seconds": dt, "wsi_per_s": args.bags / dt, "gb_per_s": gb / dt}
        print(f"{name}: {dt*1e3:.0f} ms for {args.bags} bags = {args.bags/dt:.0f} WSI/s = {gb/dt:.2f} GB/s (page cache warm)", flush=True)
        res.setdefault("_inc", []).append(inc)
    a, b = res.pop("_inc")
    print("same incidence:", bool(torch.allclose(a, b, atol=2e-6)))
    print(json.dumps({"bags": args.bags, "mean_rows": float(np.mean(sizes)), "gb": gb, "P": P, "threads": args.threads, **res}))
finally:
    shutil.rmtree(root, ignore_errors=True)
