"""Dev helper (GPU box): time the streaming kernel for each prebuilt library variant (subprocess per variant)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from vlsa_b200 import ops, synth
    dev = torch.device("cuda:0")
    cfgs = ((4, 50000, 32, torch.float32), (12, 50000, 32, torch.float32), (12, 10000, 32, torch.float32))
    if os.environ.get("DEV_CFGS") == "simt":
        cfgs = ((4, 50000, 32, torch.float32), (4, 50000, 32, torch.bfloat16), (12, 50000, 32, torch.bfloat16))
    for (P, N, B, dt) in cfgs:
        pr = synth.make_params(P, P, 1)
        Xs = [(torch.randn(N * B, 512, device=dev) * 1.1).to(dt) for _ in range(2 if B > 1 else 8)]
        Q = (0.5 * pr["residual_features"] + pr["prompt_features"]).to(dev)
        plan = ops.make_plan([N] * B, dev)
        ws = ops._workspace(plan, P, dev)
        for i in range(3):
            ops.aggregate_partial_only(Xs[i % len(Xs)], plan, Q, ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for i in range(iters):
            ops.aggregate_partial_only(Xs[i % len(Xs)], plan, Q, ws)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        gb = N * B * 512 * Xs[0].element_size() / 1e9
        print(f"   P={P:2d} N={N} B={B:2d} {str(dt)[6:]:8s}: {ms*1e3:8.1f} us  {gb/ms*1e3:6.0f} GB/s ({gb/ms*1e3/6650.0*100:5.1f}%)  chunks={plan.total_chunks}x{plan.chunk_rows}", flush=True)
else:
    libs = sorted(glob.glob(os.path.join(ROOT, "vlsa_b200/lib/variants/*.so")))
    for lib in libs:
        print("==", os.path.basename(lib), flush=True)
        env = dict(os.environ, VLSA_B200_LIB=lib, VLSA_AGG_VARIANT=os.environ.get("VLSA_AGG_VARIANT", "tc"))
        subprocess.run([sys.executable, __file__, "child"], env=env)
