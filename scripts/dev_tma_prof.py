"""Dev helper (GPU box): per-role wait accounting of the tcgen05 kernels (agg_tc_kernel / agg_bf16_kernel) (library built with -DVLSA_TMA_PROF)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vlsa_b200 import ops, synth, _lib
dev = torch.device("cuda:0")
P, N, B = int(os.environ.get("DEV_P", 12)), int(os.environ.get("DEV_N", 50000)), int(os.environ.get("DEV_B", 32))
pr = synth.make_params(P, P, 1)
X = torch.randn(N * B, 512, device=dev) * 1.1 + 0.7
if os.environ.get('DEV_DTYPE') == 'bf16':
    X = X.to(torch.bfloat16)
Q = (0.5 * pr["residual_features"] + pr["prompt_features"]).to(dev)
plan = ops.make_plan([N] * B, dev)
ws = ops._workspace(plan, P, dev)
ops.set_agg_variant(sys.argv[1] if len(sys.argv) > 1 else "tc")
for _ in range(3):
    ops.aggregate_partial_only(X, plan, Q, ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.aggregate_partial_only(X, plan, Q, ws); e1.record(); torch.cuda.synchronize()
print(f"kernel {e0.elapsed_time(e1)*1e3:.1f} us")
buf = (C.c_longlong * 32)()
L = _lib.lib()
L.vlsa_debug_read_prof.restype = C.c_int
assert L.vlsa_debug_read_prof(buf) == 0
v = list(buf)
tiles = max(v[19], 1)
names = {0: "conv: wait landed | prod: wait empty", 1: "conv: group barrier | prod: loads + norm", 2: "conv: proxy fence | prod: split + STS", 3: "conv | prod: TOTAL",
         4: "tma: wait empty", 5: "tma: TOTAL", 6: "gemm1: wait s_free", 7: "gemm1: wait full", 8: "gemm1: TOTAL",
         9: "gemm2: wait w_ready", 10: "gemm2: wait d2_free", 11: "gemm2: TOTAL", 12: "wt: wait s_ready", 13: "wt: wait full",
         14: "wt: wait decided", 15: "wt: bar.red", 16: "wt: wait w_free", 17: "wt: wait d2_done", 18: "wt: TOTAL (set 0: every other tile)"}
print(f"tiles per CTA: {tiles}")
for k in range(19):
    print(f"  {names[k]:40s} {v[k]/tiles:9.1f} cycles / tile")
print(f"block 0: Qn staged + TMEM allocated {v[22]/1e3:.1f} us, TMEM staged (warp 0) {v[23]/1e3:.1f} us, prologue {v[21]/1e3:.1f} us, kernel entry -> exit {v[20]/1e3:.1f} us")
