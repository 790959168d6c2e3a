#!/usr/bin/env python
"""Benchmark of the VLSA language-guided aggregation forward on B200 (BASELINE.json metric:
"WSIs/sec at N=50k patches D=512 (1/2/4/8 GPU); fused-kernel HBM GB/s vs peak").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A step = one pass of the hot path (VLSA.forward for every bag) over one batch of 32 synthetic bags of
N=50k CONCH-like rows (D=512, P=R=4, fp32) per GPU; bags shard across ranks with no data-path
collective (weak scaling).  Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bag-rows", type=int, default=50000, help="N, patches per bag")
    ap.add_argument("--bags", type=int, default=32, help="bags per step per GPU")
    ap.add_argument("--P", type=int, default=4, help="text prototypes (BASELINE 'K')")
    ap.add_argument("--R", type=int, default=4, help="ordinal ranks")
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"], help="storage dtype of X")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step measurement (config 3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def _find_hbm_peak(obj, path=""):
    """All (key path, GB/s) candidates for an HBM bandwidth figure in a (possibly nested) MEASURED_PEAKS.json."""
    found = []
    if isinstance(obj, dict):
        for k, v in obj.items():
            kp = f"{path}.{k}" if path else str(k)
            if isinstance(v, (int, float)) and not isinstance(v, bool):
                name = kp.lower()
                if "hbm" in name or "copy" in name or "dram" in name or "bandwidth" in name:
                    val = float(v)
                    if "tb" in name and val < 100:          # a TB/s figure
                        val *= 1000.0
                    if 1000.0 < val < 20000.0:              # plausible GB/s for one B200
                        found.append((kp, val))
            else:
                found += _find_hbm_peak(v, kp)
    elif isinstance(obj, list):
        for i, v in enumerate(obj):
            found += _find_hbm_peak(v, f"{path}[{i}]")
    return found


def measured_peaks():
    """HBM peak for the roofline: the driver-written MEASURED_PEAKS.json when present (the streaming kernel is timed
    alone, so the burst figure is preferred over a sustained one), else the profiling guide's fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        if os.path.exists(path):
            with open(path) as fh:
                data = json.load(fh)
            if isinstance(data, dict) and isinstance(data.get("hbm_gbs"), (int, float)):
                return float(data["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
            cands = _find_hbm_peak(data)
            if cands:
                burst = [c for c in cands if "burst" in c[0].lower()]
                key, val = (burst or cands)[0]
                return val, f"measured (MEASURED_PEAKS.json {key})"
    except (OSError, ValueError, TypeError):
        pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as ex:  # pragma: no cover
            self.err = repr(ex)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def synth_batch_device(n_bags, rows, dev, seed, dtype):
    """G1 'CONCH-like' rows generated on the device: row norm ~25, pairwise cos ~0.7, rank ~64."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    m = torch.randn(512, generator=g, device=dev)
    m = 21.0 * m / m.norm()
    decay = 0.85 ** torch.arange(64, device=dev, dtype=torch.float32)
    bm = torch.randn(64, 512, generator=g, device=dev) * decay[:, None]
    X = torch.empty(n_bags * rows, 512, device=dev, dtype=dtype)
    step = 200000
    for r0 in range(0, n_bags * rows, step):
        r1 = min(n_bags * rows, r0 + step)
        z = torch.randn(r1 - r0, 64, generator=g, device=dev)
        blk = m + 0.30 * (z @ bm) + 0.15 * torch.randn(r1 - r0, 512, generator=g, device=dev)
        X[r0:r1] = blk.to(dtype)
    return X


def cpu_port_throughput(rows, P, R, seconds, threads):
    """The oracle port (torch CPU restatement of the reference forward) on a bounded sample:
    G1 bags of `rows` rows, forward only, repeated until ~`seconds` s of CPU work."""
    from oracle import vlsa_oracle as O
    from vlsa_b200 import synth
    torch.set_num_threads(threads)
    pr = synth.make_params(P, R, 1)
    Q = O.task_res_query(pr["prompt_features"], pr["residual_features"], pr["res_ratio"])
    bags = [synth.make_bag("g1", rows, 10 + i).unsqueeze(0) for i in range(2)]
    with torch.no_grad():
        for X in bags:                                    # warm-up
            O.softmax_converter(O.vlsa_forward(X, Q, pr["W"], pr["b"], pr["text_features"], pr["logit_scale"])[0])
        n, t0 = 0, time.perf_counter()
        while True:
            X = bags[n % len(bags)]
            O.softmax_converter(O.vlsa_forward(X, Q, pr["W"], pr["b"], pr["text_features"], pr["logit_scale"])[0])
            n += 1
            el = time.perf_counter() - t0
            if el >= seconds or n >= 4096:
                break
    return n / el, n, el


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle port: same ATen op
    sequence as model/deepmil.py:187-204 + model/vlsa.py:185-192), all host threads, rank 0 only."""
    if rank != 0:
        return
    from oracle import vlsa_oracle as O
    from vlsa_b200 import synth
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    P, R, rows = args.P, args.R, args.bag_rows
    sample_bags = 2                                        # a step = 2 of the 32 bags of the workload
    pr = synth.make_params(P, R, 1)
    Q = O.task_res_query(pr["prompt_features"], pr["residual_features"], pr["res_ratio"])
    bags = [synth.make_bag("g1", rows, 10 + i).unsqueeze(0) for i in range(sample_bags)]
    if args.dtype == "bf16":
        bags = [b.to(torch.bfloat16).float() for b in bags]

    def step():
        for X in bags:
            O.softmax_converter(O.vlsa_forward(X, Q, pr["W"], pr["b"], pr["text_features"], pr["logit_scale"])[0])

    with torch.no_grad():
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        el = time.perf_counter() - t0
    value = args.steps * sample_bags / el
    sample = f"{sample_bags} of the {args.bags} bags per step (N={rows}, P={P}, R={R}), {args.steps} steps"
    line = {
        "impl": "reference", "metric": "WSIs/sec at N=50k patches D=512", "value": value, "unit": "WSI/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": "WSI/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "WSI/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {
        "workload": f"configs[3] variable-N sweep at N={args.bag_rows}: batch={args.bags} bags/GPU/step, D=512, "
                    f"P={args.P} prototypes, R={args.R} ranks, X {args.dtype}, VLSA.forward (aggregation + adapter + "
                    f"cosine head + incidence softmax)",
        "bag_rows": args.bag_rows, "bags_per_gpu_per_step": args.bags, "global_bags_per_step": args.bags * world,
        "P": args.P, "R": args.R, "D": 512, "x_dtype": args.dtype, "parallelism": f"bag-sharded x{world} (no data-path collective)",
        "l2_policy": "inputs larger than L2: each step reads a 3.3 GB batch, 2 distinct batches alternate",
    }


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (use --impl reference for the CPU arm)"
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from vlsa_b200 import ops, synth
    from vlsa_b200.dataset import AsyncBagLoader
    from vlsa_b200.model import VLSA

    P, R, rows, nb = args.P, args.R, args.bag_rows, args.bags
    xdtype = torch.float32 if args.dtype == "fp32" else torch.bfloat16
    esize = 4 if args.dtype == "fp32" else 2
    pr = synth.make_params(P, R, 1)

    # ---- the model, through the public API -----------------------------------------------------------
    net = VLSA(text_encoder_cfg={"name": "mahmoodlab/conch"},
               image_encoder_cfg=dict(name="VLFAN", dim_in=512, dim_hid=256, use_feat_proj=False, query="Text",
                                      num_query=P, gated_query=False, query_pooling="mean", pred_head="default",
                                      query_text_method="TaskRes", query_text_res_ratio=0.5),
               prompt_learner_cfg={"name": "CoOp"}, text_features=pr["text_features"],
               query_prompt_features=pr["prompt_features"], logit_scale_init=float(pr["logit_scale"]),
               vlsa_api="CONCH", path_clip_model=None).to(dev)
    with torch.no_grad():
        net.mil_encoder.Q.residual_features.copy_(pr["residual_features"])
        net.mil_encoder.visual_adapter.weight.copy_(pr["W"])
        net.mil_encoder.visual_adapter.bias.copy_(pr["b"])
    net.eval()

    # ---- device-resident inputs: 2 distinct batches (each >> L2) -------------------------------------
    n_batches = 2
    batches = [synth_batch_device(nb, rows, dev, 1234 + 100 * rank + i, xdtype) for i in range(n_batches)]
    plan = ops.make_plan([rows] * nb, dev)
    Q = net.mil_encoder.get_query().detach().contiguous()
    W, b = net.mil_encoder.visual_adapter.weight.detach(), net.mil_encoder.visual_adapter.bias.detach()
    T, ls = net.forward_text_only().contiguous(), net.logit_scale.detach()
    ws = ops._workspace(plan, P, dev)
    use_tc = args.dtype == "fp32" and P > 5 and os.environ.get("VLSA_AGG_VARIANT", "")[:1] != "s" \
        or os.environ.get("VLSA_AGG_VARIANT", "")[:1] == "t" and args.dtype == "fp32"
    kernel_name = "agg_tc_kernel<false> (tcgen05)" if use_tc else "agg_simt_kernel<P,0,XT>"
    two_level = plan.total_chunks >= 8 * nb
    # streaming kernel, [merge level 1,] merge, adapter, head
    launches_per_step = 4 + (1 if two_level else 0)

    def step(i):
        return ops.aggregate_forward_raw(batches[i % n_batches], plan, Q, W, b, T, ls, need_bwd=False, workspace=ws)

    def sync_all():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for i in range(max(args.warmup, 3)):
        out = step(i)
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = step(i)
    e1.record()
    sync_all()
    ms_total = e0.elapsed_time(e1)

    # ---- dominant kernel alone (roofline): same inputs, events on the launching stream ---------------
    for i in range(3):
        ops.aggregate_partial_only(batches[i % n_batches], plan, Q, ws)
    torch.cuda.synchronize(dev)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for i in range(args.steps):
        ops.aggregate_partial_only(batches[i % n_batches], plan, Q, ws)
    k1.record()
    torch.cuda.synchronize(dev)
    kernel_ms = k0.elapsed_time(k1) / args.steps
    clocks = sampler.stop()

    # ---- one optimizer step (config 3): forward + fused loss + backward + ONE flat all-reduce + Adam ------
    train = None
    if not args.no_train:
        from vlsa_b200.runner import VLSAHandler
        cfg = {"task": "vlsa", "arch": "VLSA", "loss_type": "SurvIFMLE-SurvEMD", "opt_name": "adam", "opt_lr": 2e-4}
        net.train()
        handler = VLSAHandler(cfg, net=net, device=dev)
        t_lab, e_lab = synth.make_labels(nb, R, 77 + rank)
        label = torch.stack([t_lab, e_lab], 1).to(dev)
        n_global = nb * world

        def train_step(i):
            handler.bucket.zero()
            logits, _, _, _ = handler.net.forward_packed(batches[i % n_batches], plan)
            loss = handler.calc_objective_loss(logits, label, norm=n_global)
            loss.backward()
            handler.bucket.pack(loss.detach().reshape(1))
            handler.bucket.all_reduce()
            handler.bucket.unpack()
            handler.optimizer.step()

        n_train = max(3, min(args.steps, 20))
        for i in range(5):
            train_step(i)
        sync_all()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        for i in range(n_train):
            train_step(i)
        t1e.record()
        sync_all()
        train_ms = t0e.elapsed_time(t1e) / n_train
        tt_ = torch.tensor([train_ms], device=dev)
        if dist is not None:
            dist.all_reduce(tt_, op=dist.ReduceOp.MAX)
        train_ms = float(tt_[0])
        train = {"value": nb * world / (train_ms * 1e-3), "unit": "WSI/s", "ms_per_step": train_ms, "steps": n_train,
                 "what": "VLSAHandler step on device-resident bags: forward_packed + SurvIFMLE/SurvEMD + backward "
                         "(X read twice) + one flat-bucket all-reduce + Adam",
                 "hbm_gbs_over_two_reads": 2 * nb * rows * 512 * esize / (train_ms * 1e-3) / 1e9}   # per GPU
        net.eval()

    # ---- e2e: host buffers -> public API -> host result, copies inside the timed region --------------
    e2e = None
    if not args.no_e2e:
        # one pinned 3.28 GB batch per rank, re-sent every step (8 ranks would otherwise pin 52 GB of host memory)
        host = [torch.empty(nb * rows, 512, dtype=xdtype).pin_memory()]
        host[0].copy_(batches[0])
        res_host = torch.empty(nb, R, dtype=torch.float32).pin_memory()
        sizes = [rows] * nb
        n_e2e = max(3, min(args.steps, 10))

        def source(n):
            for i in range(n):
                yield host[i % len(host)], sizes, None, None

        def run_e2e(n):
            loader = AsyncBagLoader(source(n), dev, depth=2, max_rows=nb * rows, dtype=xdtype)
            with torch.no_grad():
                for batch in loader:
                    batch.wait()
                    logits, g, Tn, inc = net.forward_packed(batch.X, batch.plan, T)
                    loader.release(batch)
                    res_host.copy_(inc, non_blocking=True)
            torch.cuda.synchronize(dev)
            return loader.h2d_bytes / max(n, 1)

        del batches[1:]                                   # leave room: the loader owns its own ring
        run_e2e(2)
        sync_all()
        t0 = time.perf_counter()
        h2d = run_e2e(n_e2e)
        torch.cuda.synchronize(dev)
        e2e_s = time.perf_counter() - t0
        e2e_t = torch.tensor([e2e_s], device=dev)
        if dist is not None:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e = {"value": n_e2e * nb * world / float(e2e_t.item()), "unit": "WSI/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": nb * R * 4, "steps": n_e2e,
               "path": "pinned host batch -> AsyncBagLoader (copy stream, 2-slot ring) -> VLSA.forward_packed -> pinned host incidence"}

    # ---- reduce over ranks ----------------------------------------------------------------------------
    tt = torch.tensor([ms_total, kernel_ms], device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total, kernel_ms = float(tt[0]), float(tt[1])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    algo_bytes = nb * rows * 512 * esize                  # X read exactly once
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    value = args.steps * nb * world / (ms_total * 1e-3)
    line = {
        "metric": "WSIs/sec at N=50k patches D=512", "value": value, "unit": "WSI/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.dtype == "fp32" else "f32 accumulate, bf16 storage", "data": "synthetic",
        "config": workload_config(args, world),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": kernel_name, "kernel_ms": kernel_ms,
                     "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src,
                     "kernel_share_of_step": kernel_ms / (ms_total / args.steps),
                     "read_only_ceiling_gbs": 7300.0,
                     "read_only_ceiling_source": "scripts/dev_readbw.cu on this pool: a pure cp.async.bulk read stream "
                                                 "of the same 3.28 GB (profiles/readbw_r01.txt)"},
        "clocks": clocks, "gpu_launches": launches_per_step * args.steps, "e2e": e2e, "train_step": train,
    }
    if train is not None:
        train["frac_of_peak"] = train["hbm_gbs_over_two_reads"] / peak   # per GPU
    traffic_file = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as fh:
                tj = json.load(fh)
            key = f"P{P}_N{rows}_B{nb}_{args.dtype}"
            if key in tj:
                line["roofline"]["traffic"] = tj[key]["dram_bytes_per_launch"]
                line["roofline"]["traffic_source"] = tj[key].get("source")
        except Exception:
            pass
    if not args.no_cpu_baseline and world == 1:            # rank 0 at N=1 only
        cores = os.cpu_count() or 1
        v, n, el = cpu_port_throughput(rows, P, R, args.cpu_seconds, cores)
        line["cpu_baseline"] = {"value": v, "unit": "WSI/s", "cores": cores, "kind": "port",
                                "sample": f"{n} forwards of one N={rows} bag (P={P}, R={R}, fp32, torch CPU oracle port) in {el:.1f} s"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
