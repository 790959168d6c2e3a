#!/usr/bin/env python
"""BASELINE configs[2] as a runnable example: slide-sharded VLSA training on a synthetic TCGA-BLCA-like cohort.

    python examples/train_synthetic_blca.py                                  # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        examples/train_synthetic_blca.py --patients 373                      # 8 GPUs, bags sharded per step

Pipeline (all of it is the accelerated path): synthetic patients -> flat memory-mapped feature store
(`vlsa_b200.dataset.build_store`) -> `WSIPatchSurvStore` (same items as the reference's `WSIPatchSurv`) ->
`VLSAHandler._train_each_epoch` with `cfg_vlsa_conch.yaml`-style settings (batch of 32 bags per optimizer step, bags
LPT-sharded over the ranks, ONE flat-bucket NCCL all-reduce, Adam) -> `test_model` -> concordance index.

The cohort has a planted signal so that learning is visible: a patient's latent risk tilts a fraction of its patches
towards one of the prototype directions, and the (discretised) survival time decreases with the risk; 45 % of the
patients are events (BLCA: 169 / 373).  There is no network access in this image, so CONCH text features and real
slides are replaced by fixed random tensors of the same shapes (P = 12 prototypes, R = 12 time bins).
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def concordance_index(risk: np.ndarray, t: np.ndarray, e: np.ndarray) -> float:
    """Harrell's C on (time bin, event) labels: comparable pairs = (i an event, t_i < t_j)."""
    num = den = 0.0
    for i in np.where(e > 0)[0]:
        later = t > t[i]
        den += later.sum()
        num += (risk[i] > risk[later]).sum() + 0.5 * (risk[i] == risk[later]).sum()
    return float(num / max(den, 1.0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patients", type=int, default=128)
    ap.add_argument("--min-n", type=int, default=1000)
    ap.add_argument("--max-n", type=int, default=20000)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--P", type=int, default=12)
    ap.add_argument("--R", type=int, default=12)
    ap.add_argument("--seed", type=int, default=0)
    # VLFAN switches no shipped config enables; all of them run on the same streaming kernels
    ap.add_argument("--gated-query", action="store_true")
    ap.add_argument("--query-pooling", default="mean", choices=["mean", "max", "weight", "attention", "gated_attention"])
    ap.add_argument("--feat-proj", action="store_true")
    ap.add_argument("--handler-loop", action="store_true",
                    help="run every epoch through VLSAHandler._train_each_epoch(loader) — the reference's own loop "
                         "(runner/vlsa_handler.py:189-239) — with cfg vlsa_device_cohort = --cohort: the handler keeps the bags it "
                         "has seen resident in HBM and the dataset stops reading them after the first epoch (1 GPU)")
    ap.add_argument("--print-steps", action="store_true", help="print the loss of every optimizer step")
    ap.add_argument("--autograd-step", action="store_true",
                    help="run every optimizer step through torch autograd and torch.optim.Adam instead of the fused C-call step "
                         "and the bucket Adam kernel (same kernels underneath; for cross-checking the two)")
    ap.add_argument("--cohort", default="none", choices=["none", "rows", "split16"],
                    help="keep every bag of this rank's shard resident in HBM after epoch 0 (DeviceCohort): 'rows' = fp32 rows, "
                         "'split16' = pre-split tile records (the tensor-core kernel then converts nothing per epoch)")
    args = ap.parse_args()

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from vlsa_b200 import synth
    from vlsa_b200.dataset import PatchFeatureStore, WSIPatchSurvStore, build_store, DeviceCohort
    from vlsa_b200.model import VLSA
    from vlsa_b200.runner import VLSAHandler

    P, R = args.P, args.R
    torch.manual_seed(args.seed)
    pr = synth.make_params(P, R, 1)
    g = torch.Generator().manual_seed(args.seed)
    # ---- cohort (identical on every rank: same seed) ---------------------------------------------------
    sizes = np.exp(np.random.default_rng(args.seed).uniform(np.log(args.min_n), np.log(args.max_n), args.patients)).astype(int)
    risk = torch.rand(args.patients, generator=g)
    event = (torch.rand(args.patients, generator=g) < 0.45).float()
    tbin = ((1.0 - risk) * R * (0.6 + 0.4 * torch.rand(args.patients, generator=g))).clamp(0, R - 1).floor()
    direction = torch.nn.functional.normalize(pr["prompt_features"][0], dim=0)

    def slides():
        for i in range(args.patients):
            x = synth.make_bag("g1", int(sizes[i]), 5000 + i)
            k = max(1, int(0.05 * sizes[i]))
            x[:k] += (6.0 * risk[i]) * direction * x[:k].norm(dim=-1, keepdim=True) / 25.0      # planted signal
            yield f"slide{i:04d}", x

    store_dir = os.path.join(tempfile.gettempdir(), f"vlsa_b200_store_{args.seed}_{args.patients}_{args.max_n}")
    if rank == 0 and not os.path.exists(os.path.join(store_dir, "index.json")):
        t0 = time.time()
        meta = build_store(store_dir, slides())
        print(f"[store] {meta['rows']} rows ({meta['rows'] * 2048 / 1e9:.2f} GB) in {time.time() - t0:.1f} s -> {store_dir}", flush=True)
    if world > 1:
        dist.barrier()
    store = PatchFeatureStore(store_dir)
    pids = [f"p{i:04d}" for i in range(args.patients)]
    pid2sids = {p: [f"slide{i:04d}"] for i, p in enumerate(pids)}
    pid2label = {p: (float(tbin[i]), float(event[i])) for i, p in enumerate(pids)}
    ds = WSIPatchSurvStore(store, pids, pid2sids, pid2label)

    class OneBagLoader:            # DataLoader(batch_size=1) of the reference: (idx [1], (feats [1,N,512], extra), label [1,2])
        dataset = ds
        def __len__(self): return len(ds)
        def __iter__(self):
            for i in range(len(ds)):
                idx, (feats, _), label = ds[i]
                yield idx, (feats.unsqueeze(0),), label.unsqueeze(0)

    cfg = {"task": "vlsa", "arch": "VLSA", "loss_type": "SurvIFMLE-SurvEMD", "opt_name": "adam", "opt_lr": 2e-4,
           "opt_weight_decay": 1e-5, "bp_every_batch": 32, "net_output_converter": "softmax"}
    net = VLSA(text_encoder_cfg={"name": "mahmoodlab/conch"},
               image_encoder_cfg=dict(name="VLFAN", dim_in=512, dim_hid=256, use_feat_proj=args.feat_proj, query="Text",
                                      num_query=P, gated_query=args.gated_query, query_pooling=args.query_pooling,
                                      pred_head="default", query_text_method="TaskRes", query_text_res_ratio=0.5),
               prompt_learner_cfg={"name": "CoOp"}, text_features=pr["text_features"],
               query_prompt_features=pr["prompt_features"], vlsa_api="CONCH", path_clip_model=None,
               query_neg_prompt_features=torch.nn.functional.normalize(torch.randn(1, 512, generator=g), dim=-1)
               if args.gated_query else None)
    if args.autograd_step:
        # (torch's foreach Adam is the implementation the bucket kernel follows to the last bit or two; its fused=True variant
        # drifts from both by ~lr within a handful of steps on this model)
        cfg.update(vlsa_fused_step=False, vlsa_bucket_adam=False, vlsa_torch_adam_fused=False)
    if args.handler_loop and args.cohort != "none":
        cfg["vlsa_device_cohort"] = args.cohort
    handler = VLSAHandler(cfg, net=net, device=dev)
    loader = OneBagLoader()
    bs = cfg["bp_every_batch"]
    sizes_all = [store.n_rows(pid2sids[p]) for p in pids]
    cohort = None
    if args.cohort != "none":
        # static partition of the split: patient i lives on rank i % world (every epoch's steps then find their bags there)
        mine_all = [i for i in range(len(pids)) if i % world == rank]
        cap = sum((sizes_all[i] + 15) // 16 * 16 for i in mine_all)
        cohort = DeviceCohort(dev, cap, layout=args.cohort)
    for epoch in range(args.epochs):
        torch.cuda.synchronize(); t0 = time.time()
        handler.net.train()
        losses = []
        if args.handler_loop:
            assert world == 1, "--handler-loop is the single-process loop of the reference"
            losses = handler._train_each_epoch(epoch, loader)["loss"]
            torch.cuda.synchronize(); dt = time.time() - t0
            if args.print_steps and rank == 0:
                print("   steps", [round(float(l), 6) for l in losses], flush=True)
            ds.skip_features(())                                # evaluation below reads every bag again
            pred = handler.test_model(handler.net, loader)["pred"]
            if handler.cohort is not None:
                ds.skip_features(handler.cohort.index.keys())
            inc = pred["y_hat"].numpy()
            score = (inc * np.arange(R)[None, :]).sum(1)
            c = concordance_index(-score, pred["y"][:, 0].numpy(), pred["y"][:, 1].numpy())
            print(f"[epoch {epoch}] loss {np.mean(losses):.4f}  C-index {c:.3f}  train {dt:.3f} s "
                  f"({args.patients / dt:.0f} bags/s through VLSAHandler._train_each_epoch, cohort {cfg.get('vlsa_device_cohort')})", flush=True)
            continue
        if cohort is not None and epoch == 0:
            for i in mine_all:                               # the one upload of the run (epoch 0 pays the store reads and H2D)
                cohort.add(i, ds[i][1][0])
        for s0 in range(0, len(pids), bs):                   # one optimizer step = 32 patients (vlsa_handler.py:260-289)
            ids = list(range(s0, min(s0 + bs, len(pids))))
            # lazy bags: a rank only reads the patients of its own shard from the store
            xs = [(lambda i=i: ds[i][1][0].unsqueeze(0)) for i in ids]
            ys = [torch.tensor(pid2label[pids[i]]).reshape(1, 2) for i in ids]
            # sync=False: loss and predictions stay on the device, the epoch reads them back once (as _train_each_epoch does)
            if cohort is not None:
                loss, _ = handler.update_network_cached(cohort, ids, ys, sync=False)
            else:
                loss, _ = handler._update_network(xs, ys, sizes=[sizes_all[i] for i in ids], sync=False)
            losses.append(loss)
        torch.cuda.synchronize(); dt = time.time() - t0
        losses = [float(l) for l in losses]
        if args.print_steps and rank == 0:
            print("   steps", [round(l, 6) for l in losses], flush=True)
        pred = handler.test_model(handler.net, loader)["pred"]
        inc = pred["y_hat"].numpy()                                            # incidence function [n, R]
        score = (inc * np.arange(R)[None, :]).sum(1)                           # expected time bin: low = high risk
        c = concordance_index(-score, pred["y"][:, 0].numpy(), pred["y"][:, 1].numpy())
        if rank == 0:
            print(f"[epoch {epoch}] loss {np.mean(losses):.4f}  C-index {c:.3f}  train {dt:.2f} s "
                  f"({args.patients / dt:.0f} bags/s incl. store reads and H2D, {world} GPU(s))", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
