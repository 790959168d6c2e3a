"""``VLSAHandler`` on B200 — the call sites of the hot path (runner/vlsa_handler.py:87-151,189-345 and the
bits of runner/base_handler.py they need), re-designed around one packed launch per optimizer step.

Kept from the reference: the flat config keys (``cfg_vlsa_conch.yaml`` parses unchanged; prefixes routed
like ``utils/func.py:136-147``), ``_update_network(xs, ys) -> (loss, preds)``, ``calc_objective_loss``,
``test_model`` returning ``{'pred': {'y','raw_y_hat','y_hat','uid'}}``, Adam with decay / no-decay groups
(optim/optim_factory.py:25-37), checkpoint = ``{'epoch','model','optimizer'}`` with the reference's
state-dict key names.  Changed on purpose: the 32 bags of a step are ONE varlen launch, the
text features are evaluated once per step instead of once per bag, and with torch.distributed initialised
the bags shard across ranks with a single flat-bucket all-reduce per step.

Out of scope here (use the reference's own code): wandb, csv split loading, SurvivalEVAL metrics.
"""
from __future__ import annotations

import os
from typing import Sequence

import numpy as np
import torch
import torch.nn.functional as F

from .. import ops
from ..dataset.loader import pack_bags
from ..loss import SurvObjective
from ..model import VLSA, PromptAdapter, deepmil
from . import dist as vdist
from .optim import BucketAdam


def fetch_kws(d: dict, prefix: str = "") -> dict:
    """utils/func.py:136-147: strip `prefix_` from matching keys."""
    if not prefix:
        return dict(d)
    pre = prefix + "_"
    return {k[len(pre):]: v for k, v in d.items() if k.startswith(pre)}


def create_output_converter(converter=None):
    """utils/func.py:40-48."""
    if converter == "sigmoid":
        return torch.sigmoid
    if converter == "softmax":
        return lambda x: F.softmax(x, dim=-1)
    return lambda x: x


def param_groups_weight_decay(model: torch.nn.Module, weight_decay: float):
    """optim/optim_factory.py:25-37 (`add_weight_decay`): no weight decay on parameters with ONE dimension and on biases.
    `len(param.shape) == 1` as in the reference: the 0-dim `logit_scale` is decayed there, so it is here."""
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        if len(p.shape) == 1 or name.endswith(".bias"):
            no_decay.append(p)
        else:
            decay.append(p)
    return [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": weight_decay}]


class VLSAHandler:
    def __init__(self, cfg: dict, net: VLSA | None = None, *, text_features=None, query_prompt_features=None,
                 device=None, balance_shards: bool = True):
        self.cfg = cfg
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        assert cfg.get("task", "vlsa") == "vlsa" and cfg.get("arch", "VLSA") == "VLSA"      # vlsa_handler.py:33-41
        assert cfg.get("net_output_converter", "softmax") == "softmax", "VLSA needs the softmax converter"
        if net is None:
            img_cfg = fetch_kws(cfg, "vlsa_img_encoder")
            txt_cfg = fetch_kws(cfg, "vlsa_txt_encoder")
            pmt_name = cfg.get("vlsa_pmt_learner_name", "CoOp")
            pmt_cfg = fetch_kws(cfg, "vlsa_pmt_learner_" + pmt_name.lower())
            pmt_cfg["name"] = pmt_name
            net = VLSA(txt_cfg, img_cfg, pmt_cfg, text_features=text_features,
                       query_prompt_features=query_prompt_features, vlsa_api=cfg.get("vlsa_api", "CONCH"),
                       path_clip_model=cfg.get("path_clip_model"))
        self.net = net.to(self.device)
        if cfg.get("vlsa_frozen_logit_scale", False):
            self.net.logit_scale.requires_grad_(False)
        if cfg.get("vlsa_img_encoder_frozen", False):
            for p in self.net.mil_encoder.parameters():
                p.requires_grad_(False)
        # losses: 'SurvIFMLE-SurvEMD' with per-loss weights (cfg_vlsa_conch.yaml:102-105)
        names = [n for n in str(cfg.get("loss_type", "SurvIFMLE-SurvEMD")).split("-") if n]
        for n in names:
            if n not in ("SurvIFMLE", "SurvEMD", "QueryDiv"):
                raise NotImplementedError(f"loss {n} is not part of the accelerated VLSA path")
        self.loss_weight = {n: float(cfg.get(f"loss_{n.lower()}_weight", 1.0)) for n in names}
        # 'QueryDiv' (runner/vlsa_handler.py:181-187,255-256 + model/deepmil.py:157-168): a regulariser on the P query rows alone,
        # once per optimizer step — N-independent, so it stays a [P, 512] torch expression next to the fused step
        self.query_div_kws = ({k: v for k, v in fetch_kws(cfg, "loss_querydiv").items() if k != "weight"}
                              if "QueryDiv" in names else None)
        # loss options the fused kernel does not implement must not be ignored silently (loss/loss_surv_ext.py:58-69,
        # loss/loss_surv.py:127-143): the shipped configs use p = 2, raw distance, mean reduction, eps 1e-7
        for key, ok in (("loss_survemd_p", lambda v: int(v) == 2), ("loss_survemd_raw_distance", lambda v: bool(v)),
                        ("loss_survemd_reduction", lambda v: v == "mean"), ("loss_survifmle_reduction", lambda v: v == "mean"),
                        ("loss_survemd_eps", lambda v: abs(float(v) - 1e-7) < 1e-12)):
            if key in cfg and not ok(cfg[key]):
                raise NotImplementedError(f"{key}={cfg[key]!r} is not implemented by the fused survival loss kernel")
        self.objective = SurvObjective(self.loss_weight.get("SurvIFMLE", 0.0), self.loss_weight.get("SurvEMD", 0.0),
                                       alpha=float(cfg.get("loss_survifmle_alpha", 0.0)),
                                       eps=float(cfg.get("loss_survifmle_eps", 1e-7)))
        self.output_converter = create_output_converter(cfg.get("net_output_converter", "softmax"))
        assert cfg.get("opt_name", "adam") == "adam", "only Adam is wired (cfg_vlsa_conch.yaml:111)"
        self.rank, self.world_size = vdist.world()
        vdist.broadcast_module(self.net)          # identical replicas before the first step
        self.balance_shards = balance_shards
        # tail = (total, ifmle, emd): the loss kernel of the fused step writes its three values there; segments 16-byte aligned
        # so that the backward kernels write W / bias / logit_scale gradients straight into the bucket
        self.bucket = vdist.FlatBucket(self.net.parameters(), extra=3, align=4)
        self.bucket.attach()                      # gradients live in the all-reduce bucket: no pack / unpack copies
        groups = param_groups_weight_decay(self.net, float(cfg.get("opt_weight_decay", 1e-5)))
        if self.device.type == "cuda" and cfg.get("vlsa_bucket_adam", True):
            # one launch over the bucket, untouched parameters skipped on the device (runner/optim.py)
            self.optimizer = BucketAdam(groups, self.bucket, lr=float(cfg.get("opt_lr", 2e-4)))
        else:
            self.optimizer = torch.optim.Adam(groups, lr=float(cfg.get("opt_lr", 2e-4)),
                                              fused=self.device.type == "cuda" and bool(cfg.get("vlsa_torch_adam_fused", True)))
        self.fused_step = bool(cfg.get("vlsa_fused_step", True))
        # 'rows' | 'split16': `_train_each_epoch` keeps every bag it has seen resident in HBM (dataset/cohort.py) and steps from there
        self.cohort_layout = cfg.get("vlsa_device_cohort") or None
        if self.cohort_layout not in (None, "rows", "split16"):
            raise ValueError("vlsa_device_cohort is 'rows' or 'split16'")
        self.cohort = None
        self._fused = ops.FusedTrainStep()
        self._flags_known = None                  # (this rank's touched pattern, reduced flags) of the last synchronous step

    # ------------------------------------------------------------------------------------------------
    def calc_objective_loss(self, raw_pred, label, norm: int | None = None):
        """runner/vlsa_handler.py:241-258, fused: softmax + weighted SurvIFMLE + SurvEMD (+ gradient)."""
        t, e = label[:, 0], label[:, 1]
        total, _, _, _ = self.objective(raw_pred, t, e, self.net.logit_scale, norm=norm)
        return total

    def _pack_local(self, xs: Sequence[torch.Tensor], idx: Sequence[int]):
        bags = [xs[i][0] if xs[i].dim() == 3 else xs[i] for i in idx]
        sizes = [int(b.shape[0]) for b in bags]
        if bags and not bags[0].is_cuda:
            # two pinned staging buffers, reused alternately: the async copy of one step may still be reading its
            # buffer while the next step is packed (cudaHostAlloc per step would dominate small steps)
            self._pin_slot = 1 - getattr(self, "_pin_slot", 0)
            pins = self.__dict__.setdefault("_pins", [None, None])
            ev = self.__dict__.setdefault("_pin_events", [None, None])
            if ev[self._pin_slot] is not None:
                ev[self._pin_slot].synchronize()
            host, _ = pack_bags(bags, pins[self._pin_slot])
            pins[self._pin_slot] = host
            X = host[: sum(sizes)].to(self.device, non_blocking=True)
            if self.device.type == "cuda":
                ev[self._pin_slot] = torch.cuda.Event()
                ev[self._pin_slot].record()
        elif bags:
            X = torch.cat(bags, 0) if len(bags) > 1 else bags[0].contiguous()
        else:
            X = torch.empty(0, ops.D_FEAT, device=self.device)
        return X, ops.make_plan(sizes, self.device)

    def _labels(self, ys, mine: Sequence[int] | None = None) -> torch.Tensor:
        """Labels of this rank's bags as [2, B_local] int64 on the device: row 0 = time bin, row 1 = event indicator
        (selected and converted where the labels live: one small copy)."""
        if isinstance(ys, torch.Tensor):
            lab = ys.reshape(-1, 2)
            if mine is not None and len(mine) != lab.shape[0]:
                lab = lab[torch.as_tensor(list(mine), device=lab.device, dtype=torch.long)]
        else:
            picked = ys if mine is None or len(mine) == len(ys) else [ys[i] for i in mine]
            lab = torch.cat(list(picked), dim=0).reshape(-1, 2) if len(picked) else torch.zeros(0, 2)
        if lab.is_cuda or self.device.type != "cuda":
            return lab.t().to(torch.int64).contiguous().to(self.device, non_blocking=True)
        # host labels: through the pinned ring (a copy from pageable memory would synchronise the stream)
        host = lab.t().to(torch.int64).contiguous().numpy()
        n = host.shape[1]
        return ops.upload_small(lambda st: st.__setitem__(slice(None), host.reshape(-1)), 2 * n, self.device).view(2, n)

    def step_packed(self, X: torch.Tensor, plan: "ops.BagPlan", labels: torch.Tensor, n_sample: int | None = None):
        """One optimizer step on this rank's bags, already packed on the device (`X`, `plan`), with their labels [B, 2] (or the
        [2, B] int64 of `_labels`); `n_sample` = bags of the step over ALL ranks (default: this rank's).  No host
        synchronisation: returns (loss, raw predictions of the local bags) as device tensors."""
        B = plan.num_bags
        if not (labels.dtype == torch.int64 and labels.dim() == 2 and labels.shape[0] == 2 and labels.is_cuda):
            labels = self._labels(labels)
        mine = list(range(B))
        if n_sample is None or (self.world_size == 1 and n_sample == B):
            return self._step(X, plan, labels, mine, B, sync=False)
        loss, _ = self._step(X, plan, labels, mine, n_sample, sync=False, gather_preds=False)
        return loss, None

    def _update_network(self, xs, ys, sizes: Sequence[int] | None = None, sync: bool = True):
        """One optimizer step on the bags `xs` (list of [1,N_i,512]) with labels `ys` (list of [1,2]).

        `xs[i]` may also be a zero-argument callable returning the bag (then `sizes` gives the N_i): only the bags of
        this rank's shard are fetched, so with a lazy dataset every rank reads 1/world of the step from storage.
        Returns (batch loss, raw predictions [n, R]) — a float and a host tensor as in the reference
        (runner/vlsa_handler.py:262-289), or with `sync=False` two DEVICE tensors and no host synchronisation in the step
        (the epoch loop converts them once, at its end)."""
        n_sample = len(xs)
        if sizes is None:
            sizes = [int(x.shape[-2]) for x in xs]
        mine = vdist.shard_indices(sizes, self.rank, self.world_size, self.balance_shards)
        X = plan = None
        if mine:
            local = {i: (xs[i]() if callable(xs[i]) else xs[i]) for i in mine}
            X, plan = self._pack_local(local, mine)
        return self._step(X, plan, self._labels(ys, mine), mine, n_sample, sync)

    def update_network_cached(self, cohort, keys: Sequence, ys, sync: bool = True):
        """`_update_network` on bags that are already resident in a `DeviceCohort` (keys = their cohort keys, e.g. the
        dataset indices): no staging, no H2D of rows — the step's plan points into the cohort buffer.  With several
        ranks the cohort is partitioned STATICALLY (every bag lives in exactly one rank's cohort, e.g. key % world ==
        rank, LPT-balanced over the whole split): a rank processes the bags of the step it holds."""
        n_sample = len(keys)
        mine = [i for i, k in enumerate(keys) if k in cohort]
        if self.world_size == 1 and len(mine) != n_sample:
            raise KeyError("update_network_cached: a bag of the step is not in the cohort")
        plan = cohort.plan([keys[i] for i in mine]) if mine else None
        return self._step(cohort.X if mine else None, plan, self._labels(ys, mine), mine, n_sample, sync)

    def _fused_ok(self, attached: bool | None = None) -> bool:
        """The step can run as three C calls (ops.FusedTrainStep): shipped VLFAN shape (mean over P + Linear adapter, raw rows
        in) and gradients attached to the bucket.  Everything else takes the autograd path."""
        enc = self.net.mil_encoder
        return (self.fused_step and self.device.type == "cuda" and isinstance(enc, deepmil.VLFAN) and enc.fused_tail
                and (self.bucket.attached() if attached is None else attached)
                and not (self.objective.w_ifmle == 0.0 and self.objective.w_emd == 0.0))

    def _taskres_residual(self):
        """The residual rows of a plain (ungated) TaskRes prompt adapter, whose gradient is `ratio * dQ` — or None."""
        enc = self.net.mil_encoder
        adapter = enc.Q if isinstance(enc.Q, PromptAdapter) else None
        if adapter is not None and adapter.method == "TaskRes" and not enc.gated_query:
            return adapter, adapter.residual_features
        return None, None

    def _step_overwrites_all(self) -> bool:
        """True when the fused step WRITES (not accumulates) the gradient of every parameter in the bucket — W, bias, logit_scale
        by the kernels, the TaskRes residual by `ratio * dQ` — so that the bucket needs no memset before the step."""
        enc = self.net.mil_encoder
        _, res = self._taskres_residual()
        written = {id(enc.visual_adapter.weight), id(enc.visual_adapter.bias), id(self.net.logit_scale)}
        if res is not None:
            written.add(id(res))
        return all(id(p) in written for p in self.bucket.params)

    def _fused_local_step(self, X, plan, t, e, n_sample, loss_out=None):
        """forward + loss + backward of this rank's bags; the three loss values land in `loss_out` (default: the bucket tail)."""
        net, enc, bk = self.net, self.net.mil_encoder, self.bucket
        adapter, res = self._taskres_residual()
        if res is not None and res.requires_grad and res.grad is not None:
            # plain TaskRes rows (prompt_adapter.py:125-126): Q = ratio * residual + prompt, so d residual = ratio * dQ — one
            # kernel into the attached view instead of a trip through the autograd engine (same arithmetic, same bits)
            with torch.no_grad():
                Qd, prenorm = enc.query_directions()
        else:
            res = None
            Qd, prenorm = enc.query_directions()               # small autograd graph over the prompt adapter (grad mode is on)
        T = net._text_features_for_kernels()
        W, bias, ls = enc.visual_adapter.weight, enc.visual_adapter.bias, net.logit_scale
        leaves = {"W": W, "bias": bias, "logit_scale": ls}
        out = self._fused(X, plan, Qd, W, bias, T, ls, t, e, w_ifmle=self.objective.w_ifmle, w_emd=self.objective.w_emd,
                          alpha=self.objective.alpha, eps=self.objective.eps, norm=n_sample, scale=enc.coattn_scale_float(),
                          q_prenorm=prenorm, grad_out={k: p.grad for k, p in leaves.items() if p.requires_grad},
                          loss_out=bk.tail if loss_out is None else loss_out)
        for k, buf in (("W", out["dW"]), ("bias", out["db"]), ("logit_scale", out["dls"])):
            p = leaves[k]
            if not p.requires_grad:
                continue
            if out["wrote"][k]:
                bk.mark_touched(p)                             # written in place of the zeroed view
            else:
                p.grad.copy_(buf.reshape(-1)[: p.numel()].view_as(p))      # the kernels' outputs are whole gradients
                bk.mark_touched(p)
        if res is not None:
            torch.mul(out["dQ"], float(adapter.res_ratio), out=res.grad)
            bk.mark_touched(res)
        roots = [(z, g) for z, g in ((Qd, out["dQ"]), (T, out["dT"])) if z.requires_grad]
        if roots:
            torch.autograd.backward([z for z, _ in roots], [g for _, g in roots])
        in_tail = out["loss"].data_ptr() == bk.tail.data_ptr()
        own = loss_out is not None and out["loss"].data_ptr() == loss_out.data_ptr()
        return out["logits"], (None if in_tail else (out["loss"][:3] if own else out["loss"][:3].clone())), in_tail

    def _query_div_term(self):
        """weight * query_div_loss() of this step, split evenly over the ranks (the bucket all-reduce sums it back)."""
        if self.query_div_kws is None or self.loss_weight.get("QueryDiv", 0.0) == 0.0:
            return None
        with torch.enable_grad():
            return (self.loss_weight["QueryDiv"] / self.world_size) * self.net.mil_encoder.query_div_loss(**self.query_div_kws)

    def _step(self, X, plan, lab, mine, n_sample, sync: bool = True, gather_preds: bool = True):
        """`lab`: labels of the local bags, [2, len(mine)] int64 on the device (`_labels`); `mine`: their positions among the
        `n_sample` bags of the step."""
        att = self.bucket.attached()                               # asked once per step (the check walks every parameter)
        fused = bool(mine) and self._fused_ok(att)
        single = self.world_size == 1
        if fused and self._step_overwrites_all():
            self.bucket.begin_step()                               # every gradient is overwritten below: no memset
        else:
            self.bucket.zero()
            att = self.bucket.attached()
            fused = bool(mine) and self._fused_ok(att)
        loss_in_tail = False
        local_loss = None
        extra = self._query_div_term()
        if mine:
            if fused:
                # one process: nothing rides an all-reduce, the losses go to a tensor of their own (no copy out of the bucket)
                own_loss = torch.empty(3, dtype=torch.float32, device=self.device) if single else None
                with torch.enable_grad():
                    local_pred, local_loss, loss_in_tail = self._fused_local_step(X, plan, lab[0], lab[1], n_sample, own_loss)
                if extra is not None:
                    extra.backward()                               # accumulates on top of what the kernels wrote
                    (self.bucket.tail if loss_in_tail else local_loss)[0:1].add_(extra.detach().reshape(1))
                    extra = None
            else:
                logits, _, _, _ = self.net.forward_packed(X, plan)                   # [B_local, R]
                pred_loss = self.calc_objective_loss(logits, lab.t(), norm=n_sample)  # sum_local / n_sample
                if extra is not None:
                    pred_loss = pred_loss + extra
                    extra = None
                pred_loss.backward()
                local_loss, local_pred = pred_loss.detach().reshape(1), logits.detach()
        else:
            local_pred = torch.zeros(0, self.net.forward_text_only().shape[0], device=self.device)
        if extra is not None:                                      # a rank without bags still owes its share of the regulariser
            extra.backward()
            local_loss = extra.detach().reshape(1)
        # the one exchange of the step: gradients + losses in one flat bucket
        if not fused:
            att = att and self.bucket.attached()                   # (the autograd path may have replaced a .grad)
        bucket_adam = isinstance(self.optimizer, BucketAdam)
        # one process, every parameter has a gradient: nobody needs the flags (with several ranks they always travel: a rank
        # without bags learns from the reduced flags which parameters the others touched)
        need_flags = not (single and bucket_adam and bool(mine) and all(self.bucket._touched))
        if fused and single and not need_flags:
            loss_dev = local_loss[0]                               # the step's own tensor: nothing to pack, reduce or copy
            self.optimizer.step(use_flags=False, attached=att)
            return self._finish_step(loss_dev, local_pred, mine, n_sample, sync, gather_preds)
        self.bucket.pack(local_loss, extra_in_place=loss_in_tail, attached=att, with_flags=need_flags)
        self.bucket.all_reduce()
        self.bucket.unpack(attached=att)
        n_par = len(self.bucket.params)
        if bucket_adam:
            # the kernel reads the reduced flags itself: untouched parameters are skipped on the device
            self.optimizer.step(use_flags=need_flags, attached=att)
            loss_dev = self.bucket.tail[0].clone()
        else:
            pattern = (tuple(self.bucket._touched), bool(mine))
            known = self._flags_known
            if not sync and known is not None and known[0] == pattern and mine:
                # which parameters a step reaches is a property of the model, not of the step: with the same local pattern as
                # the last synchronous step (and bags on this rank) the reduced flags are the same — no device -> host read
                flags = known[1]
                loss_dev = self.bucket.tail[0].clone()
            else:
                # losses + per-parameter "touched" flags; a real copy (on a CPU device `.cpu()` would alias the live bucket)
                tail = self.bucket.flat[-(3 + n_par):].detach().to("cpu", copy=True)
                flags = tail[3:]
                self._flags_known = (pattern, flags)
                loss_dev = tail[0]
            self.bucket.drop_untouched(flags)
            self.optimizer.step()
        return self._finish_step(loss_dev, local_pred, mine, n_sample, sync, gather_preds)

    def _finish_step(self, loss_dev, local_pred, mine, n_sample, sync, gather_preds):
        if not gather_preds:
            preds = None
        elif self.world_size == 1 and len(mine) == n_sample:
            preds = local_pred
        else:
            preds = vdist.all_reduce_rows(local_pred, mine, n_sample)
        if sync:
            return float(loss_dev), (preds.cpu() if preds is not None else None)
        return loss_dev, preds

    def _to_cohort(self, data_idx, x) -> int:
        """Key of the bag in the handler's device cohort; its rows are uploaded (and, for 'split16', packed) the first time the
        bag is seen.  A bag that is already resident may arrive with an empty feature tensor (the dataset skipped the read).
        With several ranks every rank keeps the bags of its own shard only (key % world == rank)."""
        from ..dataset.cohort import DeviceCohort
        key = int(data_idx.reshape(-1)[0]) if isinstance(data_idx, torch.Tensor) else int(data_idx)
        if self.cohort is None:
            self.cohort = DeviceCohort(self.device, int(self.cfg.get("vlsa_device_cohort_rows", 1 << 20)), layout=self.cohort_layout)
        if self.world_size > 1 and key % self.world_size != self.rank:
            return key
        bag = x[0] if x.dim() == 3 else x
        if key not in self.cohort:
            if bag.shape[0] == 0:
                raise KeyError(f"bag {key} arrived without rows but is not resident in the device cohort")
            need = self.cohort.rows + (bag.shape[0] + 15) // 16 * 16
            self.cohort.grow(need)
            self.cohort.add(key, bag)
        return key

    def _train_each_epoch(self, epoch, train_loader, name_loader="train"):
        self.net.train()
        bp_every_batch = int(self.cfg.get("bp_every_batch", 32))
        all_raw_pred, all_gt, all_idx, losses = [], [], [], []
        idx_c, x_c, y_c = [], [], []
        num_samples = len(train_loader)
        for i_batch, (data_idx, data_x, data_y) in enumerate(train_loader, start=1):
            x_c.append(self._to_cohort(data_idx, data_x[0]) if self.cohort_layout else data_x[0])
            y_c.append(data_y)
            idx_c.append(data_idx)
            if i_batch % bp_every_batch == 0 or i_batch == num_samples:
                if self.cohort_layout:
                    # x_c holds cohort keys: the step is a row-range plan into the resident buffer (no staging, no H2D of rows)
                    batch_loss, batch_pred = self.update_network_cached(self.cohort, x_c, y_c, sync=False)
                else:
                    batch_loss, batch_pred = self._update_network(x_c, y_c, sync=False)   # device tensors: one sync per epoch
                losses.append(batch_loss)
                all_raw_pred.append(batch_pred)
                all_gt.append(torch.cat([y.reshape(1, 2) for y in y_c], dim=0).cpu())
                all_idx.append(torch.cat([i.reshape(-1) for i in idx_c], dim=0).cpu())
                idx_c, x_c, y_c = [], [], []
        if self.cohort_layout and hasattr(getattr(train_loader, "dataset", None), "skip_features"):
            # the dataset may stop reading what is resident now (dataset/store.py: WSIPatchSurvStore.skip_features)
            train_loader.dataset.skip_features(self.cohort.index.keys())
        raw = torch.cat(all_raw_pred, 0).cpu()
        losses = torch.stack([torch.as_tensor(l, dtype=torch.float32).reshape(()).to(raw.device) for l in losses]).tolist() if losses else []
        return {"pred": {"y": torch.cat(all_gt, 0), "raw_y_hat": raw, "y_hat": self.output_converter(raw),
                         "uid": torch.cat(all_idx, 0)}, "loss": losses}

    @torch.no_grad()
    def test_model(self, model, loader, loader_name="test", ckpt_path=None, bags_per_launch: int = 32):
        """runner/vlsa_handler.py:315-345: every bag -> raw prediction + incidence; here `bags_per_launch` bags
        share one launch and one evaluation of the text features."""
        if ckpt_path is not None:
            net_ckpt = torch.load(ckpt_path, map_location=self.device)
            model.load_state_dict(net_ckpt["model"], strict=False)
        model.eval()
        all_idx, all_raw, all_pred, all_gt = [], [], [], []
        T = model.forward_text_only()
        buf_x, buf_y, buf_i = [], [], []

        def flush():
            if not buf_x:
                return
            X, plan = self._pack_local(buf_x, range(len(buf_x)))
            logits, _, _, inc = model.forward_packed(X, plan, T)
            all_raw.append(logits.cpu())
            all_pred.append(inc.cpu())
            all_gt.extend(buf_y)
            all_idx.extend(buf_i)
            buf_x.clear(); buf_y.clear(); buf_i.clear()

        for data_idx, data_x, data_y in loader:
            buf_x.append(data_x[0])
            buf_y.append(data_y.reshape(1, 2).cpu())
            buf_i.append(data_idx.reshape(-1).cpu())
            if len(buf_x) == bags_per_launch:
                flush()
        flush()
        return {"pred": {"y": torch.cat(all_gt, 0), "raw_y_hat": torch.cat(all_raw, 0),
                         "y_hat": torch.cat(all_pred, 0), "uid": torch.cat(all_idx, 0)}}

    @torch.no_grad()
    def test_model_cached(self, model, cohort, keys: Sequence, labels: Sequence | None = None, bags_per_launch: int = 32):
        """``test_model`` on bags that are resident in a ``DeviceCohort`` (keys = their cohort keys, in evaluation order): no
        loader, no staging, no H2D of rows — every launch draws ``bags_per_launch`` bags by a row-range plan.  Same result
        dict; ``labels[i]`` is the [2] / [1,2] label of ``keys[i]`` (``y`` is omitted when labels are not given)."""
        model.eval()
        T = model.forward_text_only()
        raw, pred = [], []
        for s0 in range(0, len(keys), bags_per_launch):
            logits, _, _, inc = model.forward_packed(cohort.X, cohort.plan(list(keys[s0:s0 + bags_per_launch])), T)
            raw.append(logits)
            pred.append(inc)
        out = {"raw_y_hat": torch.cat(raw, 0).cpu(), "y_hat": torch.cat(pred, 0).cpu(),
               "uid": torch.as_tensor([int(k) if isinstance(k, (int, np.integer)) else i for i, k in enumerate(keys)])}
        if labels is not None:
            out["y"] = torch.cat([torch.as_tensor(y, dtype=torch.float32).reshape(1, 2) for y in labels], 0)
        return {"pred": out}

    # ---- checkpoint (runner/base_handler.py:641-682) ---------------------------------------------------
    def save_model(self, path: str, epoch: int, module_filter: str | None = "prompt_encoder") -> None:
        state = {k: v for k, v in self.net.state_dict().items() if not (module_filter and module_filter in k)}
        if self.rank == 0:
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
            torch.save({"epoch": epoch, "model": state, "optimizer": self.optimizer.state_dict()}, path)

    def load_model(self, path: str):
        ckpt = torch.load(path, map_location=self.device)
        return self.net.load_state_dict(ckpt["model"], strict=False)
