"""``BucketAdam`` — the optimizer step of the path (runner/vlsa_handler.py:283; optim/optim_factory.py:25-37 builds
``torch.optim.Adam`` with decay / no-decay groups for cfg_vlsa_conch.yaml:111-118) as ONE kernel launch over the flat gradient
bucket (``vlsa_adam_step``).  Same update rule as ``torch.optim.Adam`` (L2 decay added to the gradient, bias-corrected moments,
amsgrad off); a parameter that received no gradient in a step is skipped exactly as torch skips ``grad is None`` — decided on
the device from the bucket's all-reduced "touched" flags, so a training step never reads the device back.

``state_dict()`` / ``load_state_dict()`` speak ``torch.optim.Adam``'s format (``{'state': {i: {'step', 'exp_avg',
'exp_avg_sq'}}, 'param_groups': [...]}``), so checkpoints written by either load into the other.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib, ops
from .dist import FlatBucket


class BucketAdam:
    def __init__(self, param_groups: list[dict], bucket: FlatBucket, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        self.bucket = bucket
        self.defaults = dict(lr=float(lr), betas=tuple(float(b) for b in betas), eps=float(eps), weight_decay=0.0, amsgrad=False,
                             maximize=False)
        self.param_groups = []
        for g in param_groups:
            grp = dict(self.defaults)
            grp.update(g)
            grp["params"] = list(grp["params"])
            self.param_groups.append(grp)
        index = {id(p): i for i, p in enumerate(bucket.params)}
        self._order = [p for g in self.param_groups for p in g["params"]]           # torch's numbering of the state
        if sorted(index[id(p)] for p in self._order) != list(range(len(bucket.params))):
            raise ValueError("BucketAdam: the parameter groups must hold exactly the bucket's parameters")
        dev = bucket.flat.device
        self.exp_avg = torch.zeros(bucket._grad_floats, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros_like(self.exp_avg)
        # per-tensor step counts, two arrays used alternately (the kernel reads one and writes the other)
        self._steps = torch.zeros(2, max(len(bucket.params), 1), dtype=torch.float32, device=dev)
        self._cur = 0
        self._seg_key = None
        self._segs = None
        self._max_n = max([p.numel() for p in bucket.params], default=0)

    # ---- device table of the segments (rebuilt when a learning rate / decay or a parameter's storage changes) ---------------
    def _segments(self):
        bk = self.bucket
        rows = []
        for g in self.param_groups:
            for p in g["params"]:
                rows.append((id(p), p.data_ptr(), float(g["lr"]), float(g["weight_decay"])))
        key = tuple(rows)
        if key != self._seg_key:
            rec = np.dtype([("param", np.uint64), ("offset", np.int64), ("n", np.int64), ("wd", np.float32), ("lr", np.float32)])
            assert rec.itemsize == _lib.lib().vlsa_adam_segment_bytes()
            table = np.zeros(len(bk.params), dtype=rec)
            by_id = {r[0]: r for r in rows}
            for i, (p, off) in enumerate(zip(bk.params, bk.offsets)):
                _, ptr, lr, wd = by_id[id(p)]
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise TypeError("BucketAdam: parameters must be contiguous fp32 tensors")
                table[i] = (ptr, off, p.numel(), wd, lr)
            self._segs = torch.from_numpy(table.view(np.uint8).copy()).to(bk.flat.device)
            self._seg_key = key
        return self._segs

    def step(self, use_flags: bool = True, attached: bool | None = None) -> None:
        """One update from the gradients in the bucket.  ``use_flags``: skip the parameters whose (all-reduced) flag is zero.
        ``attached``: what ``bucket.attached()`` returned earlier in this step."""
        bk = self.bucket
        if not bk.params:
            return
        if not (bk.attached() if attached is None else attached):
            raise RuntimeError("BucketAdam: the parameters' gradients must be attached to the bucket (FlatBucket.attach)")
        g0 = self.param_groups[0]
        for g in self.param_groups[1:]:
            if tuple(g["betas"]) != tuple(g0["betas"]) or g["eps"] != g0["eps"]:
                raise NotImplementedError("BucketAdam: betas / eps must be the same for every group")
        segs = self._segments()
        src, dst = self._steps[self._cur], self._steps[1 - self._cur]
        rc = _lib.lib().vlsa_adam_step(segs.data_ptr(), len(bk.params), self._max_n, bk.flat.data_ptr(), self.exp_avg.data_ptr(),
                                       self.exp_avg_sq.data_ptr(), src.data_ptr(), dst.data_ptr(),
                                       bk.flags.data_ptr() if use_flags else None, float(g0["betas"][0]), float(g0["betas"][1]),
                                       float(g0["eps"]), ops._stream())
        _lib.check(rc, "vlsa_adam_step")
        self._cur = 1 - self._cur
        # the kernel wrote the parameters through raw pointers: tell autograd (and every cache keyed on a tensor's version
        # counter, e.g. VLFAN.query_directions_cached) that they changed, as an in-place torch op would
        torch.autograd.graph.increment_version(bk.params)

    @property
    def step_count(self) -> torch.Tensor:
        """Steps every tensor of the bucket has taken (float, on the device)."""
        return self._steps[self._cur][: len(self.bucket.params)]

    def zero_grad(self, set_to_none: bool = True) -> None:
        self.bucket.zero()

    # ---- torch.optim.Adam-compatible checkpoints -----------------------------------------------------------------------------
    def _bucket_slot(self):
        return {id(p): (i, off, n) for i, (p, off, n) in enumerate(zip(self.bucket.params, self.bucket.offsets, self.bucket.sizes))}

    def state_dict(self) -> dict:
        slot = self._bucket_slot()
        steps = self.step_count.cpu()
        state, groups, at = {}, [], 0
        for g in self.param_groups:
            ids = []
            for p in g["params"]:
                i, off, n = slot[id(p)]
                if float(steps[i]) > 0:
                    state[at] = {"step": steps[i].clone(), "exp_avg": self.exp_avg[off:off + n].view_as(p).clone(),
                                 "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p).clone()}
                ids.append(at)
                at += 1
            groups.append({**{k: v for k, v in g.items() if k != "params"}, "params": ids})
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd: dict) -> None:
        slot = self._bucket_slot()
        if [len(g["params"]) for g in sd["param_groups"]] != [len(g["params"]) for g in self.param_groups]:
            raise ValueError("BucketAdam: the checkpoint's parameter groups do not match")
        for g, saved in zip(self.param_groups, sd["param_groups"]):
            for k, v in saved.items():
                if k != "params" and k in g:
                    g[k] = tuple(v) if k == "betas" else v
        self.exp_avg.zero_(); self.exp_avg_sq.zero_(); self._steps.zero_()
        for at, p in enumerate(self._order):
            st = sd["state"].get(at, sd["state"].get(str(at)))
            if st is None:
                continue
            i, off, n = slot[id(p)]
            self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            self._steps[self._cur, i] = float(st["step"])
        self._seg_key = None
