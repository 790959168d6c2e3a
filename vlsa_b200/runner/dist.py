"""Host-side pieces of the multi-GPU path: bags are independent, so they shard across
ranks with NO data-path collective; per optimizer step there is exactly one exchange — an all-reduce
(SUM) of one flat gradient bucket (+ the scalar loss riding in the same bucket).

Everything here is device-agnostic torch.distributed code so that it is covered on CPU with gloo
(tests/test_dist_gloo.py) and runs unchanged over NCCL / NVLink on the B200 box.
"""
from __future__ import annotations

from typing import Iterable, Sequence

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(sizes: Sequence[int], rank: int, world_size: int, balance: bool = True) -> list[int]:
    """Indices of the bags rank `rank` processes out of one step's bags.

    balance=False: round robin, bag i -> rank i mod world.
    balance=True : longest-processing-time greedy on the row counts (deterministic: ties broken by index),
                   so ragged bags (1k..100k rows) give every GPU about the same number of rows to stream.
    Every rank computes the same assignment from the same `sizes`; no communication."""
    n = len(sizes)
    if world_size <= 1:
        return list(range(n))
    if not balance:
        return [i for i in range(n) if i % world_size == rank]
    order = sorted(range(n), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world_size
    owner = [0] * n
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += int(sizes[i])
    return [i for i in range(n) if owner[i] == rank]


class FlatBucket:
    """One contiguous fp32 buffer holding every trainable gradient (+ `extra` scalars at the tail).

    `pack()` copies the .grad tensors in (zeros where a parameter got no gradient), `all_reduce()` sums
    it over the ranks in ONE collective, `unpack()` writes the reduced gradients back.  For the BLCA model
    this is 284 161 floats = 1.14 MB: latency-bound, so one bucket and one launch.

    With `attach()` the parameters' .grad tensors ARE views of the bucket: autograd accumulates straight into
    it, `zero()` clears all gradients with one memset, and pack / unpack copy nothing (a step on small bags is
    launch-bound: this removes a dozen tiny kernels per step)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], extra: int = 1, align: int = 1):
        self.params = [p for p in params if p.requires_grad]
        self.sizes = [p.numel() for p in self.params]
        self.extra = extra
        # `align` (in floats): every segment starts at a multiple of it, so that kernels may write gradients straight into the
        # attached views with 128-bit stores (the handler asks for 4); the padding floats stay zero and ride the all-reduce
        self.offsets, at = [], 0
        for n in self.sizes:
            self.offsets.append(at)
            at += -(-n // align) * align
        self._grad_floats = at
        # tail = `extra` caller scalars, then one "received a gradient this step" flag per parameter (summed over the
        # ranks by the same all-reduce): a parameter nobody touched keeps grad = None for the optimizer, as after the
        # reference's zero_grad(set_to_none) — Adam must not decay its moments or apply weight decay to it
        total = self._grad_floats + extra + len(self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self._touched = [False] * len(self.params)
        self._hooks = []
        self._flag_stage = torch.zeros(len(self.params), dtype=torch.float32)
        if dev.type == "cuda":
            self._flag_stage = self._flag_stage.pin_memory()
        self._flag_cache: dict[tuple, torch.Tensor] = {}     # touched pattern -> its flags on the device (attached buckets)
        self._index = {id(p): i for i, p in enumerate(self.params)}

    @property
    def tail(self) -> torch.Tensor:
        """The caller's `extra` scalars."""
        n = len(self.flat) - self.extra - len(self.params)
        return self.flat[n:n + self.extra]

    @property
    def flags(self) -> torch.Tensor:
        return self.flat[len(self.flat) - len(self.params):]

    def _views(self):
        for p, n, at in zip(self.params, self.sizes, self.offsets):
            yield p, self.flat[at:at + n]

    def mark_touched(self, *params: torch.nn.Parameter) -> None:
        """A kernel wrote this step's gradient of `params` straight into their attached views (no autograd hook fired)."""
        for p in params:
            i = self._index.get(id(p))
            if i is not None:
                self._touched[i] = True

    def attached(self) -> bool:
        """True when every parameter's .grad is (still) its view of the bucket."""
        return bool(self._hooks) and all(p.grad is not None and p.grad.data_ptr() == self.flat.data_ptr() + 4 * at
                                         for p, at in zip(self.params, self.offsets))

    def attach(self) -> None:
        """Make every parameter's .grad a view of the bucket (keeps the current gradient values, if any)."""
        for p, seg in self._views():
            g = seg.view_as(p)
            if p.grad is not None and p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad)
            p.grad = g
        if not self._hooks:
            for i, p in enumerate(self.params):
                self._hooks.append(p.register_post_accumulate_grad_hook(lambda _p, i=i: self._touched.__setitem__(i, True)))

    def zero(self) -> None:
        """Zero all gradients.  Attached: one memset of the bucket (parameters detached by `drop_untouched` are
        re-attached); otherwise `grad = None` like zero_grad()."""
        self._touched = [False] * len(self.params)
        att = self.attached()
        if self._hooks and not att:
            for p, seg in self._views():
                if p.grad is None or p.grad.data_ptr() != seg.data_ptr():
                    p.grad = seg.view_as(p)
            att = True
        if att or (not self._hooks and all(p.grad is not None and p.grad.data_ptr() == seg.data_ptr() for p, seg in self._views())):
            self.flat.zero_()
        else:
            for p in self.params:
                p.grad = None

    def drop_untouched(self, flags_host) -> int:
        """After the all-reduce: `grad = None` for every parameter whose flag summed to zero over the ranks (nobody's
        backward reached it).  `flags_host` is a host copy of `self.flags`.  Returns how many were dropped."""
        n = 0
        for p, f in zip(self.params, flags_host.tolist()):
            if f == 0.0:
                p.grad = None
                n += 1
        return n

    def begin_step(self) -> None:
        """`zero()` without the memset, for steps whose kernels OVERWRITE every gradient of the bucket (the caller vouches for
        it): resets the touched marks only."""
        self._touched = [False] * len(self.params)

    def pack(self, extra_values: torch.Tensor | None = None, extra_in_place: bool = False, attached: bool | None = None,
             with_flags: bool = True) -> None:
        """`extra_in_place`: the caller's scalars were already written into `tail` (by a kernel).  `attached`: what
        `self.attached()` returned earlier in this step (the check walks every parameter; a step asks once).  `with_flags=False`:
        nobody will read the flags of this step (every parameter is known to have a gradient)."""
        if self._hooks and (self.attached() if attached is None else attached):
            # gradients already live here; the flags of a touched pattern are uploaded once and copied on the device after that
            if self.extra and not extra_in_place:
                if extra_values is None:
                    self.tail.zero_()
                else:
                    vals = extra_values.reshape(-1).to(self.flat.dtype)
                    self.tail.zero_() if vals.numel() < self.extra else None
                    self.tail[: vals.numel()].copy_(vals)
            if self.params and with_flags:
                key = tuple(self._touched)
                dev_flags = self._flag_cache.get(key)
                if dev_flags is None:
                    dev_flags = torch.tensor(key, dtype=torch.float32).to(self.flat.device)
                    self._flag_cache[key] = dev_flags
                self.flags.copy_(dev_flags, non_blocking=True)
            return
        for p, seg in self._views():
            if p.grad is None:
                seg.zero_()
            elif p.grad.data_ptr() != seg.data_ptr():             # attached gradients already live here
                seg.copy_(p.grad.reshape(-1))
        if self.extra:
            if extra_values is None:
                self.tail.zero_()
            else:
                vals = extra_values.reshape(-1).to(self.flat.dtype)
                self.tail.zero_() if vals.numel() < self.extra else None
                self.tail[: vals.numel()].copy_(vals)
        if self.params:
            if self._hooks:
                self._flag_stage.copy_(torch.tensor(self._touched, dtype=torch.float32))
            else:                                   # not attached: a parameter is touched iff it has a gradient
                self._flag_stage.copy_(torch.tensor([p.grad is not None for p in self.params], dtype=torch.float32))
            self.flags.copy_(self._flag_stage, non_blocking=True)

    def all_reduce(self) -> None:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)

    def unpack(self, attached: bool | None = None) -> None:
        if self._hooks and (self.attached() if attached is None else attached):
            return
        for p, seg in self._views():
            g = seg.view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            elif p.grad.data_ptr() != g.data_ptr():
                p.grad.copy_(g)


def broadcast_module(module: torch.nn.Module, src: int = 0) -> None:
    """Make every rank start from rank `src`'s parameters and buffers (one flat broadcast), as DDP does at
    construction: replicas that were initialised from different random states must not average gradients."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return
    tensors = [t for t in list(module.parameters()) + list(module.buffers()) if t.is_floating_point()]
    if not tensors:
        return
    flat = torch.cat([t.detach().reshape(-1).float() for t in tensors])
    dist.broadcast(flat, src=src)
    at = 0
    with torch.no_grad():
        for t in tensors:
            n = t.numel()
            t.copy_(flat[at:at + n].view_as(t).to(t.dtype))
            at += n


def all_reduce_rows(local_rows: torch.Tensor, local_idx: Sequence[int], n_total: int) -> torch.Tensor:
    """Assemble the [n_total, R] prediction matrix from every rank's rows (each row owned by one rank)."""
    out = torch.zeros(n_total, local_rows.shape[-1], dtype=local_rows.dtype, device=local_rows.device)
    if len(local_idx):
        out[torch.as_tensor(list(local_idx), device=local_rows.device)] = local_rows
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out
