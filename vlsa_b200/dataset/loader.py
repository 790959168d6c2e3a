"""Pinned, double-buffered host->device feeder for packed bags.

Replaces the reference's ``DataLoader(batch_size=1) -> feats.cuda()`` hop (dataset/PatchWSI.py:197-215,
runner/vlsa_handler.py:205: a synchronous copy from pageable memory per bag, plus ``empty_cache()`` per
step) with: bags of one step concatenated into ONE pinned staging buffer, ONE async H2D copy on a
dedicated copy stream into a device ring slot, and a CUDA event the compute stream waits on.  With
``depth`` >= 2 the copy of step i+1 overlaps the kernels of step i, so end-to-end throughput is the PCIe
rate rather than copy + compute.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Iterator, Sequence

import numpy as np
import torch

from .. import ops


@dataclass
class PackedBatch:
    X: torch.Tensor                 # device [total_rows, D]
    plan: "ops.BagPlan"
    labels: torch.Tensor | None     # device [B, 2] (t, e) or None
    index: torch.Tensor | None      # host   [B] dataset indices or None
    ready: torch.cuda.Event         # recorded on the copy stream after the H2D copy
    slot: int

    def wait(self, stream: torch.cuda.Stream | None = None) -> "PackedBatch":
        (stream or torch.cuda.current_stream()).wait_event(self.ready)
        return self


def pack_bags(bags: Sequence[torch.Tensor], out: torch.Tensor | None = None) -> tuple[torch.Tensor, list[int]]:
    """Concatenate host bags ([N_i, D] or [1, N_i, D]) into one (pinned) [sum N_i, D] tensor."""
    flat = [b[0] if b.dim() == 3 else b for b in bags]
    sizes = [int(b.shape[0]) for b in flat]
    total = sum(sizes)
    D = flat[0].shape[1] if flat else ops.D_FEAT
    dtype = flat[0].dtype if flat else torch.float32
    if out is None or out.shape[0] < total or out.dtype != dtype:
        out = torch.empty(max(total, 1), D, dtype=dtype)
        if torch.cuda.is_available():
            out = out.pin_memory()
    at = 0
    for b, n in zip(flat, sizes):
        out[at:at + n].copy_(b)
        at += n
    return out, sizes


class AsyncBagLoader:
    """Iterate over steps (lists of host bags) yielding device-resident ``PackedBatch`` objects.

    ``source`` yields ``(bags, labels, index)`` with ``bags`` a list of CPU tensors [N_i, D] (fp32 or
    bf16), ``labels`` a [B, 2] tensor of (time bin, event) or None.  If ``source`` yields already packed
    pinned tensors ``(X_pinned, sizes, labels, index)`` the staging copy is skipped.
    """

    def __init__(self, source: Iterable, device, depth: int = 2, max_rows: int | None = None,
                 dtype: torch.dtype = torch.float32, cohort=None):
        """``cohort``: a ``DeviceCohort``.  Every step is then copied straight into its final place in the cohort buffer
        (no ring slot is used, nothing to release) and registered under its ``index`` entries (dataset indices, required);
        the batch's plan is the cohort's row-range plan, so epoch 1 already runs on the resident rows."""
        self.source = source
        self.cohort = cohort
        self.device = torch.device(device)
        self.depth = max(2, int(depth))
        self.copy_stream = torch.cuda.Stream(device=self.device)
        if cohort is not None:
            # the cohort buffer comes from the current stream's allocator pool: whatever used that memory before is
            # ordered on the current stream, the first writer of the rows is the copy stream
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        self.dtype = dtype
        self._dev = [None] * self.depth          # device ring slots
        self._pin = [None] * self.depth          # pinned staging per slot
        self._free = [None] * self.depth         # event: consumer finished with the slot
        self._copied = [None] * self.depth       # event: the H2D copy out of the slot's pinned staging buffer is done
        self._busy = [False] * self.depth        # slot handed to the consumer and not released yet
        self._max_rows = max_rows
        self.h2d_bytes = 0

    def _slot_buffers(self, slot: int, rows: int):
        cap = max(rows, self._max_rows or 0, 1)
        if self._dev[slot] is None or self._dev[slot].shape[0] < rows:
            self._dev[slot] = torch.empty(cap, ops.D_FEAT, dtype=self.dtype, device=self.device)
            # the block comes from the allocator pool of the CURRENT stream and may still be in use there by whoever
            # freed it; its first writer is the copy stream
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        return self._dev[slot]

    def release(self, batch: PackedBatch, stream: torch.cuda.Stream | None = None) -> None:
        """Mark the batch's ring slot reusable once ``stream`` (default: current) has consumed it.  Called
        automatically (on the current stream) when the iterator is asked for the batch after this one; call it
        yourself only when the batch was consumed on another stream."""
        ev = torch.cuda.Event()
        ev.record(stream or torch.cuda.current_stream(self.device))
        self._free[batch.slot] = ev
        self._busy[batch.slot] = False

    def __iter__(self) -> Iterator[PackedBatch]:
        pending: list[PackedBatch] = []
        it = iter(self.source)
        slot = 0
        exhausted = False
        while True:
            while not exhausted and len(pending) < self.depth - 1:
                try:
                    item = next(it)
                except StopIteration:
                    exhausted = True
                    break
                pending.append(self._stage(item, slot))
                slot = (slot + 1) % self.depth
            if not pending:
                return
            batch = pending.pop(0)
            self._busy[batch.slot] = True
            yield batch
            # the consumer asked for the next batch: everything it enqueued on the current stream so far is what used
            # this one (an explicit release() on another stream has already cleared the flag)
            if self._busy[batch.slot]:
                self.release(batch)

    def _stage(self, item, slot: int) -> PackedBatch:
        if len(item) == 4 and isinstance(item[1], (list, tuple, np.ndarray)) and isinstance(item[0], torch.Tensor) \
                and item[0].dim() == 2:
            host, sizes, labels, index = item                       # pre-packed pinned tensor
            sizes = [int(s) for s in sizes]
        else:
            bags, labels, index = item
            if self._copied[slot] is not None:
                self._copied[slot].synchronize()                     # the previous H2D copy still reads this staging buffer
            host, sizes = pack_bags(bags, self._pin[slot])
            self._pin[slot] = host
        if self.cohort is None and self._busy[slot]:
            raise RuntimeError("AsyncBagLoader: ring slot reused while its batch is still held by the consumer "
                               "(raise depth= or release() the batch)")
        rows = sum(sizes)
        if self.cohort is not None:
            if index is None:
                raise ValueError("AsyncBagLoader(cohort=...) needs the dataset indices of every step as cohort keys")
            keys = [int(i) for i in (index.tolist() if isinstance(index, torch.Tensor) else index)]
            dst = self.cohort.reserve_step(keys, sizes)
            # the cohort buffer was allocated on the consumer's stream; nothing there touches the freshly reserved rows
        else:
            dst = self._slot_buffers(slot, rows)
        with torch.cuda.stream(self.copy_stream):
            if self.cohort is None and self._free[slot] is not None:
                self.copy_stream.wait_event(self._free[slot])        # consumer done with this slot
            if rows:
                dst[:rows].copy_(host[:rows], non_blocking=True)
                self._copied[slot] = torch.cuda.Event()
                self._copied[slot].record(self.copy_stream)
            plan = self.cohort.plan(keys) if self.cohort is not None else ops.make_plan(sizes, self.device)
            lab = labels.to(self.device, non_blocking=True) if labels is not None else None
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        self.h2d_bytes += rows * ops.D_FEAT * host.element_size()
        if self.cohort is not None:
            return PackedBatch(self.cohort.X, plan, lab, index, ready, slot)
        return PackedBatch(dst[:rows], plan, lab, index, ready, slot)
