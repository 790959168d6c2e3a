"""Device-resident cohort: every patient bag of a split uploaded to HBM ONCE, steps drawn from it without a copy.

The reference re-reads and re-uploads every bag every epoch (dataset/PatchWSI.py:197-215: ``torch.load`` + ``cat`` +
``.float()``; runner/vlsa_handler.py:205: a synchronous pageable ``.cuda()``), so its training loop is bound by disk and
PCIe from the first epoch to the last.  A TCGA cohort is small next to a B200: 373 BLCA patients x 3-20k rows x 2 KB is
2-8 GB of the 180 GB of HBM.  ``DeviceCohort`` keeps the rows of all bags back to back in one device buffer and hands
out *row-range plans* (``ops.make_plan_ranges`` -> ``VLSA_ROWS_RANGES``): the kernels read the bags of a (shuffled)
optimizer step straight out of the cohort buffer, wherever they lie — no gather, no H2D beyond the 16 bytes per bag of
the range table.  Epoch 1 fills the cohort through the asynchronous loader (``AsyncBagLoader(cohort=...)`` copies each
step's pinned rows directly into their final place); from epoch 2 on a step costs what the kernels cost.

``layout="split16"`` additionally stores the rows the way the tensor cores read them (``vlsa_split16_pack``: pre-split fp16
(hi, lo) planes, pre-swizzled 16-row records, 2 056 bytes per row instead of 2 048): for P > 5 the fp32 -> fp16 split is
what keeps the fp32-row tensor-core kernel at ~0.7 of the HBM roofline, and a resident cohort only has to pay it once.  Bags
then start at multiples of 16 (padded) rows; results are bit-identical to the fp32-row tensor-core kernel's.
"""
from __future__ import annotations

from typing import Hashable, Iterable, Sequence

import numpy as np
import torch

from .. import ops


class DeviceCohort:
    def __init__(self, device, capacity_rows: int, dtype: torch.dtype = torch.float32, layout: str = "rows"):
        """``capacity_rows``: rows the cohort can hold (``layout="split16"``: count every bag rounded up to 16 rows)."""
        if layout not in ("rows", "split16") or (layout == "split16" and dtype != torch.float32):
            raise ValueError("layout is 'rows' (fp32 / bf16) or 'split16' (fp32 only)")
        self.device = torch.device(device)
        self.dtype = dtype
        self.layout = layout
        cols = ops.SPLIT16_COLS if layout == "split16" else ops.D_FEAT
        cap = max(int(capacity_rows), 1)
        if layout == "split16":
            cap = (cap + 15) // 16 * 16
        self.X = torch.empty(cap, cols, dtype=dtype, device=self.device)
        self.rows = 0                                    # rows handed out so far
        self.index: dict[Hashable, tuple[int, int]] = {}   # key -> (first row, one past the last row)

    def grow(self, min_rows: int) -> None:
        """Make room for at least ``min_rows`` rows in total (at least doubling): one device-to-device copy of what is resident."""
        if min_rows <= self.X.shape[0]:
            return
        cap = max(int(min_rows), 2 * self.X.shape[0])
        if self.layout == "split16":
            cap = (cap + 15) // 16 * 16
        new = torch.empty(cap, self.X.shape[1], dtype=self.X.dtype, device=self.device)
        new[: self.rows].copy_(self.X[: self.rows])
        self.X = new

    # ---- filling ------------------------------------------------------------------------------------
    def reserve(self, key: Hashable, n_rows: int) -> torch.Tensor:
        """Claim the next ``n_rows`` rows for bag ``key`` and return the view to copy its rows into."""
        if self.layout != "rows":
            raise RuntimeError("a split16 cohort is filled with add(): its rows are packed, not copied")
        if key in self.index:
            raise KeyError(f"bag {key!r} is already in the cohort")
        if self.rows + n_rows > self.X.shape[0]:
            raise MemoryError(f"cohort capacity exceeded: {self.rows} + {n_rows} > {self.X.shape[0]} rows")
        self.index[key] = (self.rows, self.rows + n_rows)
        self.rows += n_rows
        return self.X[self.index[key][0]:self.index[key][1]]

    def reserve_step(self, keys: Sequence[Hashable], sizes: Sequence[int]) -> torch.Tensor:
        """Claim one contiguous span for the bags of a packed step (in order); returns the [sum sizes, 512] view."""
        start = self.rows
        for k, n in zip(keys, sizes):
            self.reserve(k, int(n))
        return self.X[start:self.rows]

    def add(self, key: Hashable, bag: torch.Tensor) -> None:
        """Upload one bag ([N, 512] or [1, N, 512], host or device) into the cohort (current stream)."""
        b = bag[0] if bag.dim() == 3 else bag
        if self.layout == "split16":
            if key in self.index:
                raise KeyError(f"bag {key!r} is already in the cohort")
            n, padded = b.shape[0], (b.shape[0] + 15) // 16 * 16
            if self.rows + padded > self.X.shape[0]:
                raise MemoryError(f"cohort capacity exceeded: {self.rows} + {padded} > {self.X.shape[0]} rows")
            ops.split16_pack(b.to(self.device, torch.float32, non_blocking=True).contiguous(), self.X, self.rows)
            self.index[key] = (self.rows, self.rows + n)
            self.rows += padded
            return
        self.reserve(key, b.shape[0]).copy_(b.to(self.dtype), non_blocking=True)

    # ---- drawing steps ------------------------------------------------------------------------------
    def __contains__(self, key: Hashable) -> bool:
        return key in self.index

    def __len__(self) -> int:
        return len(self.index)

    def sizes(self, keys: Iterable[Hashable]) -> list[int]:
        return [self.index[k][1] - self.index[k][0] for k in keys]

    def plan(self, keys: Sequence[Hashable]) -> "ops.BagPlan":
        """Row-range plan of the step made of the bags ``keys`` (any order, repeats allowed).  Use it with ``self.X``:
        ``net.forward_packed(cohort.X, cohort.plan(keys))``."""
        spans = np.array([self.index[k] for k in keys], dtype=np.int64).reshape(-1, 2)
        return ops.make_plan_ranges(spans[:, 0], spans[:, 1], self.X.shape[0], self.device)

    @property
    def nbytes(self) -> int:
        return self.rows * self.X.shape[1] * self.X.element_size()
