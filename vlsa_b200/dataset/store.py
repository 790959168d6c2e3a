"""Flat-file patch-feature store and the patient dataset on top of it.

Replaces the reference's one-pickle-per-slide layout (``torch.load`` of ``<sid>.pt`` per slide and per epoch,
``torch.cat`` + ``.float()`` per patient: dataset/PatchWSI.py:197-215, utils/io.py:16-42) with ONE row-major
file of all slides' [N_s, 512] rows (fp32, or bf16 for half the bytes on disk and over PCIe) plus a small JSON
index.  Reading a patient is then a handful of contiguous ``memcpy``s out of the page cache straight into the
pinned staging buffer of ``AsyncBagLoader`` — no unpickling, no intermediate tensors, no dtype pass — and a
32-patient step becomes one packed, pinned tensor that goes to the GPU in one async copy.

Layout on disk (directory):
    features.bin     rows of all slides back to back, 512 x dtype each, slides in index order
    index.json       {"dtype": "float32"|"bfloat16", "dim": 512, "slides": {sid: [row_offset, n_rows]}, "version": 1}
"""
from __future__ import annotations

import json
import os
from concurrent.futures import ThreadPoolExecutor
from typing import Iterable, Iterator, Sequence

import numpy as np
import torch

from .. import ops

_NP_DTYPE = {"float32": np.float32, "bfloat16": np.uint16}          # bf16 is stored as its 16 raw bits
_TORCH_DTYPE = {"float32": torch.float32, "bfloat16": torch.bfloat16}


def read_patch_data(path: str) -> torch.Tensor:
    """utils/io.py:16-42 for the formats that need no extra dependency ('.pt', '.npy')."""
    ext = os.path.splitext(path)[1]
    if ext == ".pt":
        data = torch.load(path, map_location="cpu")
    elif ext == ".npy":
        data = torch.from_numpy(np.load(path))
    else:
        raise ValueError(f"Not support {ext}")
    return data


def build_store(out_dir: str, slides: Iterable[tuple[str, torch.Tensor]], dtype: str = "float32") -> dict:
    """Write ``features.bin`` + ``index.json`` from (slide id, [N, 512] features) pairs.  fp32 features are stored
    bit-exactly (what the reference's ``.to(torch.float)`` yields); ``dtype='bfloat16'`` rounds to nearest even."""
    if dtype not in _NP_DTYPE:
        raise ValueError(f"dtype must be one of {sorted(_NP_DTYPE)}")
    os.makedirs(out_dir, exist_ok=True)
    index: dict[str, list[int]] = {}
    offset = 0
    with open(os.path.join(out_dir, "features.bin"), "wb") as fh:
        for sid, feats in slides:
            if sid in index:
                raise ValueError(f"duplicate slide id {sid}")
            feats = torch.as_tensor(feats)
            if feats.dim() != 2 or feats.shape[1] != ops.D_FEAT:
                raise ValueError(f"slide {sid}: expected [N, {ops.D_FEAT}], got {tuple(feats.shape)}")
            t = feats.to(torch.float).contiguous()                    # PatchWSI.py:212
            if dtype == "bfloat16":
                t = t.to(torch.bfloat16).view(torch.int16)
            fh.write(t.numpy().tobytes())
            index[sid] = [offset, int(feats.shape[0])]
            offset += int(feats.shape[0])
    meta = {"version": 1, "dtype": dtype, "dim": ops.D_FEAT, "rows": offset, "slides": index}
    with open(os.path.join(out_dir, "index.json"), "w") as fh:
        json.dump(meta, fh)
    return meta


def build_store_from_files(out_dir: str, patch_path: str, slide_ids: Sequence[str], read_format: str = "pt",
                           dtype: str = "float32") -> dict:
    """Convert a reference-style directory of ``<sid>.<read_format>`` files."""
    def gen():
        for sid in slide_ids:
            full = os.path.join(patch_path, sid + "." + read_format)
            if not os.path.exists(full):
                print(f"[store] warning: not found slide {sid}.")
                continue
            yield sid, read_patch_data(full)
    return build_store(out_dir, gen(), dtype)


class PatchFeatureStore:
    """Memory-mapped read access to a store written by ``build_store``."""

    def __init__(self, path: str):
        with open(os.path.join(path, "index.json")) as fh:
            self.meta = json.load(fh)
        if self.meta.get("version") != 1 or self.meta.get("dim") != ops.D_FEAT:
            raise ValueError("unsupported store (version / dim)")
        self.dtype_name = self.meta["dtype"]
        self.dtype = _TORCH_DTYPE[self.dtype_name]
        self.slides: dict[str, list[int]] = self.meta["slides"]
        rows = int(self.meta["rows"])
        self._mm = np.memmap(os.path.join(path, "features.bin"), dtype=_NP_DTYPE[self.dtype_name], mode="r",
                             shape=(rows, ops.D_FEAT)) if rows else np.zeros((0, ops.D_FEAT), _NP_DTYPE[self.dtype_name])

    def __contains__(self, sid: str) -> bool:
        return sid in self.slides

    def n_rows(self, sids: Sequence[str]) -> int:
        return sum(self.slides[s][1] for s in sids if s in self.slides)

    def _view(self, t: torch.Tensor) -> np.ndarray:
        return (t.view(torch.int16) if t.dtype == torch.bfloat16 else t).numpy()

    def read_into(self, sids: Sequence[str], out: torch.Tensor) -> int:
        """Copy the rows of the slides ``sids`` (in that order: PatchWSI.py:205-212) into ``out`` [>= n, 512];
        returns the number of rows written.  Missing slides are skipped with the reference's warning."""
        if out.dtype != self.dtype:
            # a bf16 store holds raw 16-bit patterns: copying them into an fp32 tensor would convert the BITS numerically
            raise TypeError(f"read_into: destination is {out.dtype} but the store holds {self.dtype}; read into a "
                            f"{self.dtype} tensor and convert explicitly")
        dst = self._view(out)
        r = 0
        for sid in sids:
            if sid not in self.slides:
                print(f"[WSIPatchSurv] warning: not found slide {sid}.")
                continue
            off, n = self.slides[sid]
            # same element width on both sides; 'unsafe' only reinterprets int16 <-> uint16 bit patterns of bf16
            np.copyto(dst[r:r + n], self._mm[off:off + n], casting="unsafe" if dst.dtype != self._mm.dtype else "same_kind")
            r += n
        return r

    def read(self, sids: Sequence[str]) -> torch.Tensor:
        out = torch.empty(self.n_rows(sids), ops.D_FEAT, dtype=self.dtype)
        self.read_into(sids, out)
        return out


class WSIPatchSurvStore(torch.utils.data.Dataset):
    """``WSIPatchSurv`` in 'patch' mode (dataset/PatchWSI.py:147-215) on a ``PatchFeatureStore``: same item
    ``(index, (feats [N,512] fp32, tensor([0])), label [2])``, plus ``steps()`` that packs whole optimizer steps for
    ``AsyncBagLoader`` without materialising per-patient tensors."""

    def __init__(self, store: PatchFeatureStore, pids: Sequence[str], pid2sids: dict, pid2label: dict):
        self.store, self.pids, self.pid2sids, self.pid2label = store, list(pids), pid2sids, pid2label
        self.uid = self.pids
        print(f"[Dataset] WSIPatchSurv: in patch mode, avaiable patients count {len(self)}.")

    def __len__(self) -> int:
        return len(self.pids)

    def skip_features(self, indices) -> None:
        """Items whose rows already live on the device (a ``DeviceCohort`` keyed by item index, e.g. the one
        ``VLSAHandler._train_each_epoch`` fills with ``vlsa_device_cohort``) come back with an EMPTY feature tensor [0, 512]:
        from the second epoch on the loader no longer reads or copies a single row for them."""
        self._skip = set(int(i) for i in indices)

    def __getitem__(self, index: int):
        pid = self.pids[index]
        if index in getattr(self, "_skip", ()):
            feats = torch.empty(0, ops.D_FEAT, dtype=torch.float)
        else:
            feats = self.store.read(self.pid2sids[pid]).to(torch.float)
        label = torch.Tensor(self.pid2label[pid]).to(torch.float)
        return torch.Tensor([index]).to(torch.int), (feats, torch.Tensor([0])), label

    def steps(self, batch_size: int = 32, order: Sequence[int] | None = None, pin: bool = True, threads: int = 8,
              ring: int = 3) -> Iterator[tuple[torch.Tensor, list[int], torch.Tensor, torch.Tensor]]:
        """Yield ``(X_pinned [rows, 512], sizes, labels [B, 2], index [B])`` per optimizer step
        (``bp_every_batch`` patients, runner/vlsa_handler.py:260-289), the pre-packed form ``AsyncBagLoader`` takes.
        Patients of a step are copied concurrently (memcpy releases the GIL) into one of ``ring`` staging buffers."""
        order = list(range(len(self))) if order is None else list(order)
        bufs: list[torch.Tensor | None] = [None] * ring
        with ThreadPoolExecutor(max_workers=max(1, threads)) as pool:
            for k, s in enumerate(range(0, len(order), batch_size)):
                ids = order[s:s + batch_size]
                sizes = [self.store.n_rows(self.pid2sids[self.pids[i]]) for i in ids]
                total = sum(sizes)
                slot = k % ring
                if bufs[slot] is None or bufs[slot].shape[0] < total:
                    b = torch.empty(max(total, 1), ops.D_FEAT, dtype=self.store.dtype)
                    bufs[slot] = b.pin_memory() if pin and torch.cuda.is_available() else b
                host = bufs[slot]
                starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
                list(pool.map(lambda j: self.store.read_into(self.pid2sids[self.pids[ids[j]]],
                                                            host[starts[j]:starts[j + 1]]), range(len(ids))))
                labels = torch.tensor([list(self.pid2label[self.pids[i]]) for i in ids], dtype=torch.float32)
                yield host[:total], sizes, labels, torch.tensor(ids, dtype=torch.int64)
