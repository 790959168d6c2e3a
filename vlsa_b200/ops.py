"""Torch-facing wrappers of the C-ABI (device memory, streams and autograd glue only).

``aggregate(...)`` is the differentiable fused forward of ``VLSA.forward`` (model/vlsa.py:181-198
+ model/deepmil.py:170-215 of the reference) over a *packed* batch of bags.  All arithmetic runs
in libvlsa_b200.so; there is no PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import functools
import math
import threading
from collections import OrderedDict
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

D_FEAT = 512
MAX_P = 16
MAX_R = 32


@functools.lru_cache(maxsize=1)
def coattn_scale() -> float:
    """exp(fp32(log 100)) as the reference computes it (model/deepmil.py:122,125)."""
    return float((torch.ones([]) * np.log(100)).exp())


_KERNEL_FLAG = {None: 0, "auto": 0, "simt": 0x100, "tc": 0x200}
_agg_variant_flag = 0


def set_agg_variant(variant: str | None) -> None:
    """Cross-check hook of the parity tests: force the streaming kernel of fp32 passes — 'simt' = CUDA cores, 'tc' = the
    tcgen05 kernel (register-staged for fp32 rows, TMA-fed for bf16 rows) — or None for the automatic choice (fp32: 'tc' for
    P > 5; bf16: always 'tc').  The choice travels with every call as a VLSA_KERNEL_* bit of x_dtype (include/vlsa_b200.h); the library
    itself keeps no switch."""
    global _agg_variant_flag
    _agg_variant_flag = _KERNEL_FLAG[variant]


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """Handle of torch's current stream on the current device (the C entry costs ~0.3 us; torch.cuda.current_stream()
    builds a Python Stream object per call: ~20 us, three times per optimizer step)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _check_cuda(t: torch.Tensor, name: str, dtype=torch.float32) -> None:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (vlsa_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


_SM_COUNT: dict[int, int] = {}


def sm_count(device=None) -> int:
    if device is not None and torch.device(device).type != "cuda":
        return 148                      # planning only (host logic / CPU tests); kernels never run there
    idx = torch.cuda.current_device() if device is None else torch.device(device).index or 0
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SM_COUNT[idx]


@dataclass
class BagPlan:
    """Row offsets of a packed batch of bags plus the chunk schedule of the streaming kernels."""
    cu_rows_host: np.ndarray        # int64 [B+1]
    chunk_start_host: np.ndarray    # int32 [B+1]
    chunk_rows: int
    cu_rows: torch.Tensor           # device int64 [B+1]  (ranges: [2B] = (first row, one past the last) per bag)
    chunk_start: torch.Tensor       # device int32 [B+1]
    ranges: bool = False            # the bags lie anywhere inside a larger X (a step drawn from a DeviceCohort)
    x_rows: int = -1                # ranges: rows of the buffer the ranges index into

    @property
    def num_bags(self) -> int:
        return len(self.cu_rows_host) - 1

    @property
    def total_rows(self) -> int:
        """Rows of the X this plan goes with (packed: the sum of the bag sizes; ranges: the whole cohort buffer)."""
        return self.x_rows if self.ranges else int(self.cu_rows_host[-1])

    @property
    def total_chunks(self) -> int:
        return int(self.chunk_start_host[-1])


_PLAN_CACHE: "OrderedDict[tuple, tuple[BagPlan, torch.cuda.Event | None]]" = OrderedDict()
_PLAN_CACHE_MAX = 4096          # an entry is two small host arrays + one 512-byte device block


def make_plan(bag_sizes, device, sms: int | None = None) -> BagPlan:
    """Host-side schedule for a batch of bags with the given row counts (vlsa_agg_plan).

    Plans are immutable and cached by (sizes, device, sms): the per-bag loops of the reference call ``VLSA.forward``
    once per slide and epoch, and at a few thousand rows the schedule (numpy + ctypes + one H2D copy, ~45 us) costs more
    than the kernels.  A hit from another stream than the one that uploaded the plan waits on the upload's event."""
    dev = torch.device(device)
    sizes_t = tuple(int(n) for n in bag_sizes)
    key = (sizes_t, dev.type, dev.index if dev.index is not None else (torch.cuda.current_device() if dev.type == "cuda" else -1), sms)
    hit = _PLAN_CACHE.get(key)
    if hit is not None:
        _PLAN_CACHE.move_to_end(key)
        plan, uploaded = hit
        if uploaded is not None:
            if uploaded.query():
                _PLAN_CACHE[key] = (plan, None)            # the upload has completed: later hits skip the query
            else:
                torch.cuda.current_stream(dev).wait_event(uploaded)
        return plan
    plan = _build_plan(sizes_t, dev, sms)
    uploaded = None
    if dev.type == "cuda":
        uploaded = torch.cuda.Event()
        uploaded.record(torch.cuda.current_stream(dev))
    _PLAN_CACHE[key] = (plan, uploaded)
    if len(_PLAN_CACHE) > _PLAN_CACHE_MAX:
        _PLAN_CACHE.popitem(last=False)
    return plan


class _PinnedRing:
    """Small pinned staging slots for the per-step tables (row ranges, chunk table, labels): a truly asynchronous H2D copy
    without a pinned allocation per step, and without the stream synchronisation a copy from pageable memory implies.  Slots
    are protected in GROUPS: one event is recorded after the last slot of a group has been used, and the first slot of a group
    is reused only after that event (recorded a whole ring cycle earlier) has completed — an event per upload costs more host
    time than the upload itself."""

    GROUP = 16

    def __init__(self, slots: int = 64, words: int = 2048):
        assert slots % self.GROUP == 0 and slots >= 2 * self.GROUP
        self.words = words
        flat = torch.empty(slots * words, dtype=torch.int64).pin_memory()
        self.buf = [flat[i * words:(i + 1) * words] for i in range(slots)]
        self.np = [b.numpy() for b in self.buf]
        self.ev: list = [None] * (slots // self.GROUP)
        self.at = 0

    def upload(self, fill, n_words: int, device) -> torch.Tensor:
        """`fill(np_int64_view[n_words])` writes the payload; returns it as an int64 device tensor (current stream)."""
        if n_words > self.words:
            stage = torch.empty(n_words, dtype=torch.int64).pin_memory()
            fill(stage.numpy())
            return stage.to(device, non_blocking=True)
        i = self.at
        self.at = (i + 1) % len(self.buf)
        g, first, last = i // self.GROUP, i % self.GROUP == 0, i % self.GROUP == self.GROUP - 1
        if first and self.ev[g] is not None:
            self.ev[g].synchronize()                  # every copy that read this group's slots in the previous cycle is done
        fill(self.np[i][:n_words])
        out = torch.empty(n_words, dtype=torch.int64, device=device)
        out.copy_(self.buf[i][:n_words], non_blocking=True)
        if last:
            if self.ev[g] is None:
                self.ev[g] = torch.cuda.Event()
            self.ev[g].record()
        return out


_RINGS: dict[tuple, _PinnedRing] = {}        # one ring per (device, stream)
_RING_LOCK = threading.Lock()


def upload_small(fill, n_words: int, device) -> torch.Tensor:
    """n_words int64 words written by `fill` into pinned staging -> device tensor, asynchronously on the current stream of
    `device`.  Serialised by a lock: loader threads and the training thread share the rings."""
    dev = torch.device(device)
    cur = torch.cuda.current_device()
    idx = dev.index if dev.index is not None else cur
    if idx != cur:
        with torch.cuda.device(idx):                      # events are recorded on the current stream of the ring's device
            return upload_small(fill, n_words, dev)
    key = (idx, _stream())                                # one ring per (device, stream): a group's event covers all its copies
    with _RING_LOCK:
        ring = _RINGS.get(key)
        if ring is None:
            if len(_RINGS) >= 32:
                _RINGS.clear()                            # streams come and go; a dropped ring's copies have long completed
            ring = _RINGS[key] = _PinnedRing()
        return ring.upload(fill, n_words, dev)


def _chunk_schedule(sizes: np.ndarray, sms: int):
    """(cu_rows [B+1] int64, chunk_start [B+1] int32, chunk_rows) of the bags with the given row counts (vlsa_agg_plan)."""
    if (sizes < 0).any():
        raise ValueError("negative bag size")
    cu = np.zeros(len(sizes) + 1, dtype=np.int64)
    np.cumsum(sizes, out=cu[1:])
    cs = np.zeros(len(sizes) + 1, dtype=np.int32)
    chunk_rows = C.c_int(0)
    rc = _lib.lib().vlsa_agg_plan(cu.ctypes.data_as(C.POINTER(C.c_int64)), len(sizes), int(sms), C.byref(chunk_rows),
                                  cs.ctypes.data_as(C.POINTER(C.c_int32)))
    _lib.check(rc, "vlsa_agg_plan")
    cu.flags.writeable = False          # plans are shared through the cache: nobody may edit one in place
    cs.flags.writeable = False
    return cu, cs, int(chunk_rows.value)


def _build_plan(bag_sizes, device, sms: int | None) -> BagPlan:
    sizes = np.asarray(list(bag_sizes), dtype=np.int64)
    if sms is None:
        sms = sm_count(device)
    cu, cs, chunk_rows_v = _chunk_schedule(sizes, sms)
    chunk_rows = C.c_int(chunk_rows_v)
    # one small staging array -> one H2D copy on the current stream.  Batched steps keep a pinned staging buffer (a
    # truly asynchronous copy next to the big X copy of the loader); the per-bag calls of the reference's loops use
    # pageable memory, which the runtime stages itself and which is cheaper than a pinned allocation per plan.
    def fill(st):
        st[: len(cu)] = cu
        st[len(cu):].view(np.int32)[: len(cs)] = cs

    if torch.device(device).type == "cuda" and len(sizes) > 4:
        dev = upload_small(fill, len(cu) * 2, device)          # pinned ring: asynchronous, no allocation per plan
    else:
        stage = torch.empty(len(cu) * 2, dtype=torch.int64)
        fill(stage.numpy())
        dev = stage.to(device, non_blocking=True)
    return BagPlan(cu, cs, int(chunk_rows.value), dev[: len(cu)], dev[len(cu):].view(torch.int32)[: len(cs)])


SPLIT16_COLS = 514      # fp32 words per row of a pre-split tile image (vlsa_split16_row_bytes() / 4, include/vlsa_b200.h)


def _is_split16(x: torch.Tensor) -> bool:
    return x.dtype == torch.float32 and x.dim() == 2 and x.shape[1] == SPLIT16_COLS


def _x_dtype_code(x: torch.Tensor) -> int:
    if _is_split16(x):
        return 2
    if x.dtype == torch.float32:
        return 0
    if x.dtype == torch.bfloat16:
        return 1
    raise ValueError(f"X must be float32 or bfloat16, got {x.dtype}")


def _agg_dtype_code(x: torch.Tensor, plan: "BagPlan | None" = None) -> int:
    """x_dtype of the vlsa_agg_* calls: storage type + the kernel the cross-check hook asks for (if any) + the
    row-ranges bit of a cohort plan."""
    return _x_dtype_code(x) | _agg_variant_flag | (0x1000 if plan is not None and plan.ranges else 0)


def make_plan_ranges(begins, ends, x_rows: int, device, sms: int | None = None) -> BagPlan:
    """Plan for bags that lie anywhere inside a larger buffer X [x_rows, 512] (VLSA_ROWS_RANGES): bag b is rows
    begins[b] .. ends[b] - 1.  One small H2D copy (the 2 B ranges + the chunk table); no gather of the rows."""
    begins = np.asarray(begins if isinstance(begins, np.ndarray) else list(begins), dtype=np.int64)
    ends = np.asarray(ends if isinstance(ends, np.ndarray) else list(ends), dtype=np.int64)
    if begins.shape != ends.shape or (ends < begins).any() or (begins < 0).any() or (len(ends) and ends.max() > x_rows):
        raise ValueError("bad row ranges")
    cu, cs, chunk_rows = _chunk_schedule(ends - begins, sm_count(device) if sms is None else sms)
    B = len(begins)
    n_words = 2 * B + (B + 2) // 2 + 1

    def fill(st):
        st[0:2 * B:2] = begins
        st[1:2 * B:2] = ends
        st[2 * B:].view(np.int32)[: B + 1] = cs

    if torch.device(device).type == "cuda":
        dev = upload_small(fill, n_words, device)
    else:
        stage = torch.empty(n_words, dtype=torch.int64)
        fill(stage.numpy())
        dev = stage
    return BagPlan(cu, cs, chunk_rows, dev[: 2 * B], dev[2 * B:].view(torch.int32)[: B + 1], ranges=True, x_rows=int(x_rows))


def split16_pack(X: torch.Tensor, image: torch.Tensor, first_row: int) -> None:
    """vlsa_split16_pack: rows of X [n, 512] fp32 (device) -> the records of `image` [rows_padded, 514] that start at padded
    row `first_row` (a multiple of 16).  Current stream."""
    _check_cuda(X, "X")
    if X.dim() != 2 or X.shape[1] != D_FEAT or not _is_split16(image) or not image.is_cuda or not image.is_contiguous():
        raise ValueError("split16_pack: X [n, 512] fp32 and image [rows, 514] fp32, both on the device")
    n = X.shape[0]
    if first_row % 16 or first_row + (n + 15) // 16 * 16 > image.shape[0]:
        raise ValueError("split16_pack: first_row must be a multiple of 16 and the records must fit the image")
    _lib.check(_lib.lib().vlsa_split16_pack(X.data_ptr(), n, image.data_ptr(), first_row, _stream()), "vlsa_split16_pack")


def _workspace(plan: BagPlan, P: int, device) -> torch.Tensor:
    nbytes = _lib.lib().vlsa_agg_workspace_bytes(plan.total_chunks, plan.num_bags, P)
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def aggregate_forward_raw(X, plan: BagPlan, Q, W, bias, T, logit_scale, need_bwd: bool = True, scale: float | None = None,
                          workspace: torch.Tensor | None = None, want_if: bool = True, q_prenorm: bool = False):
    """Launch vlsa_agg_fwd.  Returns a dict of fresh output tensors (no autograd)."""
    L = _lib.lib()
    B, P, R = plan.num_bags, Q.shape[0], T.shape[0]
    if X.dim() != 2 or (X.shape[1] != D_FEAT and not _is_split16(X)):
        raise ValueError(f"packed X must be [total_rows, {D_FEAT}], got {tuple(X.shape)}")
    if _is_split16(X) and not plan.ranges:
        raise ValueError("a pre-split cohort image goes with a row-range plan (DeviceCohort.plan)")
    if X.shape[0] != plan.total_rows:
        raise ValueError(f"packed X has {X.shape[0]} rows but the plan covers {plan.total_rows}")
    if not (1 <= P <= MAX_P):
        raise ValueError(f"num_query P={P} outside 1..{MAX_P}")
    if not (1 <= R <= MAX_R):
        raise ValueError(f"num_ranks R={R} outside 1..{MAX_R}")
    _check_cuda(X, "X", None)
    for name, t in (("Q", Q), ("W", W), ("bias", bias), ("T", T), ("logit_scale", logit_scale)):
        _check_cuda(t, name)
    if Q.shape[1] != D_FEAT or T.shape[1] != D_FEAT or tuple(W.shape) != (D_FEAT, D_FEAT) or bias.numel() != D_FEAT:
        raise ValueError("parameter shapes do not match D=512")
    dev = X.device
    f32 = dict(dtype=torch.float32, device=dev)
    out = {
        "v": torch.empty(B, D_FEAT, **f32), "f": torch.empty(B, D_FEAT, **f32), "g": torch.empty(B, D_FEAT, **f32),
        "logits": torch.empty(B, R, **f32), "incidence": torch.empty(B, R, **f32) if want_if else None,
        "ml": torch.empty(B, P, 2, **f32), "O": torch.empty(B, P, D_FEAT, **f32) if need_bwd else None,
        "Tn": torch.empty(R, D_FEAT, **f32),
    }
    ws = workspace if workspace is not None else _workspace(plan, P, dev)
    rc = L.vlsa_agg_fwd(X.data_ptr(), _agg_dtype_code(X, plan), plan.total_rows, plan.cu_rows.data_ptr(), plan.chunk_start.data_ptr(), B,
                        plan.chunk_rows, plan.total_chunks, Q.data_ptr(), P, int(bool(q_prenorm)),
                        coattn_scale() if scale is None else float(scale), W.data_ptr(), bias.data_ptr(),
                        T.data_ptr(), R, logit_scale.data_ptr(), ws.data_ptr(), ws.numel(),
                        out["v"].data_ptr(), out["f"].data_ptr(), out["g"].data_ptr(), out["logits"].data_ptr(),
                        _ptr(out["incidence"]), out["ml"].data_ptr(), _ptr(out["O"]), out["Tn"].data_ptr(), _stream())
    _lib.check(rc, "vlsa_agg_fwd")
    out["_workspace"] = ws
    return out


_WS_BYTES: dict[tuple, int] = {}


def aggregate_infer(X, plan: BagPlan, Q, W, bias, T, logit_scale, scale: float | None = None, q_prenorm: bool = False,
                    want_if: bool = True):
    """The inference call of ``VLSA.forward`` / ``forward_packed`` with as little host work as the contract allows: the same
    launches as ``aggregate_forward_raw(need_bwd=False)``, but the workspace and the by-products nobody reads (v, f, the
    softmax statistics) share ONE allocation and the parameters are taken as they are (no detach / contiguous copies).
    At one bag of a few thousand rows per call the GPU work is ~25 us and every allocator call ~3 us.
    Returns (logits [B,R], g [B,512], Tn [R,512], incidence [B,R] or None)."""
    L = _lib.lib()
    B, P, R = plan.num_bags, Q.shape[0], T.shape[0]
    if X.dim() != 2 or (X.shape[1] != D_FEAT and not _is_split16(X)) or X.shape[0] != plan.total_rows:
        raise ValueError(f"packed X {tuple(X.shape)} does not go with a plan over {plan.total_rows} rows of {D_FEAT}")
    if _is_split16(X) and not plan.ranges:
        raise ValueError("a pre-split cohort image goes with a row-range plan (DeviceCohort.plan)")
    if not (1 <= P <= MAX_P and 1 <= R <= MAX_R):
        raise ValueError(f"P={P} / R={R} outside 1..{MAX_P} / 1..{MAX_R}")
    _check_cuda(X, "X", None)
    for name, z in (("Q", Q), ("W", W), ("bias", bias), ("T", T), ("logit_scale", logit_scale)):
        _check_cuda(z, name)
    if Q.shape[1] != D_FEAT or T.shape[1] != D_FEAT or W.shape[0] != D_FEAT or W.shape[1] != D_FEAT or bias.numel() != D_FEAT:
        raise ValueError("parameter shapes do not match D=512")
    key = (plan.total_chunks, B, P)
    ws_bytes = _WS_BYTES.get(key)
    if ws_bytes is None:
        ws_bytes = (max(int(L.vlsa_agg_workspace_bytes(plan.total_chunks, B, P)), 256) + 255) // 256 * 256
        if len(_WS_BYTES) > 4096:
            _WS_BYTES.clear()
        _WS_BYTES[key] = ws_bytes
    dev = X.device
    side = B * (2 * D_FEAT + 2 * P)                       # v | f | ml, fp32
    scratch = torch.empty(ws_bytes + 4 * side, dtype=torch.uint8, device=dev)
    logits = torch.empty(B, R, dtype=torch.float32, device=dev)
    g = torch.empty(B, D_FEAT, dtype=torch.float32, device=dev)
    Tn = torch.empty(R, D_FEAT, dtype=torch.float32, device=dev)
    inc = torch.empty(B, R, dtype=torch.float32, device=dev) if want_if else None
    base = scratch.data_ptr()
    v_ptr = base + ws_bytes
    f_ptr = v_ptr + 4 * B * D_FEAT
    ml_ptr = f_ptr + 4 * B * D_FEAT
    rc = L.vlsa_agg_fwd(X.data_ptr(), _agg_dtype_code(X, plan), plan.total_rows, plan.cu_rows.data_ptr(), plan.chunk_start.data_ptr(),
                        B, plan.chunk_rows, plan.total_chunks, Q.data_ptr(), P, int(bool(q_prenorm)),
                        coattn_scale() if scale is None else float(scale), W.data_ptr(), bias.data_ptr(), T.data_ptr(), R,
                        logit_scale.data_ptr(), base, ws_bytes, v_ptr, f_ptr, g.data_ptr(), logits.data_ptr(),
                        None if inc is None else inc.data_ptr(), ml_ptr, None, Tn.data_ptr(), _stream())
    _lib.check(rc, "vlsa_agg_fwd")
    return logits, g, Tn, inc


def aggregate_partial_only(X, plan: BagPlan, Q, workspace: torch.Tensor, scale: float | None = None) -> None:
    """Launch only the streaming kernel (vlsa_agg_partial_fwd); used to time the dominant kernel alone."""
    rc = _lib.lib().vlsa_agg_partial_fwd(X.data_ptr(), _agg_dtype_code(X, plan), plan.total_rows, plan.cu_rows.data_ptr(),
                                         plan.chunk_start.data_ptr(), plan.num_bags, plan.chunk_rows,
                                         plan.total_chunks, Q.data_ptr(), Q.shape[0],
                                         coattn_scale() if scale is None else float(scale), workspace.data_ptr(),
                                         workspace.numel(), _stream())
    _lib.check(rc, "vlsa_agg_partial_fwd")


class _AggregateFn(torch.autograd.Function):
    """Differentiable w.r.t. (Q, W, bias, T, logit_scale).  X is data (the reference never asks for dX)."""

    @staticmethod
    def forward(ctx, X, plan, Q, W, bias, T, logit_scale, scale, q_prenorm=False):
        Qc, Wc, bc, Tc, lsc = (t.detach().contiguous() for t in (Q, W, bias, T, logit_scale))
        need_bwd = any(t.requires_grad for t in (Q, W, bias, T, logit_scale))
        out = aggregate_forward_raw(X, plan, Qc, Wc, bc, Tc, lsc, need_bwd=need_bwd, scale=scale, q_prenorm=q_prenorm)
        ctx.plan, ctx.scale, ctx.need_bwd, ctx.prenorm = plan, scale, need_bwd, int(bool(q_prenorm))
        ctx.set_materialize_grads(False)
        ctx.ws = out["_workspace"]
        if need_bwd:
            ctx.save_for_backward(X, Qc, Wc, Tc, lsc, out["v"], out["f"], out["g"], out["logits"], out["ml"], out["O"])
        ctx.mark_non_differentiable(out["incidence"], out["ml"])
        return out["logits"], out["g"], out["Tn"], out["incidence"], out["ml"]

    @staticmethod
    def backward(ctx, d_logits, d_g, d_Tn, _d_if, _d_ml):
        if d_Tn is not None:
            raise NotImplementedError("gradient through the returned normalised text features is not supported; "
                                      "differentiate through the logits")
        X, Q, W, T, ls, v, f, g, logits, ml, O = ctx.saved_tensors
        plan = ctx.plan
        L = _lib.lib()
        B, P, R = plan.num_bags, Q.shape[0], T.shape[0]
        dev = X.device
        f32 = dict(dtype=torch.float32, device=dev)
        d_logits = (torch.zeros(B, R, **f32) if d_logits is None else d_logits.contiguous().float())
        d_g = None if d_g is None else d_g.contiguous().float()
        dQ, dW, db = torch.empty(P, D_FEAT, **f32), torch.empty(D_FEAT, D_FEAT, **f32), torch.empty(D_FEAT, **f32)
        dT, dls = torch.empty(R, D_FEAT, **f32), torch.empty((), **f32)
        rc = L.vlsa_agg_bwd(X.data_ptr(), _agg_dtype_code(X, plan), plan.total_rows, plan.cu_rows.data_ptr(), plan.chunk_start.data_ptr(), B,
                            plan.chunk_rows, plan.total_chunks, Q.data_ptr(), P, ctx.prenorm,
                            coattn_scale() if ctx.scale is None else float(ctx.scale), W.data_ptr(), T.data_ptr(), R,
                            ls.data_ptr(), v.data_ptr(), f.data_ptr(), g.data_ptr(), logits.data_ptr(), ml.data_ptr(),
                            O.data_ptr(), d_logits.data_ptr(), _ptr(d_g), None, ctx.ws.data_ptr(), ctx.ws.numel(),
                            dQ.data_ptr(), dW.data_ptr(), db.data_ptr(), dT.data_ptr(), dls.data_ptr(), _stream())
        _lib.check(rc, "vlsa_agg_bwd")
        return None, None, dQ, dW, db, dT, dls, None, None


class _EncodeFn(torch.autograd.Function):
    """VLFAN.forward alone (deepmil.py:170-215): packed bags -> f [B, D]; differentiable w.r.t. (Q, W, bias) and,
    when the rows come out of a trainable feat_proj (X.requires_grad), w.r.t. X."""

    @staticmethod
    def forward(ctx, X, plan, Q, W, bias, scale, q_prenorm=False):
        L = _lib.lib()
        Qc, Wc, bc = (t.detach().contiguous() for t in (Q, W, bias))
        B, P = plan.num_bags, Qc.shape[0]
        if X.dim() != 2 or (X.shape[1] != D_FEAT and not _is_split16(X)) or X.shape[0] != plan.total_rows:
            raise ValueError(f"packed X must be [{plan.total_rows}, {D_FEAT}], got {tuple(X.shape)}")
        if not (1 <= P <= MAX_P):
            raise ValueError(f"num_query P={P} outside 1..{MAX_P}")
        _check_cuda(X, "X", None)
        for name, t in (("Q", Qc), ("W", Wc), ("bias", bc)):
            _check_cuda(t, name)
        f32 = dict(dtype=torch.float32, device=X.device)
        v, f = torch.empty(B, D_FEAT, **f32), torch.empty(B, D_FEAT, **f32)
        ml, O = torch.empty(B, P, 2, **f32), torch.empty(B, P, D_FEAT, **f32)
        ws = _workspace(plan, P, X.device)
        sc = coattn_scale() if scale is None else float(scale)
        rc = L.vlsa_agg_fwd(X.data_ptr(), _agg_dtype_code(X, plan), plan.total_rows, plan.cu_rows.data_ptr(), plan.chunk_start.data_ptr(), B,
                            plan.chunk_rows, plan.total_chunks, Qc.data_ptr(), P, int(bool(q_prenorm)), sc, Wc.data_ptr(),
                            bc.data_ptr(), None, 0, None, ws.data_ptr(), ws.numel(), v.data_ptr(), f.data_ptr(), None,
                            None, None, ml.data_ptr(), O.data_ptr(), None, _stream())
        _lib.check(rc, "vlsa_agg_fwd")
        ctx.plan, ctx.scale, ctx.ws, ctx.prenorm = plan, sc, ws, int(bool(q_prenorm))
        if ctx.needs_input_grad[0] and X.dtype != torch.float32:
            raise ValueError("a gradient w.r.t. the patch rows needs fp32 rows")
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(X.detach(), Qc, Wc, v, ml, O)
        ctx.mark_non_differentiable(ml)
        return f, ml

    @staticmethod
    def backward(ctx, d_f, _d_ml):
        if d_f is None:
            return (None,) * 7
        X, Q, W, v, ml, O = ctx.saved_tensors
        plan = ctx.plan
        P = Q.shape[0]
        f32 = dict(dtype=torch.float32, device=X.device)
        d_f = d_f.contiguous().float()
        dQ, dW, db = torch.empty(P, D_FEAT, **f32), torch.empty(D_FEAT, D_FEAT, **f32), torch.empty(D_FEAT, **f32)
        rc = _lib.lib().vlsa_agg_bwd(X.data_ptr(), _agg_dtype_code(X, plan), plan.total_rows, plan.cu_rows.data_ptr(),
                                     plan.chunk_start.data_ptr(), plan.num_bags, plan.chunk_rows, plan.total_chunks,
                                     Q.data_ptr(), P, ctx.prenorm, ctx.scale, W.data_ptr(), None, 0, None, v.data_ptr(),
                                     None, None, None, ml.data_ptr(), O.data_ptr(), None, None, d_f.data_ptr(),
                                     ctx.ws.data_ptr(), ctx.ws.numel(), dQ.data_ptr(), dW.data_ptr(), db.data_ptr(), None,
                                     None, _stream())
        _lib.check(rc, "vlsa_agg_bwd")
        dX = None
        if ctx.needs_input_grad[0]:
            # mean over P: every prototype sees the same gradient row dv / P, dv = W^T d_f ([B,512] x [512,512])
            dO = ((d_f @ W) / P).unsqueeze(1).expand(-1, P, -1).contiguous()
            dX = _pooled_dx(X, plan, Q, ctx.prenorm, ctx.scale, ml, O, dO)
        return dX, None, dQ, dW, db, None, None


class _PooledFn(torch.autograd.Function):
    """Per-prototype pooled features O [B, P, D] (deepmil.py:187-200) for the VLFAN variants whose tail is not
    "mean over P -> Linear".  Differentiable w.r.t. the query directions — and w.r.t. X when X itself requires a
    gradient (rows produced by a trainable feat_proj); the gradient d_O may be anything."""

    @staticmethod
    def forward(ctx, X, plan, Q, q_prenorm, scale):
        Qc = Q.detach().contiguous()
        B, P = plan.num_bags, Qc.shape[0]
        if X.dim() != 2 or (X.shape[1] != D_FEAT and not _is_split16(X)) or X.shape[0] != plan.total_rows:
            raise ValueError(f"packed X must be [{plan.total_rows}, {D_FEAT}], got {tuple(X.shape)}")
        if not (1 <= P <= MAX_P):
            raise ValueError(f"num_query P={P} outside 1..{MAX_P}")
        _check_cuda(X, "X", None)
        _check_cuda(Qc, "Q")
        if Qc.shape[1] != D_FEAT:
            raise ValueError("Q must be [P, 512]")
        f32 = dict(dtype=torch.float32, device=X.device)
        ml, O = torch.empty(B, P, 2, **f32), torch.empty(B, P, D_FEAT, **f32)
        ws = _workspace(plan, P, X.device)
        sc = coattn_scale() if scale is None else float(scale)
        rc = _lib.lib().vlsa_agg_pooled_fwd(X.data_ptr(), _agg_dtype_code(X, plan), plan.total_rows, plan.cu_rows.data_ptr(),
                                            plan.chunk_start.data_ptr(), B, plan.chunk_rows, plan.total_chunks,
                                            Qc.data_ptr(), P, int(bool(q_prenorm)), sc, ws.data_ptr(), ws.numel(),
                                            ml.data_ptr(), O.data_ptr(), _stream())
        _lib.check(rc, "vlsa_agg_pooled_fwd")
        ctx.plan, ctx.scale, ctx.ws, ctx.prenorm = plan, sc, ws, int(bool(q_prenorm))
        if ctx.needs_input_grad[0] and X.dtype != torch.float32:
            raise ValueError("a gradient w.r.t. the patch rows needs fp32 rows")
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(X.detach(), Qc, ml, O)
        ctx.mark_non_differentiable(ml)
        return O, ml

    @staticmethod
    def backward(ctx, d_O, _d_ml):
        if d_O is None:
            return (None,) * 5
        X, Q, ml, O = ctx.saved_tensors
        plan = ctx.plan
        P = Q.shape[0]
        d_O = d_O.contiguous().float()
        dQ = torch.empty(P, D_FEAT, dtype=torch.float32, device=X.device)
        rc = _lib.lib().vlsa_agg_pooled_bwd(X.data_ptr(), _agg_dtype_code(X, plan), plan.total_rows, plan.cu_rows.data_ptr(),
                                            plan.chunk_start.data_ptr(), plan.num_bags, plan.chunk_rows,
                                            plan.total_chunks, Q.data_ptr(), P, ctx.prenorm, ctx.scale, ml.data_ptr(),
                                            O.data_ptr(), d_O.data_ptr(), ctx.ws.data_ptr(), ctx.ws.numel(),
                                            dQ.data_ptr(), _stream())
        _lib.check(rc, "vlsa_agg_pooled_bwd")
        dX = _pooled_dx(X, plan, Q, ctx.prenorm, ctx.scale, ml, O, d_O) if ctx.needs_input_grad[0] else None
        return dX, None, dQ, None, None


def _pooled_dx(X, plan: BagPlan, Q, prenorm: int, scale: float, ml, O, d_O):
    """vlsa_agg_pooled_bwd_dx: gradient of the pooled aggregation w.r.t. the (fp32) patch rows."""
    if plan.ranges:
        raise NotImplementedError("a gradient w.r.t. the patch rows needs a packed batch (not a cohort row-range plan)")
    dX = torch.empty_like(X)
    sizes = np.diff(plan.cu_rows_host)
    rc = _lib.lib().vlsa_agg_pooled_bwd_dx(X.data_ptr(), plan.cu_rows.data_ptr(), plan.num_bags,
                                           int(sizes.max()) if len(sizes) else 0, Q.data_ptr(), Q.shape[0], int(prenorm),
                                           float(scale), ml.data_ptr(), O.data_ptr(), d_O.data_ptr(), dX.data_ptr(),
                                           _stream())
    _lib.check(rc, "vlsa_agg_pooled_bwd_dx")
    return dX


def pooled(X, plan: BagPlan, Q, q_prenorm: bool = False, scale: float | None = None):
    """Packed bags -> (O [B,P,D], ml [B,P,2]): the P softmax-weighted sums of every bag, before any pooling over P.
    ``q_prenorm=True``: the rows of Q are used as score directions without normalisation (gated queries)."""
    return _PooledFn.apply(X, plan, Q, q_prenorm, scale)


def encode(X, plan: BagPlan, Q, W, bias, scale: float | None = None, q_prenorm: bool = False):
    """Fused VLFAN forward on a packed batch: returns (f [B,D], ml [B,P,2])."""
    return _EncodeFn.apply(X, plan, Q, W, bias, scale, q_prenorm)


def _needs_graph(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t.requires_grad for t in tensors)


def aggregate(X, plan: BagPlan, Q, W, bias, T, logit_scale, scale: float | None = None, q_prenorm: bool = False):
    """Fused VLSA forward on a packed batch.  Returns (logits [B,R], g [B,D], Tn [R,D], incidence [B,R], ml).
    ``q_prenorm``: the rows of Q are the gated query's difference rows, used without normalisation."""
    if not _needs_graph(Q, W, bias, T, logit_scale):
        # inference: same launches, no autograd node, nothing kept for a backward
        out = aggregate_forward_raw(X, plan, *(t.detach().contiguous() for t in (Q, W, bias, T, logit_scale)),
                                    need_bwd=False, scale=scale, q_prenorm=q_prenorm)
        return out["logits"], out["g"], out["Tn"], out["incidence"], out["ml"]
    return _AggregateFn.apply(X, plan, Q, W, bias, T, logit_scale, scale, q_prenorm)


def attention_scores(X, Q, ml, scale: float | None = None, q_prenorm: bool = False):
    """A [P,N] for ONE bag.  With ``ml`` (the (max, sum) saved by the forward): softmax over the N patches of
    scale * cos(Q, X), the ``ret_with_attn=True`` output of VLFAN.forward (model/deepmil.py:206-213).  With
    ``ml=None``: softmax over the P prototypes per patch (utils/model_inference.py:104-113, axis_softmax='L')."""
    L = _lib.lib()
    _check_cuda(X, "X", None)
    _check_cuda(Q, "Q")
    if ml is not None:
        _check_cuda(ml, "ml")
    N, P = X.shape[0], Q.shape[0]
    A = torch.empty(P, N, dtype=torch.float32, device=X.device)
    rc = L.vlsa_attn_fwd(X.data_ptr(), _x_dtype_code(X), N, Q.data_ptr(), P, int(bool(q_prenorm)),
                         coattn_scale() if scale is None else float(scale), _ptr(ml), A.data_ptr(), _stream())
    _lib.check(rc, "vlsa_attn_fwd")
    return A


def decoupled_similarity(O, W, bias, T, f, logit_scale):
    """Interpretation path (utils/model_inference.py:115-131) from the pooled per-prototype features O [B,P,512]
    of the forward: returns (sim [B,P,R], decoupled_imp [B,P,R], probs_2 [B,R]).  No pass over X."""
    for name, t in (("O", O), ("W", W), ("bias", bias), ("T", T), ("f", f), ("logit_scale", logit_scale)):
        _check_cuda(t, name)
    B, P, R = O.shape[0], O.shape[1], T.shape[0]
    f32 = dict(dtype=torch.float32, device=O.device)
    sim, imp, probs = torch.empty(B, P, R, **f32), torch.empty(B, P, R, **f32), torch.empty(B, R, **f32)
    rc = _lib.lib().vlsa_interp_fwd(O.data_ptr(), W.data_ptr(), bias.data_ptr(), T.data_ptr(), R, f.data_ptr(),
                                    logit_scale.data_ptr(), B, P, sim.data_ptr(), imp.data_ptr(), probs.data_ptr(),
                                    _stream())
    _lib.check(rc, "vlsa_interp_fwd")
    return sim, imp, probs


class _SurvLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, t, e, logit_scale, w_ifmle, w_emd, alpha, eps, inv_norm, input_is_prob):
        L = _lib.lib()
        B, R = logits.shape
        lg = logits.detach().contiguous().float()
        dev = lg.device
        loss = torch.empty(3, dtype=torch.float32, device=dev)
        inc = torch.empty(B, R, dtype=torch.float32, device=dev)
        dlog = torch.empty(B, R, dtype=torch.float32, device=dev)
        per = torch.empty(B, 2, dtype=torch.float32, device=dev)
        rc = L.vlsa_surv_loss_fwd_bwd(lg.data_ptr(), t.data_ptr(), e.data_ptr(), B, R, logit_scale.data_ptr(),
                                      float(w_ifmle), float(w_emd), float(alpha), float(eps), float(inv_norm),
                                      int(bool(input_is_prob)), loss.data_ptr(), inc.data_ptr(), dlog.data_ptr(), per.data_ptr(), _stream())
        _lib.check(rc, "vlsa_surv_loss_fwd_bwd")
        ctx.save_for_backward(dlog)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(inc, per)
        return loss[0], loss[1], loss[2], inc, per

    @staticmethod
    def backward(ctx, d_total, d_ifmle, d_emd, _a, _b):
        (dlog,) = ctx.saved_tensors
        if d_ifmle is not None or d_emd is not None:
            raise NotImplementedError("differentiate the total loss, not its components")
        if d_total is None:
            return (None,) * 10
        return dlog * d_total, None, None, None, None, None, None, None, None, None


def surv_loss(logits, t, e, logit_scale, w_ifmle: float = 1.0, w_emd: float = 1.0, alpha: float = 0.0,
              eps: float = 1e-7, norm: int | None = None, input_is_prob: bool = False):
    """softmax -> w_ifmle * SurvIFMLE + w_emd * SurvEMD with mean over ``norm`` samples (default: B), fused
    forward + gradient (runner/vlsa_handler.py:241-258, loss/loss_surv.py:144-169, loss/loss_surv_ext.py:70-109).
    ``logit_scale`` is the log-space parameter; SurvEMD uses exp(logit_scale).detach().
    Returns (total, ifmle, emd, incidence [B,R], per_sample [B,2])."""
    _check_cuda(logits, "logits")
    t = t.reshape(-1).to(device=logits.device, dtype=torch.int64).contiguous()
    e = e.reshape(-1).to(device=logits.device, dtype=torch.int64).contiguous()
    B = logits.shape[0]
    if t.numel() != B or e.numel() != B:
        raise ValueError("t and e must have one entry per row of logits")
    ls = logit_scale.detach().reshape(()).float().contiguous()
    inv_norm = 1.0 / float(B if norm is None else norm)
    return _SurvLossFn.apply(logits, t, e, ls, w_ifmle, w_emd, alpha, eps, inv_norm, input_is_prob)


class FusedTrainStep:
    """Forward + fused survival loss + backward of ONE packed optimizer step as three C calls (vlsa_agg_fwd,
    vlsa_surv_loss_fwd_bwd, vlsa_agg_bwd) on buffers that live across steps — the same kernels and numbers as
    ``aggregate`` -> ``surv_loss`` -> ``backward()`` through autograd (runner/vlsa_handler.py:262-283 of the reference),
    without the per-step Python that path costs: two autograd nodes, ~25 allocations and as many small tensor ops.  At
    TCGA bag sizes (3-20k rows) the kernels of a step take 0.2-0.3 ms and that Python 0.5 ms.

    Gradients of the leaf parameters the kernels serve directly (W, bias, logit_scale) are WRITTEN (not accumulated) into the
    tensors given in ``grad_out`` — the handler passes the parameters' views of its all-reduce bucket, zeroed at the top of the
    step; ``dQ`` / ``dT`` come back as buffers for the caller to push through whatever small graph produced Q and T
    (prompt adapter, prompt learner).  Buffers are keyed by (B, P, R, stream): the next step on the same stream reuses them in
    stream order, so everything this returns except ``logits`` is valid until the next call."""

    def __init__(self):
        self._bufs: dict = {}
        self._ws: dict = {}

    def _buffers(self, B, P, R, dev, stream):
        key = (B, P, R, dev.index, stream)
        b = self._bufs.get(key)
        if b is None:
            f32 = dict(dtype=torch.float32, device=dev)
            b = {"v": torch.empty(B, D_FEAT, **f32), "f": torch.empty(B, D_FEAT, **f32), "g": torch.empty(B, D_FEAT, **f32),
                 "inc": torch.empty(B, R, **f32), "ml": torch.empty(B, P, 2, **f32), "O": torch.empty(B, P, D_FEAT, **f32),
                 "Tn": torch.empty(R, D_FEAT, **f32), "dlog": torch.empty(B, R, **f32), "per": torch.empty(B, 2, **f32),
                 "loss": torch.empty(4, **f32), "dQ": torch.empty(P, D_FEAT, **f32), "dW": torch.empty(D_FEAT, D_FEAT, **f32),
                 "db": torch.empty(D_FEAT, **f32), "dT": torch.empty(R, D_FEAT, **f32), "dls": torch.empty(4, **f32)}
            if len(self._bufs) > 64:
                self._bufs.clear()
            self._bufs[key] = b
        return b

    def _workspace(self, plan, P, dev, stream):
        nbytes = max(int(_lib.lib().vlsa_agg_workspace_bytes(plan.total_chunks, plan.num_bags, P)), 256)
        key = (dev.index, stream)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(nbytes + nbytes // 4, dtype=torch.uint8, device=dev)
            self._ws[key] = ws
        return ws

    @staticmethod
    def _target(t, shape):
        """A gradient tensor a kernel may write with 128-bit stores."""
        return t is not None and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == shape \
            and t.data_ptr() % 16 == 0

    def __call__(self, X, plan: BagPlan, Q, W, bias, T, logit_scale, t, e, *, w_ifmle: float = 1.0, w_emd: float = 1.0,
                 alpha: float = 0.0, eps: float = 1e-7, norm: int | None = None, scale: float | None = None,
                 q_prenorm: bool = False, grad_out: dict | None = None, loss_out: torch.Tensor | None = None):
        """-> dict(logits [B,R] (fresh), loss [>=3] = (total, ifmle, emd), incidence, dQ, dT, dW, db, dls).  ``grad_out`` may
        hold targets for "W", "bias", "logit_scale"; ``loss_out`` a float32 device tensor of >= 3 elements for the losses."""
        L = _lib.lib()
        B, P, R = plan.num_bags, Q.shape[0], T.shape[0]
        if X.dim() != 2 or (X.shape[1] != D_FEAT and not _is_split16(X)) or X.shape[0] != plan.total_rows:
            raise ValueError(f"packed X {tuple(X.shape)} does not go with a plan over {plan.total_rows} rows of {D_FEAT}")
        if _is_split16(X) and not plan.ranges:
            raise ValueError("a pre-split cohort image goes with a row-range plan (DeviceCohort.plan)")
        if not (1 <= P <= MAX_P and 1 <= R <= MAX_R):
            raise ValueError(f"P={P} / R={R} outside 1..{MAX_P} / 1..{MAX_R}")
        _check_cuda(X, "X", None)
        Qc, Wc, bc, Tc, lsc = (z.detach().contiguous() for z in (Q, W, bias, T, logit_scale))
        for name, z in (("Q", Qc), ("W", Wc), ("bias", bc), ("T", Tc), ("logit_scale", lsc)):
            _check_cuda(z, name)
        if Qc.shape[1] != D_FEAT or Tc.shape[1] != D_FEAT or tuple(Wc.shape) != (D_FEAT, D_FEAT) or bc.numel() != D_FEAT:
            raise ValueError("parameter shapes do not match D=512")
        if t.dtype != torch.int64 or e.dtype != torch.int64 or t.numel() != B or e.numel() != B or not t.is_cuda or not e.is_cuda:
            raise ValueError("t and e must be int64 device tensors with one entry per bag")
        dev, st = X.device, _stream()
        b = self._buffers(B, P, R, dev, st)
        ws = self._workspace(plan, P, dev, st)
        logits = torch.empty(B, R, dtype=torch.float32, device=dev)
        sc = coattn_scale() if scale is None else float(scale)
        code, pre = _agg_dtype_code(X, plan), int(bool(q_prenorm))
        xp, cu, cs, wsp, wsn = X.data_ptr(), plan.cu_rows.data_ptr(), plan.chunk_start.data_ptr(), ws.data_ptr(), ws.numel()
        rc = L.vlsa_agg_fwd(xp, code, plan.total_rows, cu, cs, B, plan.chunk_rows, plan.total_chunks, Qc.data_ptr(), P, pre, sc,
                            Wc.data_ptr(), bc.data_ptr(), Tc.data_ptr(), R, lsc.data_ptr(), wsp, wsn, b["v"].data_ptr(),
                            b["f"].data_ptr(), b["g"].data_ptr(), logits.data_ptr(), None, b["ml"].data_ptr(), b["O"].data_ptr(),
                            b["Tn"].data_ptr(), st)
        _lib.check(rc, "vlsa_agg_fwd")
        loss = loss_out if (loss_out is not None and loss_out.is_cuda and loss_out.dtype == torch.float32
                            and loss_out.is_contiguous() and loss_out.numel() >= 3) else b["loss"]
        rc = L.vlsa_surv_loss_fwd_bwd(logits.data_ptr(), t.data_ptr(), e.data_ptr(), B, R, lsc.data_ptr(), float(w_ifmle),
                                      float(w_emd), float(alpha), float(eps), 1.0 / float(B if norm is None else norm), 0,
                                      loss.data_ptr(), b["inc"].data_ptr(), b["dlog"].data_ptr(), b["per"].data_ptr(), st)
        _lib.check(rc, "vlsa_surv_loss_fwd_bwd")
        g = grad_out or {}
        dW = g["W"] if self._target(g.get("W"), (D_FEAT, D_FEAT)) else b["dW"]
        db = g["bias"] if self._target(g.get("bias"), (D_FEAT,)) else b["db"]
        dls = g["logit_scale"] if (g.get("logit_scale") is not None and g["logit_scale"].is_cuda
                                   and g["logit_scale"].dtype == torch.float32 and g["logit_scale"].numel() == 1) else b["dls"]
        rc = L.vlsa_agg_bwd(xp, code, plan.total_rows, cu, cs, B, plan.chunk_rows, plan.total_chunks, Qc.data_ptr(), P, pre, sc,
                            Wc.data_ptr(), Tc.data_ptr(), R, lsc.data_ptr(), b["v"].data_ptr(), b["f"].data_ptr(),
                            b["g"].data_ptr(), logits.data_ptr(), b["ml"].data_ptr(), b["O"].data_ptr(), b["dlog"].data_ptr(),
                            None, None, wsp, wsn, b["dQ"].data_ptr(), dW.data_ptr(), db.data_ptr(), b["dT"].data_ptr(),
                            dls.data_ptr(), st)
        _lib.check(rc, "vlsa_agg_bwd")
        return {"logits": logits, "loss": loss, "incidence": b["inc"], "dQ": b["dQ"], "dT": b["dT"], "dW": dW, "db": db, "dls": dls,
                "wrote": {"W": dW is g.get("W"), "bias": db is g.get("bias"), "logit_scale": dls is g.get("logit_scale")}}


def logit_pool(X, T, logit_scale, pooling: str):
    """Zero-shot arm: per-patch logits exp(ls) * cos(x_n, T_r) pooled over N per class
    (model/vlsa.py:189-196 with FeatMIL identity + model/deepmil.py:16-37).  Returns (preds [1] int64, pooled [1,R])."""
    L = _lib.lib()
    _check_cuda(X, "X", None)
    _check_cuda(T, "T")
    if pooling[:9] in ("logit_max", "logit_top"):
        k = 1 if pooling == "logit_max" else int(pooling.split("top")[-1])
        mode = 1
    elif pooling == "logit_mean":
        k, mode = 0, 0
    else:
        raise NotImplementedError(f"The pooling ({pooling}) is not implemented.")
    N, R = X.shape[0], T.shape[0]
    nbytes = L.vlsa_logit_pool_workspace_bytes(N, R, k)
    ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=X.device)
    pooled = torch.empty(1, R, dtype=torch.float32, device=X.device)
    pred = torch.empty(1, dtype=torch.int64, device=X.device)
    ls = logit_scale.detach().reshape(()).float().contiguous()
    rc = L.vlsa_logit_pool_fwd(X.data_ptr(), _x_dtype_code(X), N, T.data_ptr(), R, ls.data_ptr(), mode, k,
                               ws.data_ptr(), ws.numel(), pooled.data_ptr(), pred.data_ptr(), _stream())
    _lib.check(rc, "vlsa_logit_pool_fwd")
    return pred, pooled


def feat_pool(X, T, logit_scale, pooling: str):
    """Zero-shot arm with FeatMIL pooling 'mean' | 'max' (model/deepmil.py:57-60 + model/vlsa.py:188-192), one bag:
    returns (logits [1,R], image_features g [1,512], Tn [R,512]).  A one-row bag under ANY pooling is the 'mean' case."""
    L = _lib.lib()
    _check_cuda(X, "X", None)
    _check_cuda(T, "T")
    mode = {"mean": 0, "max": 2}[pooling]
    N, R = X.shape[0], T.shape[0]
    if N < 1:
        raise ValueError("empty bag")
    f32 = dict(dtype=torch.float32, device=X.device)
    ws = torch.empty(int(L.vlsa_feat_pool_workspace_bytes()), dtype=torch.uint8, device=X.device)
    f, g, logits, Tn = torch.empty(1, D_FEAT, **f32), torch.empty(1, D_FEAT, **f32), torch.empty(1, R, **f32), torch.empty(R, D_FEAT, **f32)
    ls = logit_scale.detach().reshape(()).float().contiguous()
    rc = L.vlsa_feat_pool_fwd(X.data_ptr(), _x_dtype_code(X), N, mode, T.data_ptr(), R, ls.data_ptr(), ws.data_ptr(),
                              ws.numel(), f.data_ptr(), g.data_ptr(), logits.data_ptr(), Tn.data_ptr(), _stream())
    _lib.check(rc, "vlsa_feat_pool_fwd")
    return logits, g, Tn


def row_normalize(X):
    """F.normalize(X, dim=-1) in fp32 for packed rows [N, 512] (image_features of the logit-pooling zero-shot modes)."""
    _check_cuda(X, "X", None)
    out = torch.empty(X.shape[0], D_FEAT, dtype=torch.float32, device=X.device)
    rc = _lib.lib().vlsa_row_normalize(X.data_ptr(), _x_dtype_code(X), X.shape[0], out.data_ptr(), _stream())
    _lib.check(rc, "vlsa_row_normalize")
    return out


def forward_host(X_host: torch.Tensor, bag_sizes, Q, W, bias, T, logit_scale, out_if_host: torch.Tensor | None = None,
                 workspace: torch.Tensor | None = None, copy_stream: torch.cuda.Stream | None = None,
                 scale: float | None = None, device=None, q_prenorm: bool = False):
    """vlsa_forward_host: packed HOST bags (pinned for async copies) -> incidence on the HOST.  Enqueues the
    H2D copy on `copy_stream`, the kernels and the D2H copy on the current stream; returns
    (out_if_host [B,R], workspace) without synchronising.  The returned workspace may be passed to the next call
    straight away: the library orders the copy stream behind the kernels that still read it."""
    L = _lib.lib()
    device = Q.device if device is None else torch.device(device)
    sizes = np.asarray(list(bag_sizes), dtype=np.int64)
    cu = np.zeros(len(sizes) + 1, dtype=np.int64)
    np.cumsum(sizes, out=cu[1:])
    B, P, R = len(sizes), Q.shape[0], T.shape[0]
    if X_host.is_cuda or X_host.dim() != 2 or X_host.shape[1] != D_FEAT or X_host.shape[0] < int(cu[-1]):
        raise ValueError("X_host must be a CPU tensor [>= total_rows, 512]")
    code = _x_dtype_code(X_host)
    need = L.vlsa_forward_host_workspace_bytes(int(cu[-1]), B, P, code)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(int(need), dtype=torch.uint8, device=device)
    if out_if_host is None:
        out_if_host = torch.empty(B, R, dtype=torch.float32).pin_memory()
    cs = copy_stream.cuda_stream if copy_stream is not None else _stream()
    rc = L.vlsa_forward_host(X_host.data_ptr(), code, cu.ctypes.data_as(C.POINTER(C.c_int64)), B, Q.data_ptr(), P,
                             int(bool(q_prenorm)), coattn_scale() if scale is None else float(scale), W.data_ptr(), bias.data_ptr(),
                             T.data_ptr(), R, logit_scale.data_ptr(), workspace.data_ptr(), workspace.numel(),
                             out_if_host.data_ptr(), None, _stream(), cs)
    _lib.check(rc, "vlsa_forward_host")
    return out_if_host, workspace
