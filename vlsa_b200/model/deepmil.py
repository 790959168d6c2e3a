"""Vision end of VLSA on B200: ``VLFAN`` (language-guided aggregation), ``FeatMIL`` and ``logit_pooling``.

Same constructor arguments, attributes, method names and state-dict keys as model/deepmil.py:16-215 of
liupei101/VLSA; the arithmetic over the N patches runs in libvlsa_b200.so (no PyTorch fallback).

Shipped configuration (mean over P -> Linear): everything up to the visual feature is one fused CUDA path
(``ops.encode`` / ``ops.aggregate``); ``gated_query`` rides the same path (its P + 1 unit rows collapse to P difference
rows the kernels take as they are).  The other config-reachable variants no shipped VLSA config enables —
``query_pooling`` in {max, weight, attention, gated_attention}, ``pred_head='Identity'`` — share the streaming kernels
through ``ops.pooled`` (O [B,P,512] with a gradient row per prototype on the way back); only their P x 512 tail (the
pooling modules below, state-dict compatible with model/layers.py:85-155) is torch ops on the GPU.
``use_feat_proj=True`` puts the reference's Linear + LayerNorm over all N rows (a plain library GEMM) in front and gets
its gradient from ``vlsa_agg_pooled_bwd_dx``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops

__all__ = ["logit_pooling", "FeatMIL", "VLFAN"]


def logit_pooling(logits: torch.Tensor, method: str):
    """deepmil.py:16-37 on ALREADY materialised [N, C] logits (kept for API parity; the fused zero-shot
    path never materialises them, see ``ops.logit_pool``).  Small tensor ops only."""
    if method[:9] in ["logit_max", "logit_top"]:
        topk = 1 if method == "logit_max" else int(method.split("top")[-1])
        maxk = min(topk, logits.size(0))
        values, _ = logits.topk(maxk, 0, True, True)
        pooled_logits = values.mean(dim=0, keepdim=True)
    elif method == "logit_mean":
        pooled_logits = logits.mean(dim=0, keepdim=True)
    else:
        raise NotImplementedError(f"The pooling ({method}) is not implemented.")
    preds = pooled_logits.argmax(dim=1)
    return preds, pooled_logits


class FeatMIL(nn.Module):
    """Identity encoder of the zero-shot arm (deepmil.py:40-67).  ``VLSA.forward`` recognises it and runs
    the fused per-patch-logit + pooling kernel instead of materialising [N, R] logits."""

    def __init__(self, pooling="mean", **kwargs):
        super().__init__()
        self.network = nn.Identity()
        self.pooling = pooling

    def forward(self, X):
        assert X.shape[0] == 1
        if self.pooling == "mean":
            return torch.mean(X, dim=1)
        if self.pooling == "max":
            return torch.max(X, dim=1)[0]
        return self.network(X.squeeze(0))


class Feat_Projecter(nn.Module):
    """model/layers.py:65-82: Linear + LayerNorm over every patch row ([B, N, C] or [N, C])."""

    def __init__(self, in_dim=1024, out_dim=1024):
        super().__init__()
        self.projecter = nn.Sequential(nn.Linear(in_dim, out_dim), nn.LayerNorm(out_dim))

    def forward(self, x):
        if x.dim() == 3:
            L1, L2, L3 = x.shape
            return self.projecter(x.reshape(-1, L3)).view(L1, L2, -1)
        return self.projecter(x)


class Gated_Attention_Pooling(nn.Module):
    """model/layers.py:85-123 (Ilse et al. 2018) over the P per-prototype features: [B, P, d] -> [B, d]."""

    def __init__(self, in_dim, hid_dim, dropout=0.5):
        super().__init__()
        self.fc1 = nn.Sequential(nn.Linear(in_dim, hid_dim), nn.Tanh(), nn.Dropout(dropout))
        self.score = nn.Sequential(nn.Linear(in_dim, hid_dim), nn.Sigmoid(), nn.Dropout(dropout))
        self.fc2 = nn.Linear(hid_dim, 1)

    def forward(self, x, ret_raw_attn=False):
        if x.dim() == 2:
            x = x.unsqueeze(0)
        A_ = self.fc2(self.fc1(x).mul(self.score(x))).transpose(2, 1)      # [B, 1, P]
        A = F.softmax(A_, dim=2)
        out = torch.matmul(A, x).squeeze(1)
        return (out, A_.squeeze(1)) if ret_raw_attn else (out, A.squeeze(1))


class Attention_Pooling(nn.Module):
    """model/layers.py:126-155: [B, P, d] -> [B, d] and the RAW attention logits [B, P] (ret_raw_attn defaults True)."""

    def __init__(self, in_dim=1024, hid_dim=512):
        super().__init__()
        self.attention = nn.Sequential(nn.Linear(in_dim, hid_dim), nn.Tanh(), nn.Linear(hid_dim, 1))

    def forward(self, x, ret_raw_attn=True):
        if x.dim() == 2:
            x = x.unsqueeze(0)
        A_ = self.attention(x).transpose(2, 1)                              # [B, 1, P]
        attn = F.softmax(A_, dim=2)
        out = torch.matmul(attn, x).squeeze(1)
        return (out, A_.squeeze(1)) if ret_raw_attn else (out, attn.squeeze(1))


class VLFAN(nn.Module):
    def __init__(self, dim_in=1024, dim_hid=256, use_feat_proj=True, drop_rate=0.25, query="Parameter", num_query=10,
                 gated_query=False, query_pooling="mean", pred_head="default", dim_reduction=4, keep_ratio=0.8,
                 **kwargs):
        super().__init__()
        if dim_in != ops.D_FEAT:
            raise NotImplementedError(f"the B200 kernels are built for dim_in={ops.D_FEAT} (CONCH), got {dim_in}")
        assert query in ["Parameter", "Text"]
        assert query_pooling in ["mean", "max", "weight", "attention", "gated_attention"]
        if not (1 <= num_query <= ops.MAX_P):
            raise NotImplementedError(f"num_query must be in 1..{ops.MAX_P}, got {num_query}")
        self._pos_gated_query = -1
        self.feat_proj = Feat_Projecter(dim_in, dim_in) if use_feat_proj else None     # deepmil.py:81-84
        self.num_query = num_query
        self.query_type = query
        self.gated_query = gated_query
        if self.query_type != "Parameter":
            self.Q = None                                   # call reset_query later (deepmil.py:94-96)
        else:
            self.Q = nn.Parameter(torch.randn(num_query + 1 if gated_query else num_query, dim_in))
        if query_pooling == "attention":                    # deepmil.py:102-109
            self.query_pooling = Attention_Pooling(dim_in, dim_hid)
        elif query_pooling == "gated_attention":
            self.query_pooling = Gated_Attention_Pooling(dim_in, dim_hid, dropout=drop_rate)
        elif query_pooling == "weight":
            self.query_pooling = nn.Parameter(torch.randn(1, num_query))
        else:
            self.query_pooling = query_pooling
        self.pred_head = pred_head
        self.visual_adapter = nn.Identity() if pred_head == "Identity" else nn.Linear(dim_in, dim_in)
        self.use_custom_coattn = True
        self.coattn_logit_scale = torch.ones([]) * np.log(100)      # CPU scalar, not a buffer (deepmil.py:122)

    # ---- reference API -------------------------------------------------------------------------
    def get_coattn_logit_scale(self):
        return self.coattn_logit_scale.exp()

    def coattn_scale_float(self) -> float:
        """exp(coattn_logit_scale) as the Python float the kernels take; re-evaluated when the tensor is replaced or
        modified in place."""
        t = self.coattn_logit_scale
        key = (id(t), t._version)
        if getattr(self, "_scale_key", None) != key:
            self._scale_key, self._scale_val = key, float(t.exp())
        return self._scale_val

    def reset_query(self, query_network):
        assert self.query_type != "Parameter", f"Cannot override Q (query) for query_type ({self.query_type})."
        self.Q = query_network

    def forward_query_pooling(self, X):
        """[B, P, C] -> [B, C] (deepmil.py:133-150); P x 512 work."""
        if isinstance(self.query_pooling, str):
            if self.query_pooling == "mean":
                return torch.mean(X, dim=1), None
            return torch.max(X, dim=1)[0], None
        if callable(self.query_pooling):
            return self.query_pooling(X)
        weight = F.softmax(self.query_pooling, dim=-1).unsqueeze(0)         # [1, 1, P]
        return torch.matmul(weight, X).squeeze(1), None

    @property
    def mean_linear_tail(self) -> bool:
        """Mean over P and the Linear adapter: the tail the CUDA path fuses (the shipped configuration)."""
        return (isinstance(self.query_pooling, str) and self.query_pooling == "mean"
                and isinstance(self.visual_adapter, nn.Linear))

    @property
    def fused_tail(self) -> bool:
        """True when VLSA.forward can run as ONE fused call (ops.aggregate): fused tail and raw rows as input."""
        return self.mean_linear_tail and self.feat_proj is None

    def query_directions(self):
        """(rows that enter the scores, prenorm flag).  Gated query (deepmil.py:192-195): A_[:, :-1] - A_[:, -1:] is
        linear in the normalised query, so the P + 1 unit rows collapse to P difference rows used as they are."""
        Q = self.get_query()
        if not self.gated_query:
            return Q, False
        assert self._pos_gated_query == -1, "The gated query is placed at the end by default."
        assert Q.shape[0] == self.num_query + 1, f"Query number is expected to be {self.num_query + 1}."
        Qn = F.normalize(Q, dim=-1)
        return Qn[:-1] - Qn[-1:], True

    def query_directions_cached(self):
        """`query_directions()` for calls that need no gradient, evaluated once per state of the query network: the prompt
        adapter is a few [P, 512] tensor ops (~10 us of launches per call) whose inputs only change at optimizer steps.  The
        cache key is the version counter of every parameter and buffer behind Q (in-place updates bump it) and the
        train / eval mode (dropout in the FC adapter)."""
        Q = self.Q
        if isinstance(Q, nn.Module):
            key = (Q.training, tuple((id(t), t._version) for t in Q.parameters()), tuple((id(t), t._version) for t in Q.buffers()))
            if Q.training and any(isinstance(m, nn.Dropout) and m.p > 0 for m in Q.modules()):
                key = None
        elif isinstance(Q, torch.Tensor):
            key = (id(Q), Q._version)
        else:
            key = None
        hit = getattr(self, "_qdir_cache", None)
        if key is not None and hit is not None and hit[0] == key:
            return hit[1], hit[2]
        with torch.no_grad():
            Qd, prenorm = self.query_directions()
            Qd = Qd.contiguous()
        if key is not None:
            self._qdir_cache = (key, Qd, prenorm)
        return Qd, prenorm

    def get_query(self):
        assert self.Q is not None, f"You have to call `reset_query` to reset query for query_type ({self.query_type})."
        return self.Q() if callable(self.Q) else self.Q

    def query_div_loss(self, last_div=True, **kws):
        Q = self.get_query()
        norm_Q = F.normalize(Q, dim=-1)
        if len(Q) == self.num_query + 1 and last_div:           # deepmil.py:160-162: the gate row against the rest
            sim = norm_Q[-1:] @ norm_Q[:-1].T
        else:
            sim = norm_Q @ norm_Q.T
            sim = sim[~torch.eye(len(Q), dtype=torch.bool, device=sim.device)]
        return sim.abs().mean()

    # ---- fused path ----------------------------------------------------------------------------
    def encode_packed(self, X: torch.Tensor, plan: "ops.BagPlan"):
        """Packed bags [total_rows, D] -> visual features f [B, D] (differentiable w.r.t. Q, W, b)."""
        f, ml, _, _ = self.encode_packed_ext(X, plan)
        return f, ml

    def encode_packed_ext(self, X: torch.Tensor, plan: "ops.BagPlan"):
        """-> (f [B, D], ml, pooling scores or None, the rows the aggregation saw).  Mean + Linear tails run fused
        (ops.encode); the others go streaming kernels -> O [B, P, D] -> pooling over P -> adapter."""
        if self.feat_proj is not None:
            X = self.feat_proj(X.float())                                  # deepmil.py:176-179, [sum N_i, D]
        Qd, prenorm = self.query_directions()
        scale = self.coattn_scale_float()
        if self.mean_linear_tail:
            f, ml = ops.encode(X, plan, Qd, self.visual_adapter.weight, self.visual_adapter.bias, scale, prenorm)
            return f, ml, None, X
        O, ml = ops.pooled(X, plan, Qd, prenorm, scale)
        pooled_out, pooled_ext = self.forward_query_pooling(O)
        return self.visual_adapter(pooled_out), ml, pooled_ext, X

    def forward(self, X, ret_with_attn=False):
        """X [1, N, C] -> visual_features [1, C] (and A [1, P, N] detached), deepmil.py:170-215."""
        assert X.shape[0] == 1
        Xp = X[0].contiguous()
        plan = ops.make_plan([Xp.shape[0]], Xp.device)
        f, ml, pooled_ext, Xs = self.encode_packed_ext(Xp, plan)
        if ret_with_attn:
            Qd, prenorm = self.query_directions()
            A = ops.attention_scores(Xs.detach(), Qd.detach().contiguous(), ml[0], self.coattn_scale_float(),
                                     q_prenorm=prenorm).unsqueeze(0)
            if pooled_ext is not None:
                return f, (A, pooled_ext.detach())                  # deepmil.py:208-209
            return f, A
        return f
