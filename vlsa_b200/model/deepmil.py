"""Vision end of VLSA on B200: ``VLFAN`` (language-guided aggregation), ``FeatMIL`` and ``logit_pooling``.

Same constructor arguments, attributes, method names and state-dict keys as model/deepmil.py:16-215 of
liupei101/VLSA; the arithmetic runs in libvlsa_b200.so (no PyTorch fallback).  Configurations of VLFAN that
no shipped VLSA config enables (feat_proj, gated_query, query_pooling != 'mean', pred_head 'Identity')
raise NotImplementedError instead of silently running something else.
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


class VLFAN(nn.Module):
    def __init__(self, dim_in=1024, dim_hid=256, use_feat_proj=True, drop_rate=0.25, query="Parameter", num_query=10,
                 gated_query=False, query_pooling="mean", pred_head="default", dim_reduction=4, keep_ratio=0.8,
                 **kwargs):
        super().__init__()
        if dim_in != ops.D_FEAT:
            raise NotImplementedError(f"the B200 kernels are built for dim_in={ops.D_FEAT} (CONCH), got {dim_in}")
        if use_feat_proj:
            raise NotImplementedError("use_feat_proj=True (Feat_Projecter) is not on the accelerated path "
                                      "(cfg_vlsa_conch.yaml:49 sets it False)")
        if gated_query:
            raise NotImplementedError("gated_query=True is not on the accelerated path")
        if query_pooling != "mean":
            raise NotImplementedError(f"query_pooling={query_pooling!r}: only 'mean' is on the accelerated path "
                                      "(cfg_vlsa_conch.yaml:58)")
        if pred_head == "Identity":
            raise NotImplementedError("pred_head='Identity' is not on the accelerated path")
        assert query in ["Parameter", "Text"]
        if not (1 <= num_query <= ops.MAX_P):
            raise NotImplementedError(f"num_query must be in 1..{ops.MAX_P}, got {num_query}")
        self._pos_gated_query = -1
        self.feat_proj = None
        self.num_query = num_query
        self.query_type = query
        self.gated_query = gated_query
        if self.query_type != "Parameter":
            self.Q = None                                   # call reset_query later (deepmil.py:94-96)
        else:
            self.Q = nn.Parameter(torch.randn(num_query, dim_in))
        self.query_pooling = query_pooling
        self.pred_head = pred_head
        self.visual_adapter = nn.Linear(dim_in, dim_in)
        self.use_custom_coattn = True
        self.coattn_logit_scale = torch.ones([]) * np.log(100)      # CPU scalar, not a buffer (deepmil.py:122)

    # ---- reference API -------------------------------------------------------------------------
    def get_coattn_logit_scale(self):
        return self.coattn_logit_scale.exp()

    def reset_query(self, query_network):
        assert self.query_type != "Parameter", f"Cannot override Q (query) for query_type ({self.query_type})."
        self.Q = query_network

    def forward_query_pooling(self, X):
        return torch.mean(X, dim=1), None

    def get_query(self):
        assert self.Q is not None, f"You have to call `reset_query` to reset query for query_type ({self.query_type})."
        return self.Q() if callable(self.Q) else self.Q

    def query_div_loss(self, last_div=True, **kws):
        Q = self.get_query()
        norm_Q = F.normalize(Q, dim=-1)
        sim = norm_Q @ norm_Q.T
        sim = sim[~torch.eye(len(Q), dtype=torch.bool, device=sim.device)]
        return sim.abs().mean()

    # ---- fused path ----------------------------------------------------------------------------
    def encode_packed(self, X: torch.Tensor, plan: "ops.BagPlan"):
        """Packed bags [total_rows, D] -> visual features f [B, D] (differentiable w.r.t. Q, W, b)."""
        Q = self.get_query()
        f, ml = ops.encode(X, plan, Q, self.visual_adapter.weight, self.visual_adapter.bias,
                           float(self.get_coattn_logit_scale()))
        return f, ml

    def forward(self, X, ret_with_attn=False):
        """X [1, N, C] -> visual_features [1, C] (and A [1, P, N] detached), deepmil.py:170-215."""
        assert X.shape[0] == 1
        Xp = X[0].contiguous()
        plan = ops.make_plan([Xp.shape[0]], Xp.device)
        f, ml = self.encode_packed(Xp, plan)
        if ret_with_attn:
            A = ops.attention_scores(Xp, self.get_query().detach().contiguous(), ml[0],
                                     float(self.get_coattn_logit_scale()))
            return f, A.unsqueeze(0)
        return f
