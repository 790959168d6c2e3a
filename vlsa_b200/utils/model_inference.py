"""Interpretation helpers of VLSA on the accelerated path (mirror of utils/model_inference.py:21-144).

``calc_text_img_similarity`` keeps the reference's name, arguments and return tuple.  The reference runs
``visual_adapter`` over all N patches (26 GFLOP at N=50k) and re-reads X three times; here ONE streaming pass
produces the pooled per-prototype features O [P,512] and everything else is N-independent:
cottn_score @ ((visual_adapter(X)/|f|) @ Tn^T) = (W O_p + b) . Tn_r / |f| because attention rows sum to one.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import ops


def evaluate_prototype_shap_imp(decoupled_similarity, logit_scale: float, verbose: bool = False) -> torch.Tensor:
    """Exact Shapley value of every prototype for the predicted survival risk (utils/model_inference.py:21-78).

    value(S) = sum_r (R - r) softmax_r(logit_scale * mean_{p in S} sim[p, r]) for a non-empty subset S of the P
    prototypes, value({}) = 1.  All 2^P subset values are evaluated at once from subset sums (host arithmetic on
    a [P,R] matrix; P <= 16)."""
    sim = np.asarray(decoupled_similarity.detach().cpu() if torch.is_tensor(decoupled_similarity) else decoupled_similarity,
                     dtype=np.float32)
    P, R = sim.shape
    n = 1 << P
    masks = np.arange(n)
    member = ((masks[:, None] >> np.arange(P)[None, :]) & 1).astype(np.float32)           # [2^P, P]
    size = member.sum(1)
    mean = (member @ sim) / np.maximum(size, 1.0)[:, None]                                   # [2^P, R]
    z = np.float32(logit_scale) * mean
    z = z - z.max(1, keepdims=True)
    prob = np.exp(z); prob /= prob.sum(1, keepdims=True)
    V = (prob * (R - np.arange(R, dtype=np.float32))[None, :]).sum(1).astype(np.float32)
    V[0] = 1.0
    if verbose:
        print("[SHAP] Survival risk (base) =", V[0])
        print("[SHAP] Survival risk (full) =", V[n - 1])
    fac = [math.factorial(i) for i in range(P + 1)]
    w = np.array([fac[k] * fac[P - k - 1] / fac[P] for k in range(P)], dtype=np.float64)
    shap = np.zeros(P, dtype=np.float64)
    sizes = size.astype(np.int64)
    for i in range(P):
        without = masks[(masks >> i) & 1 == 0]
        shap[i] = np.sum(w[sizes[without]] * (V[without + (1 << i)].astype(np.float64) - V[without].astype(np.float64)))
    out = torch.from_numpy(shap.astype(np.float32))
    if verbose:
        print("[SHAP] Sum over SHAP values =", out.sum())
    return out


def calc_text_img_similarity(model, X_feats, axis_softmax: str = "V", verbose: bool = False):
    """utils/model_inference.py:81-144.  Returns (None, A, cottn_score, probs, probs_2, decoupled_imp,
    decoupled_shap_imp) as CPU tensors: A [P,N] softmax of the co-attention scores over N ('V') or over P ('L'),
    cottn_score [P,N], probs [1,R] (the model's own prediction), probs_2 [1,R], decoupled_imp [P,R], shap [P]."""
    assert axis_softmax in ["L", "V"]
    model.eval()
    if X_feats.dim() == 3 and X_feats.shape[0] == 1:
        X_feats = X_feats[0]
    dev = next(model.parameters()).device
    X = X_feats.to(dev).contiguous()
    enc = model.mil_encoder
    if not getattr(enc, "fused_tail", False) or enc.gated_query:
        # the decoupling (utils/model_inference.py:115-131) rests on the mean over the prototypes and the Linear adapter:
        # with another pooling, a gate row or a projection in front the per-prototype similarities no longer add up to
        # the model's prediction, and the reference's own routine is undefined for them (P + 1 query rows)
        raise NotImplementedError("calc_text_img_similarity needs the shipped VLFAN configuration "
                                  "(query_pooling='mean', pred_head='default', no gated_query / feat_proj)")
    with torch.no_grad():
        T = model.forward_text_only().detach().contiguous()
        Q = enc.get_query().detach().contiguous()
        W, b = enc.visual_adapter.weight.detach(), enc.visual_adapter.bias.detach()
        ls = model.logit_scale.detach()
        scale = float(enc.get_coattn_logit_scale())
        if verbose:
            print("pred_logit_scale:", float(ls.exp()))
            print("coattn_logit_scale:", scale)
        plan = ops.make_plan([X.shape[0]], dev)
        out = ops.aggregate_forward_raw(X, plan, Q, W, b, T, ls, need_bwd=True, scale=scale)
        cottn = ops.attention_scores(X, Q, out["ml"][0], scale)
        A = cottn if axis_softmax == "V" else ops.attention_scores(X, Q, None, scale)
        sim, imp, probs2 = ops.decoupled_similarity(out["O"], W, b, T, out["f"], ls)
        probs = out["incidence"]
    sim_cpu = sim[0].cpu()
    shap = evaluate_prototype_shap_imp(sim_cpu, float(ls.exp()), verbose=verbose)
    return None, A.cpu(), cottn.cpu(), probs.cpu(), probs2.cpu(), imp[0].cpu(), shap
