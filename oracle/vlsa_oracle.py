"""TEST INFRASTRUCTURE — CPU oracle for the VLSA language-guided aggregation path.

This file is a *restatement* (not a copy) of the reference arithmetic for the hot
path named in BASELINE.json, written against the same ATen op sequence the
reference calls so that fp32 rounding behaviour is the reference's.  Every
function cites the reference file:line it follows (paths relative to
liupei101/VLSA @ 915f37a).

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the reference's
own, unmodified ``model.deepmil.VLFAN`` / ``logit_pooling`` / ``model.vlsa.VLSA.forward``
/ ``loss.loss_surv.SurvIFMLE`` / ``loss.loss_surv_ext.SurvEMD`` (stub-import harness,
SURVEY.md §8c) and stores their outputs in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this oracle against every one of them.
The reference itself ships no tests for this path; its only known-answer vector
(notebook cell 12) needs the gated CONCH weights and cannot be reproduced offline.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``vlsa_b200/``) must never import it.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

__all__ = [
    "coattn_scale", "task_res_query", "vlfan_forward", "vlsa_forward", "logit_pooling",
    "vlsa_forward_zero_shot", "softmax_converter", "surv_ifmle", "convert_survival_label",
    "cdf_loss_p2_raw", "surv_emd", "objective_loss", "decoupled_similarity", "forward_with_grads",
    "query_pooling", "vlfan_forward_variant",
]


def coattn_scale() -> float:
    """model/deepmil.py:122,125 — ``(torch.ones([]) * np.log(100)).exp()`` as an fp32 scalar."""
    return float((torch.ones([]) * np.log(100)).exp())


def task_res_query(prompt_features: torch.Tensor, residual_features: torch.Tensor, res_ratio: float = 0.5):
    """model/prompt_learners/prompt_adapter.py:125-126 (TaskRes arm)."""
    return res_ratio * residual_features + prompt_features


def vlfan_forward(X, Q, W, b, ret_with_attn: bool = False, scale: torch.Tensor | None = None):
    """model/deepmil.py:170-215 with use_feat_proj=False, gated_query=False, query_pooling='mean',
    pred_head='default'.  X [1,N,D]; Q [P,D]; W [D,D]; b [D]  ->  f [1,D] (+ A [1,P,N])."""
    assert X.shape[0] == 1                                   # deepmil.py:175
    if scale is None:
        scale = (torch.ones([]) * np.log(100)).exp()         # deepmil.py:122,197 (CPU 0-dim fp32)
    scale = scale.to(X.dtype)
    Qn = F.normalize(Q.unsqueeze(0), dim=-1)                 # deepmil.py:187
    norm_X = F.normalize(X, dim=-1)                          # deepmil.py:189
    A_ = torch.matmul(Qn, norm_X.transpose(1, 2))            # deepmil.py:190   [1,P,N]
    A_ = scale * A_                                          # deepmil.py:197
    A = F.softmax(A_, dim=-1)                                # deepmil.py:198
    out = torch.matmul(A, X)                                 # deepmil.py:200   [1,P,D]
    pooled = torch.mean(out, dim=1)                          # deepmil.py:136
    f = F.linear(pooled, W, b)                               # deepmil.py:117,204
    if ret_with_attn:
        return f, A.detach()                                 # deepmil.py:206-213
    return f


def query_pooling(out, method: str, pool_params=None):
    """model/deepmil.py:133-150 over the P co-attention outputs: out [1,P,D] -> (pooled [1,D], scores | None).
    ``pool_params``: 'weight' -> {"weight": [1,P]}; 'attention' (model/layers.py:126-155, returns the RAW logits) ->
    {"attention.0.weight", "attention.0.bias", "attention.2.weight", "attention.2.bias"}; 'gated_attention'
    (model/layers.py:85-123, eval mode: dropout off; returns the softmaxed scores) -> {"fc1.0.weight", "fc1.0.bias",
    "score.0.weight", "score.0.bias", "fc2.weight", "fc2.bias"}."""
    if method == "mean":
        return torch.mean(out, dim=1), None                  # deepmil.py:135-137
    if method == "max":
        return torch.max(out, dim=1)[0], None                # deepmil.py:139-141
    if method == "weight":
        weight = F.softmax(pool_params["weight"], dim=-1).unsqueeze(0)     # deepmil.py:148  [1,1,P]
        return torch.matmul(weight, out).squeeze(1), None    # deepmil.py:149
    if method == "attention":
        h = torch.tanh(F.linear(out, pool_params["attention.0.weight"], pool_params["attention.0.bias"]))
        A_ = F.linear(h, pool_params["attention.2.weight"], pool_params["attention.2.bias"])   # layers.py:144
        A_ = torch.transpose(A_, 2, 1)                       # layers.py:145  [1,1,P]
        attn = F.softmax(A_, dim=2)                          # layers.py:146
        return torch.matmul(attn, out).squeeze(1), A_.squeeze(1)          # layers.py:147-150
    if method == "gated_attention":
        emb = torch.tanh(F.linear(out, pool_params["fc1.0.weight"], pool_params["fc1.0.bias"]))        # layers.py:111
        scr = torch.sigmoid(F.linear(out, pool_params["score.0.weight"], pool_params["score.0.bias"]))  # layers.py:112
        A_ = F.linear(emb.mul(scr), pool_params["fc2.weight"], pool_params["fc2.bias"])               # layers.py:113-114
        A_ = torch.transpose(A_, 2, 1)                       # layers.py:115
        A = F.softmax(A_, dim=2)                             # layers.py:116
        return torch.matmul(A, out).squeeze(1), A.squeeze(1)              # layers.py:117-123
    raise ValueError(method)


def vlfan_forward_variant(X, Q, W, b, gated_query: bool = False, pooling: str = "mean", pool_params=None,
                          pred_head: str = "default", scale: torch.Tensor | None = None, proj_params=None):
    """model/deepmil.py:170-215 with the config-reachable switches no shipped VLSA config enables:
    gated_query (Q has P+1 rows, the last one is the gate, deepmil.py:192-195), query_pooling (deepmil.py:133-150),
    pred_head 'Identity' (deepmil.py:111-114), use_feat_proj (``proj_params``: the Feat_Projecter state dict,
    model/layers.py:65-82, deepmil.py:176-179).  Returns (f [1,D], A [1,P,N], pooling scores | None, out [1,P,D])."""
    assert X.shape[0] == 1                                   # deepmil.py:175
    if proj_params is not None:
        h = F.linear(X.view(-1, X.shape[2]), proj_params["projecter.0.weight"], proj_params["projecter.0.bias"])
        h = F.layer_norm(h, (h.shape[-1],), proj_params["projecter.1.weight"], proj_params["projecter.1.bias"])
        X = h.view(1, -1, h.shape[-1])                       # layers.py:74-79
    if scale is None:
        scale = (torch.ones([]) * np.log(100)).exp()         # deepmil.py:122
    scale = scale.to(X.dtype)
    Qn = F.normalize(Q.unsqueeze(0), dim=-1)                 # deepmil.py:187
    norm_X = F.normalize(X, dim=-1)                          # deepmil.py:189
    A_ = torch.matmul(Qn, norm_X.transpose(1, 2))            # deepmil.py:190
    if gated_query:
        A_ = A_[:, :-1, :] - A_[:, -1:, :]                   # deepmil.py:195
    A_ = scale * A_                                          # deepmil.py:197
    A = F.softmax(A_, dim=-1)                                # deepmil.py:198
    out = torch.matmul(A, X)                                 # deepmil.py:200
    pooled, ext = query_pooling(out, pooling, pool_params)   # deepmil.py:203
    f = pooled if pred_head == "Identity" else F.linear(pooled, W, b)     # deepmil.py:111-118,204
    return f, A.detach(), (None if ext is None else ext.detach()), out


def vlsa_forward(X, Q, W, b, T, logit_scale):
    """model/vlsa.py:181-198 (VLFAN arm): returns (logits [1,R], g [1,D], Tn [R,D])."""
    Tn = F.normalize(T, dim=-1)                              # vlsa.py:185-186
    f = vlfan_forward(X, Q, W, b)                            # vlsa.py:188
    g = F.normalize(f, dim=-1)                               # vlsa.py:189
    ls = logit_scale.exp()                                   # vlsa.py:191
    logits = ls * g @ Tn.t()                                 # vlsa.py:192
    return logits, g, Tn


def logit_pooling(logits, method: str):
    """model/deepmil.py:16-37."""
    if method[:9] in ("logit_max", "logit_top"):
        topk = 1 if method == "logit_max" else int(method.split("top")[-1])
        maxk = min(topk, logits.size(0))
        values, _ = logits.topk(maxk, 0, True, True)
        pooled = values.mean(dim=0, keepdim=True)
    elif method == "logit_mean":
        pooled = logits.mean(dim=0, keepdim=True)
    else:
        raise NotImplementedError(f"The pooling ({method}) is not implemented.")
    preds = pooled.argmax(dim=1)
    return preds, pooled


def vlsa_forward_zero_shot(X, T, logit_scale, pooling: str):
    """model/vlsa.py:181-198 with mil_encoder = FeatMIL(pooling) (deepmil.py:51-67).  pooling 'mean' | 'max': the patch
    FEATURES are pooled to one vector (deepmil.py:57-60) and the logits stay [1,R]; otherwise (identity): per-patch
    logits [N,R] then logit_pooling.  Returns (preds [1], pooled [1,R], g [1|N,D], Tn)."""
    assert X.shape[0] == 1
    Tn = F.normalize(T, dim=-1)
    if pooling in ("mean", "max"):
        vec = torch.mean(X, dim=1) if pooling == "mean" else torch.max(X, dim=1)[0]
        g = F.normalize(vec, dim=-1)
        logits = logit_scale.exp() * g @ Tn.t()
        return logits.argmax(dim=1), logits, g, Tn
    g = F.normalize(X.squeeze(0), dim=-1)                    # FeatMIL identity arm, vlsa.py:189
    logits = logit_scale.exp() * g @ Tn.t()                  # [N,R]
    if logits.shape[0] > 1:                                  # vlsa.py:195
        preds, pooled = logit_pooling(logits, pooling)
    else:
        preds, pooled = logits.argmax(dim=1), logits
    return preds, pooled, g, Tn


def softmax_converter(x):
    """utils/func.py:44."""
    return F.softmax(x, dim=-1)


def surv_ifmle(incidence_hat, t, e, alpha: float = 0.0, eps: float = 1e-7, reduction: str = "mean"):
    """loss/loss_surv.py:144-169."""
    bsz = len(t)
    t = t.view(bsz, 1).long()
    c = 1 - e.view(bsz, 1).float()
    cif = torch.cumsum(incidence_hat, dim=1)
    unc = -(1 - c) * torch.log(torch.gather(incidence_hat, 1, t).clamp(min=eps))
    cen = -c * torch.log((1 - torch.gather(cif, 1, t)).clamp(min=eps))
    neg_l = cen + unc
    loss = (1.0 - alpha) * neg_l + alpha * unc
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    return loss


def convert_survival_label(t, e, n_bins: int):
    """loss/loss_surv_ext.py:42-55 (vectorised; same values as the reference's per-sample loop)."""
    t, e = t.view(-1, 1), e.view(-1, 1)
    ar = torch.arange(n_bins, device=t.device).view(1, -1)
    vec = (ar == t).to(t.dtype)
    vec = vec + (ar > t).to(t.dtype) * (1 - e)
    return vec


def cdf_loss_p2_raw(pred_dist, target_dist):
    """loss/loss_surv_ext.py:13-40 with p=2, normalize_dist=False, ret_raw=True."""
    return torch.sum(torch.pow(torch.cumsum(pred_dist, -1) - torch.cumsum(target_dist, -1), 2), dim=-1)


def surv_emd(y_hat, t, e, cur_logit_scale, reduction: str = "mean"):
    """loss/loss_surv_ext.py:70-109 (p=2, raw_distance=True)."""
    n_bins = y_hat.shape[-1]
    ls = cur_logit_scale.detach() if isinstance(cur_logit_scale, torch.Tensor) else cur_logit_scale
    t = t.view(-1, 1).long()
    e = e.view(-1, 1).long()
    target = convert_survival_label(t, e, n_bins)
    target_dist = torch.softmax((2 * target - 1) * ls, dim=-1)
    pred = (1 - e) * ((1 - target) * y_hat + target * ls) + e * y_hat
    pred_dist = torch.softmax(pred, dim=-1)
    loss = cdf_loss_p2_raw(pred_dist, target_dist)
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    return loss


def objective_loss(raw_pred, t, e, logit_scale_exp, w_ifmle: float = 1.0, w_emd: float = 1.0):
    """runner/vlsa_handler.py:241-258 with loss_type = SurvIFMLE-SurvEMD."""
    p = softmax_converter(raw_pred)
    return w_ifmle * surv_ifmle(p, t, e) + w_emd * surv_emd(p, t, e, logit_scale_exp)


def decoupled_similarity(X, Q, W, b, T, logit_scale):
    """utils/model_inference.py:81-144 ("Approach 2"): returns (A [P,N], probs [1,R], probs_2 [1,R],
    decoupled [P,R])."""
    f, A = vlfan_forward(X, Q, W, b, ret_with_attn=True)
    Tn = F.normalize(T, dim=-1)
    ls = float(logit_scale.exp())
    L = f.norm(dim=-1)
    probs = F.softmax(ls * (f / L) @ Tn.t(), dim=-1)
    enc = F.linear(X, W, b).squeeze(0) / L
    dec = A.squeeze(0) @ (enc @ Tn.t())
    probs_2 = F.softmax(ls * dec.mean(dim=0, keepdim=True), dim=-1)
    return A.squeeze(0), probs, probs_2, dec


def prototype_shap_imp(decoupled_similarity, logit_scale: float):
    """utils/model_inference.py:21-78, subset by subset as the reference does (P <= ~10 in tests)."""
    sim = torch.as_tensor(decoupled_similarity)
    num_p, num_cls = sim.shape

    def risk(sel):
        prob = F.softmax(logit_scale * sim[sel].mean(dim=0), dim=0)
        return float(torch.sum((num_cls - torch.arange(0, num_cls)) * prob))

    def members(code):
        return [i for i in range(num_p) if (code >> i) & 1]

    n_cases = 2 ** num_p
    V = [1.0] + [risk(members(i)) for i in range(1, n_cases)]
    fac = [math.factorial(i) for i in range(num_p + 1)]
    wgt = [fac[i] * fac[num_p - i - 1] / fac[num_p] for i in range(num_p)]
    out = torch.zeros(num_p)
    for i in range(num_p):
        acc = 0.0
        for j in range(n_cases):
            sel = members(j)
            if i in sel:
                continue
            acc += wgt[len(sel)] * (V[j + 2 ** i] - V[j])
        out[i] = acc
    return out


def forward_with_grads(bags, prompt_features, residual, W, b, T, logit_scale, t, e,
                       res_ratio: float = 0.5, w_ifmle: float = 1.0, w_emd: float = 1.0,
                       dtype=torch.float32):
    """One ``VLSAHandler._update_network`` minus the optimizer (runner/vlsa_handler.py:260-283):
    per-bag forward, cat, objective loss, backward via torch autograd.  Returns a dict of the
    loss, per-bag logits and gradients w.r.t. (residual_features, W, b, T, logit_scale)."""
    cast = lambda z: z.detach().to(dtype).clone()
    residual = cast(residual).requires_grad_(True)
    W = cast(W).requires_grad_(True)
    b = cast(b).requires_grad_(True)
    T = cast(T).requires_grad_(True)
    logit_scale = cast(logit_scale).requires_grad_(True)
    pf = cast(prompt_features)
    preds = []
    for X in bags:
        Q = task_res_query(pf, residual, res_ratio)
        logits, _, _ = vlsa_forward(X.to(dtype).view(1, -1, X.shape[-1]), Q, W, b, T, logit_scale)
        preds.append(logits)
    raw = torch.cat(preds, dim=0)
    loss = objective_loss(raw, t, e, logit_scale.exp(), w_ifmle, w_emd)
    loss.backward()
    return {
        "loss": loss.detach(), "logits": raw.detach(),
        "d_residual": residual.grad, "d_W": W.grad, "d_b": b.grad, "d_T": T.grad,
        "d_logit_scale": logit_scale.grad,
    }
