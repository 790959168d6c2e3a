"""Survival objective of the VLSA path on B200 (mirror of loss/loss_surv.py, loss/loss_surv_ext.py, loss/utils.py).

``SurvIFMLE`` and ``SurvEMD`` keep the reference signatures (they take the *converted* incidence);
``SurvObjective`` is the fused form of ``VLSAHandler.calc_objective_loss`` (runner/vlsa_handler.py:241-258):
softmax + both losses + gradient in one launch on the raw logits.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import ops

__all__ = ["SurvIFMLE", "SurvEMD", "SurvObjective", "load_loss"]


def _log_of(cur_logit_scale, device):
    """The loss modules receive exp(logit_scale); the kernel wants the log-space value on the device."""
    if isinstance(cur_logit_scale, torch.Tensor):
        return cur_logit_scale.detach().float().log().reshape(()).to(device)
    return torch.tensor(math.log(float(cur_logit_scale)), dtype=torch.float32, device=device)


def _reduce(total, per_col, reduction, bsz):
    if reduction == "mean":
        return total
    if reduction == "sum":
        return total * bsz
    return per_col            # 'none': per-sample values (no gradient path; the handler never uses it)


class SurvIFMLE(nn.Module):
    """loss/loss_surv.py:127-169."""

    def __init__(self, alpha=0.0, eps=1e-7, reduction="mean", **kws):
        super().__init__()
        assert reduction in ["sum", "mean", "none"]
        self.alpha, self.eps, self.reduction = alpha, eps, reduction

    def forward(self, incidence_hat, t, e, cur_alpha=None):
        alpha = self.alpha if cur_alpha is None else cur_alpha
        zero = torch.zeros((), device=incidence_hat.device)
        total, _, _, _, per = ops.surv_loss(incidence_hat, t, e, zero, 1.0, 0.0, alpha, self.eps, input_is_prob=True)
        return _reduce(total, per[:, 0], self.reduction, incidence_hat.shape[0])


class SurvEMD(nn.Module):
    """loss/loss_surv_ext.py:58-109 (p=2, raw squared CDF distance — the only variant the VLSA configs use)."""

    def __init__(self, p=2, raw_distance=True, reduction="mean", **kws):
        super().__init__()
        if p != 2 or not raw_distance:
            raise NotImplementedError("only SurvEMD(p=2, raw_distance=True) is on the accelerated path")
        assert reduction in ["mean", "sum", "none"]
        self.p, self.raw_distance, self.reduction = p, raw_distance, reduction

    def forward(self, y_hat, t, e, cur_logit_scale=10.0):
        ls = _log_of(cur_logit_scale, y_hat.device)
        total, _, _, _, per = ops.surv_loss(y_hat, t, e, ls, 0.0, 1.0, input_is_prob=True)
        return _reduce(total, per[:, 1], self.reduction, y_hat.shape[0])


class SurvObjective(nn.Module):
    """softmax -> w_ifmle * SurvIFMLE + w_emd * SurvEMD on raw logits, one kernel (+ gradient)."""

    def __init__(self, w_ifmle=1.0, w_emd=1.0, alpha=0.0, eps=1e-7):
        super().__init__()
        self.w_ifmle, self.w_emd, self.alpha, self.eps = w_ifmle, w_emd, alpha, eps

    def forward(self, raw_pred, t, e, logit_scale_param, norm=None):
        """``logit_scale_param`` is the log-space parameter (``net.logit_scale``).  Returns
        (total, ifmle, emd, incidence)."""
        total, l1, l2, inc, _ = ops.surv_loss(raw_pred, t, e, logit_scale_param, self.w_ifmle, self.w_emd,
                                              self.alpha, self.eps, norm=norm)
        return total, l1, l2, inc


def load_loss(task, *args, **kws):
    """loss/utils.py:12-22 for task 'vlsa' with loss_type in {SurvIFMLE, SurvEMD}."""
    if task not in ("sa", "vlsa"):
        raise NotImplementedError(f"cannot recognize the task {task}.")
    table = {"SurvIFMLE": SurvIFMLE, "SurvEMD": SurvEMD}
    out = {}
    for name in kws["loss_type"]:
        if name not in table:
            raise NotImplementedError(f"loss {name} is not part of the accelerated VLSA path")
        out[name] = table[name](**kws.get(name, {}))
    return out
