"""``VLSA`` on B200 — drop-in for model/vlsa.py:21-198 of liupei101/VLSA (hot path only).

``forward(X)`` keeps the reference contract: ``X [1, N, 512]`` on the GPU ->
``(logits [1, R], image_features [1, 512], text_features [R, 512])``, differentiable w.r.t.
``logit_scale``, the visual adapter, the query residuals and whatever produced the text features.
``forward_packed`` is the batched entry: one launch for a whole optimizer step of
ragged bags, text features evaluated once.

The language end (CoOp prompt learner + frozen CONCH text tower, model/prompt_encoder.py) is outside
the accelerated path: its output enters as a tensor / callable (``text_features=``), exactly like the
reference's own ``pretrained_text_features`` shortcut (model/vlsa.py:58-61,160-161).
"""
from __future__ import annotations

from typing import Callable

import torch
import torch.nn as nn

from .. import ops
from . import deepmil
from .prompt_adapter import PromptAdapter


class VLSA(nn.Module):
    def __init__(self, text_encoder_cfg=None, image_encoder_cfg=None, prompt_learner_cfg=None,
                 pretrained_prompt_learner_cfg=None, info_prefix="VLSA-B200", *,
                 text_features: torch.Tensor | Callable[[], torch.Tensor] | nn.Module | None = None,
                 query_prompt_features: torch.Tensor | None = None, logit_scale_init: float = 4.6052,
                 **kwargs) -> None:
        super().__init__()
        self.kwargs = kwargs
        image_encoder_cfg = dict(image_encoder_cfg or {})
        self.text_encoder_cfg = text_encoder_cfg
        self.image_encoder_cfg = image_encoder_cfg
        self.prompt_learner_cfg = prompt_learner_cfg
        self.pmt_learner_name = (prompt_learner_cfg or {}).get("name", "CoOp")

        # Vision end: same selection rule as model/utils_vl.py:129-137 (class looked up by cfg['name'])
        enc_name = image_encoder_cfg.get("name", "VLFAN")
        enc_cls = getattr(deepmil, enc_name, None)
        if enc_cls is None:
            raise NotImplementedError(f"image encoder {enc_name!r} is not part of the accelerated path")
        self.mil_encoder = enc_cls(**{k: v for k, v in image_encoder_cfg.items() if k != "name"})

        if enc_name == "VLFAN" and image_encoder_cfg.get("query", "Parameter") == "Text":
            query_text_cfg = {k.split("query_text_")[-1]: v for k, v in image_encoder_cfg.items()
                              if k.startswith("query_text")}                         # model/vlsa.py:82-87
            query_text_cfg.update(num_prompts=image_encoder_cfg["num_query"],
                                  load_negative_prompts=image_encoder_cfg.get("gated_query", False),
                                  pretrained_prompt_features=query_prompt_features,
                                  pretrained_neg_prompt_features=kwargs.get("query_neg_prompt_features"))
            for drop in ("load_path", "load_idx"):
                query_text_cfg.pop(drop, None)               # prototype sentences are encoded outside
            self.mil_encoder.reset_query(PromptAdapter(None, **query_text_cfg))

        # Language end: a tensor (frozen), a Parameter / module / callable (trainable elsewhere)
        self._text_fn = None
        if isinstance(text_features, nn.Module):
            self.prompt_learner = text_features               # keeps reference attribute name for ckpt keys
        elif callable(text_features) and not isinstance(text_features, torch.Tensor):
            self._text_fn = text_features
        elif text_features is not None:
            self.register_buffer("pretrained_text_features", text_features.detach().clone().float(), persistent=False)

        # CLIP-style learnable temperature (model/vlsa.py:102; CONCH initialises it at log(1/0.07)-ish;
        # the shipped BLCA checkpoint holds 4.0309)
        self.logit_scale = nn.Parameter(torch.tensor(float(logit_scale_init)))

    # ---- reference API -------------------------------------------------------------------------
    def forward_text_only(self):
        if hasattr(self, "pretrained_text_features"):
            return self.pretrained_text_features.clone()          # model/vlsa.py:160-161
        if hasattr(self, "prompt_learner"):
            return self.prompt_learner()
        if self._text_fn is not None:
            return self._text_fn()
        raise RuntimeError("no source of ordinal prompt embeddings: pass text_features= to VLSA(...)")

    def _text_features_for_kernels(self):
        """forward_text_only() without the defensive clone of the frozen buffer (the kernels only read it)."""
        if hasattr(self, "pretrained_text_features"):
            return self.pretrained_text_features
        return self.forward_text_only()

    def encode_instances(self, X):
        return self.mil_encoder(X)

    def get_logit_scale(self):
        return self.logit_scale.exp()

    def forward(self, X):
        """X: one bag, [1, N, feat_dim] (model/vlsa.py:181-198)."""
        if X.dim() != 3 or X.shape[0] != 1:
            raise AssertionError("X must be [1, N, feat_dim]")       # deepmil.py:175
        Xp = X[0].contiguous()
        text_features = self._text_features_for_kernels()
        if isinstance(self.mil_encoder, deepmil.FeatMIL):
            # zero-shot arm (model/vlsa.py:188-198 with model/deepmil.py:51-67): inference only, nothing to differentiate
            T = text_features.detach().contiguous()
            pooling = self.mil_encoder.pooling
            if pooling in ("mean", "max"):
                # FeatMIL pools the patch FEATURES: one [1, 512] vector through the cosine head, logits stay [1, R]
                return ops.feat_pool(Xp, T, self.logit_scale, pooling)
            if Xp.shape[0] == 1:
                # identity encoder, one patch: logits.shape[0] == 1, so the reference skips logit_pooling (vlsa.py:195)
                return ops.feat_pool(Xp, T, self.logit_scale, "mean")
            _, pooled = ops.logit_pool(Xp, T, self.logit_scale, self.image_encoder_cfg["pooling"])
            Tn = torch.nn.functional.normalize(text_features, dim=-1)
            # image_features of this arm are the N normalised patches (vlsa.py:188-189); callers that only want the pooled
            # logits can switch the extra [N, 512] write off with `zero_shot_image_features = False`
            feats = ops.row_normalize(Xp) if getattr(self, "zero_shot_image_features", True) else None
            return pooled, feats, Tn
        plan = ops.make_plan([Xp.shape[0]], Xp.device)
        lean = self._infer(Xp, plan, text_features, want_if=False)
        if lean is not None:
            return lean[:3]
        logits, g, Tn, _, _ = self._fused(Xp, plan, text_features)
        return logits, g, Tn

    def graphed(self, n_rows: int, dtype: torch.dtype = torch.float32):
        """CUDA-graph replay of ``forward`` for bags of ``n_rows`` rows (model/graphed.py): one ``cudaGraphLaunch`` per call."""
        from .graphed import GraphedForward
        return GraphedForward(self, n_rows, dtype)

    # ---- batched entry -------------------------------------------------------------
    def _fused(self, Xp, plan, text_features):
        enc = self.mil_encoder
        if not enc.fused_tail:
            # VLFAN variants: streaming kernels -> O [B,P,512]; pooling over P, adapter and the cosine
            # head (model/vlsa.py:186-192) are B x 512 torch ops
            f, ml = enc.encode_packed(Xp, plan)
            Tn = torch.nn.functional.normalize(text_features, dim=-1)
            g = torch.nn.functional.normalize(f, dim=-1)
            logits = self.logit_scale.exp() * g @ Tn.t()
            return logits, g, Tn, torch.softmax(logits.detach(), dim=-1), ml
        Qd, prenorm = enc.query_directions()
        return ops.aggregate(Xp, plan, Qd, enc.visual_adapter.weight, enc.visual_adapter.bias,
                             text_features, self.logit_scale, enc.coattn_scale_float(), prenorm)

    def _infer(self, Xp, plan, text_features, want_if: bool):
        """The fused forward when nothing asks for a gradient (eval loops, `torch.no_grad()`): the lean call
        (``ops.aggregate_infer``) with the query rows cached across calls.  None when autograd has to see the call."""
        enc = self.mil_encoder
        if not enc.fused_tail:
            return None
        W, b, ls = enc.visual_adapter.weight, enc.visual_adapter.bias, self.logit_scale
        grad = torch.is_grad_enabled()
        if grad and (W.requires_grad or b.requires_grad or ls.requires_grad or text_features.requires_grad):
            return None
        Qd, prenorm = enc.query_directions_cached() if not grad else enc.query_directions()
        if grad and Qd.requires_grad:
            return None
        if not (Qd.is_contiguous() and text_features.is_contiguous() and W.is_contiguous()):
            return None
        return ops.aggregate_infer(Xp, plan, Qd, W, b, text_features, ls, enc.coattn_scale_float(), prenorm, want_if)

    def forward_packed(self, X_packed: torch.Tensor, plan: "ops.BagPlan", text_features: torch.Tensor | None = None):
        """All bags of one step in one launch: X_packed [sum N_i, 512] + plan -> (logits [B,R], g [B,512], Tn,
        incidence [B,R]).  Numerically identical to looping ``forward`` over the bags."""
        if text_features is None:
            text_features = self._text_features_for_kernels()
        lean = self._infer(X_packed, plan, text_features, want_if=True)
        if lean is not None:
            return lean
        logits, g, Tn, inc, _ = self._fused(X_packed, plan, text_features)
        return logits, g, Tn, inc
