"""Source of the aggregation queries Q (mirror of model/prompt_learners/prompt_adapter.py of the reference).

The frozen prototype embeddings come in as a tensor (``pretrained_prompt_features``; in the reference they are
the CONCH encoding of the prototype sentences, prompt_adapter.py:60-70 — the text tower is out of scope here).  All four
methods of the reference are kept with their state-dict keys — 'TaskRes' (``residual_features``, the shipped choice:
the trainable parameter that receives dQ from the CUDA backward), 'Adapter' (``adapter.fc.{0,2}.weight``,
model/layers.py:50-62), 'FC' (``fc.0.weight``) and 'default'; they are [P, 512] work, N-independent.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class Adapter(nn.Module):
    """model/layers.py:50-62 (CLIP-Adapter bottleneck)."""

    def __init__(self, c_in, reduction=4):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(c_in, c_in // reduction, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(c_in // reduction, c_in, bias=False), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.fc(x)


class PromptAdapter(nn.Module):
    def __init__(self, prompt_encoder=None, tokenizer=None, method: str = "default", num_prompts: int = 4,
                 pretrained_prompt_features: torch.Tensor | None = None, res_ratio: float = 0.5,
                 load_negative_prompts: bool = False, pretrained_neg_prompt_features: torch.Tensor | None = None,
                 dim_reduction: int = 4, keep_ratio: float = 0.8, **kwargs) -> None:
        super().__init__()
        assert method in ["default", "FC", "Adapter", "TaskRes"]
        if load_negative_prompts and pretrained_neg_prompt_features is None:
            raise RuntimeError("gated query: pass `pretrained_neg_prompt_features` ([1, 512], the mean encoding of the "
                               "negative texts, prompt_adapter.py:73-81)")
        if pretrained_prompt_features is None:
            raise RuntimeError("vlsa_b200 does not run the CONCH text tower: pass `pretrained_prompt_features` "
                               "([num_prompts, 512], the encoded prototype sentences)")
        assert len(pretrained_prompt_features) == num_prompts, \
            f"Expected {num_prompts} initial texts, but got {len(pretrained_prompt_features)}."
        self.method = method
        self.register_buffer("prompt_features", pretrained_prompt_features.detach().clone().float(), persistent=False)
        if method == "TaskRes":
            # prompt_adapter.py:94 — randn init, res_ratio 0.5
            self.residual_features = nn.Parameter(torch.randn(num_prompts, self.prompt_features.shape[-1]))
            self.neg_residual_features = nn.Parameter(torch.randn(1, self.prompt_features.shape[-1])) \
                if load_negative_prompts else None                                 # prompt_adapter.py:95-99
            self.res_ratio = res_ratio
        elif method == "Adapter":                                                  # prompt_adapter.py:86-90
            self.adapter = Adapter(self.prompt_features.shape[-1], dim_reduction)
            assert 0 <= keep_ratio <= 1.0
            self.keep_ratio = keep_ratio
        elif method == "FC":                                                       # prompt_adapter.py:101-105
            d = self.prompt_features.shape[-1]
            self.fc = nn.Sequential(nn.Linear(d, d, bias=False), nn.Dropout(0.25))
        if load_negative_prompts:
            neg = pretrained_neg_prompt_features.detach().clone().float().reshape(-1, self.prompt_features.shape[-1])
            self.register_buffer("neg_prompt_features", neg.mean(0, keepdim=True), persistent=False)

    def get_raw_prompt_features(self):
        raw = self.prompt_features.clone()
        if hasattr(self, "neg_prompt_features"):
            raw = torch.cat([raw, self.neg_prompt_features.clone()], dim=0)        # [P + 1, d]
        return raw

    def forward(self):
        # (the reference clones the buffer first; every branch below builds a new tensor anyway, except 'default')
        prompt_features = self.prompt_features
        if self.method == "TaskRes":
            text_features = self.res_ratio * self.residual_features + prompt_features      # prompt_adapter.py:125-126
            if hasattr(self, "neg_prompt_features"):                                       # prompt_adapter.py:127-134
                neg = self.res_ratio * self.neg_residual_features + self.neg_prompt_features.clone()
                text_features = torch.cat([text_features, neg], dim=0)                    # [P + 1, d]
            return text_features
        if self.method == "Adapter":                                                       # prompt_adapter.py:121-123
            return (1 - self.keep_ratio) * self.adapter(prompt_features) + self.keep_ratio * prompt_features
        if self.method == "FC":                                                            # prompt_adapter.py:136-144
            if hasattr(self, "neg_prompt_features"):
                prompt_features = torch.cat([prompt_features, self.neg_prompt_features.clone()], dim=0)
            return self.fc(prompt_features)
        return prompt_features.clone()
