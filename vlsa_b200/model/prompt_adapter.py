"""Source of the aggregation queries Q (mirror of model/prompt_learners/prompt_adapter.py of the reference).

Only what the hot path needs: the frozen prototype embeddings come in as a tensor
(``pretrained_prompt_features``; in the reference they are the CONCH encoding of the prototype
sentences, prompt_adapter.py:60-70 — the text tower is out of scope here) and the TaskRes residual is the
trainable parameter that receives dQ from the CUDA backward.  State-dict key: ``residual_features``.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class PromptAdapter(nn.Module):
    def __init__(self, prompt_encoder=None, tokenizer=None, method: str = "default", num_prompts: int = 4,
                 pretrained_prompt_features: torch.Tensor | None = None, res_ratio: float = 0.5,
                 load_negative_prompts: bool = False, **kwargs) -> None:
        super().__init__()
        assert method in ["default", "FC", "Adapter", "TaskRes"]
        if method in ("FC", "Adapter"):
            raise NotImplementedError(f"PromptAdapter method {method!r} is outside the accelerated path "
                                      "(every shipped VLSA config uses TaskRes for the query and 'default' for ranks)")
        if load_negative_prompts:
            raise NotImplementedError("gated_query / negative prompts are not part of the accelerated path")
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
            self.neg_residual_features = None
            self.res_ratio = res_ratio

    def get_raw_prompt_features(self):
        return self.prompt_features.clone()

    def forward(self):
        prompt_features = self.prompt_features.clone()
        if self.method == "TaskRes":
            return self.res_ratio * self.residual_features + prompt_features      # prompt_adapter.py:125-126
        return prompt_features
