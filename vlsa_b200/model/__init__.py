from .deepmil import VLFAN, FeatMIL, logit_pooling
from .graphed import GraphedForward
from .prompt_adapter import PromptAdapter
from .utils import load_model
from .vlsa import VLSA

__all__ = ["VLSA", "VLFAN", "FeatMIL", "logit_pooling", "PromptAdapter", "load_model", "GraphedForward"]
