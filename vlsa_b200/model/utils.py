"""``load_model`` name kept from model/utils.py:13-36 of the reference (VLSA arm only)."""
from .vlsa import VLSA


def load_model(arch: str, **kws):
    if arch == "VLSA":
        return VLSA(**kws)
    raise NotImplementedError(f"Architecture {arch} is not part of vlsa_b200 (only the VLSA hot path is built).")
