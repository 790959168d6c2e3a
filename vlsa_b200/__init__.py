"""vlsa_b200 — B200-native (sm_100a) implementation of VLSA's language-guided patch-aggregation path.

Host side: Python mirror of the reference's VLSA / VLFAN / losses / handler API.  Arithmetic: hand-written
CUDA in libvlsa_b200.so behind the C ABI declared in include/vlsa_b200.h.  No CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
