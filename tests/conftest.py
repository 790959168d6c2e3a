import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_cases(prefix):
    names = []
    for index in ("INDEX.txt", "INDEX_r02.txt"):
        with open(os.path.join(GOLDEN, index)) as fh:
            names += [ln.strip() for ln in fh if ln.strip()]
    return [n for n in names if n.startswith(prefix)]


@pytest.fixture(scope="session")
def ckpt_params():
    import numpy as np
    import torch
    z = np.load(os.path.join(GOLDEN, "blca_ckpt_params.npz"))
    return {k: torch.from_numpy(z[k].copy()) for k in z.files}


@pytest.fixture(scope="session")
def real_bag():
    import numpy as np
    import torch
    return torch.from_numpy(np.load(os.path.join(GOLDEN, "blca_bag_A9ST.npy")))
