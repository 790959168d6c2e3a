"""Vendor the reference's own Python sources for the bench's reference arm (authoring container only).

    python baseline/install_ref.py [--ref /root/reference]

The reference (liupei101/VLSA) is a pure-Python research repo without packaging (no setup.py / pyproject.toml), so
``pip install --target baseline/_ref /root/reference`` has nothing to build; what the reference arm needs are the
files of ``model/``, ``loss/`` and ``utils/`` (3 MB, mostly two tokenizer vocabularies read at import), copied UNMODIFIED into the git-ignored ``baseline/_ref/``
(it travels to the GPU box with the snapshot; it never enters the history).  ``baseline/ref_harness.py`` imports them
with the same stub harness ``tests/golden/make_golden.py`` uses (the CONCH / CLIP towers need gated weights and
packages that are not installed; the hot path — ``VLFAN.forward``, ``VLSA.forward`` — runs verbatim).
"""
from __future__ import annotations

import argparse
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def install(ref_root: str = "/root/reference") -> str | None:
    if not os.path.isdir(os.path.join(ref_root, "model")):
        return DEST if os.path.isdir(os.path.join(DEST, "model")) else None
    for sub in ("model", "loss", "utils"):
        for root, _dirs, files in os.walk(os.path.join(ref_root, sub)):
            rel = os.path.relpath(root, ref_root)
            for name in files:
                if name.endswith(".pyc"):
                    continue
                os.makedirs(os.path.join(DEST, rel), exist_ok=True)        # incl. the tokenizer vocabularies model/clip and
                shutil.copyfile(os.path.join(root, name), os.path.join(DEST, rel, name))   # model/conch read at import
    with open(os.path.join(DEST, "SOURCE.txt"), "w") as fh:
        fh.write(f"unmodified files of model/, loss/, utils/ copied from {ref_root} by baseline/install_ref.py\n")
    return DEST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    print(install(ap.parse_args().ref))
