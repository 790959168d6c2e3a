"""Round-2 golden vectors, again from the UNMODIFIED reference run on CPU (authoring container only):

  interp_*     the reference's own ``utils.model_inference.calc_text_img_similarity`` and
               ``evaluate_prototype_shap_imp`` (utils/model_inference.py:21-144) on a reference VLSA assembled as in
               make_golden.py (RefBundle).  ``runner.vlsa_handler`` (imported at the top of that file, pulls wandb /
               the evaluators) is stubbed; nothing of it is used by the two functions.
  zeroshot2_*  the remaining branches of the zero-shot arm through the reference's ``VLSA.forward``
               (model/vlsa.py:181-198) with ``FeatMIL``: pooling 'mean' | 'max' (model/deepmil.py:57-60), one-patch
               bags, and the returned ``image_features`` (the N normalised patches) of the logit-pooling modes.

    python tests/golden/make_golden_r02.py [--ref /root/reference]
"""
from __future__ import annotations

import argparse
import importlib
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from vlsa_b200 import synth  # noqa: E402
import make_golden as MG  # noqa: E402

INTERP_CASES = [  # (name, P, R, kind, N, axis_softmax, seed)
    ("interp_real_P12_R12_V", 12, 12, "real", 2798, "V", 41001),
    ("interp_P4_R4_g0_N1000_L", 4, 4, "g0", 1000, "L", 41002),
    ("interp_P7_R13_g1_N5000_V", 7, 13, "g1", 5000, "V", 41003),
    ("interp_P12_R12_g1_N20000_V", 12, 12, "g1", 20000, "V", 41004),
    ("interp_P8_R8_g1_N300_L", 8, 8, "g1", 300, "L", 41005),
]
ZS_CASES = [  # (R, kind, N, pooling)
    (4, "g1", 1000, "mean"), (12, "g0", 2798, "mean"), (4, "g1", 1000, "max"), (16, "g0", 513, "max"),
    (8, "g1", 1, "mean"), (8, "g1", 1, "max"), (8, "g1", 1, "logit_top10"), (4, "g0", 1, "logit_mean"),
    (4, "g1", 1000, "logit_top10"), (12, "g0", 257, "logit_mean"),
]


def interp_inputs(P, R, kind, N, seed, ck):
    """Seeded inputs of an interpretation case (shared with tests/golden_util.py::rebuild_interp)."""
    params = synth.make_params(P, R, seed, w=ck["W"], b=ck["b"])
    if kind == "real":
        X = ck["real"]
        if P == 12:
            params["residual_features"] = ck["residual_features"].clone()
        params["logit_scale"] = ck["logit_scale"].clone()
    else:
        X = synth.make_bag(kind, N, seed + 7)
    return X, params


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    torch.set_num_threads(8)
    mods = MG.import_reference(args.ref)
    deepmil, vlsa_mod = mods[0], mods[1]
    # utils/model_inference.py imports the handler (wandb, evaluators, ...) only for load_vlsa_model
    stub = types.ModuleType("runner.vlsa_handler")
    stub.VLSAHandler = object
    pkg = types.ModuleType("runner")
    pkg.__path__ = []
    sys.modules["runner"] = pkg
    sys.modules["runner.vlsa_handler"] = stub
    MI = importlib.import_module("utils.model_inference")

    z = np.load(os.path.join(HERE, "blca_ckpt_params.npz"))
    ck = {k: torch.from_numpy(z[k].copy()) for k in z.files}
    ck["real"] = torch.from_numpy(np.load(os.path.join(HERE, "blca_bag_A9ST.npy")))
    index = []

    for name, P, R, kind, N, axis, seed in INTERP_CASES:
        X, params = interp_inputs(P, R, kind, N, seed, ck)
        rec = {}
        for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
            rb = MG.RefBundle(mods, params, P, dtype)
            net = rb.net
            _, A, cottn, probs, probs2, imp, shap = MI.calc_text_img_similarity(net, X.to(dtype).unsqueeze(0), axis_softmax=axis)
            k = min(64, A.shape[1])
            rec.update({f"A_head_{tag}": A[:, :k].numpy(), f"A_colsum_head_{tag}": A[:, :k].sum(0).numpy(),
                        f"A_rowsum_{tag}": A.sum(1).numpy(), f"A_max_{tag}": A.max(1).values.numpy(),
                        f"cottn_head_{tag}": cottn[:, :k].numpy(), f"cottn_max_{tag}": cottn.max(1).values.numpy(),
                        f"cottn_argmax_{tag}": cottn.argmax(1).numpy(),
                        f"probs_{tag}": probs.numpy(), f"probs2_{tag}": probs2.numpy(), f"imp_{tag}": imp.numpy(),
                        f"shap_{tag}": shap.numpy()})
        np.savez_compressed(os.path.join(HERE, name + ".npz"), P=P, R=R, kind=kind, n=np.array([X.shape[0]]), seed=seed,
                            axis=axis, x_sum=X.double().sum().numpy(), **rec)
        index.append(name)
        print("[make_golden_r02]", name, "probs[:4] =", rec["probs_f32"][0, :4], "shap[:4] =", rec["shap_f32"][:4])

    for (R, kind, n, pooling) in ZS_CASES:
        seed = synth.BASE_SEED + 30000 + R + n
        X = synth.make_bag(kind, n, seed)
        params = synth.make_params(1, R, seed + 100000, w=ck["W"], b=ck["b"])
        net = vlsa_mod.VLSA.__new__(vlsa_mod.VLSA)
        nn.Module.__init__(net)
        net.mil_encoder = deepmil.FeatMIL(pooling=pooling)
        net.logit_scale = nn.Parameter(params["logit_scale"].clone())
        net.image_encoder_cfg = {"name": "FeatMIL", "pooling": pooling}
        net.pmt_learner_name = "CoOp"
        net.register_buffer("pretrained_text_features", params["text_features"].clone(), persistent=False)
        with torch.no_grad():
            logits, feats, Tn = net(X.unsqueeze(0))
            net64 = net.double()
            logits64, feats64, _ = net64(X.double().unsqueeze(0))
        name = f"zeroshot2_R{R}_{kind}_N{n}_{pooling}"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), R=R, kind=kind, n=np.array([n]), seed=seed, pooling=pooling,
                            x_sum=X.double().sum().numpy(), logits_f32=logits.numpy(), logits_f64=logits64.numpy(),
                            feats_shape=np.array(feats.shape), feats_head_f64=feats64[:8].numpy(),
                            feats_sum_f64=feats64.sum(0).numpy(), Tn_f32=Tn.numpy())
        index.append(name)
        print("[make_golden_r02]", name, "logits =", logits.numpy()[0, :4], "feats", tuple(feats.shape))

    with open(os.path.join(HERE, "INDEX_r02.txt"), "w") as fh:
        fh.write("\n".join(index) + "\n")
    print(f"[make_golden_r02] wrote {len(index)} cases")


if __name__ == "__main__":
    main()
