"""Golden vectors for the VLFAN variants from the UNMODIFIED reference on CPU.

    python tests/golden/make_golden_variants.py [--ref /root/reference]

Runs the reference's own ``model.deepmil.VLFAN`` (stub-import harness of make_golden.py) with gated_query,
query_pooling in {max, weight, attention, gated_attention} and pred_head 'Identity', in eval mode (dropout off),
in fp32 and fp64, and stores f, the attention head, the pooling scores and the autograd gradients of
sum(f * G) w.r.t. Q and the pooling parameters.  Every parameter of a case is regenerated from its seed by
``tests/golden_util.variant_inputs`` (W, b are the shipped checkpoint's), so the fixtures stay small.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, HERE)

from golden_util import VARIANT_CASES, variant_inputs, variant_name  # noqa: E402
from make_golden import import_reference  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    torch.set_num_threads(8)
    deepmil = import_reference(args.ref)[0]
    names = []
    for case in VARIANT_CASES:
        inp = variant_inputs(case)
        P, hid = case["P"], case["hid"]
        rec = {}
        for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
            enc = deepmil.VLFAN(dim_in=512, dim_hid=hid, use_feat_proj=bool(case.get("feat_proj")), drop_rate=0.25, query="Parameter",
                                num_query=P, gated_query=case["gated"], query_pooling=case["pooling"],
                                pred_head=case["pred_head"])
            enc.eval()
            with torch.no_grad():
                enc.Q.copy_(inp["Q"])
                if case["pred_head"] != "Identity":
                    enc.visual_adapter.weight.copy_(inp["W"])
                    enc.visual_adapter.bias.copy_(inp["b"])
                if case["pooling"] == "weight":
                    enc.query_pooling.copy_(inp["pool"]["weight"])
                elif case["pooling"] in ("attention", "gated_attention"):
                    enc.query_pooling.load_state_dict(inp["pool"])
                if case.get("feat_proj"):
                    enc.feat_proj.load_state_dict(inp["proj"])
            enc.to(dtype)
            enc.coattn_logit_scale = enc.coattn_logit_scale.exp().to(dtype).log()   # keep the fp32 value of exp(log 100)
            fs = []
            loss = 0
            for X, G in zip(inp["bags"], inp["G"]):
                f, attn = enc(X.to(dtype).unsqueeze(0), ret_with_attn=True)
                fs.append(f.detach())
                loss = loss + (f * G.to(dtype)).sum()
            loss.backward()
            A, ext = (attn if isinstance(attn, tuple) else (attn, None))       # of the LAST bag
            rec[f"f_{tag}"] = torch.cat(fs, 0).numpy()
            rec[f"d_Q_{tag}"] = enc.Q.grad.numpy()
            rec[f"attn_head_{tag}"] = A[0, :, :64].numpy()
            if ext is not None:
                rec[f"pool_scores_{tag}"] = ext.numpy()
            if case["pooling"] == "weight":
                rec[f"d_pool_weight_{tag}"] = enc.query_pooling.grad.numpy()
            elif case["pooling"] in ("attention", "gated_attention"):
                last = "attention.2.weight" if case["pooling"] == "attention" else "fc2.weight"
                rec[f"d_pool_last_{tag}"] = dict(enc.query_pooling.named_parameters())[last].grad.numpy()
            if case["pred_head"] != "Identity":
                rec[f"d_b_{tag}"] = enc.visual_adapter.bias.grad.numpy()
            if case.get("feat_proj"):
                pg = dict(enc.feat_proj.named_parameters())
                rec[f"d_proj_w_rows_{tag}"] = pg["projecter.0.weight"].grad[:4].numpy()
                rec[f"d_proj_w_fro_{tag}"] = pg["projecter.0.weight"].grad.double().norm().numpy()
                rec[f"d_proj_ln_w_{tag}"] = pg["projecter.1.weight"].grad.numpy()
        name = variant_name(case)
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            x_sum=np.array([x.double().sum().item() for x in inp["bags"]]), **rec)
        names.append(name)
        print("[make_golden_variants]", name, "f[0,:3] =", rec["f_f32"][0, :3],
              "ref fp32 err f %.2e dQ %.2e" % (np.abs(rec["f_f32"] - rec["f_f64"]).max(),
                                              np.abs(rec["d_Q_f32"] - rec["d_Q_f64"]).max()))
    with open(os.path.join(HERE, "INDEX_variants.txt"), "w") as fh:
        fh.write("\n".join(names) + "\n")


if __name__ == "__main__":
    main()
