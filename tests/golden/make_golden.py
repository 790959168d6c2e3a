"""Generate golden vectors by running the UNMODIFIED reference (liupei101/VLSA) on CPU.

Run only where the reference checkout exists (the authoring container):

    python tests/golden/make_golden.py [--ref /root/reference]

Writes ``tests/golden/*.npz`` (small) + ``blca_bag_A9ST.npy`` / ``blca_ckpt_params.npz``
(the reference's shipped data assets, re-serialised).  The GPU box has no
/root/reference; tests read only the committed files.

Stub-import harness: the reference's ``model/__init__.py`` pulls CONCH/CLIP
(timm, ftfy, …, not installed), so we register a bare ``model`` package and stub the two
third-party modules ``model/layers.py`` / ``model/deepmil.py`` import at top level.  The classes
that then run — VLFAN, FeatMIL, logit_pooling, VLSA.forward, PromptAdapter.forward,
SurvIFMLE, SurvEMD — are the reference's own code.
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

from vlsa_b200 import synth  # noqa: E402


def import_reference(ref_root: str):
    import transformers  # noqa: F401  (must be imported before the stubs shadow anything)
    sys.path.insert(0, ref_root)

    class _Missing(nn.Module):
        def __init__(self, *a, **k):
            raise RuntimeError("stubbed third-party module")

    class _Permissive(types.ModuleType):
        def __getattr__(self, item):  # any other name resolves to a class that refuses to construct
            if item.startswith("__"):
                raise AttributeError(item)
            return _Missing

    def stub(name, **attrs):
        m = _Permissive(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    stub("nystrom_attention", Nystromformer=_Missing, NystromAttention=_Missing)
    tg = stub("torch_geometric")
    tg.nn = stub("torch_geometric.nn", GENConv=_Missing, DeepGCNLayer=_Missing)
    stub("h5py")
    stub("ftfy")
    timm = stub("timm")
    timm.models = stub("timm.models")
    timm.models.layers = stub("timm.models.layers", trunc_normal_=None, DropPath=_Missing, to_2tuple=None)
    timm.models.vision_transformer = stub("timm.models.vision_transformer", VisionTransformer=_Missing)

    pkg = types.ModuleType("model")
    pkg.__path__ = [os.path.join(ref_root, "model")]
    sys.modules["model"] = pkg
    deepmil = importlib.import_module("model.deepmil")
    loss_surv = importlib.import_module("loss.loss_surv")
    loss_ext = importlib.import_module("loss.loss_surv_ext")
    try:
        vlsa_mod = importlib.import_module("model.vlsa")
    except Exception as ex:  # pragma: no cover
        print("[make_golden] model.vlsa import failed:", repr(ex))
        vlsa_mod = None
    try:
        pa_mod = importlib.import_module("model.prompt_learners.prompt_adapter")
    except Exception as ex:  # pragma: no cover
        print("[make_golden] prompt_adapter import failed:", repr(ex))
        pa_mod = None
    return deepmil, vlsa_mod, pa_mod, loss_surv, loss_ext


class RefBundle:
    """Reference VLSA assembled without its __init__ (needs gated CONCH weights): the
    reference's own shortcut ``pretrained_text_features`` (model/vlsa.py:58-61,160-161)."""

    def __init__(self, mods, params, P, dtype=torch.float32):
        deepmil, vlsa_mod, pa_mod, _, _ = mods
        D = params["W"].shape[0]
        enc = deepmil.VLFAN(dim_in=D, dim_hid=256, use_feat_proj=False, drop_rate=0.25, query="Text",
                            num_query=P, gated_query=False, query_pooling="mean", pred_head="default")
        with torch.no_grad():
            enc.visual_adapter.weight.copy_(params["W"])
            enc.visual_adapter.bias.copy_(params["b"])
        if pa_mod is not None:
            qnet = pa_mod.PromptAdapter(None, method="TaskRes", num_prompts=P,
                                        pretrained_prompt_features=params["prompt_features"].clone(),
                                        res_ratio=params["res_ratio"])
            with torch.no_grad():
                qnet.residual_features.copy_(params["residual_features"])
        else:
            raise RuntimeError("PromptAdapter import failed")
        enc.reset_query(qnet)
        self.enc = enc
        self.qnet = qnet
        if vlsa_mod is not None:
            net = vlsa_mod.VLSA.__new__(vlsa_mod.VLSA)
            nn.Module.__init__(net)
            net.mil_encoder = enc
            net.logit_scale = nn.Parameter(params["logit_scale"].clone())
            net.image_encoder_cfg = {"name": "VLFAN", "pooling": "logit_top10"}
            net.pmt_learner_name = "CoOp"
            # a Parameter instead of the buffer so d/dT is observable; forward_text_only clones it
            net.pretrained_text_features = nn.Parameter(params["text_features"].clone())
            self.net = net
        else:
            raise RuntimeError("model.vlsa import failed")
        if dtype != torch.float32:
            self.net.to(dtype)
            # the fp64 "truth" run keeps the fp32 value of exp(log 100) the fp32 run multiplies by
            self.enc.coattn_logit_scale = self.enc.coattn_logit_scale.exp().to(dtype).log()


def run_case(mods, params, bags, P, R, t, e, with_attn_rows=64):
    """Forward every bag through the reference, then the reference losses + autograd."""
    _, _, _, loss_surv, loss_ext = mods
    out = {}
    for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        rb = RefBundle(mods, params, P, dtype)
        net = rb.net
        preds, gs, fs = [], [], []
        for X in bags:
            Xc = X.to(dtype).unsqueeze(0)
            logits, g, Tn = net(Xc)
            preds.append(logits)
            gs.append(g.detach())
            with torch.no_grad():
                fs.append(rb.enc(Xc))
        raw = torch.cat(preds, dim=0)
        conv = torch.softmax(raw, dim=-1)                      # utils/func.py:44
        ifmle = loss_surv.SurvIFMLE()
        emd = loss_ext.SurvEMD(p=2)
        tt, ee = t.view(-1, 1).to(dtype), e.view(-1, 1).to(dtype)   # labels arrive as float [B,2]
        l1 = ifmle(conv, tt, ee)
        l2 = emd(conv, tt, ee, net.get_logit_scale())
        loss = 1.0 * l1 + 1.0 * l2
        loss.backward()
        out[f"logits_{tag}"] = raw.detach().numpy()
        out[f"if_{tag}"] = conv.detach().numpy()
        out[f"g_{tag}"] = torch.cat(gs, 0).numpy()
        out[f"f_{tag}"] = torch.cat(fs, 0).numpy()
        out[f"loss_ifmle_{tag}"] = l1.detach().numpy()
        out[f"loss_emd_{tag}"] = l2.detach().numpy()
        out[f"loss_{tag}"] = loss.detach().numpy()
        out[f"d_residual_{tag}"] = rb.qnet.residual_features.grad.numpy()
        dW = rb.enc.visual_adapter.weight.grad
        # d_W is [512,512] (3 MB per case in f32+f64): keep a row/column band + two checksums
        out[f"d_W_rows_{tag}"] = dW[:8].numpy()
        out[f"d_W_cols_{tag}"] = dW[:, :8].numpy()
        out[f"d_W_sum_{tag}"] = dW.double().sum().numpy()
        out[f"d_W_fro_{tag}"] = dW.double().norm().numpy()
        out[f"d_b_{tag}"] = rb.enc.visual_adapter.bias.grad.numpy()
        out[f"d_T_{tag}"] = net.pretrained_text_features.grad.numpy()
        out[f"d_logit_scale_{tag}"] = net.logit_scale.grad.numpy()
        # attention of the first bag (first/last rows only + row sums) for ret_with_attn parity
        with torch.no_grad():
            _, A = rb.enc(bags[0].to(dtype).unsqueeze(0), ret_with_attn=True)
        A = A[0]
        k = min(with_attn_rows, A.shape[1])
        out[f"attn_head_{tag}"] = A[:, :k].numpy()
        out[f"attn_max_{tag}"] = A.max(dim=1).values.numpy()
        out[f"attn_argmax_{tag}"] = A.argmax(dim=1).numpy()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    torch.set_num_threads(8)
    torch.backends.cuda.matmul.allow_tf32 = False
    mods = import_reference(args.ref)
    deepmil = mods[0]

    # ---- shipped assets, re-serialised (data, not source) --------------------------------
    ck = torch.load(os.path.join(args.ref, "assert/blca-train-VLSA/train_model-last.pth"), map_location="cpu")["model"]
    W = ck["mil_encoder.visual_adapter.weight"].float()
    b = ck["mil_encoder.visual_adapter.bias"].float()
    res12 = ck["mil_encoder.Q.residual_features"].float()
    ls = ck["logit_scale"].float()
    np.savez(os.path.join(HERE, "blca_ckpt_params.npz"), W=W.numpy(), b=b.numpy(),
             residual_features=res12.numpy(), logit_scale=ls.numpy())
    real = torch.load(os.path.join(args.ref, "assert/blca-test-WSI-TCGA-XF-A9ST.pt"), map_location="cpu").float()
    np.save(os.path.join(HERE, "blca_bag_A9ST.npy"), real.numpy())

    # the shipped experiment config the host side must parse unchanged (a config file, not source)
    import shutil
    shutil.copyfile(os.path.join(args.ref, "config/IFMLE/tcga_blca/cfg_vlsa_conch.yaml"),
                    os.path.join(HERE, "cfg_vlsa_conch.yaml"))
    shutil.copyfile(os.path.join(args.ref, "assert/blca-train-VLSA/config.yaml"),
                    os.path.join(HERE, "blca_train_config.yaml"))

    index = []
    # ---- case family 1: single bags, shapes x generators ----------------------------------
    cases = []
    for (P, R) in ((4, 4), (8, 8), (12, 12), (16, 16), (7, 13), (1, 1)):
        for kind, n in (("g1", 1), ("g1", 7), ("g0", 33), ("g1", 1000), ("g0", 1000), ("g1", 2798), ("g0", 5000)):
            if (P, R) in ((7, 13), (1, 1)) and n not in (7, 1000):
                continue
            cases.append((P, R, kind, n))
    for ci, (P, R, kind, n) in enumerate(cases):
        seed = synth.BASE_SEED + ci
        X = synth.make_bag(kind, n, seed)
        params = synth.make_params(P, R, seed + 100000, w=W, b=b)
        t, e = synth.make_labels(1, R, seed + 200000)
        out = run_case(mods, params, [X], P, R, t, e)
        name = f"single_P{P}_R{R}_{kind}_N{n}"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), P=P, R=R, kind=kind, n=np.array([n]), seed=seed,
                 x_sum=X.double().sum().numpy(), t=t.numpy(), e=e.numpy(), **out)
        index.append(name)
        print("[make_golden]", name, "IF[0,:4] =", out["if_f32"][0, :4])

    # ---- case family 2: the real shipped bag + shipped checkpoint -------------------------
    for (P, R) in ((12, 12), (4, 4)):
        seed = synth.BASE_SEED + 5000 + P
        params = synth.make_params(P, R, seed, w=W, b=b)
        if P == 12:
            params["residual_features"] = res12.clone()
        params["logit_scale"] = ls.clone()
        t, e = synth.make_labels(1, R, seed + 1)
        out = run_case(mods, params, [real], P, R, t, e)
        name = f"real_P{P}_R{R}"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), P=P, R=R, kind="real", n=np.array([real.shape[0]]), seed=seed,
                 x_sum=real.double().sum().numpy(), t=t.numpy(), e=e.numpy(),
                 uses_ckpt_residual=(P == 12), **out)
        index.append(name)
        print("[make_golden]", name, "IF =", out["if_f32"][0])

    # ---- case family 3: one optimizer step's worth of ragged bags (handler path) ----------
    for (P, R, ns, kind) in ((12, 12, (1000, 37, 2798, 1, 513, 4096, 255, 1500), "g1"),
                             (4, 4, (300, 2000, 64, 129), "g0")):
        seed = synth.BASE_SEED + 9000 + P
        bags = [synth.make_bag(kind, n, seed + i) for i, n in enumerate(ns)]
        params = synth.make_params(P, R, seed + 100000, w=W, b=b)
        t, e = synth.make_labels(len(ns), R, seed + 200000)
        out = run_case(mods, params, bags, P, R, t, e)
        name = f"batch_P{P}_R{R}_{kind}_B{len(ns)}"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), P=P, R=R, kind=kind, n=np.array(ns), seed=seed,
                 x_sum=np.array([x.double().sum().item() for x in bags]), t=t.numpy(), e=e.numpy(), **out)
        index.append(name)
        print("[make_golden]", name, "loss =", out["loss_f32"])

    # ---- case family 4: zero-shot arm (FeatMIL + logit_pooling) ---------------------------
    vlsa_mod = mods[1]
    for (R, kind, n, pooling) in ((4, "g1", 1000, "logit_top10"), (4, "g1", 7, "logit_top10"),
                                  (12, "g0", 2798, "logit_top10"), (8, "g1", 513, "logit_mean"),
                                  (16, "g0", 1000, "logit_max"), (4, "g1", 1000, "logit_top3"),
                                  (4, "g1", 1, "logit_top10")):
        seed = synth.BASE_SEED + 20000 + R + n
        X = synth.make_bag(kind, n, seed)
        params = synth.make_params(1, R, seed + 100000, w=W, b=b)
        net = vlsa_mod.VLSA.__new__(vlsa_mod.VLSA)
        nn.Module.__init__(net)
        net.mil_encoder = deepmil.FeatMIL(pooling=pooling)
        net.logit_scale = nn.Parameter(params["logit_scale"].clone())
        net.image_encoder_cfg = {"name": "FeatMIL", "pooling": pooling}
        net.pmt_learner_name = "CoOp"
        net.register_buffer("pretrained_text_features", params["text_features"].clone(), persistent=False)
        with torch.no_grad():
            logits, g, Tn = net(X.unsqueeze(0))
            # what the caller discards (vlsa.py:196) but north_star wants bit-exact: preds
            per_patch = net.logit_scale.exp() * g @ Tn.t()
            if per_patch.shape[0] > 1:
                preds, _ = deepmil.logit_pooling(per_patch, pooling)
            else:
                preds = per_patch.argmax(dim=1)
        name = f"zeroshot_R{R}_{kind}_N{n}_{pooling}"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), R=R, kind=kind, n=np.array([n]), seed=seed, pooling=pooling,
                 x_sum=X.double().sum().numpy(), logits_f32=logits.numpy(), preds=preds.numpy(),
                 per_patch_head=per_patch[:64].numpy())
        index.append(name)
        print("[make_golden]", name, "pooled =", logits.numpy()[0, :4], "pred =", preds.numpy())

    # ---- case family 5: the losses alone on [B,R] (edge labels incl. t=0, t=R-1) -----------
    _, _, _, loss_surv, loss_ext = mods
    for R in (4, 8, 12, 13, 16):
        g = torch.Generator().manual_seed(777 + R)
        Bsz = 32
        raw = 3.0 * torch.randn(Bsz, R, generator=g)
        t, e = synth.make_labels(Bsz, R, 888 + R)
        t[0], t[1], e[0], e[1] = 0, R - 1, 0, 0
        t[2], t[3], e[2], e[3] = 0, R - 1, 1, 1
        rec = {}
        for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
            rawd = raw.detach().to(dtype).clone().requires_grad_(True)
            lsd = torch.tensor(4.0309, dtype=dtype)
            conv = torch.softmax(rawd, dim=-1)
            l1 = loss_surv.SurvIFMLE()(conv, t.view(-1, 1).to(dtype), e.view(-1, 1).to(dtype))
            l2 = loss_ext.SurvEMD(p=2)(conv, t.view(-1, 1).to(dtype), e.view(-1, 1).to(dtype), lsd.exp())
            (l1 + l2).backward()
            rec[f"ifmle_{tag}"] = l1.detach().numpy()
            rec[f"emd_{tag}"] = l2.detach().numpy()
            rec[f"d_raw_{tag}"] = rawd.grad.numpy()
        name = f"loss_R{R}"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), R=R, raw=raw.numpy(), t=t.numpy(), e=e.numpy(),
                 logit_scale=np.float32(4.0309), **rec)
        index.append(name)
        print("[make_golden]", name, rec["ifmle_f32"], rec["emd_f32"])

    with open(os.path.join(HERE, "INDEX.txt"), "w") as fh:
        fh.write("\n".join(index) + "\n")
    print(f"[make_golden] wrote {len(index)} cases")


if __name__ == "__main__":
    main()
