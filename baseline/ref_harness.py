"""Import the vendored, unmodified reference (baseline/_ref, see install_ref.py) and assemble its own modules for the
bench's reference arm and the GPU eager-torch comparator.  Nothing here is on the product path."""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "model", "deepmil.py")) and os.path.isfile(os.path.join(REF, "model", "vlsa.py"))


_mods = None


def import_reference():
    """-> (model.deepmil, model.vlsa, model.prompt_learners.prompt_adapter) of the reference, imported with stubs for
    the third-party packages that are not installed; the classes that run are the reference's own."""
    global _mods
    if _mods is not None:
        return _mods
    import transformers  # noqa: F401  (before the stubs shadow anything)
    sys.path.insert(0, REF)

    class _Missing(nn.Module):
        def __init__(self, *a, **k):
            raise RuntimeError("stubbed third-party module")

    class _Permissive(types.ModuleType):
        def __getattr__(self, item):
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
    pkg.__path__ = [os.path.join(REF, "model")]
    sys.modules["model"] = pkg
    deepmil = importlib.import_module("model.deepmil")
    vlsa_mod = importlib.import_module("model.vlsa")
    pa_mod = importlib.import_module("model.prompt_learners.prompt_adapter")
    _mods = (deepmil, vlsa_mod, pa_mod)
    return _mods


def build_reference_vlsa(params: dict, P: int, device="cpu"):
    """The reference's VLSA with its VLFAN encoder and TaskRes prompt adapter, assembled without ``VLSA.__init__``
    (it needs the gated CONCH weights): the reference's own ``pretrained_text_features`` shortcut (model/vlsa.py:58-61,
    160-161) supplies the ordinal prompt embeddings.  ``net(X[1,N,512])`` then runs model/vlsa.py:181-198 verbatim."""
    deepmil, vlsa_mod, pa_mod = import_reference()
    D = params["W"].shape[0]
    enc = deepmil.VLFAN(dim_in=D, dim_hid=256, use_feat_proj=False, drop_rate=0.25, query="Text", num_query=P,
                        gated_query=False, query_pooling="mean", pred_head="default")
    qnet = pa_mod.PromptAdapter(None, method="TaskRes", num_prompts=P,
                                pretrained_prompt_features=params["prompt_features"].clone(), res_ratio=params["res_ratio"])
    with torch.no_grad():
        enc.visual_adapter.weight.copy_(params["W"])
        enc.visual_adapter.bias.copy_(params["b"])
        qnet.residual_features.copy_(params["residual_features"])
    enc.reset_query(qnet)
    net = vlsa_mod.VLSA.__new__(vlsa_mod.VLSA)
    nn.Module.__init__(net)
    net.mil_encoder = enc
    net.logit_scale = nn.Parameter(params["logit_scale"].clone())
    net.image_encoder_cfg = {"name": "VLFAN", "pooling": "logit_top10"}
    net.pmt_learner_name = "CoOp"
    net.register_buffer("pretrained_text_features", params["text_features"].clone(), persistent=False)
    net = net.to(device).eval()
    # PromptAdapter keeps the frozen prototype embeddings in a plain attribute / buffer: make sure they follow
    for mod in net.modules():
        for name, val in list(vars(mod).items()):
            if isinstance(val, torch.Tensor) and not isinstance(val, nn.Parameter) and val.device != torch.device(device):
                setattr(mod, name, val.to(device))
    return net
