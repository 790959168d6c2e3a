"""GPU: the VLFAN variants of SURVEY §8 f4 (gated_query, query_pooling max / weight / attention / gated_attention,
pred_head Identity) through the reference-facing module, against golden vectors produced by the reference's own VLFAN
(tests/golden/make_golden_variants.py) and against the oracle at other sizes."""
import numpy as np
import pytest
import torch

from golden_util import VARIANT_CASES, load_case, variant_inputs, variant_name

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def build_encoder(case, inp, dev):
    from vlsa_b200.model.deepmil import VLFAN
    enc = VLFAN(dim_in=512, dim_hid=case["hid"], use_feat_proj=bool(case.get("feat_proj")), drop_rate=0.25, query="Parameter",
                num_query=case["P"], gated_query=case["gated"], query_pooling=case["pooling"],
                pred_head=case["pred_head"]).eval()
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
    return enc.to(dev)


def close(got, gold, key, slack=4.0, floor=2e-6):
    """|ours - fp64 reference| <= slack * |fp32 reference - fp64 reference| + floor * max|fp64 reference|."""
    r64, r32 = gold[f"{key}_f64"], gold[f"{key}_f32"]
    err = np.abs(np.asarray(got, dtype=np.float64) - r64).max()
    bound = slack * np.abs(r32.astype(np.float64) - r64).max() + floor * max(1.0, np.abs(r64).max())
    assert err <= bound, f"{key}: err {err:.3e} > bound {bound:.3e}"


@pytest.mark.parametrize("case", VARIANT_CASES, ids=variant_name)
def test_variant_matches_reference_golden(case, dev):
    gold = load_case(variant_name(case))
    inp = variant_inputs(case)
    enc = build_encoder(case, inp, dev)
    assert enc.fused_tail == (case["pooling"] == "mean" and case["pred_head"] == "default" and not case.get("feat_proj"))
    fs, loss = [], 0
    for X, G in zip(inp["bags"], inp["G"]):
        f, attn = enc(X.unsqueeze(0).to(dev), ret_with_attn=True)
        fs.append(f.detach())
        loss = loss + (f * G.to(dev)).sum()
    loss.backward()
    A, ext = attn if isinstance(attn, tuple) else (attn, None)
    assert A.shape == (1, case["P"], inp["bags"][-1].shape[0])
    close(torch.cat(fs, 0).cpu().numpy(), gold, "f")
    close(enc.Q.grad.cpu().numpy(), gold, "d_Q")
    np.testing.assert_allclose(A[0, :, :64].cpu().numpy(), gold["attn_head_f64"], rtol=3e-4, atol=1e-9)
    if ext is not None:
        close(ext.cpu().numpy(), gold, "pool_scores")
    if case["pooling"] == "weight":
        close(enc.query_pooling.grad.cpu().numpy(), gold, "d_pool_weight")
    elif case["pooling"] in ("attention", "gated_attention"):
        last = "attention.2.weight" if case["pooling"] == "attention" else "fc2.weight"
        close(dict(enc.query_pooling.named_parameters())[last].grad.cpu().numpy(), gold, "d_pool_last")
    if case["pred_head"] != "Identity":
        close(enc.visual_adapter.bias.grad.cpu().numpy(), gold, "d_b")
    if case.get("feat_proj"):
        pg = dict(enc.feat_proj.named_parameters())
        close(pg["projecter.0.weight"].grad[:4].cpu().numpy(), gold, "d_proj_w_rows")
        close(pg["projecter.1.weight"].grad.cpu().numpy(), gold, "d_proj_ln_w")
        close(pg["projecter.0.weight"].grad.double().norm().cpu().numpy(), gold, "d_proj_w_fro", floor=1e-5)


@pytest.mark.parametrize("P,dtype", [(1, torch.float32), (3, torch.float32), (4, torch.float32), (5, torch.float32),
                                     (7, torch.float32), (12, torch.float32), (16, torch.float32),
                                     (4, torch.bfloat16), (12, torch.bfloat16)])
@pytest.mark.parametrize("prenorm", [False, True])
def test_pooled_op_packed_batch_against_oracle(P, dtype, prenorm, dev):
    """ops.pooled on a ragged packed batch (incl. an empty and a 1-row bag) with a random gradient per prototype,
    against fp64 autograd of the oracle's formula; both streaming kernels in the forward."""
    from oracle import vlsa_oracle as O
    from vlsa_b200 import ops, synth
    g = torch.Generator().manual_seed(900 + P)
    sizes = [700, 0, 1, 2500, 33]
    bags = [synth.make_bag("g1", n, 4000 + i).to(dtype) for i, n in enumerate(sizes)]
    Q = torch.nn.functional.normalize(torch.randn(P, 512, generator=g), dim=-1) + 0.5 * torch.randn(P, 512, generator=g)
    if prenorm:                                         # a difference of unit rows, as the gated query passes it
        Qn = torch.nn.functional.normalize(torch.randn(P + 1, 512, generator=g) + Q.mean(0), dim=-1)
        Q = (Qn[:-1] - Qn[-1:]).contiguous()
    dO = torch.randn(len(sizes), P, 512, generator=g)
    # oracle: fp64 on the same (possibly bf16-rounded) inputs
    Q64 = Q.double().requires_grad_(True)
    qdir = Q64 if prenorm else torch.nn.functional.normalize(Q64, dim=-1)
    outs = []
    for X in bags:
        X64 = X.double()
        if X64.shape[0] == 0:
            outs.append(torch.zeros(P, 512, dtype=torch.float64))
            continue
        S = float(O.coattn_scale()) * qdir @ torch.nn.functional.normalize(X64, dim=-1).t()
        outs.append(torch.softmax(S, dim=-1) @ X64)
    O64 = torch.stack(outs)
    (O64 * dO.double()).sum().backward()
    Xp = torch.cat(bags).to(dev)
    plan = ops.make_plan(sizes, dev)
    variants = ("simt", "tc") if dtype == torch.float32 else ("simt",)
    try:
        for variant in variants:
            ops.set_agg_variant(variant)
            Qd = Q.to(dev).requires_grad_(True)
            Og, ml = ops.pooled(Xp, plan, Qd, prenorm)
            (Og * dO.to(dev)).sum().backward()
            scale_o = O64.abs().max().item()
            assert (Og.detach().cpu().double() - O64.detach()).abs().max().item() <= 3e-6 * scale_o, variant
            err = (Qd.grad.cpu().double() - Q64.grad).abs().max().item()
            assert err <= 2e-4 * Q64.grad.abs().max().item(), (variant, err, Q64.grad.abs().max().item())
    finally:
        ops.set_agg_variant(None)


def test_vlsa_module_with_variant_encoder_trains(dev):
    """VLSA.forward / forward_packed with a gated, attention-pooled encoder: same logits per bag and packed, loss
    gradients reach the gate row, the pooling module and the adapter."""
    from vlsa_b200 import ops, synth
    from vlsa_b200.model import VLSA
    P, R = 4, 4
    pr = synth.make_params(P, R, 77)
    img = dict(name="VLFAN", dim_in=512, dim_hid=64, use_feat_proj=False, query="Parameter", num_query=P,
               gated_query=True, query_pooling="attention", pred_head="default")
    net = VLSA({"name": "mahmoodlab/conch"}, img, {"name": "CoOp"}, text_features=pr["text_features"],
               logit_scale_init=float(pr["logit_scale"])).to(dev)
    sizes = [900, 130, 2048]
    bags = [synth.make_bag("g1", n, 600 + i) for i, n in enumerate(sizes)]
    single = torch.cat([net(b.unsqueeze(0).to(dev))[0] for b in bags])
    plan = ops.make_plan(sizes, dev)
    logits, g, Tn, inc = net.forward_packed(torch.cat(bags).to(dev), plan)
    np.testing.assert_allclose(logits.detach().cpu().numpy(), single.detach().cpu().numpy(), rtol=1e-5, atol=1e-5)
    t = torch.tensor([0, 3, 1], device=dev)
    e = torch.tensor([1, 0, 1], device=dev)
    total = ops.surv_loss(logits, t, e, net.logit_scale)[0]
    total.backward()
    pool = dict(net.mil_encoder.query_pooling.named_parameters())
    grads = [net.mil_encoder.Q.grad, net.mil_encoder.visual_adapter.weight.grad, net.logit_scale.grad,
             pool["attention.0.weight"].grad, pool["attention.2.weight"].grad]
    for gr in grads:
        assert gr is not None and torch.isfinite(gr).all() and gr.abs().max() > 0
    assert pool["attention.2.bias"].grad.abs().max() <= 1e-6       # a softmax ignores a common shift of its logits
    assert net.mil_encoder.Q.grad.shape == (P + 1, 512) and net.mil_encoder.Q.grad[-1].abs().max() > 0


@pytest.mark.parametrize("P,prenorm", [(1, False), (4, False), (12, True), (16, False)])
def test_gradient_wrt_patch_rows_against_oracle(P, prenorm, dev):
    """vlsa_agg_pooled_bwd_dx through ops.pooled (any gradient per prototype) and through ops.encode (mean + Linear):
    dX against fp64 autograd of the oracle's formula, ragged bags incl. an empty and a 1-row bag."""
    from oracle import vlsa_oracle as O
    from vlsa_b200 import ops, synth
    g = torch.Generator().manual_seed(1200 + P)
    sizes = [300, 0, 1, 700, 129]
    bags = [synth.make_bag("g1" if i % 2 else "g0", n, 5000 + i) for i, n in enumerate(sizes)]
    Q = torch.nn.functional.normalize(torch.randn(P, 512, generator=g), dim=-1) + 0.5 * torch.randn(P, 512, generator=g)
    if prenorm:
        Qn = torch.nn.functional.normalize(torch.randn(P + 1, 512, generator=g), dim=-1)
        Q = (Qn[:-1] - Qn[-1:]).contiguous()
    W = torch.randn(512, 512, generator=g) / 22.0
    bias = 0.1 * torch.randn(512, generator=g)
    dO = torch.randn(len(sizes), P, 512, generator=g)
    df = torch.randn(len(sizes), 512, generator=g)
    X64 = torch.cat(bags).double().requires_grad_(True)
    qdir = Q.double() if prenorm else torch.nn.functional.normalize(Q.double(), dim=-1)
    outs, at = [], 0
    for n in sizes:
        Xb = X64[at:at + n]
        at += n
        if n == 0:
            outs.append(torch.zeros(P, 512, dtype=torch.float64))
            continue
        S = float(O.coattn_scale()) * qdir @ torch.nn.functional.normalize(Xb, dim=-1).t()
        outs.append(torch.softmax(S, dim=-1) @ Xb)
    O64 = torch.stack(outs)
    f64 = O64.mean(dim=1) @ W.double().t() + bias.double()
    ref_pooled, = torch.autograd.grad((O64 * dO.double()).sum(), X64, retain_graph=True)
    ref_encode, = torch.autograd.grad((f64 * df.double()).sum(), X64)
    plan = ops.make_plan(sizes, dev)
    Xg = torch.cat(bags).to(dev).requires_grad_(True)
    Og, _ = ops.pooled(Xg, plan, Q.to(dev), prenorm)
    got_pooled, = torch.autograd.grad((Og * dO.to(dev)).sum(), Xg)
    fg, _ = ops.encode(Xg, plan, Q.to(dev), W.to(dev), bias.to(dev), None, prenorm)
    got_encode, = torch.autograd.grad((fg * df.to(dev)).sum(), Xg)
    for got, ref, what in ((got_pooled, ref_pooled, "pooled"), (got_encode, ref_encode, "encode")):
        err = (got.cpu().double() - ref).abs().max().item()
        assert err <= 2e-5 * ref.abs().max().item(), (what, err, ref.abs().max().item())
