"""CPU: the oracle restatement vs golden vectors produced by the reference itself."""
import numpy as np
import pytest
import torch

from conftest import golden_cases
from golden_util import load_case, rebuild_inputs
from oracle import vlsa_oracle as O

torch.set_num_threads(4)

FWD_CASES = golden_cases("single_") + golden_cases("real_")


@pytest.mark.parametrize("name", FWD_CASES)
def test_forward_matches_reference(name):
    case = load_case(name)
    bags, pr, t, e = rebuild_inputs(name, case)
    X = bags[0].unsqueeze(0)
    Q = O.task_res_query(pr["prompt_features"], pr["residual_features"], pr["res_ratio"])
    logits, g, Tn = O.vlsa_forward(X, Q, pr["W"], pr["b"], pr["text_features"], pr["logit_scale"])
    # same ATen ops in the same order: the fp32 oracle reproduces the reference bit for bit
    # (thread-count dependent reductions allowed for: tolerance 2e-6 instead of 0)
    np.testing.assert_allclose(logits.numpy(), case["logits_f32"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(O.softmax_converter(logits).numpy(), case["if_f32"], atol=2e-6)
    np.testing.assert_allclose(g.numpy(), case["g_f32"], atol=2e-6)
    f, A = O.vlfan_forward(X, Q, pr["W"], pr["b"], ret_with_attn=True)
    np.testing.assert_allclose(f.numpy(), case["f_f32"], rtol=1e-5, atol=1e-6)
    k = case["attn_head_f32"].shape[1]
    np.testing.assert_allclose(A[0, :, :k].numpy(), case["attn_head_f32"], rtol=1e-4, atol=1e-9)
    assert (A[0].argmax(dim=1).numpy() == case["attn_argmax_f32"]).all()
    # and the fp64 run of the oracle agrees with the fp64 run of the reference
    c64 = lambda z: z.double()
    Q64 = O.task_res_query(c64(pr["prompt_features"]), c64(pr["residual_features"]), pr["res_ratio"])
    logits64, _, _ = O.vlsa_forward(c64(X), Q64, c64(pr["W"]), c64(pr["b"]), c64(pr["text_features"]),
                                    c64(pr["logit_scale"]))
    np.testing.assert_allclose(O.softmax_converter(logits64).numpy(), case["if_f64"], atol=1e-12)


@pytest.mark.parametrize("name", golden_cases("batch_") + FWD_CASES[:6] + golden_cases("real_"))
def test_grads_match_reference(name):
    case = load_case(name)
    bags, pr, t, e = rebuild_inputs(name, case)
    for tag, dtype, tol in (("f32", torch.float32, 2e-5), ("f64", torch.float64, 1e-11)):
        out = O.forward_with_grads(bags, pr["prompt_features"], pr["residual_features"], pr["W"], pr["b"],
                                   pr["text_features"], pr["logit_scale"], t, e, dtype=dtype)
        scale = lambda ref: tol * max(1.0, float(np.abs(ref).max()))
        np.testing.assert_allclose(out["loss"].numpy(), case[f"loss_{tag}"], rtol=tol, atol=tol)
        for key in ("d_residual", "d_b", "d_T", "d_logit_scale"):
            ref = case[f"{key}_{tag}"]
            np.testing.assert_allclose(out[key].numpy(), ref, atol=scale(ref), rtol=10 * tol, err_msg=key)
        ref = case[f"d_W_rows_{tag}"]
        np.testing.assert_allclose(out["d_W"][:8].numpy(), ref, atol=scale(ref), rtol=10 * tol)
        ref = case[f"d_W_cols_{tag}"]
        np.testing.assert_allclose(out["d_W"][:, :8].numpy(), ref, atol=scale(ref), rtol=10 * tol)
        np.testing.assert_allclose(out["d_W"].double().norm().item(), case[f"d_W_fro_{tag}"], rtol=100 * tol)


@pytest.mark.parametrize("name", golden_cases("zeroshot_"))
def test_zero_shot_matches_reference(name):
    case = load_case(name)
    bags, pr, _, _ = rebuild_inputs(name, case)
    preds, pooled, g, Tn = O.vlsa_forward_zero_shot(bags[0].unsqueeze(0), pr["text_features"], pr["logit_scale"],
                                                     str(case["pooling"]))
    np.testing.assert_allclose(pooled.numpy(), case["logits_f32"], rtol=1e-6, atol=1e-6)
    assert (preds.numpy() == case["preds"]).all()


@pytest.mark.parametrize("name", golden_cases("loss_"))
def test_losses_match_reference(name):
    case = load_case(name)
    t, e = torch.from_numpy(case["t"]), torch.from_numpy(case["e"])
    for tag, dtype, tol in (("f32", torch.float32, 1e-6), ("f64", torch.float64, 1e-13)):
        raw = torch.from_numpy(case["raw"]).to(dtype).requires_grad_(True)
        ls = torch.tensor(float(case["logit_scale"]), dtype=dtype).exp()
        p = O.softmax_converter(raw)
        l1 = O.surv_ifmle(p, t, e)
        l2 = O.surv_emd(p, t, e, ls)
        (l1 + l2).backward()
        np.testing.assert_allclose(l1.item(), case[f"ifmle_{tag}"], rtol=tol * 10)
        np.testing.assert_allclose(l2.item(), case[f"emd_{tag}"], rtol=tol * 10)
        np.testing.assert_allclose(raw.grad.numpy(), case[f"d_raw_{tag}"], atol=tol)


def test_decoupled_identity(real_bag, ckpt_params):
    """notebook cell 12 == cell 17 identity (utils/model_inference.py:118-131): weight-free property."""
    from vlsa_b200 import synth
    pr = synth.make_params(12, 12, 99, w=ckpt_params["W"], b=ckpt_params["b"])
    Q = O.task_res_query(pr["prompt_features"], ckpt_params["residual_features"], 0.5)
    c = lambda z: z.double()
    A, probs, probs_2, dec = O.decoupled_similarity(c(real_bag).unsqueeze(0), c(Q), c(pr["W"]), c(pr["b"]),
                                                    c(pr["text_features"]), c(ckpt_params["logit_scale"]))
    # identical up to the bias term handling: A rows sum to 1 so b passes through the mean
    np.testing.assert_allclose(probs.numpy(), probs_2.numpy(), atol=1e-10)
    np.testing.assert_allclose(A.sum(dim=1).numpy(), np.ones(12), atol=1e-12)


# ---- VLFAN variants ----------------------------------------------------------------------------
from golden_util import VARIANT_CASES, variant_inputs, variant_name  # noqa: E402


def oracle_variant(case, inp, dtype):
    """The oracle's variant forward over the bags of a case + autograd of sum(f * G): -> dict like the golden record."""
    c = lambda z: z.to(dtype)
    Q = c(inp["Q"]).clone().requires_grad_(True)
    pool = {k: c(v).clone().requires_grad_(True) for k, v in inp["pool"].items()}
    b = c(inp["b"]).clone().requires_grad_(True)
    proj = None if inp["proj"] is None else {k: c(v).clone().requires_grad_(True) for k, v in inp["proj"].items()}
    scale = torch.tensor(O.coattn_scale(), dtype=dtype)
    fs, loss = [], 0
    for X, G in zip(inp["bags"], inp["G"]):
        f, A, ext, _ = O.vlfan_forward_variant(c(X).unsqueeze(0), Q, c(inp["W"]), b, case["gated"], case["pooling"],
                                               pool, case["pred_head"], scale=scale, proj_params=proj)
        fs.append(f.detach())
        loss = loss + (f * c(G)).sum()
    loss.backward()
    rec = {"f": torch.cat(fs, 0).numpy(), "d_Q": Q.grad.numpy(), "attn_head": A[0, :, :64].numpy()}
    if ext is not None:
        rec["pool_scores"] = ext.numpy()
    if case["pooling"] == "weight":
        rec["d_pool_weight"] = pool["weight"].grad.numpy()
    elif case["pooling"] in ("attention", "gated_attention"):
        rec["d_pool_last"] = pool["attention.2.weight" if case["pooling"] == "attention" else "fc2.weight"].grad.numpy()
    if case["pred_head"] != "Identity":
        rec["d_b"] = b.grad.numpy()
    if proj is not None:
        rec["d_proj_w_rows"] = proj["projecter.0.weight"].grad[:4].numpy()
        rec["d_proj_w_fro"] = proj["projecter.0.weight"].grad.double().norm().numpy()
        rec["d_proj_ln_w"] = proj["projecter.1.weight"].grad.numpy()
    return rec


@pytest.mark.parametrize("case", VARIANT_CASES, ids=variant_name)
def test_variant_oracle_matches_reference(case):
    gold = load_case(variant_name(case))
    inp = variant_inputs(case)
    np.testing.assert_allclose([x.double().sum().item() for x in inp["bags"]], gold["x_sum"], rtol=1e-12)
    for tag, dtype, tol in (("f32", torch.float32, 2e-5), ("f64", torch.float64, 1e-11)):
        rec = oracle_variant(case, inp, dtype)
        for key, val in rec.items():
            ref = gold[f"{key}_{tag}"]
            np.testing.assert_allclose(val, ref, rtol=10 * tol, atol=tol * max(1.0, float(np.abs(ref).max())),
                                       err_msg=f"{key}_{tag}")


@pytest.mark.parametrize("name", golden_cases("interp_"))
def test_interpretation_oracle_matches_reference_own_routine(name):
    """The oracle's decoupled similarities / Shapley values vs the reference's OWN calc_text_img_similarity and
    evaluate_prototype_shap_imp (utils/model_inference.py:21-144), run unmodified by make_golden_r02.py."""
    from golden_util import rebuild_interp
    case = load_case(name)
    X, pr = rebuild_interp(case)
    c = lambda z: z.double()
    Q64 = O.task_res_query(c(pr["prompt_features"]), c(pr["residual_features"]), pr["res_ratio"])
    A, probs, probs2, dec = O.decoupled_similarity(c(X).unsqueeze(0), Q64, c(pr["W"]), c(pr["b"]), c(pr["text_features"]),
                                                   c(pr["logit_scale"]))
    k = case["cottn_head_f64"].shape[1]
    np.testing.assert_allclose(A[:, :k].numpy(), case["cottn_head_f64"], rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(probs.numpy(), case["probs_f64"], atol=1e-12)
    np.testing.assert_allclose(probs2.numpy(), case["probs2_f64"], atol=1e-12)
    ls = float(pr["logit_scale"].double().exp())
    np.testing.assert_allclose(torch.softmax(ls * dec, dim=0).numpy(), case["imp_f64"], atol=1e-12)
    if int(case["P"]) <= 8:                       # the subset-by-subset port is O(P 4^P) Python
        np.testing.assert_allclose(O.prototype_shap_imp(dec.float(), ls).numpy(), case["shap_f64"], atol=2e-5)
    # the reference's identity (notebook cell 12 == cell 17): both routes give the same prediction
    np.testing.assert_allclose(case["probs_f64"], case["probs2_f64"], atol=1e-10)


@pytest.mark.parametrize("name", golden_cases("zeroshot2_"))
def test_zero_shot_feature_pooling_matches_reference(name):
    from golden_util import rebuild_zeroshot2
    case = load_case(name)
    X, pr = rebuild_zeroshot2(case)
    preds, pooled, g, Tn = O.vlsa_forward_zero_shot(X.unsqueeze(0), pr["text_features"], pr["logit_scale"], str(case["pooling"]))
    np.testing.assert_allclose(pooled.numpy(), case["logits_f32"], rtol=2e-5, atol=2e-5)
    assert tuple(g.shape) == tuple(case["feats_shape"])
    np.testing.assert_allclose(g[:8].numpy(), case["feats_head_f64"], atol=2e-6)
    np.testing.assert_allclose(Tn.numpy(), case["Tn_f32"], atol=1e-6)
