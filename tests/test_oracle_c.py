"""CPU: the plain-C fp64 oracle vs the golden vectors produced by the reference."""
import numpy as np
import pytest
import torch

from conftest import golden_cases
from golden_util import load_case, rebuild_inputs
from oracle import c_binding as OC

SCALE = float((torch.ones([]) * np.log(100)).exp())
CASES = [n for n in golden_cases("single_") + golden_cases("real_") if "N5000" not in n]


@pytest.mark.parametrize("name", CASES)
def test_c_forward_matches_reference_fp64(name):
    case = load_case(name)
    bags, pr, t, e = rebuild_inputs(name, case)
    Q = (pr["res_ratio"] * pr["residual_features"].double() + pr["prompt_features"].double()).float()
    # the fp64 reference run builds Q in fp64; rounding Q to fp32 costs ~1e-7 on the incidence
    f, g, logits, inc, A = OC.forward(bags[0].numpy(), Q.numpy(), pr["W"].numpy(), pr["b"].numpy(),
                                      pr["text_features"].numpy(), SCALE, float(pr["logit_scale"]), want_attn=True)
    np.testing.assert_allclose(inc, case["if_f64"][0], atol=2e-6)
    np.testing.assert_allclose(g, case["g_f64"][0], atol=5e-7)
    k = case["attn_head_f64"].shape[1]
    np.testing.assert_allclose(A[:, :k], case["attn_head_f64"], rtol=2e-4, atol=1e-12)
    assert (A.argmax(1) == case["attn_argmax_f64"]).all()


@pytest.mark.parametrize("name", golden_cases("loss_"))
def test_c_losses_match_reference_fp64(name):
    case = load_case(name)
    p = torch.softmax(torch.from_numpy(case["raw"]).double(), -1).numpy()
    out = OC.losses(p, case["t"], case["e"], float(np.exp(np.float64(case["logit_scale"]))))
    np.testing.assert_allclose(out[0], case["ifmle_f64"], rtol=1e-9)
    np.testing.assert_allclose(out[1], case["emd_f64"], rtol=1e-9)


@pytest.mark.parametrize("name", golden_cases("zeroshot_"))
def test_c_logit_pool_matches_reference(name):
    case = load_case(name)
    bags, pr, _, _ = rebuild_inputs(name, case)
    pooling = str(case["pooling"])
    mode, k = (0, 0) if pooling == "logit_mean" else (1, 1 if pooling == "logit_max" else int(pooling.split("top")[-1]))
    pooled, pred = OC.logit_pool(bags[0].numpy(), pr["text_features"].numpy(), float(pr["logit_scale"]), mode, k)
    np.testing.assert_allclose(pooled, case["logits_f32"][0], rtol=2e-5, atol=2e-5)
    assert pred == int(case["preds"][0])
