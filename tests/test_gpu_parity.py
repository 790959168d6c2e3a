"""GPU parity tests proper: the CUDA path (through the C ABI) vs golden vectors of the reference and vs the
CPU oracle on the same seeded inputs.

Tolerances (north_star): incidence function <= 1e-4 absolute in fp32 (we assert 2e-5), bf16 storage <= 1e-3;
index outputs (argmax / preds) bit-exact.
"""
import numpy as np
import pytest
import torch

from conftest import golden_cases
from golden_util import load_case, rebuild_inputs

pytestmark = pytest.mark.gpu

IF_TOL = 2e-5          # north_star bar is 1e-4
GRAD_RTOL = 2e-4       # relative to the largest |entry| of the reference gradient


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _query(pr):
    return pr["res_ratio"] * pr["residual_features"] + pr["prompt_features"]


def _fwd(ops, bags, pr, dev, dtype=torch.float32):
    X = torch.cat(bags, 0).to(dev).to(dtype).contiguous()
    plan = ops.make_plan([b.shape[0] for b in bags], dev)
    out = ops.aggregate_forward_raw(X, plan, _query(pr).to(dev), pr["W"].to(dev), pr["b"].to(dev),
                                    pr["text_features"].to(dev), pr["logit_scale"].to(dev))
    torch.cuda.synchronize()
    return X, plan, out


@pytest.mark.parametrize("name", golden_cases("single_") + golden_cases("real_"))
def test_forward_vs_reference_golden(name, dev):
    from vlsa_b200 import ops
    case = load_case(name)
    bags, pr, t, e = rebuild_inputs(name, case)
    X, plan, out = _fwd(ops, bags, pr, dev)
    inc = out["incidence"].cpu().numpy()
    # vs the fp64 run of the reference (truth) and vs its fp32 run
    assert np.abs(inc - case["if_f64"]).max() <= IF_TOL
    assert np.abs(inc - case["if_f32"]).max() <= IF_TOL
    np.testing.assert_allclose(out["f"].cpu().numpy(), case["f_f64"], atol=5e-6, rtol=1e-5)
    np.testing.assert_allclose(out["g"].cpu().numpy(), case["g_f64"], atol=2e-6)
    np.testing.assert_allclose(out["logits"].cpu().numpy(), case["logits_f64"], atol=2e-4, rtol=1e-5)
    assert int(inc.argmax()) == int(case["if_f64"].argmax())
    np.testing.assert_allclose(inc.sum(-1), 1.0, atol=1e-6)
    # attention read-out (ret_with_attn=True)
    A = ops.attention_scores(X, _query(pr).to(dev), out["ml"][0]).cpu().numpy()
    k = case["attn_head_f64"].shape[1]
    np.testing.assert_allclose(A[:, :k], case["attn_head_f64"], rtol=2e-4, atol=1e-9)
    np.testing.assert_allclose(A.sum(1), 1.0, atol=1e-5)
    np.testing.assert_allclose(A.max(1), case["attn_max_f64"], rtol=2e-4)
    assert (A.argmax(1) == case["attn_argmax_f64"]).all()


@pytest.mark.parametrize("name", golden_cases("batch_") + golden_cases("real_") + golden_cases("single_")[:7])
def test_forward_backward_loss_vs_reference_golden(name, dev):
    """One `_update_network` worth of work: packed ragged bags -> logits -> fused loss -> gradients."""
    from vlsa_b200 import ops
    case = load_case(name)
    bags, pr, t, e = rebuild_inputs(name, case)
    X = torch.cat(bags, 0).to(dev)
    plan = ops.make_plan([b.shape[0] for b in bags], dev)
    leaf = lambda z: z.detach().clone().to(dev).requires_grad_(True)
    res, W, b, T, ls = (leaf(pr[k]) for k in ("residual_features", "W", "b", "text_features", "logit_scale"))
    Q = pr["res_ratio"] * res + pr["prompt_features"].to(dev)          # prompt_adapter.py:125-126 (torch autograd)
    logits, g, Tn, inc, ml = ops.aggregate(X, plan, Q, W, b, T, ls)
    total, l_ifmle, l_emd, inc2, per = ops.surv_loss(logits, t.to(dev), e.to(dev), ls)
    total.backward()
    torch.cuda.synchronize()
    np.testing.assert_allclose(logits.detach().cpu().numpy(), case["logits_f64"], atol=2e-4, rtol=1e-5)
    assert np.abs(inc.cpu().numpy() - case["if_f64"]).max() <= IF_TOL
    assert np.abs(inc2.cpu().numpy() - case["if_f64"]).max() <= IF_TOL
    np.testing.assert_allclose(total.item(), case["loss_f64"], rtol=2e-5)
    np.testing.assert_allclose(l_ifmle.item(), case["loss_ifmle_f64"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(l_emd.item(), case["loss_emd_f64"], rtol=2e-5, atol=1e-6)

    def close(got, ref, what):
        ref = np.asarray(ref)
        tol = max(GRAD_RTOL * np.abs(ref).max(), 1e-7)      # floor: exact-zero gradients (N=1) carry fp32 noise
        err = np.abs(got.detach().cpu().numpy() - ref).max()
        assert err <= tol, f"{what}: max err {err:.3e} > {tol:.3e}"

    close(res.grad, case["d_residual_f64"], "d_residual")
    close(b.grad, case["d_b_f64"], "d_b")
    close(T.grad, case["d_T_f64"], "d_T")
    close(ls.grad, case["d_logit_scale_f64"], "d_logit_scale")
    close(W.grad[:8], case["d_W_rows_f64"], "d_W rows")
    close(W.grad[:, :8], case["d_W_cols_f64"], "d_W cols")
    np.testing.assert_allclose(W.grad.double().norm().item(), case["d_W_fro_f64"], rtol=1e-4, atol=1e-12)


@pytest.mark.parametrize("name", golden_cases("loss_"))
def test_loss_kernel_vs_reference_golden(name, dev):
    from vlsa_b200 import ops
    case = load_case(name)
    raw = torch.from_numpy(case["raw"]).to(dev).requires_grad_(True)
    ls = torch.tensor(float(case["logit_scale"]), device=dev)
    total, l1, l2, inc, per = ops.surv_loss(raw, torch.from_numpy(case["t"]).to(dev), torch.from_numpy(case["e"]).to(dev), ls)
    total.backward()
    np.testing.assert_allclose(l1.item(), case["ifmle_f64"], rtol=5e-6)
    np.testing.assert_allclose(l2.item(), case["emd_f64"], rtol=5e-6)
    # fp32 conditioning of 1/p_t and 1/(1-CIF_t) at extreme logits: bound relative to the largest entry
    ref = case["d_raw_f64"]
    assert np.abs(raw.grad.cpu().numpy() - ref).max() <= GRAD_RTOL * np.abs(ref).max()
    np.testing.assert_allclose(raw.grad.cpu().numpy(), ref, atol=2e-7, rtol=1e-3)
    np.testing.assert_allclose(per.cpu().numpy().mean(0), [case["ifmle_f64"], case["emd_f64"]], rtol=5e-6)


@pytest.mark.parametrize("name", golden_cases("zeroshot_"))
def test_zero_shot_vs_reference_golden(name, dev):
    from vlsa_b200 import ops
    case = load_case(name)
    bags, pr, _, _ = rebuild_inputs(name, case)
    pred, pooled = ops.logit_pool(bags[0].to(dev), pr["text_features"].to(dev), pr["logit_scale"].to(dev),
                                  str(case["pooling"]))
    np.testing.assert_allclose(pooled.cpu().numpy(), case["logits_f32"], rtol=2e-5, atol=2e-5)
    assert (pred.cpu().numpy() == case["preds"]).all()          # index op: bit-exact


@pytest.mark.parametrize("P,N,kind", [(4, 10000, "g0"), (4, 10000, "g1"), (12, 20000, "g0"), (16, 3001, "g1"),
                                      (8, 50000, "g0"), (5, 257, "g1")])
def test_forward_vs_oracle_seeded(P, N, kind, dev):
    """BASELINE config 2 (N=10k, K=4, fp32) and friends: CUDA vs the CPU oracle in fp64 on the same inputs."""
    from oracle import vlsa_oracle as O
    from vlsa_b200 import ops, synth
    X = synth.make_bag(kind, N, 4242 + N + P)
    pr = synth.make_params(P, P, 99 + P)
    _, plan, out = _fwd(ops, [X], pr, dev)
    c = lambda z: z.double()
    Q64 = O.task_res_query(c(pr["prompt_features"]), c(pr["residual_features"]), pr["res_ratio"])
    logits64, g64, _ = O.vlsa_forward(c(X).unsqueeze(0), Q64, c(pr["W"]), c(pr["b"]), c(pr["text_features"]),
                                      c(pr["logit_scale"]))
    inc64 = O.softmax_converter(logits64).numpy()
    assert np.abs(out["incidence"].cpu().numpy() - inc64).max() <= IF_TOL
    np.testing.assert_allclose(out["g"].cpu().numpy(), g64.numpy(), atol=2e-6)


@pytest.mark.parametrize("P,N,kind", [(4, 50000, "g0"), (4, 50000, "g1"), (4, 100000, "g0"), (4, 100000, "g1"),
                                      (12, 50000, "g1"), (12, 100000, "g0")])
def test_headline_shapes_vs_oracle(P, N, kind, dev):
    """The driver-benchmarked configurations against the fp64 oracle at FULL size: P = R = 4 (BASELINE 'K = 4', the
    CUDA-core kernel) and the shipped P = R = 12 (tcgen05 kernel), N = 50k and 100k, both generators."""
    from oracle import vlsa_oracle as O
    from vlsa_b200 import ops, synth
    X = synth.make_bag(kind, N, 90210 + N + P)
    pr = synth.make_params(P, P, 17 + P)
    _, plan, out = _fwd(ops, [X], pr, dev)
    c = lambda z: z.double()
    Q64 = O.task_res_query(c(pr["prompt_features"]), c(pr["residual_features"]), pr["res_ratio"])
    logits64, g64, _ = O.vlsa_forward(c(X).unsqueeze(0), Q64, c(pr["W"]), c(pr["b"]), c(pr["text_features"]),
                                      c(pr["logit_scale"]))
    assert np.abs(out["incidence"].cpu().numpy() - O.softmax_converter(logits64).numpy()).max() <= IF_TOL
    np.testing.assert_allclose(out["g"].cpu().numpy(), g64.numpy(), atol=2e-6)


@pytest.mark.parametrize("P", [4, 12])
def test_packed_step_of_32_bags_matches_per_bag_calls(P, dev):
    """bench.py's workload (32 bags x 50k rows in ONE launch) against 32 one-bag calls and, for three of the bags, the
    fp64 oracle.  The packed plan cuts chunks differently from a one-bag plan, so the comparison is to summation-order
    tolerance; two copies of the same bag inside the batch must agree bit for bit."""
    from oracle import vlsa_oracle as O
    from vlsa_b200 import ops, synth
    B, N = 32, 50000
    pr = synth.make_params(P, P, 5 + P)
    g = torch.Generator(device=dev).manual_seed(1234 + P)
    X = torch.randn(B * N, 512, generator=g, device=dev) * 1.1
    X[: 3 * N] += 0.7                                            # three bags with a common direction (peaky softmax)
    X[31 * N:] = X[:N]                                           # bag 31 duplicates bag 0
    args = tuple(z.to(dev) for z in (_query(pr), pr["W"], pr["b"], pr["text_features"], pr["logit_scale"]))
    packed = ops.aggregate_forward_raw(X, ops.make_plan([N] * B, dev), *args)
    torch.cuda.synchronize()
    assert torch.equal(packed["incidence"][0], packed["incidence"][31]) and torch.equal(packed["f"][0], packed["f"][31])
    one = ops.make_plan([N], dev)
    for i in range(B):
        single = ops.aggregate_forward_raw(X[i * N:(i + 1) * N], one, *args)
        assert (single["incidence"][0] - packed["incidence"][i]).abs().max().item() <= 2e-6, i
    c = lambda z: z.double()
    Q64 = O.task_res_query(c(pr["prompt_features"]), c(pr["residual_features"]), pr["res_ratio"])
    for i in (0, 7, 30):
        Xi = X[i * N:(i + 1) * N].cpu()
        logits64, _, _ = O.vlsa_forward(c(Xi).unsqueeze(0), Q64, c(pr["W"]), c(pr["b"]), c(pr["text_features"]), c(pr["logit_scale"]))
        assert np.abs(packed["incidence"][i].cpu().numpy() - O.softmax_converter(logits64).numpy()[0]).max() <= IF_TOL, i


@pytest.mark.parametrize("P", [4, 8, 12, 16])
def test_bf16_storage_tolerance(P, dev):
    """BASELINE config 5 (K in {4, 8, 16}, N = 50k, bf16 vs fp32): bf16 storage of X, fp32 accumulate; incidence <= 1e-3
    vs the fp32-input oracle and <= 2e-5 vs the oracle run on the SAME bf16-rounded values (storage rounding is the
    caller's choice, the kernel must add < 1e-4)."""
    from oracle import vlsa_oracle as O
    from vlsa_b200 import ops, synth
    N = 50000
    X = synth.make_bag("g1", N, 777 + P)
    pr = synth.make_params(P, P, 55 + P)
    Xb = X.to(torch.bfloat16)
    _, plan, out = _fwd(ops, [X], pr, dev, dtype=torch.bfloat16)
    c = lambda z: z.double()
    Q64 = O.task_res_query(c(pr["prompt_features"]), c(pr["residual_features"]), pr["res_ratio"])
    args = (Q64, c(pr["W"]), c(pr["b"]), c(pr["text_features"]), c(pr["logit_scale"]))
    inc_same = O.softmax_converter(O.vlsa_forward(c(Xb).unsqueeze(0), *args)[0]).numpy()
    inc_f32 = O.softmax_converter(O.vlsa_forward(c(X).unsqueeze(0), *args)[0]).numpy()
    got = out["incidence"].cpu().numpy()
    assert np.abs(got - inc_same).max() <= IF_TOL
    assert np.abs(got - inc_f32).max() <= 1e-3


def test_properties_full_size(dev):
    """BASELINE full sizes (N=50k/100k, batch of bags): size-independent properties instead of an oracle run.
    (a) split invariance: a bag fed as one bag == the same rows under a different chunk schedule;
    (b) permutation invariance over N; (c) duplicate-bag equality inside a batch; (d) incidence sums to 1;
    (e) run-to-run bit stability (fixed split schedule)."""
    from vlsa_b200 import ops, synth
    P = R = 12
    pr = synth.make_params(P, R, 3)
    X = synth.make_bag("g1", 100000, 31337)
    Xd = X.to(dev)
    args = tuple(z.to(dev) for z in (_query(pr), pr["W"], pr["b"], pr["text_features"], pr["logit_scale"]))
    plan_a = ops.make_plan([100000], dev)
    plan_b = ops.make_plan([100000], dev, sms=8)           # different chunking
    a = ops.aggregate_forward_raw(Xd, plan_a, *args)
    b = ops.aggregate_forward_raw(Xd, plan_b, *args)
    assert plan_a.chunk_rows != plan_b.chunk_rows
    assert (a["incidence"] - b["incidence"]).abs().max().item() <= 2e-6
    a2 = ops.aggregate_forward_raw(Xd, plan_a, *args)
    assert torch.equal(a["incidence"], a2["incidence"]) and torch.equal(a["f"], a2["f"])
    perm = torch.randperm(100000, generator=torch.Generator().manual_seed(1))
    c = ops.aggregate_forward_raw(Xd[perm.to(dev)].contiguous(), plan_a, *args)
    assert (a["incidence"] - c["incidence"]).abs().max().item() <= 2e-6
    np.testing.assert_allclose(a["incidence"].sum(-1).cpu().numpy(), 1.0, atol=1e-6)
    # batch of 4 bags: [50k | 100k(dup of X) | 1 | 50k dup of first]
    X50 = Xd[:50000]
    packed = torch.cat([X50, Xd, Xd[:1], X50], 0).contiguous()
    plan = ops.make_plan([50000, 100000, 1, 50000], dev)
    d = ops.aggregate_forward_raw(packed, plan, *args)
    assert torch.equal(d["incidence"][0], d["incidence"][3])
    assert (d["incidence"][1] - a["incidence"][0]).abs().max().item() <= 2e-6
    single = ops.aggregate_forward_raw(Xd[:1].contiguous(), ops.make_plan([1], dev), *args)
    assert (d["incidence"][2] - single["incidence"][0]).abs().max().item() <= 1e-6


def test_empty_and_ragged_bags(dev):
    """Edge cases: empty bag inside a batch (reference: matmul over empty N gives zeros -> f = b), N=1, N=TN±1."""
    from oracle import vlsa_oracle as O
    from vlsa_b200 import ops, synth
    P = R = 4
    pr = synth.make_params(P, R, 11)
    sizes = [0, 1, 31, 32, 33, 0, 65]
    bags = [synth.make_bag("g1", n, 500 + i) if n else torch.zeros(0, 512) for i, n in enumerate(sizes)]
    _, plan, out = _fwd(ops, bags, pr, dev)
    Q = _query(pr)
    for i, X in enumerate(bags):
        logits, g, _ = O.vlsa_forward(X.double().unsqueeze(0), Q.double(), pr["W"].double(), pr["b"].double(),
                                      pr["text_features"].double(), pr["logit_scale"].double())
        inc = O.softmax_converter(logits).numpy()[0]
        assert np.abs(out["incidence"][i].cpu().numpy() - inc).max() <= IF_TOL, f"bag {i} (N={sizes[i]})"


def test_argument_errors(dev):
    from vlsa_b200 import ops, synth
    from vlsa_b200._lib import VlsaLibraryError
    pr = synth.make_params(4, 4, 1)
    X = torch.randn(10, 512, device=dev)
    args = tuple(z.to(dev) for z in (_query(pr), pr["W"], pr["b"], pr["text_features"], pr["logit_scale"]))
    plan = ops.make_plan([10], dev)
    with pytest.raises(ValueError):
        ops.aggregate_forward_raw(X[:, :256].contiguous(), plan, *args)
    with pytest.raises(ValueError):
        ops.aggregate_forward_raw(X.cpu(), plan, *args)
    with pytest.raises(ValueError):
        ops.aggregate_forward_raw(X, ops.make_plan([11], dev), *args)
    with pytest.raises(ValueError):
        ops.aggregate_forward_raw(X, plan, torch.randn(17, 512, device=dev), *args[1:])
    with pytest.raises(VlsaLibraryError):
        ops.aggregate_forward_raw(X, plan, *args, workspace=torch.empty(16, dtype=torch.uint8, device=dev))
    with pytest.raises(NotImplementedError):
        ops.logit_pool(X, args[3], args[4], "logit_median")


@pytest.mark.parametrize("P,sizes,kind,dtype", [(4, [10000], "g1", "fp32"), (12, [2798, 1000, 37], "g1", "fp32"),
                                                (16, [5000, 33], "g0", "fp32"), (1, [700], "g0", "fp32"), (7, [50000], "g1", "fp32"),
                                                (4, [10000, 33], "g1", "bf16"), (12, [2798, 1000, 37, 1, 17], "g1", "bf16"),
                                                (16, [5000, 33], "g0", "bf16"), (8, [50000], "g1", "bf16")])
def test_tensor_core_and_cuda_core_kernels_agree(P, sizes, kind, dtype, dev):
    """The streaming kernels (tcgen05: register-staged for fp32 rows, TMA-fed for bf16 rows; CUDA-core) are
    complete implementations of the same pass; every case runs forward + loss + backward through each and the results
    must agree to the parity tolerances, and each must match the fp64 oracle (bf16: on the same rounded values)."""
    from oracle import vlsa_oracle as O
    from vlsa_b200 import ops, synth
    bags = [synth.make_bag(kind, n, 900 + i + P) for i, n in enumerate(sizes)]
    if dtype == "bf16":
        bags = [b.to(torch.bfloat16).float() for b in bags]          # the stored values ARE the inputs
    pr = synth.make_params(P, P, 21 + P)
    t, e = synth.make_labels(len(sizes), P, 5)
    X = torch.cat(bags, 0).to(dev)
    if dtype == "bf16":
        X = X.to(torch.bfloat16)
    plan = ops.make_plan(sizes, dev)
    res = {}
    try:
        for variant in ("simt", "tc"):
            ops.set_agg_variant(variant)
            leaf = lambda z: z.detach().clone().to(dev).requires_grad_(True)
            r, W, b, T, ls = (leaf(pr[k]) for k in ("residual_features", "W", "b", "text_features", "logit_scale"))
            Q = pr["res_ratio"] * r + pr["prompt_features"].to(dev)
            logits, g, Tn, inc, ml = ops.aggregate(X, plan, Q, W, b, T, ls)
            total, *_ = ops.surv_loss(logits, t.to(dev), e.to(dev), ls)
            total.backward()
            torch.cuda.synchronize()
            res[variant] = dict(inc=inc.detach().cpu().numpy(), g=g.detach().cpu().numpy(), loss=total.item(),
                                d_res=r.grad.cpu().numpy(), d_W=W.grad.cpu().numpy(), d_T=T.grad.cpu().numpy())
    finally:
        ops.set_agg_variant(None)
    ref = O.forward_with_grads(bags, pr["prompt_features"], pr["residual_features"], pr["W"], pr["b"],
                               pr["text_features"], pr["logit_scale"], t, e, dtype=torch.float64)
    inc64 = torch.softmax(ref["logits"], -1).numpy()
    for variant, out in res.items():
        assert np.abs(out["inc"] - inc64).max() <= IF_TOL, variant
        np.testing.assert_allclose(out["loss"], ref["loss"].item(), rtol=2e-5, err_msg=variant)
        for key, rk in (("d_res", "d_residual"), ("d_W", "d_W"), ("d_T", "d_T")):
            if rk not in ref:
                continue
            r64 = ref[rk].numpy()
            tol = max(GRAD_RTOL * np.abs(r64).max(), 1e-7)
            assert np.abs(out[key] - r64).max() <= tol, f"{variant} {key}"
    a, b_ = res["simt"], res["tc"]
    assert np.abs(a["inc"] - b_["inc"]).max() <= IF_TOL
    np.testing.assert_allclose(a["g"], b_["g"], atol=2e-6)
    assert np.abs(a["d_res"] - b_["d_res"]).max() <= max(GRAD_RTOL * np.abs(a["d_res"]).max(), 1e-7)


def test_backward_accuracy_regression_case(dev):
    """An ill-conditioned optimizer step: a censored sample with almost no probability mass left after its time bin
    makes d loss / d logits amplify a 1e-6 difference of a logit about a thousand times, so the reference's own fp32
    arithmetic is 5e-4 away from its fp64 run here — and so is any fp32-grade forward, by the luck of its last bit
    (round 1 bounded the tcgen05 path at 2e-5 on exactly this case; the same kernel is 7.5e-4 off on [2798, 37]).
    What the kernels can be held to, and are:
      (a) with the UPSTREAM gradient d loss / d logits fixed (taken from the fp64 reference), the gradients of both
          streaming kernels are within 2e-5 of the fp64 reference — the backward itself loses nothing (this is the
          check that exposed a 700x loss of accuracy when the lo-plane products shared a TMEM accumulator with the
          hi-plane ones);
      (b) through the real loss, both stay within a small multiple of what fp32 arithmetic in the reference's order
          gives."""
    from oracle import vlsa_oracle as O
    from vlsa_b200 import ops, synth
    P = R = 12
    sizes = [2798, 1000, 37]
    bags = [synth.make_bag("g1", n, 100 + i) for i, n in enumerate(sizes)]
    pr = synth.make_params(P, R, 7)
    t, e = synth.make_labels(len(sizes), R, 9)
    ref = O.forward_with_grads(bags, pr["prompt_features"], pr["residual_features"], pr["W"], pr["b"],
                               pr["text_features"], pr["logit_scale"], t, e, dtype=torch.float64)
    ref32 = O.forward_with_grads(bags, pr["prompt_features"], pr["residual_features"], pr["W"], pr["b"],
                                 pr["text_features"], pr["logit_scale"], t, e, dtype=torch.float32)
    gref = ref["d_residual"].numpy()
    err_ref32 = np.abs(ref32["d_residual"].double().numpy() - gref).max() / np.abs(gref).max()
    assert err_ref32 > 1e-4                      # the case is ill-conditioned for fp32 arithmetic (measured 5e-4)
    # fp64 reference of (a): d (sum logits * G) with G = d loss / d logits of the fp64 run
    c = lambda z: z.double()
    res64 = c(pr["residual_features"]).requires_grad_(True)
    Q64 = O.task_res_query(c(pr["prompt_features"]), res64, pr["res_ratio"])
    lg64 = torch.cat([O.vlsa_forward(c(X).unsqueeze(0), Q64, c(pr["W"]), c(pr["b"]), c(pr["text_features"]),
                                     c(pr["logit_scale"]))[0] for X in bags], 0)
    lg_leaf = lg64.detach().clone().requires_grad_(True)
    O.objective_loss(lg_leaf, t, e, c(pr["logit_scale"]).exp()).backward()
    G = lg_leaf.grad.clone()
    (lg64 * G).sum().backward()
    gfix = res64.grad.numpy()
    X = torch.cat(bags, 0).to(dev)
    plan = ops.make_plan(sizes, dev)
    try:
        for variant in ("tc", "simt"):
            ops.set_agg_variant(variant)
            leaf = lambda z: z.detach().clone().to(dev).requires_grad_(True)
            res, W, b, T, ls = (leaf(pr[k]) for k in ("residual_features", "W", "b", "text_features", "logit_scale"))
            Q = pr["res_ratio"] * res + pr["prompt_features"].to(dev)
            logits, g, Tn, inc, ml = ops.aggregate(X, plan, Q, W, b, T, ls)
            (logits * G.float().to(dev)).sum().backward()
            torch.cuda.synchronize()
            err = np.abs(res.grad.cpu().numpy() - gfix).max() / np.abs(gfix).max()
            assert err <= 2e-5, f"{variant}: fixed upstream gradient, d_residual relative error {err:.2e}"
            res.grad = None
            logits, g, Tn, inc, ml = ops.aggregate(X, plan, pr["res_ratio"] * res + pr["prompt_features"].to(dev), W, b, T, ls)
            total, *_ = ops.surv_loss(logits, t.to(dev), e.to(dev), ls)
            total.backward()
            torch.cuda.synchronize()
            err = np.abs(res.grad.cpu().numpy() - gref).max() / np.abs(gref).max()
            assert err <= 3 * err_ref32, f"{variant}: d_residual relative error {err:.2e} (fp32 reference ops: {err_ref32:.2e})"
            assert abs(total.item() - ref["loss"].item()) <= 1e-5 * abs(ref["loss"].item()), variant
    finally:
        ops.set_agg_variant(None)
