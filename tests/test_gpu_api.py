"""GPU: the reference-facing Python API (VLSA / VLFAN / handler / loader / host-buffer entry) on top of the C ABI."""
import numpy as np
import pytest
import torch

from conftest import golden_cases
from golden_util import load_case, rebuild_inputs

pytestmark = pytest.mark.gpu
IF_TOL = 2e-5


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def build_net(pr, P, R, dev, encoder="VLFAN", pooling=None):
    from vlsa_b200.model import VLSA
    if encoder == "VLFAN":
        img = dict(name="VLFAN", dim_in=512, dim_hid=256, use_feat_proj=False, drop_rate=0.25, query="Text", num_query=P,
                   gated_query=False, query_pooling="mean", pred_head="default", query_text_method="TaskRes",
                   query_text_res_ratio=pr["res_ratio"])
    else:
        img = dict(name="FeatMIL", pooling=pooling)
    net = VLSA({"name": "mahmoodlab/conch"}, img, {"name": "CoOp"}, text_features=pr["text_features"],
               query_prompt_features=pr["prompt_features"], logit_scale_init=float(pr["logit_scale"]),
               vlsa_api="CONCH", path_clip_model=None).to(dev)
    if encoder == "VLFAN":
        with torch.no_grad():
            net.mil_encoder.Q.residual_features.copy_(pr["residual_features"])
            net.mil_encoder.visual_adapter.weight.copy_(pr["W"])
            net.mil_encoder.visual_adapter.bias.copy_(pr["b"])
    return net


@pytest.mark.parametrize("name", golden_cases("real_") + ["single_P12_R12_g1_N2798", "single_P7_R13_g0_N1000"])
def test_vlsa_forward_module_matches_reference(name, dev):
    """VLSA.forward(X[1,N,512]) -> (logits, image_features, text_features) and the handler's softmax (config 1)."""
    case = load_case(name)
    bags, pr, t, e = rebuild_inputs(name, case)
    P, R = int(case["P"]), int(case["R"])
    net = build_net(pr, P, R, dev)
    X = bags[0].unsqueeze(0).to(dev)
    logits, g, Tn = net(X)
    assert logits.shape == (1, R) and g.shape == (1, 512) and Tn.shape == (R, 512)
    inc = torch.softmax(logits, -1)
    assert np.abs(inc.detach().cpu().numpy() - case["if_f64"]).max() <= IF_TOL
    np.testing.assert_allclose(g.detach().cpu().numpy(), case["g_f64"], atol=2e-6)
    np.testing.assert_allclose(Tn.detach().norm(dim=-1).cpu().numpy(), 1.0, atol=1e-6)
    # encoder alone + attention (mil_encoder(X, ret_with_attn=True), utils/model_inference.py:118)
    f, A = net.mil_encoder(X, ret_with_attn=True)
    assert f.shape == (1, 512) and A.shape == (1, P, X.shape[1])
    np.testing.assert_allclose(f.detach().cpu().numpy(), case["f_f64"], atol=5e-6, rtol=1e-5)
    k = case["attn_head_f64"].shape[1]
    np.testing.assert_allclose(A[0, :, :k].cpu().numpy(), case["attn_head_f64"], rtol=2e-4, atol=1e-9)
    # gradients flow to the reference's trainable tensors
    loss = logits.square().sum() + f.sum()
    loss.backward()
    for p in (net.logit_scale, net.mil_encoder.visual_adapter.weight, net.mil_encoder.visual_adapter.bias,
              net.mil_encoder.Q.residual_features):
        assert p.grad is not None and torch.isfinite(p.grad).all()


def test_encoder_only_backward_matches_autograd_of_oracle(dev):
    from oracle import vlsa_oracle as O
    from vlsa_b200 import synth
    P = 8
    pr = synth.make_params(P, P, 21)
    X = synth.make_bag("g1", 3000, 77)
    net = build_net(pr, P, P, dev)
    f = net.mil_encoder(X.unsqueeze(0).to(dev))
    wvec = torch.linspace(-1, 1, 512, device=dev)
    (f * wvec).sum().backward()
    res = pr["residual_features"].double().requires_grad_(True)
    W = pr["W"].double().requires_grad_(True)
    b = pr["b"].double().requires_grad_(True)
    Q = O.task_res_query(pr["prompt_features"].double(), res, pr["res_ratio"])
    f64 = O.vlfan_forward(X.double().unsqueeze(0), Q, W, b)
    (f64 * wvec.cpu().double()).sum().backward()
    for got, ref, what in ((net.mil_encoder.Q.residual_features.grad, res.grad, "dQ"),
                           (net.mil_encoder.visual_adapter.weight.grad, W.grad, "dW"),
                           (net.mil_encoder.visual_adapter.bias.grad, b.grad, "db")):
        err = (got.cpu().double() - ref).abs().max().item()
        assert err <= 2e-4 * ref.abs().max().item(), what


@pytest.mark.parametrize("name", golden_cases("zeroshot_"))
def test_zero_shot_module(name, dev):
    case = load_case(name)
    bags, pr, _, _ = rebuild_inputs(name, case)
    net = build_net(pr, 1, int(case["R"]), dev, encoder="FeatMIL", pooling=str(case["pooling"]))
    logits, feats, Tn = net(bags[0].unsqueeze(0).to(dev))        # one-patch bags included (vlsa.py:195: no pooling)
    assert logits.shape == (1, int(case["R"]))
    np.testing.assert_allclose(logits.cpu().numpy(), case["logits_f32"], rtol=2e-5, atol=2e-5)
    assert feats.shape == (bags[0].shape[0], 512)                  # image_features: the N normalised patches
    np.testing.assert_allclose(feats.norm(dim=-1).cpu().numpy(), 1.0, atol=1e-5)


@pytest.mark.parametrize("name", golden_cases("zeroshot2_"))
def test_zero_shot_all_featmil_branches_vs_reference(name, dev):
    """VLSA.forward with FeatMIL through every branch the reference has (model/vlsa.py:181-198, deepmil.py:51-67):
    feature pooling 'mean' | 'max', one-patch bags, and the returned image_features, against the reference's own
    forward (tests/golden/make_golden_r02.py)."""
    from golden_util import rebuild_zeroshot2
    case = load_case(name)
    X, pr = rebuild_zeroshot2(case)
    R, pooling = int(case["R"]), str(case["pooling"])
    net = build_net(pr, 1, R, dev, encoder="FeatMIL", pooling=pooling)
    logits, feats, Tn = net(X.unsqueeze(0).to(dev))
    assert logits.shape == (1, R) and tuple(feats.shape) == tuple(case["feats_shape"]) and Tn.shape == (R, 512)
    np.testing.assert_allclose(logits.cpu().numpy(), case["logits_f64"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(feats[:8].cpu().numpy(), case["feats_head_f64"], atol=2e-6)
    np.testing.assert_allclose(feats.double().sum(0).cpu().numpy(), case["feats_sum_f64"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(Tn.cpu().numpy(), case["Tn_f32"], atol=1e-6)
    if pooling == "max":                                            # max is exact: same bits as the fp32 reference
        np.testing.assert_allclose(logits.cpu().numpy(), case["logits_f32"], rtol=1e-6, atol=2e-6)
    # bf16 storage of the patches goes through the same entries
    lb, fb, _ = net(X.to(torch.bfloat16).unsqueeze(0).to(dev))
    assert lb.shape == (1, R) and torch.isfinite(lb).all() and (lb - logits).abs().max().item() < 0.5


@pytest.mark.parametrize("name", golden_cases("batch_"))
def test_handler_update_network_and_test_model(name, dev):
    """`_update_network` (one optimizer step on ragged bags) and `test_model` against the reference golden."""
    from vlsa_b200.runner import VLSAHandler
    case = load_case(name)
    bags, pr, t, e = rebuild_inputs(name, case)
    P, R = int(case["P"]), int(case["R"])
    net = build_net(pr, P, R, dev)
    net.pretrained_text_features.requires_grad_(False)
    cfg = dict(task="vlsa", arch="VLSA", net_output_converter="softmax", loss_type="SurvIFMLE-SurvEMD",
               loss_survifmle_weight=1.0, loss_survemd_weight=1.0, opt_name="adam", opt_lr=2e-4, opt_weight_decay=1e-5,
               bp_every_batch=32)
    h = VLSAHandler(cfg, net, device=dev)
    xs = [b.unsqueeze(0).to(dev) for b in bags]
    ys = [torch.stack([t[i], e[i]]).float().reshape(1, 2).to(dev) for i in range(len(bags))]

    # eval path first (parameters untouched): DataLoader-style iterable of (idx, (feats, coords), label)
    loader = [(torch.tensor([[i]]), (xs[i], torch.zeros(1)), ys[i]) for i in range(len(bags))]
    cltor = h.test_model(h.net, loader, "test", bags_per_launch=3)
    assert np.abs(cltor["pred"]["y_hat"].numpy() - case["if_f64"]).max() <= IF_TOL
    np.testing.assert_allclose(cltor["pred"]["raw_y_hat"].numpy(), case["logits_f64"], atol=2e-4, rtol=1e-5)
    assert cltor["pred"]["uid"].tolist() == list(range(len(bags)))
    h.net.train()

    before = {k: v.detach().clone() for k, v in h.net.state_dict().items()}
    loss, preds = h._update_network(xs, ys)
    np.testing.assert_allclose(loss, case["loss_f64"], rtol=2e-5)
    np.testing.assert_allclose(preds.numpy(), case["logits_f64"], atol=2e-4, rtol=1e-5)
    # the reduced gradient bucket holds what the reference's autograd produced
    named = [(n, p) for n, p in h.net.named_parameters() if p.requires_grad]
    got = {}
    for (n, p), sz, at in zip(named, h.bucket.sizes, h.bucket.offsets):
        got[n] = h.bucket.flat[at:at + sz].view_as(p).cpu().numpy()
    for key, ref in (("mil_encoder.Q.residual_features", case["d_residual_f64"]),
                     ("mil_encoder.visual_adapter.bias", case["d_b_f64"]), ("logit_scale", case["d_logit_scale_f64"])):
        assert np.abs(got[key] - ref).max() <= max(2e-4 * np.abs(ref).max(), 1e-7), key
    assert np.abs(got["mil_encoder.visual_adapter.weight"][:8] - case["d_W_rows_f64"]).max() <= 2e-4 * np.abs(case["d_W_rows_f64"]).max()
    # Adam moved every trainable tensor by at most lr per element
    after = h.net.state_dict()
    for k in ("logit_scale", "mil_encoder.visual_adapter.weight", "mil_encoder.Q.residual_features"):
        delta = (after[k] - before[k]).abs().max().item()
        assert 0 < delta <= 2e-4 * 1.01 + 1e-5 * 2e-4 * before[k].abs().max().item() + 1e-9, (k, delta)


def test_forward_host_and_async_loader_match_device_path(dev):
    """Host-buffer C-ABI entry and the pinned ring loader give bit-identical results to device-resident inputs."""
    from vlsa_b200 import ops, synth
    from vlsa_b200.dataset import AsyncBagLoader
    P = R = 12
    pr = synth.make_params(P, R, 4)
    net = build_net(pr, P, R, dev).eval()
    steps = []
    for s in range(3):
        sizes = [1000 + 37 * s, 1, 5000, 64 + s]
        steps.append([synth.make_bag("g1", n, 300 + 10 * s + i) for i, n in enumerate(sizes)])
    Q = net.mil_encoder.get_query().detach().contiguous()
    W, b = net.mil_encoder.visual_adapter.weight.detach(), net.mil_encoder.visual_adapter.bias.detach()
    T, ls = net.forward_text_only(), net.logit_scale.detach()
    direct = []
    for bags in steps:
        X = torch.cat(bags, 0).to(dev)
        plan = ops.make_plan([x.shape[0] for x in bags], dev)
        direct.append(ops.aggregate_forward_raw(X, plan, Q, W, b, T, ls)["incidence"].cpu())
    # (a) C-ABI host entry; the workspace of one call is handed to the next WITHOUT a synchronisation in between (the
    #     library orders the copy stream behind the kernels still reading it), several rounds to give a race a chance
    copy_stream = torch.cuda.Stream()
    hosts = [torch.cat(bags, 0).pin_memory() for bags in steps]
    for _ in range(5):
        ws, outs = None, []
        for bags, host in zip(steps, hosts):
            out, ws = ops.forward_host(host, [x.shape[0] for x in bags], Q, W, b, T, ls, workspace=ws, copy_stream=copy_stream)
            outs.append(out)
        torch.cuda.synchronize()
        for out, ref in zip(outs, direct):
            assert torch.equal(out, ref)
    # (a') pre-normalised query rows (the gated query's difference rows) through the same entry
    Qd = torch.nn.functional.normalize(Q, dim=-1)
    Qd = (Qd[:-1] - Qd[-1:]).contiguous()
    Xd = torch.cat(steps[0], 0).to(dev)
    ref_g = ops.aggregate_forward_raw(Xd, ops.make_plan([x.shape[0] for x in steps[0]], dev), Qd, W, b, T, ls,
                                      q_prenorm=True)["incidence"].cpu()
    out_g, _ = ops.forward_host(hosts[0], [x.shape[0] for x in steps[0]], Qd, W, b, T, ls, copy_stream=copy_stream, q_prenorm=True)
    torch.cuda.synchronize()
    assert torch.equal(out_g, ref_g) and not torch.equal(out_g[:, :], direct[0][:, :])
    # (b) loader + module API; a batch is released automatically when the next one is requested
    loader = AsyncBagLoader(((bags, None, None) for bags in steps), dev, depth=2)
    got = []
    with torch.no_grad():
        for batch in loader:
            batch.wait()
            logits, g, Tn, inc = net.forward_packed(batch.X, batch.plan)
            got.append(inc)
    got = [g_.cpu() for g_ in got]
    assert len(got) == 3
    for a, ref in zip(got, direct):
        assert torch.equal(a, ref)
    assert loader.h2d_bytes == sum(x.shape[0] for bags in steps for x in bags) * 512 * 4


def test_loss_modules_keep_reference_signature(dev):
    """SurvIFMLE / SurvEMD take the converted incidence (loss_surv.py:144, loss_surv_ext.py:70)."""
    from oracle import vlsa_oracle as O
    from vlsa_b200.loss import SurvEMD, SurvIFMLE, load_loss
    case = load_case("loss_R12")
    raw = torch.from_numpy(case["raw"])
    t, e = torch.from_numpy(case["t"]), torch.from_numpy(case["e"])
    p = torch.softmax(raw.to(dev), -1).requires_grad_(True)
    ls = torch.tensor(float(case["logit_scale"]), device=dev)
    fns = load_loss("vlsa", loss_type=["SurvIFMLE", "SurvEMD"], SurvIFMLE={}, SurvEMD={"p": 2})
    l1 = fns["SurvIFMLE"](p, t.view(-1, 1).float().to(dev), e.view(-1, 1).float().to(dev))
    l2 = fns["SurvEMD"](p, t.view(-1, 1).float().to(dev), e.view(-1, 1).float().to(dev), ls.exp())
    (l1 + l2).backward()
    p64 = torch.softmax(raw.double(), -1).requires_grad_(True)
    r1 = O.surv_ifmle(p64, t, e)
    r2 = O.surv_emd(p64, t, e, torch.tensor(float(case["logit_scale"]), dtype=torch.float64).exp())
    (r1 + r2).backward()
    np.testing.assert_allclose(l1.item(), r1.item(), rtol=1e-5)
    np.testing.assert_allclose(l2.item(), r2.item(), rtol=1e-5)
    ref = p64.grad.numpy()
    assert np.abs(p.grad.cpu().numpy() - ref).max() <= 2e-4 * np.abs(ref).max()
    assert isinstance(SurvIFMLE(reduction="sum"), torch.nn.Module) and isinstance(SurvEMD(), torch.nn.Module)
    with pytest.raises(NotImplementedError):
        SurvEMD(p=1)


@pytest.mark.parametrize("P,R,N,kind,axis", [(12, 12, 2798, "g1", "V"), (4, 4, 1000, "g0", "L"), (7, 13, 5000, "g1", "V")])
def test_interpretation_path_matches_reference_formula(P, R, N, kind, axis, dev):
    """utils/model_inference.py:81-144 (`calc_text_img_similarity`): attention, decoupled similarities and their
    softmaxes from ONE streaming pass vs the reference's formula evaluated the long way (visual_adapter over all N
    patches) in fp64."""
    import numpy as np
    from oracle import vlsa_oracle as O
    from vlsa_b200 import synth
    from vlsa_b200.utils import calc_text_img_similarity
    pr = synth.make_params(P, R, 40 + P)
    net = build_net(pr, P, R, dev)
    X = synth.make_bag(kind, N, 77 + N)
    _, A, cottn, probs, probs2, imp, shap = calc_text_img_similarity(net, X.unsqueeze(0), axis_softmax=axis)
    c = lambda z: z.double()
    Q64 = O.task_res_query(c(pr["prompt_features"]), c(pr["residual_features"]), pr["res_ratio"])
    A64, probs64, probs2_64, dec64 = O.decoupled_similarity(c(X).unsqueeze(0), Q64, c(pr["W"]), c(pr["b"]),
                                                            c(pr["text_features"]), c(pr["logit_scale"]))
    np.testing.assert_allclose(cottn.numpy(), A64.numpy(), rtol=2e-4, atol=1e-9)
    if axis == "V":
        np.testing.assert_allclose(A.numpy(), A64.numpy(), rtol=2e-4, atol=1e-9)
    else:
        Qn = Q64 / Q64.norm(dim=-1, keepdim=True); Xn = c(X) / c(X).norm(dim=-1, keepdim=True)
        AL = torch.softmax(O.coattn_scale() * Qn @ Xn.T, dim=0)
        np.testing.assert_allclose(A.numpy(), AL.numpy(), rtol=2e-4, atol=1e-7)
    assert np.abs(probs.numpy() - probs64.numpy()).max() <= 2e-5
    assert np.abs(probs2.numpy() - probs2_64.numpy()).max() <= 2e-5
    ls = float(pr["logit_scale"].exp())
    imp64 = torch.softmax(ls * dec64, dim=0)
    assert np.abs(imp.numpy() - imp64.numpy()).max() <= 5e-5
    shap64 = O.prototype_shap_imp(dec64.float(), ls) if P <= 9 else None
    if shap64 is not None:
        np.testing.assert_allclose(shap.numpy(), shap64.numpy(), atol=2e-4)
    assert shap.shape == (P,)


@pytest.mark.parametrize("name", golden_cases("interp_"))
def test_interpretation_path_vs_reference_own_routine(name, dev):
    """`calc_text_img_similarity` + `evaluate_prototype_shap_imp` against the outputs of the reference's OWN functions
    (utils/model_inference.py:21-144) run unmodified on CPU in fp64 (tests/golden/make_golden_r02.py)."""
    from golden_util import rebuild_interp
    from vlsa_b200.utils import calc_text_img_similarity
    case = load_case(name)
    X, pr = rebuild_interp(case)
    P, R, axis = int(case["P"]), int(case["R"]), str(case["axis"])
    net = build_net(pr, P, R, dev)
    none, A, cottn, probs, probs2, imp, shap = calc_text_img_similarity(net, X.unsqueeze(0), axis_softmax=axis)
    assert none is None and A.shape == (P, X.shape[0]) and cottn.shape == (P, X.shape[0])
    k = case["A_head_f64"].shape[1]
    np.testing.assert_allclose(A[:, :k].numpy(), case["A_head_f64"], rtol=2e-4, atol=1e-7 if axis == "L" else 1e-9)
    np.testing.assert_allclose(cottn[:, :k].numpy(), case["cottn_head_f64"], rtol=2e-4, atol=1e-9)
    np.testing.assert_allclose(cottn.max(1).values.numpy(), case["cottn_max_f64"], rtol=2e-4)
    assert (cottn.argmax(1).numpy() == case["cottn_argmax_f64"]).all()
    np.testing.assert_allclose(A.sum(1).numpy(), case["A_rowsum_f64"], rtol=1e-4)
    assert np.abs(probs.numpy() - case["probs_f64"]).max() <= IF_TOL
    assert np.abs(probs2.numpy() - case["probs2_f64"]).max() <= IF_TOL
    assert np.abs(imp.numpy() - case["imp_f64"]).max() <= 5e-5
    np.testing.assert_allclose(shap.numpy(), case["shap_f64"], atol=2e-4)        # all 2^P subsets, P up to 12
    np.testing.assert_allclose(shap.numpy(), case["shap_f32"], atol=5e-4)


def test_flat_store_to_async_loader_to_forward(tmp_path, dev):
    """f2 of SURVEY §8: flat memory-mapped store -> pinned packed steps -> AsyncBagLoader (copy stream, ring) ->
    forward_packed gives what the reference-style per-bag loop gives on the same rows."""
    import numpy as np
    from vlsa_b200 import synth
    from vlsa_b200.dataset import AsyncBagLoader, PatchFeatureStore, WSIPatchSurvStore, build_store
    P = R = 12
    pr = synth.make_params(P, R, 5)
    net = build_net(pr, P, R, dev)
    slides = {f"s{i}": synth.make_bag("g1", n, 300 + i) for i, n in enumerate([700, 33, 1500, 260, 1, 999, 4100])}
    build_store(str(tmp_path / "st"), slides.items())
    st = PatchFeatureStore(str(tmp_path / "st"))
    pid2sids = {"a": ["s0", "s1"], "b": ["s2"], "c": ["s3", "s4", "s5"], "d": ["s6"], "e": ["s1"]}
    pid2label = {k: (float(i % R), float(i % 2)) for i, k in enumerate(pid2sids)}
    ds = WSIPatchSurvStore(st, list(pid2sids), pid2sids, pid2label)
    got, labs = [], []
    loader = AsyncBagLoader(ds.steps(batch_size=2), dev, depth=2)
    with torch.no_grad():
        for batch in loader:
            batch.wait()
            logits, g, Tn, inc = net.forward_packed(batch.X, batch.plan)
            loader.release(batch)
            got.append(inc.cpu()); labs.append(batch.labels.cpu())
    got = torch.cat(got, 0).numpy()
    assert torch.cat(labs, 0).tolist() == [list(pid2label[k]) for k in pid2sids]
    with torch.no_grad():
        for i, pid in enumerate(pid2sids):
            X = torch.cat([slides[s] for s in pid2sids[pid]], 0).unsqueeze(0).to(dev)
            ref = torch.softmax(net(X)[0], -1).cpu().numpy()
            assert np.abs(got[i] - ref[0]).max() <= 2e-6


def test_device_cohort_row_range_plans_match_packed_batches(dev):
    """Steps drawn from a device-resident cohort (row-range plans, VLSA_ROWS_RANGES) against the same bags packed:
    forward bit-identical for both streaming kernels, gradients equal, loader-filled cohort included."""
    from vlsa_b200 import ops, synth
    from vlsa_b200.dataset import AsyncBagLoader, DeviceCohort
    from vlsa_b200.runner import VLSAHandler
    sizes = [1000, 37, 2798, 1, 513, 64, 4096, 255]
    bags = [synth.make_bag("g1", n, 700 + i) for i, n in enumerate(sizes)]
    for P in (4, 12):
        pr = synth.make_params(P, P, 60 + P)
        net = build_net(pr, P, P, dev).eval()
        # cohort filled by the loader in two steps (bags 0-3, 4-7), each copied straight into its final place
        cohort = DeviceCohort(dev, sum(sizes) + 100)
        steps = [(bags[:4], None, torch.arange(0, 4)), (bags[4:], None, torch.arange(4, 8))]
        loader = AsyncBagLoader(iter(steps), dev, depth=2, cohort=cohort)
        with torch.no_grad():
            first = []
            for batch in loader:
                batch.wait()
                first.append(net.forward_packed(batch.X, batch.plan)[3])
        assert len(cohort) == 8 and cohort.rows == sum(sizes)
        # a shuffled step with a repeated bag, drawn from the cohort vs packed by hand
        order = [6, 1, 3, 0, 6, 7]
        with torch.no_grad():
            inc_c = net.forward_packed(cohort.X, cohort.plan(order))[3]
            Xp = torch.cat([bags[i] for i in order], 0).to(dev)
            inc_p = net.forward_packed(Xp, ops.make_plan([sizes[i] for i in order], dev))[3]
        assert torch.equal(inc_c, inc_p)
        assert torch.equal(first[0], net.forward_packed(torch.cat(bags[:4], 0).to(dev), ops.make_plan(sizes[:4], dev))[3])
        # one optimizer step from the cohort == one optimizer step on the packed bags (same initial weights)
        cfg = dict(task="vlsa", arch="VLSA", loss_type="SurvIFMLE-SurvEMD", opt_name="adam", opt_lr=2e-4)
        t, e = synth.make_labels(len(order), P, 5)
        ys = [torch.stack([t[i], e[i]]).float().reshape(1, 2) for i in range(len(order))]
        net_a, net_b = build_net(pr, P, P, dev), build_net(pr, P, P, dev)
        for n_ in (net_a, net_b):
            n_.pretrained_text_features.requires_grad_(False)
        ha, hb = VLSAHandler(cfg, net_a, device=dev), VLSAHandler(cfg, net_b, device=dev)
        la, pa = ha.update_network_cached(cohort, order, ys)
        lb, pb = hb._update_network([bags[i].unsqueeze(0).to(dev) for i in order], ys)
        assert la == lb and torch.equal(pa, pb)
        for (k, va), (_, vb) in zip(net_a.state_dict().items(), net_b.state_dict().items()):
            assert torch.equal(va, vb), k
    with pytest.raises(KeyError):
        cohort.plan([99])
    with pytest.raises(MemoryError):
        cohort.reserve("x", 10 ** 9)


@pytest.mark.gpu
@pytest.mark.parametrize("P", [7, 12, 16, 4])
def test_split16_cohort_is_bit_identical_to_the_fp32_tensor_core_kernel(P, dev):
    """A device cohort stored as pre-split tile images (layout='split16', vlsa_split16_pack + agg_split_kernel) against the
    same bags as fp32 rows through the register-staged tensor-core kernel: same planes, same summation orders — the FORWARD
    (incidence, logits, per-prototype softmax statistics) must be bit-identical, for shuffled steps with a repeated bag, bags
    of 1 row, of tiny and of huge rows (the power-of-two row scale).  The backward takes u = dv . x / P from the (hi + lo)
    planes instead of the raw row (another summation order): gradients within 2e-5 of their largest entry, and an optimizer
    step leaves the same weights to 1e-6."""
    from vlsa_b200 import ops, synth
    from vlsa_b200.dataset import DeviceCohort
    from vlsa_b200.runner import VLSAHandler
    sizes = [1000, 37, 2798, 1, 513, 64, 4096, 255, 16, 17]
    bags = [synth.make_bag("g1", n, 300 + i) for i, n in enumerate(sizes)]
    bags[2] = bags[2] * 1e-3
    bags[5] = bags[5] * 3e3
    pr = synth.make_params(P, P, 40 + P)
    cohort = DeviceCohort(dev, sum((n + 15) // 16 * 16 for n in sizes), layout="split16")
    for i, b in enumerate(bags):
        cohort.add(i, b)
    assert len(cohort) == len(sizes) and cohort.X.shape[1] == ops.SPLIT16_COLS
    order = [6, 1, 3, 0, 6, 7, 2, 5, 9, 8, 4]
    t, e = synth.make_labels(len(order), P, 5)
    res = {}
    try:
        ops.set_agg_variant("tc")                       # fp32 rows on the tensor-core kernel whatever P
        for name in ("rows", "split16"):
            leaf = lambda z: z.detach().clone().to(dev).requires_grad_(True)
            r, W, b, T, ls = (leaf(pr[k]) for k in ("residual_features", "W", "b", "text_features", "logit_scale"))
            Q = pr["res_ratio"] * r + pr["prompt_features"].to(dev)
            if name == "rows":
                X, plan = torch.cat([bags[i] for i in order], 0).to(dev), ops.make_plan([sizes[i] for i in order], dev)
            else:
                X, plan = cohort.X, cohort.plan(order)
            logits, g, Tn, inc, ml = ops.aggregate(X, plan, Q, W, b, T, ls)
            total, *_ = ops.surv_loss(logits, t.to(dev), e.to(dev), ls)
            total.backward()
            torch.cuda.synchronize()
            res[name] = dict(inc=inc.detach(), logits=logits.detach(), ml=ml.detach(), d_res=r.grad, d_W=W.grad, d_b=b.grad,
                             d_T=T.grad, d_ls=ls.grad)
    finally:
        ops.set_agg_variant(None)
    for k in ("inc", "logits", "ml"):
        assert torch.equal(res["rows"][k], res["split16"][k]), k
    for k in ("d_res", "d_W", "d_b", "d_T", "d_ls"):
        a, b_ = res["rows"][k], res["split16"][k]
        assert float((a - b_).abs().max()) <= 2e-5 * max(float(a.abs().max()), 1e-30), k
    # the handler on the cohort: default dispatch (P > 5: the same tensor-core arithmetic -> identical weights)
    if P > 5:
        cfg = dict(task="vlsa", arch="VLSA", loss_type="SurvIFMLE-SurvEMD", opt_name="adam", opt_lr=2e-4)
        ys = [torch.stack([t[i], e[i]]).float().reshape(1, 2) for i in range(len(order))]
        net_a, net_b = build_net(pr, P, P, dev), build_net(pr, P, P, dev)
        for n_ in (net_a, net_b):
            n_.pretrained_text_features.requires_grad_(False)
        ha, hb = VLSAHandler(cfg, net_a, device=dev), VLSAHandler(cfg, net_b, device=dev)
        la, pa = ha.update_network_cached(cohort, order, ys)
        lb, pb = hb._update_network([bags[i].unsqueeze(0).to(dev) for i in order], ys)
        assert la == lb and torch.equal(pa, pb)
        for (k, va), (_, vb) in zip(net_a.state_dict().items(), net_b.state_dict().items()):
            assert torch.allclose(va, vb, rtol=0, atol=1e-6), k
    with pytest.raises(RuntimeError):
        cohort.reserve("x", 10)
    with pytest.raises(ValueError):
        ops.aggregate(cohort.X, ops.make_plan([16], dev), Q.detach(), W.detach(), b.detach(), T.detach(), ls.detach())


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["rows", "split16"])
def test_test_model_cached_matches_test_model(layout, dev):
    """`VLSAHandler.test_model_cached` (bags drawn from a DeviceCohort by row-range plans) against `test_model` on a loader of
    the same bags: identical raw predictions and incidence (P = 12: both run the tensor-core arithmetic)."""
    from vlsa_b200 import synth
    from vlsa_b200.dataset import DeviceCohort
    from vlsa_b200.runner import VLSAHandler
    P = 12
    sizes = [700, 33, 1, 2049, 512, 90, 4000]
    bags = [synth.make_bag("g1", n, 800 + i) for i, n in enumerate(sizes)]
    pr = synth.make_params(P, P, 13)
    net = build_net(pr, P, P, dev)
    h = VLSAHandler(dict(task="vlsa", arch="VLSA", loss_type="SurvIFMLE-SurvEMD", opt_name="adam", opt_lr=2e-4), net, device=dev)
    t, e = synth.make_labels(len(sizes), P, 3)
    ys = [torch.stack([t[i], e[i]]).float().reshape(1, 2) for i in range(len(sizes))]
    loader = [(torch.tensor([[i]]), (bags[i].unsqueeze(0), torch.zeros(1)), ys[i]) for i in range(len(sizes))]
    ref = h.test_model(h.net, loader, "test", bags_per_launch=3)["pred"]
    cohort = DeviceCohort(dev, sum((n + 15) // 16 * 16 for n in sizes), layout=layout)
    for i, b in enumerate(bags):
        cohort.add(i, b)
    got = h.test_model_cached(h.net, cohort, list(range(len(sizes))), ys, bags_per_launch=3)["pred"]
    for k in ("raw_y_hat", "y_hat", "y"):
        assert torch.equal(ref[k], got[k]), k
    assert got["uid"].tolist() == list(range(len(sizes)))


@pytest.mark.gpu
@pytest.mark.parametrize("P,gated", [(4, False), (12, False), (7, True)])
def test_fused_train_step_equals_the_autograd_step(P, gated, dev):
    """`ops.FusedTrainStep` (three C calls on persistent buffers, gradients written into the bucket, losses into its tail, no
    host synchronisation) against the autograd path of the same handler: the SAME kernels run in the same order, so losses,
    predictions and the weights after two optimizer steps must be bit-identical — for fp32 rows on both dispatches (P = 4:
    CUDA cores, P = 12: tcgen05), for the gated query (difference rows, no normalisation), from a split16 cohort, with a
    trainable text tensor, and through `step_packed` / `sync=False`."""
    from vlsa_b200 import ops, synth
    from vlsa_b200.dataset import DeviceCohort
    from vlsa_b200.model import VLSA
    from vlsa_b200.runner import VLSAHandler
    sizes = [1000, 37, 2798, 1, 513, 64, 4096, 255]
    bags = [synth.make_bag("g1", n, 300 + i) for i, n in enumerate(sizes)]
    pr = synth.make_params(P, P, 20 + P)
    cfg = dict(task="vlsa", arch="VLSA", loss_type="SurvIFMLE-SurvEMD", opt_name="adam", opt_lr=2e-4,
               loss_survifmle_weight=0.7, loss_survemd_weight=1.3)
    t, e = synth.make_labels(len(sizes), P, 5)
    ys = [torch.stack([t[i], e[i]]).float().reshape(1, 2) for i in range(len(sizes))]
    xs = [b.unsqueeze(0).to(dev) for b in bags]

    def make(fused, train_text):
        torch.manual_seed(1234)                      # the gate's residual row is drawn at construction
        if gated:
            img = dict(name="VLFAN", dim_in=512, use_feat_proj=False, query="Text", num_query=P, gated_query=True,
                       query_text_method="TaskRes")
            g = torch.Generator().manual_seed(9)
            net = VLSA({"name": "mahmoodlab/conch"}, img, {"name": "CoOp"}, text_features=pr["text_features"],
                       query_prompt_features=pr["prompt_features"], query_neg_prompt_features=torch.randn(1, 512, generator=g),
                       logit_scale_init=float(pr["logit_scale"])).to(dev)
        else:
            net = build_net(pr, P, P, dev)
        if train_text:
            net = net.to(dev)
            del net.pretrained_text_features
            net._text_param = torch.nn.Parameter(pr["text_features"].clone().to(dev))
            net._text_fn = lambda: net._text_param
        h = VLSAHandler(dict(cfg, vlsa_fused_step=fused), net, device=dev)
        assert h._fused_ok() == fused
        return h

    for train_text in (False, True):
        ha, hb = make(True, train_text), make(False, train_text)
        ha.net.eval()
        with torch.no_grad():                        # an evaluation BEFORE the steps fills the query-row cache
            ha.net.forward_packed(torch.cat(bags, 0).to(dev), ops.make_plan(sizes, dev))
        ha.net.train()
        for step in range(2):
            la, pa = ha._update_network(xs, ys)
            lb, pb = hb._update_network(xs, ys)
            assert la == lb and torch.equal(pa, pb), (step, la, lb)
        for (k, va), (_, vb) in zip(ha.net.state_dict().items(), hb.net.state_dict().items()):
            assert torch.equal(va, vb), k
        if train_text:
            assert torch.equal(ha.net._text_param, hb.net._text_param)
            assert not torch.equal(ha.net._text_param, pr["text_features"].to(dev))
        # the optimizer kernel writes through raw pointers: version counters must still move (caches keyed on them, e.g. the
        # query rows of the lean inference call, would otherwise serve the weights of before the step)
        ha.net.eval()
        with torch.no_grad():
            eval_a = ha.net.forward_packed(torch.cat(bags, 0).to(dev), ops.make_plan(sizes, dev))[0]
            eval_b = ops.aggregate(torch.cat(bags, 0).to(dev), ops.make_plan(sizes, dev), *ha.net.mil_encoder.query_directions()[:1],
                                   ha.net.mil_encoder.visual_adapter.weight, ha.net.mil_encoder.visual_adapter.bias,
                                   ha.net._text_features_for_kernels(), ha.net.logit_scale,
                                   q_prenorm=ha.net.mil_encoder.query_directions()[1])[0]
        assert torch.equal(eval_a, eval_b)
        ha.net.train()
        # no-sync entry points: device tensors back, same numbers; component losses ride the bucket tail
        X = torch.cat(bags, 0).to(dev)
        plan = ops.make_plan(sizes, dev)
        lab = torch.cat(ys, 0)
        l1, p1 = ha.step_packed(X, plan, lab)
        l2, p2 = hb._update_network(xs, ys, sync=False)
        assert l1.is_cuda and p1.is_cuda and float(l1) == float(l2) and torch.equal(p1, p2)
        tail = ha.bucket.tail.cpu()
        assert abs(float(tail[0]) - (0.7 * float(tail[1]) + 1.3 * float(tail[2]))) <= 1e-5 * abs(float(tail[0]))
    # from a split16 cohort
    if not gated:
        cohort = DeviceCohort(dev, sum((n + 15) // 16 * 16 for n in sizes), layout="split16")
        for i, b in enumerate(bags):
            cohort.add(i, b)
        ha, hb = make(True, False), make(False, False)
        order = [6, 1, 3, 0, 7, 2]
        la, pa = ha.update_network_cached(cohort, order, [ys[i] for i in order])
        lb, pb = hb.update_network_cached(cohort, order, [ys[i] for i in order])
        assert la == lb and torch.equal(pa, pb)
        for (k, va), (_, vb) in zip(ha.net.state_dict().items(), hb.net.state_dict().items()):
            assert torch.equal(va, vb), k


@pytest.mark.gpu
def test_bucket_adam_follows_torch_adam(dev):
    """`BucketAdam` (one launch over the flat gradient bucket, untouched parameters skipped on the device) against
    torch.optim.Adam on the same gradients: decay / no-decay groups, a 0-dim parameter, sizes that are no multiple of four, a
    parameter without a gradient in some steps (torch: grad None -> no update, no step count), and checkpoints that load
    into each other."""
    from vlsa_b200.runner.dist import FlatBucket
    from vlsa_b200.runner.optim import BucketAdam
    torch.manual_seed(3)
    shapes = [(), (513,), (37, 19), (512, 512), (5,)]

    def make():
        torch.manual_seed(11)
        return [torch.nn.Parameter(torch.randn(s, device=dev)) for s in shapes]
    pa, pb = make(), make()
    groups = lambda ps: [{"params": [ps[1], ps[4]], "weight_decay": 0.0}, {"params": [ps[0], ps[2], ps[3]], "weight_decay": 1e-2}]
    ref = torch.optim.Adam(groups(pa), lr=3e-3)
    bk = FlatBucket(pb, extra=3, align=4)
    bk.attach()
    opt = BucketAdam(groups(pb), bk, lr=3e-3)

    def one_step(step):
        bk.zero()
        for i, (a, b) in enumerate(zip(pa, pb)):
            if i == 2 and step % 2 == 1:            # no gradient for this tensor on odd steps
                a.grad = None
                continue
            g = torch.randn_like(a) * (0.1 + step)
            a.grad = g.clone()
            b.grad.copy_(g)
            bk.mark_touched(b)
        bk.pack(None)
        bk.all_reduce()
        opt.step()
        ref.step()

    for step in range(6):
        one_step(step)
    for a, b in zip(pa, pb):
        torch.testing.assert_close(b, a, rtol=2e-6, atol=2e-7)
    sd = opt.state_dict()
    assert float(sd["state"][2]["step"]) == 6 and float(sd["state"][3]["step"]) == 3      # torch numbering: group by group
    assert float(ref.state_dict()["state"][3]["step"]) == 3
    # checkpoints interchange: each continues from the other's state
    ref2 = torch.optim.Adam(groups(pa), lr=3e-3)
    ref2.load_state_dict(sd)
    opt.load_state_dict(ref.state_dict())
    ref = ref2
    for step in range(6, 9):
        one_step(step)
    for a, b in zip(pa, pb):
        torch.testing.assert_close(b, a, rtol=2e-6, atol=2e-7)


@pytest.mark.gpu
def test_train_each_epoch_on_the_device(dev):
    """`VLSAHandler._train_each_epoch` with the real kernels: groups of `bp_every_batch` bags, one fused step each with no host
    synchronisation inside the epoch, one read-back at its end — same losses, predictions and weights as a twin handler stepping
    the same groups with synchronous `_update_network` calls."""
    from vlsa_b200 import synth
    from vlsa_b200.runner import VLSAHandler
    P = 12
    sizes = [1000, 37, 2798, 1, 513, 64, 4096, 255]
    bags = [synth.make_bag("g1", n, 900 + i) for i, n in enumerate(sizes)]          # host bags, as a DataLoader yields them
    pr = synth.make_params(P, P, 33)
    t, e = synth.make_labels(len(sizes), P, 5)
    ys = [torch.stack([t[i], e[i]]).float().reshape(1, 2) for i in range(len(sizes))]
    cfg = dict(task="vlsa", arch="VLSA", loss_type="SurvIFMLE-SurvEMD", opt_name="adam", opt_lr=2e-4, bp_every_batch=3)
    ha, hb = VLSAHandler(cfg, build_net(pr, P, P, dev), device=dev), VLSAHandler(cfg, build_net(pr, P, P, dev), device=dev)
    loader = [(torch.tensor([[10 + i]]), (bags[i].unsqueeze(0), torch.zeros(1)), ys[i]) for i in range(len(bags))]
    out = ha._train_each_epoch(0, loader)
    ref_loss, ref_pred = [], []
    for g in ([0, 1, 2], [3, 4, 5], [6, 7]):
        l, p = hb._update_network([bags[i].unsqueeze(0) for i in g], [ys[i] for i in g])
        ref_loss.append(l); ref_pred.append(p)
    assert len(out["loss"]) == 3 and all(isinstance(v, float) for v in out["loss"]) and out["loss"] == ref_loss
    assert not out["pred"]["raw_y_hat"].is_cuda and torch.equal(out["pred"]["raw_y_hat"], torch.cat(ref_pred, 0))
    assert out["pred"]["uid"].tolist() == [10 + i for i in range(len(bags))]
    for (k, va), (_, vb) in zip(ha.net.state_dict().items(), hb.net.state_dict().items()):
        assert torch.equal(va, vb), k


@pytest.mark.gpu
def test_query_div_regulariser_rides_the_fused_step(dev):
    """loss_type 'SurvIFMLE-SurvEMD-QueryDiv' (runner/vlsa_handler.py:181-187,255-256: weight * VLFAN.query_div_loss() added once per
    optimizer step): the loss of a step is the survival objective plus the weighted regulariser, and the fused step (kernels write
    d residual, autograd adds the regulariser's share on top) leaves the same weights as the autograd path."""
    from vlsa_b200 import ops, synth
    from vlsa_b200.runner import VLSAHandler
    P = 12
    sizes = [1000, 37, 2798, 513]
    bags = [synth.make_bag("g1", n, 70 + i).to(dev) for i, n in enumerate(sizes)]
    pr = synth.make_params(P, P, 44)
    t, e = synth.make_labels(len(sizes), P, 5)
    ys = [torch.stack([t[i], e[i]]).float().reshape(1, 2) for i in range(len(sizes))]
    xs = [b.unsqueeze(0) for b in bags]
    base = dict(task="vlsa", arch="VLSA", opt_name="adam", opt_lr=2e-4)
    h0 = VLSAHandler(dict(base, loss_type="SurvIFMLE-SurvEMD"), build_net(pr, P, P, dev), device=dev)
    ha = VLSAHandler(dict(base, loss_type="SurvIFMLE-SurvEMD-QueryDiv", loss_querydiv_weight=0.5), build_net(pr, P, P, dev), device=dev)
    hb = VLSAHandler(dict(base, loss_type="SurvIFMLE-SurvEMD-QueryDiv", loss_querydiv_weight=0.5, vlsa_fused_step=False),
                     build_net(pr, P, P, dev), device=dev)
    assert ha._fused_ok() and not hb._fused_ok()
    with torch.no_grad():
        reg = float(ha.net.mil_encoder.query_div_loss())
    l0, _ = h0._update_network(xs, ys)
    la, pa = ha._update_network(xs, ys)
    lb, pb = hb._update_network(xs, ys)
    assert abs(la - (l0 + 0.5 * reg)) <= 1e-5 * abs(la) and reg > 0
    assert la == lb and torch.equal(pa, pb)
    for (k, va), (_, vb) in zip(ha.net.state_dict().items(), hb.net.state_dict().items()):
        assert torch.equal(va, vb), k
    assert not torch.equal(ha.net.mil_encoder.Q.residual_features, h0.net.mil_encoder.Q.residual_features)


@pytest.mark.gpu
@pytest.mark.parametrize("P,dtype", [(4, torch.float32), (12, torch.float32), (12, torch.bfloat16)])
def test_graphed_forward_replays_the_same_call(P, dtype, dev):
    """`VLSA.graphed(N)` (CUDA-graph replay of `VLSA.forward` for bags of one size): bit-identical to the eager call for every
    bag, sees in-place weight updates, re-captures when the prompt adapter's tensors change, rejects other shapes."""
    from vlsa_b200 import synth
    N = 2798
    pr = synth.make_params(P, P, 50 + P)
    net = build_net(pr, P, P, dev).eval()
    gf = net.graphed(N, dtype)
    bags = [synth.make_bag("g1", N, 400 + i).to(dev).to(dtype).unsqueeze(0) for i in range(3)]
    with torch.no_grad():
        for X in bags:
            want = [z.clone() for z in net(X)]
            got = gf(X)
            assert all(torch.equal(a, b) for a, b in zip(want, got))
        # the caller may land a bag in `.input` itself (e.g. as the target of its H2D copy) and replay
        gf.input.copy_(bags[1])
        assert torch.equal(gf.replay()[0], net(bags[1])[0])
        # in-place weight updates are seen through the pointers; a changed residual re-captures (new query rows)
        net.mil_encoder.visual_adapter.weight.mul_(1.01)
        net.logit_scale.add_(0.05)
        assert torch.equal(gf(bags[0])[0], net(bags[0])[0])
        net.mil_encoder.Q.residual_features.add_(0.01)
        assert torch.equal(gf(bags[2])[0], net(bags[2])[0])
        with pytest.raises(ValueError):
            gf(bags[0][:, :100])
    net.train()
    with pytest.raises(RuntimeError):
        gf(bags[0])


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["rows", "split16"])
def test_train_each_epoch_with_an_automatic_device_cohort(layout, dev):
    """cfg `vlsa_device_cohort`: `_train_each_epoch` uploads every bag once, steps from the resident buffer by row-range plans, and
    tells a cooperating dataset to stop reading what is resident (`skip_features`): the second epoch sees empty feature tensors
    and trains all the same.  'rows' must match the plain handler bit for bit; 'split16' (pre-split tensor-core records, another
    summation order in the backward) to 1e-4."""
    from vlsa_b200 import synth
    from vlsa_b200.runner import VLSAHandler
    P = 12
    sizes = [1000, 37, 2798, 1, 513, 64, 4096, 255]
    bags = [synth.make_bag("g1", n, 600 + i) for i, n in enumerate(sizes)]
    pr = synth.make_params(P, P, 21)
    t, e = synth.make_labels(len(sizes), P, 5)
    ys = [torch.stack([t[i], e[i]]).float().reshape(1, 2) for i in range(len(sizes))]

    class DS:
        def __init__(self): self.skip, self.reads = set(), 0
        def skip_features(self, idx): self.skip = set(int(i) for i in idx)
        def item(self, i):
            if i in self.skip:
                return torch.tensor([[i]]), (torch.empty(1, 0, 512),), ys[i]
            self.reads += 1
            return torch.tensor([[i]]), (bags[i].unsqueeze(0),), ys[i]

    class Loader:
        def __init__(self, ds): self.dataset = ds
        def __len__(self): return len(bags)
        def __iter__(self): return (self.dataset.item(i) for i in range(len(bags)))

    cfg = dict(task="vlsa", arch="VLSA", loss_type="SurvIFMLE-SurvEMD", opt_name="adam", opt_lr=2e-4, bp_every_batch=3)
    ha = VLSAHandler(dict(cfg, vlsa_device_cohort=layout, vlsa_device_cohort_rows=2000), build_net(pr, P, P, dev), device=dev)
    hb = VLSAHandler(cfg, build_net(pr, P, P, dev), device=dev)
    la, lb = Loader(DS()), Loader(DS())
    for epoch in range(3):
        oa, ob = ha._train_each_epoch(epoch, la), hb._train_each_epoch(epoch, lb)
        if layout == "rows":
            assert oa["loss"] == ob["loss"] and torch.equal(oa["pred"]["raw_y_hat"], ob["pred"]["raw_y_hat"])
        else:
            np.testing.assert_allclose(oa["loss"], ob["loss"], rtol=1e-4)
            torch.testing.assert_close(oa["pred"]["raw_y_hat"], ob["pred"]["raw_y_hat"], rtol=1e-4, atol=1e-4)
    assert la.dataset.reads == len(bags) and lb.dataset.reads == 3 * len(bags)        # the cohort's dataset read every bag once
    assert len(ha.cohort) == len(bags) and ha.cohort.X.shape[0] >= sum(sizes)           # grown from the 2 000 rows it started with
