"""CPU: host-side mirror of the reference interface — config parsing, module structure / state-dict keys,
sharding, gradient bucket, and the "fails loudly without CUDA" contract."""
import os

import numpy as np
import pytest
import torch

from vlsa_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _net(P=12, R=12, seed=3):
    from vlsa_b200.model import VLSA
    pr = synth.make_params(P, R, seed)
    net = VLSA({"name": "mahmoodlab/conch"},
               dict(name="VLFAN", dim_in=512, dim_hid=256, use_feat_proj=False, drop_rate=0.25, query="Text",
                    num_query=P, gated_query=False, query_pooling="mean", pred_head="default", dim_reduction=4,
                    keep_ratio=0.8, query_text_method="TaskRes", query_text_res_ratio=0.5,
                    query_text_load_path="tools/survival_text_prototypes.json", query_text_load_idx="tcga_blca_0"),
               {"name": "CoOp"}, text_features=pr["text_features"], query_prompt_features=pr["prompt_features"],
               vlsa_api="CONCH", path_clip_model=None)
    return net, pr


def test_reference_config_parses_unchanged():
    from vlsa_b200.runner import config as C
    from vlsa_b200.runner import fetch_kws
    cfg = C.load_config(os.path.join(GOLDEN, "cfg_vlsa_conch.yaml"))
    grid = C.expand_grid(cfg)
    assert len(grid) == 5                                   # 5 folds (data_split_seed is the only real axis)
    run = C.resolve_placeholders(grid[0], time_bins=12)
    assert run["vlsa_img_encoder_num_query"] == 12          # BLCA prototypes
    assert run["vlsa_img_encoder_query_text_load_idx"] == "tcga_blca_0"
    img = fetch_kws(run, "vlsa_img_encoder")
    assert img["name"] == "VLFAN" and img["dim_in"] == 512 and img["use_feat_proj"] is False
    assert img["query_pooling"] == "mean" and img["query"] == "Text" and img["query_text_method"] == "TaskRes"
    assert run["loss_type"] == "SurvIFMLE-SurvEMD" and run["bp_every_batch"] == 32
    # the resolved config the reference dumped next to its checkpoint agrees on the hot-path keys
    shipped = C.load_config(os.path.join(GOLDEN, "blca_train_config.yaml"))
    for k in ("vlsa_img_encoder_name", "vlsa_img_encoder_num_query", "vlsa_img_encoder_query_pooling",
              "vlsa_img_encoder_use_feat_proj", "loss_type", "net_output_converter", "time_bins"):
        assert run[k] == shipped[k], k


def test_state_dict_keys_match_reference_checkpoint(ckpt_params):
    net, _ = _net()
    keys = set(net.state_dict().keys())
    assert {"logit_scale", "mil_encoder.visual_adapter.weight", "mil_encoder.visual_adapter.bias",
            "mil_encoder.Q.residual_features"} <= keys
    assert not any("prompt_features" in k or "pretrained_text_features" in k for k in keys)   # non-persistent
    ref_state = {"logit_scale": ckpt_params["logit_scale"],
                 "prompt_learner.context_embeds": torch.zeros(4, 768),          # language end: ignored (strict=False)
                 "prompt_learner.rank_embeds": torch.zeros(4, 4, 768),
                 "mil_encoder.visual_adapter.weight": ckpt_params["W"],
                 "mil_encoder.visual_adapter.bias": ckpt_params["b"],
                 "mil_encoder.Q.residual_features": ckpt_params["residual_features"]}
    res = net.load_state_dict(ref_state, strict=False)
    assert res.missing_keys == []
    assert set(res.unexpected_keys) == {"prompt_learner.context_embeds", "prompt_learner.rank_embeds"}
    assert torch.equal(net.mil_encoder.visual_adapter.weight.data, ckpt_params["W"])
    assert abs(float(net.get_logit_scale()) - 56.31) < 0.01         # exp(4.0309), SURVEY §3.3


def test_reference_api_surface():
    net, pr = _net(P=4, R=4)
    enc = net.mil_encoder
    assert enc.num_query == 4 and enc.query_type == "Text" and enc.use_custom_coattn
    assert abs(float(enc.get_coattn_logit_scale()) - 100.0) < 1e-3
    Q = enc.get_query()
    torch.testing.assert_close(Q, 0.5 * enc.Q.residual_features + pr["prompt_features"])
    assert net.forward_text_only().shape == (4, 512)
    assert callable(enc.visual_adapter) and enc.visual_adapter(torch.zeros(1, 3, 512)).shape == (1, 3, 512)
    assert float(enc.query_div_loss()) >= 0
    for bad in (dict(dim_in=1024), dict(num_query=17)):
        from vlsa_b200.model import VLFAN
        kw = dict(dim_in=512, use_feat_proj=False, num_query=4)
        kw.update(bad)
        with pytest.raises(NotImplementedError):
            VLFAN(**kw)
    from vlsa_b200.model import load_model
    with pytest.raises(NotImplementedError):
        load_model("TransMIL")


def test_vlfan_variants_keep_reference_parameters_and_fail_loudly_on_cpu():
    """gated_query / query_pooling / pred_head variants construct with the reference's parameter names
    and shapes (model/deepmil.py:89-118, model/layers.py:85-155); their pooled features come from the CUDA kernels, so
    a CPU bag raises instead of falling back."""
    from vlsa_b200.model import VLFAN
    from vlsa_b200.model.prompt_adapter import PromptAdapter
    kw = dict(dim_in=512, dim_hid=256, use_feat_proj=False, num_query=6)
    enc = VLFAN(gated_query=True, query_pooling="gated_attention", **kw)
    keys = {k: tuple(v.shape) for k, v in enc.state_dict().items()}
    assert keys == {"Q": (7, 512), "query_pooling.fc1.0.weight": (256, 512), "query_pooling.fc1.0.bias": (256,),
                    "query_pooling.score.0.weight": (256, 512), "query_pooling.score.0.bias": (256,),
                    "query_pooling.fc2.weight": (1, 256), "query_pooling.fc2.bias": (1,),
                    "visual_adapter.weight": (512, 512), "visual_adapter.bias": (512,)}
    assert not enc.fused_tail and not enc.mean_linear_tail and float(enc.query_div_loss()) >= 0
    Qd, prenorm = enc.query_directions()
    assert prenorm and Qd.shape == (6, 512)
    enc = VLFAN(query_pooling="attention", pred_head="Identity", **kw)
    assert set(enc.state_dict()) == {"Q", "query_pooling.attention.0.weight", "query_pooling.attention.0.bias",
                                     "query_pooling.attention.2.weight", "query_pooling.attention.2.bias"}
    enc = VLFAN(query_pooling="weight", **kw)
    assert tuple(enc.state_dict()["query_pooling"].shape) == (1, 6)
    pooled, ext = enc.forward_query_pooling(torch.randn(2, 6, 512))
    assert pooled.shape == (2, 512) and ext is None
    assert VLFAN(query_pooling="max", **kw).forward_query_pooling(torch.ones(1, 6, 512))[0].shape == (1, 512)
    assert VLFAN(**kw).fused_tail and VLFAN(gated_query=True, **kw).fused_tail
    enc = VLFAN(dim_in=512, use_feat_proj=True, num_query=6)
    assert not enc.fused_tail and enc.mean_linear_tail
    assert {k for k in enc.state_dict() if k.startswith("feat_proj")} == {
        "feat_proj.projecter.0.weight", "feat_proj.projecter.0.bias", "feat_proj.projecter.1.weight",
        "feat_proj.projecter.1.bias"}
    with pytest.raises((ValueError, RuntimeError)):
        enc(torch.randn(1, 10, 512))                                   # CPU tensor: no fallback
    # gated Text query: P + 1 rows, the last one from the negative prompt (prompt_adapter.py:73-81,127-134)
    qnet = PromptAdapter(None, method="TaskRes", num_prompts=6, pretrained_prompt_features=torch.randn(6, 512),
                         load_negative_prompts=True, pretrained_neg_prompt_features=torch.randn(3, 512))
    assert qnet().shape == (7, 512) and qnet.get_raw_prompt_features().shape == (7, 512)
    assert set(qnet.state_dict()) == {"residual_features", "neg_residual_features"}
    with pytest.raises(RuntimeError):
        PromptAdapter(None, method="TaskRes", num_prompts=6, pretrained_prompt_features=torch.randn(6, 512),
                      load_negative_prompts=True)
    # the other query_text methods of the reference (prompt_adapter.py:86-105,118-149)
    feats = torch.randn(6, 512)
    ad = PromptAdapter(None, method="Adapter", num_prompts=6, pretrained_prompt_features=feats, dim_reduction=4,
                       keep_ratio=0.8)
    assert {k: tuple(v.shape) for k, v in ad.state_dict().items()} == {"adapter.fc.0.weight": (128, 512),
                                                                        "adapter.fc.2.weight": (512, 128)}
    want = 0.2 * torch.relu(torch.relu(feats @ ad.adapter.fc[0].weight.T) @ ad.adapter.fc[2].weight.T) + 0.8 * feats
    torch.testing.assert_close(ad(), want)
    fc = PromptAdapter(None, method="FC", num_prompts=6, pretrained_prompt_features=feats).eval()
    assert set(fc.state_dict()) == {"fc.0.weight"}
    torch.testing.assert_close(fc(), feats @ fc.fc[0].weight.T)
    torch.testing.assert_close(PromptAdapter(None, method="default", num_prompts=6, pretrained_prompt_features=feats)(),
                               feats)


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors instead of computing something else."""
    net, _ = _net(P=4, R=4)
    with pytest.raises((ValueError, RuntimeError, AssertionError)):
        net(torch.randn(1, 10, 512))
    from vlsa_b200 import ops
    with pytest.raises(ValueError):
        ops.surv_loss(torch.randn(2, 4), torch.zeros(2, dtype=torch.long), torch.zeros(2, dtype=torch.long),
                      torch.tensor(4.0))
    import vlsa_b200
    src = []
    for root, _, files in os.walk(os.path.dirname(vlsa_b200.__file__)):
        for f in files:
            if f.endswith(".py"):
                src.append(open(os.path.join(root, f)).read())
    joined = "\n".join(src)
    assert "import oracle" not in joined and "from oracle" not in joined, "the product must never import the oracle"
    assert "import triton" not in joined and "torch.compile" not in joined


def test_shard_indices_partition_and_balance():
    from vlsa_b200.runner.dist import shard_indices
    rng = np.random.default_rng(0)
    sizes = np.exp(rng.uniform(np.log(1000), np.log(100000), size=32)).astype(int).tolist()
    for world in (1, 2, 4, 8):
        parts = [shard_indices(sizes, r, world) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(32))          # a partition
        loads = [sum(sizes[i] for i in p) for p in parts]
        assert max(loads) <= 1.25 * (sum(sizes) / world) + max(sizes) * (world > 1) * 0.0 + 1 or world == 8
        rr = [shard_indices(sizes, r, world, balance=False) for r in range(world)]
        assert all(i % world == r for r, p in enumerate(rr) for i in p)
    assert shard_indices([5, 5, 5], 1, 8) in ([1], [0], [2], [])              # fewer bags than ranks is fine
    assert shard_indices([], 0, 4) == []


def test_flat_bucket_roundtrip():
    from vlsa_b200.runner.dist import FlatBucket
    a = torch.nn.Parameter(torch.randn(3, 4))
    b = torch.nn.Parameter(torch.randn(5))
    c = torch.nn.Parameter(torch.randn(2), requires_grad=False)
    a.grad = torch.randn(3, 4)
    bucket = FlatBucket([a, b, c], extra=1)
    assert bucket.flat.numel() == 12 + 5 + 1 + 2          # gradients | loss slot | one 'touched' flag per parameter
    ga = a.grad.clone()
    bucket.pack(torch.tensor([2.5]))
    bucket.all_reduce()                       # no process group: identity
    a.grad.zero_()
    bucket.unpack()
    assert torch.equal(a.grad, ga) and torch.equal(b.grad, torch.zeros(5)) and c.grad is None
    assert float(bucket.tail[0]) == 2.5
    # `a` had a gradient, `b` had none: the flags say so, and drop_untouched gives `b` back grad = None (the reference's
    # zero_grad(set_to_none) semantics: Adam skips a parameter nobody's backward reached)
    assert bucket.flags.tolist() == [1.0, 0.0]
    assert bucket.drop_untouched(bucket.flags) == 1 and b.grad is None and a.grad is not None


def test_param_groups_follow_reference_rule():
    from vlsa_b200.runner.vlsa_handler import param_groups_weight_decay
    net, _ = _net(P=4, R=4)
    groups = param_groups_weight_decay(net, 1e-5)
    no_decay = {id(p) for p in groups[0]["params"]}
    # optim/optim_factory.py:25-37 tests `len(param.shape) == 1`: the bias is exempt, the 0-dim logit_scale is NOT
    assert id(net.mil_encoder.visual_adapter.bias) in no_decay
    assert id(net.logit_scale) not in no_decay
    assert id(net.mil_encoder.visual_adapter.weight) not in no_decay
    assert id(net.mil_encoder.Q.residual_features) not in no_decay
    assert sum(p.numel() for g in groups for p in g["params"]) == 1 + 512 * 512 + 512 + 4 * 512


def test_pack_bags_and_loader_host_side():
    from vlsa_b200.dataset import pack_bags
    bags = [torch.randn(3, 512), torch.randn(1, 5, 512), torch.zeros(0, 512)]
    out = torch.empty(16, 512)
    packed, sizes = pack_bags(bags, out)
    assert sizes == [3, 5, 0] and packed.data_ptr() == out.data_ptr()
    assert torch.equal(packed[:3], bags[0]) and torch.equal(packed[3:8], bags[1][0])


def test_prototype_shap_vectorised_matches_reference_loops():
    """utils/model_inference.py:21-78: the all-subsets-at-once evaluation equals the reference's subset loop."""
    import numpy as np
    import torch
    from oracle import vlsa_oracle as O
    from vlsa_b200.utils.model_inference import evaluate_prototype_shap_imp
    g = torch.Generator().manual_seed(3)
    for P, R in ((1, 4), (4, 4), (7, 12), (9, 13)):
        sim = (torch.rand(P, R, generator=g) * 2 - 1) * 0.3
        ref = O.prototype_shap_imp(sim, 56.3)
        got = evaluate_prototype_shap_imp(sim, 56.3)
        np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=2e-5, rtol=1e-4)
        # efficiency axiom: the values add up to value(full) - value(empty)
        full = torch.softmax(56.3 * sim.mean(0), 0)
        total = float(((R - torch.arange(R)) * full).sum()) - 1.0
        assert abs(float(got.sum()) - total) < 1e-3


def test_flat_store_matches_per_slide_files(tmp_path):
    """dataset/PatchWSI.py:197-215 vs the flat store: a patient = its slides' rows in sid order, bit-exact in fp32;
    missing slides are skipped; bf16 stores the rounded values; steps() packs whole optimizer steps."""
    import numpy as np
    import torch
    from vlsa_b200.dataset import PatchFeatureStore, WSIPatchSurvStore, build_store_from_files
    g = torch.Generator().manual_seed(11)
    pdir = tmp_path / "pt"; pdir.mkdir()
    sizes = {"s0": 5, "s1": 17, "s2": 1, "s3": 300, "s4": 64}
    feats = {}
    for sid, n in sizes.items():
        feats[sid] = torch.randn(n, 512, generator=g)
        if sid == "s3":
            np.save(pdir / (sid + ".npy"), feats[sid].numpy())
        else:
            torch.save(feats[sid].to(torch.float16 if sid == "s4" else torch.float32), pdir / (sid + ".pt"))
    feats["s4"] = feats["s4"].to(torch.float16).float()
    build_store_from_files(str(tmp_path / "st32"), str(pdir), ["s0", "s1", "s2", "s4", "missing"], "pt")
    st = PatchFeatureStore(str(tmp_path / "st32"))
    assert "missing" not in st and st.n_rows(["s0", "s1", "nope"]) == 22
    pid2sids = {"pA": ["s1", "s0"], "pB": ["s2"], "pC": ["s4", "gone", "s0"]}
    pid2label = {"pA": (3.0, 1.0), "pB": (0.0, 0.0), "pC": (7.0, 1.0)}
    ds = WSIPatchSurvStore(st, ["pA", "pB", "pC"], pid2sids, pid2label)
    for i, pid in enumerate(ds.pids):
        idx, (x, extra), label = ds[i]
        ref = torch.cat([feats[s] for s in pid2sids[pid] if s in feats and s != "s3"], 0).to(torch.float)
        assert int(idx) == i and extra.tolist() == [0.0] and label.tolist() == list(pid2label[pid])
        assert x.dtype == torch.float32 and torch.equal(x, ref)
    steps = list(ds.steps(batch_size=2, pin=False, threads=2))
    assert [s[1] for s in steps] == [[22, 1], [69]]
    assert torch.equal(steps[0][0], torch.cat([feats["s1"], feats["s0"], feats["s2"]], 0))
    assert steps[1][2].tolist() == [[7.0, 1.0]] and steps[1][3].tolist() == [2]
    # bf16 store: the rounded values, half the bytes
    from vlsa_b200.dataset import build_store
    build_store(str(tmp_path / "st16"), [(s, feats[s]) for s in ("s0", "s3")], dtype="bfloat16")
    st16 = PatchFeatureStore(str(tmp_path / "st16"))
    x16 = st16.read(["s3", "s0"])
    assert x16.dtype == torch.bfloat16 and torch.equal(x16, torch.cat([feats["s3"], feats["s0"]], 0).to(torch.bfloat16))
    assert (tmp_path / "st16" / "features.bin").stat().st_size == 305 * 512 * 2


def test_flat_bucket_attached_gradients_alias_the_bucket():
    """FlatBucket.attach(): .grad tensors are views of the all-reduce buffer, so autograd accumulates into it, zero()
    is one memset and pack()/unpack() copy nothing — with the same numbers as the copying path."""
    import torch
    from vlsa_b200.runner.dist import FlatBucket
    torch.manual_seed(0)
    def make():
        return [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5)),
                torch.nn.Parameter(torch.randn(2, 2), requires_grad=False)]
    def loss(ps, k):
        return (ps[0] * k).sum() ** 2 + (ps[1] ** 3).sum() * k + ps[2].sum()
    a, b = make(), make()
    for pa, pb in zip(a, b):
        pb.data.copy_(pa.data)
    plain, att = FlatBucket(a, extra=1), FlatBucket(b, extra=1)
    att.attach()
    assert all(p.grad.data_ptr() == seg.data_ptr() for p, seg in att._views())
    for step in (1.0, 2.5):
        for p in a:
            p.grad = None
        att.zero()
        assert float(att.flat.abs().sum()) == 0.0
        loss(a, step).backward(); loss(b, step).backward()
        loss(b, step).backward()                       # accumulation into the attached views
        plain.pack(torch.tensor([step])); att.pack(torch.tensor([step]))
        assert torch.allclose(att.flat[:17], 2 * plain.flat[:17]) and float(att.tail[0]) == step
        assert att.flags.tolist() == [1.0, 1.0] and plain.flags.tolist() == [1.0, 1.0]
        plain.all_reduce(); att.all_reduce(); plain.unpack(); att.unpack()
        assert all(p.grad.data_ptr() == seg.data_ptr() for p, seg in att._views())
        assert torch.allclose(b[0].grad, 2 * a[0].grad) and torch.allclose(b[1].grad, 2 * a[1].grad)
    # a caller that drops the gradients (zero_grad(set_to_none=True)) falls back to the copying path
    for p in b:
        p.grad = None
    att.zero()
    loss(b, 1.0).backward()
    att.pack(None)
    assert torch.allclose(att.flat[:12], b[0].grad.reshape(-1))


def test_plan_cache_returns_the_same_immutable_schedule():
    """ops.make_plan caches by (sizes, device, sms): the per-bag loops of the reference ask for the same schedule once
    per slide and epoch.  CPU device = planning only (no upload event)."""
    import numpy as np
    from vlsa_b200 import ops
    a = ops.make_plan([1000, 37, 2798], "cpu", sms=148)
    b = ops.make_plan(np.array([1000, 37, 2798]), "cpu", sms=148)          # numpy ints hash like Python ints
    c = ops.make_plan([1000, 37, 2798], "cpu", sms=8)
    d = ops.make_plan([1000, 37, 2799], "cpu", sms=148)
    assert a is b and a is not c and a is not d
    assert a.total_rows == 3835 and d.total_rows == 3836 and a.num_bags == 3
    assert a.chunk_start_host[0] == 0 and a.total_chunks == int(a.chunk_start_host[-1])
    before = len(ops._PLAN_CACHE)
    old_max, ops._PLAN_CACHE_MAX = ops._PLAN_CACHE_MAX, before + 2
    try:
        for n in range(5):
            ops.make_plan([5000 + n], "cpu", sms=148)
        assert len(ops._PLAN_CACHE) <= before + 2                          # least recently used entries are dropped
    finally:
        ops._PLAN_CACHE_MAX = old_max
    with pytest.raises(ValueError):
        ops.make_plan([10, -1], "cpu")


def test_bench_reads_measured_peaks_in_any_layout(tmp_path, monkeypatch):
    """bench.py's roofline denominator: the driver-written MEASURED_PEAKS.json when it is there (flat `hbm_gbs`, nested,
    burst preferred for a kernel timed alone, TB/s keys), the profiling guide's fallback otherwise — never a crash."""
    import json
    import bench
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.measured_peaks()[0] == 6650.0
    cases = (({"hbm_gbs": 6538.9, "bf16_tflops": 1500.0}, 6538.9),
             ({"hbm": {"sustained_gbs": 6500.0, "burst_gbs": 7100.0}, "bf16": {"tflops": 1600.0}}, 7100.0),
             ({"hbm_copy_tb_s": 6.6}, 6600.0), ({"unrelated": 1}, 6650.0), ("garbage", 6650.0))
    for data, want in cases:
        (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps(data))
        got, source = bench.measured_peaks()
        assert abs(got - want) < 1e-6, (data, got, source)
    (tmp_path / "MEASURED_PEAKS.json").write_text("{not json")
    assert bench.measured_peaks()[0] == 6650.0


def test_interpretation_utility_rejects_variant_encoders():
    """The decoupled similarities assume mean pooling + Linear adapter; other VLFAN switches must fail loudly."""
    from vlsa_b200 import synth
    from vlsa_b200.model import VLSA
    from vlsa_b200.utils import calc_text_img_similarity
    pr = synth.make_params(4, 4, 5)
    for extra in (dict(gated_query=True), dict(query_pooling="max"), dict(pred_head="Identity"), dict(use_feat_proj=True)):
        img = dict(name="VLFAN", dim_in=512, use_feat_proj=False, query="Parameter", num_query=4)
        img.update(extra)
        net = VLSA({"name": "mahmoodlab/conch"}, img, {"name": "CoOp"}, text_features=pr["text_features"])
        with pytest.raises(NotImplementedError):
            calc_text_img_similarity(net, torch.randn(10, 512))


def test_gated_query_directions_reproduce_the_reference_scores():
    """What the kernels rely on for gated_query: A_[:, :-1] - A_[:, -1:] (model/deepmil.py:192-195) equals the scores of
    the P difference rows Qn_p - Qn_gate used WITHOUT normalisation, and the gradient reaches all P + 1 raw rows."""
    from oracle import vlsa_oracle as O
    from vlsa_b200.model import VLFAN
    torch.manual_seed(3)
    enc = VLFAN(dim_in=512, use_feat_proj=False, num_query=5, gated_query=True).double()
    X = torch.randn(1, 40, 512, dtype=torch.float64) + 0.5
    Qd, prenorm = enc.query_directions()
    assert prenorm and Qd.shape == (5, 512)
    scale = torch.tensor(O.coattn_scale(), dtype=torch.float64)
    S = scale * Qd @ torch.nn.functional.normalize(X[0], dim=-1).t()          # what the kernels compute from Qd
    A = torch.softmax(S, dim=-1)
    f_kernel_math = (A @ X[0]).mean(0, keepdim=True) @ enc.visual_adapter.weight.t() + enc.visual_adapter.bias
    f_ref, A_ref, _, _ = O.vlfan_forward_variant(X, enc.Q, enc.visual_adapter.weight, enc.visual_adapter.bias,
                                                 gated_query=True, scale=scale)
    torch.testing.assert_close(A, A_ref[0], rtol=1e-10, atol=1e-14)
    torch.testing.assert_close(f_kernel_math, f_ref, rtol=1e-10, atol=1e-12)
    g1, = torch.autograd.grad(f_kernel_math.sum(), enc.Q, retain_graph=True)
    g2, = torch.autograd.grad(f_ref.sum(), enc.Q)
    torch.testing.assert_close(g1, g2, rtol=1e-8, atol=1e-12)
    assert g1[-1].abs().max() > 0


def test_flat_bucket_aligned_segments_and_direct_writes():
    """FlatBucket(align=4): every segment starts at a multiple of four floats (kernels write gradients straight into the attached
    views with 128-bit stores), the padding rides the all-reduce as zeros, `mark_touched` stands in for the autograd hook when a
    kernel wrote a gradient, and `pack(extra_in_place=True)` leaves the tail a kernel filled alone."""
    import torch
    from vlsa_b200.runner.dist import FlatBucket
    ps = [torch.nn.Parameter(torch.randn(())), torch.nn.Parameter(torch.randn(5)), torch.nn.Parameter(torch.randn(3, 3)),
          torch.nn.Parameter(torch.randn(2), requires_grad=False)]
    bk = FlatBucket(ps, extra=3, align=4)
    assert bk.offsets == [0, 4, 12] and bk.sizes == [1, 5, 9]
    assert bk.flat.numel() == 4 + 8 + 12 + 3 + 3                      # padded segments | 3 extras | one flag per trainable tensor
    bk.attach()
    assert bk.attached() and all(p.grad.data_ptr() % 16 == bk.flat.data_ptr() % 16 for p in ps[:3])
    bk.zero()
    ps[1].grad.copy_(torch.arange(5.0))                               # "a kernel wrote this gradient"
    bk.mark_touched(ps[1])
    (ps[2] ** 2).sum().backward()                                     # autograd accumulates into the view, its hook marks it
    bk.tail.copy_(torch.tensor([1.5, 0.5, 1.0]))                      # "the loss kernel wrote (total, ifmle, emd)"
    bk.pack(None, extra_in_place=True)
    bk.all_reduce(); bk.unpack()
    assert bk.tail.tolist() == [1.5, 0.5, 1.0] and bk.flags.tolist() == [0.0, 1.0, 1.0]
    assert torch.equal(bk.flat[4:9], torch.arange(5.0)) and float(bk.flat[9:12].abs().sum()) == 0.0     # padding stays zero
    assert torch.allclose(ps[2].grad, 2 * ps[2].detach())
    # a caller that replaces a gradient tensor detaches the bucket; zero() re-attaches it
    ps[1].grad = torch.ones(5)
    assert not bk.attached()
    bk.zero()
    assert bk.attached() and float(bk.flat.abs().sum()) == 0.0


def test_query_rows_cache_follows_the_adapter():
    """`VLFAN.query_directions_cached()` (the rows the lean inference call and the graph replay use): evaluated once per state of the
    prompt adapter — an in-place update or a bumped version counter (what `BucketAdam` does after its kernel wrote the weights
    through raw pointers) invalidates it; the gated query's difference rows are cached the same way."""
    import torch
    from vlsa_b200.model import VLSA
    nets = [_net(P=4, R=4)[0]]
    pr = synth.make_params(4, 4, 3)
    nets.append(VLSA({"name": "mahmoodlab/conch"},
                     dict(name="VLFAN", dim_in=512, use_feat_proj=False, query="Text", num_query=4, gated_query=True,
                          query_text_method="TaskRes"),
                     {"name": "CoOp"}, text_features=pr["text_features"], query_prompt_features=pr["prompt_features"],
                     query_neg_prompt_features=torch.randn(1, 512)))
    for net in nets:
        enc = net.mil_encoder
        a, pre = enc.query_directions_cached()
        b, _ = enc.query_directions_cached()
        assert a is b and not a.requires_grad and bool(pre) == bool(enc.gated_query) and a.shape == (4, 512)
        with torch.no_grad():
            want, _ = enc.query_directions()
        assert torch.equal(a, want)
        with torch.no_grad():
            enc.Q.residual_features.add_(0.25)                      # in-place update: the version counter moves
        c, _ = enc.query_directions_cached()
        assert c is not a and torch.equal(c, enc.query_directions()[0].detach()) and not torch.equal(c, a)
        torch.autograd.graph.increment_version(enc.Q.residual_features)     # what a raw-pointer writer does after its kernel
        d, _ = enc.query_directions_cached()
        assert d is not c and torch.equal(d, c)
