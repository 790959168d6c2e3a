"""CPU, world_size 2, gloo: the N>1 path of `VLSAHandler._update_network` — bag sharding, loss normalisation over
the GLOBAL batch, the single flat-bucket all-reduce and the identical Adam step — must reproduce the
single-process step.  The CUDA arithmetic is replaced by the CPU oracle through test-only injection (the oracle
is test infrastructure); everything else is the product's host code, unchanged."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vlsa_b200 import synth

P = R = 4
SIZES = [300, 40, 1200, 7, 513, 64, 900, 33]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_handler():
    from oracle import vlsa_oracle as O
    from vlsa_b200.model import VLSA
    from vlsa_b200.runner import VLSAHandler
    pr = synth.make_params(P, R, 5)
    net = VLSA({"name": "mahmoodlab/conch"},
               dict(name="VLFAN", dim_in=512, use_feat_proj=False, query="Text", num_query=P, gated_query=False,
                    query_pooling="mean", pred_head="default", query_text_method="TaskRes", query_text_res_ratio=0.5),
               {"name": "CoOp"}, text_features=pr["text_features"], query_prompt_features=pr["prompt_features"],
               logit_scale_init=4.0309, vlsa_api="CONCH", path_clip_model=None)
    with torch.no_grad():
        net.mil_encoder.Q.residual_features.copy_(pr["residual_features"])
        net.mil_encoder.visual_adapter.weight.copy_(pr["W"])
        net.mil_encoder.visual_adapter.bias.copy_(pr["b"])
    cfg = dict(task="vlsa", arch="VLSA", net_output_converter="softmax", loss_type="SurvIFMLE-SurvEMD",
               loss_survifmle_weight=1.0, loss_survemd_weight=1.0, opt_name="adam", opt_lr=2e-4, opt_weight_decay=1e-5,
               bp_every_batch=len(SIZES))
    h = VLSAHandler(cfg, net, device="cpu")

    # ---- test-only injection of the CPU oracle in place of the CUDA kernels -------------------------
    def forward_packed(X, plan, text_features=None):
        T = net.forward_text_only() if text_features is None else text_features
        enc = net.mil_encoder
        outs = []
        cu = plan.cu_rows_host
        for i in range(plan.num_bags):
            Xi = X[int(cu[i]):int(cu[i + 1])].unsqueeze(0)
            outs.append(O.vlsa_forward(Xi, enc.get_query(), enc.visual_adapter.weight, enc.visual_adapter.bias, T,
                                       net.logit_scale)[0])
        logits = torch.cat(outs, 0)
        return logits, None, None, torch.softmax(logits, -1)

    def calc_objective_loss(raw_pred, label, norm=None):
        p = O.softmax_converter(raw_pred)
        t, e = label[:, 0].long(), label[:, 1].long()
        per = O.surv_ifmle(p, t, e, reduction="none").reshape(-1) + O.surv_emd(p, t, e, net.get_logit_scale(), reduction="none")
        return per.sum() / float(norm or raw_pred.shape[0])

    net.forward_packed = forward_packed
    h.calc_objective_loss = calc_objective_loss
    return h


def _batch():
    xs = [synth.make_bag("g1", n, 900 + i).unsqueeze(0) for i, n in enumerate(SIZES)]
    t, e = synth.make_labels(len(SIZES), R, 17)
    ys = [torch.stack([t[i], e[i]]).float().reshape(1, 2) for i in range(len(SIZES))]
    return xs, ys


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h = _make_handler()
        assert (h.rank, h.world_size) == (rank, world)
        xs, ys = _batch()
        losses = []
        for _ in range(2):                                    # two optimizer steps
            loss, preds = h._update_network(xs, ys)
            losses.append(loss)
        state = {k: v.detach().clone() for k, v in h.net.state_dict().items()}
        torch.save({"loss": losses, "preds": preds, "state": state}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_step_equals_single_process_step(tmp_path):
    torch.set_num_threads(4)
    h = _make_handler()
    xs, ys = _batch()
    ref_losses = []
    for _ in range(2):
        loss, ref_preds = h._update_network(xs, ys)
        ref_losses.append(loss)
    ref_state = {k: v.detach().clone() for k, v in h.net.state_dict().items()}

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    # both ranks end with identical replicas ...
    for k in ref_state:
        assert torch.equal(r0["state"][k], r1["state"][k]), k
    # ... that match the single-process result (fp32 summation order differs across the shard boundary)
    for k in ref_state:
        torch.testing.assert_close(r0["state"][k], ref_state[k], rtol=1e-5, atol=1e-7, msg=k)
    np.testing.assert_allclose(r0["loss"], ref_losses, rtol=1e-5)
    np.testing.assert_allclose(r1["loss"], ref_losses, rtol=1e-5)
    torch.testing.assert_close(r0["preds"], ref_preds, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(r1["preds"], r0["preds"], rtol=0, atol=0)
    # and the parameters actually moved
    assert not torch.equal(ref_state["mil_encoder.visual_adapter.weight"], synth.make_params(P, R, 5)["W"])


def _worker_bcast(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vlsa_b200.model import VLSA
        from vlsa_b200.runner import VLSAHandler
        torch.manual_seed(1000 + rank)                        # replicas start from DIFFERENT random states
        pr = synth.make_params(P, R, 5)
        net = VLSA({"name": "mahmoodlab/conch"},
                   dict(name="VLFAN", dim_in=512, use_feat_proj=False, query="Text", num_query=P, gated_query=False,
                        query_pooling="mean", pred_head="default", query_text_method="TaskRes", query_text_res_ratio=0.5),
                   {"name": "CoOp"}, text_features=pr["text_features"], query_prompt_features=pr["prompt_features"],
                   vlsa_api="CONCH", path_clip_model=None)
        before = net.mil_encoder.visual_adapter.weight.detach().clone()
        h = VLSAHandler(dict(task="vlsa", arch="VLSA", opt_name="adam"), net, device="cpu")
        state = {k: v.detach().clone() for k, v in h.net.state_dict().items()}
        torch.save({"before": before, "state": state}, os.path.join(out_dir, f"b{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_handler_broadcasts_rank0_parameters(tmp_path):
    """Replicas built from different random states (nn.Linear / randn initialisers) must be identical before the
    first step: the handler broadcasts rank 0's parameters at construction, like DDP."""
    port = _free_port()
    mp.spawn(_worker_bcast, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    b0, b1 = torch.load(tmp_path / "b0.pt"), torch.load(tmp_path / "b1.pt")
    assert not torch.equal(b0["before"], b1["before"])                    # they did start differently
    for k in b0["state"]:
        assert torch.equal(b0["state"][k], b1["state"][k]), k
    assert torch.equal(b0["state"]["mil_encoder.visual_adapter.weight"], b0["before"])   # rank 0 is the source


@pytest.mark.timeout(600)
def test_train_each_epoch_groups_steps_and_reads_back_once():
    """`_train_each_epoch` (runner/vlsa_handler.py:189-239 of the reference): bags grouped `bp_every_batch` at a time, one optimizer
    step per group through the no-synchronisation path, losses and raw predictions converted once at the end of the epoch —
    against a twin handler stepping the same groups one `_update_network` call at a time."""
    torch.set_num_threads(4)
    ha, hb = _make_handler(), _make_handler()
    ha.cfg["bp_every_batch"] = 3
    xs, ys = _batch()
    loader = [(torch.tensor([[10 + i]]), (xs[i], torch.zeros(1)), ys[i]) for i in range(len(xs))]
    out = ha._train_each_epoch(0, loader)
    groups = [list(range(0, 3)), list(range(3, 6)), list(range(6, 8))]
    ref_loss, ref_pred = [], []
    for g in groups:
        l, p = hb._update_network([xs[i] for i in g], [ys[i] for i in g])
        ref_loss.append(l); ref_pred.append(p)
    assert isinstance(out["loss"], list) and len(out["loss"]) == 3 and all(isinstance(v, float) for v in out["loss"])
    np.testing.assert_allclose(out["loss"], ref_loss, rtol=1e-6)
    torch.testing.assert_close(out["pred"]["raw_y_hat"], torch.cat(ref_pred, 0), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(out["pred"]["y_hat"], torch.softmax(torch.cat(ref_pred, 0), -1), rtol=1e-6, atol=1e-6)
    assert out["pred"]["uid"].tolist() == [10 + i for i in range(len(xs))]
    torch.testing.assert_close(out["pred"]["y"], torch.cat([y.reshape(1, 2) for y in ys], 0))
    for (k, va), (_, vb) in zip(ha.net.state_dict().items(), hb.net.state_dict().items()):
        torch.testing.assert_close(va, vb, rtol=1e-6, atol=1e-8, msg=k)
