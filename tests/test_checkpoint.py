"""Host-side bookkeeping: EarlyStopping mirrors against traces produced by the
reference's own classes (tests/golden/make_earlystopping_golden.py), checkpoint dict keys."""
import json
import os

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden():
    with open(os.path.join(HERE, "golden", "earlystopping_golden.json")) as f:
        return json.load(f)


def test_early_stopping_follows_reference_trace(tmp_path, monkeypatch, capsys):
    from bigcn_b200.checkpoint import EarlyStopping
    monkeypatch.chdir(tmp_path)
    model = torch.nn.Linear(2, 2)
    for case in _golden()["four"]:
        for f in os.listdir(tmp_path):
            os.remove(f)
        e = EarlyStopping(patience=case["patience"], verbose=True)
        for ep, t in enumerate(case["trace"]):
            ck = {"fold": 2, "iter": 0, "epoch": ep, "loss": round(t["val_loss"] * 1.1, 6)}
            e(t["val_loss"], *t["scores"], model, "BiGCN", "Twitter16", checkpoint=ck)
            assert (e.counter, e.best_score, e.early_stop) == (t["counter"], t["best_score"], t["early_stop"])
            assert [e.accs, e.F1, e.F2, e.F3, e.F4] == t["kept"]
            assert e.checkpoint["epoch"] == t["kept_epoch"]
            assert sorted(os.listdir(tmp_path)) == t["files"]
        assert case["trace"][-1]["early_stop"] == e.early_stop
        if e.early_stop:
            saved = torch.load(e.saved_path)
            assert saved["epoch"] == case["trace"][-1]["kept_epoch"]
    assert "BEST Accuracy" in capsys.readouterr().out


def test_early_stopping_2class_follows_reference_trace(tmp_path, monkeypatch):
    from bigcn_b200.checkpoint import EarlyStopping2class
    monkeypatch.chdir(tmp_path)
    model = torch.nn.Linear(2, 2)
    for case in _golden()["two"]:
        for f in os.listdir(tmp_path):
            os.remove(f)
        e = EarlyStopping2class(patience=case["patience"], verbose=True)
        for t in case["trace"]:
            e(t["val_loss"], *t["scores"], model, "BiGCN", "Weibo")
            assert (e.counter, e.best_score, e.early_stop) == (t["counter"], t["best_score"], t["early_stop"])
            assert [e.accs, e.acc1, e.acc2, e.pre1, e.pre2, e.rec1, e.rec2, e.F1, e.F2] == t["kept"]
            assert e.val_loss_min == t["val_loss_min"]
            assert sorted(os.listdir(tmp_path)) == t["files"]
        assert set(torch.load("BiGCNWeibo.m").keys()) == {"weight", "bias"}


def test_make_checkpoint_keys_and_torch_optimizer():
    """BiGCN_Twitter.py:253-261: same keys; a plain torch optimizer is accepted as well."""
    from bigcn_b200.checkpoint import make_checkpoint
    model = torch.nn.Linear(3, 2)
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
    ck = make_checkpoint(model, opt, fold=1, iter=0, epoch=7, loss=0.25, res=["acc:0.5"])
    assert list(ck.keys()) == ["fold", "iter", "epoch", "model_state_dict", "optimizer_state_dict", "loss", "res"]
    with torch.no_grad():
        model.weight.add_(1.0)                  # the kept dict does not follow later updates
    assert not torch.equal(ck["model_state_dict"]["weight"], model.weight)


@pytest.mark.gpu
def test_fused_trainer_optimizer_state_interchanges_with_torch_adam():
    """optimizer_state_dict() loads into the Adam the reference builds (:146-153) and that optimizer's
    next step equals FusedTrainer's; load_optimizer_state_dict() restores a trainer."""
    import bigcn_b200 as bb
    from bigcn_b200.data import make_batch
    dev = torch.device("cuda:0")
    torch.manual_seed(0)

    def to_dev(b):
        for k in ("x", "edge_index", "BU_edge_index", "batch", "rootindex", "y"):
            setattr(b, k, getattr(b, k).to(dev))
        return b
    batches = [to_dev(make_batch("twitter15", 12, seed=s, train=True, in_feats=300, num_classes=4)) for s in range(3)]
    model = bb.BiGCN(300, 64, 64, dev).to(dev)
    model.train()
    tr = bb.FusedTrainer(model, lr=5e-4, weight_decay=1e-4)
    for b in batches[:2]:
        tr.step(b, seed=11)
    sd = tr.optimizer_state_dict()
    ck = bb.make_checkpoint(model, tr, 0, 0, 1, 0.5, [])
    assert set(ck["model_state_dict"]) == {f"{d}.{c}.{p}" for d in ("TDrumorGCN", "BUrumorGCN") for c in ("conv1", "conv2")
                                           for p in ("lin.weight", "bias")} | {"fc.weight", "fc.bias"}
    # the reference's optimizer, built the reference's way, on a copy of the model
    ref = bb.BiGCN(300, 64, 64, dev).to(dev)
    ref.load_state_dict(ck["model_state_dict"])
    ref.train()
    bu = list(map(id, ref.BUrumorGCN.conv1.parameters())) + list(map(id, ref.BUrumorGCN.conv2.parameters()))
    base = [p for p in ref.parameters() if id(p) not in bu]
    opt = torch.optim.Adam([{"params": base}, {"params": ref.BUrumorGCN.conv1.parameters(), "lr": 5e-4 / 5},
                            {"params": ref.BUrumorGCN.conv2.parameters(), "lr": 5e-4 / 5}], lr=5e-4, weight_decay=1e-4)
    import copy
    opt.load_state_dict(copy.deepcopy(sd))        # torch keeps the tensors it is handed and updates them in place
    assert [g["lr"] for g in opt.param_groups] == [5e-4, 1e-4, 1e-4]
    ref.TDrumorGCN.seed, ref.TDrumorGCN._calls = 11, 0        # the module path then draws the mask of seed 11
    b = batches[2]
    loss = torch.nn.functional.nll_loss(ref(b), b.y)
    opt.zero_grad()
    loss.backward()
    opt.step()
    # second trainer restored from the exported state
    m2 = bb.BiGCN(300, 64, 64, dev).to(dev)
    m2.load_state_dict(ck["model_state_dict"])
    m2.train()
    tr2 = bb.FusedTrainer(m2, lr=1.0, weight_decay=1e-4)      # lr comes back from the state
    tr2.load_optimizer_state_dict(sd)
    tr.step(b, seed=11)
    tr2.step(b, seed=11)
    for (n, p), (_, q) in zip(model.named_parameters(), m2.named_parameters()):
        assert torch.equal(p, q), n
    assert tr2.optimizer_state_dict()["state"][0]["step"].item() == 3.0
    sd3 = tr.optimizer_state_dict()
    for i, st in opt.state_dict()["state"].items():
        assert float(st["step"]) == 3.0
        for k in ("exp_avg", "exp_avg_sq"):
            d = (st[k] - sd3["state"][i][k]).abs().max() / st[k].abs().max().clamp_min(1e-30)
            assert float(d) < 1e-4, (i, k)
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        assert float((p - q).abs().max() / q.abs().max()) < 2e-5, n
