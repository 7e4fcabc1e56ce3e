"""The training loop around the path (BiGCN_Twitter.py:134-275) on a small Twitter16-shaped fold:
device dataset -> batches with DropEdge -> FusedTrainer -> device metrics -> early stopping -> checkpoint."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_train_gcn_learns_and_stops(tmp_path, monkeypatch):
    import bigcn_b200
    from bigcn_b200.data import make_trees_shard
    monkeypatch.chdir(tmp_path)
    dev = torch.device("cuda:0")
    K = 400
    trees = make_trees_shard("twitter16", 96, seed=5, in_feats=K)[0]
    # make the label learnable: the root's bag of words carries a class-specific column
    for t in trees:
        y = int(t.y)
        t.x[int(t.rootindex), :] = 0
        t.x[int(t.rootindex), y * 7:(y * 7) + 5] = 3.0
    forest = bigcn_b200.DeviceForest.from_data_list(trees, dev)
    ids = np.arange(96)
    train_ids, test_ids = ids[:72], ids[72:]
    torch.manual_seed(0)
    model = bigcn_b200.BiGCN(K, 64, 64, dev, gemm_mode="sparse").to(dev)
    logs = []
    out = bigcn_b200.train_GCN(model, forest, train_ids, test_ids, 0.2, 0.2, lr=5e-3, weight_decay=1e-4, patience=4,
                               n_epochs=25, batchsize=24, datasetname="Twitter16", iter=0, fold=1, log=logs.append)
    train_losses, val_losses, train_accs, val_accs, accs, F1, F2, F3, F4 = out
    assert len(train_losses) == len(val_losses) >= 5
    assert train_losses[-1] < 0.6 * train_losses[0]          # it learns
    assert max(val_accs) > 0.6 and accs >= 0.5
    assert all(np.isfinite(train_losses)) and all(np.isfinite(val_losses))
    assert any(l.startswith("Fold 1 | Epoch") for l in logs)
    saved = [f for f in os.listdir(tmp_path) if f.startswith("best_BiGCN_Twitter16_f1_i0_e")]
    if "Early stopping" in logs:
        assert len(saved) == 1
        ck = torch.load(saved[0], weights_only=False)
        assert set(ck) == {"fold", "iter", "epoch", "model_state_dict", "optimizer_state_dict", "loss", "res"}
        m2 = bigcn_b200.BiGCN(K, 64, 64, dev, gemm_mode="sparse").to(dev)
        m2.load_state_dict(ck["model_state_dict"])
        assert len(ck["optimizer_state_dict"]["param_groups"]) == 3


def test_final_checkpoint_when_epochs_run_out(tmp_path, monkeypatch):
    """BiGCN_Twitter.py:311-325: a run that exhausts n_epochs without early stopping leaves
    checkpoints/final_bigcn_f{fold}_i{iter}_e{epoch}_l{loss}.pt (the for/else branch)."""
    import bigcn_b200
    from bigcn_b200.data import make_trees_shard
    monkeypatch.chdir(tmp_path)
    dev = torch.device("cuda:0")
    K = 64
    trees = make_trees_shard("twitter16", 24, seed=6, in_feats=K)[0]
    forest = bigcn_b200.DeviceForest.from_data_list(trees, dev)
    torch.manual_seed(0)
    model = bigcn_b200.BiGCN(K, 64, 64, dev, gemm_mode="sparse").to(dev)
    out = bigcn_b200.train_GCN(model, forest, np.arange(16), np.arange(16, 24), 0.2, 0.2, lr=5e-4, weight_decay=1e-4,
                               patience=100, n_epochs=2, batchsize=8, datasetname="Twitter16", iter=3, fold=2, log=lambda s: None)
    assert len(out[0]) == 2
    files = os.listdir(tmp_path / "checkpoints")
    assert len(files) == 1 and files[0].startswith("final_bigcn_f2_i3_e00001_l") and files[0].endswith(".pt")
    ck = torch.load(tmp_path / "checkpoints" / files[0], weights_only=False)
    assert ck["epoch"] == 1 and "model_state_dict" in ck and "optimizer_state_dict" in ck
    assert bigcn_b200.train_GCN.last_final_checkpoint.endswith(files[0])
