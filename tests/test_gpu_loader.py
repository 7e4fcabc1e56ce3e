"""On-device batch assembly + DropEdge against the host collate of
bigcn_b200.data (PyG's offset rules) and the DropEdge contract of Process/dataset.py:68-90."""
import numpy as np
import pytest
import torch

from oracle import bigcn_oracle, gcn_oracle
from bigcn_b200.data import Batch, collate, make_tree

pytestmark = pytest.mark.gpu


def trees_for(k=300, sizes=(1, 60, 2, 350, 17, 90, 1, 5), shape="twitter15", seed=0):
    rng = np.random.default_rng(seed)
    return [make_tree(shape, int(n), rng, in_feats=k) for n in sizes]


def test_assembly_without_dropedge_equals_collate(dev):
    import bigcn_b200
    trees = trees_for()
    forest = bigcn_b200.DeviceForest.from_data_list(trees, dev)
    ids = [3, 0, 7, 1, 1, 5, 2]                       # arbitrary order, a repeated tree, single-node trees
    got = forest.batch(ids)
    want = collate([trees[i] for i in ids])
    assert torch.equal(got.edge_index.cpu(), want.edge_index)
    assert torch.equal(got.BU_edge_index.cpu(), want.BU_edge_index)
    assert torch.equal(got.batch.cpu(), want.batch)
    assert torch.equal(got.rootindex.cpu(), want.rootindex)
    assert torch.equal(got.y.cpu(), want.y)
    assert torch.equal(got.x.to_dense().cpu(), want.x)
    empty = forest.batch([])
    assert empty.batch.numel() == 0 and empty.x.shape == (0, 300)


def test_dropedge_contract(dev):
    """keep int(e*(1-rate)) positions, a subsequence of the original list (order preserved), TD and
    BU drawn independently, reproducible per seed, every position equally likely to be dropped."""
    import bigcn_b200
    trees = trees_for(k=16, sizes=(6, 40, 301, 2, 1))
    forest = bigcn_b200.DeviceForest.from_data_list(trees, dev)
    ids = [0, 1, 2, 3, 4]
    full = collate([trees[i] for i in ids])
    node_off = np.concatenate([[0], np.cumsum([trees[i].num_nodes for i in ids])])

    def per_tree(ei, b):
        m = (ei[1] >= node_off[b]) & (ei[1] < node_off[b + 1])
        return ei[:, m]

    def is_subsequence(sub, seq):
        it = iter(map(tuple, seq.T.tolist()))
        return all(any(s == t for t in it) for s in map(tuple, sub.T.tolist()))

    a = forest.batch(ids, 0.2, 0.35, seed=7)
    a2 = forest.batch(ids, 0.2, 0.35, seed=7)
    c = forest.batch(ids, 0.2, 0.35, seed=8)
    assert torch.equal(a.edge_index, a2.edge_index) and torch.equal(a.BU_edge_index, a2.BU_edge_index)
    assert not torch.equal(a.edge_index, c.edge_index)
    td, bu = a.edge_index.cpu().numpy(), a.BU_edge_index.cpu().numpy()
    for b, i in enumerate(ids):
        e = trees[i].edge_index.shape[1]
        ftd = per_tree(full.edge_index.numpy(), b)
        fbu = per_tree(full.BU_edge_index.numpy()[::-1], b)[::-1]      # BU list = [child; parent]
        ttd = per_tree(td, b)
        tbu = per_tree(bu[::-1], b)[::-1]
        assert ttd.shape[1] == int(e * (1 - 0.2)) and tbu.shape[1] == int(e * (1 - 0.35))
        assert is_subsequence(ttd, ftd) and is_subsequence(tbu, fbu)
    # TD and BU masks of the 301-node tree differ (independent draws)
    kept_td = set(map(tuple, per_tree(td, 2).T.tolist()))
    kept_bu = set(map(tuple, per_tree(bu[::-1], 2).T.tolist()))
    assert kept_td != kept_bu
    # uniformity: 5 edges, keep 4 -> each position dropped ~1/5 of the time
    drops = np.zeros(5)
    full0 = list(map(tuple, trees[0].edge_index.numpy().T.tolist()))
    n_draws = 600
    for s in range(n_draws):
        ei = forest.batch([0], 0.2, 0.0, seed=1000 + s).edge_index.cpu().numpy()
        kept = set(map(tuple, ei.T.tolist()))
        for p, ed in enumerate(full0):
            drops[p] += ed not in kept
    assert drops.sum() == n_draws
    assert np.all(np.abs(drops / n_draws - 0.2) < 0.06), drops


def test_model_on_assembled_batch_matches_oracle(dev):
    """Train-mode step on a device-assembled, DropEdge'd batch: the oracle run on the very same
    kept edges (read back) and the same dropout mask agrees to 1e-5 / 1e-4."""
    import bigcn_b200
    K = 400
    trees = trees_for(k=K, sizes=(70, 33, 120, 9, 250, 64))
    forest = bigcn_b200.DeviceForest.from_data_list(trees, dev)
    ids = [4, 2, 0, 5, 1, 3]
    bd = forest.batch(ids, 0.2, 0.2, seed=3)
    torch.manual_seed(1)
    ref = bigcn_oracle.BiGCN(K, 64, 64).train()
    m = bigcn_b200.BiGCN(K, 64, 64, dev, gemm_mode="sparse").to(dev).train()
    m.load_state_dict(ref.state_dict())
    got = m(bd)
    m.check_inputs()
    host = Batch(x=bd.x.to_dense().cpu(), edge_index=bd.edge_index.cpu(), BU_edge_index=bd.BU_edge_index.cpu(),
                 batch=bd.batch.cpu(), rootindex=bd.rootindex.cpu(), y=bd.y.cpu())
    n = host.x.shape[0]
    s = m.TDrumorGCN.last_seed
    ktd = torch.from_numpy(gcn_oracle.dropout_keep_mask(s, 0, np.arange(n), 64 + K, 0.5))
    kbu = torch.from_numpy(gcn_oracle.dropout_keep_mask(s, 1, np.arange(n), 64 + K, 0.5))
    want = ref(host, keep_td=ktd, keep_bu=kbu)
    err = float((got.detach().cpu().double() - want.detach().double()).abs().max() / want.detach().abs().max())
    assert err < 1e-5
    torch.nn.functional.nll_loss(want, host.y).backward()
    torch.nn.functional.nll_loss(got, bd.y).backward()
    for (name, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        e = float((p.grad.cpu().double() - q.grad.double()).abs().max() / max(float(q.grad.abs().max()), 1e-30))
        assert e < 1e-4, name


def test_eval_counts_match_reference_metric_functions(dev):
    """Device-side confusion counts accumulated over several batches give the tuple the reference's
    evaluation4class / evaluationclass return for the concatenated predictions."""
    from bigcn_b200.metrics import EvalCounts
    from oracle import evaluate_oracle
    torch.manual_seed(0)
    for C, fn in ((4, evaluate_oracle.evaluation4class), (2, evaluate_oracle.evaluationclass)):
        ev = EvalCounts(C, dev)
        preds, ys, nll = [], [], 0.0
        for bsz in (128, 1, 77):
            logp = torch.log_softmax(torch.randn(bsz, C), 1)
            if bsz == 77:
                logp[:5] = logp[:5, :1].expand(5, C)                  # ties: first maximum wins
            y = torch.randint(0, C, (bsz,))
            if C == 4:
                y[y == 3] = 0                                          # a class that never occurs
            ev.update(logp.to(dev), y.to(dev))
            preds += logp.max(dim=-1)[1].tolist()
            ys += y.tolist()
            nll += float(-logp[torch.arange(bsz), y].sum())
        assert ev.result() == fn(preds, ys)
        # what the reference's validation loop logs: the mean over batches of the per-batch tuples
        per, o = [], 0
        for bsz in (128, 1, 77):
            per.append(fn(preds[o:o + bsz], ys[o:o + bsz]))
            o += bsz
        _, vacc, means = ev.epoch_means()
        assert means == tuple(sum(t[i] for t in per) / 3 for i in range(len(per[0])))
        assert abs(vacc - sum(t[0] for t in per) / 3) < 1e-3
        acc, loss = ev.accuracy_and_loss()
        assert abs(acc - sum(p == t for p, t in zip(preds, ys)) / len(ys)) < 1e-12
        assert abs(loss - nll / len(ys)) < 1e-4
        ev.reset()
        assert int(ev.totals.sum().item()) == 0


def test_eval_counts_match_reference_golden(dev):
    """Device counts against tuples produced by the reference's own tools/evaluate.py
    (tests/golden/evaluate_golden.json); predictions enter as log-probs whose argmax is the
    golden prediction, split over two updates."""
    import json
    import os
    from bigcn_b200.metrics import EvalCounts
    with open(os.path.join(os.path.dirname(__file__), "golden", "evaluate_golden.json")) as f:
        cases = json.load(f)["cases"]
    g = torch.Generator().manual_seed(1)
    for c in cases:
        C, n = c["classes"], len(c["y"])
        logit = torch.rand(n, C, generator=g)
        logit[torch.arange(n), torch.tensor(c["pred"])] += 2.0
        logp = torch.log_softmax(logit, 1).to(dev)
        y = torch.tensor(c["y"]).to(dev)
        ev = EvalCounts(C, dev)
        h = n // 2
        ev.update(logp[:h], y[:h])                       # an empty first half (n = 1) is a no-op
        ev.update(logp[h:], y[h:])
        assert list(ev.result()) == c["want"], c["fn"]


def test_unsynchronised_epoch_loop_keeps_every_batch_intact(dev):
    """The epoch loop never synchronises the host (forest.batch, FusedTrainer.step and EvalCounts.update are all
    asynchronous): with the device many batches behind, a staging slot of the per-batch offsets must not be rewritten
    while the copy that reads it is still pending.  Hold the device up with a long kernel queue, assemble many
    batches back to back, and compare each one with the host collate afterwards."""
    import bigcn_b200
    trees = trees_for(k=32, sizes=tuple(int(s) for s in np.random.default_rng(3).integers(1, 400, 40)))
    forest = bigcn_b200.DeviceForest.from_data_list(trees, dev)
    rng = np.random.default_rng(9)
    busy = torch.randn(4096, 4096, device=dev)
    for _ in range(20):                                   # ~100 ms of queued device work ahead of the batches
        busy = busy @ busy * 1e-4
    got, want = [], []
    for i in range(24):                                   # six times the depth of the staging ring
        ids = rng.integers(0, len(trees), int(rng.integers(1, 12))).tolist()
        got.append(forest.batch(ids))
        want.append(collate([trees[j] for j in ids]))
    torch.cuda.synchronize()
    for g, w in zip(got, want):
        assert torch.equal(g.edge_index.cpu(), w.edge_index) and torch.equal(g.batch.cpu(), w.batch)
        assert torch.equal(g.rootindex.cpu(), w.rootindex) and torch.equal(g.x.to_dense().cpu(), w.x)


def test_synthetic_forest_generator_is_a_valid_dataset(dev):
    """data.synth_forest_device (the vectorised generator bench.py uses for the Weibo / power-law / PHEME configurations):
    every tree is a tree (each non-root node has exactly one parent inside its tree), the root sits at root_local, TD
    edges are sorted by (parent, child), bag-of-words rows have distinct ascending columns and counts in 1..3."""
    import bigcn_b200
    from bigcn_b200.data import synth_forest_device, forest_slice_batch
    for shape in ("twitter16", "powerlaw"):
        f = synth_forest_device(shape, 300, dev, seed=4)
        forest = bigcn_b200.DeviceForest.from_device_arrays(f)
        b = forest.batch(list(range(300)))
        n = int(b.batch.numel())
        assert n == int(f["node_ptr"][-1]) and int(b.edge_index.shape[1]) == n - 300
        ei = b.edge_index
        assert bool((b.batch[ei[0]] == b.batch[ei[1]]).all())
        indeg = torch.bincount(ei[1], minlength=n)
        indeg[b.rootindex] += 1
        assert bool((indeg == 1).all())
        key = (b.batch[ei[0]] << 40) | (ei[0] << 20) | ei[1]
        assert bool((key[1:] > key[:-1]).all())
        x = b.x
        rows = torch.repeat_interleave(torch.arange(n, device=dev), (x.ptr[1:] - x.ptr[:-1]).long())
        k2 = rows * 5000 + x.col.long()
        assert bool((k2[1:] > k2[:-1]).all()) and float(x.val.min()) >= 1 and float(x.val.max()) <= 3
    f = synth_forest_device("pheme", 200, dev, seed=5)
    b = forest_slice_batch(f, 17, 60)
    assert b.x.shape[1] == 768 and int(b.rootindex.numel()) == 43 and bool((b.batch[b.rootindex] == torch.arange(43, device=dev)).all())


def _ref_dataset():
    import os
    gold = os.path.join(os.path.dirname(__file__), "golden")
    return os.path.join(gold, "ref_dataset"), np.load(os.path.join(gold, "ref_dataset.npz"))


def _ref_items(ref, tag, n):
    from bigcn_b200.data import Data
    return [Data(**{k: torch.from_numpy(ref[f"{tag}/{j}/{k}"]) for k in ("x", "edge_index", "BU_edge_index", "rootindex", "y")})
            for j in range(n)]


def test_npz_dataset_against_the_reference_dataset_class(dev):
    """tests/golden/ref_dataset.npz: what the reference's own ``BiGraphDataset.__getitem__`` (Process/dataset.py:64-99)
    returned for the ``.npz`` trees under tests/golden/ref_dataset/ (tests/golden/make_ref_dataset_golden.py).
    ``DeviceForest.from_npz_dir`` reads the same files: same id filter, and without DropEdge the assembled batch is the
    PyG collate of the reference's items, bit for bit; with DropEdge 0.2 / 0.3 every tree keeps exactly as many edges
    as the reference kept, as an order-preserving subset, BU drawn from the flipped list independently of TD."""
    import bigcn_b200
    d, ref = _ref_dataset()
    fold_x = [str(i) for i in ref["fold_x"]]
    sizes = {"1001": 9, "1002": 1, "1003": 14, "1004": 5, "1005": 2, "1006": 23, "1007": 7}
    treeDic = {i: {j: {} for j in range(n)} for i, n in sizes.items()}
    forest = bigcn_b200.DeviceForest.from_npz_dir(fold_x, d, dev, treeDic=treeDic, lower=2)
    assert forest.ids == [str(i) for i in ref["nodrop/kept_ids"]] and forest.in_feats == int(ref["K"])
    nt = len(forest.ids)
    items = _ref_items(ref, "nodrop", nt)
    for order in (list(range(nt)), [4, 0, 5, 2]):
        got, want = forest.batch(order), collate([items[i] for i in order])
        for k in ("edge_index", "BU_edge_index", "batch", "rootindex", "y"):
            assert torch.equal(getattr(got, k).cpu(), getattr(want, k)), k
        assert torch.equal(got.x.to_dense().cpu(), want.x)
    dropped = _ref_items(ref, "drop", nt)
    got = forest.batch(list(range(nt)), 0.2, 0.3, seed=11)
    node_off = np.concatenate([[0], np.cumsum([int(t.x.shape[0]) for t in items])])
    for name, flip in (("edge_index", False), ("BU_edge_index", True)):
        e = getattr(got, name).cpu().numpy()
        tree_of_edge = np.searchsorted(node_off, e[0], side="right") - 1
        for t in range(nt):
            mine = e[:, tree_of_edge == t] - node_off[t]
            assert mine.shape[1] == getattr(dropped[t], name).shape[1], (name, t)      # the reference's count
            full = items[t].edge_index.numpy()[::-1] if flip else items[t].edge_index.numpy()
            pos, j = [], 0
            for c in range(mine.shape[1]):      # an order-preserving subset of the tree's own list
                while j < full.shape[1] and not (full[:, j] == mine[:, c]).all():
                    j += 1
                assert j < full.shape[1], (name, t)
                pos.append(j)
                j += 1
    # the reference's own dropped lists are such subsets too (the fixture is what it claims to be)
    for t in range(nt):
        assert dropped[t].edge_index.shape[1] == int(items[t].edge_index.shape[1] * (1 - 0.2))
        assert dropped[t].BU_edge_index.shape[1] == int(items[t].edge_index.shape[1] * (1 - 0.3))


def test_batches_lookahead_loop_equals_plain_loop(dev):
    """DeviceForest.batches hands every step its successor (next_data): the successor's weight-independent half runs
    underneath the step -- same parameters, moments and losses, bit for bit, as assembling and stepping one by one."""
    import bigcn_b200
    trees = trees_for(k=5000, sizes=(30, 1, 200, 17, 90, 5, 2, 350, 44, 8, 120, 60))
    forest = bigcn_b200.DeviceForest.from_data_list(trees, dev)
    rng = np.random.default_rng(2)
    lists = [rng.choice(len(trees), 6, replace=False) for _ in range(7)]
    seeds = list(range(50, 57))
    res = []
    for look in (False, True):
        torch.manual_seed(3)
        m = bigcn_b200.BiGCN(5000, 64, 64, dev, gemm_mode="sparse", validate="off").to(dev).train()
        tr = bigcn_b200.FusedTrainer(m)
        losses = []
        if look:
            seen = 0
            for b, nxt in forest.batches(lists, 0.2, 0.2, seeds):
                losses.append(tr.step(b, next_data=nxt).clone())
                seen += 1
                assert (nxt is None) == (seen == len(lists))
        else:
            for ids, sd in zip(lists, seeds):
                losses.append(tr.step(forest.batch(ids, 0.2, 0.2, seed=sd)).clone())
        tr.check_inputs()
        res.append((tr.flat.clone(), tr.exp_avg.clone(), torch.cat(losses)))
    for a, b in zip(*res):
        assert torch.equal(a, b)
    assert list(forest.batches([])) == []
