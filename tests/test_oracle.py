"""CPU tests: the oracle against the golden vectors and its own invariants."""
import copy
import json
import os

import numpy as np
import pytest
import torch

from oracle import bigcn_oracle, gcn_oracle
from bigcn_b200.data import Data, collate, make_batch, drop_edge, make_tree, tree_sizes

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def unhex(lst):
    return np.array([float.fromhex(v) for v in lst], dtype=np.float32)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert [int(v) for v in gcn_oracle.philox4x32_10(0, 0, 0, 0, 0, 0)] == \
        [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert [int(v) for v in gcn_oracle.philox4x32_10(f, f, f, f, f, f)] == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert [int(v) for v in gcn_oracle.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344,
                                                      0xa4093822, 0x299f31d0)] == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_dropout_mask_rate_and_determinism():
    m1 = gcn_oracle.dropout_keep_mask(123, 0, np.arange(200), 97, 0.5)
    m2 = gcn_oracle.dropout_keep_mask(123, 0, np.arange(200), 97, 0.5)
    m3 = gcn_oracle.dropout_keep_mask(123, 1, np.arange(200), 97, 0.5)
    assert m1.shape == (200, 97) and (m1 == m2).all() and (m1 != m3).any()
    assert abs(m1.mean() - 0.5) < 0.02
    # keyed on global node id: a shard sees the same bits
    assert (gcn_oracle.dropout_keep_mask(123, 0, np.arange(50, 80), 97, 0.5) == m1[50:80]).all()
    assert gcn_oracle.dropout_threshold(0.5) == 0x80000000


def test_kat_tree5_survey_8c():
    """The hand-checkable vector of SURVEY.md 8c: exact fp32 bit patterns."""
    kat = json.load(open(os.path.join(GOLD, "kat_tree5.json")))
    ei = torch.tensor([[0, 0, 1, 1], [1, 2, 3, 4]])
    for name, e in (("TD", ei), ("BU", ei.flip(0))):
        for deg_by in ("target", "source"):
            k = kat[f"{name}_{deg_by}"]
            e2, w = gcn_oracle.gcn_norm(e, 5, deg_by=deg_by)
            assert e2.tolist() == k["coo"]
            assert [float(v).hex() for v in w] == k["w"]
    td = kat["TD_target"]
    assert td["deg"] == [1, 2, 2, 2, 2]
    assert td["w"][0] == float.fromhex("0x1.6a09e6p-1").hex() and td["w"][2] == float.fromhex("0x1.fffffep-2").hex()
    assert unhex(td["rowsum"]).tolist() == pytest.approx([1.0, 1.2071067, 1.2071067, 0.99999994, 0.99999994], abs=1e-7)
    bu = kat["BU_target"]
    assert bu["deg"] == [3, 3, 1, 1, 1]
    assert bu["w"][0] == float.fromhex("0x1.555554p-2").hex() and bu["w"][1] == float.fromhex("0x1.279a74p-1").hex()
    # 1.3.2 convention swaps the degree vectors
    assert kat["TD_source"]["deg"] == [3, 3, 1, 1, 1] and kat["BU_source"]["deg"] == [1, 2, 2, 2, 2]


def test_deg_inv_sqrt_is_ieee_div_sqrt():
    d = torch.arange(1, 20001, dtype=torch.float32)
    a = d.pow(-0.5).numpy()
    b = (np.float32(1.0) / np.sqrt(d.numpy())).astype(np.float32)
    assert (a.view(np.uint32) == b.view(np.uint32)).all()


def test_graph_prep_matches_gcn_norm_on_random_graphs():
    rng = np.random.default_rng(0)
    for n, e in ((1, 0), (7, 0), (10, 25), (60, 200)):
        ei = rng.integers(0, n, (2, e))
        for deg_by in ("target", "source"):
            g = gcn_oracle.graph_prep(ei, n, np.zeros(n, np.int64), 1, deg_by)
            e2, w = gcn_oracle.gcn_norm(torch.from_numpy(ei), n, deg_by=deg_by)
            h = torch.randn(n, 3)
            want = gcn_oracle.propagate_sum(h, e2, w).numpy()
            got = np.zeros((n, 3), np.float32)
            for i in range(n):
                acc = np.zeros(3, np.float32)
                for j in g["in_idx"][g["in_ptr"][i]:g["in_ptr"][i + 1]]:
                    acc = acc + np.float32(g["dis"][j] * g["dis"][i]) * h[j].numpy()
                acc = acc + np.float32(g["dis"][i] * g["dis"][i]) * h[i].numpy()
                got[i] = acc
            assert (got.view(np.uint32) == want.view(np.uint32)).all()
            # rowsum == A-hat applied to ones, same order
            ones = gcn_oracle.propagate_sum(torch.ones(n, 1), e2, w).numpy()[:, 0]
            assert (ones.view(np.uint32) == g["rowsum"].view(np.uint32)).all()
            # out-CSR is the transpose of in-CSR
            a = np.zeros((n, n), np.int64); b = np.zeros((n, n), np.int64)
            for i in range(n):
                for j in g["in_idx"][g["in_ptr"][i]:g["in_ptr"][i + 1]]:
                    a[i, j] += 1
                for j in g["out_idx"][g["out_ptr"][i]:g["out_ptr"][i + 1]]:
                    b[j, i] += 1
            assert (a == b).all()


def test_golden_small_batch_reproduces():
    gold = json.load(open(os.path.join(GOLD, "bigcn_small.json")))
    b = Data(x=torch.from_numpy(unhex(gold["x"]).reshape(gold["N"], gold["K"])),
             edge_index=torch.tensor(gold["edge_index"]).reshape(2, -1),
             BU_edge_index=torch.tensor(gold["BU_edge_index"]).reshape(2, -1),
             batch=torch.tensor(gold["batch"]), rootindex=torch.tensor(gold["rootindex"]))
    m = bigcn_oracle.BiGCN(gold["K"], 64, 64, num_classes=gold["C"]).eval()
    m.load_state_dict({k: torch.from_numpy(unhex(v["data"]).reshape(v["shape"])) for k, v in gold["state"].items()})
    with torch.no_grad():
        logp = m(b).numpy()
    np.testing.assert_allclose(logp.ravel(), unhex(gold["logp"]), rtol=0, atol=2e-6)
    g = gcn_oracle.graph_prep(b.edge_index.numpy(), gold["N"], b.batch.numpy(), 5)
    assert g["in_idx"].tolist() == gold["graph_td"]["in_idx"]
    assert g["deg"].tolist() == gold["graph_td"]["deg"]


def _wiring_case(name):
    """tests/golden/ref_wiring.npz: the reference's own model files run over dense stand-ins of the two absent wheels
    (tests/golden/make_ref_wiring_golden.py) on the bigcn_small.json batch."""
    ref = np.load(os.path.join(GOLD, "ref_wiring.npz"))
    gold = json.load(open(os.path.join(GOLD, "bigcn_small.json")))
    b = Data(x=torch.from_numpy(unhex(gold["x"]).reshape(gold["N"], gold["K"])),
             edge_index=torch.tensor(gold["edge_index"]).reshape(2, -1),
             BU_edge_index=torch.tensor(gold["BU_edge_index"]).reshape(2, -1),
             batch=torch.tensor(gold["batch"]), rootindex=torch.tensor(gold["rootindex"]),
             y=torch.from_numpy(ref[f"{name}/y"]))
    state = {k[len(name) + 7:]: torch.from_numpy(ref[k]) for k in ref.files if k.startswith(f"{name}/state/")}
    return ref, gold, b, state


@pytest.mark.parametrize("name,hid,cls", [("twitter", 64, bigcn_oracle.BiGCN), ("weibo", 16, bigcn_oracle.Net)])
@pytest.mark.parametrize("loops", [False, True])
def test_oracle_matches_reference_model_code(name, hid, cls, loops):
    """The restated modules against the outputs of BiGCN_Twitter.py:19-131 / BiGCN_Weibo.py:16-89 themselves: eval and
    training mode (the same torch generator state feeds F.dropout: TD's mask is drawn first, then BU's), loss and all
    ten gradients.  Differences left: summation order inside GCNConv (dense matrix there, index_add_ here)."""
    ref, gold, b, state = _wiring_case(name)
    m = cls(gold["K"], hid, hid, reference_loops=loops)
    m.load_state_dict(state)
    for mode in ("eval", "train"):
        m.train(mode == "train")
        m.zero_grad()
        torch.manual_seed(int(ref[f"{name}/{mode}/dropout_seed"].item()))
        logp = m(b)
        loss = torch.nn.functional.nll_loss(logp, b.y)
        loss.backward()
        np.testing.assert_allclose(logp.detach().numpy(), ref[f"{name}/{mode}/logp"], rtol=0, atol=2e-6)
        assert abs(float(loss.detach()) - float(ref[f"{name}/{mode}/loss"].item())) <= 2e-6
        for k, p in m.named_parameters():
            want = ref[f"{name}/{mode}/grad/{k}"]
            np.testing.assert_allclose(p.grad.numpy(), want, rtol=0, atol=1e-6 + 1e-5 * float(np.abs(want).max()),
                                       err_msg=f"{name} {mode} {k}")


def test_reference_loops_equal_vectorised_form():
    b = make_batch("twitter15", 3, seed=3, train=True, in_feats=64)
    torch.manual_seed(0)
    m = bigcn_oracle.BiGCN(64, 64, 64).eval()
    m2 = bigcn_oracle.BiGCN(64, 64, 64, reference_loops=True).eval()
    m2.load_state_dict(m.state_dict())
    assert torch.equal(m(b), m2(b))


def test_copy_copy_detaches_second_root_extend():
    """BiGCN_Twitter.py:44: copy.copy(non-leaf) is a detached leaf sharing storage, so no
    gradient flows from the readout's root half into conv1."""
    x = torch.randn(4, 3, requires_grad=True) * 2.0
    y = copy.copy(x)
    assert y.grad_fn is None and y.is_leaf
    b = make_batch("twitter15", 2, seed=5, train=False, in_feats=32)
    torch.manual_seed(1)
    m = bigcn_oracle.TDrumorGCN(32, 64, 64).eval()
    out = m(b)
    out[:, 64:].sum().backward()                    # only the root-extend half
    assert m.conv1.lin.weight.grad is None or float(m.conv1.lin.weight.grad.abs().max()) == 0.0


def test_collate_offsets_and_dropedge():
    rng = np.random.default_rng(0)
    t1, t2 = make_tree("twitter15", 60, rng, in_feats=8), make_tree("twitter15", 70, rng, in_feats=8)
    b = collate([t1, t2])
    assert b.x.shape == (130, 8) and b.batch.tolist() == [0] * 60 + [1] * 70
    assert b.rootindex.tolist() == [int(t1.rootindex), 60 + int(t2.rootindex)]
    assert torch.equal(b.edge_index[:, 59:], t2.edge_index + 60)
    assert torch.equal(b.BU_edge_index, b.edge_index.flip(0))
    d = drop_edge(t1, 0.2, 0.2, rng)
    assert d.edge_index.shape[1] == int(59 * 0.8) and d.BU_edge_index.shape[1] == int(59 * 0.8)
    # every tree node but the root has exactly one parent; edges sorted by (parent, child)
    ei = t1.edge_index.numpy()
    assert sorted(ei[1].tolist() + [int(t1.rootindex)]) == list(range(60))
    assert (np.lexsort((ei[1], ei[0])) == np.arange(59)).all()


def test_tree_size_distributions():
    rng = np.random.default_rng(0)
    s = tree_sizes("twitter15", 4000, rng)
    assert s.min() >= 55 and s.max() <= 1768 and 180 < s.mean() < 260
    p = tree_sizes("pheme", 4000, rng)
    assert p.min() == 1 and p.max() <= 108 and 0.15 < (p == 1).mean() < 0.27
    w = tree_sizes("weibo", 4000, rng)
    assert w.min() >= 10 and 500 < w.mean() < 1200


def test_evaluate_oracle_hand_checked():
    """tools/evaluate.py restatement on a case small enough to check by hand."""
    from oracle import evaluate_oracle
    y = [0, 0, 1, 1, 1, 2]
    pred = [0, 1, 1, 1, 0, 3]
    r = evaluate_oracle.evaluation4class(pred, y)
    assert r[0] == round(3 / 6, 4)
    # class 1 (label 0): TP 1, FN 1, FP 1, TN 3
    assert r[1:5] == (round(4 / 6, 4), 0.5, 0.5, 0.5)
    # class 2 (label 1): TP 2, FN 1, FP 1, TN 2
    assert r[5:9] == (round(4 / 6, 4), round(2 / 3, 4), round(2 / 3, 4), round(2 * 0.6667 * 0.6667 / (0.6667 + 0.6667), 4))
    # class 4 (label 3): never true, predicted once: precision 0, recall 0 (no positives), F 0
    assert r[13:17] == (round(5 / 6, 4), 0.0, 0, 0)
    assert evaluate_oracle.evaluationclass([0, 1, 1], [0, 1, 0])[0] == round(2 / 3, 4)


def test_evaluate_oracle_matches_reference_golden():
    """The golden tuples were produced by the reference's own tools/evaluate.py
    (tests/golden/make_evaluate_golden.py): this pins the restatement."""
    import json
    import os
    from oracle import evaluate_oracle
    with open(os.path.join(os.path.dirname(__file__), "golden", "evaluate_golden.json")) as f:
        cases = json.load(f)["cases"]
    assert len(cases) >= 12
    for c in cases:
        got = getattr(evaluate_oracle, c["fn"])(c["pred"], c["y"])
        assert list(got) == c["want"], c["fn"]


def test_gcnconv_matches_dense_normalised_adjacency():
    """An independent derivation of what GCNConv computes -- the published formula
    out = D^-1/2 (A + I) D^-1/2 X W^T + b with A[target, source] = #edges source -> target (self-loops of the
    input dropped, one unit loop per node), D = row sums, as dense fp64 matrices (scipy) -- against the
    oracle's edge-list restatement (add_remaining_self_loops / scatter_add degree / index_add_ aggregation,
    the PyG formulation).  Not a pin on PyG itself (it is not installable here, oracle/__init__.py), but the two
    forms share no code: duplicate edges, pre-existing self-loops, isolated nodes and both directions of a
    reply tree are covered, and so is scatter_mean against a dense averaging matrix."""
    import scipy.sparse as sp
    rng = np.random.default_rng(7)
    cases = []
    for n, e in ((1, 0), (6, 0), (12, 30), (80, 300)):
        cases.append((n, rng.integers(0, n, (2, e))))
    tree = np.array([[0, 0, 1, 1, 3, 3, 3], [1, 2, 3, 4, 5, 6, 7]])
    cases += [(8, tree), (8, tree[::-1].copy())]                 # TD [parent; child] and BU [child; parent]
    for n, ei in cases:
        x = rng.standard_normal((n, 5))
        conv = gcn_oracle.GCNConv(5, 4).double()
        with torch.no_grad():
            conv.bias.uniform_(-1, 1)
        got = conv(torch.from_numpy(x), torch.from_numpy(ei)).detach().numpy()
        keep = ei[0] != ei[1]
        a = sp.coo_matrix((np.ones(int(keep.sum())), (ei[1][keep], ei[0][keep])), shape=(n, n)).toarray() + np.eye(n)
        d = a.sum(1)                                                # degree by target incl. the loop (PyG 2.x)
        a_hat = a / np.sqrt(d)[:, None] / np.sqrt(d)[None, :]
        want = a_hat @ (x @ conv.lin.weight.detach().numpy().T) + conv.bias.detach().numpy()
        assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
    # scatter_mean == dense averaging matrix (empty trees give zero rows)
    batch = np.array([0, 0, 0, 2, 2, 3])
    src = rng.standard_normal((6, 3))
    m = np.zeros((5, 6))
    for i, b in enumerate(batch):
        m[b, i] = 1.0
    cnt = np.maximum(m.sum(1, keepdims=True), 1.0)
    got = gcn_oracle.scatter_mean(torch.from_numpy(src), torch.from_numpy(batch), 5).numpy()
    assert np.abs(got - (m / cnt) @ src).max() <= 1e-14


def test_reference_dataset_fixture_contract():
    """tests/golden/ref_dataset.npz (outputs of the reference's BiGraphDataset, Process/dataset.py:45-99): the input
    contract bigcn_b200.data restates -- BU = flipped TD list, DropEdge keeps int(e * (1 - rate)) positions as an
    order-preserving subset per direction, the id filter drops single-post trees and unknown ids -- and the host-side
    ``drop_edge`` follows the same count rule."""
    ref = np.load(os.path.join(GOLD, "ref_dataset.npz"))
    kept = [str(i) for i in ref["nodrop/kept_ids"]]
    assert kept == ["1001", "1003", "1004", "1005", "1006", "1007"]
    rng = np.random.default_rng(0)
    for j in range(len(kept)):
        ei, bu = ref[f"nodrop/{j}/edge_index"], ref[f"nodrop/{j}/BU_edge_index"]
        assert (bu == ei[::-1]).all() and ref[f"nodrop/{j}/x"].dtype == np.float32
        z = np.load(os.path.join(GOLD, "ref_dataset", kept[j] + ".npz"))
        assert (z["edgeindex"] == ei).all() and int(z["rootindex"]) == int(ref[f"nodrop/{j}/rootindex"][0])
        d_td, d_bu = ref[f"drop/{j}/edge_index"], ref[f"drop/{j}/BU_edge_index"]
        e = ei.shape[1]
        assert d_td.shape[1] == int(e * (1 - 0.2)) and d_bu.shape[1] == int(e * (1 - 0.3))
        for sub, full in ((d_td, ei), (d_bu, bu)):
            pos = 0
            for c in range(sub.shape[1]):
                while not (full[:, pos] == sub[:, c]).all():
                    pos += 1
                pos += 1
        mine = drop_edge(Data(x=torch.from_numpy(ref[f"nodrop/{j}/x"]), edge_index=torch.from_numpy(ei.copy()),
                              BU_edge_index=torch.from_numpy(bu.copy()), rootindex=torch.tensor([0]), y=torch.tensor([0])),
                         0.2, 0.3, rng)
        assert mine.edge_index.shape[1] == d_td.shape[1] and mine.BU_edge_index.shape[1] == d_bu.shape[1]
