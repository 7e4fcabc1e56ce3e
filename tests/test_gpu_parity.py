"""GPU parity tests: the CUDA path (through the C-ABI in libbigcn_b200.so) against the CPU
oracle on the same seeded inputs.  Bars: graph prep / propagate / dropout mask bit-exact;
fp32 log-probs max|d| <= 1e-5 * max|ref| (BASELINE.json north_star); gradients
max|d| <= 1e-4 * max|ref| per tensor."""
import copy
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

from oracle import bigcn_oracle, gcn_oracle
from bigcn_b200.data import Data, Batch, collate, make_batch, make_tree, drop_edge

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
LOGP_TOL = 1e-5      # north_star: fp32 logits within 1e-5 relative
GRAD_TOL = 1e-4


def unhex(lst):
    return np.array([float.fromhex(v) for v in lst], dtype=np.float32)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def rel_err(got, want):
    want = want.detach().cpu().double()
    got = got.detach().cpu().double()
    scale = max(float(want.abs().max()), 1e-30)
    return float((got - want).abs().max()) / scale


def assert_logp_parity(got, ref, b, keep_td=None, keep_bu=None, what=""):
    """north_star bar: fp32 log-probs within 1e-5 (max|d| / max|ref|).  Checked against the
    fp64 oracle (truth) outright, and against the fp32 oracle with the fp32 oracle's own
    distance from the truth as allowance (its sequential index_add_ sums drift by ~1e-5 on
    trees of thousands of nodes; triangle inequality)."""
    want32 = ref(b, keep_td, keep_bu).detach()
    ref64 = copy.deepcopy(ref).double()
    want64 = ref64(b, keep_td, keep_bu).detach()
    own = rel_err(want32, want64)
    e64, e32 = rel_err(got, want64), rel_err(got, want32)
    assert e64 < LOGP_TOL, f"{what}: vs fp64 oracle {e64:.3e}"
    assert e32 < LOGP_TOL + own, f"{what}: vs fp32 oracle {e32:.3e} (oracle fp32-vs-fp64 {own:.3e})"
    assert torch.equal(got.argmax(1).cpu(), want32.argmax(1)), what
    return want32


def clone_batch(b, dev):
    return Batch(x=b.x.to(dev), edge_index=b.edge_index.to(dev), BU_edge_index=b.BU_edge_index.to(dev),
                 batch=b.batch.to(dev), rootindex=b.rootindex.to(dev), y=b.y.to(dev))


def edge_cases():
    """The shapes the reference's data actually produces."""
    rng = np.random.default_rng(42)
    K = 40
    cases = {}
    cases["twitter_train_dropedge"] = make_batch("twitter15", 5, seed=1, train=True, in_feats=K)
    cases["twitter_eval"] = make_batch("twitter16", 4, seed=2, train=False, in_feats=K)
    cases["pheme_singletons"] = make_batch("pheme", 24, seed=3, train=True, in_feats=K)
    cases["single_tree"] = make_batch("twitter15", 1, seed=4, train=False, in_feats=K)
    # all single-node trees: E = 0
    cases["no_edges"] = collate([make_tree("pheme", 1, rng, in_feats=K) for _ in range(3)])
    # star with a very high-degree root + deep chain
    n = 700
    star = make_tree("twitter15", n, rng, in_feats=K)
    ei = np.stack([np.zeros(n - 1, np.int64), np.arange(1, n)])
    star.edge_index = torch.from_numpy(ei); star.BU_edge_index = torch.from_numpy(ei[::-1].copy())
    star.rootindex = torch.tensor([0])
    chain = make_tree("twitter15", 90, rng, in_feats=K)
    ce = np.stack([np.arange(0, 89), np.arange(1, 90)])
    chain.edge_index = torch.from_numpy(ce); chain.BU_edge_index = torch.from_numpy(ce[::-1].copy())
    chain.rootindex = torch.tensor([0])
    cases["star_and_chain"] = collate([star, chain])
    # general graph: shuffled edge order, duplicates, self-loops, multi-parent
    g = make_tree("twitter15", 80, rng, in_feats=K)
    e = g.edge_index.numpy()
    extra = rng.integers(0, 80, (2, 40))
    loops = np.stack([np.arange(0, 80, 9), np.arange(0, 80, 9)])
    e = np.concatenate([e, extra, loops, e[:, :10]], 1)
    e = e[:, rng.permutation(e.shape[1])]
    g.edge_index = torch.from_numpy(e.copy())
    g.BU_edge_index = torch.from_numpy(e[::-1][:, rng.permutation(e.shape[1])].copy())
    cases["general_graph"] = collate([g, make_tree("twitter15", 60, rng, in_feats=K)])
    return K, cases


# ------------------------------------------------------------------------------- graph prep
def check_graph_prep(b, dev, deg_by="target"):
    from bigcn_b200 import ops
    n, nb = b.x.shape[0], int(b.rootindex.numel())
    graphs, node_ptr, flags = ops.graph_prep([b.edge_index.to(dev), b.BU_edge_index.to(dev)], n,
                                             b.batch.to(dev), nb, deg_by)
    assert int(flags.item()) == 0
    for g, ei in zip(graphs, (b.edge_index, b.BU_edge_index)):
        want = gcn_oracle.graph_prep(ei.numpy(), n, b.batch.numpy(), nb, deg_by)
        ne = want["n_edges"]
        assert g["in_ptr"].cpu().numpy().tolist() == want["in_ptr"].tolist()
        assert g["out_ptr"].cpu().numpy().tolist() == want["out_ptr"].tolist()
        assert g["in_idx"].cpu().numpy()[:ne].tolist() == want["in_idx"].tolist()
        assert g["out_idx"].cpu().numpy()[:ne].tolist() == want["out_idx"].tolist()
        if n:
            assert g["deg"].cpu().numpy()[:n].tolist() == want["deg"].tolist()
            assert (bits(g["dis"].cpu().numpy()[:n]) == bits(want["dis"])).all()
            assert (bits(g["rowsum"].cpu().numpy()[:n]) == bits(want["rowsum"])).all()
        assert node_ptr.cpu().numpy().tolist() == want["node_ptr"].tolist()


def test_graph_prep_bit_exact_edge_cases(dev):
    _, cases = edge_cases()
    for name, b in cases.items():
        for deg_by in ("target", "source"):
            check_graph_prep(b, dev, deg_by)


def test_graph_prep_kat_tree5(dev):
    from bigcn_b200 import ops
    kat = json.load(open(os.path.join(GOLD, "kat_tree5.json")))
    ei = torch.tensor([[0, 0, 1, 1], [1, 2, 3, 4]], device=dev)
    for deg_by in ("target", "source"):
        graphs, node_ptr, _ = ops.graph_prep([ei, ei.flip(0).contiguous()], 5,
                                             torch.zeros(5, dtype=torch.int64, device=dev), 1, deg_by)
        for g, name in zip(graphs, ("TD", "BU")):
            k = kat[f"{name}_{deg_by}"]
            assert g["deg"].cpu().tolist() == k["deg"]
            assert [float(v).hex() for v in g["dis"].cpu()] == k["dis"]
            assert [float(v).hex() for v in g["rowsum"].cpu()] == k["rowsum"]
            assert g["in_ptr"].cpu().tolist() == k["in_ptr"] and g["in_idx"].cpu().tolist()[:4] == k["in_idx"]
            assert g["out_ptr"].cpu().tolist() == k["out_ptr"] and g["out_idx"].cpu().tolist()[:4] == k["out_idx"]
        assert node_ptr.cpu().tolist() == [0, 5]


def test_graph_prep_full_size_batch_and_large_n(dev):
    """BASELINE config sizes: a 128-tree Twitter15-shaped batch (3 radix passes are not needed,
    N < 65536) and a 300k-node forest (3 passes, multi-block scan)."""
    b = make_batch("twitter15", 128, seed=0, train=True, in_feats=8)
    check_graph_prep(b, dev)
    from bigcn_b200 import ops
    from bigcn_b200.data import make_device_forest
    f = make_device_forest(3000, 100, dev, seed=1)
    n = 300000
    graphs, node_ptr, flags = ops.graph_prep([f.edge_index, f.BU_edge_index], n, f.batch, 3000)
    assert int(flags.item()) == 0
    want = gcn_oracle.graph_prep(f.edge_index.cpu().numpy(), n, f.batch.cpu().numpy(), 3000)
    g = graphs[0]
    assert torch.equal(g["in_ptr"].cpu(), torch.from_numpy(want["in_ptr"]))
    assert torch.equal(g["in_idx"].cpu()[:want["n_edges"]], torch.from_numpy(want["in_idx"]))
    assert torch.equal(g["out_idx"].cpu()[:want["n_edges"]], torch.from_numpy(want["out_idx"]))
    assert (bits(g["dis"].cpu().numpy()) == bits(want["dis"])).all()
    # size-independent property: the BU structure is the transpose of the TD structure
    assert torch.equal(graphs[1]["in_ptr"], g["out_ptr"]) and torch.equal(graphs[1]["out_ptr"], g["in_ptr"])
    assert torch.equal(node_ptr.cpu(), torch.arange(0, n + 1, 100, dtype=torch.int32))


def test_invalid_inputs_raise_flags(dev):
    import bigcn_b200
    from bigcn_b200 import ops
    ei = torch.tensor([[0, 1, 7], [1, 2, 0]], device=dev)
    _, _, flags = ops.graph_prep([ei], 3, torch.tensor([0, 1, 0], device=dev), 2)
    v = int(flags.item())
    assert v & 1 and v & 2
    with pytest.raises(IndexError):
        ops.raise_on_flags(flags)
    b = make_batch("twitter15", 2, seed=0, train=False, in_feats=16)
    b.rootindex = torch.tensor([0, 10 ** 6])
    m = bigcn_b200.BiGCN(16, 64, 64, dev, validate="sync").to(dev).eval()
    with pytest.raises(IndexError):
        m(clone_batch(b, dev))


# ------------------------------------------------------------------------------- dropout mask
def test_dropout_mask_matches_philox_spec(dev):
    from bigcn_b200 import ops
    for seed, stream, base, n, cols, p in ((0, 0, 0, 37, 64 + 40, 0.5), (2 ** 40 + 12345, 1, 2 ** 33 + 5, 19, 131, 0.5),
                                           (99, 1, 7, 50, 70, 0.25)):
        got = ops.dropout_mask(seed, stream, base, n, cols, p, dev).cpu().numpy().astype(bool)
        want = gcn_oracle.dropout_keep_mask(seed, stream, base + np.arange(n), cols, p)
        assert (got == want).all()


# ------------------------------------------------------------------------------- X W and propagate
def test_xw_matches_fp32_reference(dev):
    from bigcn_b200 import ops
    torch.manual_seed(0)
    b = make_batch("twitter15", 6, seed=9, train=False)           # K = 5000 BoW
    w_td, w_bu = torch.randn(64, 5000) * 0.02, torch.randn(64, 5000) * 0.02
    want = b.x.double() @ torch.cat([w_td, w_bu]).double().t()
    got = ops.xw(b.x.to(dev), [w_td.to(dev), w_bu.to(dev)])
    assert rel_err(got, want) < 2e-6
    got1 = ops.xw(b.x.to(dev), [w_bu.to(dev)])
    assert rel_err(got1, want[:, 64:]) < 2e-6
    # dense signed features, K not a multiple of 4 (scalar path), empty input
    x = torch.tanh(torch.randn(300, 771))
    w = torch.randn(64, 771) * 0.05
    assert rel_err(ops.xw(x.to(dev), [w.to(dev)]), x.double() @ w.double().t()) < 2e-6
    assert ops.xw(torch.zeros(0, 64, device=dev), [torch.zeros(64, 64, device=dev)]).shape == (0, 64)


def test_propagate_bit_exact_and_transpose(dev):
    """Rows are summed in COO' order with separate multiply and add: bit-identical to the CPU
    index_add_ of the oracle.  Hub rows (more than BIGCN_LONG_ROW = 32 in-edges) go through the
    deterministic split-row path when the hub lists are present: same value to fp32 rounding
    of a different summation order, bit-identical from run to run; without the lists
    (long_rows=False) every row is walked sequentially and matches bit for bit."""
    from bigcn_b200 import ops
    _, cases = edge_cases()
    torch.manual_seed(1)
    for name, b in cases.items():
        n = b.x.shape[0]
        for ei in (b.edge_index, b.BU_edge_index):
            h = torch.randn(n, 64)
            bias = torch.randn(64)
            e2, w = gcn_oracle.gcn_norm(ei, n)
            want = gcn_oracle.propagate_sum(h, e2, w) + bias
            g_seq, _, _ = ops.graph_prep([ei.to(dev)], n, long_rows=False)
            got = ops.propagate(g_seq[0], h.to(dev), bias.to(dev)).cpu()
            assert (bits(got.numpy()) == bits(want.numpy())).all(), name
            graphs, _, _ = ops.graph_prep([ei.to(dev)], n)
            got_l = ops.propagate(graphs[0], h.to(dev), bias.to(dev)).cpu()
            indeg = np.diff(graphs[0]["in_ptr"].cpu().numpy())
            short = torch.from_numpy(indeg <= 32)
            assert (bits(got_l[short].numpy()) == bits(want[short].numpy())).all(), name
            if (~short).any():
                assert rel_err(got_l[~short], want[~short]) < 2e-6, name
                again = ops.propagate(graphs[0], h.to(dev), bias.to(dev)).cpu()
                assert torch.equal(again, got_l), name            # deterministic
            # A-hat^T: linearity / adjoint property  <A h, g> == <h, A^T g>
            gq = torch.randn(n, 64)
            at_g = ops.propagate(graphs[0], gq.to(dev), transpose=True).cpu().double()
            a_h = gcn_oracle.propagate_sum(h.double(), e2, w.double())
            lhs, rhs = float((a_h * gq.double()).sum()), float((h.double() * at_g).sum())
            assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)
            relu = ops.propagate(graphs[0], h.to(dev), bias.to(dev), relu=True).cpu()
            assert torch.equal(relu, torch.relu(got_l))


def test_propagate_hub_rows_split_path(dev):
    """Weibo-shaped skew: one root with 20,000 children (64 chunks), mid-size hubs just above and
    below the split threshold, many small rows around them; both CSR orientations."""
    from bigcn_b200 import ops
    rng = np.random.default_rng(5)
    n = 24000
    parent = np.zeros(n, np.int64)
    parent[1:20001] = 0                                   # the root's 20,000 children
    parent[20001:20034] = 1                               # 33 children: just above the threshold
    parent[20034:20066] = 2                               # 32 children: stays on the sequential path
    parent[20066:20400] = 3                               # 334 children
    parent[20400:] = rng.integers(4, 20000, n - 20400)    # the rest hang off random nodes
    child = np.arange(1, n)
    ei = torch.from_numpy(np.stack([parent[1:], child]))
    perm = torch.from_numpy(rng.permutation(n))           # arbitrary node numbering
    ei = perm[ei]
    torch.manual_seed(3)
    h = torch.randn(n, 64)
    for e in (ei, ei.flip(0).contiguous()):
        graphs, _, flags = ops.graph_prep([e.to(dev)], n)
        assert int(flags.item()) == 0
        e2, w = gcn_oracle.gcn_norm(e, n)
        want = gcn_oracle.propagate_sum(h.double(), e2, w.double())
        for transpose in (False, True):
            if transpose:
                e2t = e2.flip(0)
                want_t = gcn_oracle.propagate_sum(h.double(), e2t, w.double())
            got = ops.propagate(graphs[0], h.to(dev), transpose=transpose).cpu()
            ref = want_t if transpose else want
            assert rel_err(got, ref) < 2e-6
            assert torch.equal(got, ops.propagate(graphs[0], h.to(dev), transpose=transpose).cpu())


def test_readout_matches_scatter_mean(dev):
    """scatter_mean + second root-extend (BiGCN_Twitter.py:58-65) on ragged trees: single nodes,
    trees inside one 512-row slice, trees straddling slices, a 70k-node Weibo-sized tree."""
    from bigcn_b200 import ops
    torch.manual_seed(4)
    sizes = [1, 1, 3, 600, 5000, 2, 70000, 1, 511, 513, 1024, 7]
    n, nb = sum(sizes), len(sizes)
    batch = torch.repeat_interleave(torch.arange(nb), torch.tensor(sizes))
    node_ptr = torch.zeros(nb + 1, dtype=torch.int32)
    node_ptr[1:] = torch.cumsum(torch.tensor(sizes), 0)
    root = torch.tensor([int(node_ptr[b]) + (sizes[b] * 7) // 11 for b in range(nb)])
    h2 = torch.relu(torch.randn(n, 64))
    h1 = torch.randn(n, 64)
    feat, pos = ops.readout(h2.to(dev), h1.to(dev), node_ptr.to(dev), root.to(dev), want_pos=True)
    want = gcn_oracle.scatter_mean(h2.double(), batch, nb)
    assert rel_err(feat[:, :64], want) < 1e-6
    assert torch.equal(feat[:, 64:].cpu(), h1[root])
    want_pos = torch.zeros(nb, 64, dtype=torch.float64).index_add_(0, batch, (h2 > 0).double())
    assert torch.equal(pos.cpu().double(), want_pos)
    again = ops.readout(h2.to(dev), h1.to(dev), node_ptr.to(dev), root.to(dev))
    assert torch.equal(again, feat)                                 # deterministic


def test_gcnconv_forward_backward(dev):
    import bigcn_b200
    torch.manual_seed(2)
    K, cases = edge_cases()
    for name in ("twitter_train_dropedge", "general_graph", "no_edges"):
        b = cases[name]
        ref = gcn_oracle.GCNConv(K, 64)
        with torch.no_grad():
            ref.bias.uniform_(-0.1, 0.1)
        conv = bigcn_b200.GCNConv(K, 64).to(dev)
        conv.load_state_dict(ref.state_dict())
        want = ref(b.x, b.BU_edge_index)
        got = conv(b.x.to(dev), b.BU_edge_index.to(dev))
        assert rel_err(got, want) < LOGP_TOL
        g = torch.randn_like(want)
        want.backward(g)
        got.backward(g.to(dev))
        assert rel_err(conv.lin.weight.grad, ref.lin.weight.grad) < GRAD_TOL
        assert rel_err(conv.bias.grad, ref.bias.grad) < GRAD_TOL


# ------------------------------------------------------------------------------- the model
def make_pair(K, C, dev, deg_by="target", seed=0):
    import bigcn_b200
    torch.manual_seed(seed)
    ref = bigcn_oracle.BiGCN(K, 64, 64, num_classes=C, deg_by=deg_by)
    with torch.no_grad():
        for p in ref.parameters():
            if p.dim() == 1:
                p.uniform_(-0.1, 0.1)
    m = bigcn_b200.BiGCN(K, 64, 64, dev, num_classes=C, deg_by=deg_by).to(dev)
    m.load_state_dict(ref.state_dict())
    return ref, m


def masks_for(m, b, K):
    n = b.x.shape[0]
    seed = m.TDrumorGCN.last_seed
    ktd = torch.from_numpy(gcn_oracle.dropout_keep_mask(seed, 0, np.arange(n), 64 + K, 0.5))
    kbu = torch.from_numpy(gcn_oracle.dropout_keep_mask(seed, 1, np.arange(n), 64 + K, 0.5))
    return ktd, kbu


def compare_grads(m, ref, tol=GRAD_TOL):
    worst = 0.0
    for (name, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, name
        e = rel_err(p.grad, q.grad if q.grad is not None else torch.zeros_like(q))
        assert e < tol, f"{name}: grad rel err {e:.3e}"
        worst = max(worst, e)
    return worst


def test_golden_small_batch(dev):
    import bigcn_b200
    gold = json.load(open(os.path.join(GOLD, "bigcn_small.json")))
    b = Batch(x=torch.from_numpy(unhex(gold["x"]).reshape(gold["N"], gold["K"])),
              edge_index=torch.tensor(gold["edge_index"]).reshape(2, -1),
              BU_edge_index=torch.tensor(gold["BU_edge_index"]).reshape(2, -1),
              batch=torch.tensor(gold["batch"]), rootindex=torch.tensor(gold["rootindex"]),
              y=torch.tensor(gold["y"]))
    m = bigcn_b200.BiGCN(gold["K"], 64, 64, dev, num_classes=gold["C"]).to(dev).eval()
    m.load_state_dict({k: torch.from_numpy(unhex(v["data"]).reshape(v["shape"])) for k, v in gold["state"].items()})
    got = m(b.to(dev))
    m.check_inputs()
    want = torch.from_numpy(unhex(gold["logp"]).reshape(-1, gold["C"]))
    assert rel_err(got, want) < LOGP_TOL
    check_graph_prep(Batch(x=b.x.cpu(), edge_index=b.edge_index.cpu(), BU_edge_index=b.BU_edge_index.cpu(),
                           batch=b.batch.cpu(), rootindex=b.rootindex.cpu()), dev)


def test_reference_model_code_fixture(dev):
    """tests/golden/ref_wiring.npz: outputs of the reference's own BiGCN_Twitter.py:19-131 (run over dense stand-ins of
    the two absent wheels, tests/golden/make_ref_wiring_golden.py) -- eval-mode log-probs, nll loss and all ten
    gradients -- against the CUDA path directly, no oracle in between."""
    import bigcn_b200
    ref = np.load(os.path.join(GOLD, "ref_wiring.npz"))
    gold = json.load(open(os.path.join(GOLD, "bigcn_small.json")))
    b = Batch(x=torch.from_numpy(unhex(gold["x"]).reshape(gold["N"], gold["K"])),
              edge_index=torch.tensor(gold["edge_index"]).reshape(2, -1),
              BU_edge_index=torch.tensor(gold["BU_edge_index"]).reshape(2, -1),
              batch=torch.tensor(gold["batch"]), rootindex=torch.tensor(gold["rootindex"]),
              y=torch.from_numpy(ref["twitter/y"])).to(dev)
    for mode in ("fp32", "auto"):
        m = bigcn_b200.BiGCN(gold["K"], 64, 64, dev, num_classes=4, gemm_mode=mode).to(dev).eval()
        m.load_state_dict({k[14:]: torch.from_numpy(ref[k]) for k in ref.files if k.startswith("twitter/state/")})
        got = m(b)
        loss = torch.nn.functional.nll_loss(got, b.y)
        loss.backward()
        m.check_inputs()
        assert rel_err(got, torch.from_numpy(ref["twitter/eval/logp"])) < LOGP_TOL, mode
        assert abs(float(loss.detach()) - float(ref["twitter/eval/loss"].item())) < 1e-5
        for k, p in m.named_parameters():
            e = rel_err(p.grad, torch.from_numpy(ref[f"twitter/eval/grad/{k}"]))
            assert e < GRAD_TOL, f"{mode} {k}: {e:.3e}"


@pytest.mark.parametrize("deg_by", ["target", "source"])
def test_eval_logits_and_grads_edge_cases(dev, deg_by):
    K, cases = edge_cases()
    for i, (name, b) in enumerate(cases.items()):
        ref, m = make_pair(K, 4 if i % 2 == 0 else 2, dev, deg_by, seed=i)
        ref.eval(); m.eval()
        want = ref(b)
        got = m(clone_batch(b, dev))
        m.check_inputs()
        assert got.shape == want.shape
        e = rel_err(got, want)
        assert e < LOGP_TOL, f"{name}: {e:.3e}"
        assert torch.equal(got.argmax(1).cpu(), want.argmax(1))
        g = torch.randn_like(want)
        want.backward(g)
        got.backward(g.to(dev))
        compare_grads(m, ref)


def test_train_mode_with_injected_philox_mask(dev):
    """Train-mode parity: the kernel's Philox mask (spec in oracle/gcn_oracle.py) is injected
    into the oracle, then log-probs and all ten gradients must agree."""
    K, cases = edge_cases()
    for i, name in enumerate(("twitter_train_dropedge", "pheme_singletons", "star_and_chain", "general_graph")):
        b = cases[name]
        ref, m = make_pair(K, 4, dev, seed=10 + i)
        ref.train(); m.train()
        got = m(clone_batch(b, dev))
        ktd, kbu = masks_for(m, b, K)
        want = ref(b, keep_td=ktd, keep_bu=kbu)
        e = rel_err(got, want)
        assert e < LOGP_TOL, f"{name}: {e:.3e}"
        y = b.y
        torch.nn.functional.nll_loss(want, y).backward()
        torch.nn.functional.nll_loss(got, y.to(dev)).backward()
        compare_grads(m, ref)


def test_twitter_shaped_bow_batch_k5000(dev):
    """BASELINE configs[0] shape at reduced tree count: K=5000 BoW, C=4, DropEdge 0.2/0.2, train mode."""
    b = make_batch("twitter15", 12, seed=0, train=True)
    ref, m = make_pair(5000, 4, dev, seed=3)
    ref.train(); m.train()
    got = m(clone_batch(b, dev))
    ktd, kbu = masks_for(m, b, 5000)
    want = ref(b, keep_td=ktd, keep_bu=kbu)
    assert rel_err(got, want) < LOGP_TOL
    assert torch.equal(got.argmax(1).cpu(), want.argmax(1))
    torch.nn.functional.nll_loss(want, b.y).backward()
    torch.nn.functional.nll_loss(got, b.y.to(dev)).backward()
    compare_grads(m, ref)
    # second call draws a different mask; eval is deterministic and mask-free
    got2 = m(clone_batch(b, dev))
    assert not torch.equal(got2, got)
    m.eval(); ref.eval()
    assert rel_err(m(clone_batch(b, dev)), ref(b)) < LOGP_TOL


def test_weibo_and_pheme_shapes(dev):
    # Weibo: C=2, no DropEdge, reference batch 16, heavy-tailed sizes (capped for oracle time)
    sizes = np.array([10, 2500, 37, 800, 12, 90, 300, 55, 1200, 20, 64, 33, 410, 77, 150, 18])
    b = make_batch("weibo", 16, seed=5, train=True, in_feats=200, sizes=sizes)
    assert torch.equal(b.BU_edge_index, b.edge_index.flip(0))
    ref, m = make_pair(200, 2, dev, seed=4)
    ref.eval(); m.eval()
    assert_logp_parity(m(clone_batch(b, dev)), ref, b, what="weibo")
    # PHEME: K=768 dense signed features, B=24, 21 % single-node trees
    b = make_batch("pheme", 24, seed=6, train=True)
    ref, m = make_pair(768, 4, dev, seed=5)
    ref.train(); m.train()
    got = m(clone_batch(b, dev))
    ktd, kbu = masks_for(m, b, 768)
    want = ref(b, keep_td=ktd, keep_bu=kbu)
    assert rel_err(got, want) < LOGP_TOL
    torch.nn.functional.nll_loss(want, b.y).backward()
    torch.nn.functional.nll_loss(got, b.y.to(dev)).backward()
    compare_grads(m, ref)


def test_weibo_full_size_tree_and_powerlaw_batch(dev):
    """BASELINE configs[2] / [4] at their maximum tree sizes (feature width reduced so the CPU
    oracle finishes in seconds): a 59,318-node Weibo tree inside a reference-sized batch of 16,
    and a power-law batch with a 10,000-node tree.  Hub rows go through the split path, the
    readout spreads the big tree over >100 CTAs; log-probs still within 1e-5 of the fp64 oracle,
    gradients within 1e-4, and two runs are bit-identical."""
    for shape, sizes, C in (("weibo", [59318, 10, 816, 37, 2500, 12, 90, 300, 55, 1200, 20, 64, 33, 410, 77, 150], 2),
                            ("powerlaw", [10000, 2, 3, 2, 7, 120, 2, 2, 45, 4, 2, 900, 2, 6, 2, 3], 4)):
        b = make_batch(shape, len(sizes), seed=21, train=(shape == "powerlaw"), in_feats=96, num_classes=C,
                       sizes=np.array(sizes))
        ref, m = make_pair(96, C, dev, seed=9)
        ref.eval(); m.eval()
        got = m(clone_batch(b, dev))
        m.check_inputs()
        assert_logp_parity(got, ref, b, what=shape)
        again = m(clone_batch(b, dev))
        assert torch.equal(again, got)
        # gradients against the fp64 oracle: the fp32 oracle's own sequential sums over a
        # 59k-node tree are off by ~1e-3
        ref64 = copy.deepcopy(ref).double()
        want = ref64(b)
        g = torch.randn_like(want)
        want.backward(g)
        got.backward(g.float().to(dev))
        compare_grads(m, ref64)


def test_single_direction_modules(dev):
    import bigcn_b200
    K, cases = edge_cases()
    b = cases["twitter_train_dropedge"]
    for cls_ref, cls in ((bigcn_oracle.TDrumorGCN, bigcn_b200.TDrumorGCN), (bigcn_oracle.BUrumorGCN, bigcn_b200.BUrumorGCN)):
        torch.manual_seed(7)
        ref = cls_ref(K, 64, 64).eval()
        m = cls(K, 64, 64, dev).to(dev).eval()
        m.load_state_dict(ref.state_dict())
        want = ref(b)
        got = m(clone_batch(b, dev))
        assert got.shape == (want.shape[0], 128)
        assert rel_err(got, want) < LOGP_TOL
        g = torch.randn_like(want)
        want.backward(g); got.backward(g.to(dev))
        compare_grads(m, ref)


def test_dp_reduce_adam_two_ranks_emulated_on_one_device(dev):
    """bigcn_dp_reduce_adam (reduce-scatter in rank order + Adam on the owned shard + parameter
    all-gather through peer pointers) with both ranks' buffers on this device: after each rank
    ran its slice, both parameter copies equal torch.optim.Adam on the summed gradient."""
    import ctypes as C
    from bigcn_b200 import _lib as L
    lib = L.lib()
    torch.manual_seed(11)
    n, world = 10007, 3                      # n not a multiple of 4: the last slice has a scalar tail
    p0 = torch.randn(n)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, weight_decay=1e-4)
    params = [p0.clone().to(dev) for _ in range(world)]
    ms = [torch.zeros(n, device=dev) for _ in range(world)]
    vs = [torch.zeros(n, device=dev) for _ in range(world)]
    steps = [torch.zeros(4, dtype=torch.int64, device=dev) for _ in range(world)]
    seg_end = torch.tensor([n], dtype=torch.int64, device=dev)
    seg_lr = torch.tensor([1e-3], dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for it in range(3):
        grads = [torch.randn(n) for _ in range(world)]
        ref.grad = sum(g.double() for g in grads).float()
        opt.step()
        gd = [g.to(dev) for g in grads]
        gp = (C.c_void_p * world)(*[g.data_ptr() for g in gd])
        pp = (C.c_void_p * world)(*[p.data_ptr() for p in params])
        for r in range(world):
            L.check(lib.bigcn_dp_reduce_adam(gp, pp, world, r, ms[r].data_ptr(), vs[r].data_ptr(), n,
                                             seg_end.data_ptr(), seg_lr.data_ptr(), 1, 0.9, 0.999, 1e-8, 1e-4, 1.0,
                                             steps[r].data_ptr(), None, None, st))
        for r in range(1, world):
            assert torch.equal(params[r], params[0])           # every copy got the owner's value
        assert rel_err(params[0], ref.data) < 2e-6
    lo, hi = C.c_int64(), C.c_int64()
    cover = []
    for r in range(world):
        lib.bigcn_dp_slice(n, world, r, C.byref(lo), C.byref(hi))
        cover.append((lo.value, hi.value))
    assert cover[0][0] == 0 and cover[-1][1] == n and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))


def test_determinism_run_to_run(dev):
    b = make_batch("twitter15", 10, seed=8, train=True, in_feats=300)
    _, m = make_pair(300, 4, dev, seed=6)
    m.train()
    outs, grads = [], []
    for _ in range(2):
        m.TDrumorGCN._calls = 0
        m.zero_grad(set_to_none=True)
        o = m(clone_batch(b, dev))
        torch.nn.functional.nll_loss(o, b.y.to(dev)).backward()
        outs.append(o.detach().clone())
        grads.append([p.grad.clone() for p in m.parameters()])
    assert torch.equal(outs[0], outs[1])
    for a, c in zip(*grads):
        assert torch.equal(a, c)


def test_fused_trainer_matches_reference_adam(dev):
    """Three steps of FusedTrainer (flat buffer, fused Adam, BU convs at lr/5) against
    torch.optim.Adam on the oracle, eval-mode forward so no mask is involved."""
    import bigcn_b200
    K = 48
    ref, m = make_pair(K, 4, dev, seed=9)
    ref.eval(); m.eval()
    opt = bigcn_oracle.make_optimizer(ref, lr=5e-4, weight_decay=1e-4)
    tr = bigcn_b200.FusedTrainer(m, lr=5e-4, weight_decay=1e-4)
    for step in range(3):
        b = make_batch("twitter15", 6, seed=20 + step, train=True, in_feats=K)
        loss_ref = torch.nn.functional.nll_loss(ref(b), b.y)
        opt.zero_grad(); loss_ref.backward(); opt.step()
        loss = tr.step(clone_batch(b, dev))
        tr.check_inputs()
        assert abs(float(loss.item()) - float(loss_ref)) < 1e-5 * max(1.0, abs(float(loss_ref)))
    for (name, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        assert rel_err(p, q) < 2e-5, name
    assert int(tr.step_count[0].item()) == 3


def test_dense_root_features_training(dev):
    """PHEME-shaped batches (K = 768 dense features, 21 % single-node trees) in training mode: the tiled masked products
    for the root half of conv2.lin and its weight gradient (csrc/rootdense.cu, opts.dense_roots -- what gemm_mode 'auto'
    selects for dense features) against (a) the oracle with the kernel's Philox mask injected, (b) the list walk
    (dense_roots off): the forward is the same sum in the same order, bit for bit; of the gradients only conv2's weight
    may differ, by summation order."""
    import bigcn_b200
    K = 768
    for seed, (shape, nt, kk) in enumerate((("pheme", 24, 768), ("pheme", 90, 768), ("twitter15", 3, 70))):   # K = 70: K % 4 != 0
        b = make_batch(shape, nt, seed=70 + seed, train=True, in_feats=kk)
        if kk != 768:
            b.x = torch.tanh(torch.randn(b.x.shape[0], kk))           # dense, about half positive
        ref, m = make_pair(kk, 4, dev, seed=20 + seed)
        ref.train(); m.train()
        bd = clone_batch(b, dev)
        for mod in (m.TDrumorGCN, m.BUrumorGCN):
            mod.gemm_mode, mod.dense_roots = "fp32", True               # exact products: isolates the root part
        got = m(bd)
        ktd, kbu = masks_for(m, b, kk)
        want = ref(b, keep_td=ktd, keep_bu=kbu)
        assert rel_err(got, want) < LOGP_TOL, (shape, nt, rel_err(got, want))
        torch.nn.functional.nll_loss(want, b.y).backward()
        torch.nn.functional.nll_loss(got, bd.y).backward()
        compare_grads(m, ref)
        m.check_inputs()
        # the list walk with the same seed: identical forward, identical gradients except (by rounding) conv2.lin.weight
        m2 = copy.deepcopy(m)
        m2.zero_grad()
        for mod in (m2.TDrumorGCN, m2.BUrumorGCN):
            mod.dense_roots = False
            mod.seed, mod._calls = m.TDrumorGCN.last_seed, 0
        got2 = m2(bd)
        assert m2.TDrumorGCN.last_seed == m.TDrumorGCN.last_seed
        assert rel_err(got2, got) < 2e-6          # small batches split the columns over CTAs: partial sums, other rounding
        torch.nn.functional.nll_loss(got2, bd.y).backward()
        # ... and with ONE column split the tiled product is the list walk's sum, term by term: bit-identical forward,
        # bit-identical gradients except (summation order) conv2.lin.weight
        lib = bigcn_b200.lib()
        lib.bigcn_debug_set.argtypes = [C.c_int, C.c_int]
        lib.bigcn_debug_set.restype = None
        lib.bigcn_debug_set(15, 1)
        try:
            m3 = copy.deepcopy(m)
            m3.zero_grad()
            for mod in (m3.TDrumorGCN, m3.BUrumorGCN):
                mod.seed, mod._calls = m.TDrumorGCN.last_seed, 0
            got3 = m3(bd)
            torch.nn.functional.nll_loss(got3, bd.y).backward()
        finally:
            lib.bigcn_debug_set(15, 0)
        assert torch.equal(got3, got2)
        for (name, p), (_, q) in zip(m3.named_parameters(), m2.named_parameters()):
            if name.endswith("conv2.lin.weight"):
                assert rel_err(p.grad, q.grad) < 1e-5, name
            else:
                assert torch.equal(p.grad, q.grad), name
    # gemm_mode 'auto' on dense features turns it on by itself; bag-of-words features keep the list walk
    m = bigcn_b200.BiGCN(768, 64, 64, dev).to(dev).train()
    assert m.TDrumorGCN.resolved_dense_roots(clone_batch(make_batch("pheme", 24, seed=1, train=True), dev).x) is True
    m = bigcn_b200.BiGCN(5000, 64, 64, dev).to(dev).train()
    assert m.TDrumorGCN.resolved_dense_roots(clone_batch(make_batch("twitter15", 3, seed=1, train=True), dev).x) is False
