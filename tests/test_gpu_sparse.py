"""GPU tests of gemm_mode='sparse' (SURVEY.md 8f, N1): the forward X*W scan captures the
non-zeros of data.x, a stable radix sort builds the column-sorted copy on the side stream and
the conv1 weight gradient is a CSR sweep over the columns -- no second pass over X.  Same bars
as the dense fp32 path: log-probs 1e-5, all ten gradients 1e-4, bit-reproducible."""
import numpy as np
import pytest
import torch

from oracle import bigcn_oracle, gcn_oracle
from bigcn_b200.data import Batch, make_batch

pytestmark = pytest.mark.gpu


def rel_err(got, want):
    want = want.detach().cpu().double()
    got = got.detach().cpu().double()
    return float((got - want).abs().max()) / max(float(want.abs().max()), 1e-30)


def to_dev(b, dev):
    return Batch(**{k: getattr(b, k).to(dev) for k in Batch._tensor_keys})


def pair(K, C, dev, seed):
    import bigcn_b200
    torch.manual_seed(seed)
    ref = bigcn_oracle.BiGCN(K, 64, 64, num_classes=C)
    with torch.no_grad():
        for p in ref.parameters():
            if p.dim() == 1:
                p.uniform_(-0.1, 0.1)
    m = bigcn_b200.BiGCN(K, 64, 64, dev, num_classes=C, gemm_mode="sparse").to(dev)
    m.load_state_dict(ref.state_dict())
    return ref, m


def check_train_step(b, K, C, dev, seed, what):
    ref, m = pair(K, C, dev, seed)
    ref.train(); m.train()
    n = b.x.shape[0]
    got = m(to_dev(b, dev))
    m.check_inputs()
    s = m.TDrumorGCN.last_seed
    ktd = torch.from_numpy(gcn_oracle.dropout_keep_mask(s, 0, np.arange(n), 64 + K, 0.5))
    kbu = torch.from_numpy(gcn_oracle.dropout_keep_mask(s, 1, np.arange(n), 64 + K, 0.5))
    want = ref(b, keep_td=ktd, keep_bu=kbu)
    assert rel_err(got, want) < 1e-5, what
    torch.nn.functional.nll_loss(want, b.y).backward()
    torch.nn.functional.nll_loss(got, b.y.to(dev)).backward()
    for (name, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        e = rel_err(p.grad, q.grad)
        assert e < 1e-4, f"{what} {name}: {e:.3e}"
    return m


def test_sparse_mode_twitter_bow(dev):
    b = make_batch("twitter15", 12, seed=0, train=True)
    check_train_step(b, 5000, 4, dev, 3, "twitter15 K=5000")


def test_sparse_mode_long_rows_hot_columns_negatives(dev):
    """Rows with more non-zeros than the 32 ELL slots (re-read from the dense row), a column set
    in EVERY row (hub column: split path of the sweep), negative and fractional values, an
    all-zero row, K not a multiple of 4 (scalar load path of the scan)."""
    rng = np.random.default_rng(1)
    for K in (5000, 333):
        b = make_batch("twitter16", 6, seed=2, train=True, in_feats=K, num_classes=2)
        x = b.x.numpy().copy()
        n = x.shape[0]
        x[:, 7] = rng.integers(1, 4, n)                       # hot column
        x[:, K - 1] = np.where(rng.random(n) < 0.5, -1.5, 0)  # last column, negatives
        heavy = rng.choice(n, 9, replace=False)
        for r in heavy:                                       # 40..90 non-zeros in a row
            cols = rng.choice(K, int(rng.integers(40, 91)), replace=False)
            x[r, cols] = rng.normal(0, 1, cols.size).astype(np.float32)
        x[int(heavy[0]) ^ 1] = 0.0                            # an empty row
        b.x = torch.from_numpy(x)
        check_train_step(b, K, 2, dev, 4, f"K={K}")


def test_sparse_mode_eval_and_edge_shapes(dev):
    # single-node trees only (E = 0), PHEME-shaped small trees with K <= 48 (capacity = N*K)
    for shape, nt, K in (("pheme", 24, 40), ("twitter15", 1, 64)):
        b = make_batch(shape, nt, seed=5, train=False, in_feats=K)
        if shape == "pheme":
            b.x = torch.where(torch.rand_like(b.x) < 0.3, b.x, torch.zeros_like(b.x))
        ref, m = pair(K, 4, dev, 6)
        ref.eval(); m.eval()
        want = ref(b)
        got = m(to_dev(b, dev))
        m.check_inputs()
        assert rel_err(got, want) < 1e-5
        g = torch.randn_like(want)
        want.backward(g); got.backward(g.to(dev))
        for (name, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
            assert rel_err(p.grad, q.grad) < 1e-4, name
    # inference under no_grad skips the column sort altogether
    with torch.no_grad():
        assert rel_err(m(to_dev(b, dev)), want) < 1e-5


def test_sparse_mode_is_deterministic_and_matches_dense_modes(dev):
    import bigcn_b200
    b = make_batch("twitter15", 10, seed=8, train=True, in_feats=1200)
    ref, m = pair(1200, 4, dev, 7)
    m.train()
    runs = []
    for mode in ("sparse", "sparse", "fp32"):
        m.gemm_mode = mode
        m.TDrumorGCN._calls = 0
        m.zero_grad(set_to_none=True)
        o = m(to_dev(b, dev))
        torch.nn.functional.nll_loss(o, b.y.to(dev)).backward()
        runs.append((o.detach().clone(), [p.grad.clone() for p in m.parameters()]))
    assert torch.equal(runs[0][0], runs[1][0])
    for a, c in zip(runs[0][1], runs[1][1]):
        assert torch.equal(a, c)
    assert torch.equal(runs[0][0], runs[2][0])                 # same forward kernel
    for a, c in zip(runs[0][1], runs[2][1]):
        assert rel_err(a, c) < 2e-6                            # dW1: another summation order only


def test_sparse_mode_rejects_dense_features(dev):
    """Dense 768-d features hold N*768 non-zeros > N*48: the violation flag is raised and the
    conv1 weight gradient is NaN rather than silently wrong."""
    b = make_batch("pheme", 24, seed=9, train=True)
    _, m = pair(768, 4, dev, 8)
    m.train()
    m.validate = "off"
    o = m(to_dev(b, dev))
    torch.nn.functional.nll_loss(o, b.y.to(dev)).backward()
    assert torch.isnan(m.TDrumorGCN.conv1.lin.weight.grad).all()
    with pytest.raises(IndexError):
        m.check_inputs()


def test_sparse_input_x_never_dense_on_device(dev):
    """data.x handed over as CSR (ops.SparseX from the host compaction, or a torch sparse
    tensor): forward through the CSR product, root rows and weight gradient from the same CSR.
    Same bars against the oracle run on the dense matrix."""
    from bigcn_b200 import ops
    b = make_batch("twitter15", 9, seed=12, train=True, in_feats=2000)
    x = b.x.numpy().copy()
    x[:, 11] = 2.0                                   # hub column
    x[5, :100] = np.linspace(-1, 1, 100)             # a 100-entry row
    b.x = torch.from_numpy(x)
    K, n = 2000, b.x.shape[0]
    ref, m = pair(K, 4, dev, 21)
    ref.train(); m.train()
    sx = ops.host_dense_to_csr(b.x, n_threads=2, cap=n * 64)
    bd = to_dev(b, dev)
    bd.x = sx.to(dev)
    got = m(bd)
    m.check_inputs()
    s = m.TDrumorGCN.last_seed
    ktd = torch.from_numpy(gcn_oracle.dropout_keep_mask(s, 0, np.arange(n), 64 + K, 0.5))
    kbu = torch.from_numpy(gcn_oracle.dropout_keep_mask(s, 1, np.arange(n), 64 + K, 0.5))
    want = ref(b, keep_td=ktd, keep_bu=kbu)
    assert rel_err(got, want) < 1e-5
    torch.nn.functional.nll_loss(want, b.y).backward()
    torch.nn.functional.nll_loss(got, b.y.to(dev)).backward()
    for (name, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        assert rel_err(p.grad, q.grad) < 1e-4, name
    # torch sparse tensor as data.x, eval mode
    m.eval(); ref.eval()
    bd.x = b.x.to_sparse_csr().to(dev)
    assert rel_err(m(bd), ref(b)) < 1e-5
    # a column index outside [0, K) is reported, not read out of bounds
    bad = ops.SparseX(sx.ptr.clone(), sx.col.clone(), sx.val.clone(), sx.shape)
    bad.col[3] = K + 5
    bd.x = bad.to(dev)
    m.train()
    m.validate = "off"
    out = m(bd)
    out.sum().backward()
    with pytest.raises(IndexError):
        m.check_inputs()


def test_fused_trainer_sparse_input_matches_dense_input(dev):
    """FusedTrainer steps fed with SparseX batches track the same steps fed with dense x."""
    import bigcn_b200
    from bigcn_b200 import ops
    b = make_batch("twitter16", 8, seed=13, train=True, in_feats=1500)
    losses = []
    for sparse_in in (False, True):
        torch.manual_seed(5)
        m = bigcn_b200.BiGCN(1500, 64, 64, dev, gemm_mode="sparse", validate="off").to(dev).train()
        tr = bigcn_b200.FusedTrainer(m, lr=5e-4, weight_decay=1e-4)
        bd = to_dev(b, dev)
        if sparse_in:
            bd.x = ops.host_dense_to_csr(b.x).to(dev)
        ls = [float(tr.step(bd, seed=100 + i).item()) for i in range(4)]
        tr.check_inputs()
        losses.append(ls)
    for a, c in zip(*losses):
        assert abs(a - c) < 2e-6 * max(1.0, abs(a))
    assert losses[0][-1] < losses[0][0]
