"""CPU tests of the host side of the sparse input path: bigcn_host_dense_to_csr (threaded
one-pass compaction of a dense host matrix, csrc/host_compact.cpp) against torch's CSR."""
import numpy as np
import pytest
import torch

import bigcn_b200
from bigcn_b200 import ops
from bigcn_b200.data import make_batch


def check(x, n_threads):
    sx = ops.host_dense_to_csr(x, n_threads=n_threads, cap=max(x.numel(), 1))
    want = x.to_sparse_csr()
    assert sx.shape == tuple(x.shape)
    assert torch.equal(sx.ptr.long(), want.crow_indices())
    assert torch.equal(sx.col.long(), want.col_indices())          # ascending columns per row
    assert torch.equal(sx.val, want.values())
    assert torch.equal(sx.to_dense(), x)


@pytest.mark.parametrize("n_threads", [1, 3, 0])
def test_host_dense_to_csr_matches_torch(n_threads):
    b = make_batch("twitter15", 6, seed=0, train=False)            # K = 5000 BoW
    check(b.x, n_threads)
    rng = np.random.default_rng(0)
    for n, k in ((1, 1), (7, 33), (40, 31), (5, 64), (300, 771)):  # tails shorter than a 32-float line
        x = torch.from_numpy(np.where(rng.random((n, k)) < 0.1, rng.normal(0, 1, (n, k)), 0).astype(np.float32))
        x[0] = 0
        x[-1, -1] = -2.5
        check(x, n_threads)
    check(torch.zeros(0, 16), n_threads)
    check(torch.zeros(9, 16), n_threads)


def test_host_dense_to_csr_capacity_and_negative_zero():
    x = torch.ones(4, 100)
    with pytest.raises(bigcn_b200.BigcnError):
        ops.host_dense_to_csr(x)                                   # 400 non-zeros > 4 * 48
    sx = ops.host_dense_to_csr(x, cap=400)
    assert int(sx.ptr[-1]) == 400
    z = torch.zeros(2, 40)
    z[0, 3] = -0.0                                                 # -0.0 == 0: not an entry
    assert int(ops.host_dense_to_csr(z).ptr[-1]) == 0


def test_sparse_x_roundtrip_from_torch_sparse():
    b = make_batch("twitter16", 2, seed=1, train=False, in_feats=300)
    sx = ops.SparseX.from_torch_csr(b.x.to_sparse_csr())
    assert sx.ptr.dtype == torch.int32 and sx.col.dtype == torch.int32
    assert torch.equal(sx.to_dense(), b.x)
    sx2 = ops.SparseX.from_torch_csr(b.x.to_sparse())
    assert torch.equal(sx2.to_dense(), b.x)
