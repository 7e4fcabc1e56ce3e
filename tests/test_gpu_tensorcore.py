"""GPU tests of the tcgen05/TMA GEMM modes of X * W^T (north_star: TF32 mode within 1e-2
relative with identical argmax; the split mode restores fp32-class accuracy on BoW counts)."""
import numpy as np
import pytest
import torch

from oracle import bigcn_oracle
from bigcn_b200.data import make_batch

pytestmark = pytest.mark.gpu


def rel_err(got, want):
    want = want.detach().cpu().double()
    got = got.detach().cpu().double()
    return float((got - want).abs().max()) / max(float(want.abs().max()), 1e-30)


@pytest.mark.parametrize("n_trees,k", [(3, 5000), (9, 5000), (2, 64), (5, 768)])
def test_xw_tensor_core_matches_fp64(dev, n_trees, k):
    from bigcn_b200 import ops
    torch.manual_seed(0)
    shape = "pheme" if k == 768 else "twitter15"
    b = make_batch(shape, n_trees, seed=11, train=False, in_feats=k)
    if shape == "pheme":                       # dense signed features: make them TF32-exact for the split mode
        b.x = (b.x * 64).round() / 64
    w_td, w_bu = torch.randn(64, k) * 0.05, torch.randn(64, k) * 0.05
    want = b.x.double() @ torch.cat([w_td, w_bu]).double().t()
    x = b.x.to(dev)
    got = ops.xw(x, [w_td.to(dev), w_bu.to(dev)], "tf32")
    assert got.shape == want.shape
    assert rel_err(got, want) < 2e-3           # one TF32 pass: 10-bit mantissa on W
    got3 = ops.xw(x, [w_td.to(dev), w_bu.to(dev)], "tf32x3")
    assert rel_err(got3, want) < 2e-6          # hi/lo split: fp32-class
    one = ops.xw(x, [w_bu.to(dev)], "tf32x3")  # single direction, N = 64 MMA
    assert rel_err(one, want[:, 64:]) < 2e-6
    # fp32 scan and tensor-core split agree
    ref = ops.xw(x, [w_td.to(dev), w_bu.to(dev)], "fp32")
    assert rel_err(got3, ref) < 2e-6


def test_model_in_tf32_mode(dev):
    import bigcn_b200
    torch.manual_seed(1)
    b = make_batch("twitter15", 16, seed=2, train=False)
    ref = bigcn_oracle.BiGCN(5000, 64, 64).eval()
    want = ref(b)
    for mode, tol in (("tf32", 1e-2), ("tf32x3", 1e-5)):
        m = bigcn_b200.BiGCN(5000, 64, 64, dev, gemm_mode=mode).to(dev).eval()
        m.load_state_dict(ref.state_dict())
        got = m(b.to(dev) if not b.x.is_cuda else b)
        m.check_inputs()
        e = rel_err(got, want)
        assert e < tol + (1.2e-5 if mode == "tf32x3" else 0), f"{mode}: {e:.3e}"
        assert torch.equal(got.argmax(1).cpu(), want.argmax(1)), mode
