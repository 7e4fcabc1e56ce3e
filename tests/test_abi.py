"""CPU tests: the C-ABI library loads and exports every symbol include/bigcn_b200.h declares
(no compute call is made without a GPU), and the product refuses to run on the CPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "bigcn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bigcn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built_lib):
    lib = ctypes.CDLL(built_lib)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/bigcn_b200.h but not exported"


def test_python_binding_covers_header(built_lib):
    from bigcn_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()
    assert _lib.lib().bigcn_version() >= 100


def test_no_cpu_fallback(built_lib):
    import bigcn_b200
    from bigcn_b200.data import make_batch
    if torch.cuda.is_available():
        pytest.skip("this check is for GPU-less hosts")
    m = bigcn_b200.BiGCN(16, 64, 64)
    b = make_batch("twitter15", 1, seed=0, train=False, in_feats=16)
    with pytest.raises(bigcn_b200.BigcnError):
        m(b)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "bigcn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_state_dict_keys_match_reference_layout(built_lib):
    import bigcn_b200
    m = bigcn_b200.BiGCN(32, 64, 64, "cpu")
    keys = sorted(m.state_dict())
    want = sorted([f"{d}.{c}.{p}" for d in ("TDrumorGCN", "BUrumorGCN") for c in ("conv1", "conv2")
                   for p in ("lin.weight", "bias")] + ["fc.weight", "fc.bias"])
    assert keys == want
    assert m.TDrumorGCN.conv2.lin.weight.shape == (64, 64 + 32) and m.fc.weight.shape == (4, 256)
    assert sum(p.numel() for p in bigcn_b200.BiGCN(5000, 64, 64).parameters()) == 1289476
    assert bigcn_b200.Net(5000, 64, 64).fc.weight.shape == (2, 256)
    # the optimizer construction of BiGCN_Twitter.py:146-153 works unchanged
    bu = list(map(id, m.BUrumorGCN.conv1.parameters())) + list(map(id, m.BUrumorGCN.conv2.parameters()))
    assert len(bu) == 4
    # PyG-1.3.2 checkpoints (weight [in,out]) load
    sd = m.state_dict()
    old = {}
    for k, v in sd.items():
        old[k.replace("lin.weight", "weight")] = v.t().contiguous() if k.endswith("lin.weight") else v
    m2 = bigcn_b200.BiGCN(32, 64, 64, "cpu")
    m2.load_state_dict(old)
    assert torch.equal(m2.TDrumorGCN.conv1.lin.weight, m.TDrumorGCN.conv1.lin.weight)
