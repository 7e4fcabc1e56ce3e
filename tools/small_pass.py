#!/usr/bin/env python
"""A small pass through every kernel of the training step (run on the GPU box; written for
`compute-sanitizer --tool memcheck python tools/small_pass.py`, which this pool refuses -- the bounds are held by the
flag words (check_inputs) and the parity tests instead).  Dense BoW batches (the capture pass, ELL product, CSC sort, sweeps), the pipelined step with next_data, a CSR-input
batch from a DeviceForest, inference, and the dense tensor-core mode on a PHEME-shaped batch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bigcn_b200  # noqa: E402
from bigcn_b200.data import Batch, make_batch, make_tree  # noqa: E402

dev = torch.device("cuda", 0)


def to_dev(b):
    return Batch(**{k: getattr(b, k).to(dev) for k in Batch._tensor_keys})


torch.manual_seed(0)
batches = [to_dev(make_batch("twitter16", 6, seed=40 + i, train=True)) for i in range(3)]
m = bigcn_b200.BiGCN(5000, 64, 64, dev, gemm_mode="sparse", validate="off").to(dev).train()
tr = bigcn_b200.FusedTrainer(m, graphs=False)
for i in range(4):
    tr.step(batches[i % 3], next_data=batches[(i + 1) % 3])
tr.step(batches[0])
tr.check_inputs()
rng = np.random.default_rng(0)
forest = bigcn_b200.DeviceForest.from_data_list([make_tree("twitter15", int(n), rng, in_feats=5000) for n in (30, 1, 200, 17, 90, 5)], dev)
for b, nxt in forest.batches([[0, 1, 2], [3, 4, 5], [5, 0, 2]], 0.2, 0.2):
    tr.step(b, next_data=nxt)
tr.check_inputs()
m.eval()
with torch.no_grad():
    out = m(batches[1])
m.check_inputs()
ph = to_dev(make_batch("pheme", 24, seed=7, train=True))
mp = bigcn_b200.BiGCN(768, 64, 64, dev, validate="off").to(dev).train()
trp = bigcn_b200.FusedTrainer(mp, graphs=False)
for _ in range(2):
    trp.step(ph)
trp.check_inputs()
torch.cuda.synchronize()
print("small_pass ok", float(out.sum()))
