#!/usr/bin/env python
"""A few eager PHEME-shaped training steps (B from argv, default 4096) for ncu: dense features, gemm_mode auto."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bigcn_b200  # noqa: E402
from bigcn_b200.data import Batch, make_batch_shard  # noqa: E402

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = make_batch_shard("pheme", B, 2000)[0]
bd = Batch(**{k: getattr(b, k).to(dev) for k in Batch._tensor_keys})
torch.manual_seed(0)
m = bigcn_b200.BiGCN(768, 64, 64, dev, validate="off").to(dev).train()
tr = bigcn_b200.FusedTrainer(m, graphs=False)
for _ in range(4):
    tr.step(bd)
tr.check_inputs()
torch.cuda.synchronize()
print("ok", bd.x.shape)
