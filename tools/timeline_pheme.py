#!/usr/bin/env python
"""Kernel timeline of one replayed PHEME-shaped training step (B = 24, K = 768 dense, gemm_mode auto -> tf32x3)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bigcn_b200
from bigcn_b200.data import Batch, make_batch_shard
from torch.profiler import ProfilerActivity, profile
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
bs = []
for i in range(3):
    b = make_batch_shard("pheme", B, 2000 + i)[0]
    bs.append(Batch(**{k: getattr(b, k).to(dev) for k in Batch._tensor_keys}))
torch.manual_seed(0)
m = bigcn_b200.BiGCN(768, 64, 64, dev, validate="off").to(dev).train()
tr = bigcn_b200.FusedTrainer(m)
for i in range(24):
    tr.step(bs[i % 3], next_data=bs[(i + 1) % 3])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(24, 28):
        tr.step(bs[i % 3], next_data=bs[(i + 1) % 3])
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
starts = [i for i, e in enumerate(ev) if "k_transpose_jobs" in e.name]
ev = ev[starts[-2]:starts[-1]]
t0 = ev[0].time_range.start
for e in ev:
    print(f"{e.time_range.start - t0:8.1f} {e.time_range.end - e.time_range.start:7.1f}  {e.name[:90]}")
