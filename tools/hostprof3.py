#!/usr/bin/env python
"""Where the HOST time of the device-dataset loop goes (run on the GPU box): DeviceForest.batch + FusedTrainer.step on
freshly assembled batches (no graph replay), wall clock per call and a cProfile of 300 iterations by own time."""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bigcn_b200  # noqa: E402
from bigcn_b200.data import synth_forest_device  # noqa: E402

dev = torch.device("cuda", 0)
f = synth_forest_device("twitter16", 818, dev, seed=16)
forest = bigcn_b200.DeviceForest.from_device_arrays(f)
torch.manual_seed(0)
m = bigcn_b200.BiGCN(5000, 64, 64, dev, gemm_mode="sparse", validate="off").to(dev).train()
tr = bigcn_b200.FusedTrainer(m)
rng = np.random.default_rng(0)
ids = [rng.choice(818, 128, replace=False) for _ in range(64)]
for i in range(10):
    tr.step(forest.batch(ids[i % 64], 0.2, 0.2, seed=i))
torch.cuda.synchronize()
tb = ts = 0.0
n = 300
t00 = time.perf_counter()
for i in range(n):
    t0 = time.perf_counter()
    b = forest.batch(ids[i % 64], 0.2, 0.2, seed=i)
    t1 = time.perf_counter()
    tr.step(b)
    t2 = time.perf_counter()
    tb += t1 - t0
    ts += t2 - t1
t3 = time.perf_counter()
torch.cuda.synchronize()
t4 = time.perf_counter()
print(f"host: batch {1e3 * tb / n:.4f} ms, step {1e3 * ts / n:.4f} ms, loop {1e3 * (t3 - t00) / n:.4f} ms; incl. drain {1e3 * (t4 - t00) / n:.4f} ms", flush=True)
# the same loop with the next batch handed to the step (its weight-independent half runs underneath)
lists = [ids[i % 64] for i in range(n)]
t0 = time.perf_counter()
for b, nxt in forest.batches(lists, 0.2, 0.2):
    tr.step(b, next_data=nxt)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"with next_data: host loop {1e3 * (t1 - t0) / n:.4f} ms, incl. drain {1e3 * (t2 - t0) / n:.4f} ms", flush=True)
for name, look in (("plain loop", False), ("with next_data", True)):
    pr = cProfile.Profile()
    pr.enable()
    if look:
        for b, nxt in forest.batches(lists, 0.2, 0.2):
            tr.step(b, next_data=nxt)
    else:
        for i in range(n):
            tr.step(forest.batch(ids[i % 64], 0.2, 0.2, seed=i))
    pr.disable()
    torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(16)
    print(name)
    print(s.getvalue()[:3600])
