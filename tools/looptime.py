#!/usr/bin/env python
"""Device-dataset loop, host vs device time per step (run on the GPU box): plain loop and the lookahead loop
(DeviceForest.batches + next_data), each twice; FusedTrainer(graphs=...) from argv."""
import os
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
graphs = {"auto": "auto", "0": False}[sys.argv[1] if len(sys.argv) > 1 else "auto"]
tr = bigcn_b200.FusedTrainer(m, graphs=graphs)
rng = np.random.default_rng(0)
n = 300
lists = [rng.choice(818, 128, replace=False) for _ in range(n)]


def run(look):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    if look:
        for b, nxt in forest.batches(lists, 0.2, 0.2):
            tr.step(b, next_data=nxt)
    else:
        for i in range(n):
            tr.step(forest.batch(lists[i], 0.2, 0.2, seed=i))
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return 1e3 * (t1 - t0) / n, 1e3 * (t2 - t0) / n, e0.elapsed_time(e1) / n


for rep in range(2):
    for look in (False, True):
        h, w, d = run(look)
        print(f"rep {rep} {'lookahead' if look else 'plain    '}: host loop {h:.4f} ms, wall {w:.4f} ms, device {d:.4f} ms per step", flush=True)
tr.check_inputs()
