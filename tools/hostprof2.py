import sys, time, cProfile, pstats, io
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bigcn_b200
from bigcn_b200.data import synth_forest_device
dev = torch.device("cuda", 0)
f = synth_forest_device("twitter16", 818, dev, seed=16)
forest = bigcn_b200.DeviceForest.from_device_arrays(f)
ids = np.arange(128)
for i in range(5):
    forest.batch(ids, 0.2, 0.2, seed=i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(300):
    forest.batch(ids, 0.2, 0.2, seed=i)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"forest.batch host {1e3*(t1-t0)/300:.4f} ms, total {1e3*(time.perf_counter()-t0)/300:.4f} ms")
pr = cProfile.Profile(); pr.enable()
for i in range(300):
    forest.batch(ids, 0.2, 0.2, seed=i)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:3500])
