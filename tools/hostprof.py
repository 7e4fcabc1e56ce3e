import sys, time, cProfile, pstats, io
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bigcn_b200
from bigcn_b200.data import Batch, make_batch_shard
dev = torch.device("cuda", 0)
b = make_batch_shard("twitter16", 128, 1000)[0]
bd = Batch(**{k: getattr(b, k).to(dev) for k in Batch._tensor_keys})
bd.x = bigcn_b200.host_dense_to_csr(b.x).to(dev)
torch.manual_seed(0)
m = bigcn_b200.BiGCN(5000, 64, 64, dev, gemm_mode="sparse", validate="off").to(dev).train()
for graphs in (False, "fresh"):
    tr = bigcn_b200.FusedTrainer(m, graphs=bool(graphs))
    def mk():
        return Batch(x=bd.x, edge_index=bd.edge_index, BU_edge_index=bd.BU_edge_index, batch=bd.batch, rootindex=bd.rootindex, y=bd.y)
    for i in range(5):
        tr.step(mk() if graphs else bd)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(200):
        tr.step(mk() if graphs else bd)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"graphs={graphs}: host enqueue {1e3*(t1-t0)/200:.4f} ms/step, total {1e3*(t2-t0)/200:.4f} ms/step", flush=True)
pr = cProfile.Profile()
pr.enable()
for i in range(200):
    tr.step(mk())
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25)
print(s.getvalue()[:5000])
