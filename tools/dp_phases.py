#!/usr/bin/env python
"""Where the data-parallel optimiser kernel spends its time (run under torchrun): k_dp_reduce_adam stamps globaltimer at its
start, after the 'all gradients complete' barrier, when the rank's own slice is done and after the 'every rank done' barrier."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist, bigcn_b200
from bigcn_b200.data import Batch, make_batch_shard
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
bs = []
for i in range(3):
    b = make_batch_shard("twitter16", 128 * world, 1000 + i, rank=rank, world=world)[0]
    bs.append(Batch(**{k: getattr(b, k).to(dev) for k in Batch._tensor_keys}))
torch.manual_seed(0)
m = bigcn_b200.BiGCN(5000, 64, 64, dev, gemm_mode="sparse", validate="off").to(dev).train()
tr = bigcn_b200.FusedTrainer(m, process_group=dist.group.WORLD, world_size=world)
rows = []
for i in range(60):
    tr.step(bs[i % 3], b_global=128 * world, next_data=bs[(i + 1) % 3])
    if i >= 30:
        torch.cuda.synchronize()
        t = tr._sig[32:36].cpu().tolist()
        rows.append([t[0]] + [(t[k] - t[0]) / 1e3 for k in (1, 2, 3)])
import numpy as np
a = np.array(rows)
starts = torch.tensor(a[:, 0], dtype=torch.float64, device=dev)
allst = [torch.empty_like(starts) for _ in range(world)]
dist.all_gather(allst, starts)
skew = (torch.stack(allst) - torch.stack(allst).min(0).values).cpu().numpy() / 1e3      # us after the first rank's start
med = np.median(a[:, 1:], axis=0)
out = torch.tensor(list(med) + [float(np.median(skew[rank]))], dtype=torch.float64, device=dev)
g = [torch.empty_like(out) for _ in range(world)]
dist.all_gather(g, out)
if rank == 0:
    print("rank: us from kernel start to [grads of all ranks complete | own slice done | every rank done]; start skew vs first rank (medians of 30 steps)")
    for r, v in enumerate(g):
        print(r, [round(float(x), 1) for x in v])
dist.barrier()
dist.destroy_process_group()
