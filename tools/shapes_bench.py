#!/usr/bin/env python
"""Trees/s of the hot path on every dataset shape BASELINE.json names (run on the GPU box; under
torchrun the global batch is sharded over the ranks like bench.py does):

  c1/c2 twitter15 / twitter16   B = 128, K = 5000, C = 4      train + inference
  c3    weibo                   B = 16 (reference) and 128    train, exact fp32 and tensor-core GEMM modes
  c4    pheme                   B = 24 (reference) and 4096   inference, K = 768 dense features
  c5    powerlaw                B = 128 and 2048              train (Pareto(2) sizes up to 10k nodes)

Batches are device-resident (3 in rotation; training steps prepare the next batch beside the current one and replay
CUDA graphs, inference replays CUDA graphs), timed with CUDA events after warm-up, max over ranks.
Prints one JSON line per (shape, batch, mode, phase); `python tools/shapes_bench.py > profiles/...`."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bigcn_b200  # noqa: E402
from bigcn_b200.data import SHAPES, Batch, make_batch_shard  # noqa: E402

CASES = [
    ("twitter15", 128, "sparse", ("train", "infer")),
    ("twitter16", 128, "sparse", ("train", "infer")),
    ("weibo", 16, "sparse", ("train",)),
    ("weibo", 128, "sparse", ("train", "infer")),
    ("weibo", 128, "tf32x3", ("train",)),
    ("weibo", 128, "tf32", ("train",)),
    ("pheme", 24, "auto", ("infer", "train")),          # auto -> tf32x3 (fp32-class, X split in shared memory)
    ("pheme", 4096, "auto", ("infer", "train")),
    ("pheme", 4096, "tf32", ("infer",)),
    ("pheme", 4096, "fp32", ("infer",)),                # the exact FFMA scan on dense features, for comparison
    ("powerlaw", 128, "sparse", ("train",)),
    ("powerlaw", 2048, "sparse", ("train", "infer")),
]


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    only = sys.argv[1:] or None
    cache = {}
    for shape, bsz, mode, phases in CASES:
        if only and shape not in only:
            continue
        cfg = SHAPES[shape]
        if (shape, bsz) not in cache:
            cache.clear()
            torch.cuda.empty_cache()
            shards = [make_batch_shard(shape, bsz * world, seed=2000 + i, rank=rank, world=world, train=True)
                      for i in range(3)]
            cache[(shape, bsz)] = ([Batch(**{k: getattr(s[0], k).to(dev) for k in Batch._tensor_keys}) for s in shards],
                                   [s[1] for s in shards])
        res, base = cache[(shape, bsz)]
        nodes = [int(b.x.shape[0]) for b in res]
        torch.manual_seed(0)
        model = bigcn_b200.BiGCN(cfg["in_feats"], 64, 64, dev, num_classes=cfg["num_classes"], gemm_mode=mode,
                                 validate="off", graphs=True).to(dev)
        tr = bigcn_b200.FusedTrainer(model, process_group=pg, world_size=world)
        for phase in phases:
            model.train(phase == "train")

            def one(i):
                if phase == "train":
                    return tr.step(res[i % 3], b_global=bsz * world, node_id_base=base[i % 3], next_data=res[(i + 1) % 3])
                with torch.no_grad():
                    return model(res[i % 3])
            steps = 30
            for i in range(24):          # every (batch, next batch, buffer) combination: enqueued once, then captured
                one(i)
            torch.cuda.synchronize()
            if world > 1:
                torch.distributed.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(24, 24 + steps):
                one(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
                ms = float(t.item())
            tr.check_inputs()
            model.check_inputs()
            if rank == 0:
                print(json.dumps({"shape": shape, "trees_per_gpu": bsz, "n_gpus": world, "gemm_mode": mode, "phase": phase,
                                  "ms_per_step": round(ms, 4), "trees_per_s": round(bsz * world / (ms * 1e-3), 1),
                                  "nodes_per_step": nodes, "nodes_per_s": round(sum(nodes) / 3 * world / (ms * 1e-3), 1),
                                  "x_bytes_per_step": int(sum(nodes) / 3 * cfg["in_feats"] * 4)}), flush=True)
        del tr, model
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
