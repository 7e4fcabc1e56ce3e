"""The other BASELINE.json configurations, measured by bench.py into the `configs` object of its JSON line
(the headline `value` stays configs[1]'s training step):

  c2_epoch        Twitter16-shaped FULL TRAINING EPOCH on 1 B200: train_GCN's loop (BiGCN_Twitter.py:160-246) over a
                  654 / 164 fold of 818 trees -- 6 train batches (the last one partial), 2 validation batches,
                  DropEdge + collate on the device per batch, metrics on the device, one host read per phase
  c3_weibo        Weibo-shaped training (C = 2, 4664 trees averaging ~800 nodes, DropEdge 0/0) at the reference's
                  batch 16 and at 128, in the fp32-class mode ('sparse': exact fp32) and the tensor-core modes
  c4_pheme_infer  PHEME-shaped 9-fold inference (K = 768 dense, 6425 trees in 9 event-sized folds), batches of 24
                  (the reference's) and 4096, trees sharded over the ranks, no communication
  c5_powerlaw     1 M power-law reply trees streamed once through DeviceForest.batch + FusedTrainer.step,
                  data parallel over the ranks (strong scaling: the million trees are split across the GPUs)

Every number is device-timed with CUDA events between host synchronisations unless it says wall clock."""
from __future__ import annotations

import time

import numpy as np

PHEME_EVENT_SIZES = (2079, 1221, 1143, 890, 469, 238, 233, 138, 14)     # the 9 PHEME events, 6425 threads


def _timed(torch, fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3, out


def c2_epoch(torch, bigcn_b200, dev, epochs=4, warm=2):
    from bigcn_b200.data import synth_forest_device
    f = synth_forest_device("twitter16", 818, dev, seed=16)
    forest = bigcn_b200.DeviceForest.from_device_arrays(f)
    perm = np.random.default_rng(16).permutation(818)
    train_ids, test_ids = perm[:654], perm[654:]
    torch.manual_seed(0)
    model = bigcn_b200.BiGCN(5000, 64, 64, dev, num_classes=4, gemm_mode="sparse", validate="off").to(dev)
    tr = bigcn_b200.FusedTrainer(model, lr=5e-4, weight_decay=1e-4)
    tev, vev = bigcn_b200.EvalCounts(4, dev), bigcn_b200.EvalCounts(4, dev)
    rng = np.random.default_rng(0)
    state = {"epoch": 0, "val_acc": 0.0, "train_loss": 0.0}

    def epoch():
        ep = state["epoch"]
        model.train()
        tev.reset()
        order = rng.permutation(train_ids)
        lists = [order[lo:lo + 128] for lo in range(0, len(order), 128)]
        for data, nxt in forest.batches(lists, 0.2, 0.2, seeds=[ep * 4096 + bi for bi in range(len(lists))]):
            tr.step(data, next_data=nxt)
            tev.update(tr.last_logp, data.y)
        state["train_loss"], _, _ = tev.epoch_means()            # host read (BiGCN_Twitter.py:199-200)
        model.eval()
        vev.reset()
        order = rng.permutation(test_ids)
        with torch.no_grad():
            for lo in range(0, len(order), 128):
                data = forest.batch(order[lo:lo + 128], 0.0, 0.0)
                vev.update(model(data), data.y)
        _, state["val_acc"], _ = vev.epoch_means()               # host read (:226-246)
        state["epoch"] = ep + 1

    for _ in range(warm):
        epoch()
    dev_ms, wall_ms, _ = _timed(torch, lambda: [epoch() for _ in range(epochs)])
    tr.check_inputs()
    nodes = int(f["node_ptr"][-1])
    return {"workload": "Twitter16-shaped epoch: 654 train trees (5 x 128 + 14, DropEdge 0.2/0.2, train mode, Adam) + 164 "
                        "validation trees (128 + 36, eval mode), batches assembled on the device, metrics on the device",
            "epoch_ms": wall_ms / epochs, "epoch_ms_device": dev_ms / epochs, "epochs_timed": epochs,
            "train_trees_per_s": 654 * epochs / (wall_ms * 1e-3), "trees_per_s_incl_validation": 818 * epochs / (wall_ms * 1e-3),
            "dataset_nodes": nodes, "timing": "wall clock incl. the two host reads per epoch", "gemm_mode": "sparse",
            "final_train_loss": state["train_loss"], "val_acc": state["val_acc"]}


def _train_steps(torch, bigcn_b200, dev, batches, k, c, mode, steps):
    torch.manual_seed(0)
    model = bigcn_b200.BiGCN(k, 64, 64, dev, num_classes=c, gemm_mode=mode, validate="off").to(dev).train()
    tr = bigcn_b200.FusedTrainer(model, lr=5e-4, weight_decay=1e-4)
    nb = len(batches)
    # the loader loop of bench.py: step i is handed batch i+1 (its weight-independent half runs beside step i); every
    # (batch, next batch, buffer) combination is enqueued once and captured once before the timed steps replay
    for i in range(4 * nb + 2):
        tr.step(batches[i % nb], next_data=batches[(i + 1) % nb])
    tr.check_inputs()
    i0 = 4 * nb + 2
    ms, _, _ = _timed(torch, lambda: [tr.step(batches[i % nb], next_data=batches[(i + 1) % nb]) for i in range(i0, i0 + steps)])
    return ms / steps


def c3_weibo(torch, bigcn_b200, dev, steps=24):
    from bigcn_b200.data import Batch, synth_forest_device
    f = synth_forest_device("weibo", 4664, dev, seed=3)
    forest = bigcn_b200.DeviceForest.from_device_arrays(f)
    rng = np.random.default_rng(3)
    out = {"workload": "Weibo-shaped training step (graph prep+fwd+nll+bwd+Adam), C=2, K=5000, DropEdge 0/0, dropout 0.5, "
                       "4664 trees, dense fp32 data.x resident in HBM", "dataset_nodes": int(f["node_ptr"][-1]), "by_batch": {}}
    for bsz, nbatch in ((16, 6), (128, 3)):
        batches = []
        for j in range(nbatch):
            ids = rng.choice(4664, bsz, replace=False)
            b = forest.batch(ids, 0.0, 0.0, seed=j)
            batches.append(Batch(x=b.x.to_dense(), edge_index=b.edge_index, BU_edge_index=b.BU_edge_index, batch=b.batch,
                                 rootindex=b.rootindex, y=b.y))
        nodes = [int(b.x.shape[0]) for b in batches]
        row = {"nodes_per_batch": nodes, "feature_mb_per_batch": round(sum(nodes) / len(nodes) * 5000 * 4 / 1e6)}
        for mode, label in (("sparse", "fp32_class_sparse"), ("tf32x3", "tf32x3_tcgen05"), ("tf32", "tf32_tcgen05")):
            ms = _train_steps(torch, bigcn_b200, dev, batches, 5000, 2, mode, steps)
            row[label] = {"ms_per_step": ms, "trees_per_s": bsz / (ms * 1e-3)}
        out["by_batch"][str(bsz)] = row
        del batches
    return out


def c4_pheme_infer(torch, bigcn_b200, dev, rank, world, max_over_ranks):
    from bigcn_b200.data import forest_slice_batch, synth_forest_device
    n_trees = sum(PHEME_EVENT_SIZES)
    f = synth_forest_device("pheme", n_trees, dev, seed=9)          # every rank replays the same dataset
    fold_lo = np.concatenate([[0], np.cumsum(PHEME_EVENT_SIZES)])
    out = {"workload": "PHEME-shaped 9-fold inference: 6425 trees (mean ~9.6 nodes, 21% single-node) in 9 event-sized test "
                       "folds, K=768 dense features, eval mode, no DropEdge; every fold's trees split contiguously over the "
                       "ranks, no communication", "dataset_nodes": int(f["node_ptr"][-1]), "by_batch": {}}
    for bsz in (24, 4096):
        batches = []
        for k in range(9):
            lo, hi = int(fold_lo[k]), int(fold_lo[k + 1])
            per = -(-(hi - lo) // world)
            a, b = min(hi, lo + rank * per), min(hi, lo + (rank + 1) * per)
            for t0 in range(a, b, bsz):
                batches.append(forest_slice_batch(f, t0, min(b, t0 + bsz)))
        row = {"batches_per_rank": len(batches)}
        for mode, label in (("auto", "fp32_class_tf32x3"), ("tf32", "tf32")):     # auto resolves to tf32x3 on dense features
            torch.manual_seed(0)
            model = bigcn_b200.BiGCN(768, 64, 64, dev, num_classes=4, gemm_mode=mode, validate="off", graphs=True,
                                     max_graphs=len(batches) + 1).to(dev).eval()
            with torch.no_grad():
                for _ in range(2):                     # enqueue, then capture
                    for b in batches:
                        model(b)
                ms, _, _ = _timed(torch, lambda: [model(b) for b in batches])
                ms2, _, _ = _timed(torch, lambda: [model(b) for b in batches])
            ms = max_over_ranks(min(ms, ms2))
            row[label] = {"ms_per_pass": ms, "trees_per_s": n_trees / (ms * 1e-3),
                          "ms_per_batch": ms / max(1, len(batches))}
            del model
        out["by_batch"][str(bsz)] = row
    return out


def c5_powerlaw(torch, bigcn_b200, dev, rank, world, pg, max_over_ranks, total_trees=1_000_000):
    from bigcn_b200.data import synth_forest_device
    per_rank = total_trees // world
    f = synth_forest_device("powerlaw", per_rank, dev, seed=50 + rank)
    forest = bigcn_b200.DeviceForest.from_device_arrays(f)
    out = {"workload": f"{total_trees} power-law reply trees (Pareto(2) sizes, 2..10000 nodes), K=5000 BoW kept as CSR in HBM, "
                       f"one pass: DeviceForest.batches (collate + DropEdge 0.2/0.2 on the device) + FusedTrainer.step(batch, next_data=next batch); the trees "
                       f"are split over the {world} rank(s) (strong scaling), gradients reduced every step",
           "trees_per_rank": per_rank, "nodes_per_rank": int(f["node_ptr"][-1]), "by_batch": {}}
    for bsz in (128, 4096):
        torch.manual_seed(0)
        model = bigcn_b200.BiGCN(5000, 64, 64, dev, num_classes=4, gemm_mode="sparse", validate="off").to(dev).train()
        tr = bigcn_b200.FusedTrainer(model, lr=5e-4, weight_decay=1e-4, process_group=pg, world_size=world, graphs=False)
        nsteps = per_rank // bsz
        if world > 1:      # every rank must run the same number of steps (a collective per step)
            t = torch.tensor([nsteps], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN)
            nsteps = int(t.item())
        ids = np.arange(per_rank, dtype=np.int64)

        def run(n0, n1):
            lists = [ids[s * bsz:(s + 1) * bsz] for s in range(n0, n1)]
            for data, nxt in forest.batches(lists, 0.2, 0.2, seeds=range(n0, n1)):
                tr.step(data, b_global=bsz * world, next_data=nxt)
        run(0, min(20, nsteps))                   # warm-up on the first batches (they are streamed again below)
        tr.check_inputs()
        if world > 1:
            torch.distributed.barrier()
        ms, wall, _ = _timed(torch, lambda: run(0, nsteps))
        ms = max_over_ranks(max(ms, wall))
        out["by_batch"][str(bsz)] = {"steps": nsteps, "ms_total": ms, "ms_per_step": ms / max(1, nsteps),
                                     "trees_per_s": nsteps * bsz * world / (ms * 1e-3),
                                     "timing": "max(device, wall) over the whole pass, max over ranks"}
        tr.check_inputs()
        del tr, model
    return out
