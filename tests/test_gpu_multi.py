"""Multi-GPU data parallelism; needs >= 2 GPUs on the box, skipped otherwise
(the CPU-side logic is covered by tests/test_dist_cpu.py on gloo).  Two ranks train the same model
on different tree shards with (a) the fused peer-memory optimiser step (bigcn_dp_reduce_adam between
symmetric-memory barriers) and (b) NCCL all-reduce + Adam: parameters must be bit-identical across
ranks in both, and agree between the two to fp32 summation order."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    import torch.distributed as dist
    import bigcn_b200
    from bigcn_b200.data import Batch, make_batch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    b = make_batch("twitter15", 6, seed=100 + rank, train=True, in_feats=600)
    bd = Batch(**{k: getattr(b, k).to(dev) for k in Batch._tensor_keys})
    res = {}
    # symm: in-kernel barriers (default) and torch's symmetric-memory barriers around the kernel; nccl: all-reduce + Adam
    # symm: gradient slices PUSHED to their owners + in-kernel barriers (default); symm_pull: peer loads + in-kernel
    # barriers; symm_ext: peer loads between torch's symmetric-memory barriers; nccl: all-reduce + Adam on every rank
    for name, comm, fused, push in (("symm", "symm", True, True), ("symm_pull", "symm", True, False),
                                    ("symm_ext", "symm", False, False), ("nccl", "nccl", True, False)):
        torch.manual_seed(0)
        m = bigcn_b200.BiGCN(600, 64, 64, dev, gemm_mode="sparse", validate="off").to(dev).train()
        tr = bigcn_b200.FusedTrainer(m, process_group=dist.group.WORLD, world_size=world, comm=comm, fused_sync=fused,
                                     dp_push=push, graphs=False)
        assert tr.comm == comm
        for i in range(3):
            tr.step(bd, b_global=6 * world, seed=7 + i)
        tr.check_inputs()
        torch.cuda.synchronize()
        flat = tr.flat.detach().clone()
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        res[name] = (flat.cpu(), all(torch.equal(g, gathered[0]) for g in gathered))
    # CUDA-graph replay of the whole DP step (symm, in-kernel barriers) with the next batch prepared a step ahead:
    # same batches, same order as an enqueued run -> bit-identical parameters, on every rank
    b2 = make_batch("twitter15", 5, seed=200 + rank, train=True, in_feats=600)
    bd2 = Batch(**{k: getattr(b2, k).to(dev) for k in Batch._tensor_keys})
    pair = [bd, bd2]
    flats = {}
    for graphs in (False, True):
        torch.manual_seed(0)
        m = bigcn_b200.BiGCN(600, 64, 64, dev, gemm_mode="sparse", validate="off").to(dev).train()
        tr = bigcn_b200.FusedTrainer(m, process_group=dist.group.WORLD, world_size=world, comm="symm", graphs=graphs)
        for i in range(10):
            tr.step(pair[i % 2], b_global=11 * world // 2 + 1, next_data=pair[(i + 1) % 2])
        tr.check_inputs()
        torch.cuda.synchronize()
        flats[graphs] = tr.flat.detach().clone()
        if graphs:
            res["graph_replays"] = tr.graph_replays
    gathered = [torch.empty_like(flats[True]) for _ in range(world)]
    dist.all_gather(gathered, flats[True])
    res["graph_same_ranks"] = all(torch.equal(g, gathered[0]) for g in gathered)
    res["graph_equals_eager"] = bool(torch.equal(flats[True], flats[False]))
    if rank == 0:
        d = (res["symm"][0].double() - res["nccl"][0].double()).abs().max().item()
        torch.save({"same_symm": res["symm"][1], "same_nccl": res["nccl"][1], "same_symm_ext": res["symm_ext"][1],
                    "fused_equals_ext": bool(torch.equal(res["symm"][0], res["symm_ext"][0])) and
                                        bool(torch.equal(res["symm_pull"][0], res["symm_ext"][0])) and res["symm_pull"][1], "diff": d,
                    "scale": res["nccl"][0].abs().max().item(), "graph_replays": res["graph_replays"],
                    "graph_same_ranks": res["graph_same_ranks"], "graph_equals_eager": res["graph_equals_eager"]}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_symm_and_nccl_agree(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["same_symm"] and r["same_nccl"] and r["same_symm_ext"], r
    assert r["fused_equals_ext"], r
    assert r["diff"] <= 2e-6 * r["scale"], r
    assert r["graph_replays"] >= 4 and r["graph_same_ranks"] and r["graph_equals_eager"], r
