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
    for comm in ("symm", "nccl"):
        torch.manual_seed(0)
        m = bigcn_b200.BiGCN(600, 64, 64, dev, gemm_mode="sparse", validate="off").to(dev).train()
        tr = bigcn_b200.FusedTrainer(m, process_group=dist.group.WORLD, world_size=world, comm=comm)
        assert tr.comm == comm
        for i in range(3):
            tr.step(bd, b_global=6 * world, seed=7 + i)
        tr.check_inputs()
        torch.cuda.synchronize()
        flat = tr.flat.detach().clone()
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        res[comm] = (flat.cpu(), all(torch.equal(g, gathered[0]) for g in gathered))
    if rank == 0:
        d = (res["symm"][0].double() - res["nccl"][0].double()).abs().max().item()
        torch.save({"same_symm": res["symm"][1], "same_nccl": res["nccl"][1], "diff": d,
                    "scale": res["nccl"][0].abs().max().item()}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_symm_and_nccl_agree(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["same_symm"] and r["same_nccl"]
    assert r["diff"] <= 2e-6 * r["scale"], r
