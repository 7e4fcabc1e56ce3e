"""CPU tests of the data-parallel host logic with a world_size-2 gloo group: tree sharding by
node count, loss scaling by 1/B_global and the flat-gradient all-reduce reproduce the
single-process gradient."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

from bigcn_b200.dist import shard_trees, node_id_base, allreduce_flat_


def test_shard_trees_balances_nodes_and_covers_all():
    rng = np.random.default_rng(0)
    sizes = np.clip(np.rint(np.exp(rng.normal(5.9, 1.25, 300))), 10, 59318).astype(int)
    for world in (1, 2, 4, 8):
        parts = shard_trees(sizes, world)
        assert parts[0][0] == 0 and parts[-1][1] == len(sizes)
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        loads = [int(sizes[lo:hi].sum()) for lo, hi in parts]
        assert sum(loads) == int(sizes.sum())
        assert max(loads) <= sizes.sum() / world + sizes.max()
    assert shard_trees([5, 5], 4)[-1][1] == 2          # more ranks than trees: some ranks are empty
    assert node_id_base(sizes, 3) == int(sizes[:3].sum())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import bigcn_oracle
    from bigcn_b200.data import make_tree, collate
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    rng = np.random.default_rng(5)
    sizes = [60, 300, 70, 90, 150, 80]
    trees = [make_tree("twitter15", n, rng, in_feats=48) for n in sizes]
    torch.manual_seed(0)
    model = bigcn_oracle.BiGCN(48, 64, 64).eval()
    lo, hi = shard_trees(sizes, world)[rank]
    b = collate(trees[lo:hi])
    out = model(b)
    loss = torch.nn.functional.nll_loss(out, b.y, reduction="sum") / len(sizes)   # 1 / B_global
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    allreduce_flat_(flat)
    if rank == 0:
        q.put(flat.numpy())
    dist.destroy_process_group()


def test_two_rank_gradient_equals_single_process():
    from oracle import bigcn_oracle
    from bigcn_b200.data import make_tree, collate
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    sizes = [60, 300, 70, 90, 150, 80]
    trees = [make_tree("twitter15", n, rng, in_feats=48) for n in sizes]
    torch.manual_seed(0)
    model = bigcn_oracle.BiGCN(48, 64, 64).eval()
    b = collate(trees)
    torch.nn.functional.nll_loss(model(b), b.y).backward()
    want = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).numpy()
    assert np.abs(got - want).max() <= 1e-5 * max(np.abs(want).max(), 1e-30)
