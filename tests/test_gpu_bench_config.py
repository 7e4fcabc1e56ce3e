"""Parity AT the bench configuration (BASELINE.json configs[1]): 128 Twitter16-shaped trees, K = 5000,
gemm_mode='sparse', one FusedTrainer.step in train mode -- the exact call bench.py times -- against the CPU
oracle handed the kernel's own Philox mask (reference: BiGCN_Twitter.py:183-189 forward, nll_loss, backward).
Bars: loss and log-probs 1e-5, all ten gradients 1e-4.  Plus: a CUDA-graph replay of the step is
bit-identical to the enqueued step, and draws a fresh dropout mask on every replay."""
import numpy as np
import pytest
import torch

from oracle import bigcn_oracle, gcn_oracle
from bigcn_b200.data import Batch, make_batch, make_batch_shard

pytestmark = pytest.mark.gpu


def rel_err(got, want):
    want = want.detach().cpu().double()
    got = got.detach().cpu().double()
    return float((got - want).abs().max()) / max(float(want.abs().max()), 1e-30)


def to_dev(b, dev):
    return Batch(**{k: getattr(b, k).to(dev) for k in Batch._tensor_keys})


def test_fused_trainer_step_at_bench_configuration(dev):
    import bigcn_b200
    from bigcn_b200.trainer import _ORDER
    K, C, trees = 5000, 4, 128
    b = make_batch_shard("twitter16", trees, 1000)[0]          # bench.py's first batch (seed 1000), DropEdge 0.2/0.2
    n = int(b.x.shape[0])
    assert n > 20000 and int(b.rootindex.numel()) == trees
    torch.manual_seed(0)
    ref = bigcn_oracle.BiGCN(K, 64, 64, num_classes=C).train()
    with torch.no_grad():
        for p in ref.parameters():
            if p.dim() == 1:
                p.uniform_(-0.1, 0.1)
    m = bigcn_b200.BiGCN(K, 64, 64, dev, num_classes=C, gemm_mode="sparse", validate="off").to(dev).train()
    m.load_state_dict(ref.state_dict())
    tr = bigcn_b200.FusedTrainer(m, lr=5e-4, weight_decay=1e-4, graphs=False)
    seed = 12345
    loss = tr.step(to_dev(b, dev), seed=seed)
    tr.check_inputs()
    got_logp = tr.last_logp.cpu()
    got_grads = {name: tr.gviews[name].detach().cpu().clone() for name in _ORDER}
    ktd = torch.from_numpy(gcn_oracle.dropout_keep_mask(seed, 0, np.arange(n), 64 + K, 0.5))
    kbu = torch.from_numpy(gcn_oracle.dropout_keep_mask(seed, 1, np.arange(n), 64 + K, 0.5))
    want = ref(b, keep_td=ktd, keep_bu=kbu)
    want_loss = torch.nn.functional.nll_loss(want, b.y)
    want_loss.backward()
    assert rel_err(got_logp, want) < 1e-5
    assert abs(float(loss.item()) - float(want_loss)) <= 1e-5 * max(1.0, abs(float(want_loss)))
    for name, q in ref.named_parameters():
        e = rel_err(got_grads[name], q.grad)
        assert e < 1e-4, f"{name}: {e:.3e}"


def _trainer(dev, graphs, K=5000):
    import bigcn_b200
    torch.manual_seed(3)
    m = bigcn_b200.BiGCN(K, 64, 64, dev, num_classes=4, gemm_mode="sparse", validate="off").to(dev).train()
    return m, bigcn_b200.FusedTrainer(m, lr=5e-4, weight_decay=1e-4, graphs=graphs)


def test_graph_replay_is_bit_identical_to_enqueued_steps(dev):
    """Same batches, same order: the trainer that replays CUDA graphs ends with exactly the parameters,
    moments and losses of the trainer that enqueues every launch (deterministic kernels, no atomics)."""
    batches = [to_dev(make_batch("twitter16", 24, seed=40 + i, train=True), dev) for i in range(2)]
    out = {}
    for graphs in (False, True):
        m, tr = _trainer(dev, graphs)
        losses = []
        for i in range(8):
            losses.append(tr.step(batches[i % 2]).clone())
        tr.check_inputs()
        out[graphs] = (tr.flat.clone(), tr.exp_avg.clone(), tr.exp_avg_sq.clone(), torch.cat(losses), tr)
    tr = out[True][4]
    assert tr.graph_captures == 2 and tr.graph_replays == 6          # first sighting enqueued, second captured
    assert int(tr.step_count[0].item()) == 8 and int(tr.step_count[2].item()) == 8
    for a, c in zip(out[False][:4], out[True][:4]):
        assert torch.equal(a, c)
    # fresh masks per replay: the same batch twice in a row must not reproduce the same loss
    l = out[True][3]
    assert float(l[2]) != float(l[4]) and float(l[4]) != float(l[6])


def test_graph_replay_inference_matches_eager(dev):
    import bigcn_b200
    torch.manual_seed(5)
    b = to_dev(make_batch("pheme", 24, seed=7, train=False), dev)
    m = bigcn_b200.BiGCN(768, 64, 64, dev, gemm_mode="fp32", graphs=True).to(dev).eval()
    e = bigcn_b200.BiGCN(768, 64, 64, dev, gemm_mode="fp32").to(dev).eval()
    e.load_state_dict(m.state_dict())
    with torch.no_grad():
        want = e(b).clone()
        outs = [m(b).clone() for _ in range(4)]       # enqueue, capture + replay, replay, replay
    m.check_inputs()
    for o in outs:
        assert torch.equal(o, want)
    # parameters changed in place (a training step between validation passes): the replay must see them
    with torch.no_grad():
        for p, q in zip(m.parameters(), e.parameters()):
            p.mul_(1.01)
            q.mul_(1.01)
        assert torch.equal(m(b), e(b))


def test_prefetched_batches_match_unprefetched_steps(dev):
    """step(data, next_data=...) prepares the next batch a step ahead (bigcn_batch_prepare: graph prep, root
    columns, x -> CSR -> CSC); the result must be the step's own -- bit-identical for bag-of-words rows (the
    product over the prepared CSR runs in the order of the fused scan), with or without CUDA graphs."""
    batches = [to_dev(make_batch("twitter16", 20, seed=60 + i, train=True), dev) for i in range(3)]
    out = {}
    for name, graphs, prefetch in (("plain", False, False), ("prefetch", False, True), ("prefetch+graphs", True, True)):
        m, tr = _trainer(dev, graphs)
        losses = []
        for i in range(14):
            nxt = batches[(i + 1) % 3] if prefetch else None
            losses.append(tr.step(batches[i % 3], next_data=nxt).clone())
        tr.check_inputs()
        out[name] = (tr.flat.clone(), tr.exp_avg.clone(), torch.cat(losses), tr)
    assert out["prefetch+graphs"][3].graph_replays >= 2
    for name in ("prefetch", "prefetch+graphs"):
        for a, c in zip(out["plain"][:3], out[name][:3]):
            assert torch.equal(a, c), name
    # a batch that was NOT announced is prepared inline and still right; so is a dense-mode model
    m, tr = _trainer(dev, False)
    l0 = tr.step(batches[0], next_data=batches[2]).clone()
    l1 = tr.step(batches[1]).clone()            # buffer holds batch 2: batch 1 goes the self-contained way
    l2 = tr.step(batches[2]).clone()            # prepared two steps ago
    m2, tr2 = _trainer(dev, False)
    for l, b in zip((l0, l1, l2), batches):
        assert torch.equal(l, tr2.step(b))
    assert torch.equal(tr.flat, tr2.flat)


def test_prefetch_pheme_dense_features(dev):
    import bigcn_b200
    batches = [to_dev(make_batch("pheme", 400, seed=80 + i, train=True), dev) for i in range(3)]     # ~3.8 k rows each
    small = to_dev(make_batch("pheme", 24, seed=90, train=True), dev)
    res = []
    for prefetch in (False, True):
        torch.manual_seed(4)
        m = bigcn_b200.BiGCN(768, 64, 64, dev, validate="off").to(dev).train()      # auto -> tf32x3 for dense features
        tr = bigcn_b200.FusedTrainer(m, graphs=prefetch)
        ls = [tr.step(batches[i % 3], next_data=batches[(i + 1) % 3] if prefetch else None).clone() for i in range(9)]
        tr.check_inputs()
        assert m.TDrumorGCN.resolved_gemm_mode(batches[0].x) == "tf32x3"
        assert m.TDrumorGCN.resolved_gemm_mode(small.x) == "tf32x3"     # small dense batches too: the FFMA scan walks dense
        #                                                                   rows one element at a time (0.24 vs 0.085 ms at B = 24)
        res.append((tr.flat.clone(), torch.cat(ls)))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
