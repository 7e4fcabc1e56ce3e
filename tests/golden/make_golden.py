"""Regenerates the golden fixtures from the CPU oracle (run from the repo root):

    python tests/golden/make_golden.py

kat_tree5.json   -- the hand-checkable 5-node known-answer vector of SURVEY.md 8c
                    (gcn_norm weights as exact fp32 hex, both degree conventions).
bigcn_small.json -- a 5-tree batch (K=24, C=4; single-node tree, root not at local
                    index 0, unsorted edges, a self-loop and a duplicate edge), the oracle's
                    parameters, graph_prep arrays and eval-mode log-probs, fp32 hex.
The reference itself cannot be imported here (torch_geometric / torch_scatter absent),
so these pin the restatement, not PyG: parity stays "unpinned" (oracle/__init__.py).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bigcn_oracle, gcn_oracle  # noqa: E402
from bigcn_b200.data import Data, collate  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def hexf(a):
    return [float(v).hex() for v in np.asarray(a, dtype=np.float32).ravel()]


def kat_tree5():
    ei = torch.tensor([[0, 0, 1, 1], [1, 2, 3, 4]])
    out = {}
    for name, e in (("TD", ei), ("BU", ei.flip(0))):
        for deg_by in ("target", "source"):
            e2, w = gcn_oracle.gcn_norm(e, 5, deg_by=deg_by)
            g = gcn_oracle.graph_prep(e.numpy(), 5, np.zeros(5, np.int64), 1, deg_by)
            out[f"{name}_{deg_by}"] = dict(coo=e2.tolist(), w=hexf(w), deg=g["deg"].tolist(),
                                            dis=hexf(g["dis"]), rowsum=hexf(g["rowsum"]),
                                            in_ptr=g["in_ptr"].tolist(), in_idx=g["in_idx"].tolist(),
                                            out_ptr=g["out_ptr"].tolist(), out_idx=g["out_idx"].tolist())
    json.dump(out, open(os.path.join(HERE, "kat_tree5.json"), "w"), indent=1)


def small_batch():
    rng = np.random.default_rng(7)
    K, C = 24, 4
    trees = []
    specs = [  # (n, edges [parent, child], root)
        (1, [], 0),
        (4, [(2, 0), (2, 1), (1, 3)], 2),
        (6, [(5, 0), (0, 3), (5, 1), (3, 2), (0, 4)], 5),            # unsorted by parent
        (3, [(0, 1), (0, 1), (1, 1), (1, 2)], 0),                    # duplicate edge and a self-loop
        (7, [(0, i) for i in range(1, 7)], 0),                       # star
    ]
    for n, edges, root in specs:
        x = np.zeros((n, K), np.float32)
        for i in range(n):
            cols = rng.choice(K, rng.integers(1, 5), replace=False)
            x[i, cols] = rng.integers(1, 4, len(cols))
        e = np.array(edges, np.int64).reshape(-1, 2).T
        trees.append(Data(x=torch.from_numpy(x), edge_index=torch.from_numpy(e.copy()),
                          BU_edge_index=torch.from_numpy(e[::-1].copy()),
                          rootindex=torch.tensor([root]), y=torch.tensor([int(rng.integers(0, C))])))
    b = collate(trees)
    torch.manual_seed(11)
    m = bigcn_oracle.BiGCN(K, 64, 64, num_classes=C).eval()
    with torch.no_grad():
        for p in m.parameters():      # non-zero biases so they are exercised
            if p.dim() == 1:
                p.uniform_(-0.1, 0.1)
        logp = m(b)
    N = b.x.shape[0]
    g_td = gcn_oracle.graph_prep(b.edge_index.numpy(), N, b.batch.numpy(), 5)
    g_bu = gcn_oracle.graph_prep(b.BU_edge_index.numpy(), N, b.batch.numpy(), 5)

    def gdump(g):
        return {k: (hexf(v) if v.dtype == np.float32 else v.tolist()) for k, v in g.items()
                if isinstance(v, np.ndarray)}
    out = dict(K=K, C=C, x=hexf(b.x), N=N, edge_index=b.edge_index.tolist(),
               BU_edge_index=b.BU_edge_index.tolist(), batch=b.batch.tolist(),
               rootindex=b.rootindex.tolist(), y=b.y.tolist(),
               state={k: dict(shape=list(v.shape), data=hexf(v)) for k, v in m.state_dict().items()},
               logp=hexf(logp), graph_td=gdump(g_td), graph_bu=gdump(g_bu))
    json.dump(out, open(os.path.join(HERE, "bigcn_small.json"), "w"))


if __name__ == "__main__":
    kat_tree5()
    small_batch()
    print("golden fixtures written")
