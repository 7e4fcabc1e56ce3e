"""Golden vectors from the REFERENCE'S OWN DATASET CODE (input contract of the hot path):

    python tests/golden/make_ref_dataset_golden.py       # needs /root/reference; writes ref_dataset/*.npz + ref_dataset.npz

Seven small trees are written in the reference's on-disk format (one ``<id>.npz`` with x / edgeindex / rootindex / y /
root / cls / tweetids, as Process/getTwittergraph.py:67-72 and getPHEMEgraph.py save them) and read back by the
reference's ``BiGraphDataset`` (Process/dataset.py:45-99, imported unmodified; ``torch_geometric.data.Data`` -- absent
from this image -- is stood in for by a plain attribute bag, the class holds no logic the dataset uses).  Stored:
what ``__getitem__`` returns for every kept id without DropEdge, and with DropEdge 0.2 / 0.3 under ``random.seed(7)``
(so that the count rule ``int(e * (1 - rate))`` and the order-preserving subset are the reference's).  The id filter
(``lower=2``: the single-node tree is dropped, so is an id missing from treeDic) is exercised too.
tests/test_gpu_loader.py holds ``DeviceForest.from_npz_dir(...).batch(...)`` to these files.
"""
import os
import random
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
K = 12


def main():
    from bigcn_b200.data import make_tree
    tg, tg_data = types.ModuleType("torch_geometric"), types.ModuleType("torch_geometric.data")
    tg_data.Data = type("Data", (), {"__init__": lambda self, **kw: self.__dict__.update(kw)})
    tg.data = tg_data
    sys.modules.update({"torch_geometric": tg, "torch_geometric.data": tg_data})
    sys.path.insert(0, REF)
    from Process.dataset import BiGraphDataset          # the reference's class

    d = os.path.join(HERE, "ref_dataset")
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(5)
    ids = ["1001", "1002", "1003", "1004", "1005", "1006", "1007"]
    sizes = [9, 1, 14, 5, 2, 23, 7]                      # 1002 is a single post: filtered out by lower=2
    treeDic = {}
    for i, n in zip(ids, sizes):
        t = make_tree("twitter15", n, rng, in_feats=K)
        ei = t.edge_index.numpy()
        np.savez(os.path.join(d, i + ".npz"), x=t.x.numpy(), root=t.x.numpy()[int(t.rootindex)][None, :], edgeindex=ei,
                 rootindex=int(t.rootindex), y=int(t.y), cls=np.zeros((1, 4), np.float32),
                 tweetids=np.arange(n) + 10 * int(i))
        treeDic[i] = {j: {} for j in range(n)}
    fold_x = ids + ["9999"]                              # not in treeDic: filtered out
    out = {"K": np.array(K), "fold_x": np.array(fold_x)}
    for tag, td, bu in (("nodrop", 0, 0), ("drop", 0.2, 0.3)):
        ds = BiGraphDataset(fold_x, treeDic, lower=2, upper=100000, tddroprate=td, budroprate=bu, data_path=d)
        out[f"{tag}/kept_ids"] = np.array(ds.fold_x)
        random.seed(7)
        for j in range(len(ds)):
            data, _ = ds[j]
            for key in ("x", "edge_index", "BU_edge_index", "rootindex", "y"):
                out[f"{tag}/{j}/{key}"] = getattr(data, key).numpy()
    np.savez(os.path.join(HERE, "ref_dataset.npz"), **out)
    print("written:", len(out), "arrays;", "kept", list(out["nodrop/kept_ids"]))


if __name__ == "__main__":
    main()
