"""A dataset resident on the device and batch assembly there.

Stands in for ``BiGraphDataset`` (/root/reference/Process/dataset.py:45-99: one ``.npz`` per tree
loaded per epoch, DropEdge per direction) plus PyG's ``DataLoader`` collate
(/root/reference/model/Twitter/BiGCN_Twitter.py:162-169) when the whole dataset fits in HBM --
at ~14 non-zeros per node even the 1 M-tree synthetic scale-out config does.  Trees are packed
once (features as CSR, never densified); ``batch(ids)`` builds the batch on the device with one
call into libbigcn_b200.so and hands ``forward(data)`` a sparse ``data.x``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from ._lib import check, lib
from .data import Batch
from .ops import SparseX, _p, _stream


class DeviceForest:
    def __init__(self, node_ptr, edge_ptr, edge_src, edge_dst, x_ptr, x_col, x_val, root_local, y, in_feats, device):
        # host copies of the (small) size arrays: batch offsets are computed without touching the device
        self.h_node_ptr = np.asarray(node_ptr, np.int64)
        self.h_edge_ptr = np.asarray(edge_ptr, np.int64)
        self.h_x_ptr_at_tree = np.asarray(x_ptr, np.int64)[self.h_node_ptr]   # nnz offset of each tree's first node
        self.in_feats = int(in_feats)
        self.device = torch.device(device)
        dev = self.device
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)  # noqa: E731
        self.node_ptr, self.edge_ptr = t(node_ptr, torch.int64), t(edge_ptr, torch.int64)
        self.edge_src, self.edge_dst = t(edge_src, torch.int32), t(edge_dst, torch.int32)
        self.x_ptr, self.x_col, self.x_val = t(x_ptr, torch.int64), t(x_col, torch.int32), t(x_val, torch.float32)
        self.root_local, self.y = t(root_local, torch.int32), t(y, torch.int64)
        self.num_trees = len(self.h_node_ptr) - 1
        self._stage, self._turn = None, 0

    @staticmethod
    def from_device_arrays(f):
        """From ``data.synth_forest_device`` (sparse features): everything already sits on the device; only the
        size arrays live on the host."""
        self = DeviceForest.__new__(DeviceForest)
        dev = f["edge_src"].device
        self.h_node_ptr = np.asarray(f["node_ptr"], np.int64)
        self.h_edge_ptr = np.asarray(f["edge_ptr"], np.int64)
        self.node_ptr = torch.from_numpy(self.h_node_ptr).to(dev)
        self.edge_ptr = torch.from_numpy(self.h_edge_ptr).to(dev)
        self.h_x_ptr_at_tree = f["x_ptr"][self.node_ptr].cpu().numpy()
        self.in_feats, self.device = int(f["in_feats"]), dev
        self.edge_src, self.edge_dst = f["edge_src"], f["edge_dst"]
        self.x_ptr, self.x_col, self.x_val = f["x_ptr"], f["x_col"], f["x_val"]
        self.root_local, self.y = f["root_local"], f["y"].to(torch.int64)
        self.num_trees = len(self.h_node_ptr) - 1
        self._stage, self._turn = None, 0
        return self

    @staticmethod
    def from_data_list(trees, device):
        """Pack ``Data`` objects with the attribute layout of dataset.py:91-98 (x dense or SparseX,
        edge_index [2,e] = [parent; child] WITHOUT DropEdge, rootindex, y)."""
        node_ptr, edge_ptr, x_ptr = [0], [0], [0]
        es, ed, xc, xv, roots, ys = [], [], [], [], [], []
        k = None
        for d in trees:
            x = d.x
            if isinstance(x, SparseX):
                ptr, col, val, n = x.ptr.numpy().astype(np.int64), x.col.numpy(), x.val.numpy(), x.shape[0]
                k = x.shape[1]
            else:
                xn = x.numpy()
                n, k = xn.shape
                r, c = np.nonzero(xn)
                col, val = c.astype(np.int32), xn[r, c]
                ptr = np.concatenate([[0], np.cumsum(np.bincount(r, minlength=n))]).astype(np.int64)
            ei = d.edge_index.numpy()
            es.append(ei[0].astype(np.int32)); ed.append(ei[1].astype(np.int32))
            xc.append(col); xv.append(val)
            x_ptr.extend((x_ptr[-1] + ptr[1:]).tolist())
            node_ptr.append(node_ptr[-1] + n)
            edge_ptr.append(edge_ptr[-1] + ei.shape[1])
            roots.append(int(d.rootindex)); ys.append(int(d.y))
        cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)  # noqa: E731
        return DeviceForest(node_ptr, edge_ptr, cat(es, np.int32), cat(ed, np.int32), x_ptr, cat(xc, np.int32),
                            cat(xv, np.float32), roots, ys, k, device)

    @staticmethod
    def from_npz_dir(fold_x, data_path, device, treeDic=None, lower=2, upper=100000):
        """Pack the reference's preprocessed dataset: one ``<id>.npz`` per tree under ``data_path`` with ``x`` [n, K],
        ``edgeindex`` [2, e] = [parent; child], ``rootindex`` and ``y`` (written by Process/getTwittergraph.py:67-72,
        read per item and per epoch by ``BiGraphDataset.__getitem__``, Process/dataset.py:64-99 -- here once).  The id
        filter is ``BiGraphDataset.__init__``'s (dataset.py:48-59): PHEME paths take ``fold_x`` as is, the others keep
        the ids found in ``treeDic`` whose tree has ``lower .. upper`` nodes.  ``self.ids`` holds the kept ids in
        order: tree t of the forest is ``ids[t]``."""
        import os
        from .data import Data
        if str(data_path).find("PHEME") == -1 and treeDic is not None:
            fold_x = [i for i in fold_x if i in treeDic and lower <= len(treeDic[i]) <= upper]
        trees = []
        for i in fold_x:
            z = np.load(os.path.join(data_path, str(i) + ".npz"), allow_pickle=True)
            ei = np.asarray(z["edgeindex"], np.int64).reshape(2, -1)
            trees.append(Data(x=torch.tensor(np.asarray(z["x"]), dtype=torch.float32), edge_index=torch.from_numpy(ei),
                              rootindex=torch.tensor([int(z["rootindex"])]), y=torch.tensor([int(z["y"])])))
        forest = DeviceForest.from_data_list(trees, device)
        forest.ids = list(fold_x)
        return forest

    def batch(self, tree_ids, td_droprate=0.0, bu_droprate=0.0, seed=0) -> Batch:
        """The collated batch of ``tree_ids`` (host sequence, in batch order) with DropEdge at the
        given rates, on the device; ``data.x`` is a SparseX."""
        L.require_device()
        ids = np.asarray(tree_ids, np.int64)
        b = len(ids)
        n_t = self.h_node_ptr[ids + 1] - self.h_node_ptr[ids]
        e_t = self.h_edge_ptr[ids + 1] - self.h_edge_ptr[ids]
        z_t = self.h_x_ptr_at_tree[ids + 1] - self.h_x_ptr_at_tree[ids]
        # int(length * (1 - droprate)), dataset.py:72,84 (Python float arithmetic = IEEE double)
        keep = lambda rate: e_t if rate <= 0 else np.floor(e_t.astype(np.float64) * (1.0 - rate)).astype(np.int64)  # noqa: E731
        offs = np.zeros((5, b + 1), np.int64)
        offs[0, :b] = ids
        for row, v in zip(range(1, 5), (n_t, keep(td_droprate), keep(bu_droprate), z_t)):
            offs[row, 1:] = np.cumsum(v)
        n, e_td, e_bu, nnz = (int(offs[r, -1]) for r in range(1, 5))
        dev = self.device
        # offsets through a ring of pinned staging buffers; a slot is rewritten only after the H2D copy that last
        # read it has completed (the epoch loop never synchronises the host, so the device may lag many batches)
        need = 5 * (b + 1)
        if self._stage is None or self._stage[0][0].numel() < need:
            if self._stage is not None:
                for _, ev in self._stage:
                    if ev is not None:
                        ev.synchronize()
            self._stage = [[torch.empty(max(need, 4096), dtype=torch.int64).pin_memory(), None] for _ in range(4)]
        self._turn = (self._turn + 1) % len(self._stage)
        slot = self._stage[self._turn]
        if slot[1] is not None:
            slot[1].synchronize()
        hbuf = slot[0]
        hbuf[:need].copy_(torch.from_numpy(offs.reshape(-1)))
        d_offs = torch.empty(need, dtype=torch.int64, device=dev)
        d_offs.copy_(hbuf[:need], non_blocking=True)
        if slot[1] is None:
            slot[1] = torch.cuda.Event()
        slot[1].record()
        d_offs = d_offs.view(5, b + 1)
        # one int64 block [edge_index | BU_edge_index | batch | rootindex | y], one int32 block [x ptr | x col]
        blk = torch.empty(2 * e_td + 2 * e_bu + n + 2 * b, dtype=torch.int64, device=dev)
        o = 0
        ei = blk[o:o + 2 * e_td].view(2, e_td); o += 2 * e_td
        bu = blk[o:o + 2 * e_bu].view(2, e_bu); o += 2 * e_bu
        batch = blk[o:o + n]; o += n
        root = blk[o:o + b]; o += b
        y = blk[o:o + b]
        iblk = torch.empty(n + 1 + nnz, dtype=torch.int32, device=dev)
        if b == 0:
            iblk.zero_()
        ox_ptr, ox_col = iblk[:n + 1], iblk[n + 1:]
        ox_val = torch.empty(nnz, dtype=torch.float32, device=dev)
        row = lambda r: d_offs[r].data_ptr()  # noqa: E731
        check(lib().bigcn_assemble_batch(_p(self.node_ptr), _p(self.edge_ptr), _p(self.edge_src), _p(self.edge_dst),
                                         _p(self.x_ptr), _p(self.x_col), _p(self.x_val), _p(self.root_local), _p(self.y),
                                         row(0), row(1), row(2), row(3), row(4), b, e_td, e_bu,
                                         int(seed) & ((1 << 64) - 1), _p(ei), _p(bu), _p(batch), _p(root), _p(y),
                                         _p(ox_ptr), _p(ox_col), _p(ox_val), _stream()), "assemble_batch")
        out = Batch(x=SparseX(ox_ptr, ox_col, ox_val, (n, self.in_feats)), edge_index=ei, BU_edge_index=bu,
                    batch=batch, rootindex=root, y=y)
        out._keep = d_offs
        return out
