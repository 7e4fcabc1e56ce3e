"""CPU restatement of the reference's BiGCN modules (TEST INFRASTRUCTURE).

Follows /root/reference/model/Twitter/BiGCN_Twitter.py:19-131 statement by
statement (the Weibo twin, model/Weibo/BiGCN_Weibo.py:16-89, is the same maths
with a 2-class head).  ``GCNConv`` and ``scatter_mean`` are the restated
library routines of oracle/gcn_oracle.py -- parity unpinned, see
oracle/__init__.py.

Two execution styles share the numerics:
  * ``reference_loops=True``  keeps the reference's Python ``max(data.batch)``
    and per-tree boolean-mask loops (BiGCN_Twitter.py:46-50,59-62).  This is
    what the CPU baseline times, because it is what the reference executes.
  * ``reference_loops=False`` replaces them by the equivalent gather
    ``x1[rootindex[batch]]`` so parity tests run in seconds.

Train-mode parity: ``forward(data, keep=...)`` takes an explicit keep-mask
[N, hid+in] per direction instead of drawing one from torch's generator
(BiGCN_Twitter.py:54), because bit-matching torch's RNG stream is not a goal.
"""
from __future__ import annotations

import copy

import torch
import torch.nn.functional as F

from .gcn_oracle import GCNConv, scatter_mean


def _dropout(x, p, training, keep):
    if not training:
        return x
    if keep is None:
        return F.dropout(x, p=p, training=True)
    scale = torch.tensor(1.0, dtype=torch.float32) / torch.tensor(1.0 - p, dtype=torch.float32)  # fp32 1/(1-p), as F.dropout
    return x * (keep.to(x.dtype) * scale.to(x.dtype))


class _RumorGCN(torch.nn.Module):
    """One direction; BiGCN_Twitter.py:19-67 (TD) and :70-114 (BU)."""

    edge_attr = "edge_index"

    def __init__(self, in_feats, hid_feats, out_feats, device=None, deg_by="target",
                 reference_loops=False):
        super().__init__()
        self.conv1 = GCNConv(in_feats, hid_feats, deg_by)             # :22 / :73
        self.conv2 = GCNConv(hid_feats + in_feats, out_feats, deg_by)  # :23 / :74
        self.device = device
        self.reference_loops = reference_loops
        self.p = 0.5                                                   # F.dropout default, :54

    def _root_extend(self, src, data):
        if self.reference_loops:                                       # :45-50 / :59-62
            root_extend = torch.zeros(len(data.batch), src.size(1))
            batch_size = max(data.batch) + 1
            for num_batch in range(batch_size):
                index = torch.eq(data.batch, num_batch)
                root_extend[index] = src[data.rootindex[num_batch]]
            return root_extend
        return src[data.rootindex[data.batch]]

    def forward(self, data, keep=None):
        x, edge_index = data.x, getattr(data, self.edge_attr)         # :27 / :78
        x1 = copy.copy(x.to(self.conv1.lin.weight.dtype))              # :28 x.float(); fp64 when the oracle is .double() (truth mode)
        x = x1
        x = self.conv1(x, edge_index)                                  # :42
        x2 = copy.copy(x)                                              # :44  (detached leaf sharing storage)
        if x2.requires_grad and x2.grad_fn is not None:                # torch<2 semantics guard
            x2 = x.detach()
        root_extend = self._root_extend(x1, data)                      # :45-50
        x = torch.cat((x, root_extend), 1)                             # :51
        x = F.relu(x)                                                  # :53
        x = _dropout(x, self.p, self.training, keep)                   # :54
        x = self.conv2(x, edge_index)                                  # :56
        x = F.relu(x)                                                  # :57
        root_extend = self._root_extend(x2, data)                      # :58-62
        x = torch.cat((x, root_extend), 1)                             # :63
        nb = int(data.rootindex.numel())
        x = scatter_mean(x, data.batch, nb)                            # :65
        return x


class TDrumorGCN(_RumorGCN):
    edge_attr = "edge_index"


class BUrumorGCN(_RumorGCN):
    edge_attr = "BU_edge_index"


class BiGCN(torch.nn.Module):
    """BiGCN_Twitter.py:117-131 (4 classes) / BiGCN_Weibo.py:76-89 (``Net``, 2 classes)."""

    def __init__(self, in_feats, hid_feats, out_feats, device=None, num_classes=4,
                 deg_by="target", reference_loops=False):
        super().__init__()
        self.TDrumorGCN = TDrumorGCN(in_feats, hid_feats, out_feats, device, deg_by, reference_loops)
        self.BUrumorGCN = BUrumorGCN(in_feats, hid_feats, out_feats, device, deg_by, reference_loops)
        self.fc = torch.nn.Linear((out_feats + hid_feats) * 2, num_classes)   # :122
        self.device = device

    def forward(self, data, keep_td=None, keep_bu=None):
        TD_x = self.TDrumorGCN(data, keep_td)                          # :126
        BU_x = self.BUrumorGCN(data, keep_bu)                          # :127
        x = torch.cat((BU_x, TD_x), 1)                                 # :128  (BU first)
        x = self.fc(x)                                                 # :129
        x = F.log_softmax(x, dim=1)                                    # :130
        return x


class Net(BiGCN):
    """BiGCN_Weibo.py:76-89."""

    def __init__(self, in_feats, hid_feats, out_feats, **kw):
        kw.setdefault("num_classes", 2)
        super().__init__(in_feats, hid_feats, out_feats, None, **kw)


def make_optimizer(model, lr=5e-4, weight_decay=1e-4):
    """BiGCN_Twitter.py:146-153: Adam, BU convs at lr/5, coupled L2."""
    bu = list(map(id, model.BUrumorGCN.conv1.parameters())) + \
        list(map(id, model.BUrumorGCN.conv2.parameters()))
    base = [p for p in model.parameters() if id(p) not in bu]
    return torch.optim.Adam([
        {"params": base},
        {"params": model.BUrumorGCN.conv1.parameters(), "lr": lr / 5},
        {"params": model.BUrumorGCN.conv2.parameters(), "lr": lr / 5},
    ], lr=lr, weight_decay=weight_decay)
