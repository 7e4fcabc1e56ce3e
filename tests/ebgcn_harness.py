"""TEST HARNESS (not part of the product package): EBGCN (the fork's second GCN model,
/root/reference/model/Twitter/EBGCN.py) assembled on this library's edge-weighted GCNConv -- SURVEY.md 8f N3 asks for
the weighted propagate only; this scaffolding (edge-inference sub-networks, BatchNorm, KL loss: plain torch modules
as in the reference) exists to exercise it end to end.  The module-level shape: ``TDrumorGCN(args)``, ``BUrumorGCN(args)``,
``EBGCN(args)`` with the reference's sub-module names (state_dict keys match) and the same
``forward(data) -> (log-probs, TD_edge_loss, BU_edge_loss)``.

What runs where:
  conv1 / conv2 (torch_geometric GCNConv, EBGCN.py:26-27,84; conv2 with ``edge_weight=edge_pred``)
        -> bigcn_b200.GCNConv: graph prep, X W^T, (edge-weighted) propagate and their gradients
           (weights, bias, edge weights, and x for conv2) in libbigcn_b200.so
  the Python root-extend loops (:71-76, :86-89) -> one gather x1[rootindex[batch]]
  scatter_mean (:92)                            -> torch.ops.bigcn_b200.readout (+ its backward)
  edge inference sub-networks, BatchNorm1d, KL loss (:29-59, :95-117) -> the same torch modules as
        in the reference (they are not part of the GCN path)
The N x (hidden + input) tensor IS materialised here, as in the reference: BatchNorm1d over the
concatenation makes every column dense."""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

from bigcn_b200 import ops
from bigcn_b200.nn import GCNConv


def _create_network(hidden, name):
    """EBGCN.py:33-48: Conv1d(k=1, no bias) -> BatchNorm1d -> LeakyReLU, then Conv1d(hidden -> 1, k=1)."""
    layers = OrderedDict()
    layers[name + "conv0"] = torch.nn.Conv1d(hidden, hidden, kernel_size=1, bias=False)
    layers[name + "norm0"] = torch.nn.BatchNorm1d(num_features=hidden)
    layers[name + "relu0"] = torch.nn.LeakyReLU()
    layers[name + "conv_out"] = torch.nn.Conv1d(hidden, 1, kernel_size=1)
    return layers


class _EdgeRumorGCN(torch.nn.Module):
    _edge_key = "edge_index"
    _flag = "edge_infer_td"

    def __init__(self, args):
        super().__init__()
        self.args = args
        hid = args.hidden_features
        self.conv1 = GCNConv(args.input_features, hid)
        self.conv2 = GCNConv(args.input_features + hid, args.output_features)
        self.device = getattr(args, "device", None)
        self.sim_network = torch.nn.Sequential(_create_network(hid, "sim_val"))
        self.W_mean = torch.nn.Sequential(_create_network(hid, "W_mean"))
        self.W_bias = torch.nn.Sequential(_create_network(hid, "W_bias"))
        self.B_mean = torch.nn.Sequential(_create_network(hid, "B_mean"))
        self.B_bias = torch.nn.Sequential(_create_network(hid, "B_bias"))
        self.fc1 = torch.nn.Linear(hid, args.edge_num, bias=False)
        self.fc2 = torch.nn.Linear(hid, args.edge_num, bias=False)
        self.dropout = torch.nn.Dropout(args.dropout)
        self.eval_loss = torch.nn.KLDivLoss(reduction="batchmean")
        self.bn1 = torch.nn.BatchNorm1d(hid + args.input_features)

    def forward(self, data):
        x1 = data.x.float()                                             # :62-63
        edge_index = getattr(data, self._edge_key)
        x = self.conv1(x1, edge_index)                                  # :64
        x2 = x.detach()                                                 # copy.copy(x), :65: a detached leaf
        if getattr(self.args, self._flag, False):
            edge_loss, edge_pred = self.edge_infer(x, edge_index)       # :67-70
        else:
            edge_loss, edge_pred = None, None
        n = x1.shape[0]
        num_graphs = int(data.rootindex.numel())
        root_extend = x1[data.rootindex[data.batch]]                    # the loop of :72-76
        x = torch.cat((x, root_extend), 1)                              # :77
        if x.shape[0] != 1:
            x = self.bn1(x)                                             # :79-80
        x = F.relu(x)                                                   # :81
        x = self.conv2(x, edge_index, edge_weight=edge_pred)            # :84
        x = F.relu(x)                                                   # :85
        _, node_ptr, flags = ops.graph_prep([edge_index], n, data.batch, num_graphs, rowsum=False, long_rows=False)
        self.last_flags = flags
        # cat(x, x2[root]) + scatter_mean, :86-92
        return torch.ops.bigcn_b200.readout(x, x2, node_ptr, data.rootindex, data.batch), edge_loss

    def edge_infer(self, x, edge_index):
        """EBGCN.py:95-117, unchanged (including its ``row - 1`` / ``col - 1`` indexing)."""
        row, col = edge_index[0], edge_index[1]
        x_i = x[row - 1].unsqueeze(2)
        x_j = x[col - 1].unsqueeze(1)
        x_ij = torch.abs(x_i - x_j)
        sim_val = self.sim_network(x_ij)
        edge_pred = torch.sigmoid(self.fc1(sim_val))
        w_mean, w_bias = self.W_mean(x_ij), self.W_bias(x_ij)
        b_mean, b_bias = self.B_mean(x_ij), self.B_bias(x_ij)
        logit_mean = w_mean * sim_val + b_mean
        logit_var = torch.relu(torch.log((sim_val ** 2) * torch.exp(w_bias) + torch.exp(b_bias)))
        edge_y = torch.sigmoid(torch.normal(logit_mean, logit_var))
        edge_y = self.fc2(edge_y)
        logp_x = F.log_softmax(edge_pred, dim=-1)
        p_y = F.softmax(edge_y, dim=-1)
        return self.eval_loss(logp_x, p_y), torch.mean(edge_pred, dim=-1).squeeze(1)


class TDrumorGCN(_EdgeRumorGCN):
    """EBGCN.py:23-117."""
    _edge_key, _flag = "edge_index", "edge_infer_td"


class BUrumorGCN(_EdgeRumorGCN):
    """EBGCN.py:120-214."""
    _edge_key, _flag = "BU_edge_index", "edge_infer_bu"


class EBGCN(torch.nn.Module):
    """EBGCN.py:217-233: ``forward(data) -> (log-probs [B, num_class], TD_edge_loss, BU_edge_loss)``."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.TDrumorGCN = TDrumorGCN(args)
        self.BUrumorGCN = BUrumorGCN(args)
        self.fc = torch.nn.Linear((args.hidden_features + args.output_features) * 2, args.num_class)

    def forward(self, data):
        TD_x, TD_edge_loss = self.TDrumorGCN(data)
        BU_x, BU_edge_loss = self.BUrumorGCN(data)
        self.x = torch.cat((BU_x, TD_x), 1)
        out = F.log_softmax(self.fc(self.x), dim=1)
        return out, TD_edge_loss, BU_edge_loss
