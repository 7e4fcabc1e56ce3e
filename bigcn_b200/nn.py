"""Drop-in modules for the reference's hot path.

Same class names, constructor signatures, sub-module names and state_dict keys as
/root/reference/model/Twitter/BiGCN_Twitter.py:19-131 (``TDrumorGCN``, ``BUrumorGCN``,
``BiGCN``) and /root/reference/model/Weibo/BiGCN_Weibo.py:16-89 (``Net``), and a
``GCNConv`` that stands in for ``torch_geometric.nn.GCNConv`` on this path
(``conv(x, edge_index)`` as called at explain_PHEME.py:95,132).  ``forward(data)``
is duck-typed on ``data.x, data.edge_index, data.BU_edge_index, data.batch,
data.rootindex``.  All arithmetic runs in libbigcn_b200.so; nothing here falls back
to PyTorch or the CPU.
"""
from __future__ import annotations

import math

import torch

from . import _lib as L
from .ops import FeaturesFunction, GCNConvFunction, GCNConvWeightedFunction, HeadFunction, raise_on_flags

H = L.H


class _Lin(torch.nn.Module):
    """``GCNConv.lin`` of PyG 2.x: Linear(in, out, bias=False), glorot-uniform weight [out,in]."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels))
        a = math.sqrt(6.0 / (in_channels + out_channels))
        with torch.no_grad():
            self.weight.uniform_(-a, a)


class GCNConv(torch.nn.Module):
    """GCNConv(in, 64): state_dict keys ``lin.weight`` [64,in] and ``bias`` [64] (PyG 2.x);
    PyG-1.3.2 checkpoints (``weight`` [in,64]) are accepted on load."""

    def __init__(self, in_channels, out_channels, deg_by="target", gemm_mode="fp32"):
        super().__init__()
        if out_channels != H:
            raise L.BigcnError(f"bigcn_b200 kernels are specialised for out_channels = {H} "
                               f"(the reference's only configuration), got {out_channels}")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = _Lin(in_channels, out_channels)
        self.bias = torch.nn.Parameter(torch.zeros(out_channels))
        self.deg_by, self.gemm_mode = deg_by, gemm_mode

    def forward(self, x, edge_index, edge_weight=None):
        if edge_weight is None and torch.is_tensor(x) and x.requires_grad:
            # x is an activation (EBGCN's conv2 input is batch-normalised): the generic form returns dx
            edge_weight = torch.ones(edge_index.shape[1], dtype=torch.float32, device=x.device)
        if edge_weight is not None:      # EBGCN.py:84,181
            return GCNConvWeightedFunction.apply(x, edge_index, edge_weight, self.lin.weight, self.bias, self.deg_by,
                                                 self.gemm_mode)
        return GCNConvFunction.apply(x, edge_index, self.lin.weight, self.bias, self.deg_by, self.gemm_mode)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kw):
        old = prefix + "weight"      # torch_geometric 1.3.2: weight [in,out]
        if old in state_dict and prefix + "lin.weight" not in state_dict:
            state_dict[prefix + "lin.weight"] = state_dict.pop(old).t().contiguous()
        super()._load_from_state_dict(state_dict, prefix, *args, **kw)


class _RumorGCN(torch.nn.Module):
    _dir = L.DIR_TD

    def __init__(self, in_feats, hid_feats, out_feats, device=None, deg_by="target", gemm_mode="fp32"):
        super().__init__()
        if hid_feats != H or out_feats != H:
            raise L.BigcnError(f"bigcn_b200 kernels are specialised for hid_feats = out_feats = {H}")
        self.conv1 = GCNConv(in_feats, hid_feats, deg_by, gemm_mode)
        self.conv2 = GCNConv(hid_feats + in_feats, out_feats, deg_by, gemm_mode)
        self.device = device
        self.p = 0.5                        # F.dropout default (BiGCN_Twitter.py:54)
        self.deg_by, self.gemm_mode = deg_by, gemm_mode
        self.seed = torch.initial_seed() & ((1 << 63) - 1)
        self._calls = 0
        self.node_id_base = 0
        self.last_flags = None

    def _conv_params(self):
        return (self.conv1.lin.weight, self.conv1.bias, self.conv2.lin.weight, self.conv2.bias)

    def _opts(self, dir_mask):
        seed = (self.seed + self._calls) & ((1 << 64) - 1)
        if self.training:
            self._calls += 1
        self.last_seed = seed
        # want_grad: autograd.Function.forward cannot see torch.no_grad() (needs_input_grad is True for
        # parameters either way); inference then skips the capture / column sort of x that only dW1 needs
        return dict(training=self.training, p=self.p, seed=seed, deg_by=self.deg_by,
                    gemm_mode=self.gemm_mode, dir_mask=dir_mask, node_id_base=self.node_id_base,
                    want_grad=torch.is_grad_enabled())

    def forward(self, data):
        none4 = (None,) * 4
        params = self._conv_params() + none4 if self._dir == L.DIR_TD else none4 + self._conv_params()
        feat = FeaturesFunction.apply(data.x, data.edge_index, data.BU_edge_index, data.batch,
                                      data.rootindex, self._opts(self._dir), *params)
        return feat[:, 2 * H:] if self._dir == L.DIR_TD else feat[:, :2 * H]


class TDrumorGCN(_RumorGCN):
    """BiGCN_Twitter.py:19-67."""
    _dir = L.DIR_TD


class BUrumorGCN(_RumorGCN):
    """BiGCN_Twitter.py:70-114."""
    _dir = L.DIR_BU


class BiGCN(torch.nn.Module):
    """BiGCN_Twitter.py:117-131: ``BiGCN(in_feats, hid_feats, out_feats, device)`` -> log-probs [B,4]."""

    def __init__(self, in_feats, hid_feats, out_feats, device=None, num_classes=4, deg_by="target",
                 gemm_mode="fp32", validate="lazy"):
        super().__init__()
        self.TDrumorGCN = TDrumorGCN(in_feats, hid_feats, out_feats, device, deg_by, gemm_mode)
        self.BUrumorGCN = BUrumorGCN(in_feats, hid_feats, out_feats, device, deg_by, gemm_mode)
        self.fc = torch.nn.Linear((out_feats + hid_feats) * 2, num_classes)
        self.device = device
        self.validate = validate            # "lazy": check the previous call's flags; "sync"; "off"
        self.last_flags = None

    @property
    def gemm_mode(self):
        return self.TDrumorGCN.gemm_mode

    @gemm_mode.setter
    def gemm_mode(self, mode):
        for m in (self.TDrumorGCN, self.BUrumorGCN):
            m.gemm_mode = mode

    def check_inputs(self):
        """Raise if the last forward saw an invalid graph (synchronises)."""
        if self.last_flags is not None:
            flags, self.last_flags = self.last_flags, None
            raise_on_flags(flags)

    def forward(self, data):
        if self.validate == "lazy":
            self.check_inputs()
        td = self.TDrumorGCN
        opts = td._opts(L.DIR_TD | L.DIR_BU)
        flags = torch.zeros(1, dtype=torch.int32, device=data.x.device)
        opts["flags"] = flags
        feat = FeaturesFunction.apply(data.x, data.edge_index, data.BU_edge_index, data.batch,
                                      data.rootindex, opts, *td._conv_params(),
                                      *self.BUrumorGCN._conv_params())
        out = HeadFunction.apply(feat, self.fc.weight, self.fc.bias)
        self.last_flags = flags
        if self.validate == "sync":
            self.check_inputs()
        return out


class Net(BiGCN):
    """BiGCN_Weibo.py:76-89: ``Net(in_feats, hid_feats, out_feats)`` -> log-probs [B,2]."""

    def __init__(self, in_feats, hid_feats, out_feats, num_classes=2, **kw):
        super().__init__(in_feats, hid_feats, out_feats, None, num_classes=num_classes, **kw)
