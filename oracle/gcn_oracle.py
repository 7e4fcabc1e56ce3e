"""Restatement of the library routines the reference's hot path calls.

TEST INFRASTRUCTURE (see oracle/__init__.py; parity unpinned).

Each function names the reference call site it stands in for and the
third-party routine whose published behaviour it restates
(torch_geometric 2.x / torch_scatter; pinned versions in
/root/reference/readme.md:21-30 are 1.3.2 / 1.4.0, kept as ``deg_by="source"``).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# gcn_norm  (reached from BiGCN_Twitter.py:42,56,92,105 through GCNConv.forward;
#            called directly at explain_PHEME.py:62-63)
# --------------------------------------------------------------------------
def add_remaining_self_loops(edge_index: torch.Tensor, edge_weight, fill_value, num_nodes: int):
    """[PyG utils.loop.add_remaining_self_loops] drop nothing, move existing
    self-loops to the appended block: result = [non-loop edges in order | (i,i) for i<N].
    An existing self-loop's weight is carried into the appended loop."""
    row, col = edge_index[0], edge_index[1]
    mask = row != col
    loop_index = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    loop_index = loop_index.unsqueeze(0).repeat(2, 1)
    if edge_weight is not None:
        loop_weight = edge_weight.new_full((num_nodes,), fill_value)
        inv = ~mask
        loop_weight[row[inv]] = edge_weight[inv]
        edge_weight = torch.cat([edge_weight[mask], loop_weight], dim=0)
    edge_index = torch.cat([edge_index[:, mask], loop_index], dim=1)
    return edge_index, edge_weight


def gcn_norm(edge_index: torch.Tensor, num_nodes: int, edge_weight=None,
             deg_by: str = "target", dtype=torch.float32):
    """[PyG nn.conv.gcn_conv.gcn_norm, improved=False, add_self_loops=True,
    flow='source_to_target'].  Returns (edge_index', w) with
    w_e = deg^-1/2[row] * w * deg^-1/2[col]; deg is summed at ``col`` (2.x) or
    at ``row`` (1.3.2, deg_by='source')."""
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype, device=edge_index.device)
    edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, 1.0, num_nodes)
    row, col = edge_index[0], edge_index[1]
    idx = col if deg_by == "target" else row
    deg = torch.zeros(num_nodes, dtype=edge_weight.dtype).scatter_add_(0, idx, edge_weight)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    w = dis[row] * edge_weight * dis[col]
    return edge_index, w


def glorot_(w: torch.Tensor):
    """[PyG nn.inits.glorot] U(-a, a), a = sqrt(6/(fan_in+fan_out))."""
    a = math.sqrt(6.0 / (w.size(-2) + w.size(-1)))
    with torch.no_grad():
        w.uniform_(-a, a)
    return w


class GCNConv(torch.nn.Module):
    """[PyG nn.GCNConv 2.x] lin = Linear(in, out, bias=False) glorot; bias zeros.
    forward: x' = lin(x); out[i] = sum_{(j->i) in COO'} w_e x'[j]; out += bias.
    Stands in for the objects built at BiGCN_Twitter.py:22-23,73-74."""

    def __init__(self, in_channels: int, out_channels: int, deg_by: str = "target"):
        super().__init__()
        self.lin = torch.nn.Linear(in_channels, out_channels, bias=False)
        self.bias = torch.nn.Parameter(torch.zeros(out_channels))
        self.deg_by = deg_by
        glorot_(self.lin.weight)

    def forward(self, x, edge_index, edge_weight=None):
        n = x.size(0)
        ei, w = gcn_norm(edge_index, n, edge_weight, self.deg_by, x.dtype)
        h = self.lin(x)
        return propagate_sum(h, ei, w) + self.bias


def propagate_sum(h, ei, w):
    """message w_e * h[row_e], sum-aggregated at col_e, in COO' order (CPU index_add_)."""
    msg = w.view(-1, 1) * h.index_select(0, ei[0])
    return torch.zeros(h.size(0), h.size(1), dtype=h.dtype).index_add_(0, ei[1], msg)


# --------------------------------------------------------------------------
# scatter_mean (BiGCN_Twitter.py:65,113; BiGCN_Weibo.py:43,73)
# --------------------------------------------------------------------------
def scatter_mean(src: torch.Tensor, index: torch.Tensor, dim_size: int | None = None):
    """[torch_scatter.scatter_mean, dim=0] sum, count.clamp(min=1), true divide."""
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    out = torch.zeros(dim_size, src.size(1), dtype=src.dtype).index_add_(0, index, src)
    cnt = torch.bincount(index, minlength=dim_size).clamp(min=1).to(src.dtype)
    return out / cnt.view(-1, 1)


# --------------------------------------------------------------------------
# graph_prep: the integer structure the CUDA path emits instead of COO'
# (canonical order = COO' order: in-edges of a node in edge-list order).
# --------------------------------------------------------------------------
def graph_prep(edge_index: np.ndarray, num_nodes: int, batch: np.ndarray, num_graphs: int,
               deg_by: str = "target"):
    """Bit-exact spec for ``bigcn_graph_prep`` (include/bigcn_b200.h).

    Returns dict with int32 ``in_ptr[N+1]``, ``in_idx[E']`` (sources of the
    in-edges of each node, edge-list order), ``out_ptr[N+1]``, ``out_idx[E']``
    (targets of the out-edges, edge-list order), ``deg[N]`` (incl. the unit
    self-loop), fp32 ``dis[N]``, ``rowsum[N]`` (sum of A-hat row i in COO'
    order: in-edges then loop), int32 ``node_ptr[B+1]``, ``n_edges`` (E')."""
    ei = np.asarray(edge_index, dtype=np.int64).reshape(2, -1)
    row, col = ei[0], ei[1]
    keep = row != col
    row, col = row[keep], col[keep]
    n = int(num_nodes)
    indeg = np.bincount(col, minlength=n).astype(np.int64)
    outdeg = np.bincount(row, minlength=n).astype(np.int64)
    in_ptr = np.zeros(n + 1, np.int64); in_ptr[1:] = np.cumsum(indeg)
    out_ptr = np.zeros(n + 1, np.int64); out_ptr[1:] = np.cumsum(outdeg)
    order_in = np.argsort(col, kind="stable")
    order_out = np.argsort(row, kind="stable")
    in_idx = row[order_in]
    out_idx = col[order_out]
    deg = (indeg if deg_by == "target" else outdeg) + 1
    # torch CPU deg.pow(-0.5) == 1.0f / sqrtf(deg) with IEEE sqrt and divide (hard part 7)
    dis = (np.float32(1.0) / np.sqrt(deg.astype(np.float32))).astype(np.float32)
    rowsum = np.zeros(n, np.float32)
    for i in range(n):  # COO' order: in-edges (edge-list order), then the self-loop
        acc = np.float32(0.0)
        for j in in_idx[in_ptr[i]:in_ptr[i + 1]]:
            acc = np.float32(acc + np.float32(dis[j] * dis[i]))
        acc = np.float32(acc + np.float32(dis[i] * dis[i]))
        rowsum[i] = acc
    b = np.asarray(batch, dtype=np.int64)
    node_ptr = np.searchsorted(b, np.arange(num_graphs + 1), side="left").astype(np.int32)
    return dict(in_ptr=in_ptr.astype(np.int32), in_idx=in_idx.astype(np.int32),
                out_ptr=out_ptr.astype(np.int32), out_idx=out_idx.astype(np.int32),
                deg=deg.astype(np.int32), dis=dis, rowsum=rowsum, node_ptr=node_ptr,
                n_edges=int(keep.sum()))


# --------------------------------------------------------------------------
# Philox4x32-10 counter-based RNG: the dropout-mask spec of the CUDA path.
# (The reference draws F.dropout masks from torch's global generator,
#  BiGCN_Twitter.py:54; bit-matching that stream is not a goal -- train-mode
#  parity injects THIS mask into the oracle.)
# --------------------------------------------------------------------------
_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11).  Inputs are uint32
    arrays/scalars (broadcast); returns four uint32 arrays."""
    c0 = np.asarray(c0, np.uint64); c1 = np.asarray(c1, np.uint64)
    c2 = np.asarray(c2, np.uint64); c3 = np.asarray(c3, np.uint64)
    k0 = int(k0) & 0xFFFFFFFF; k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def dropout_threshold(p: float) -> int:
    """keep iff r >= thresh, thresh = round(p * 2^32) clamped to uint32."""
    return min(0xFFFFFFFF, max(0, int(round(p * 4294967296.0))))


def dropout_keep_mask(seed: int, stream: int, node_ids: np.ndarray, n_cols: int, p: float) -> np.ndarray:
    """Mask spec: keep[i, c] = philox(ctr=(c>>2, node_lo, node_hi, stream), key=seed)[c&3] >= thresh.
    ``node_ids`` are GLOBAL node ids (so masks do not depend on the world size);
    ``stream`` is 0 for the TD direction, 1 for BU.  Column c indexes the
    concatenated [h1 | root_extend] tensor of BiGCN_Twitter.py:51-54."""
    node_ids = np.asarray(node_ids, np.uint64).reshape(-1, 1)
    nblk = (n_cols + 3) // 4
    blk = np.arange(nblk, dtype=np.uint64).reshape(1, -1)
    r = philox4x32_10(blk, node_ids & _MASK32, node_ids >> np.uint64(32), np.uint64(stream),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    bits = np.stack(np.broadcast_arrays(*r), axis=-1).reshape(node_ids.shape[0], nblk * 4)[:, :n_cols]
    return bits >= np.uint32(dropout_threshold(p))
