"""EBGCN on this library's convs (tests/ebgcn_harness.py) against the same module built from the oracle's
GCNConv / scatter_mean and the reference's Python loops (model/Twitter/EBGCN.py:61-93, 217-233)."""
import copy
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

from oracle import gcn_oracle

pytestmark = pytest.mark.gpu


class OracleDir(torch.nn.Module):
    """EBGCN.py TDrumorGCN / BUrumorGCN restated with the oracle's GCNConv and scatter_mean."""

    def __init__(self, proto, key, flag):
        super().__init__()
        a = proto.args
        self.args, self.key, self.flag = a, key, flag
        self.conv1 = gcn_oracle.GCNConv(a.input_features, a.hidden_features)
        self.conv2 = gcn_oracle.GCNConv(a.input_features + a.hidden_features, a.output_features)
        for name in ("sim_network", "W_mean", "W_bias", "B_mean", "B_bias", "fc1", "fc2", "bn1"):
            setattr(self, name, copy.deepcopy(getattr(proto, name)).cpu())
        self.eval_loss = torch.nn.KLDivLoss(reduction="batchmean")

    edge_infer = None

    def forward(self, data):
        x, edge_index = data.x, getattr(data, self.key)
        x1 = copy.copy(x)
        x = self.conv1(x, edge_index)
        x2 = copy.copy(x)
        edge_pred = None
        if getattr(self.args, self.flag):
            row, col = edge_index[0], edge_index[1]
            x_ij = torch.abs(x[row - 1].unsqueeze(2) - x[col - 1].unsqueeze(1))
            edge_pred = torch.mean(torch.sigmoid(self.fc1(self.sim_network(x_ij))), dim=-1).squeeze(1)
        root_extend = torch.zeros(len(data.batch), x1.size(1), dtype=x1.dtype)
        batch_size = int(max(data.batch)) + 1
        for b in range(batch_size):
            root_extend[data.batch == b] = x1[data.rootindex[b]]
        x = torch.cat((x, root_extend), 1)
        if x.shape[0] != 1:
            x = self.bn1(x)
        x = F.relu(x)
        x = F.relu(self.conv2(x, edge_index, edge_weight=edge_pred))
        root_extend = torch.zeros(len(data.batch), x2.size(1), dtype=x2.dtype)
        for b in range(batch_size):
            root_extend[data.batch == b] = x2[data.rootindex[b]].detach()
        x = torch.cat((x, root_extend), 1)
        return gcn_oracle.scatter_mean(x, data.batch)


@pytest.mark.parametrize("infer", [True, False])
def test_ebgcn_matches_oracle_module(infer, monkeypatch):
    import bigcn_b200
    import ebgcn_harness as ebgcn
    from bigcn_b200.data import make_batch
    dev = torch.device("cuda:0")
    # the edge-inference sub-networks are plain torch modules: keep cuDNN / cuBLAS from running them in TF32,
    # or their 1e-3 noise hides what is compared here
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    K = 40
    args = SimpleNamespace(input_features=K, hidden_features=64, output_features=64, edge_num=2, dropout=0.5, num_class=4,
                           edge_infer_td=infer, edge_infer_bu=infer, device=dev)
    torch.manual_seed(4)
    m = ebgcn.EBGCN(args)
    ref_td = OracleDir(m.TDrumorGCN, "edge_index", "edge_infer_td")
    ref_bu = OracleDir(m.BUrumorGCN, "BU_edge_index", "edge_infer_bu")
    ref_fc = copy.deepcopy(m.fc)
    for r, d in ((ref_td, m.TDrumorGCN), (ref_bu, m.BUrumorGCN)):
        r.conv1.load_state_dict(d.conv1.state_dict())
        r.conv2.load_state_dict(d.conv2.state_dict())
    m = m.to(dev)
    m.train()
    # the oracle runs in fp64: through BatchNorm and the degree normalisation of hub rows (a BU root sums hundreds of
    # weighted in-edges) fp32 autograd on the CPU is itself ~1e-3 away from the truth for the edge-inference parameters
    ref_td, ref_bu, ref_fc = ref_td.double(), ref_bu.double(), ref_fc.double()
    b = make_batch("twitter15", 5, seed=21, train=True, in_feats=K)
    b.x = b.x.double()
    bd = make_batch("twitter15", 5, seed=21, train=True, in_feats=K).to(dev)
    out, tdl, bul = m(bd)
    want = F.log_softmax(ref_fc(torch.cat((ref_bu(b), ref_td(b)), 1)), dim=1)
    assert float((out.detach().cpu().double() - want.detach()).abs().max()) <= 5e-5 * float(want.abs().max())
    if infer:
        assert torch.isfinite(tdl) and torch.isfinite(bul)
    else:
        assert tdl is None and bul is None
    y = b.y
    F.nll_loss(want, y).backward()
    F.nll_loss(out, y.to(dev)).backward()          # the edge losses sample th.normal: left out of the comparison
    gmax = max(float(p.grad.abs().max()) for p in list(ref_td.parameters()) + list(ref_bu.parameters()) if p.grad is not None)
    errs = {}
    for tag, ours, ref in (("TD", m.TDrumorGCN, ref_td), ("BU", m.BUrumorGCN, ref_bu)):
        theirs = dict(ref.named_parameters())
        for name, p in ours.named_parameters():
            q = theirs.get(name)
            if q is None or q.grad is None:
                continue
            # parameters whose exact gradient is zero (a bias in front of BatchNorm or of |h_i - h_j|) hold rounding
            # noise on both sides: errors are measured against the largest gradient of the model, floor 1e-3
            scale = max(float(q.grad.abs().max()), 1e-3 * gmax)
            errs[f"{tag}.{name}"] = float((p.grad.cpu().double() - q.grad).abs().max()) / scale
    print({k: f"{v:.2e}" for k, v in errs.items()})
    assert len(errs) >= (12 if not infer else 20)
    # this library's kernels produce the conv gradients (and, through d edge_weight and dx, what reaches bn1): tight.
    # The edge-inference parameters and conv1.bias get their gradients through torch's own GPU BatchNorm / Conv1d
    # backward over [E, 64, 64] tensors (fp32 reductions; conv1.bias is exactly zero in exact arithmetic): the
    # kernel-side inputs to that chain are pinned by conv_out / fc1, which sit right behind d edge_weight.
    tight = ("conv1.lin.weight", "conv2.lin.weight", "conv2.bias", "bn1.weight", "bn1.bias", "conv_out.weight", "fc1.weight")
    bad = {k: v for k, v in errs.items() if v >= (5e-4 if k.endswith(tight) else 5e-3)}
    assert not bad, bad
    # state_dict keys are the reference's (PyG 2.x): checkpoints interchange
    keys = set(m.state_dict())
    assert {"TDrumorGCN.conv1.lin.weight", "TDrumorGCN.conv2.bias", "BUrumorGCN.sim_network.sim_valconv0.weight",
            "TDrumorGCN.bn1.running_mean", "BUrumorGCN.fc1.weight", "fc.weight"} <= keys
