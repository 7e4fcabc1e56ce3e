"""torch.library operators of the path: registered, CUDA only, differentiable, and
equal to the oracle when composed the way torch_geometric's GCNConv composes them."""
import pytest
import torch

from oracle import gcn_oracle


def test_ops_are_registered_and_have_no_cpu_kernel():
    import bigcn_b200  # noqa: F401
    names = {"graph_prep", "xw", "xw_wgrad", "propagate", "propagate_transposed", "colsum64", "readout", "readout_backward"}
    assert names <= set(dir(torch.ops.bigcn_b200))
    s = str(torch.ops.bigcn_b200.graph_prep.default._schema)
    assert s.startswith("bigcn_b200::graph_prep(Tensor edge_index, SymInt num_nodes, Tensor? batch, SymInt num_graphs, str deg_by)")
    with pytest.raises(NotImplementedError):
        torch.ops.bigcn_b200.xw(torch.zeros(2, 4), torch.zeros(64, 4), "fp32")
    with pytest.raises(NotImplementedError):
        torch.ops.bigcn_b200.colsum64(torch.zeros(2, 64))


@pytest.mark.gpu
def test_conv_composed_from_ops_matches_oracle_with_autograd():
    from bigcn_b200.data import make_batch
    dev = torch.device("cuda:0")
    K = 48
    b = make_batch("twitter15", 6, seed=12, train=True, in_feats=K)
    n = b.x.shape[0]
    torch.manual_seed(1)
    ref = gcn_oracle.GCNConv(K, 64)
    with torch.no_grad():
        ref.bias.uniform_(-0.3, 0.3)
    w = ref.lin.weight.detach().to(dev).requires_grad_(True)
    bias = ref.bias.detach().to(dev).requires_grad_(True)
    B = torch.ops.bigcn_b200
    for ei in (b.edge_index, b.BU_edge_index):
        for relu in (False, True):
            g = B.graph_prep(ei.to(dev), n, b.batch.to(dev), int(b.rootindex.numel()), "target")
            in_ptr, in_idx, out_ptr, out_idx, deg, dis, rowsum, node_ptr, flags, in_long, out_long = g
            assert int(flags.item()) == 0
            want_ptr = gcn_oracle.graph_prep(ei.numpy(), n, b.batch.numpy(), int(b.rootindex.numel()))
            assert torch.equal(in_ptr.cpu(), torch.from_numpy(want_ptr["in_ptr"]))
            assert torch.equal(node_ptr.cpu(), torch.from_numpy(want_ptr["node_ptr"]))
            h = B.xw(b.x.to(dev), w, "fp32")
            out = B.propagate(h, in_ptr, in_idx, out_ptr, out_idx, dis, ei.shape[1], in_long, out_long, bias, relu)
            o_ref = ref(b.x, ei)
            if relu:
                o_ref = torch.relu(o_ref)
            assert float((out.cpu() - o_ref).abs().max()) <= 1e-5 * float(o_ref.abs().max())
            gout = torch.randn(n, 64, generator=torch.Generator().manual_seed(2))
            ref.zero_grad()
            w.grad = bias.grad = None
            o_ref.backward(gout)
            out.backward(gout.to(dev))
            for a, r in ((w.grad, ref.lin.weight.grad), (bias.grad, ref.bias.grad)):
                assert float((a.cpu() - r).abs().max()) <= 1e-4 * float(r.abs().max())
    # readout = second root-extend + scatter_mean
    h2, h1 = torch.randn(n, 64), torch.randn(n, 64)
    h2r, h2g = h2.clone().requires_grad_(True), h2.to(dev).requires_grad_(True)
    feat = B.readout(h2g, h1.to(dev), node_ptr, b.rootindex.to(dev), b.batch.to(dev))
    want = torch.cat([gcn_oracle.scatter_mean(h2r, b.batch), h1[b.rootindex]], 1)
    assert float((feat.detach().cpu() - want.detach()).abs().max()) <= 1e-6 * float(want.abs().max())
    gf = torch.randn_like(want)
    want.backward(gf)
    feat.backward(gf.to(dev))
    assert float((h2g.grad.cpu() - h2r.grad).abs().max()) <= 1e-6 * float(h2r.grad.abs().max())


@pytest.mark.gpu
def test_opcheck_schema_and_fake():
    """torch.library.opcheck: schema, fake (meta) kernel and autograd registration agree with the real op."""
    from bigcn_b200.data import make_batch
    dev = torch.device("cuda:0")
    b = make_batch("twitter15", 3, seed=1, train=False, in_feats=32)
    n = b.x.shape[0]
    B = torch.ops.bigcn_b200
    g = B.graph_prep(b.edge_index.to(dev), n, b.batch.to(dev), 3, "target")
    w = torch.randn(64, 32, device=dev, requires_grad=True)
    x = b.x.to(dev)
    tests = ("test_schema", "test_faketensor", "test_autograd_registration")
    torch.library.opcheck(B.xw.default, (x, w, "fp32"), test_utils=tests)
    h = torch.randn(n, 64, device=dev, requires_grad=True)
    torch.library.opcheck(B.propagate.default, (h, g[0], g[1], g[2], g[3], g[5], b.edge_index.shape[1], g[9], g[10], None, False),
                          test_utils=tests)
    torch.library.opcheck(B.colsum64.default, (h.detach(),), test_utils=tests)
