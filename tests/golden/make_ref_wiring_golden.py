"""Golden vectors from the REFERENCE'S OWN MODEL CODE, run on CPU in the build container:

    python tests/golden/make_ref_wiring_golden.py          # needs /root/reference; writes ref_wiring.npz

What runs: ``model/Twitter/BiGCN_Twitter.py`` imported unmodified as a module (classes TDrumorGCN / BUrumorGCN /
BiGCN, :19-131) and the class definitions of ``model/Weibo/BiGCN_Weibo.py`` (:16-89; that file trains at import,
so only its ``import`` and ``class`` statements are executed, straight from its AST -- nothing is copied).

What is shimmed: the two third-party wheels the reference imports and this image does not have
(``torch_geometric``, ``torch_scatter``; oracle/__init__.py).  The shim below is NOT the oracle: GCNConv is the
published dense formula  D^-1/2 (A + I) D^-1/2 (X W^T) + b  on an explicit [N, N] matrix (degree by target, one
weight-1 self-loop per node, duplicate edges counted), scatter_mean is index_add_ / count.  So the fixture pins
everything the reference's own files decide -- which tensor is x1 / x2, what ``copy.copy`` does to the autograd
graph, the order of relu / dropout / cat, root_extend, BU-before-TD concatenation, the head, where torch's
generator is consumed in training mode -- and leaves only the inside of the two library calls to the restatement
(parity of THAT stays unpinned).  tests/test_oracle.py holds the oracle to these vectors, tests/test_gpu_parity.py
the CUDA path.
"""
import ast
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def f32(a):
    return np.ascontiguousarray(a.detach().numpy(), dtype=np.float32)


# ---- stand-ins for the absent wheels (dense, independent of oracle/) -------------------------------------------
class ShimGCNConv(torch.nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin = torch.nn.Linear(in_channels, out_channels, bias=False)
        self.bias = torch.nn.Parameter(torch.zeros(out_channels))
        a = (6.0 / (in_channels + out_channels)) ** 0.5
        with torch.no_grad():
            self.lin.weight.uniform_(-a, a)

    def forward(self, x, edge_index):
        n = x.size(0)
        src, dst = edge_index[0], edge_index[1]
        keep = src != dst                                   # existing self-loops are replaced by the one added below
        adj = torch.zeros(n, n, dtype=x.dtype)
        adj.index_put_((dst[keep], src[keep]), torch.ones(int(keep.sum()), dtype=x.dtype), accumulate=True)
        adj = adj + torch.eye(n, dtype=x.dtype)
        dis = adj.sum(1).pow(-0.5)                          # in-degree including the self-loop
        return dis.view(-1, 1) * (adj @ (dis.view(-1, 1) * self.lin(x))) + self.bias


def shim_scatter_mean(src, index, dim=0):
    assert dim == 0
    nb = int(index.max()) + 1
    out = torch.zeros(nb, src.size(1), dtype=src.dtype).index_add_(0, index, src)
    cnt = torch.bincount(index, minlength=nb).clamp(min=1).to(src.dtype)
    return out / cnt.view(-1, 1)


def install_shims():
    tg = types.ModuleType("torch_geometric")
    tg_nn = types.ModuleType("torch_geometric.nn")
    tg_data = types.ModuleType("torch_geometric.data")
    tg_nn.GCNConv = ShimGCNConv
    tg_data.Data = type("Data", (), {"__init__": lambda self, **kw: self.__dict__.update(kw)})
    tg_data.DataLoader = object
    tg.nn, tg.data = tg_nn, tg_data
    ts = types.ModuleType("torch_scatter")
    ts.scatter_mean = shim_scatter_mean
    sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": tg_nn, "torch_geometric.data": tg_data,
                        "torch_scatter": ts})


def load_reference_models():
    install_shims()
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)                                           # the files append os.getcwd() to sys.path
    try:
        import importlib
        tw = importlib.import_module("model.Twitter.BiGCN_Twitter")
        path = os.path.join(REF, "model/Weibo/BiGCN_Weibo.py")
        tree = ast.parse(open(path).read(), path)
        keep = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom, ast.ClassDef))]
        ns = {"__name__": "ref_weibo_classes", "device": torch.device("cpu")}
        exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    finally:
        os.chdir(cwd)
    return tw.BiGCN, ns["Net"]


def main():
    from bigcn_b200.data import Batch
    RefBiGCN, RefNet = load_reference_models()
    small = json.load(open(os.path.join(HERE, "bigcn_small.json")))     # the 5-tree edge-case batch of make_golden.py
    K, N = small["K"], small["N"]
    x = torch.tensor([float.fromhex(h) for h in small["x"]], dtype=torch.float32).view(N, K)
    out = {"note": np.array("reference model code over dense shims of torch_geometric / torch_scatter; batch = "
                            "bigcn_small.json; torch " + torch.__version__)}
    # twitter: hidden 64 (what the CUDA path is built for); weibo: hidden 16 (oracle only), keeps the file small
    for name, ctor, C, seed in (("twitter", lambda: RefBiGCN(K, 64, 64, torch.device("cpu")), 4, 21),
                                ("weibo", lambda: RefNet(K, 16, 16), 2, 22)):
        data = Batch(x=x, edge_index=torch.tensor(small["edge_index"]), BU_edge_index=torch.tensor(small["BU_edge_index"]),
                     batch=torch.tensor(small["batch"]), rootindex=torch.tensor(small["rootindex"]),
                     y=torch.tensor(small["y"]) % C)
        torch.manual_seed(seed)
        m = ctor()
        with torch.no_grad():
            for p in m.parameters():
                if p.dim() == 1:
                    p.uniform_(-0.1, 0.1)
        out[f"{name}/y"] = data.y.numpy()
        for k, v in m.state_dict().items():
            out[f"{name}/state/{k}"] = f32(v)
        for mode in ("eval", "train"):
            m.train(mode == "train")
            m.zero_grad()
            torch.manual_seed(1000 + seed)                  # consumed by F.dropout in training mode (TD first, then BU)
            logp = m(data)
            loss = torch.nn.functional.nll_loss(logp, data.y)           # BiGCN_Twitter.py:186
            loss.backward()
            out[f"{name}/{mode}/dropout_seed"] = np.array(1000 + seed)
            out[f"{name}/{mode}/logp"] = f32(logp)
            out[f"{name}/{mode}/loss"] = f32(loss)
            for k, p in m.named_parameters():
                out[f"{name}/{mode}/grad/{k}"] = f32(p.grad)
    np.savez(os.path.join(HERE, "ref_wiring.npz"), **out)
    print("ref_wiring.npz written:", len(out), "arrays")


if __name__ == "__main__":
    main()
