#!/usr/bin/env python
"""Kernel micro-benchmarks on forests far larger than L2 (run on the GPU box): propagate in both
CSR orientations (parent gather / child segment-sum), with and without the hub-row split lists,
on uniform and hub-heavy forests.  Usage: python tools/kbench.py [--forest=name]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bigcn_b200 import _lib as L, ops  # noqa: E402
from bigcn_b200.data import make_device_forest  # noqa: E402

FORESTS = {"uniform256": (16384, 256, 1.0), "hub256": (16384, 256, 3.0), "hub4096": (1024, 4096, 3.0),
           "batch128": (128, 256, 3.0)}


def timeit(fn, iters=10):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    only_forest = [a.split("=")[1] for a in sys.argv[1:] if a.startswith("--forest=")]
    dev = torch.device("cuda", 0)
    lib = L.lib()
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    for fname, (n_trees, per, skew) in FORESTS.items():
        if only_forest and fname not in only_forest:
            continue
        n = n_trees * per
        f = make_device_forest(n_trees, per, dev, seed=3, skew=skew)
        hs = [torch.randn(n, 64, device=dev) for _ in range(2)]
        outb = torch.empty(n, 64, device=dev)
        bias = torch.randn(64, device=dev)
        e = int(f.edge_index.shape[1])
        byt = n * (2 * 256 + 8) + 4 * e + 260
        # readout (scatter_mean + root rows): N*256 in, B*512 out
        node_ptr = torch.arange(0, n + 1, per, dtype=torch.int32, device=dev)
        scr = torch.empty(lib.bigcn_readout_scratch_floats(n, n_trees), device=dev)
        feat = torch.empty(n_trees, 128, device=dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)

        def ro(i):
            L.check(lib.bigcn_readout(hs[i % 2].data_ptr(), hs[(i + 1) % 2].data_ptr(), node_ptr.data_ptr(),
                                      f.rootindex.data_ptr(), n, n_trees, feat.data_ptr(), 128, None, scr.data_ptr(),
                                      flags.data_ptr(), st))
        ms = timeit(ro)
        rbyt = n * 256 + n_trees * 256 + 4 * (n_trees + 1) + n_trees * 512
        gbs = rbyt / ms / 1e6
        res[f"{fname}/readout"] = dict(ms=round(ms, 4), gbs=round(gbs, 1), frac=round(gbs / peak, 3))
        print(f"{fname:11s} readout  {ms:8.4f} ms {gbs:8.1f} GB/s {gbs / peak:6.3f} of measured peak", flush=True)
        outs = {}
        for mode in ("split", "seq"):
            graphs, _, _ = ops.graph_prep([f.edge_index, f.BU_edge_index], n, f.batch, n_trees, rowsum=False,
                                          long_rows=(mode == "split"))
            maxdeg = int(graphs[1]["deg"].max().item())
            for dname, g in (("td", graphs[0]), ("bu", graphs[1])):
                lng = g["in_long"].data_ptr() if g["in_long"] is not None else None

                def fn(i, g=g, lng=lng):
                    L.check(lib.bigcn_propagate(g["in_ptr"].data_ptr(), g["in_idx"].data_ptr(), g["dis"].data_ptr(), n,
                                                e, lng, hs[i % 2].data_ptr(), 64, bias.data_ptr(), 1,
                                                outb.data_ptr(), 64, st))
                fn(0)
                torch.cuda.synchronize()
                outs[(mode, dname)] = outb.clone()
                ms = timeit(fn)
                gbs = byt / ms / 1e6
                err = ""
                if mode == "seq":
                    d = (outs[("split", dname)] - outs[("seq", dname)]).abs().max().item()
                    err = f" max|split-seq|={d:.2e}"
                res[f"{fname}/{dname}/{mode}"] = dict(ms=round(ms, 4), gbs=round(gbs, 1), frac=round(gbs / peak, 3))
                print(f"{fname:11s} {dname} {mode:5s} {ms:8.4f} ms {gbs:8.1f} GB/s {gbs / peak:6.3f} of measured peak"
                      f" maxdeg={maxdeg}{err}", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "kbench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
