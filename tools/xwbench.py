#!/usr/bin/env python
"""A/B of the X*W scan variants (bigcn_debug_set knob 3) on three Twitter16-shaped batches (625 MB each):
ms per launch of bigcn_xw_sparse without the CSC build, and GB/s of algorithmic bytes."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bigcn_b200 import _lib as L  # noqa: E402
from bigcn_b200.data import make_batch_shard  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    lib = L.lib()
    lib.bigcn_debug_set.argtypes = [C.c_int, C.c_int]
    lib.bigcn_debug_set.restype = None
    K = 5000
    xs = [make_batch_shard("twitter16", 128, 1000 + i)[0].x.to(dev) for i in range(3)]
    nodes = [int(x.shape[0]) for x in xs]
    w0, w1 = torch.randn(64, K, device=dev), torch.randn(64, K, device=dev)
    ys = [torch.empty(n, 128, device=dev) for n in nodes]
    ws = [torch.empty(lib.bigcn_xsparse_workspace_bytes(n, K), dtype=torch.uint8, device=dev) for n in nodes]
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def fn(i):
        j = i % 3
        L.check(lib.bigcn_xw_sparse(xs[j].data_ptr(), nodes[j], K, w0.data_ptr(), w1.data_ptr(), K, ys[j].data_ptr(), 128, 0,
                                    flags.data_ptr(), ws[j].data_ptr(), ws[j].numel(), st))
    ref = None
    for knob in [int(a) for a in sys.argv[1:]] or [0, 7, 8, 9, 10, 2, 3]:
        lib.bigcn_debug_set(3, knob)
        for i in range(6):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(30):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        by = sum(nodes) / 3 * (K * 4 + 512) + K * 512
        if ref is None:
            ref = [y.clone() for y in ys]
        same = all(torch.equal(a, b) for a, b in zip(ref, ys))
        print(f"knob3={knob}: {ms * 1e3:.1f} us  {by / ms / 1e6:.0f} GB/s  bit-identical to the first variant: {same}", flush=True)
    lib.bigcn_debug_set(3, 0)


if __name__ == "__main__":
    main()
