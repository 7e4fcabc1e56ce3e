#!/usr/bin/env python
"""The tcgen05 X*W GEMMs and the weight-gradient GEMM on their own (for ncu: k_xw_tc<1>, k_xw_tc<2,false>, k_xw_tc<2,true>,
k_dw_tc<2,true>): one Twitter16-shaped matrix (625 MB) and one PHEME-shaped one (4096 trees, 120 MB), ms per launch."""
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
    st = torch.cuda.current_stream().cuda_stream
    for shape, trees, K in (("twitter16", 128, 5000), ("pheme", 4096, 768)):
        x = make_batch_shard(shape, trees, 1000)[0].x.to(dev)
        n = int(x.shape[0])
        w0, w1 = torch.randn(64, K, device=dev) * 0.02, torch.randn(64, K, device=dev) * 0.02
        y = torch.empty(n, 128, device=dev)
        t = torch.randn(n, 128, device=dev)
        dws = [torch.empty(64, K, device=dev) for _ in range(2)]
        scr = torch.empty(lib.bigcn_xw_scratch_floats(K, 2), device=dev)
        wscr = torch.empty(lib.bigcn_xw_wgrad_scratch_floats(n, K, 2), device=dev)
        byt = n * K * 4 + K * 128 * 4 + n * 128 * 4
        for mode in ("tf32", "tf32x2", "tf32x3"):
            def fn():
                L.check(lib.bigcn_xw(x.data_ptr(), n, K, w0.data_ptr(), w1.data_ptr(), K, y.data_ptr(), 128, L.GEMM_MODE[mode],
                                     scr.data_ptr(), st))
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"{shape} N={n} K={K} xw {mode}: {ms * 1e3:.1f} us  {byt / ms / 1e6:.0f} GB/s  {2 * n * K * 128 / ms / 1e9:.0f} TFLOP/s", flush=True)

        def dfn():
            L.check(lib.bigcn_xw_wgrad(x.data_ptr(), n, K, t.data_ptr(), 2, dws[0].data_ptr(), dws[1].data_ptr(), K,
                                       L.GEMM_MODE["tf32x3"], wscr.data_ptr(), st))
        for _ in range(3):
            dfn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dfn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{shape} N={n} K={K} dw tf32x3: {ms * 1e3:.1f} us  {byt / ms / 1e6:.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
