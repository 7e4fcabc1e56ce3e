"""Feeding dense host features to the sparse device path at more than PCIe speed.

The reference's loader produces a DENSE fp32 ``data.x`` in host memory (Process/dataset.py:64-99;
~99.7 % zeros for the bag-of-words datasets) and ``Batch_data.to(device)`` ships it as is
(BiGCN_Twitter.py:171): 20 KB per node over PCIe.  ``HostFeeder.ship`` moves the same matrix with
both engines at once: the copy engine DMAs the LAST rows dense into a staging buffer in HBM, where
two kernels of this library compact them (bigcn_dense_row_counts / bigcn_dense_rows_to_csr), while
the host threads compact the FIRST rows themselves (bigcn_host_dense_to_csr) and ship only their
non-zeros.  The result is one CSR ``SparseX`` on the device -- what ``forward(data)`` takes in
``gemm_mode='sparse'`` -- and the dense matrix never exists on the device beyond the staging rows.

The split adapts: after every call the feeder compares how long the DMA and the host pass took
and moves the boundary so both finish together.  Nothing here synchronises the stream.
"""
from __future__ import annotations

import time

import torch

from . import _lib as L
from ._lib import check, lib
from .ops import SparseX, host_dense_to_csr, _p, _stream


class HostFeeder:
    def __init__(self, device, in_feats: int, max_nodes: int, n_threads: int = 0, dma_fraction: float = 0.5,
                 adapt: bool = True, slots: int = 2, nnz_per_row_cap: int = 48):
        L.require_device()
        self.device = torch.device(device)
        self.k, self.max_nodes = int(in_feats), int(max_nodes)
        self.n_threads = int(n_threads)
        self.frac, self.adapt = float(dma_fraction), bool(adapt)
        self.cap = self.max_nodes * min(self.k, int(nnz_per_row_cap))
        dev = self.device
        self.slots = []
        for _ in range(max(1, int(slots))):       # double-buffered: a batch in use by a step is not overwritten
            self.slots.append(dict(
                dense=None,
                cnt=torch.empty(self.max_nodes, dtype=torch.int32, device=dev),
                ptr=torch.empty(self.max_nodes + 1, dtype=torch.int32, device=dev),
                col=torch.empty(self.cap, dtype=torch.int32, device=dev),
                val=torch.empty(self.cap, dtype=torch.float32, device=dev),
                host=SparseX(torch.empty(self.max_nodes + 1, dtype=torch.int32, pin_memory=True),
                             torch.empty(self.cap, dtype=torch.int32, pin_memory=True),
                             torch.empty(self.cap, dtype=torch.float32, pin_memory=True), (self.max_nodes, self.k)),
                ev=None))
        self._turn = 0
        self._t_begin, self._loop_ms, self._c = None, 0.0, 0.0
        self.last, self.last_cpu = {}, {}
        self.flags = torch.zeros(1, dtype=torch.int32, device=dev)   # BIGCN_FLAG_X_NOT_SPARSE: see check()

    def check(self):
        """Raise if some shipped matrix had more non-zeros than the CSR arrays hold (synchronises)."""
        from .ops import raise_on_flags
        raise_on_flags(self.flags)

    def _settle(self, s):
        """Fold the previous use of this slot into the split (its events completed long ago)."""
        if s.get("h2d_done") is not None:
            s["h2d_done"].synchronize()        # the pinned CSR buffers of this slot are free again
        ev = s["ev"]
        if ev is None:
            return
        s["ev"] = None
        e0, e1, n_dma, n_host, host_ms, dma_first = ev
        if not e1.query():
            e1.synchronize()
        dma_ms = e0.elapsed_time(e1)
        self.last = dict(dma_ms=dma_ms, host_ms=host_ms, n_dma=n_dma, n_host=n_host, frac=self.frac)
        if self.adapt and n_dma > 0 and n_host > 0 and dma_ms > 0 and host_ms > 0:
            r_dma, r_host = n_dma / dma_ms, n_host / host_ms          # rows per ms of either engine
            # the host side of a loader loop also enqueues the step and this feeder's own work: when the
            # DMA finished before the host pass did (the loop is host-bound), that overhead c -- the loop
            # period minus the host pass -- goes on the host's side of the balance
            #     n_dma / r_dma = c + (n - n_dma) / r_host
            if dma_first and self._loop_ms > 0:
                self._c = 0.5 * self._c + 0.5 * max(0.0, min(self._loop_ms - host_ms, 0.5 * host_ms))
            else:
                self._c *= 0.8                      # DMA-bound: the estimate would include waiting; let it decay
            c = self._c
            n = n_dma + n_host
            target = (c + n / r_host) / (1.0 / r_dma + 1.0 / r_host) / n
            self.last["overhead_ms"] = c
            self.frac = min(0.95, max(0.05, 0.5 * self.frac + 0.5 * target))

    def ship(self, x_host: torch.Tensor) -> SparseX:
        """Dense fp32 [N, K] in (pinned) host memory -> CSR on the device, on the current stream."""
        return self.finish(self.begin(x_host))

    def begin(self, x_host: torch.Tensor):
        """First half of ``ship``: hands the last rows to the copy engine and returns at once, so a
        loader loop can enqueue other device work before it spends the host pass in ``finish``."""
        if x_host.is_cuda or x_host.dtype != torch.float32 or not x_host.is_contiguous() or x_host.dim() != 2:
            raise L.BigcnError("HostFeeder.ship: expects a contiguous fp32 CPU matrix")
        n, k = x_host.shape
        if k != self.k or n > self.max_nodes:
            raise L.BigcnError(f"HostFeeder.ship: built for at most {self.max_nodes} rows of {self.k} features, got {n} x {k}")
        now = time.perf_counter()
        if self._t_begin is not None:
            loop = (now - self._t_begin) * 1e3
            self._loop_ms = loop if self._loop_ms == 0 else 0.5 * self._loop_ms + 0.5 * loop
        self._t_begin = now
        s = self.slots[self._turn]
        self._turn = (self._turn + 1) % len(self.slots)
        self._settle(s)
        n_dma = int(round(n * self.frac)) if x_host.is_pinned() else 0     # pageable memory cannot overlap: host only
        n_host = n - n_dma
        e0 = e1 = incl = None
        if n_dma > 0:
            if s["dense"] is None or s["dense"].shape[0] < n_dma:
                s["dense"] = torch.empty(min(self.max_nodes, int(n_dma * 1.25) + 1), k, dtype=torch.float32,
                                         device=self.device)
            d = s["dense"][:n_dma]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            d.copy_(x_host[n_host:], non_blocking=True)                    # the copy engine works from here on
            e1.record()
            check(lib().bigcn_dense_row_counts(_p(d), n_dma, k, _p(s["cnt"]), _stream()), "dense_row_counts")
            incl = torch.cumsum(s["cnt"][:n_dma], 0, dtype=torch.int32)
        return (s, x_host, n_host, n_dma, e0, e1, incl)

    def finish(self, ticket) -> SparseX:
        """Second half of ``ship`` (same stream as ``begin``): the host pass over the first rows, their
        CSR across PCIe, and the device compaction of the DMA'd rows behind it."""
        s, x_host, n_host, n_dma, e0, e1, incl = ticket
        n, k = x_host.shape
        t0 = time.perf_counter()
        hs = host_dense_to_csr(x_host[:n_host], n_threads=self.n_threads, out=s["host"], cap=self.cap)
        host_ms = (time.perf_counter() - t0) * 1e3
        dma_first = e1.query() if e1 is not None else False      # did the copy engine finish before the host pass?
        nnz_h = int(hs.col.numel())
        s["ptr"][:n_host + 1].copy_(hs.ptr, non_blocking=True)
        s["col"][:nnz_h].copy_(hs.col, non_blocking=True)
        s["val"][:nnz_h].copy_(hs.val, non_blocking=True)
        s["h2d_done"] = torch.cuda.Event()
        s["h2d_done"].record()
        t_cp = time.perf_counter()
        if n_dma > 0:
            check(lib().bigcn_dense_rows_to_csr(_p(s["dense"]), n_dma, k, _p(incl), nnz_h, _p(s["ptr"][n_host + 1:]),
                                                _p(s["col"]), _p(s["val"]), self.cap, _p(self.flags), _stream()),
                  "dense_rows_to_csr")
            s["ev"] = (e0, e1, n_dma, n_host, host_ms, dma_first)
            s["incl"] = incl
        self.last_cpu = dict(host_ms=host_ms, copies_ms=(t_cp - t0) * 1e3 - host_ms, fill_ms=(time.perf_counter() - t_cp) * 1e3)
        # the number of non-zeros of the DMA part is only known on the device: col / val are handed
        # over at capacity and ptr[N] bounds what is read
        return SparseX(s["ptr"][:n + 1], s["col"], s["val"], (n, k))
