"""Evaluation bookkeeping without per-batch host syncs.

``EvalCounts`` accumulates, on the device, the confusion counts the reference's
``evaluation4class`` / ``evaluationclass`` (/root/reference/tools/evaluate.py:3-91, :93-139) build
with Python loops per batch, the number of correct predictions (BiGCN_Twitter.py:190-191) and the
summed nll loss.  Every ``update`` fills its own slot, so one device->host read at the end of an
epoch yields both the pooled tuple (``result``) and what the reference's validation loop logs --
the mean over batches of the per-batch tuples (``epoch_means``, BiGCN_Twitter.py:203-246)."""
from __future__ import annotations

import torch

from . import _lib as L
from ._lib import check, lib
from .ops import _p, _stream, _i64, _f32


def _tuple_from_counts(cnt, n):
    """(Acc_all, Acc1, Prec1, Recll1, F1, Acc2, ...) from per-class [TP, FN, FP, TN]: evaluate.py:33-91."""
    out = [round(float(sum(c[0] for c in cnt)) / float(n), 4)]
    for tp, fn, fp, tn in cnt:
        acc = round(float(tp + tn) / float(tp + tn + fn + fp), 4)
        prec = 0 if tp + fp == 0 else round(float(tp) / float(tp + fp), 4)
        rec = 0 if tp + fn == 0 else round(float(tp) / float(tp + fn), 4)
        f1 = 0 if prec + rec == 0 else round(2 * prec * rec / (prec + rec), 4)
        out += [acc, prec, rec, f1]
    return tuple(out)


class EvalCounts:
    def __init__(self, num_classes: int, device, slots: int = 64):
        self.c = int(num_classes)
        self.device = device
        self.counts = torch.zeros(max(int(slots), 1), self.c, 4, dtype=torch.int64, device=device)
        self.totals = torch.zeros(max(int(slots), 1), 3, dtype=torch.int64, device=device)
        self.n_batches = 0

    def reset(self):
        self.counts.zero_()
        self.totals.zero_()
        self.n_batches = 0

    def update(self, logp, y):
        """Add a batch: log-probs [B, C] (the model's output) and labels [B]; no synchronisation."""
        L.require_device()
        logp, y = _f32(logp.detach()), _i64(y)
        b, c = logp.shape
        if c != self.c:
            raise L.BigcnError(f"EvalCounts: expected {self.c} classes, got {c}")
        if b == 0:
            return
        s = self.n_batches
        if s == self.counts.shape[0]:              # grow on the device, nothing is read back
            self.counts = torch.cat([self.counts, torch.zeros_like(self.counts)])
            self.totals = torch.cat([self.totals, torch.zeros_like(self.totals)])
        check(lib().bigcn_eval_counts(_p(logp), _p(y), b, c, _p(self.counts[s]), _p(self.totals[s]), _stream()),
              "eval_counts")
        self.n_batches = s + 1

    def _read(self):
        s = self.n_batches
        return self.counts[:s].cpu().tolist(), self.totals[:s].cpu().tolist()

    def result(self):
        """The reference tuple over everything added so far (all batches pooled).  One device->host read."""
        cnt, tot = self._read()
        n = sum(t[0] for t in tot)
        if n == 0:
            return tuple([0.0] * (1 + 4 * self.c))
        pooled = [[sum(b[k][j] for b in cnt) for j in range(4)] for k in range(self.c)]
        return _tuple_from_counts(pooled, n)

    def per_batch(self):
        """[(tuple, acc, mean nll)] per update, as the reference's validation loop computes them per batch."""
        cnt, tot = self._read()
        return [(_tuple_from_counts(c, t[0]), t[1] / t[0], t[2] * 1e-6 / t[0]) for c, t in zip(cnt, tot)]

    def epoch_means(self):
        """(val_loss, val_acc, tuple of means): np.mean over batches of the per-batch values, which is
        what the reference logs, checkpoints and hands to EarlyStopping (BiGCN_Twitter.py:226-258)."""
        pb = self.per_batch()
        if not pb:
            return 0.0, 0.0, tuple([0.0] * (1 + 4 * self.c))
        k = len(pb)
        means = tuple(sum(p[0][i] for p in pb) / k for i in range(1 + 4 * self.c))
        return sum(p[2] for p in pb) / k, sum(p[1] for p in pb) / k, means

    def accuracy_and_loss(self):
        """(correct / trees, mean nll) over everything added so far (BiGCN_Twitter.py:188-191,219-222)."""
        _, tot = self._read()
        n = sum(t[0] for t in tot)
        return (sum(t[1] for t in tot) / n, sum(t[2] for t in tot) * 1e-6 / n) if n else (0.0, 0.0)
