"""Early stopping and checkpoints with the reference's names, call signatures and file formats
 -- host-side bookkeeping around the path, no arithmetic on tensors.

* ``EarlyStopping`` mirrors /root/reference/tools/earlystopping.py:4-70 (Twitter / PHEME, four F1
  scores, keeps the best epoch's checkpoint dict and writes it when patience runs out);
  ``EarlyStopping2class`` mirrors tools/earlystopping2class.py (Weibo: writes
  ``model.state_dict()`` to ``modelname + str + '.m'`` on every improvement).
* ``make_checkpoint`` builds the dict of BiGCN_Twitter.py:253-261 (keys fold / iter / epoch /
  model_state_dict / optimizer_state_dict / loss / res).  With a ``FusedTrainer`` the optimizer
  entry is a ``torch.optim.Adam`` state dict in the reference's parameter-group order
  (:146-153), so the reference's ``optimizer.load_state_dict`` accepts it and vice versa.
"""
from __future__ import annotations

import torch


class EarlyStopping:
    """tools/earlystopping.py:4-70.  ``score = -val_loss``; an epoch that does not improve on the
    best score bumps ``counter``; at ``patience`` the best epoch's checkpoint is written to
    ``best_{modelname}_{str}_f{fold}_i{iter}_e{epoch:05d}_l{loss:.5f}.pt`` (cwd, as the reference)."""

    def __init__(self, patience=7, verbose=False):
        self.patience = patience
        self.verbose = verbose
        self.counter = 0
        self.best_score = None
        self.early_stop = False
        self.accs = 0
        self.F1 = 0
        self.F2 = 0
        self.F3 = 0
        self.F4 = 0
        self.val_loss_min = float("inf")
        self.checkpoint = None
        self.saved_path = None

    def _keep(self, score, accs, F1, F2, F3, F4, checkpoint):
        self.best_score = score
        self.accs, self.F1, self.F2, self.F3, self.F4 = accs, F1, F2, F3, F4
        self.checkpoint = checkpoint

    def __call__(self, val_loss, accs, F1, F2, F3, F4, model, modelname, str, checkpoint):
        score = -val_loss
        if self.best_score is None:
            self._keep(score, accs, F1, F2, F3, F4, checkpoint)
        elif score < self.best_score:
            self.counter += 1
            if self.counter >= self.patience:
                self.early_stop = True
                print("BEST Accuracy: {:.4f}|NR F1: {:.4f}|FR F1: {:.4f}|TR F1: {:.4f}|UR F1: {:.4f}"
                      .format(self.accs, self.F1, self.F2, self.F3, self.F4))
                self.save_checkpoint(val_loss, model, modelname, str)
        else:
            self._keep(score, accs, F1, F2, F3, F4, checkpoint)
            self.counter = 0

    def save_checkpoint(self, val_loss, model, modelname, str):
        ck = self.checkpoint
        self.saved_path = "best_{}_{}_f{}_i{}_e{:05d}_l{:.5f}.pt".format(
            modelname, str, ck["fold"], ck["iter"], ck["epoch"], ck["loss"])
        torch.save(ck, self.saved_path)
        self.val_loss_min = val_loss


class EarlyStopping2class:
    """tools/earlystopping2class.py: the Weibo variant (two classes; state_dict saved on every improvement)."""

    def __init__(self, patience=7, verbose=False):
        self.patience = patience
        self.verbose = verbose
        self.counter = 0
        self.best_score = None
        self.early_stop = False
        self.accs = 0
        self.acc1 = self.acc2 = self.pre1 = self.pre2 = self.rec1 = self.rec2 = 0
        self.F1 = 0
        self.F2 = 0
        self.val_loss_min = float("inf")

    def _keep(self, score, accs, acc1, acc2, pre1, pre2, rec1, rec2, F1, F2):
        self.best_score = score
        self.accs, self.acc1, self.acc2 = accs, acc1, acc2
        self.pre1, self.pre2, self.rec1, self.rec2 = pre1, pre2, rec1, rec2
        self.F1, self.F2 = F1, F2

    def __call__(self, val_loss, accs, acc1, acc2, pre1, pre2, rec1, rec2, F1, F2, model, modelname, str):
        score = -val_loss
        if self.best_score is None:
            self._keep(score, accs, acc1, acc2, pre1, pre2, rec1, rec2, F1, F2)
            self.save_checkpoint(val_loss, model, modelname, str)
        elif score < self.best_score:
            self.counter += 1
            if self.counter >= self.patience:
                self.early_stop = True
                print("BEST LOSS:{:.4f}| Accuracy: {:.4f}|acc1: {:.4f}|acc2: {:.4f}|pre1: {:.4f}|pre2: {:.4f}"
                      "|rec1: {:.4f}|rec2: {:.4f}|F1: {:.4f}|F2: {:.4f}"
                      .format(-self.best_score, self.accs, self.acc1, self.acc2, self.pre1, self.pre2, self.rec1,
                              self.rec2, self.F1, self.F2))
        else:
            self._keep(score, accs, acc1, acc2, pre1, pre2, rec1, rec2, F1, F2)
            self.counter = 0
            self.save_checkpoint(val_loss, model, modelname, str)

    def save_checkpoint(self, val_loss, model, modelname, str):
        torch.save(model.state_dict(), modelname + str + ".m")
        self.val_loss_min = val_loss


def make_checkpoint(model, optimizer, fold, iter, epoch, loss, res):
    """The dict of BiGCN_Twitter.py:253-261.  ``optimizer``: a FusedTrainer or a torch optimizer.
    Tensors are cloned: the reference keeps the best epoch's dict while training goes on, and a
    FusedTrainer updates its flat buffers in place."""
    osd = optimizer.optimizer_state_dict() if hasattr(optimizer, "optimizer_state_dict") else optimizer.state_dict()
    return {"fold": fold, "iter": iter, "epoch": epoch,
            "model_state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()},
            "optimizer_state_dict": osd, "loss": loss, "res": res}
