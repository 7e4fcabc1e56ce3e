"""Golden traces of the reference's EarlyStopping classes, produced by the reference itself (run in
the build container):

    python tests/golden/make_earlystopping_golden.py

/root/reference/tools/earlystopping.py and earlystopping2class.py import only numpy and torch but use
``np.Inf``, which numpy 2 removed; the alias is restored before loading them (no other change).
The trace records, after every call, counter / best_score / early_stop / the kept scores and the
files written (names only) for a seeded validation-loss sequence."""
import importlib.util
import json
import os
import random
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def load(path, name):
    if not hasattr(np, "Inf"):
        np.Inf = np.inf
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    es4 = load("/root/reference/tools/earlystopping.py", "ref_es4").EarlyStopping
    es2 = load("/root/reference/tools/earlystopping2class.py", "ref_es2").EarlyStopping
    rng = random.Random(0)
    out = {"four": [], "two": []}
    model = torch.nn.Linear(2, 2)
    cwd = os.getcwd()
    for patience in (1, 3, 10):
        losses = [round(1.0 / (1 + 0.3 * i) + rng.uniform(-0.15, 0.15), 4) for i in range(14)]
        losses[5] = losses[4]                                   # a tie counts as an improvement (score == best)
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            try:
                e = es4(patience=patience, verbose=True)
                trace = []
                for ep, vl in enumerate(losses):
                    sc = [round(rng.random(), 4) for _ in range(5)]
                    ck = {"fold": 2, "iter": 0, "epoch": ep, "loss": round(vl * 1.1, 6)}
                    e(vl, *sc, model, "BiGCN", "Twitter16", checkpoint=ck)
                    trace.append(dict(val_loss=vl, scores=sc, counter=e.counter, best_score=e.best_score,
                                      early_stop=e.early_stop, kept=[e.accs, e.F1, e.F2, e.F3, e.F4],
                                      kept_epoch=e.checkpoint["epoch"], files=sorted(os.listdir(tmp))))
                    if e.early_stop:
                        break
                out["four"].append(dict(patience=patience, trace=trace))
                for f in os.listdir(tmp):
                    os.remove(f)
                e = es2(patience=patience, verbose=True)
                trace = []
                for ep, vl in enumerate(losses):
                    sc = [round(rng.random(), 4) for _ in range(9)]
                    e(vl, *sc, model, "BiGCN", "Weibo")
                    trace.append(dict(val_loss=vl, scores=sc, counter=e.counter, best_score=e.best_score,
                                      early_stop=e.early_stop,
                                      kept=[e.accs, e.acc1, e.acc2, e.pre1, e.pre2, e.rec1, e.rec2, e.F1, e.F2],
                                      val_loss_min=e.val_loss_min, files=sorted(os.listdir(tmp))))
                    if e.early_stop:
                        break
                out["two"].append(dict(patience=patience, trace=trace))
            finally:
                os.chdir(cwd)
    with open(os.path.join(HERE, "earlystopping_golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote", sum(len(c["trace"]) for k in out for c in out[k]), "calls")


if __name__ == "__main__":
    main()
