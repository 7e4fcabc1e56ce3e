"""CPU restatement of the reference's metric functions (TEST INFRASTRUCTURE, never imported by the
product): /root/reference/tools/evaluate.py:3-91 (evaluation4class) and :93-139 (evaluationclass)
-- per-class TP/FN/FP/TN with Python loops, then Acc / Prec / Recall / F1 rounded to four places.

PINNED: unlike the GCN oracle, this one is checked against outputs of the reference itself --
tests/golden/evaluate_golden.json is produced by tests/golden/make_evaluate_golden.py, which loads
the reference's own evaluate.py (pure Python, no third-party imports) in the build container."""


def evaluation_nclass(prediction, y, num_classes):
    tp = [0] * num_classes
    fn = [0] * num_classes
    fp = [0] * num_classes
    tn = [0] * num_classes
    for i in range(len(y)):
        act, pre = int(y[i]), int(prediction[i])
        for c in range(num_classes):            # evaluate.py:11-31: the four `if`s per class
            if act == c and pre == c: tp[c] += 1
            if act == c and pre != c: fn[c] += 1
            if act != c and pre == c: fp[c] += 1
            if act != c and pre != c: tn[c] += 1
    out = [round(float(sum(tp)) / float(len(y)), 4)]                      # :34
    for c in range(num_classes):
        acc = round(float(tp[c] + tn[c]) / float(tp[c] + tn[c] + fn[c] + fp[c]), 4)     # :35
        prec = 0 if (tp[c] + fp[c]) == 0 else round(float(tp[c]) / float(tp[c] + fp[c]), 4)   # :36-39
        rec = 0 if (tp[c] + fn[c]) == 0 else round(float(tp[c]) / float(tp[c] + fn[c]), 4)    # :40-43
        f1 = 0 if (prec + rec) == 0 else round(2 * prec * rec / (prec + rec), 4)              # :44-47
        out += [acc, prec, rec, f1]
    return tuple(out)


def evaluation4class(prediction, y):
    return evaluation_nclass(prediction, y, 4)


def evaluationclass(prediction, y):
    return evaluation_nclass(prediction, y, 2)
