"""Golden vectors from the REFERENCE'S OWN PREPROCESSING (raw tree files -> per-tree arrays):

    python tests/golden/make_ref_ingest_golden.py        # needs /root/reference; writes ref_ingest/ + ref_ingest.npz

Two small raw datasets are written in the reference's text formats (tests/golden/ref_ingest/Twitter15/... and
.../Weibo/...; committed, they are the test inputs), then ``Process/getTwittergraph.py:main('Twitter15')`` and
``Process/getWeibograph.py:main()`` -- imported unmodified, run in a scratch working directory because they resolve
``data/...`` against os.getcwd() -- write their dense ``<eid>.npz`` files.  Those are far too large to commit (float64
[n, 5000]); ref_ingest.npz keeps, per tree, the non-zeros of ``x`` (np.nonzero: row-major, ascending columns), the
non-zeros of ``root``, ``edgeindex``, ``rootindex`` and ``y`` exactly as written.  The raw files exercise: a repeated
word index (last value wins), an explicit zero count, words above 5000 (dropped), Weibo's ``index - 1`` shift with
index 5000 and index 0 (column -1 = the last one), nodes listed out of order, a single-post tree and an unlabeled tree
(no file), label spellings in mixed case.
"""
import os
import shutil
import sys
import tempfile

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

TWITTER_TREE = [   # eid, parent, index, max_degree, maxL, vec
    ("501", "None", 1, 3, 9, "12:2 7:1 4999:3 12:5"),          # column 12 twice: the last value (5) stays
    ("501", "1", 2, 3, 9, "0:1 33:2"),
    ("501", "1", 3, 3, 9, "7:0 8:4"),                            # an explicit zero count
    ("501", "3", 4, 3, 9, "6000:2 45:1"),                        # 6000 > 5000: dropped
    ("501", "1", 5, 3, 9, "100:1"),
    ("502", "2", 1, 1, 4, "5:1"),                                # the root is node 2 and comes second in the file
    ("502", "None", 2, 1, 4, "9:2 3:1"),
    ("502", "1", 3, 1, 4, "4:1"),
    ("503", "None", 1, 0, 1, "1:1"),                             # a single post: no .npz
    ("504", "None", 1, 2, 3, "10:1"),
    ("504", "1", 3, 2, 3, "11:1"),                               # children listed out of order
    ("504", "1", 2, 2, 3, "12:1 13:1"),
    ("504", "2", 4, 2, 3, "14:2"),
    ("505", "None", 1, 1, 2, "20:1"),                            # no label line: no .npz
    ("505", "1", 2, 1, 2, "21:1"),
]
TWITTER_LABELS = [("Non-Rumor", "501"), ("false", "502"), ("true", "503"), ("UNVERIFIED", "504"), ("news", "599")]
WEIBO_TREE = [     # eid, parent, index, vec
    ("w1", "None", 1, "1:1 5000:2 17:3"),                        # 5000 -> column 4999
    ("w1", "1", 2, "0:4 2:1"),                                   # 0 -> column -1 = 4999
    ("w1", "2", 3, "3:1 3:2"),
    ("w2", "None", 1, "8:1"),
    ("w2", "1", 2, "9:1 5001:7"),
    ("w2", "1", 3, "10:2"),
    ("w2", "1", 4, "11:1"),
]
WEIBO_LABELS = [("w1", 1), ("w2", 0)]


def write_raw(root):
    t15 = os.path.join(root, "data", "Twitter15")
    wb = os.path.join(root, "data", "Weibo")
    os.makedirs(t15, exist_ok=True)
    os.makedirs(wb, exist_ok=True)
    with open(os.path.join(t15, "data.TD_RvNN.vol_5000.txt"), "w") as f:
        for r in TWITTER_TREE:
            f.write("\t".join(str(v) for v in r) + "\n")
    with open(os.path.join(t15, "Twitter15_label_All.txt"), "w") as f:
        for lab, eid in TWITTER_LABELS:
            f.write(f"{lab}\tx\t{eid}\n")
    with open(os.path.join(wb, "weibotree.txt"), "w") as f:
        for r in WEIBO_TREE:
            f.write("\t".join(str(v) for v in r) + "\n")
    with open(os.path.join(wb, "weibo_id_label.txt"), "w") as f:
        for eid, lab in WEIBO_LABELS:
            f.write(f"{eid} {lab}\n")


def main():
    raw = os.path.join(HERE, "ref_ingest")
    if os.path.isdir(raw):
        shutil.rmtree(raw)
    write_raw(raw)
    tmp = tempfile.mkdtemp()
    shutil.copytree(os.path.join(raw, "data"), os.path.join(tmp, "data"))
    os.makedirs(os.path.join(tmp, "data", "Twitter15graph"))   # the reference's loadEid drops the tree whose save creates it
    os.makedirs(os.path.join(tmp, "data", "Weibograph"))
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(tmp)                                              # the modules read os.getcwd() at import
    try:
        import Process.getTwittergraph as tw
        import Process.getWeibograph as wb
        tw.main("Twitter15")
        wb.main()
    finally:
        os.chdir(cwd)
    out = {}
    for name, sub in (("Twitter15", "Twitter15graph"), ("Weibo", "Weibograph")):
        d = os.path.join(tmp, "data", sub)
        files = sorted(os.listdir(d))
        out[f"{name}/files"] = np.array(files)
        for fn in files:
            z = np.load(os.path.join(d, fn), allow_pickle=True)
            eid = fn[:-4]
            x = z["x"]
            assert x.dtype == np.float64 and x.shape[1] == 5000
            r, c = np.nonzero(x)
            out[f"{name}/{eid}/n"] = np.array(x.shape[0])
            out[f"{name}/{eid}/x_row"], out[f"{name}/{eid}/x_col"], out[f"{name}/{eid}/x_val"] = r, c, x[r, c]
            rc = np.nonzero(z["root"][0])[0]
            out[f"{name}/{eid}/root_col"], out[f"{name}/{eid}/root_val"] = rc, z["root"][0][rc]
            out[f"{name}/{eid}/edgeindex"] = np.asarray(z["edgeindex"]).reshape(2, -1)
            out[f"{name}/{eid}/rootindex"] = np.asarray(z["rootindex"])
            out[f"{name}/{eid}/y"] = np.asarray(z["y"])
    np.savez(os.path.join(HERE, "ref_ingest.npz"), **out)
    shutil.rmtree(tmp)
    print("written:", {k: list(v) for k, v in out.items() if k.endswith("/files")})


if __name__ == "__main__":
    main()
