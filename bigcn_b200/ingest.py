"""The reference's RAW dataset files straight into a ``DeviceForest`` (the step before the hot path).

The reference preprocesses ``data/<name>/data.TD_RvNN.vol_5000.txt`` (Twitter15/16) or ``data/Weibo/weibotree.txt`` into
one dense ``<eid>.npz`` per tree -- ``x`` [n, 5000] float64, ~99.7 % zeros -- with Process/getTwittergraph.py /
Process/getWeibograph.py, and ``BiGraphDataset`` (Process/dataset.py:64-99) re-reads those files every epoch.  Here the
``index:count`` pairs are kept as they are: each tree becomes a CSR block, nothing is densified, and the whole dataset
is packed once into HBM (``DeviceForest``).  The parsing rules are the reference's, line by line:

  tree file   Twitter: ``eid \\t parent \\t index \\t max_degree \\t maxL \\t vec`` (getTwittergraph.py:78-85)
              Weibo:   ``eid \\t parent \\t index \\t vec``                       (getWeibograph.py:75-81)
  vec         ``i:f i:f ...``; pairs with ``i <= 5000`` are kept (getTwittergraph.py:16-24); Twitter uses column ``i``,
              Weibo column ``i - 1`` (getWeibograph.py:22); a repeated column keeps the LAST value and a column
              index of -1 addresses the last column (numpy fancy assignment, getTwittergraph.py:67-72); Twitter's
              ``i == 5000`` is out of range for the reference too (IndexError there, ValueError here)
  nodes       ids 1..n, node i is row i - 1; the root is the node whose parent is the string ``None`` (:43-46)
  edges       [parent; child], ordered by (parent, child) -- the order of the reference's double loop (:55-62)
  labels      Twitter: ``label \\t ? \\t eid`` with news / non-rumor -> 0, false -> 1, true -> 2, unverified -> 3
              (:88-110); Weibo: ``eid label`` (getWeibograph.py:87-93)
  kept trees  at least 2 nodes (``loadEid`` writes nothing for smaller ones, :113-117) and present in both files

tests/golden/ref_ingest.npz holds what the reference's own ``main()`` wrote for two small raw files
(tests/golden/make_ref_ingest_golden.py); tests/test_ingest.py holds this module to it.
"""
from __future__ import annotations

import numpy as np

TWITTER_LABELS = {"news": 0, "non-rumor": 0, "false": 1, "true": 2, "unverified": 3}


def _is_weibo(dataset: str) -> bool:
    return dataset.lower().startswith("weibo")


def parse_tree_file(path: str, dataset: str) -> dict:
    """``treeDic[eid][index] = {'parent': str, 'vec': str}`` (a later line for the same (eid, index) replaces an
    earlier one, as the reference's dict assignment does)."""
    vec_field = 3 if _is_weibo(dataset) else 5
    trees: dict = {}
    with open(path) as f:
        for line in f:
            line = line.rstrip()
            if not line:
                continue
            parts = line.split("\t")
            trees.setdefault(parts[0], {})[int(parts[2])] = {"parent": parts[1], "vec": parts[vec_field]}
    return trees


def parse_labels(path: str, dataset: str):
    """``(events in file order, labelDic)``."""
    events, labels = [], {}
    with open(path) as f:
        for line in f:
            line = line.rstrip()
            if not line:
                continue
            if _is_weibo(dataset):
                eid, lab = line.split(" ")[0], line.split(" ")[1]
                labels[eid] = int(lab)
            else:
                lab, eid = line.split("\t")[0].lower(), line.split("\t")[2]
                if lab in TWITTER_LABELS:
                    labels[eid] = TWITTER_LABELS[lab]
            events.append(eid)
    return events, labels


def tree_arrays(tree: dict, dataset: str, in_feats: int = 5000):
    """One tree of ``parse_tree_file`` -> ``(n, edges int64 [2, e], rootindex, ptr int64 [n + 1], col int32, val float32)``;
    the CSR rows hold ascending columns, explicit zeros dropped (what ``np.nonzero`` of the reference's dense ``x``
    would give)."""
    n = len(tree)
    if sorted(tree) != list(range(1, n + 1)):
        raise ValueError("tree_arrays: node ids must be 1..n (the reference indexes index2node[i + 1])")
    shift = 1 if _is_weibo(dataset) else 0
    children = [[] for _ in range(n)]
    root = None
    ptr, cols, vals = [0], [], []
    for idx in range(1, n + 1):
        node = tree[idx]
        row = {}
        for pair in node["vec"].split(" "):
            i, f = int(pair.split(":")[0]), float(pair.split(":")[1])
            if i > 5000:
                continue
            c = i - shift
            if c < 0:
                c += in_feats                                  # numpy: x[r, -1] is the last column
            if not 0 <= c < in_feats:
                raise ValueError(f"tree_arrays: word index {i} addresses column {c} of {in_feats}")
            row.pop(c, None)
            row[c] = f                                         # the last pair for a column wins
        keep = sorted(c for c, f in row.items() if f != 0.0)
        cols.extend(keep)
        vals.extend(row[c] for c in keep)
        ptr.append(len(cols))
    for idx in tree:                                           # dict order = file order: a parent's children in that order
        p = tree[idx]["parent"]
        if p == "None":
            root = idx - 1
        else:
            if not 1 <= int(p) <= n:
                raise ValueError(f"tree_arrays: parent {p} of node {idx} is not a node of the tree")
            children[int(p) - 1].append(idx - 1)
    if root is None:
        raise ValueError("tree_arrays: no node with parent 'None'")
    src, dst = [], []
    for i in range(n):                                         # the reference's (i, j) double loop: sorted, no duplicates
        for j in sorted(set(children[i])):
            src.append(i)
            dst.append(j)
    edges = np.array([src, dst], np.int64).reshape(2, -1)
    return n, edges, root, np.array(ptr, np.int64), np.array(cols, np.int32), np.array(vals, np.float32)


def forest_from_raw(tree_path: str, label_path: str, dataset: str, device, fold_x=None, in_feats: int = 5000):
    """``DeviceForest`` over the trees of ``fold_x`` (default: every event of the label file, in file order) that the
    reference's preprocessing would have written an ``.npz`` for.  ``forest.ids[t]`` is the event id of tree ``t``."""
    from .loader import DeviceForest
    trees = parse_tree_file(tree_path, dataset)
    events, labels = parse_labels(label_path, dataset)
    ids = [e for e in dict.fromkeys(events if fold_x is None else fold_x)      # first occurrence, order kept
           if e in trees and e in labels and len(trees[e]) >= 2]
    node_ptr, edge_ptr, x_ptr = [0], [0], [0]
    es, ed, xc, xv, roots, ys = [], [], [], [], [], []
    for e in ids:
        n, edges, root, ptr, col, val = tree_arrays(trees[e], dataset, in_feats)
        es.append(edges[0].astype(np.int32)); ed.append(edges[1].astype(np.int32))
        xc.append(col); xv.append(val)
        x_ptr.extend((x_ptr[-1] + ptr[1:]).tolist())
        node_ptr.append(node_ptr[-1] + n)
        edge_ptr.append(edge_ptr[-1] + edges.shape[1])
        roots.append(root); ys.append(labels[e])
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)  # noqa: E731
    forest = DeviceForest(node_ptr, edge_ptr, cat(es, np.int32), cat(ed, np.int32), x_ptr, cat(xc, np.int32),
                          cat(xv, np.float32), roots, ys, in_feats, device)
    forest.ids = ids
    return forest
