"""CPU oracle for the BiGCN hot path -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the arithmetic of the reference's hot path lives in two
third-party wheels that are NOT vendored under /root/reference and are NOT
installable in this image (no network): ``torch_geometric`` (GCNConv /
gcn_norm / add_remaining_self_loops) and ``torch_scatter`` (scatter_mean).
The reference ships no test, golden vector or logged output for this path
(SURVEY.md section 4 / 8c).  This package therefore *restates* the published
algorithm of those libraries (PyG 2.x semantics, degree-by-target; the
readme-pinned 1.3.2 degree-by-source convention is kept behind a switch) and
anchors it on the reference's own call sites:

  model/Twitter/BiGCN_Twitter.py:19-131   (TDrumorGCN / BUrumorGCN / BiGCN)
  model/Weibo/BiGCN_Weibo.py:16-89        (same maths, class Net, 2 classes)
  Process/dataset.py:64-99                (input contract, DropEdge)

Probed again on the GPU box this round (profiles/r02_pyg_probe.log, one ``gpurun`` call: ``import
torch_geometric`` / ``torch_scatter`` / ``torch_sparse`` -> ModuleNotFoundError, no ``baseline/_ref``, nothing
on disk): the box runs this same image, so no fixture can be generated from the real GCNConv / scatter_mean
and the parity of the GCN path stays UNPINNED at the library boundary.  What the restatement is held to
instead: the hand-checkable 5-node vector below, torch's own CPU ops wherever the reference calls torch,
and an independent dense-matrix derivation of the published formula D^-1/2 (A + I) D^-1/2 X W in fp64
(tests/test_oracle.py::test_gcnconv_matches_dense_normalised_adjacency).

Pinned this round -- everything the reference's OWN files decide: tests/golden/ref_wiring.npz holds outputs of
model/Twitter/BiGCN_Twitter.py:19-131 (imported unmodified) and of the classes of model/Weibo/BiGCN_Weibo.py:16-89,
run in the build container over dense stand-ins of the two absent wheels (an explicit [N, N] normalised adjacency
for GCNConv, index_add_ / count for scatter_mean -- independent of this package; tests/golden/
make_ref_wiring_golden.py).  tests/test_oracle.py::test_oracle_matches_reference_model_code holds the restated
modules to them in eval AND training mode (same torch generator state -> same F.dropout masks), log-probs <= 2e-6,
loss, all ten gradients (so the autograd effect of ``copy.copy`` at :44 is the reference's, not an assumption);
tests/test_gpu_parity.py::test_reference_model_code_fixture holds the CUDA path to them directly.  What remains
unpinned is only the inside of the two library calls (gcn_norm's conventions, summation order).

Pinned exception: ``evaluate_oracle`` (tools/evaluate.py) is checked against
outputs of the reference's own file (tests/golden/make_evaluate_golden.py).
For the GCN path the only pin available is the hand-checkable 5-node known-answer vector of
SURVEY.md section 8c (tests/golden/kat_tree5.json) plus torch's own CPU ops
(Linear, relu, log_softmax, nll_loss, pow(-0.5)) which are used directly.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  The
product (``bigcn_b200``) never does; it fails loudly if its CUDA library is
missing.
"""
