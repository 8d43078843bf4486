"""CPU oracle for the RSCoTr co-training hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain-PyTorch (fp32, eager, CPU) restatement of the
algorithm the reference executes on its hot path.  It is
imported only by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``; nothing under
``rscotr_b200/`` may import it.

PARITY PINNED ONLY IN PART.  Pinned against outputs of the reference's OWN in-tree code run in this container
(``tools/make_golden.py`` -> ``tests/golden/reference_*``, checked by ``tests/test_golden_reference.py`` on CPU and,
with the same fixtures, on the CUDA path): CdnQueryGenerator (a15), the DINO decoder / sine embedding (a14),
DinoTransformer.forward and DINOHead.forward (a13), DINOHead.loss / DETRHead.loss_single / dn targets (a16, in-tree
flow), MlvlSegPixelDecoder.forward (a17), Mask2FormerHead.forward / forward_head (a18), MTL.train_step /
_parse_losses (a21), the iteration strategies and MultiDataLoader (a22), MlvlClsPixelDecoder.forward and
MlvlClsHead.pre_logits_1..8 (8f rank 4), the evaluation path DeformableDETRHead.get_bboxes / DETRHead._get_bboxes_single and
MTL.simple_test_{det,seg} / inference_seg / whole_inference_seg / forward_test (8f rank 4).  ``oracle/metrics.py`` (evaluators) and ``oracle/uper.py`` (UPerNet, a20) restate
third-party code only and say PARITY UNPINNED in their own headers.
Everything else is PARITY UNPINNED: the arithmetic of the path lives in un-vendored third-party
packages (mmcv-full 1.6.1, mmdet 2.25.1, mmsegmentation 0.28.0, mmcls) that are
not installable in this image, and the reference ships no tests, golden vectors
or fixtures.  The restatement follows the reference's
in-tree call sites (cited per function as ``file:line`` relative to
``/root/reference``) and the published algorithm of the pinned third-party
versions; it is cross-checked against the independent implementations that
are importable here (``torchvision.models.swin_transformer``, HF
``transformers`` deformable-DETR, ``torch.nn.functional.grid_sample``,
``scipy.optimize.linear_sum_assignment``) by ``tests/test_oracle_*.py``.

Everything is functional and keyed on the reference's state-dict key layout
(SURVEY.md section 8b), so a reference checkpoint can be fed to it unchanged.
"""
