"""Evaluation metrics against the loop-for-loop oracle restatement of the
evaluators (oracle/metrics.py) and against hand-computed known answers."""
import numpy as np
import pytest
import torch

from oracle import metrics as O
from rscotr_b200.mtl.data import metrics as M


# ------------------------------------------------------------------------------------------ cls
def test_accuracy_matches_oracle_and_known_answer():
    rng = np.random.default_rng(0)
    scores = rng.normal(size=(257, 45))
    gt = rng.integers(0, 45, size=257)
    assert M.accuracy(scores, gt, (1, 5)) == pytest.approx(O.accuracy_topk(scores, gt, (1, 5)))
    # known answer: 3 samples, labels ranked 1st / 2nd / last
    s = np.array([[.9, .05, .05], [.3, .6, .1], [.5, .4, .1]])
    assert M.accuracy(s, [0, 0, 2], (1, 2)) == pytest.approx([100 / 3, 200 / 3])
    res = M.evaluate_cls([torch.tensor(r) for r in s], [0, 0, 2], metric='accuracy', metric_options=dict(topk=(1, 2)))
    assert list(res) == ['accuracy_top-1', 'accuracy_top-2'] and res['accuracy_top-2'] == pytest.approx(200 / 3)
    # the threshold of mmcls: a correct but non-positive score does not count
    assert M.accuracy(np.array([[-1., -2.]]), [0], (1,), thr=0.) == [0.]
    with pytest.raises(ValueError):
        M.evaluate_cls(list(s), [0, 0, 2], metric='f1')


# ------------------------------------------------------------------------------------------ seg
@pytest.mark.parametrize('reduce_zero', [False, True])
def test_seg_areas_and_metrics_match_oracle(reduce_zero):
    g = torch.Generator().manual_seed(1)
    C, ignore = 5, 5
    pre_p, pre_o = [], []
    for _ in range(3):
        pred = torch.randint(0, C, (37, 41), generator=g)
        label = torch.randint(0, 7, (37, 41), generator=g)          # includes the ignore index and one out-of-range id
        label[0, :5] = 255
        a = M.intersect_and_union(pred, label, C, ignore, reduce_zero)
        b = O.seg_areas(pred.numpy(), label.numpy(), C, ignore, reduce_zero)
        for x, y in zip(a, b):
            assert np.array_equal(x.numpy(), y)
        pre_p.append(a)
        pre_o.append(b)
    ref = O.seg_metrics(pre_o, ('mFscore', 'mIoU'))
    names = ['imp', 'building', 'low_veg', 'tree', 'car']
    out = M.evaluate_seg(pre_p, names, metric=['mFscore', 'mIoU'])
    assert out['aAcc'] == pytest.approx(round(float(ref['aAcc']) * 100, 2) / 100)
    for k in ('IoU', 'Acc', 'Fscore', 'Precision', 'Recall'):
        assert out['m' + k] == pytest.approx(round(float(np.nanmean(ref[k])) * 100, 2) / 100), k
        for n, v in zip(names, ref[k]):
            assert out['%s.%s' % (k, n)] == pytest.approx(round(float(v) * 100, 2) / 100)


def test_seg_known_answer():
    pred = torch.tensor([[0, 0, 1, 1]])
    label = torch.tensor([[0, 1, 1, 2]])
    ai, au, ap, al = M.intersect_and_union(pred, label, 3, 255)
    assert ai.tolist() == [1, 1, 0] and au.tolist() == [2, 3, 1] and ap.tolist() == [2, 2, 0] and al.tolist() == [1, 2, 1]
    out = M.evaluate_seg([(ai, au, ap, al)], ['a', 'b', 'c'], metric='mIoU')
    assert out['aAcc'] == 0.5 and out['mIoU'] == pytest.approx(round((1 / 2 + 1 / 3 + 0) / 3 * 100, 2) / 100)
    with pytest.raises(KeyError):
        M.total_area_to_metrics(ai, au, ap, al, metrics=['mAP'])


# ------------------------------------------------------------------------------------------ det
def _random_det_case(seed, n_img=6, n_cat=3, crowd=True):
    rng = np.random.default_rng(seed)
    gts, dts = [], []
    for img in range(n_img):
        for _ in range(rng.integers(0, 6)):
            x, y = rng.uniform(0, 300, 2)
            w, h = rng.uniform(8, 150, 2)
            gts.append(dict(image_id=img, category_id=int(rng.integers(1, n_cat + 1)), bbox=[x, y, w, h], area=w * h,
                            iscrowd=int(crowd and rng.random() < 0.15)))
        for g in [g for g in gts if g['image_id'] == img]:                     # jittered copies of the gts ...
            if rng.random() < 0.8:
                b = np.array(g['bbox']) + rng.normal(0, 6, 4)
                b[2:] = np.abs(b[2:]) + 1
                dts.append(dict(image_id=img, category_id=g['category_id'] if rng.random() < 0.9 else 1, bbox=b.tolist(),
                                score=float(np.round(rng.random(), 2))))          # (rounded: ties exercise the stable sorts)
        for _ in range(rng.integers(0, 5)):                                        # ... plus clutter
            x, y = rng.uniform(0, 300, 2)
            w, h = rng.uniform(8, 150, 2)
            dts.append(dict(image_id=img, category_id=int(rng.integers(1, n_cat + 1)), bbox=[x, y, w, h],
                            score=float(np.round(rng.random(), 2))))
    return gts, dts, list(range(n_img)), list(range(1, n_cat + 1))


@pytest.mark.parametrize('seed', [0, 1, 2, 3])
@pytest.mark.parametrize('iou_thrs', [None, [0.5]])
def test_coco_bbox_eval_matches_oracle(seed, iou_thrs):
    gts, dts, img_ids, cat_ids = _random_det_case(seed)
    max_dets = (100, 300, 1000) if seed % 2 else (1, 3, 100)
    ev = M.coco_eval_bbox(gts, dts, cat_ids, img_ids, iou_thrs, max_dets)
    o = O.CocoEvalOracle(gts, dts, img_ids, cat_ids, iou_thrs, max_dets)
    o.evaluate()
    o.accumulate()
    assert np.array_equal(ev['precision'], o.precision)
    assert np.array_equal(ev['recall'], o.recall)
    assert M.coco_summarize(ev) == o.summarize()


def test_coco_bbox_known_answers():
    cat_ids, img_ids = [1, 2], [0, 1]
    gts = [dict(image_id=0, category_id=1, bbox=[10, 10, 50, 50], area=2500, iscrowd=0),
           dict(image_id=1, category_id=2, bbox=[20, 20, 100, 100], area=10000, iscrowd=0)]
    # perfect detections (mmdet result format: per image, per class (n,5) xyxy+score) -> AP = 1 everywhere defined
    perfect = [[np.array([[10, 10, 60, 60, .9]]), np.zeros((0, 5))], [np.zeros((0, 5)), np.array([[20, 20, 120, 120, .8]])]]
    out = M.evaluate_det(perfect, gts, img_ids, cat_ids, ['a', 'b'], iou_thrs=[0.5], classwise=True)
    assert out['bbox_mAP'] == 1.0 and out['bbox_mAP_50'] == 1.0
    assert out['bbox_mAP_75'] == -1.0                 # iou_thrs=[0.5]: no 0.75 threshold -> pycocotools prints -1
    assert out['bbox_mAP_s'] == -1.0 and out['bbox_mAP_m'] == 1.0 and out['bbox_mAP_l'] == 1.0
    assert out['bbox_AP.a'] == pytest.approx(1.0) and out['bbox_AP.b'] == pytest.approx(1.0)
    assert out['bbox_mAP_copypaste'].split()[0] == '1.000'
    # one true positive ranked below one false positive, one gt: precision 1/2 at every recall level -> AP 0.5
    res = [[np.array([[200, 200, 240, 240, .9], [10, 10, 60, 60, .5]]), np.zeros((0, 5))], [np.zeros((0, 5)), np.zeros((0, 5))]]
    out = M.evaluate_det(res, gts[:1], img_ids, cat_ids, iou_thrs=[0.5])
    assert out['bbox_mAP'] == pytest.approx(0.5, abs=1e-3)
    # no detections at all: AP 0 for categories that have gts
    none = [[np.zeros((0, 5)), np.zeros((0, 5))]] * 2
    assert M.evaluate_det(none, gts, img_ids, cat_ids, iou_thrs=[0.5])['bbox_mAP'] == 0.0
    with pytest.raises(KeyError):
        M.evaluate_det(none, gts, img_ids, cat_ids, metric='segm')


# ------------------------------------------------------------------------------------------ independent cross-checks
def test_cls_and_seg_metrics_against_sklearn():
    """scikit-learn as an independent implementation: top-k accuracy, per-class IoU / precision / recall / F1 from the
    confusion matrix (the evaluators themselves are absent from this image, see oracle/metrics.py)."""
    sk = pytest.importorskip('sklearn.metrics')
    rng = np.random.default_rng(3)
    scores = rng.normal(size=(400, 12))
    gt = rng.integers(0, 12, size=400)
    ours = M.accuracy(scores, gt, (1, 3, 5), thr=None)
    for k, a in zip((1, 3, 5), ours):
        assert a == pytest.approx(100 * sk.top_k_accuracy_score(gt, scores, k=k, labels=np.arange(12)))
    C = 6
    pred = rng.integers(0, C, size=(3, 50, 60))
    lab = rng.integers(0, C + 1, size=(3, 50, 60))                  # label C = ignore
    pre = [M.intersect_and_union(torch.from_numpy(p), torch.from_numpy(l), C, ignore_index=C) for p, l in zip(pred, lab)]
    out = M.evaluate_seg(pre, [str(c) for c in range(C)], metric=['mIoU', 'mFscore'])
    keep = lab.reshape(-1) != C
    y, p = lab.reshape(-1)[keep], pred.reshape(-1)[keep]
    iou = sk.jaccard_score(y, p, average=None, labels=np.arange(C))
    prec, rec, f1, _ = sk.precision_recall_fscore_support(y, p, labels=np.arange(C), zero_division=0)
    for c in range(C):
        assert out['IoU.%d' % c] == pytest.approx(round(iou[c] * 100, 2) / 100)
        assert out['Precision.%d' % c] == pytest.approx(round(prec[c] * 100, 2) / 100)
        assert out['Recall.%d' % c] == pytest.approx(round(rec[c] * 100, 2) / 100)
        assert out['Fscore.%d' % c] == pytest.approx(round(f1[c] * 100, 2) / 100)
    assert out['aAcc'] == pytest.approx(round(sk.accuracy_score(y, p) * 100, 2) / 100)
    assert out['mIoU'] == pytest.approx(round(iou.mean() * 100, 2) / 100)


def test_coco_ap_single_class_against_sklearn_style_pr_curve():
    """one category, IoU 0.5, no crowds: COCO AP = mean over the 101 recall thresholds of the monotone precision envelope of the
    score-ranked PR curve -- rebuilt here from scikit-learn's precision_recall_curve on the matched / unmatched flags."""
    sk = pytest.importorskip('sklearn.metrics')
    rng = np.random.default_rng(5)
    gts, dts = [], []
    for img in range(8):
        for k in range(3):
            x, y = rng.uniform(0, 200, 2)
            gts.append(dict(image_id=img, category_id=1, bbox=[x, y, 40, 40], area=1600, iscrowd=0))
            if rng.random() < 0.7:                                   # a good detection of this gt
                dts.append(dict(image_id=img, category_id=1, bbox=[x + 2, y - 1, 40, 40], score=float(rng.random())))
        for k in range(2):                                            # false positives far away
            dts.append(dict(image_id=img, category_id=1, bbox=[500 + 50 * k, 500, 30, 30], score=float(rng.random())))
    ev = M.coco_eval_bbox(gts, dts, [1], list(range(8)), [0.5], (100, 300, 1000))
    ap = M.coco_summarize(ev)[0]
    flags = np.array([1 if d['bbox'][0] < 400 else 0 for d in dts])                 # by construction: near boxes match, far ones do not
    scores = np.array([d['score'] for d in dts])
    npos = len(gts)
    order = np.argsort(-scores, kind='mergesort')
    tp = np.cumsum(flags[order])
    prec = tp / (np.arange(len(order)) + 1)
    rec = tp / npos
    env = np.maximum.accumulate(prec[::-1])[::-1]
    want = np.mean([env[np.searchsorted(rec, r, side='left')] if np.searchsorted(rec, r, side='left') < len(env) else 0.0
                    for r in np.linspace(0, 1, 101)])
    assert ap == pytest.approx(want, abs=1e-9)
    # scikit-learn's curve gives the same (precision, recall) points
    p_sk, r_sk, _ = sk.precision_recall_curve(flags, scores)
    assert np.allclose(sorted(set(np.round(r_sk * flags.sum() / npos, 9))), sorted(set(np.round(np.r_[0, rec], 9))))
