"""Evaluation metrics of the three tasks, with the semantics of the third-party
evaluators the reference's `dataset.evaluate(results, **eval_kwargs[task])` call reaches
(reference call site: mtl/runner/hooks/evaluation.py:130-142; eval kwargs: configs/multi/
MTL_slvlcls_swin-t-p4-w7_1x1_resisc&dior&potsdam.py:222-238):

* cls  -- mmcls 0.23 BaseDataset.evaluate(metric='accuracy'): `accuracy_top-k` in percent.
* det  -- mmdet 2.25 CocoDataset.evaluate(metric='bbox', iou_thrs=[0.5], classwise=True) = pycocotools
          COCOeval (greedy score-ordered matching, 101-point interpolated precision), keys `bbox_mAP`,
          `bbox_mAP_50`, ... ; -1 where pycocotools prints -1.
* seg  -- mmseg 0.2x CustomDataset.evaluate(metric=['mFscore','mIoU'], pre_eval=True): per-image
          intersect / union / pred / label areas, summed, then aAcc / IoU / Acc / Fscore / Precision / Recall.

Tensor inputs may live on the GPU (the per-image seg areas are ONE bincount of label*C+pred on the
device the prediction is on); the final reductions run on the host in float64 like the evaluators."""
from collections import OrderedDict, defaultdict

import numpy as np
import torch


# --------------------------------------------------------------------------------------------- cls
def accuracy(scores, target, topk=(1, 5), thr=0.0):
    """mmcls.models.losses.accuracy: percent of samples whose label is among the k best scores (and
    whose matched score exceeds `thr`).  scores (N, C), target (N,).  Returns a list, one per k."""
    scores = torch.as_tensor(np.asarray(scores) if not torch.is_tensor(scores) else scores).float()
    target = torch.as_tensor(np.asarray(target) if not torch.is_tensor(target) else target).to(scores.device).long()
    n = scores.shape[0]
    maxk = min(max(topk), scores.shape[1])
    val, idx = scores.topk(maxk, dim=1)
    hit = idx.eq(target.view(-1, 1))
    if thr is not None:
        hit = hit & (val > thr)
    out = []
    for k in topk:
        out.append(float(hit[:, :min(k, maxk)].any(1).sum().item() * 100.0 / max(n, 1)))
    return out


def evaluate_cls(results, gt_labels, metric='accuracy', metric_options=None):
    """results: list of per-sample score vectors (what simple_test_cls returns)."""
    metric_options = metric_options or dict(topk=(1, 5))
    metrics = [metric] if isinstance(metric, str) else list(metric)
    out = OrderedDict()
    scores = np.vstack([np.asarray(r.detach().cpu() if torch.is_tensor(r) else r) for r in results])
    gt = np.asarray(gt_labels)
    assert len(gt) == scores.shape[0], 'dataset testing results should be of the same length as gt_labels.'
    for m in metrics:
        if m != 'accuracy':
            raise ValueError('metric %s is not supported.' % m)
        topk = metric_options.get('topk', (1, 5))
        topk = (topk,) if isinstance(topk, int) else tuple(topk)
        for k, a in zip(topk, accuracy(scores, gt, topk, metric_options.get('thrs', 0.0))):
            out['accuracy_top-%d' % k] = a
    return out


# --------------------------------------------------------------------------------------------- seg
def intersect_and_union(pred, label, num_classes, ignore_index, reduce_zero_label=False):
    """mmseg.core.evaluation.intersect_and_union for one image: four (num_classes,) area histograms
    (intersect, union, pred, label), float64.  One confusion bincount instead of 3 histc."""
    pred = torch.as_tensor(pred).long()
    label = torch.as_tensor(label).to(pred.device).long()
    if reduce_zero_label:
        label = torch.where(label == 0, torch.full_like(label, 255), label) - 1
        label = torch.where(label == 254, torch.full_like(label, 255), label)
    keep = label != ignore_index
    pred, label = pred[keep], label[keep]
    C = num_classes
    # one (C+1)x(C+1) confusion histogram; bucket C collects the out-of-range values histc(min=0, max=C-1) drops
    pred = torch.where((pred >= 0) & (pred < C), pred, torch.full_like(pred, C))
    label = torch.where((label >= 0) & (label < C), label, torch.full_like(label, C))
    conf = torch.bincount(label * (C + 1) + pred, minlength=(C + 1) * (C + 1)).view(C + 1, C + 1)
    area_label, area_pred = conf[:C].sum(1), conf[:, :C].sum(0)
    area_intersect = conf.diagonal()[:C]
    area_union = area_pred + area_label - area_intersect
    return tuple(t.double().cpu() for t in (area_intersect, area_union, area_pred, area_label))


def _f_score(precision, recall, beta=1):
    return (1 + beta ** 2) * (precision * recall) / ((beta ** 2 * precision) + recall)


def total_area_to_metrics(ti, tu, tp, tl, metrics=('mIoU',), nan_to_num=None, beta=1):
    metrics = [metrics] if isinstance(metrics, str) else list(metrics)
    for m in metrics:
        if m not in ('mIoU', 'mDice', 'mFscore'):
            raise KeyError('metrics %s is not supported' % m)
    ti, tu, tp, tl = (np.asarray(t, dtype=np.float64) for t in (ti, tu, tp, tl))
    with np.errstate(divide='ignore', invalid='ignore'):
        ret = OrderedDict(aAcc=ti.sum() / tl.sum())
        for m in metrics:
            if m == 'mIoU':
                ret['IoU'], ret['Acc'] = ti / tu, ti / tl
            elif m == 'mDice':
                ret['Dice'], ret['Acc'] = 2 * ti / (tp + tl), ti / tl
            else:
                precision, recall = ti / tp, ti / tl
                ret['Fscore'], ret['Precision'], ret['Recall'] = _f_score(precision, recall, beta), precision, recall
    if nan_to_num is not None:
        ret = OrderedDict((k, np.nan_to_num(v, nan=nan_to_num)) for k, v in ret.items())
    return ret


def evaluate_seg(pre_eval_results, class_names, metric='mIoU'):
    """CustomDataset.evaluate with pre_eval results (a list of per-image 4-tuples): summary values
    are round(nanmean*100, 2)/100, per-class values `<Metric>.<class>` likewise (mmseg/datasets/custom.py)."""
    metrics = [metric] if isinstance(metric, str) else list(metric)
    tot = [sum(np.asarray(r[i], dtype=np.float64) for r in pre_eval_results) for i in range(4)]
    ret = total_area_to_metrics(*tot, metrics=metrics)
    out = OrderedDict()
    for k, v in ret.items():
        s = np.round(np.nanmean(v) * 100, 2)
        out[k if k == 'aAcc' else 'm' + k] = float(s) / 100.0
    for k, v in ret.items():
        if k == 'aAcc':
            continue
        per = np.round(np.asarray(v) * 100, 2)
        for name, x in zip(class_names, per):
            out['%s.%s' % (k, name)] = float(x) / 100.0
    return out


# --------------------------------------------------------------------------------------------- det
_AREA_RNG = [[0 ** 2, 1e5 ** 2], [0 ** 2, 32 ** 2], [32 ** 2, 96 ** 2], [96 ** 2, 1e5 ** 2]]    # all / s / m / l
_REC_THRS = np.linspace(.0, 1.00, int(np.round((1.00 - .0) / .01)) + 1, endpoint=True)


def _iou_xywh(d, g, crowd):
    """pycocotools maskUtils.iou on xywh boxes: (D, G); crowd gts use the detection area as the union."""
    if len(d) == 0 or len(g) == 0:
        return np.zeros((len(d), len(g)))
    d, g = np.asarray(d, dtype=np.float64), np.asarray(g, dtype=np.float64)
    iw = np.minimum(d[:, None, 0] + d[:, None, 2], g[None, :, 0] + g[None, :, 2]) - np.maximum(d[:, None, 0], g[None, :, 0])
    ih = np.minimum(d[:, None, 1] + d[:, None, 3], g[None, :, 1] + g[None, :, 3]) - np.maximum(d[:, None, 1], g[None, :, 1])
    inter = np.clip(iw, 0, None) * np.clip(ih, 0, None)
    da, ga = (d[:, 2] * d[:, 3])[:, None], (g[:, 2] * g[:, 3])[None, :]
    union = np.where(np.asarray(crowd, dtype=bool)[None, :], da, da + ga - inter)
    with np.errstate(divide='ignore', invalid='ignore'):
        out = np.where((iw > 0) & (ih > 0), inter / union, 0.0)
    return out


def _match_image(ious, g_ignore, g_crowd, iou_thrs):
    """COCOeval.evaluateImg inner loop for one (image, category, area range): dets in score order
    greedily take the best still-free gt with iou >= thr; non-ignored gts are preferred (gts are sorted
    non-ignored first and the scan stops at the first ignored gt once a real match exists); crowd gts
    can be matched repeatedly.  Returns dt_match (T, D) gt index or -1, dt_ignore (T, D)."""
    T, (D, G) = len(iou_thrs), ious.shape
    dtm = -np.ones((T, D), dtype=np.int64)
    dti = np.zeros((T, D), dtype=bool)
    for ti, t in enumerate(iou_thrs):
        taken = np.zeros(G, dtype=bool)
        for di in range(D):
            best, m = min(t, 1 - 1e-10), -1
            for gi in range(G):
                if taken[gi] and not g_crowd[gi]:
                    continue
                if m > -1 and not g_ignore[m] and g_ignore[gi]:
                    break
                if ious[di, gi] < best:
                    continue
                best, m = ious[di, gi], gi
            if m == -1:
                continue
            dti[ti, di] = g_ignore[m]
            dtm[ti, di] = m
            taken[m] = True
    return dtm, dti


def coco_eval_bbox(gts, dts, cat_ids, img_ids, iou_thrs=None, max_dets=(100, 300, 1000)):
    """pycocotools COCOeval(iouType='bbox') evaluate + accumulate.
    gts: list of dict(image_id, category_id, bbox=[x,y,w,h], area, iscrowd);
    dts: list of dict(image_id, category_id, bbox=[x,y,w,h], score).
    Returns precision (T, R, K, A, M) and recall (T, K, A, M), -1 where undefined."""
    iou_thrs = np.asarray(iou_thrs if iou_thrs is not None else
                          np.linspace(.5, 0.95, int(np.round((0.95 - .5) / .05)) + 1, endpoint=True), dtype=np.float64)
    T, R, K, A, M = len(iou_thrs), len(_REC_THRS), len(cat_ids), len(_AREA_RNG), len(max_dets)
    img_ids = sorted(set(img_ids))                     # (COCOeval.evaluate: np.unique -> tie order of equal scores)
    g_by, d_by = defaultdict(list), defaultdict(list)
    for g in gts:
        g_by[g['image_id'], g['category_id']].append(g)
    for d in dts:
        d_by[d['image_id'], d['category_id']].append(d)
    precision, recall = -np.ones((T, R, K, A, M)), -np.ones((T, K, A, M))
    maxdet = max_dets[-1]
    for k, cat in enumerate(cat_ids):
        per_img = []                                   # (scores, [(dtm, dti, n_valid_gt)] per area)
        for img in img_ids:
            g, d = g_by.get((img, cat), []), d_by.get((img, cat), [])
            if not g and not d:
                continue
            order = np.argsort([-x['score'] for x in d], kind='mergesort')[:maxdet]
            d = [d[i] for i in order]
            dbox = np.array([x['bbox'] for x in d], dtype=np.float64).reshape(-1, 4)
            darea = dbox[:, 2] * dbox[:, 3]
            crowd0 = np.array([bool(x.get('iscrowd', 0)) for x in g], dtype=bool)
            ign0 = crowd0.copy()                       # (COCOeval._prepare overwrites 'ignore' with iscrowd)
            garea = np.array([x['area'] for x in g], dtype=np.float64)
            gbox = np.array([x['bbox'] for x in g], dtype=np.float64).reshape(-1, 4)
            ious0 = _iou_xywh(dbox, gbox, crowd0)
            areas = []
            for lo, hi in _AREA_RNG:
                gi = ign0 | (garea < lo) | (garea > hi)
                gorder = np.argsort(gi, kind='mergesort')
                dtm, dti = _match_image(ious0[:, gorder], gi[gorder], crowd0[gorder], iou_thrs)
                out_rng = (darea < lo) | (darea > hi)
                dti = dti | ((dtm == -1) & out_rng[None, :])
                areas.append((dtm, dti, int((~gi).sum())))
            per_img.append((np.array([x['score'] for x in d], dtype=np.float64), areas))
        if not per_img:
            continue
        for a in range(A):
            for m, md in enumerate(max_dets):
                scores = np.concatenate([s[:md] for s, _ in per_img])
                inds = np.argsort(-scores, kind='mergesort')
                dtm = np.concatenate([ar[a][0][:, :md] for _, ar in per_img], axis=1)[:, inds]
                dti = np.concatenate([ar[a][1][:, :md] for _, ar in per_img], axis=1)[:, inds]
                npig = sum(ar[a][2] for _, ar in per_img)
                if npig == 0:
                    continue
                tps = np.cumsum((dtm != -1) & ~dti, axis=1).astype(np.float64)
                fps = np.cumsum((dtm == -1) & ~dti, axis=1).astype(np.float64)
                for t in range(T):
                    tp, fp = tps[t], fps[t]
                    nd = len(tp)
                    rc = tp / npig
                    pr = tp / (fp + tp + np.spacing(1))
                    recall[t, k, a, m] = rc[-1] if nd else 0
                    pr = np.maximum.accumulate(pr[::-1])[::-1] if nd else pr      # monotone envelope from the right
                    pos = np.searchsorted(rc, _REC_THRS, side='left')
                    q = np.zeros(R)
                    ok = pos < nd
                    q[ok] = pr[pos[ok]]
                    precision[t, :, k, a, m] = q
    return dict(precision=precision, recall=recall, iou_thrs=iou_thrs, max_dets=tuple(max_dets))


def coco_summarize(ev):
    """COCOeval.summarize with mmdet's maxDets=(100,300,1000): the 12 `stats` (AP uses maxDets[-1])."""
    P, Rc, thrs, M = ev['precision'], ev['recall'], ev['iou_thrs'], len(ev['max_dets'])

    def summ(ap, iou=None, area=0, m=M - 1):
        s = P[:, :, :, area, m] if ap else Rc[:, :, area, m]
        if iou is not None:
            s = s[np.where(np.isclose(thrs, iou))[0]]          # (empty selection -> -1, like pycocotools)
        s = s[s > -1]
        return float(np.mean(s)) if s.size else -1.0

    return [summ(1), summ(1, .5), summ(1, .75), summ(1, None, 1), summ(1, None, 2), summ(1, None, 3),
            summ(0, None, 0, 0), summ(0, None, 0, min(1, M - 1)), summ(0, None, 0, M - 1),
            summ(0, None, 1), summ(0, None, 2), summ(0, None, 3)]


def results_to_coco(results, img_ids, cat_ids):
    """CocoDataset._det2json: per image a list (one (n,5) array [x1,y1,x2,y2,score] per class)."""
    out = []
    for img_id, per_cls in zip(img_ids, results):
        for label, arr in enumerate(per_cls):
            arr = np.asarray(arr.detach().cpu() if torch.is_tensor(arr) else arr, dtype=np.float64).reshape(-1, 5)
            for x1, y1, x2, y2, s in arr:
                out.append(dict(image_id=img_id, category_id=cat_ids[label], bbox=[x1, y1, x2 - x1, y2 - y1], score=float(s)))
    return out


def evaluate_det(results, gts, img_ids, cat_ids, class_names=None, metric='bbox', iou_thrs=None, classwise=False,
                 proposal_nums=(100, 300, 1000), metric_items=None):
    """CocoDataset.evaluate(metric='bbox'): `bbox_mAP`, `bbox_mAP_50`, `bbox_mAP_75`, `bbox_mAP_s/m/l`
    (3 decimals) + `bbox_mAP_copypaste`; classwise adds `bbox_AP.<class>` entries (the reference's
    mmdet only prints that table; exposing it costs nothing)."""
    metrics = [metric] if isinstance(metric, str) else list(metric)
    for m in metrics:
        if m != 'bbox':
            raise KeyError('metric %s is not supported' % m)
    ev = coco_eval_bbox(gts, results_to_coco(results, img_ids, cat_ids), cat_ids, img_ids, iou_thrs, tuple(proposal_nums))
    stats = coco_summarize(ev)
    names = ['mAP', 'mAP_50', 'mAP_75', 'mAP_s', 'mAP_m', 'mAP_l']
    out = OrderedDict()
    for i, n in enumerate(names):
        if metric_items is None or n in metric_items:
            out['bbox_' + n] = float('%.3f' % stats[i])
    out['bbox_mAP_copypaste'] = ' '.join('%.3f' % s for s in stats[:6])
    if classwise:
        P = ev['precision']
        for k, cat in enumerate(cat_ids):
            p = P[:, :, k, 0, -1]
            p = p[p > -1]
            name = class_names[k] if class_names is not None else str(cat)
            out['bbox_AP.%s' % name] = float(np.mean(p)) if p.size else float('nan')
    return out
