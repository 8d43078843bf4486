"""CPU oracle of the evaluation metrics.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED: the evaluators are third-party and absent from /root/reference and from this image
(pycocotools 2.0.x `cocoeval.py`, mmseg 0.28 `core/evaluation/metrics.py`, mmcls 0.23
`models/losses/accuracy.py`); the reference holds no golden metric values.  Each function restates
the published algorithm, loop for loop, and is anchored on the reference's call site
(mtl/runner/hooks/evaluation.py:130-142) and eval kwargs (configs/multi/MTL_slvlcls_swin-t-p4-w7_1x1_
resisc&dior&potsdam.py:222-238).  Known-answer cases and scikit-learn (top-k accuracy, confusion-matrix IoU / precision / recall / F1, a rebuilt PR
curve for single-class AP) cross-check it in tests/test_metrics.py."""
import numpy as np


# ---- mmcls accuracy_numpy ------------------------------------------------------------------------
def accuracy_topk(scores, target, topk=(1, 5), thr=0.0):
    scores, target = np.asarray(scores, dtype=np.float64), np.asarray(target)
    res = []
    for k in topk:
        correct = 0
        for row, t in zip(scores, target):
            order = sorted(range(len(row)), key=lambda j: (-row[j], j))[:k]
            if any(j == t and row[j] > thr for j in order):
                correct += 1
        res.append(correct * 100.0 / len(target))
    return res


# ---- mmseg intersect_and_union / total_area_to_metrics -------------------------------------------------
def seg_areas(pred, label, num_classes, ignore_index, reduce_zero_label=False):
    pred, label = np.asarray(pred).astype(np.int64).copy(), np.asarray(label).astype(np.int64).copy()
    if reduce_zero_label:
        label[label == 0] = 255
        label = label - 1
        label[label == 254] = 255
    mask = label != ignore_index
    pred, label = pred[mask], label[mask]
    intersect = pred[pred == label]
    hist = lambda v: np.array([(v == c).sum() for c in range(num_classes)], dtype=np.float64)   # histc(bins=C, 0, C-1)
    ai, ap, al = hist(intersect), hist(pred), hist(label)
    return ai, ap + al - ai, ap, al


def seg_metrics(pre_eval, metrics=('mIoU',), beta=1):
    ti = sum(r[0] for r in pre_eval)
    tu = sum(r[1] for r in pre_eval)
    tp = sum(r[2] for r in pre_eval)
    tl = sum(r[3] for r in pre_eval)
    out = dict(aAcc=ti.sum() / tl.sum())
    with np.errstate(divide='ignore', invalid='ignore'):
        for m in metrics:
            if m == 'mIoU':
                out['IoU'], out['Acc'] = ti / tu, ti / tl
            elif m == 'mFscore':
                p, r = ti / tp, ti / tl
                out['Fscore'] = np.array([(1 + beta ** 2) * (a * b) / ((beta ** 2 * a) + b) for a, b in zip(p, r)])
                out['Precision'], out['Recall'] = p, r
    return out


# ---- pycocotools COCOeval (bbox) -------------------------------------------------------------------
class CocoEvalOracle:
    """evaluate() -> evalImgs, accumulate() -> eval['precision'|'recall'], summarize() -> stats."""

    def __init__(self, gts, dts, img_ids, cat_ids, iou_thrs=None, max_dets=(1, 10, 100)):
        self.gts, self.dts = {}, {}
        for g in gts:
            g = dict(g)
            g['_ignore'] = 1 if g.get('iscrowd', 0) else 0
            self.gts.setdefault((g['image_id'], g['category_id']), []).append(g)
        for i, d in enumerate(dts):
            d = dict(d)
            d['area'] = d['bbox'][2] * d['bbox'][3]
            d['id'] = i + 1
            self.dts.setdefault((d['image_id'], d['category_id']), []).append(d)
        self.img_ids, self.cat_ids = sorted(set(img_ids)), sorted(set(cat_ids))
        self.iou_thrs = list(iou_thrs) if iou_thrs is not None else [0.5 + 0.05 * i for i in range(10)]
        self.rec_thrs = [i / 100.0 for i in range(101)]
        self.max_dets = list(max_dets)
        self.area_rng = [[0, 1e10], [0, 32 ** 2], [32 ** 2, 96 ** 2], [96 ** 2, 1e10]]

    @staticmethod
    def _iou(d, g, crowd):
        dx, dy, dw, dh = d
        gx, gy, gw, gh = g
        w = min(dx + dw, gx + gw) - max(dx, gx)
        h = min(dy + dh, gy + gh) - max(dy, gy)
        if w <= 0 or h <= 0:
            return 0.0
        i = w * h
        u = dw * dh if crowd else dw * dh + gw * gh - i
        return i / u

    def compute_iou(self, img, cat):
        gt, dt = self.gts.get((img, cat), []), self.dts.get((img, cat), [])
        if not gt and not dt:
            return []
        dt = sorted(dt, key=lambda d: -d['score'])[:self.max_dets[-1]]       # (python's sort is stable, like mergesort)
        return [[self._iou(d['bbox'], g['bbox'], bool(g.get('iscrowd', 0))) for g in gt] for d in dt]

    def evaluate_img(self, img, cat, rng, max_det):
        gt, dt = self.gts.get((img, cat), []), self.dts.get((img, cat), [])
        if not gt and not dt:
            return None
        ig = [1 if (g['_ignore'] or g['area'] < rng[0] or g['area'] > rng[1]) else 0 for g in gt]
        gtind = sorted(range(len(gt)), key=lambda i: ig[i])
        gt = [gt[i] for i in gtind]
        g_ig = [ig[i] for i in gtind]
        dt = sorted(dt, key=lambda d: -d['score'])[:max_det]
        ious_all = self.ious[img, cat]
        ious = [[row[i] for i in gtind] for row in ious_all] if len(ious_all) else ious_all
        T, G, D = len(self.iou_thrs), len(gt), len(dt)
        gtm = [[0] * G for _ in range(T)]
        dtm = [[0] * D for _ in range(T)]
        dt_ig = [[0] * D for _ in range(T)]
        if G and D:
            for ti, t in enumerate(self.iou_thrs):
                for di in range(D):
                    iou, m = min(t, 1 - 1e-10), -1
                    for gi in range(G):
                        if gtm[ti][gi] > 0 and not gt[gi].get('iscrowd', 0):
                            continue
                        if m > -1 and g_ig[m] == 0 and g_ig[gi] == 1:
                            break
                        if ious[di][gi] < iou:
                            continue
                        iou, m = ious[di][gi], gi
                    if m == -1:
                        continue
                    dt_ig[ti][di] = g_ig[m]
                    dtm[ti][di] = m + 1                       # (stand-in for the gt id: anything > 0)
                    gtm[ti][m] = dt[di]['id']
        for ti in range(T):
            for di, d in enumerate(dt):
                if dtm[ti][di] == 0 and (d['area'] < rng[0] or d['area'] > rng[1]):
                    dt_ig[ti][di] = 1
        return dict(dtm=dtm, dt_ig=dt_ig, scores=[d['score'] for d in dt], g_ig=g_ig)

    def evaluate(self):
        self.ious = {(i, c): self.compute_iou(i, c) for i in self.img_ids for c in self.cat_ids}
        self.eval_imgs = [self.evaluate_img(i, c, rng, self.max_dets[-1])
                          for c in self.cat_ids for rng in self.area_rng for i in self.img_ids]

    def accumulate(self):
        T, R, K, A, M = len(self.iou_thrs), len(self.rec_thrs), len(self.cat_ids), len(self.area_rng), len(self.max_dets)
        I = len(self.img_ids)
        precision, recall = -np.ones((T, R, K, A, M)), -np.ones((T, K, A, M))
        for k in range(K):
            for a in range(A):
                for m, max_det in enumerate(self.max_dets):
                    E = [self.eval_imgs[k * A * I + a * I + i] for i in range(I)]
                    E = [e for e in E if e is not None]
                    if not E:
                        continue
                    scores = np.array([s for e in E for s in e['scores'][:max_det]])
                    inds = np.argsort(-scores, kind='mergesort')
                    npig = sum(1 for e in E for x in e['g_ig'] if x == 0)
                    if npig == 0:
                        continue
                    for t in range(T):
                        dtm = np.array([x for e in E for x in e['dtm'][t][:max_det]])[inds] if len(inds) else np.array([])
                        dig = np.array([x for e in E for x in e['dt_ig'][t][:max_det]])[inds] if len(inds) else np.array([])
                        tp = fp = 0.0
                        rc, pr = [], []
                        for x, g in zip(dtm, dig):
                            if x != 0 and not g:
                                tp += 1
                            if x == 0 and not g:
                                fp += 1
                            rc.append(tp / npig)
                            pr.append(tp / (fp + tp + np.spacing(1)))
                        nd = len(rc)
                        recall[t, k, a, m] = rc[-1] if nd else 0
                        for i in range(nd - 1, 0, -1):
                            if pr[i] > pr[i - 1]:
                                pr[i - 1] = pr[i]
                        q = [0.0] * R
                        for ri, thr in enumerate(self.rec_thrs):
                            pi = 0
                            while pi < nd and rc[pi] < thr:        # searchsorted(side='left')
                                pi += 1
                            if pi >= nd:
                                break
                            q[ri] = pr[pi]
                        precision[t, :, k, a, m] = q
        self.precision, self.recall = precision, recall

    def _summ(self, ap, iou=None, area=0, m=None):
        m = len(self.max_dets) - 1 if m is None else m
        s = self.precision[:, :, :, area, m] if ap else self.recall[:, :, area, m]
        if iou is not None:
            sel = [i for i, t in enumerate(self.iou_thrs) if abs(t - iou) < 1e-9]
            s = s[sel]
        s = s[s > -1]
        return float(np.mean(s)) if s.size else -1.0

    def summarize(self):
        M = len(self.max_dets)
        return [self._summ(1), self._summ(1, .5), self._summ(1, .75), self._summ(1, None, 1), self._summ(1, None, 2),
                self._summ(1, None, 3), self._summ(0, None, 0, 0), self._summ(0, None, 0, min(1, M - 1)),
                self._summ(0, None, 0, M - 1), self._summ(0, None, 1), self._summ(0, None, 2), self._summ(0, None, 3)]
