"""Seeded synthetic datasets with the batch formats of the three reference pipelines
(SURVEY 8d 'Synthetic inputs'): there is no network / no dataset in this environment,
so the real mmcls / mmdet / mmseg dataset builders are replaced by these.  Batches are
produced in PINNED host memory; the step engine copies them to the device."""
import torch


class _SyntheticLoader:
    """A loader that yields `length` pre-generated batches (host tensors)."""

    def __init__(self, dataset, batch_size, length, seed, pin=True):
        self.dataset, self.batch_size, self.length, self.seed, self.pin = dataset, batch_size, length, seed, pin
        self._cache = None
        dataset._loader = self

    def __len__(self):
        return self.length

    def _make(self):
        g = torch.Generator().manual_seed(self.seed)
        n_distinct = min(self.length, self.dataset.distinct)
        return [self.dataset.make_batch(self.batch_size, g, self.pin) for _ in range(n_distinct)]

    def served(self):
        """the batches of one pass, in serving order (ground truth for dataset.evaluate())."""
        if self._cache is None:
            self._cache = self._make()
        return [self._cache[i % len(self._cache)] for i in range(self.length)]

    def __iter__(self):
        for b in self.served():
            yield dict(b)


def _pin(t, pin):
    return t.pin_memory() if pin and torch.cuda.is_available() else t


class SyntheticDataset:
    distinct = 2
    CLASSES = None

    def __init__(self, task, img_size=(800, 800), num_classes=None, num_boxes=8, dtype=torch.float32):
        self.task, self.img_size, self.num_boxes, self.dtype = task, tuple(img_size), num_boxes, dtype
        self.num_classes = num_classes or dict(cls=45, det=20, seg=5)[task]

    def evaluate(self, results, metric=None, **kwargs):
        """same metric code as the real datasets (mtl/data/metrics.py), against the batches the loader served."""
        from . import metrics as M
        served = self._loader.served()
        if self.task == 'cls':
            gt = torch.cat([b['gt_label'] for b in served]).numpy()
            return M.evaluate_cls(results, gt, metric or 'accuracy', kwargs.get('metric_options'))
        if self.task == 'det':
            gts, n = [], 0
            for b in served:
                for boxes, labels in zip(b['gt_bboxes'], b['gt_labels']):
                    for (x1, y1, x2, y2), l in zip(boxes.tolist(), labels.tolist()):
                        gts.append(dict(image_id=n, category_id=l + 1, bbox=[x1, y1, x2 - x1, y2 - y1],
                                        area=(x2 - x1) * (y2 - y1), iscrowd=0))
                    n += 1
            return M.evaluate_det(results, gts, list(range(n)), list(range(1, self.num_classes + 1)), None, metric or 'bbox',
                                  kwargs.get('iou_thrs'), kwargs.get('classwise', False))
        labels = [l[0] for b in served for l in b['gt_semantic_seg']]
        pre = [M.intersect_and_union(torch.as_tensor(p), l, self.num_classes, 255) for p, l in zip(results, labels)]
        return M.evaluate_seg(pre, [str(c) for c in range(self.num_classes)], metric or 'mIoU')

    def _metas(self, B):
        H, W = self.img_size
        return [dict(img_shape=(H, W, 3), ori_shape=(H, W, 3), pad_shape=(H, W, 3), scale_factor=1.0, flip=False)
                for _ in range(B)]

    def make_batch(self, B, g, pin=True):
        H, W = self.img_size
        img = _pin(torch.randn(B, 3, H, W, generator=g).to(self.dtype), pin)
        out = dict(img=img, img_metas=self._metas(B))
        if self.task == 'cls':
            out['gt_label'] = _pin(torch.randint(0, self.num_classes, (B,), generator=g), pin)
        elif self.task == 'det':
            boxes, labels = [], []
            for _ in range(B):
                K = self.num_boxes
                x1 = torch.rand(K, generator=g) * (W - 100)
                y1 = torch.rand(K, generator=g) * (H - 100)
                w = 16 + torch.rand(K, generator=g) * 84
                h = 16 + torch.rand(K, generator=g) * 84
                boxes.append(_pin(torch.stack([x1, y1, (x1 + w).clamp(max=W), (y1 + h).clamp(max=H)], -1), pin))
                labels.append(_pin(torch.randint(0, self.num_classes, (K,), generator=g), pin))
            out['gt_bboxes'], out['gt_labels'] = boxes, labels
        else:
            out['gt_semantic_seg'] = _pin(torch.randint(0, self.num_classes, (B, 1, H, W), generator=g), pin)
        return out
