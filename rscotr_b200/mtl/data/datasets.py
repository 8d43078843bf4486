"""The three dataset types the reference configs name, on top of mtl/data/transforms.py:

* `CustomDataset` (cls, mmcls 0.23 -- configs/_base_/cls/resisc_swin_224.py:4,57-68): one folder per class
  under `data_prefix`, or an `ann_file` of "<relative path> <label>" lines.
* `CocoDataset` (det, mmdet 2.25 -- configs/_base_/det/dior.py:2-56): COCO json parsed directly (pycocotools
  is not needed for reading), training filter (images without gt / smaller than 32 px dropped), crowd boxes to
  `bboxes_ignore`, aspect-ratio group flag.
* `PotsdamDataset` (seg, mmseg 0.28 -- configs/_base_/seg/potsdam_IRRG_all.py:2,58-80): `img_dir` / `ann_dir` pairs,
  `reduce_zero_label=True`, 6 classes, `ignore_index` from the config.

Each one implements `evaluate(results, **eval_kwargs[task])` through mtl/data/metrics.py, which is what
MultiDatasetsEvalHook calls (reference mtl/runner/hooks/evaluation.py:130-142)."""
import json
import os
from collections import defaultdict

import numpy as np
import torch

from . import metrics
from .transforms import Compose

DATASETS = {}


def register(cls):
    DATASETS[cls.__name__] = cls
    return cls


def build_dataset(cfg, task, default_args=None):
    cfg = dict(cfg)
    t = cfg.pop('type')
    if t in ('RepeatDataset', 'ConcatDataset', 'ClassBalancedDataset'):
        raise KeyError('dataset wrapper %s is not used by any reference config' % t)
    if t not in DATASETS:
        raise KeyError('%s is not in the dataset registry' % t)
    for k, v in (default_args or {}).items():
        cfg.setdefault(k, v)
    ds = DATASETS[t](task=task, **cfg)
    return ds


class _Base(torch.utils.data.Dataset):
    CLASSES = None

    def __len__(self):
        return len(self.data_infos)

    def _rand_another(self, idx):
        return int(np.random.choice(len(self)))


# ---------------------------------------------------------------------------------------------------- cls
_IMG_EXT = ('.jpg', '.jpeg', '.png', '.ppm', '.bmp', '.pgm', '.tif')


@register
class CustomDataset(_Base):
    def __init__(self, data_prefix, pipeline, classes=None, ann_file=None, extensions=_IMG_EXT, test_mode=False, task='cls'):
        self.task, self.data_prefix, self.ann_file, self.test_mode = task, os.path.expanduser(data_prefix), ann_file, test_mode
        self.extensions = tuple(e.lower() for e in extensions)
        self.pipeline = Compose(pipeline, task)
        self.CLASSES = self._classes(classes)
        self.data_infos = self._load()

    @staticmethod
    def _classes(classes):
        if classes is None:
            return None
        if isinstance(classes, str):
            with open(classes) as f:
                return [l.strip() for l in f if l.strip()]
        return list(classes)

    def _load(self):
        samples = []
        if self.ann_file is None:
            folders = sorted(d for d in os.listdir(self.data_prefix) if os.path.isdir(os.path.join(self.data_prefix, d)))
            if not folders:
                raise RuntimeError('Found no valid folders in %s' % self.data_prefix)
            to_idx = {f: i for i, f in enumerate(folders)}
            for f in folders:
                for root, _, files in sorted(os.walk(os.path.join(self.data_prefix, f), followlinks=True)):
                    for name in sorted(files):
                        if name.lower().endswith(self.extensions):
                            samples.append((os.path.relpath(os.path.join(root, name), self.data_prefix), to_idx[f]))
            if not samples:
                raise RuntimeError('Found 0 files in subfolders of: %s. Supported extensions are: %s'
                                   % (self.data_prefix, ','.join(self.extensions)))
            if self.CLASSES is None:
                self.CLASSES = folders
            else:
                assert len(self.CLASSES) == len(folders), 'the number of folders does not match `classes`'
            self.folder_to_idx = to_idx
        else:
            with open(self.ann_file) as f:
                for line in f:
                    if line.strip():
                        name, label = line.strip().rsplit(' ', 1)
                        samples.append((name, int(label)))
        return [dict(img_prefix=self.data_prefix, img_info=dict(filename=n), gt_label=np.array(l, dtype=np.int64))
                for n, l in samples]

    def get_gt_labels(self):
        return np.array([d['gt_label'] for d in self.data_infos])

    def get_cat_ids(self, idx):
        return [int(self.data_infos[idx]['gt_label'])]

    def __getitem__(self, idx):
        d = self.data_infos[idx]
        return self.pipeline(dict(img_prefix=d['img_prefix'], img_info=dict(d['img_info']), gt_label=d['gt_label'].copy()))

    def evaluate(self, results, metric='accuracy', metric_options=None, indices=None, logger=None, **kwargs):
        gt = self.get_gt_labels()
        if indices is not None:
            gt = gt[indices]
        return metrics.evaluate_cls(results, gt, metric, metric_options)


# ---------------------------------------------------------------------------------------------------- det
@register
class CocoDataset(_Base):
    def __init__(self, ann_file, pipeline, classes=None, data_root=None, img_prefix='', test_mode=False, filter_empty_gt=True,
                 task='det', **kwargs):
        self.task, self.test_mode, self.filter_empty_gt = task, test_mode, filter_empty_gt
        if data_root is not None:
            ann_file = ann_file if os.path.isabs(ann_file) else os.path.join(data_root, ann_file)
            img_prefix = img_prefix if (not img_prefix or os.path.isabs(img_prefix)) else os.path.join(data_root, img_prefix)
        self.ann_file, self.img_prefix = ann_file, img_prefix
        with open(ann_file) as f:
            coco = json.load(f)
        names = tuple(classes) if classes is not None else tuple(c['name'] for c in sorted(coco['categories'], key=lambda c: c['id']))
        self.CLASSES = names
        by_name = defaultdict(list)
        for c in coco['categories']:
            by_name[c['name']].append(c['id'])
        self.cat_ids = [i for n in names for i in by_name.get(n, [])]          # COCO.get_cat_ids(cat_names=CLASSES)
        self.cat2label = {c: i for i, c in enumerate(self.cat_ids)}
        self.anns = defaultdict(list)
        for a in coco.get('annotations', []):
            self.anns[a['image_id']].append(a)
        self.data_infos = []
        for im in coco['images']:                                  # (COCO.get_img_ids(): json order)
            info = dict(im)
            info['filename'] = info['file_name']
            self.data_infos.append(info)
        self.img_ids = [i['id'] for i in self.data_infos]
        assert len(set(self.img_ids)) == len(self.img_ids), "Annotation ids in '%s' are not unique!" % ann_file
        if not test_mode:
            keep = self._filter_imgs()
            self.data_infos = [self.data_infos[i] for i in keep]
            self.img_ids = [self.img_ids[i] for i in keep]
            self.flag = np.array([1 if i['width'] / i['height'] > 1 else 0 for i in self.data_infos], dtype=np.uint8)
        self.pipeline = Compose(pipeline, task)

    def _filter_imgs(self, min_size=32):
        with_ann = {a['image_id'] for anns in self.anns.values() for a in anns}
        in_cat = {a['image_id'] for anns in self.anns.values() for a in anns if a['category_id'] in self.cat2label}
        in_cat &= with_ann
        keep = []
        for i, info in enumerate(self.data_infos):
            if self.filter_empty_gt and info['id'] not in in_cat:
                continue
            if min(info['width'], info['height']) >= min_size:
                keep.append(i)
        return keep

    def get_ann_info(self, idx):
        info = self.data_infos[idx]
        boxes, labels, ignore = [], [], []
        for a in self.anns.get(info['id'], []):
            if a.get('ignore', False):
                continue
            x1, y1, w, h = a['bbox']
            iw = max(0, min(x1 + w, info['width']) - max(x1, 0))
            ih = max(0, min(y1 + h, info['height']) - max(y1, 0))
            if iw * ih == 0 or a.get('area', w * h) <= 0 or w < 1 or h < 1:
                continue
            if a['category_id'] not in self.cat2label:
                continue
            box = [x1, y1, x1 + w, y1 + h]
            if a.get('iscrowd', False):
                ignore.append(box)
            else:
                boxes.append(box)
                labels.append(self.cat2label[a['category_id']])
        return dict(bboxes=np.array(boxes, dtype=np.float32).reshape(-1, 4), labels=np.array(labels, dtype=np.int64),
                    bboxes_ignore=np.array(ignore, dtype=np.float32).reshape(-1, 4))

    def get_cat_ids(self, idx):
        return [a['category_id'] for a in self.anns.get(self.data_infos[idx]['id'], [])]

    def _prepare(self, idx, train):
        r = dict(img_info=dict(self.data_infos[idx]), img_prefix=self.img_prefix, bbox_fields=[], seg_fields=[])
        if train:
            r['ann_info'] = self.get_ann_info(idx)
        return self.pipeline(r)

    def __getitem__(self, idx):
        if self.test_mode:
            return self._prepare(idx, False)
        while True:
            data = self._prepare(idx, True)
            if data is None:
                pool = np.where(self.flag == self.flag[idx])[0]
                idx = int(np.random.choice(pool))
                continue
            return data

    def coco_gts(self):
        gts = []
        for info in self.data_infos:
            for a in self.anns.get(info['id'], []):
                gts.append(dict(image_id=info['id'], category_id=a['category_id'], bbox=list(a['bbox']),
                                area=a.get('area', a['bbox'][2] * a['bbox'][3]), iscrowd=int(a.get('iscrowd', 0))))
        return [g for g in gts if g['category_id'] in self.cat2label]

    def evaluate(self, results, metric='bbox', logger=None, jsonfile_prefix=None, classwise=False,
                 proposal_nums=(100, 300, 1000), iou_thrs=None, metric_items=None, **kwargs):
        assert len(results) == len(self), 'The length of results is not equal to the dataset len: %d != %d' % (len(results), len(self))
        return metrics.evaluate_det(results, self.coco_gts(), self.img_ids, self.cat_ids, self.CLASSES, metric, iou_thrs,
                                    classwise, proposal_nums, metric_items)


# ---------------------------------------------------------------------------------------------------- seg
class SegCustomDataset(_Base):
    CLASSES, PALETTE = None, None

    def __init__(self, pipeline, img_dir, img_suffix='.jpg', ann_dir=None, seg_map_suffix='.png', split=None, data_root=None,
                 test_mode=False, ignore_index=255, reduce_zero_label=False, classes=None, palette=None, task='seg', **kwargs):
        self.task, self.test_mode, self.ignore_index, self.reduce_zero_label = task, test_mode, ignore_index, reduce_zero_label
        self.img_suffix, self.seg_map_suffix, self.label_map = img_suffix, seg_map_suffix, None
        if classes is not None:
            self.CLASSES = tuple(classes)
        if data_root is not None:
            img_dir = img_dir if os.path.isabs(img_dir) else os.path.join(data_root, img_dir)
            ann_dir = ann_dir if (ann_dir is None or os.path.isabs(ann_dir)) else os.path.join(data_root, ann_dir)
            split = split if (split is None or os.path.isabs(split)) else os.path.join(data_root, split)
        self.img_dir, self.ann_dir = img_dir, ann_dir
        infos = []
        if split is not None:
            with open(split) as f:
                for line in f:
                    n = line.strip()
                    if n:
                        infos.append(dict(filename=n + img_suffix, ann=dict(seg_map=n + seg_map_suffix)))
        else:
            for root, _, files in os.walk(img_dir, followlinks=True):
                for name in files:
                    if name.endswith(img_suffix):
                        rel = os.path.relpath(os.path.join(root, name), img_dir)
                        infos.append(dict(filename=rel, ann=dict(seg_map=rel[:-len(img_suffix)] + seg_map_suffix)))
            infos = sorted(infos, key=lambda x: x['filename'])
        self.data_infos = self.img_infos = infos
        self.pipeline = Compose(pipeline, task)

    def _pre(self, idx):
        info = self.data_infos[idx]
        return dict(img_info=dict(info), ann_info=info['ann'], seg_fields=[], img_prefix=self.img_dir, seg_prefix=self.ann_dir,
                    label_map=self.label_map)

    def __getitem__(self, idx):
        return self.pipeline(self._pre(idx))

    def get_gt_seg_map_by_idx(self, idx):
        from PIL import Image
        return np.array(Image.open(os.path.join(self.ann_dir, self.data_infos[idx]['ann']['seg_map']))).squeeze().astype(np.uint8)

    def pre_eval(self, preds, indices):
        preds = preds if isinstance(preds, list) else [preds]
        indices = indices if isinstance(indices, list) else [indices]
        return [metrics.intersect_and_union(p, torch.from_numpy(self.get_gt_seg_map_by_idx(i).astype(np.int64)),
                                            len(self.CLASSES), self.ignore_index, self.reduce_zero_label)
                for p, i in zip(preds, indices)]

    def evaluate(self, results, metric='mIoU', logger=None, gt_seg_maps=None, pre_eval=None, classwise=None, **kwargs):
        first = results[0]
        if not (isinstance(first, tuple) and len(first) == 4):          # raw label maps: reduce them now
            results = self.pre_eval(list(results), list(range(len(self))))
        return metrics.evaluate_seg(results, self.CLASSES, metric)


@register
class PotsdamDataset(SegCustomDataset):
    CLASSES = ('impervious_surface', 'building', 'low_vegetation', 'tree', 'car', 'clutter')
    PALETTE = [[255, 255, 255], [0, 0, 255], [0, 255, 255], [0, 255, 0], [255, 255, 0], [255, 0, 0]]

    def __init__(self, **kwargs):
        kwargs.setdefault('img_suffix', '.png')
        kwargs.setdefault('seg_map_suffix', '.png')
        kwargs.setdefault('reduce_zero_label', True)
        super().__init__(**kwargs)
