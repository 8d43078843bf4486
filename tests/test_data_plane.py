"""Real data plane on CPU: the repo's dataset configs (configs/datasets/*.py, checked
equal to the reference's configs/_base_/{cls,det,seg}/*.py when that tree is mounted) drive the
pipelines over tiny on-disk fixtures; geometric / photometric invariants, batch formats, deferred
normalisation, evaluation through the dataset classes."""
import os
import random

import numpy as np
import pytest
import torch

from rscotr_b200.config import Config
from rscotr_b200.mtl.data import transforms as T
from rscotr_b200.mtl.data.datasets import build_dataset
from rscotr_b200.mtl.data.loader import InfiniteGroupBatchSampler, build_dataloader, collate
from rscotr_b200.models.mtl import normalize_on_device
from tests import data_fixtures as FX

HERE = os.path.dirname(os.path.abspath(__file__))
CFG_ROOT = os.path.join(os.path.dirname(HERE), 'configs', 'datasets')
REF_ROOT = '/root/reference/configs/_base_'


def _cfg(rel):
    return Config.fromfile(os.path.join(CFG_ROOT, rel))._cfg_dict['data']


def _seed(s=0):
    np.random.seed(s)
    random.seed(s)
    torch.manual_seed(s)


def _plain(x):
    if isinstance(x, dict):
        return {k: _plain(v) for k, v in x.items()}
    return [_plain(v) for v in x] if isinstance(x, (list, tuple)) else x


@pytest.mark.skipif(not os.path.isdir(REF_ROOT), reason='reference tree not mounted (GPU box)')
@pytest.mark.parametrize('own,ref', [('dior.py', 'det/dior.py'), ('potsdam.py', 'seg/potsdam_IRRG_all.py'),
                                     ('resisc45.py', 'cls/resisc_swin_224.py')])
def test_dataset_configs_equal_the_reference(own, ref):
    a = Config.fromfile(os.path.join(CFG_ROOT, own))._cfg_dict
    b = Config.fromfile(os.path.join(REF_ROOT, ref))._cfg_dict
    assert _plain(dict(a['data'])) == _plain(dict(b['data']))
    if 'evaluation' in b:
        assert _plain(dict(a['evaluation'])) == _plain(dict(b['evaluation']))
    # ... and the reference's own files drive the same dataset classes unmodified
    from rscotr_b200.mtl.data.transforms import Compose
    task = dict(d='det', p='seg', r='cls')[own[0]]
    for split in ('train', 'val', 'test'):
        Compose(b['data'][split]['pipeline'], task)


# ------------------------------------------------------------------------------------------ det
@pytest.fixture(scope='module')
def dior(tmp_path_factory):
    return FX.make_dior(str(tmp_path_factory.mktemp('dior')))


def _det_cfg(root, split='train'):
    d = dict(_cfg('dior.py')[split])
    d['ann_file'] = os.path.join(root, 'coco_ann', 'DIOR_%s_coco.json' % split)
    d['img_prefix'] = os.path.join(root, 'JPEGImages-trainval')
    return d


def test_det_pipeline_keeps_boxes_on_their_objects(dior):
    ds = build_dataset(_det_cfg(dior), 'det')
    assert len(ds) == 4 and ds.CLASSES == FX.DIOR_CLASSES          # the image without gt is filtered in training
    assert set(ds.flag.tolist()) == {0, 1}
    mean, std = np.array([123.675, 116.28, 103.53]), np.array([58.395, 57.12, 57.375])
    flips = 0
    for s in range(6):
        _seed(s)
        d = ds[s % len(ds)]
        img, m = d['img'], d['img_metas']
        assert img.dtype == torch.float32 and img.shape[0] == 3 and img.shape[1] % 32 == 0 and img.shape[2] % 32 == 0
        h, w = m['img_shape'][:2]
        assert m['pad_shape'][:2] == tuple(img.shape[1:]) and max(h, w) <= 1333 and min(h, w) <= 800
        assert (img[:, h:, :] == 0).all() and (img[:, :, w:] == 0).all()              # padded AFTER normalising, with 0
        assert d['gt_bboxes'].dtype == torch.float32 and d['gt_labels'].dtype == torch.int64
        flips += int(m['flip'])
        # un-normalise: bright (230) inside every box, dark (30) just outside it
        raw = img[:, :h, :w].permute(1, 2, 0).numpy() * std + mean
        for x1, y1, x2, y2 in d['gt_bboxes'].numpy():
            inner = raw[int(y1) + 3:int(y2) - 3, int(x1) + 3:int(x2) - 3]
            assert inner.size and inner.mean() > 200, (s, x1, y1, x2, y2)
        sf = m['scale_factor']
        assert sf.shape == (4,) and abs(sf[0] - w / m['ori_shape'][1]) < 1e-6
    assert 0 < flips < 6


def test_det_test_pipeline_and_coco_evaluate(dior):
    ds = build_dataset(_det_cfg(dior, 'val'), 'det', dict(test_mode=True))
    assert len(ds) == 5                                              # no filtering in test mode
    d = ds[0]
    assert isinstance(d['img'], list) and len(d['img']) == 1 and d['img'][0].shape[0] == 3
    assert d['img_metas'][0]['flip'] is False
    # perfect detections built from the annotations -> mAP 1; shifted ones -> 0
    perfect, off = [], []
    for i in range(len(ds)):
        ann = ds.get_ann_info(i)
        per, bad = [np.zeros((0, 5), dtype=np.float32) for _ in ds.CLASSES], [np.zeros((0, 5), dtype=np.float32) for _ in ds.CLASSES]
        for b, l in zip(ann['bboxes'], ann['labels']):
            per[l] = np.vstack([per[l], np.r_[b, 0.9][None]])
            bad[l] = np.vstack([bad[l], np.r_[b + 500, 0.9][None]])
        perfect.append(per)
        off.append(bad)
    res = ds.evaluate(perfect, metric='bbox', iou_thrs=[0.5], classwise=True)
    assert res['bbox_mAP'] == 1.0 and res['bbox_mAP_50'] == 1.0
    assert ds.evaluate(off, metric='bbox', iou_thrs=[0.5])['bbox_mAP'] == 0.0


def test_det_loader_is_infinite_and_groups_by_aspect(dior):
    ds = build_dataset(_det_cfg(dior), 'det')
    loader = build_dataloader(ds, samples_per_gpu=2, workers_per_gpu=0, dist=False, seed=3, runner_type='IterBasedRunner')
    it = iter(loader)
    seen = 0
    for _ in range(5):                                               # more batches than the dataset holds
        b = next(it)
        assert b['img'].dim() == 4 and b['img'].shape[0] == 2 and len(b['gt_bboxes']) == 2 and len(b['img_metas']) == 2
        landscape = [m['ori_shape'][1] > m['ori_shape'][0] for m in b['img_metas']]
        assert landscape[0] == landscape[1]
        seen += 2
    assert seen > len(ds)
    # two ranks draw disjoint index streams from the same permutation
    a = InfiniteGroupBatchSampler(ds, 1, 2, 0, seed=5)
    c = InfiniteGroupBatchSampler(ds, 1, 2, 1, seed=5)
    ia, ic = iter(a), iter(c)
    first_a, first_c = [next(ia)[0] for _ in range(2)], [next(ic)[0] for _ in range(2)]
    assert sorted(first_a + first_c) == [0, 1, 2, 3]


def test_collate_pads_to_the_largest_image():
    a = dict(img=torch.ones(3, 32, 64), img_metas=dict(k=1), gt_bboxes=torch.zeros(2, 4), gt_labels=torch.zeros(2, dtype=torch.long))
    b = dict(img=torch.ones(3, 64, 32), img_metas=dict(k=2), gt_bboxes=torch.zeros(1, 4), gt_labels=torch.zeros(1, dtype=torch.long))
    out = collate([a, b])
    assert out['img'].shape == (2, 3, 64, 64) and float(out['img'][0, :, 32:].abs().sum()) == 0
    assert [m['k'] for m in out['img_metas']] == [1, 2] and [len(x) for x in out['gt_bboxes']] == [2, 1]
    t = collate([dict(img=[torch.ones(3, 8, 8)], img_metas=[dict(k=1)]), dict(img=[torch.ones(3, 8, 8)], img_metas=[dict(k=2)])])
    assert isinstance(t['img'], list) and t['img'][0].shape == (2, 3, 8, 8) and [m['k'] for m in t['img_metas'][0]] == [1, 2]


# ------------------------------------------------------------------------------------------ seg
@pytest.fixture(scope='module')
def potsdam(tmp_path_factory):
    return FX.make_potsdam(str(tmp_path_factory.mktemp('potsdam')))


def _seg_cfg(root, split='train', drop=()):
    d = dict(_cfg('potsdam.py')[split])
    d['data_root'] = root
    d['pipeline'] = [t for t in d['pipeline'] if t['type'] not in drop]
    return d


def test_seg_pipeline_keeps_image_and_label_aligned(potsdam):
    ds = build_dataset(_seg_cfg(potsdam, drop=('PhotoMetricDistortion',)), 'seg')
    assert len(ds) == 3 and ds.reduce_zero_label and ds.ignore_index == 5 and len(ds.CLASSES) == 6
    mean, std = np.array([123.675, 116.28, 103.53]), np.array([58.395, 57.12, 57.375])
    for s in range(5):
        _seed(s)
        d = ds[s % 3]
        img, seg, m = d['img'], d['gt_semantic_seg'], d['img_metas']
        assert img.shape == (3, 512, 512) and seg.shape == (1, 512, 512) and seg.dtype == torch.int64
        h, w = m['img_shape'][:2]
        assert (seg[0, h:, :] == 5).all() and (seg[0, :, w:] == 5).all()               # seg_pad_val of the config
        assert (img[:, h:, :] == 0).all()
        # image value = 40 * (label + 1) in the un-padded area (fixture), away from the block borders the
        # bilinear / nearest resize pair disagrees on
        raw = (img[:, :h, :w].permute(1, 2, 0).numpy() * std + mean)[..., 0]
        lab = seg[0, :h, :w].numpy()
        agree = np.abs(raw - 40.0 * (lab + 1)) < 1.0
        assert agree.mean() > 0.9, (s, agree.mean())
        assert set(np.unique(lab)).issubset({0, 1, 2, 3, 4, 5})


def test_seg_photometric_distortion_is_reproducible_and_uint8(potsdam):
    ds = build_dataset(_seg_cfg(potsdam), 'seg')
    _seed(4)
    a = ds[0]
    _seed(4)
    b = ds[0]
    assert torch.equal(a['img'], b['img']) and torch.equal(a['gt_semantic_seg'], b['gt_semantic_seg'])
    _seed(5)
    c = ds[0]
    assert not torch.equal(a['img'], c['img'])


def test_seg_val_pipeline_and_evaluate(potsdam):
    ds = build_dataset(_seg_cfg(potsdam, 'val'), 'seg', dict(test_mode=True))
    d = ds[1]
    assert isinstance(d['img'], list) and d['img'][0].shape == (3, 512, 512)      # 96 px tile rescaled to fit (512, 512)
    assert d['img_metas'][0]['ori_shape'][:2] == (96, 96)
    # predicting the ground truth (label - 1; clutter mapped to any class, it is ignored) -> every metric 1
    preds = []
    for i in range(len(ds)):
        lab = ds.get_gt_seg_map_by_idx(i).astype(np.int64) - 1
        lab[lab == 5] = 0
        preds.append(lab)
    res = ds.evaluate(preds, metric=['mFscore', 'mIoU'], pre_eval=True, classwise=True)
    assert res['mIoU'] == 1.0 and res['mFscore'] == 1.0 and res['aAcc'] == 1.0
    assert 'IoU.building' in res and np.isnan(res['IoU.clutter'])
    # pre-reduced results (what the test loop hands over) give the same numbers
    pre = ds.pre_eval(preds, list(range(len(ds))))
    assert ds.evaluate(pre, metric=['mFscore', 'mIoU']) == pytest.approx(res, nan_ok=True)
    # all-wrong predictions
    wrong = [(p + 1) % 5 for p in preds]
    assert ds.evaluate(wrong, metric='mIoU')['mIoU'] == 0.0


# ------------------------------------------------------------------------------------------ cls
@pytest.fixture(scope='module')
def resisc(tmp_path_factory):
    return FX.make_resisc(str(tmp_path_factory.mktemp('resisc')))


def _cls_cfg(root, split='train'):
    d = dict(_cfg('resisc45.py')[split])
    d['data_prefix'] = os.path.join(root, split)
    return d


def test_cls_pipeline_and_folder_dataset(resisc):
    ds = build_dataset(_cls_cfg(resisc), 'cls')
    assert ds.CLASSES == ['airport', 'beach', 'forest'] and len(ds) == 12
    assert ds.get_gt_labels().tolist() == [0] * 4 + [1] * 4 + [2] * 4
    for s in range(8):                       # RandAugment draws 2 of 15 policies: 8 seeds exercise most of them
        _seed(s)
        d = ds[s]
        assert d['img'].shape == (3, 224, 224) and d['img'].dtype == torch.float32 and torch.isfinite(d['img']).all()
        assert int(d['gt_label']) == s // 4 and d['gt_label'].dtype == torch.int64
    loader = build_dataloader(ds, samples_per_gpu=4, workers_per_gpu=0, dist=False, seed=0)
    b = next(iter(loader))
    assert b['img'].shape == (4, 3, 224, 224) and b['gt_label'].shape == (4,) and len(b['img_metas']) == 4
    val = build_dataset(_cls_cfg(resisc, 'val'), 'cls', dict(test_mode=True))
    v = val[0]
    assert v['img'].shape == (3, 224, 224) and 'gt_label' not in v
    scores = [np.eye(3)[l] for l in val.get_gt_labels()]
    assert val.evaluate(scores, metric='accuracy', metric_options=dict(topk=(1,)))['accuracy_top-1'] == 100.0


@pytest.mark.parametrize('name', ['AutoContrast', 'Equalize', 'Invert', 'Rotate', 'Posterize', 'Solarize', 'SolarizeAdd',
                                  'ColorTransform', 'Contrast', 'Brightness', 'Sharpness', 'Shear', 'Translate'])
def test_rand_augment_policies_run_and_respect_identity_magnitude(name):
    rng = np.random.default_rng(0)
    img = rng.integers(40, 200, (40, 48, 3), dtype=np.uint8)          # (reduced range: AutoContrast has work to do)
    ident = dict(Rotate=dict(angle=0.), Posterize=dict(bits=8), Solarize=dict(thr=256), SolarizeAdd=dict(magnitude=0),
                 ColorTransform=dict(magnitude=0.), Contrast=dict(magnitude=0.), Brightness=dict(magnitude=0.),
                 Sharpness=dict(magnitude=0.), Shear=dict(magnitude=0.), Translate=dict(magnitude=0.))
    strong = dict(Rotate=dict(angle=30.), Posterize=dict(bits=2), Solarize=dict(thr=64), SolarizeAdd=dict(magnitude=110),
                  ColorTransform=dict(magnitude=0.9), Contrast=dict(magnitude=0.9), Brightness=dict(magnitude=0.9),
                  Sharpness=dict(magnitude=0.9), Shear=dict(magnitude=0.3), Translate=dict(magnitude=0.45))
    if name in ident:
        out = T.build_transform(dict(type=name, prob=1.0, **ident[name]))(dict(img=img.copy()))['img']
        assert np.array_equal(out, img), name            # zero magnitude = identity
    out = T.build_transform(dict(type=name, prob=1.0, **strong.get(name, {})))(dict(img=img.copy()))['img']
    assert out.shape == img.shape and out.dtype == np.uint8 and not np.array_equal(out, img)
    assert np.array_equal(T.build_transform(dict(type=name, prob=0.0, **strong.get(name, {})))(dict(img=img.copy()))['img'], img)
    if name == 'Invert':
        assert np.array_equal(out, 255 - img)
    if name == 'Translate':                              # shifted by 0.45 * width, border filled with pad_val
        t = T.build_transform(dict(type=name, prob=1.0, random_negative_prob=0., magnitude=0.25, pad_val=7))(dict(img=img.copy()))['img']
        assert np.array_equal(t[:, 12:], img[:, :-12]) and (t[:, :12] == 7).all()


# ------------------------------------------------------------------------------------------ deferred normalise
@pytest.mark.parametrize('task', ['det', 'seg'])
def test_deferred_normalize_equals_host_normalize(task, dior, potsdam):
    base = _det_cfg(dior) if task == 'det' else _seg_cfg(potsdam)
    deferred = dict(base, pipeline=[dict(t, defer=True) if t['type'] == 'Normalize' else t for t in base['pipeline']])
    host, dev = build_dataset(base, task), build_dataset(deferred, task)
    for s in range(3):
        _seed(s)
        a = host[s]
        _seed(s)
        b = dev[s]
        assert b['img'].dtype == torch.uint8 and b['img_metas']['norm_deferred']
        x = normalize_on_device(b['img'][None], [b['img_metas']])[0]
        assert torch.allclose(x, a['img'], atol=1e-4), (task, s, float((x - a['img']).abs().max()))
    # 4x fewer bytes on the wire
    assert b['img'].numel() * b['img'].element_size() * 4 == a['img'].numel() * a['img'].element_size()


# ------------------------------------------------------------------------------------------ end to end
@pytest.mark.timeout(900)
def test_cotraining_and_evaluation_on_real_files(resisc, dior, potsdam, tmp_path):
    """load_data_cfg -> build_datasets -> MultiDataLoader -> StepEngine.train_iter on real batches of the three
    tasks (round robin), then MultiDatasetsEvalHook over the val loaders: the metric keys the reference's
    `save_best` names come out, and a best checkpoint is written."""
    import rscotr_b200.models  # noqa: F401
    from rscotr_b200.config import MODELS
    from rscotr_b200.mtl.data import build_dataloaders, build_datasets, build_multidataloader, load_data_cfg
    from rscotr_b200.mtl.data.datasets import CocoDataset, CustomDataset, PotsdamDataset
    from rscotr_b200.mtl.engine import StepEngine
    from rscotr_b200.mtl.runner import IterBasedRunner, MultiDatasetsEvalHook
    from tests.cpu_ops_shim import cpu_ops
    from tests.test_host_model import small_cfg

    def small(pipeline, **over):
        out = []
        for t in pipeline:
            t = dict(t)
            for k, v in over.items():
                typ, arg = k.split('__')
                if t['type'] == typ:
                    t[arg] = v
            if 'transforms' in t:
                t['transforms'] = small(t['transforms'], **over)
            out.append(t)
        return out

    base = {k: Config.fromfile(os.path.join(CFG_ROOT, f))._cfg_dict['data'] for k, f in
            dict(cls='resisc45.py', det='dior.py', seg='potsdam.py').items()}
    root = os.path.dirname(CFG_ROOT)
    cfg = Config(dict(seed=0, gpu_ids=[0], data=dict(
        resisc=dict(task='cls', config=os.path.join(CFG_ROOT, 'resisc45.py'), data=dict(
            samples_per_gpu=2, workers_per_gpu=0,
            train=dict(data_prefix=os.path.join(resisc, 'train'), pipeline=small(base['cls']['train']['pipeline'], RandomResizedCrop__size=64)),
            val=dict(data_prefix=os.path.join(resisc, 'val'), pipeline=small(base['cls']['val']['pipeline'], Resize__size=(64, 64))))),
        dior=dict(task='det', config=os.path.join(CFG_ROOT, 'dior.py'), data=dict(
            samples_per_gpu=1, workers_per_gpu=0,
            train=dict(ann_file=os.path.join(dior, 'coco_ann/DIOR_train_coco.json'), img_prefix=os.path.join(dior, 'JPEGImages-trainval'),
                       pipeline=small(base['det']['train']['pipeline'], Resize__img_scale=(160, 128))),
            val=dict(ann_file=os.path.join(dior, 'coco_ann/DIOR_val_coco.json'), img_prefix=os.path.join(dior, 'JPEGImages-trainval'),
                     pipeline=small(base['det']['val']['pipeline'], MultiScaleFlipAug__img_scale=(160, 128))))),
        potsdam=dict(task='seg', config=os.path.join(CFG_ROOT, 'potsdam.py'), data=dict(
            samples_per_gpu=1, workers_per_gpu=0,
            train=dict(data_root=potsdam, pipeline=small(base['seg']['train']['pipeline'], Resize__img_scale=(128, 128),
                                                         RandomCrop__crop_size=(128, 128), Pad__size=(128, 128))),
            val=dict(data_root=potsdam, pipeline=small(base['seg']['val']['pipeline'], MultiScaleFlipAug__img_scale=(128, 128))))))))
    assert root
    load_data_cfg(cfg)
    assert cfg.data.dior.config.data.test.classes == FX.DIOR_CLASSES            # untouched parts of the base survive the splice
    _seed(0)
    train_sets = build_datasets(cfg.data)                                        # files exist -> the real datasets
    assert [type(d) for d in train_sets.values()] == [CustomDataset, CocoDataset, PotsdamDataset]
    mdl = build_multidataloader(cfg, False, train_sets)
    torch.manual_seed(0)
    model = MODELS.build(small_cfg().model)
    model.init_weights()
    eng = StepEngine(model, dict(type='AdamW', lr=1e-4, weight_decay=1e-4), grad_clip=dict(max_norm=0.1, norm_type=2),
                     device='cpu', compute_dtype=torch.float32, use_graphs=False)
    runner = IterBasedRunner(eng, max_iters=3, work_dir=str(tmp_path), log_interval=0)
    from rscotr_b200.mtl.data.multi_eval_dataset import MultiEvalDatasets
    val_sets = {k: MultiEvalDatasets(v) for k, v in build_datasets(cfg.data, split='val').items()}   # as train_model does
    assert all(d.test_mode for d in val_sets.values()) and val_sets['dior'][0]['task'] == 'det'
    val_loaders = build_dataloaders(cfg, False, val_sets, train=False)
    hook = MultiDatasetsEvalHook(val_loaders, interval=3, by_epoch=False,
                                 save_best={'resisc.accuracy_top-1': 1, 'dior.bbox_mAP': 100, 'potsdam.mFscore': 100},
                                 cls=dict(metric='accuracy'), det=dict(metric='bbox', iou_thrs=[0.5], classwise=True),
                                 seg=dict(metric=['mFscore', 'mIoU'], pre_eval=True, classwise=True))
    runner.register_hook(hook)
    tasks = []
    orig = eng.train_iter
    eng.train_iter = lambda b: (tasks.append((b['task'], b['dataset_name'])), orig(b))[1]
    with cpu_ops():
        runner.run([mdl])
    assert tasks == [('cls', 'resisc'), ('det', 'dior'), ('seg', 'potsdam')]
    logs = dict(runner.log_buffer)
    for k in ('resisc.accuracy_top-1', 'resisc.accuracy_top-5', 'dior.bbox_mAP', 'dior.bbox_mAP_50', 'potsdam.mFscore',
              'potsdam.mIoU', 'potsdam.aAcc', 'potsdam.IoU.building'):
        assert k in logs, (k, sorted(logs))
    assert 0 <= logs['resisc.accuracy_top-1'] <= 100 and -1 <= logs['dior.bbox_mAP'] <= 1 and 0 <= logs['potsdam.mIoU'] <= 1
    assert hook.best_score is not None
    # single-image inference through the val pipelines (tools/inference_one_img.py)
    from rscotr_b200.mtl.apis import inference_one_img
    with cpu_ops():
        scores = inference_one_img(model, val_sets['resisc'], 'airport/airport_000.jpg')
        boxes = inference_one_img(model, val_sets['dior'], '00000.png')
        labels = inference_one_img(model, val_sets['potsdam'], 'tile_0.png')
    assert scores.shape == (45,) and len(boxes) == 20 and boxes[0].shape[1] == 5 and labels.shape == (96, 96)
    assert model.training                                                       # (the call restores the mode)
    best = [f for f in os.listdir(tmp_path) if f.startswith('best_')]
    assert best == ['best_resisc_accuracy_top-1_dior_bbox_mAP_potsdam_mFscore_iter_3.pth']


def test_worker_processes_and_seeding(resisc):
    """DataLoader workers (fork): the dynamically built RandAugment policy classes and the per-(rank, worker) seeding work
    in worker processes; the same loader seed reproduces the same first batch."""
    ds = build_dataset(_cls_cfg(resisc), 'cls')

    def first(seed):
        loader = build_dataloader(ds, samples_per_gpu=3, workers_per_gpu=2, dist=False, seed=seed, pin_memory=False)
        it = iter(loader)
        b = next(it)
        del it
        return b
    a, b, c = first(5), first(5), first(6)
    assert a['img'].shape == (3, 3, 224, 224)
    assert torch.equal(a['gt_label'], b['gt_label']) and torch.equal(a['img'], b['img'])
    assert not torch.equal(a['img'], c['img'])


def test_distributed_sampler_partitions_and_reshuffles():
    from rscotr_b200.mtl.data.loader import DistributedSampler
    data = list(range(10))
    parts = [list(DistributedSampler(data, num_replicas=4, rank=r, shuffle=True, seed=7)) for r in range(4)]
    assert all(len(p) == 3 for p in parts)                                   # rounded up to 12 = 4 x 3
    flat = sorted(i for p in parts for i in p)
    assert set(flat) == set(range(10)) and len(flat) == 12                   # every sample seen, two repeated
    s = DistributedSampler(data, num_replicas=2, rank=0, shuffle=True, seed=7)
    first = list(s)
    s.set_epoch(1)
    assert list(s) != first and sorted(list(s) + list(_other(s))) == list(range(10))
    assert list(DistributedSampler(data, num_replicas=2, rank=1, shuffle=False)) == [1, 3, 5, 7, 9]
    assert len(list(DistributedSampler(data, num_replicas=4, rank=3, shuffle=False, round_up=False))) == 2


def _other(s):
    from rscotr_b200.mtl.data.loader import DistributedSampler
    o = DistributedSampler(list(range(s.n)), num_replicas=2, rank=1, shuffle=True, seed=s.seed)
    o.set_epoch(s.epoch)
    return o


def test_geometric_transform_properties():
    """property checks with hypothesis: horizontal flip of boxes is an involution and keeps widths; keep-ratio resize maps the
    boxes with the image (same relative position), never leaves the image, and respects the (long, short) edge limits."""
    hyp = pytest.importorskip('hypothesis')
    st = pytest.importorskip('hypothesis.strategies')

    @hyp.settings(max_examples=60, deadline=None, derandomize=True, database=None)
    @hyp.given(h=st.integers(20, 200), w=st.integers(20, 200), seed=st.integers(0, 10 ** 6),
               long_edge=st.integers(64, 400), short_edge=st.integers(32, 300))
    def run(h, w, seed, long_edge, short_edge):
        hyp.assume(long_edge >= short_edge)
        rng = np.random.default_rng(seed)
        n = int(rng.integers(1, 5))
        x1, y1 = rng.uniform(0, w - 2, n), rng.uniform(0, h - 2, n)
        boxes = np.stack([x1, y1, x1 + rng.uniform(1, w - x1), y1 + rng.uniform(1, h - y1)], 1).astype(np.float32)
        img = np.zeros((h, w, 3), dtype=np.uint8)
        base = dict(img=img, img_shape=img.shape, gt_bboxes=boxes.copy(), bbox_fields=['gt_bboxes'], img_fields=['img'])
        flip = T.RandomFlip(flip_ratio=None, task='det')
        once = flip(dict(base, flip=True, flip_direction='horizontal'))
        twice = flip(dict(once, flip=True))
        assert np.allclose(twice['gt_bboxes'], boxes, atol=1e-4)
        assert np.allclose(once['gt_bboxes'][:, 2] - once['gt_bboxes'][:, 0], boxes[:, 2] - boxes[:, 0], atol=1e-4)
        r = T.Resize(img_scale=(long_edge, short_edge), keep_ratio=True, task='det')(dict(base, gt_bboxes=boxes.copy()))
        nh, nw = r['img'].shape[:2]
        assert max(nh, nw) <= long_edge + 1 and min(nh, nw) <= short_edge + 1
        assert abs(nw / w - nh / h) < 0.06 * max(nw / w, nh / h) + 2.0 / min(h, w)       # aspect kept up to pixel rounding
        b = r['gt_bboxes']
        assert (b[:, 0::2] >= 0).all() and (b[:, 0::2] <= nw).all() and (b[:, 1::2] >= 0).all() and (b[:, 1::2] <= nh).all()
        assert np.allclose(b[:, 0] / nw, boxes[:, 0] / w, atol=1e-5) and np.allclose(b[:, 3] / nh, boxes[:, 3] / h, atol=1e-5)
    run()


def test_image_ops_against_independent_implementations():
    """mmcv-style image ops (cv2) against torch / torchvision: normalisation (incl. BGR -> RGB), bilinear up-scaling with the
    half-pixel convention, zero padding, horizontal flip."""
    import torch.nn.functional as F
    tvf = pytest.importorskip('torchvision.transforms.functional')
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (23, 31, 3), dtype=np.uint8)                       # BGR, HWC
    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    ours = T.imnormalize(img, np.array(mean, dtype=np.float32), np.array(std, dtype=np.float32), to_rgb=True)
    rgb = torch.from_numpy(img[..., ::-1].copy()).permute(2, 0, 1).float()
    want = tvf.normalize(rgb, mean, std).permute(1, 2, 0).numpy()
    assert np.allclose(ours, want, atol=1e-5)
    up = T.imresize(img, (62, 46), 'bilinear')                                   # (w, h) = exactly 2x
    ref = F.interpolate(torch.from_numpy(img).permute(2, 0, 1)[None].float(), size=(46, 62), mode='bilinear', align_corners=False)
    assert np.abs(up.astype(np.float32) - ref[0].permute(1, 2, 0).numpy()).max() <= 1.0      # cv2 rounds to uint8 (fixed point)
    pad = T.impad_to_multiple(img, 32)
    assert pad.shape == (32, 32, 3) and (pad[23:] == 0).all() and (pad[:, 31:] == 0).all() and np.array_equal(pad[:23, :31], img)
    assert np.array_equal(T.imflip(img), tvf.hflip(torch.from_numpy(img).permute(2, 0, 1)).permute(1, 2, 0).numpy())


def test_random_resized_crop_distribution_matches_torchvision():
    """same sampling scheme as torchvision.transforms.RandomResizedCrop.get_params (area ~ U(0.08, 1), log-uniform aspect in
    [3/4, 4/3], 10 attempts, centre fallback): compare the first two moments of area fraction and log aspect over many draws."""
    tvt = pytest.importorskip('torchvision.transforms')
    img = np.zeros((120, 200, 3), dtype=np.uint8)
    rrc = T.RandomResizedCrop(size=32)
    random.seed(0)
    torch.manual_seed(0)
    ours = np.array([rrc._params(img) for _ in range(4000)], dtype=np.float64)              # (y, x, h, w)
    ref = np.array([tvt.RandomResizedCrop.get_params(torch.zeros(3, 120, 200), (0.08, 1.0), (3 / 4, 4 / 3)) for _ in range(4000)],
                   dtype=np.float64)                                                        # (i, j, h, w)
    for a in (ours, ref):
        assert (a[:, 0] >= 0).all() and (a[:, 1] >= 0).all() and (a[:, 0] + a[:, 2] <= 120).all() and (a[:, 1] + a[:, 3] <= 200).all()
    area_o, area_r = ours[:, 2] * ours[:, 3] / (120 * 200), ref[:, 2] * ref[:, 3] / (120 * 200)
    asp_o, asp_r = np.log(ours[:, 3] / ours[:, 2]), np.log(ref[:, 3] / ref[:, 2])
    assert abs(area_o.mean() - area_r.mean()) < 0.02 and abs(area_o.std() - area_r.std()) < 0.02
    assert abs(asp_o.mean() - asp_r.mean()) < 0.02 and abs(asp_o.std() - asp_r.std()) < 0.02
