# DIOR (20 classes) through the COCO-json reader of rscotr_b200/mtl/data/datasets.py.  The resulting `data` /
# `evaluation` dicts are checked (tests/test_data_plane.py) to equal the reference's DIOR dataset settings.
NORM = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)
ROOT = 'data/DIOR/'
SCALE = (1333, 800)          # long edge, short edge
NAMES = ('airplane airport baseballfield basketballcourt bridge chimney dam Expressway-Service-area Expressway-toll-station '
         'golffield groundtrackfield harbor overpass ship stadium storagetank tenniscourt trainstation vehicle windmill').split()


def step(kind, **kw):
    return dict(type=kind, **kw)


def pipeline(train):
    tail = [step('Normalize', **NORM), step('Pad', size_divisor=32)]
    if train:
        return [step('LoadImageFromFile'), step('LoadAnnotations', with_bbox=True), step('Resize', img_scale=SCALE, keep_ratio=True),
                step('RandomFlip', flip_ratio=0.5)] + tail + [step('DefaultFormatBundle'),
                                                              step('Collect', keys=['img', 'gt_bboxes', 'gt_labels'])]
    inner = [step('Resize', keep_ratio=True), step('RandomFlip')] + tail + [step('ImageToTensor', keys=['img']),
                                                                            step('Collect', keys=['img'])]
    return [step('LoadImageFromFile'), step('MultiScaleFlipAug', img_scale=SCALE, flip=False, transforms=inner)]


def split(name, images, train=False):
    return dict(type='CocoDataset', ann_file='%scoco_ann/DIOR_%s_coco.json' % (ROOT, name), img_prefix=ROOT + images,
                pipeline=pipeline(train), classes=tuple(NAMES))


data = dict(samples_per_gpu=1, workers_per_gpu=2, train=split('train', 'JPEGImages-trainval', True),
            val=split('val', 'JPEGImages-trainval/'), test=split('test', 'JPEGImages-test/'))
evaluation = dict(interval=1, metric='bbox', iou_thrs=[0.5], save_best='bbox_mAP_50', classwise=True)
