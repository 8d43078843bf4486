# DIOR (20 classes) through the COCO-json reader of rscotr_b200/mtl/data/datasets.py.
# Same dataset / pipeline settings as the reference's configs/_base_/det/dior.py
# (tests/test_data_plane.py compares the two `data` dicts when the reference tree is mounted).
dataset_type = 'CocoDataset'
data_root = 'data/DIOR/'
classes = ('airplane', 'airport', 'baseballfield', 'basketballcourt', 'bridge', 'chimney', 'dam', 'Expressway-Service-area',
           'Expressway-toll-station', 'golffield', 'groundtrackfield', 'harbor', 'overpass', 'ship', 'stadium', 'storagetank',
           'tenniscourt', 'trainstation', 'vehicle', 'windmill')
img_norm_cfg = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)
_scale = (1333, 800)

train_pipeline = [
    dict(type='LoadImageFromFile'),
    dict(type='LoadAnnotations', with_bbox=True),
    dict(type='Resize', img_scale=_scale, keep_ratio=True),
    dict(type='RandomFlip', flip_ratio=0.5),
    dict(type='Normalize', **img_norm_cfg),
    dict(type='Pad', size_divisor=32),
    dict(type='DefaultFormatBundle'),
    dict(type='Collect', keys=['img', 'gt_bboxes', 'gt_labels'])]

_test_steps = [
    dict(type='Resize', keep_ratio=True),
    dict(type='RandomFlip'),
    dict(type='Normalize', **img_norm_cfg),
    dict(type='Pad', size_divisor=32),
    dict(type='ImageToTensor', keys=['img']),
    dict(type='Collect', keys=['img'])]
test_pipeline = [
    dict(type='LoadImageFromFile'),
    dict(type='MultiScaleFlipAug', img_scale=_scale, flip=False, transforms=_test_steps)]


def _split(ann, prefix, pipeline):
    return dict(type=dataset_type, ann_file=data_root + 'coco_ann/DIOR_%s_coco.json' % ann, img_prefix=data_root + prefix,
                pipeline=pipeline, classes=classes)


data = dict(
    samples_per_gpu=1,
    workers_per_gpu=2,
    train=_split('train', 'JPEGImages-trainval', train_pipeline),
    val=_split('val', 'JPEGImages-trainval/', test_pipeline),
    test=_split('test', 'JPEGImages-test/', test_pipeline))
evaluation = dict(interval=1, metric='bbox', iou_thrs=[0.5], save_best='bbox_mAP_50', classwise=True)
