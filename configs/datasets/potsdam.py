# ISPRS Potsdam IRRG tiles (512 crops, 5 classes + ignored clutter) through PotsdamDataset.
# Same dataset / pipeline settings as the reference's configs/_base_/seg/potsdam_IRRG_all.py.
dataset_type = 'PotsdamDataset'
data_root = 'data/potsdam'
img_norm_cfg = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)
crop_size = (512, 512)

train_pipeline = [
    dict(type='LoadImageFromFile'),
    dict(type='LoadAnnotations', reduce_zero_label=True),
    dict(type='Resize', img_scale=(512, 512), ratio_range=(0.5, 2.0)),
    dict(type='RandomCrop', crop_size=crop_size, cat_max_ratio=0.75),
    dict(type='RandomFlip', prob=0.5),
    dict(type='PhotoMetricDistortion'),
    dict(type='Normalize', **img_norm_cfg),
    dict(type='Pad', size=crop_size, pad_val=0, seg_pad_val=5),      # clutter (5) is the ignore index
    dict(type='DefaultFormatBundle'),
    dict(type='Collect', keys=['img', 'gt_semantic_seg'])]


def _tta(**scale):
    steps = [dict(type='Resize', keep_ratio=True), dict(type='RandomFlip'), dict(type='Normalize', **img_norm_cfg),
             dict(type='ImageToTensor', keys=['img']), dict(type='Collect', keys=['img'])]
    return [dict(type='LoadImageFromFile'), dict(type='MultiScaleFlipAug', flip=False, transforms=steps, **scale)]


val_pipeline = _tta(img_scale=(512, 512))
test_pipeline = _tta(img_scale=None, img_ratios=[1.0])
dataset_kwargs = dict(type=dataset_type, data_root=data_root, ignore_index=5)

data = dict(
    samples_per_gpu=8,
    workers_per_gpu=8,
    train=dict(img_dir='img_IRRG/train', ann_dir='ann_all/train', pipeline=train_pipeline, **dataset_kwargs),
    val=dict(img_dir='img_IRRG/val', ann_dir='ann_all/val', pipeline=val_pipeline, **dataset_kwargs),
    test=dict(img_dir='img_IRRG/val', ann_dir='ann_all/val', pipeline=test_pipeline, **dataset_kwargs))
