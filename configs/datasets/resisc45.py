# NWPU-RESISC45 class folders (224 crops, RandAugment + RandomErasing) through CustomDataset.
# Same dataset / pipeline settings as the reference's configs/_base_/cls/resisc_swin_224.py + rand_aug.py.
dataset_type = 'CustomDataset'
img_norm_cfg = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)
_bgr_mean, _bgr_std = img_norm_cfg['mean'][::-1], img_norm_cfg['std'][::-1]


def _mag(type, key, lo, hi, **kw):
    return dict(type=type, magnitude_key=key, magnitude_range=(lo, hi), **kw)


# timm's "increasing" RandAugment set: every magnitude grows with the level
rand_increasing_policies = [
    dict(type='AutoContrast'), dict(type='Equalize'), dict(type='Invert'),
    _mag('Rotate', 'angle', 0, 30), _mag('Posterize', 'bits', 4, 0), _mag('Solarize', 'thr', 256, 0),
    _mag('SolarizeAdd', 'magnitude', 0, 110), _mag('ColorTransform', 'magnitude', 0, 0.9),
    _mag('Contrast', 'magnitude', 0, 0.9), _mag('Brightness', 'magnitude', 0, 0.9), _mag('Sharpness', 'magnitude', 0, 0.9),
    _mag('Shear', 'magnitude', 0, 0.3, direction='horizontal'), _mag('Shear', 'magnitude', 0, 0.3, direction='vertical'),
    _mag('Translate', 'magnitude', 0, 0.45, direction='horizontal'), _mag('Translate', 'magnitude', 0, 0.45, direction='vertical')]

train_pipeline = [
    dict(type='LoadImageFromFile'),
    dict(type='RandomResizedCrop', size=224, backend='pillow', interpolation='bicubic'),
    dict(type='RandomFlip', flip_prob=0.5, direction='horizontal'),
    dict(type='RandAugment', policies=rand_increasing_policies, num_policies=2, total_level=10, magnitude_level=9,
         magnitude_std=0.5, hparams=dict(pad_val=[round(x) for x in _bgr_mean], interpolation='bicubic')),
    dict(type='RandomErasing', erase_prob=0.25, mode='rand', min_area_ratio=0.02, max_area_ratio=1 / 3,
         fill_color=_bgr_mean, fill_std=_bgr_std),
    dict(type='Normalize', **img_norm_cfg),
    dict(type='ImageToTensor', keys=['img']),
    dict(type='ToTensor', keys=['gt_label']),
    dict(type='Collect', keys=['img', 'gt_label'])]

test_pipeline = [
    dict(type='LoadImageFromFile'),
    dict(type='Resize', size=(224, 224), backend='pillow', interpolation='bicubic'),
    dict(type='Normalize', **img_norm_cfg),
    dict(type='ImageToTensor', keys=['img']),
    dict(type='Collect', keys=['img'])]

data = dict(
    samples_per_gpu=16,
    workers_per_gpu=8,
    train=dict(type=dataset_type, data_prefix='data/NWPU-RESISC45/train', pipeline=train_pipeline),
    val=dict(type=dataset_type, data_prefix='data/NWPU-RESISC45/val', pipeline=test_pipeline),
    test=dict(type=dataset_type, data_prefix='data/NWPU-RESISC45/test', pipeline=test_pipeline))
evaluation = dict(interval=10, metric='accuracy')
