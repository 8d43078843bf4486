# ISPRS Potsdam IRRG tiles (512 crops, 5 classes + ignored clutter) through PotsdamDataset.  The resulting `data` dict is
# checked (tests/test_data_plane.py) to equal the reference's Potsdam dataset settings.
NORM = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)
CROP = (512, 512)
CLUTTER = 5                  # label of the ignored class after reduce_zero_label; also pads the label map


def step(kind, **kw):
    return dict(type=kind, **kw)


train_steps = [
    step('LoadImageFromFile'), step('LoadAnnotations', reduce_zero_label=True),
    step('Resize', img_scale=CROP, ratio_range=(0.5, 2.0)), step('RandomCrop', crop_size=CROP, cat_max_ratio=0.75),
    step('RandomFlip', prob=0.5), step('PhotoMetricDistortion'), step('Normalize', **NORM),
    step('Pad', size=CROP, pad_val=0, seg_pad_val=CLUTTER), step('DefaultFormatBundle'),
    step('Collect', keys=['img', 'gt_semantic_seg'])]


def whole_image(**scale):
    inner = [step('Resize', keep_ratio=True), step('RandomFlip'), step('Normalize', **NORM), step('ImageToTensor', keys=['img']),
             step('Collect', keys=['img'])]
    return [step('LoadImageFromFile'), step('MultiScaleFlipAug', flip=False, transforms=inner, **scale)]


def split(name, steps):
    return dict(img_dir='img_IRRG/' + name, ann_dir='ann_all/' + name, pipeline=steps, type='PotsdamDataset',
                data_root='data/potsdam', ignore_index=CLUTTER)


data = dict(samples_per_gpu=8, workers_per_gpu=8, train=split('train', train_steps),
            val=split('val', whole_image(img_scale=CROP)), test=split('val', whole_image(img_scale=None, img_ratios=[1.0])))
