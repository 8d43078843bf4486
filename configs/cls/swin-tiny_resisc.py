# Single-task scene classification, Swin-T on NWPU-RESISC45 (BASELINE.json configs[0]; the reference's
# configs/cls/swin-tiny_1xb16_resisc.py with its four _base_ files folded in).  The reference's file also
# loads unmodified (tests/test_host_units.py::test_single_task_classifier).
_base_ = '../datasets/resisc45.py'
model = dict(
    type='ImageClassifier',
    backbone=dict(type='SwinTransformer', arch='tiny', img_size=224, drop_path_rate=0.2),
    neck=dict(type='GlobalAveragePooling'),
    head=dict(type='LinearClsHead', num_classes=45, in_channels=768, init_cfg=None,
              loss=dict(type='LabelSmoothLoss', label_smooth_val=0.1, mode='original'), cal_acc=False),
    init_cfg=[dict(type='TruncNormal', layer='Linear', std=0.02, bias=0.), dict(type='Constant', layer='LayerNorm', val=1., bias=0.)],
    train_cfg=dict())
optimizer = dict(type='AdamW', lr=2e-4, weight_decay=0.0001, eps=1e-8, betas=(0.9, 0.999),
                 paramwise_cfg=dict(norm_decay_mult=0.0, bias_decay_mult=0.0,
                                    custom_keys={'.absolute_pos_embed': dict(decay_mult=0.0),
                                                 '.relative_position_bias_table': dict(decay_mult=0.0)}))
optimizer_config = dict(grad_clip=dict(max_norm=0.1, norm_type=2))
lr_config = dict(policy='step', step=[150])
runner = dict(type='EpochBasedRunner', max_epochs=200)
checkpoint_config = dict(interval=50)
evaluation = dict(interval=1, metric='accuracy')
log_config = dict(interval=100)
dist_params = dict(backend='nccl')
synthetic = dict(img_size=(256, 256))
