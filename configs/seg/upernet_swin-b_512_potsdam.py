# Segmentation-only Swin-B + UPerNet on 512x512 Potsdam tiles (BASELINE.json configs[4]).
# The reference repo ships no UPerNet config; this is the standard mmseg Swin-B/UPerNet head layout
# (PPM scales 1/2/3/6, 512 channels, FCN auxiliary head on stage 3 with weight 0.4) on the reference's
# Potsdam dataset settings (configs/datasets/potsdam.py; 6 classes, clutter ignored).
norm_cfg = dict(type='BN', requires_grad=True)       # per-rank statistics: gradients are the only collective
model = dict(
    type='EncoderDecoder',
    backbone=dict(type='SwinTransformer', pretrain_img_size=224, embed_dims=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32],
                  window_size=7, mlp_ratio=4, qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.3,
                  patch_norm=True, out_indices=(0, 1, 2, 3), with_cp=False, convert_weights=True),
    decode_head=dict(type='UPerHead', in_channels=[128, 256, 512, 1024], in_index=[0, 1, 2, 3], pool_scales=(1, 2, 3, 6),
                     channels=512, dropout_ratio=0.1, num_classes=6, norm_cfg=norm_cfg, align_corners=False, ignore_index=5,
                     loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0)),
    auxiliary_head=dict(type='FCNHead', in_channels=512, in_index=2, channels=256, num_convs=1, concat_input=False,
                        dropout_ratio=0.1, num_classes=6, norm_cfg=norm_cfg, align_corners=False, ignore_index=5,
                        loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=0.4)),
    train_cfg=dict(),
    test_cfg=dict(mode='whole'))

data = dict(potsdam=dict(task='seg', config='configs/datasets/potsdam.py', data=dict(samples_per_gpu=8, workers_per_gpu=8)))
synthetic = dict(img_size=(512, 512), seg=dict(num_classes=5))
strategy = dict(type='round_robin')

optimizer = dict(type='AdamW', lr=6e-5, betas=(0.9, 0.999), weight_decay=0.01,
                 paramwise_cfg=dict(custom_keys={'absolute_pos_embed': dict(decay_mult=0.), 'relative_position_bias_table': dict(decay_mult=0.),
                                                 'norm': dict(decay_mult=0.)}))
optimizer_config = dict()
lr_config = dict(policy='poly', warmup='linear', warmup_iters=1500, warmup_ratio=1e-6, power=1.0, min_lr=0.0, by_epoch=False)
runner = dict(type='IterBasedRunner', max_iters=80000)
checkpoint_config = dict(interval=8000)
evaluation = dict(interval=8000, save_best={'potsdam.mFscore': 100}, seg=dict(metric=['mFscore', 'mIoU'], pre_eval=True, classwise=True))
dist_params = dict(backend='nccl')
log_config = dict(interval=50)
