# Swin-T co-training (cls + det + seg, round robin), synthetic 3x800x800 inputs.
# Same schema and hyper-parameters as the reference's
# configs/multi/MTL_slvlcls_swin-t-p4-w7_1x1_resisc&dior&potsdam.py (which also loads
# unmodified, see tests/test_host_model.py); datasets are synthetic.
_base_ = 'runtime.py'

E = 256            # transformer width
FFN_CH = 2048
relu = dict(type='ReLU', inplace=True)
ffn = dict(type='FFN', feedforward_channels=FFN_CH, num_fcs=2, ffn_drop=0.0, act_cfg=relu)
msda = dict(type='MultiScaleDeformableAttention', embed_dims=E, num_levels=4, dropout=0.0)
sine_pe = dict(type='SinePositionalEncoding', num_feats=E // 2, normalize=True)

model = dict(
    type='MTL',
    backbone=dict(type='SwinTransformer', embed_dims=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=7,
                  mlp_ratio=4, qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2,
                  patch_norm=True, out_indices=(0, 1, 2, 3), with_cp=False, convert_weights=True, init_cfg=None),
    neck=dict(type='ChannelMapper', in_channels=[192, 384, 768], kernel_size=1, out_channels=E, act_cfg=None,
              norm_cfg=dict(type='GN', num_groups=32), num_outs=4),
    shared_encoder=dict(type='DetrTransformerEncoder', num_layers=6,
                        transformerlayers=dict(type='BaseTransformerLayer', attn_cfgs=msda, ffn_cfgs=ffn,
                                               operation_order=('self_attn', 'norm', 'ffn', 'norm'))),
    cls_head=dict(type='SlvlClsHead', num_classes=45, in_channels=768,
                  loss=dict(type='LabelSmoothLoss', label_smooth_val=0.1, mode='original'), cal_acc=False),
    bbox_head=dict(
        type='DINOHead', num_query=600, num_classes=20, num_feature_levels=4, in_channels=2048,
        sync_cls_avg_factor=True, as_two_stage=True, with_box_refine=True,
        dn_cfg=dict(type='CdnQueryGenerator', noise_scale=dict(label=0.5, box=1.0),
                    group_cfg=dict(dynamic=True, num_groups=None, num_dn_queries=100)),
        transformer=dict(
            type='DinoTransformer',
            decoder=dict(type='DinoTransformerDecoder', num_layers=6, return_intermediate=True,
                         transformerlayers=dict(
                             type='BaseTransformerLayer',
                             attn_cfgs=[dict(type='MultiheadAttention', embed_dims=E, num_heads=8, dropout=0.0), msda],
                             ffn_cfgs=ffn,
                             operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')))),
        positional_encoding=dict(type='SinePositionalEncoding', num_feats=E // 2, temperature=20, normalize=True),
        loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0),
        loss_bbox=dict(type='L1Loss', loss_weight=5.0),
        loss_iou=dict(type='GIoULoss', loss_weight=2.0)),
    seg_head=dict(
        type='Mask2FormerHead', in_channels=[96, 192, 384, 768], scheme=2, feat_channels=E, out_channels=E,
        num_classes=5, num_queries=100, num_transformer_feat_level=4, align_corners=False,
        pixel_decoder=dict(type='MlvlSegPixelDecoder', num_outs=4, norm_cfg=dict(type='GN', num_groups=32),
                           act_cfg=dict(type='ReLU'), positional_encoding=dict(temperature=10000, **sine_pe)),
        positional_encoding=dict(temperature=10000, **sine_pe),
        transformer_decoder=dict(
            type='DetrTransformerDecoder', num_layers=9, return_intermediate=True,
            transformerlayers=dict(
                type='BaseTransformerLayer',
                attn_cfgs=dict(type='MultiheadAttention', embed_dims=E, num_heads=8, attn_drop=0.0, proj_drop=0.0,
                               dropout_layer=None, batch_first=False),
                ffn_cfgs=ffn,
                operation_order=('cross_attn', 'norm', 'self_attn', 'norm', 'ffn', 'norm'))),
        loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0)),
    task_weight=dict(cls=1, det=1, seg=0.1),
    train_cfg=dict(
        cls=dict(augments=None),     # BatchMixup / BatchCutMix off for timing parity
        det=dict(assigner=dict(type='HungarianAssigner', cls_cost=dict(type='FocalLossCost', weight=2.0),
                               reg_cost=dict(type='BBoxL1Cost', weight=5.0, box_format='xywh'),
                               iou_cost=dict(type='IoUCost', iou_mode='giou', weight=2.0))),
        seg=dict()),
    test_cfg=dict(cls=dict(), det=dict(max_per_img=300), seg=dict(mode='whole')))

data = dict(
    resisc=dict(task='cls', config='configs/datasets/synthetic_cls.py', data=dict(samples_per_gpu=16)),
    dior=dict(task='det', config='configs/datasets/synthetic_det.py', data=dict(samples_per_gpu=1)),
    potsdam=dict(task='seg', config='configs/datasets/synthetic_seg.py', data=dict(samples_per_gpu=2)))
synthetic = dict(img_size=(800, 800), det=dict(num_boxes=8))
strategy = dict(type='round_robin')

optimizer = dict(type='AdamW', lr=5e-5, weight_decay=0.0001,
                 paramwise_cfg=dict(custom_keys={'backbone': dict(lr_mult=0.1), 'query_embed': dict(decay_mult=0.0),
                                                 'query_feat': dict(decay_mult=0.0),
                                                 'level_embed': dict(decay_mult=0.0)}))
optimizer_config = dict(grad_clip=dict(max_norm=0.1, norm_type=2))
lr_config = dict(policy='step', step=[240000, 285000])
runner = dict(type='IterBasedRunner', max_iters=300000)
evaluation = dict(
    interval=15000,
    save_best={'resisc.accuracy_top-1': 1, 'dior.bbox_mAP': 100, 'potsdam.mFscore': 100},
    cls=dict(metric='accuracy'),
    det=dict(metric='bbox', iou_thrs=[0.5], classwise=True),
    seg=dict(metric=['mFscore', 'mIoU'], pre_eval=True, classwise=True))
checkpoint_config = dict(interval=100000)
custom_imports = dict(imports='models.multi', allow_failed_imports=False)
