# Detection-only run of the co-training model (BASELINE.json configs[3]: Swin-S, DIOR-shape 3x800x800,
# global batch 16 = 2 per GPU on 8 GPUs): the same MTL schema with the cls / seg heads removed, one dataset,
# the `constant` iteration strategy.  Swin-S = depths (2, 2, 18, 2) at the Swin-T widths.
_base_ = './cotrain_swin-t_800.py'
model = dict(
    backbone=dict(depths=[2, 2, 18, 2], drop_path_rate=0.3),
    cls_head=None,
    seg_head=None)
data = dict(
    _delete_=True,
    dior=dict(task='det', config='configs/datasets/synthetic_det.py', data=dict(samples_per_gpu=2)))
strategy = dict(type='constant', idx=0)
evaluation = dict(_delete_=True, interval=15000, save_best={'dior.bbox_mAP': 100},
                  det=dict(metric='bbox', iou_thrs=[0.5], classwise=True))
