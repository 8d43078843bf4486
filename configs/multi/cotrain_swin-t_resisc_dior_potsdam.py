# The co-training run on the real datasets (NWPU-RESISC45 + DIOR + Potsdam under ./data, the directory
# layout of the reference README): same model / schedule as cotrain_swin-t_800.py, dataset settings from
# configs/datasets/{resisc45,dior,potsdam}.py, mixup / cutmix on as in the reference.
_base_ = './cotrain_swin-t_800.py'
model = dict(train_cfg=dict(cls=dict(augments=[
    dict(type='BatchMixup', alpha=0.8, num_classes=45, prob=0.5),
    dict(type='BatchCutMix', alpha=1.0, num_classes=45, prob=0.5)])))
data = dict(
    _delete_=True,
    resisc=dict(task='cls', config='configs/datasets/resisc45.py', data=dict(samples_per_gpu=16, workers_per_gpu=8)),
    dior=dict(task='det', config='configs/datasets/dior.py', data=dict(samples_per_gpu=1, workers_per_gpu=2)),
    potsdam=dict(task='seg', config='configs/datasets/potsdam.py', data=dict(samples_per_gpu=2, workers_per_gpu=4)))
synthetic = False      # insist on the files: a missing dataset is an error, not a synthetic stand-in
