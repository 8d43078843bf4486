# runtime defaults shared by the co-training configs (mmcv-style python config)
checkpoint_config = dict(interval=5000)
dist_params = dict(backend='nccl')
log_level = 'INFO'
load_from = None
resume_from = None
workflow = [('train', 1)]
log_config = dict(interval=50, hooks=[dict(type='TextLoggerHook')])
