"""train_model with the reference's signature (mtl/apis/train.py:24-120)."""
import logging
import os

import torch
import torch.distributed as dist

from ..data import build_datasets, build_dataloaders, build_multidataloader
from ..data.multi_eval_dataset import MultiEvalDatasets
from ..data.synthetic import SyntheticDataset
from ..engine import StepEngine
from ..runner import CheckpointHook, IterBasedRunner, MultiDatasetsEvalHook
from ..utils.checkpoint import find_latest_checkpoint


def train_model(model, datasets, cfg, distributed=False, validate=False, timestamp=None, meta=None, _skip_eval_tasks=()):
    logger = logging.getLogger('rscotr_b200')
    # the process group must exist BEFORE the loaders are built: their samplers read (rank, world) at construction
    if distributed and not dist.is_initialized():
        dist.init_process_group(cfg.get('dist_params', {}).get('backend', 'nccl'))
    data_loader = [build_multidataloader(cfg, distributed, datasets)]
    device = cfg.get('device', 'cuda')
    if distributed and str(device) == 'cuda':
        device = 'cuda:%d' % int(os.environ.get('LOCAL_RANK', 0))
    if str(device).startswith('cuda'):
        torch.cuda.set_device(device)
    optimizer_config = cfg.get('optimizer_config', {}) or {}
    if cfg.get('fp16', None) is not None:
        # the reference's Fp16OptimizerHook (fp16 + dynamic loss scaling); on B200 the mixed-precision type is bf16, which has
        # fp32's exponent range and needs no loss scaling
        logger.warning('cfg.fp16 is set: running bf16 mixed precision (no loss scaling needed) instead of fp16')
    engine = StepEngine(model, cfg.optimizer, grad_clip=optimizer_config.get('grad_clip'), device=device,
                        compute_dtype=cfg.get('compute_dtype', torch.bfloat16),
                        lr_config=cfg.get('lr_config'))
    engine.max_iters = (cfg.get('lr_config') or {}).get('max_iters', cfg.runner['max_iters'])
    runner = IterBasedRunner(engine, cfg.runner['max_iters'], work_dir=cfg.get('work_dir'), logger=logger, meta=meta,
                             log_interval=cfg.get('log_config', {}).get('interval', 50))
    runner.timestamp = timestamp
    if cfg.get('checkpoint_config'):
        ck = dict(cfg.checkpoint_config)
        ck.setdefault('by_epoch', False)
        runner.register_hook(CheckpointHook(**ck), priority='NORMAL')
    if validate:
        val_dataset = build_datasets(cfg.data, split='val', synthetic=cfg.get('synthetic'))
        val_dataset = {name: (ds if isinstance(ds, SyntheticDataset) else MultiEvalDatasets(ds)) for name, ds in val_dataset.items()}
        val_dataloader = build_dataloaders(cfg, distributed, val_dataset, train=False)
        val_dataloader = {n: l for n, l in val_dataloader.items() if l.dataset.task not in _skip_eval_tasks}
        eval_cfg = dict(cfg.get('evaluation', {}))
        eval_cfg['by_epoch'] = cfg.runner['type'] != 'IterBasedRunner'
        runner.register_hook(MultiDatasetsEvalHook(val_dataloader, **eval_cfg), priority='LOW')
    resume_from = cfg.get('resume_from')
    if resume_from is None and cfg.get('auto_resume') and cfg.get('work_dir'):
        resume_from = find_latest_checkpoint(cfg.work_dir)
    if resume_from:
        runner.resume(resume_from)
    elif cfg.get('load_from'):
        runner.load_checkpoint(cfg.load_from)
    runner.run(data_loader, cfg.get('workflow', [('train', 1)]))
    return runner


def train_model_without_det_eval(model, datasets, cfg, distributed=False, validate=False, timestamp=None, meta=None):
    """reference mtl/apis/train.py:123-222: train_model, with the detection datasets left out of the validation loaders."""
    return train_model(model, datasets, cfg, distributed, validate, timestamp, meta, _skip_eval_tasks=('det',))
