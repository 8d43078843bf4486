"""test_model: the body of the reference's tools/test.py:146-222 as a function:
splice the dataset configs, build the test datasets / loaders of the selected tasks, load the checkpoint,
run every loader through the model, evaluate each dataset with `cfg.evaluation[task]`."""
import torch

from ...config import MODELS
from ..data import build_dataloaders, build_datasets, load_data_cfg
from ..engine.test import multi_gpu_test, single_gpu_test
from ..utils.checkpoint import load_checkpoint

_HOOK_ARGS = ('interval', 'tmpdir', 'start', 'gpu_collect', 'save_best', 'rule', 'dynamic_intervals')


def test_model(cfg, checkpoint=None, tasks=('cls', 'det', 'seg'), split='test', distributed=False, device=None,
               synthetic=None, test_outputs=None, eval_override=None):
    """-> (metrics {dataset: {metric: value}}, outputs {dataset: results})."""
    if not all('config' in v and 'task' in v and hasattr(v['config'], 'get') for v in cfg.data.values()):
        load_data_cfg(cfg)
    datasets = {n: d for n, d in build_datasets(cfg.data, split=split, synthetic=synthetic).items() if d.task in tasks}
    loaders = build_dataloaders(cfg, distributed, datasets, train=False)
    model_cfg = cfg.model
    model_cfg['train_cfg'] = {k: {} for k in (model_cfg.get('train_cfg') or {}).keys()}
    model = MODELS.build(model_cfg)
    meta = {}
    if checkpoint is not None:
        meta = load_checkpoint(model, checkpoint)[0].get('meta', {})
    model.CLASSES = meta.get('CLASSES') or {n: d.CLASSES for n, d in datasets.items()}
    device = device or ('cuda' if torch.cuda.is_available() else 'cpu')
    model.to(device).eval()
    if test_outputs is None:
        test_outputs = (multi_gpu_test if distributed else single_gpu_test)(model, loaders)
    eval_kwargs = {k: v for k, v in dict(cfg.get('evaluation', {}) or {}).items() if k not in _HOOK_ARGS}
    metrics = {}
    for name, ds in datasets.items():
        results = test_outputs.get(name)
        if results is None:      # multi_gpu_test gathers on rank 0: the other ranks have nothing to evaluate
            metrics[name] = None
            continue
        kw = dict(eval_kwargs.get(ds.task, {}) or {})
        kw.update((eval_override or {}).get(ds.task, {}))
        metrics[name] = ds.evaluate(results, **kw)
    return metrics, test_outputs
