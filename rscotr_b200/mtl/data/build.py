"""Data-plane builders with the reference's names (mtl/data/build.py:21-100): per-dataset config
splice, dataset / dataloader builders, strategy registry, MultiDataLoader.

`build_datasets` builds the REAL datasets (mtl/data/datasets.py) when the files the dataset config names
exist, and seeded SYNTHETIC datasets of the configured shapes otherwise (this environment has no data on disk
and no network; bench.py and the tests ask for synthetic data explicitly).  Pass `synthetic=False` to
insist on the real data (missing files then raise)."""
import os
import warnings

from ...config import Config, ConfigDict, _wrap
from . import iteration_strategies as strategies
from .multi_data_loader import MultiDataLoader
from .synthetic import SyntheticDataset, _SyntheticLoader
from .datasets import build_dataset
from .loader import build_dataloader

strategies_map = {
    'constant': strategies.ConstantIterationStrategy,
    'round_robin': strategies.RoundRobinIterationStrategy,
    'random': strategies.RandomIterationStrategy,
    'size_proportional': strategies.SizeProportionalIterationStrategy,
    'repeated_sequence': strategies.RepeatedSequenceIterationStrategy,
    'weighted_random': strategies.WeightedRandomIterationStrategy,
}

# loader lengths of the reference datasets (configs/multi/slvl_strategies/batch-weighted_random.py:5)
_DEFAULT_LENGTH = dict(cls=394, det=5862, seg=1728)


def _merge(base, other):
    out = dict(base)
    for k, v in other.items():
        if k in out and isinstance(out[k], dict) and isinstance(v, dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


def load_data_cfg(cfg, config_root=None):
    data_cfg = cfg.data
    for ds in list(data_cfg.keys()):
        _cfg = dict(data_cfg[ds])
        path = _cfg.pop('config')
        task = _cfg.pop('task')
        if config_root is not None and not os.path.isabs(path):
            path = os.path.join(config_root, path)
        if not os.path.isfile(path):
            raise FileNotFoundError('dataset config %r of data.%s not found' % (path, ds))
        config = dict(Config.fromfile(path)._cfg_dict)
        base = _wrap(_merge(config, _cfg))             # (addict's recursive dict.update, like the reference's ConfigDict)
        data_cfg[ds] = ConfigDict(dict(task=task, config=base))


def _split_cfg(entry, split):
    cfg = entry.get('config') if hasattr(entry, 'get') else None
    data = cfg.get('data') if cfg is not None and hasattr(cfg, 'get') else None
    return data.get(split) if data is not None and hasattr(data, 'get') else None


def _files_exist(ds_cfg):
    root = ds_cfg.get('data_root')
    for key in ('data_prefix', 'ann_file', 'img_dir'):
        v = ds_cfg.get(key)
        if v:
            v = v if (root is None or os.path.isabs(v) or key == 'data_prefix') else os.path.join(root, v)
            if not os.path.exists(v):
                return False
    return True


def build_datasets(data_cfg, split='train', synthetic=None):
    assert split in ('train', 'val', 'test')
    datasets = dict()
    for ds in data_cfg.keys():
        task = data_cfg[ds]['task']
        ds_cfg = _split_cfg(data_cfg[ds], split)
        real = synthetic is False or (synthetic is None and ds_cfg is not None and _files_exist(ds_cfg))
        if real:
            if ds_cfg is None:
                raise KeyError('dataset %r has no data.%s config' % (ds, split))
            extra = dict(test_mode=True) if split != 'train' else None
            datasets[ds] = build_dataset(ds_cfg, task, extra)
            continue
        if synthetic is None:
            warnings.warn('dataset %r: files of data.%s not found, using a synthetic %s dataset' % (ds, split, task))
        syn = dict(synthetic or {})
        kw = dict(syn.get(task, {}))
        kw.setdefault('img_size', syn.get('img_size', (800, 800)))
        datasets[ds] = SyntheticDataset(task, **kw)
    return datasets


_LOADER_KEYS = ('train', 'val', 'test', 'train_dataloader', 'val_dataloader', 'test_dataloader')


def prepare_dataloader_args(split, task, distributed, data_cfg, seed=None, num_gpus=1):
    """the per-task DataLoader arguments of mtl/data/prepare_loader_args.py:8-211 in one function."""
    top = {k: v for k, v in data_cfg.items() if k not in _LOADER_KEYS}
    own = dict(data_cfg.get('%s_dataloader' % split, {}) or {})
    if task == 'cls':
        args = dict(num_gpus=num_gpus, dist=distributed, round_up=True, seed=seed, **top)
        if split != 'train':
            args.update(shuffle=False, sampler_cfg=None)
    elif task == 'det':
        if split == 'train':
            args = dict(samples_per_gpu=2, workers_per_gpu=2, num_gpus=num_gpus, dist=distributed, seed=seed,
                        runner_type='IterBasedRunner', persistent_workers=False)
        else:
            args = dict(samples_per_gpu=1, workers_per_gpu=2, dist=distributed, shuffle=False, persistent_workers=False)
        args.update({k: v for k, v in top.items() if k in ('samples_per_gpu', 'workers_per_gpu', 'persistent_workers')}
                    if split == 'train' else {k: v for k, v in top.items() if k == 'workers_per_gpu'})
    else:
        args = dict(num_gpus=num_gpus, dist=distributed, seed=seed, drop_last=True, **top)
        if split != 'train':
            args.update(samples_per_gpu=1, shuffle=False)
            if split == 'test':
                args.pop('drop_last', None)
                args.pop('seed', None)
    args.update(own)
    return args


def build_dataloaders(cfg, distributed, datasets, train=True):
    split = 'train' if train else 'val'
    data_loaders = dict()
    data_cfg = cfg.data
    rank = int(os.environ.get('RANK', 0)) if distributed else 0
    for i, ds in enumerate(datasets.keys()):
        dcfg = data_cfg[ds]['config'].get('data', {}) if 'config' in data_cfg[ds] else {}
        if isinstance(datasets[ds], SyntheticDataset):
            bs = dcfg.get('samples_per_gpu', 1)
            syn = cfg.get('synthetic') or {}
            length = syn.get('length', {}).get(ds, _DEFAULT_LENGTH[datasets[ds].task]) if train else \
                syn.get('val_length', {}).get(ds, 4)          # (synthetic validation: a few batches, not an epoch)
            data_loaders[ds] = _SyntheticLoader(datasets[ds], bs, length, seed=(cfg.get('seed', 0) or 0) * 1000 + 17 * i + rank)
            continue
        args = prepare_dataloader_args(split, datasets[ds].task, distributed, dcfg, seed=cfg.get('seed', None),
                                       num_gpus=len(cfg.get('gpu_ids', [0])))
        data_loaders[ds] = build_dataloader(datasets[ds], **args)
    return data_loaders


def build_iteration_strategy(cfg, data_loaders):
    if 'strategy' in cfg:
        kwargs = dict(cfg.strategy)
        strategy_type = kwargs.pop('type')
    else:
        strategy_type, kwargs = 'round_robin', dict()
    return strategies_map[strategy_type](data_loaders, **kwargs)


def build_multidataloader(cfg, distributed, datasets):
    train_loaders = build_dataloaders(cfg, distributed, datasets)
    iteration_strategy = build_iteration_strategy(cfg, train_loaders)
    return MultiDataLoader(train_loaders, iteration_strategy)
