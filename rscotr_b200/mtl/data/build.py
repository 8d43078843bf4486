"""Data-plane builders with the reference's names (mtl/data/build.py:21-100).  The
dataset builders produce SYNTHETIC datasets of the configured shapes (no data on
disk, no network); everything else -- per-dataset config splice, strategy registry,
MultiDataLoader -- follows the reference."""
import copy
import os

from ...config import Config, ConfigDict
from . import iteration_strategies as strategies
from .multi_data_loader import MultiDataLoader
from .synthetic import SyntheticDataset, _SyntheticLoader

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


def load_data_cfg(cfg, config_root=None):
    data_cfg = cfg.data
    for ds in list(data_cfg.keys()):
        _cfg = dict(data_cfg[ds])
        path = _cfg.pop('config')
        task = _cfg.pop('task')
        if config_root is not None and not os.path.isabs(path):
            path = os.path.join(config_root, path)
        config = {}
        if os.path.isfile(path):
            try:
                config = dict(Config.fromfile(path)._cfg_dict)
            except Exception:          # dataset pipelines reference mm* types that are not needed here
                config = {}
        base = ConfigDict(config)
        for k, v in _cfg.items():
            if isinstance(v, dict) and isinstance(base.get(k), dict):
                merged = dict(base[k])
                merged.update(v)
                base[k] = ConfigDict(merged)
            else:
                base[k] = v
        data_cfg[ds] = ConfigDict(dict(task=task, config=base))


def build_datasets(data_cfg, split='train', synthetic=None):
    assert split in ('train', 'val', 'test')
    synthetic = dict(synthetic or {})
    datasets = dict()
    for ds in data_cfg.keys():
        task = data_cfg[ds]['task']
        kw = dict(synthetic.get(task, {}))
        kw.setdefault('img_size', synthetic.get('img_size', (800, 800)))
        datasets[ds] = SyntheticDataset(task, **kw)
    return datasets


def build_dataloaders(cfg, distributed, datasets, train=True):
    data_loaders = dict()
    data_cfg = cfg.data
    rank = int(os.environ.get('RANK', 0)) if distributed else 0
    for i, ds in enumerate(datasets.keys()):
        dcfg = data_cfg[ds]['config'].get('data', {}) if 'config' in data_cfg[ds] else {}
        bs = dcfg.get('samples_per_gpu', 1)
        length = cfg.get('synthetic', {}).get('length', {}).get(ds, _DEFAULT_LENGTH[datasets[ds].task])
        data_loaders[ds] = _SyntheticLoader(datasets[ds], bs, length, seed=(cfg.get('seed', 0) or 0) * 1000 + 17 * i + rank)
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
