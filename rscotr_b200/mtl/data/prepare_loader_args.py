"""The nine per-task loader-argument functions of the reference (mtl/data/prepare_loader_args.py:8-211) as named
entry points over `build.prepare_dataloader_args`, plus the `prepare_dataloader_args[split][task]` table."""
from .build import prepare_dataloader_args as _prepare


def _make(split, task):
    def fn(distributed, data_cfg, num_gpus, seed):
        return _prepare(split, task, distributed, data_cfg, seed=seed, num_gpus=num_gpus)
    fn.__name__ = 'prepare_%s_%sloader_args' % (task, split)
    return fn


prepare_dataloader_args = {s: {t: _make(s, t) for t in ('cls', 'det', 'seg')} for s in ('train', 'val', 'test')}
for _s, _tab in prepare_dataloader_args.items():
    for _t, _f in _tab.items():
        globals()[_f.__name__] = _f
