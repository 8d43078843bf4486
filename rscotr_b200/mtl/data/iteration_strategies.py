"""Iteration strategies (reference mtl/data/iteration_strategies.py:13-257, an MMF
derivative) without omegaconf.  `__call__` returns the index of the dataset the next
batch is pulled from."""
import numpy as np


class _Cfg(dict):
    __getattr__ = dict.__getitem__


class IterationStrategy:
    name = None
    defaults = {}

    def __init__(self, dataloaders, config=None, *args, **kwargs):
        cfg = dict(name=self.name, **self.defaults)
        cfg.update(config or {})
        cfg.update(kwargs)
        self.config = _Cfg(cfg)
        self.dataloaders = dataloaders

    @classmethod
    def from_params(cls, dataloaders, **kwargs):
        return cls(dataloaders, kwargs)

    @property
    def should_exhaust_all_iterators(self):
        return False

    def __call__(self, *args, **kwargs):
        raise NotImplementedError("__call__ hasn't been implemented")


class ConstantIterationStrategy(IterationStrategy):
    name, defaults = 'constant', dict(idx=0)

    def __init__(self, dataloaders, config=None, *args, **kwargs):
        super().__init__(dataloaders, config, *args, **kwargs)
        self._idx = self.config.idx

    @property
    def should_exhaust_all_iterators(self):
        return True

    def __call__(self, *args, **kwargs):
        return self._idx


class RoundRobinIterationStrategy(IterationStrategy):
    name, defaults = 'round_robin', dict(start_idx=0)

    def __init__(self, dataloaders, config=None, *args, **kwargs):
        super().__init__(dataloaders, config, *args, **kwargs)
        self._current_idx = self.config.start_idx

    def __call__(self, *args, **kwargs):
        nxt = self._current_idx
        self._current_idx = (self._current_idx + 1) % len(self.dataloaders)
        return nxt


class RepeatedSequenceIterationStrategy(IterationStrategy):
    """cycles through a fixed sequence of dataset indices (configs/multi/slvl_strategies/repeated_sequence.py)."""
    name, defaults = 'repeated_sequence', dict(sequence=(0,))

    def __init__(self, dataloaders, config=None, *args, **kwargs):
        super().__init__(dataloaders, config, *args, **kwargs)
        self._seq = list(self.config.sequence)
        assert all(0 <= i < len(dataloaders) for i in self._seq)
        self._pos = 0

    def __call__(self, *args, **kwargs):
        nxt = self._seq[self._pos]
        self._pos = (self._pos + 1) % len(self._seq)
        return nxt


class RandomIterationStrategy(IterationStrategy):
    name = 'random'

    def __call__(self, *args, **kwargs):
        return int(np.random.choice(len(self.dataloaders), 1)[0])


class WeightedRandomIterationStrategy(IterationStrategy):
    """p-weighted random choice.  (The reference leaves self.p unset when p already sums
    to one, iteration_strategies.py:192-196; here p is always normalised.)"""
    name, defaults = 'weighted_random', dict(p=None)

    def __init__(self, dataloaders, config=None, *args, **kwargs):
        super().__init__(dataloaders, config, *args, **kwargs)
        p = self.config.p
        p = np.ones(len(dataloaders)) if p is None else np.asarray(p, dtype=np.float64)
        assert len(p) == len(dataloaders) and (p >= 0).all() and p.sum() > 0
        self.p = p / p.sum()

    def __call__(self, *args, **kwargs):
        return int(np.random.choice(len(self.dataloaders), 1, p=self.p)[0])


class SizeProportionalIterationStrategy(IterationStrategy):
    name = 'size_proportional'

    def __init__(self, dataloaders, config=None, *args, **kwargs):
        super().__init__(dataloaders, config, *args, **kwargs)
        # the reference weighs by DATASET size (len(loader.dataset), iteration_strategies.py:219-246), not by the
        # number of batches; loaders without a sized dataset fall back to their own length
        def size(l):
            ds = getattr(l, 'dataset', None)
            return len(ds) if ds is not None and hasattr(ds, '__len__') else len(l)
        sizes = np.array([max(size(l), 1) for l in dataloaders.values()], dtype=np.float64)
        self._p = sizes / sizes.sum()

    @property
    def should_exhaust_all_iterators(self):
        return True

    def __call__(self, *args, **kwargs):
        return int(np.random.choice(len(self.dataloaders), 1, p=self._p)[0])
