"""MultiDataLoader: the co-training batch scheduler.

Interface and observable behaviour follow the reference's `mtl/data/multi_data_loader.py:17-204` (pinned by
`tests/golden/reference_multi_data_loader.json`, produced by running the reference class in place): one loader per
dataset, an iteration strategy names the dataset of every batch, each batch is tagged with `dataset_name` and
`task`.  The implementation is this repo's own: the schedule is a generator that owns the per-dataset streams,
instead of index / iterator bookkeeping spread over the instance.

Schedule contract (what the golden file checks, draw for draw):
  * starting an epoch draws once from the strategy; every delivered batch is followed by one more draw, so the
    strategy is always one decision ahead of the consumer;
  * a strategy that does not exhaust all iterators restarts a dataset that runs dry (the epoch never ends on its
    own: the runner stops at max_iters);
  * a strategy that exhausts all iterators retires a dry dataset, re-draws until it lands on a live one, and the
    epoch ends when every dataset has been retired;
  * with a single dataset the strategy is never consulted.
"""
import warnings

from .iteration_strategies import RoundRobinIterationStrategy


class _Stream:
    """one dataset's loader with its live iterator"""

    def __init__(self, name, loader):
        self.name, self.loader = name, loader
        self.task = getattr(getattr(loader, 'dataset', None), 'task', None)
        self.it = None
        self.retired = False

    def restart(self):
        self.it = iter(self.loader)

    def pull(self):
        return next(self.it)


class MultiDataLoader:
    def __init__(self, loaders, iteration_strategy=None):
        if not loaders:
            warnings.warn('Empty loaders passed into MultiDataLoader. This can have unintended consequences.')
            loaders = loaders or {}
        self.loaders = loaders
        self.iteration_strategy = iteration_strategy or RoundRobinIterationStrategy(loaders)
        self.dataset_list = list(loaders)
        self.lengths = {name: len(loader) for name, loader in loaders.items()}
        self.samplers = {name: loader.sampler for name, loader in loaders.items() if hasattr(loader, 'sampler')}
        self._streams = [_Stream(name, loader) for name, loader in loaders.items()]
        self._schedule = None
        self.current_index = 0

    # ---- introspection the reference exposes --------------------------------------------------------------
    num_datasets = property(lambda self: len(self._streams))
    current_dataset_name = property(lambda self: self.dataset_list[self.current_index])
    current_loader = property(lambda self: self.loaders[self.current_dataset_name])
    current_dataset = property(lambda self: getattr(self.current_loader, 'dataset', None))
    first_loader = property(lambda self: next(iter(self.loaders.values())))

    def get_datasets(self):
        return [loader.dataset for loader in self.loaders.values()]

    def __len__(self):
        return sum(self.lengths.values())

    # ---- the schedule -----------------------------------------------------------------------------------
    def _draw(self):
        """next live dataset according to the strategy (not consulted when there is nothing to choose)"""
        if len(self._streams) <= 1:
            self.current_index = 0
            return
        while True:
            choice = self.iteration_strategy()
            if not self._streams[choice].retired:
                self.current_index = choice
                return

    def _run(self):
        exhaust_all = self.iteration_strategy.should_exhaust_all_iterators
        for s in self._streams:
            s.retired = False
            s.restart()
        self._draw()
        while True:
            stream = self._streams[self.current_index]
            try:
                batch = stream.pull()
            except StopIteration:
                if exhaust_all:
                    stream.retired = True
                    if all(s.retired for s in self._streams):
                        return
                    self._draw()
                else:
                    stream.restart()
                stream = self._streams[self.current_index]
                try:
                    batch = stream.pull()       # (an empty / freshly dry loader ends the epoch, as in the reference)
                except StopIteration:
                    return
            self._draw()                        # stay one decision ahead of the consumer
            batch['dataset_name'] = stream.name
            batch['task'] = stream.task
            yield batch

    def __iter__(self):
        self._schedule = self._run()
        return self

    def __next__(self):
        if self._schedule is None:
            raise TypeError('MultiDataLoader: call iter() before next()')
        return next(self._schedule)

    def seed_sampler(self, epoch):
        """distributed samplers reshuffle per epoch"""
        for sampler in self.samplers.values():
            if sampler is not None and hasattr(sampler, 'set_epoch'):
                sampler.set_epoch(epoch)
