"""MultiDataLoader (reference mtl/data/multi_data_loader.py:17-204): one loader per
dataset, an IterationStrategy picks which one the next batch comes from; exhausted
iterators are re-ignited; every batch is tagged with `dataset_name` and `task`."""
import warnings

from .iteration_strategies import RoundRobinIterationStrategy


class MultiDataLoader:
    def __init__(self, loaders, iteration_strategy=None):
        if loaders is None or len(loaders) == 0:
            warnings.warn('Empty loaders passed into MultiDataLoader. This can have unintended consequences.')
        if iteration_strategy is None:
            iteration_strategy = RoundRobinIterationStrategy(loaders)
        self._iteration_strategy = iteration_strategy
        self._loaders = loaders
        self._num_datasets = len(self.loaders)
        self.dataset_list = list(loaders.keys())
        self._iterators = {}
        self._finished_iterators = {}
        self.current_index = 0
        self.lengths = {name: len(loader) for name, loader in self.loaders.items()}
        self.samplers = {k: l.sampler for k, l in self.loaders.items() if hasattr(l, 'sampler')}

    def get_datasets(self):
        return [loader.dataset for loader in self.loaders.values()]

    loaders = property(lambda self: self._loaders)
    num_datasets = property(lambda self: self._num_datasets)
    iteration_strategy = property(lambda self: self._iteration_strategy)
    current_dataset_name = property(lambda self: self.dataset_list[self.current_index])
    current_loader = property(lambda self: self.loaders[self.current_dataset_name])
    current_iterator = property(lambda self: self._iterators[self.current_dataset_name])
    first_loader = property(lambda self: list(self.loaders.values())[0])

    @property
    def iterators(self):
        return self._iterators

    @iterators.setter
    def iterators(self, v):
        self._iterators = v

    @property
    def current_dataset(self):
        return getattr(self.current_loader, 'dataset', None)

    def __len__(self):
        return sum(self.lengths.values())

    def __iter__(self):
        self._finished_iterators = {}
        self._iterators = {key: iter(loader) for key, loader in self.loaders.items()}
        self.change_dataloader()
        return self

    def __next__(self):
        try:
            next_batch = next(self.current_iterator)
        except StopIteration:
            if self.iteration_strategy.should_exhaust_all_iterators:
                self._finished_iterators[self.current_dataset_name] = 1
                if len(self._finished_iterators) == self.num_datasets:
                    raise
                self.change_dataloader()
                next_batch = next(self.current_iterator)
            else:
                self._iterators[self.current_dataset_name] = iter(self.current_loader)
                next_batch = next(self.current_iterator)
        current_dataset_name = self.current_dataset_name
        current_task = getattr(self.current_dataset, 'task', None)
        self.change_dataloader()
        next_batch['dataset_name'] = current_dataset_name
        next_batch['task'] = current_task
        return next_batch

    def change_dataloader(self):
        if self.num_datasets <= 1:
            self.current_index = 0
            return
        choice = self.iteration_strategy()
        while self.dataset_list[choice] in self._finished_iterators:
            choice = self.iteration_strategy()
        self.current_index = choice

    def seed_sampler(self, epoch):
        for sampler in self.samplers.values():
            if sampler is not None and hasattr(sampler, 'set_epoch'):
                sampler.set_epoch(epoch)
