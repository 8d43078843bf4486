"""MultiEvalDatasets (reference mtl/data/multi_eval_dataset.py:8-33): a validation / test dataset that tags every
sample with its task, and otherwise behaves like the dataset it wraps.  (No DataContainer: the tag is a plain string
and the collate function turns it into a per-batch list.)"""
from torch.utils.data import Dataset


class MultiEvalDatasets(Dataset):
    def __init__(self, dataset):
        assert hasattr(dataset, 'task')
        self.dataset = dataset
        self.task = getattr(dataset, 'task')

    def __getitem__(self, item):
        data = self.dataset.__getitem__(item)
        data['task'] = self.task
        return data

    def __len__(self):
        return len(self.dataset)

    def __repr__(self):
        return 'task: %s ' % self.task + repr(self.dataset)

    def __getattr__(self, item):
        if item == 'dataset':
            raise AttributeError(item)
        return getattr(self.dataset, item)
