"""DataContainer-free batching for the real datasets (reference: the mm* `build_dataloader`
calls of mtl/data/build.py:52-66 with the arguments of mtl/data/prepare_loader_args.py).

A batch is a plain dict in the format the step engine already consumes from the synthetic loaders:
`img` (B,C,H,W) -- padded right/bottom to the largest image of the batch --, `img_metas` list of dicts,
`gt_label` (B,), `gt_bboxes` / `gt_labels` per-image lists, `gt_semantic_seg` (B,1,H,W).  Test batches keep
the one-entry-per-augmentation lists of MultiScaleFlipAug.  Tensors are pinned by the DataLoader, so the
engine's copy stream can prefetch them (StepEngine.prefetch)."""
import random

import numpy as np
import torch
import torch.distributed as dist
from torch.utils.data import DataLoader, Sampler

_STACK = ('img', 'gt_label', 'gt_semantic_seg')
_PAD_VALUE = dict(img=0, gt_semantic_seg=255)


def _stack_padded(tensors, value):
    if all(t.shape == tensors[0].shape for t in tensors):
        return torch.stack(tensors)
    h, w = max(t.shape[-2] for t in tensors), max(t.shape[-1] for t in tensors)
    out = tensors[0].new_full((len(tensors),) + tuple(tensors[0].shape[:-2]) + (h, w), value)
    for i, t in enumerate(tensors):
        out[i, ..., :t.shape[-2], :t.shape[-1]] = t
    return out


def collate(samples):
    first = samples[0]
    batch = {}
    for k in first:
        vals = [s[k] for s in samples]
        if isinstance(first[k], list):                          # MultiScaleFlipAug: one entry per augmentation
            batch[k] = [collate([{k: v[a]} for v in vals])[k] for a in range(len(first[k]))]
        elif k in _STACK:
            vals = [v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v)) for v in vals]
            batch[k] = _stack_padded(vals, _PAD_VALUE.get(k, 0)) if vals[0].dim() >= 2 else torch.stack(vals)
        else:
            batch[k] = vals
    return batch


def worker_init_fn(worker_id, num_workers, rank, seed):
    """mmdet/mmcls/mmseg worker_init_fn: a distinct seed per (rank, worker)."""
    try:                                   # (mmdet setup_multi_processes: OpenCV threads off inside loader workers)
        import cv2
        cv2.setNumThreads(0)
    except ImportError:
        pass
    s = num_workers * rank + worker_id + seed
    np.random.seed(s)
    random.seed(s)
    torch.manual_seed(s)


class DistributedSampler(Sampler):
    """epoch-seeded permutation, rank-strided, rounded up to a multiple of the world size."""

    def __init__(self, dataset, num_replicas=1, rank=0, shuffle=True, seed=0, round_up=True):
        self.n, self.world, self.rank, self.shuffle, self.seed, self.epoch = len(dataset), num_replicas, rank, shuffle, seed or 0, 0
        self.num_samples = -(-self.n // self.world) if round_up else self.n // self.world
        self.total = self.num_samples * self.world

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __len__(self):
        return self.num_samples

    def __iter__(self):
        if self.shuffle:
            g = torch.Generator().manual_seed(self.epoch + self.seed)
            idx = torch.randperm(self.n, generator=g).tolist()
        else:
            idx = list(range(self.n))
        idx = (idx * (self.total // max(len(idx), 1) + 1))[:self.total]
        return iter(idx[self.rank:self.total:self.world])


class InfiniteGroupBatchSampler(Sampler):
    """mmdet InfiniteGroupBatchSampler (IterBasedRunner): endless stream of batches whose images share the
    aspect-ratio flag; indices come from one seeded permutation stream, strided over the ranks."""

    def __init__(self, dataset, batch_size=1, world_size=1, rank=0, seed=0, shuffle=True):
        self.n, self.batch_size, self.world, self.rank, self.seed, self.shuffle = len(dataset), batch_size, world_size, rank, seed or 0, shuffle
        self.flag = np.asarray(getattr(dataset, 'flag', np.zeros(self.n, dtype=np.uint8)))
        self.size = self.n

    def _indices(self):
        g = torch.Generator().manual_seed(self.seed)
        while True:
            order = torch.randperm(self.n, generator=g).tolist() if self.shuffle else list(range(self.n))
            for i in order[self.rank::self.world]:
                yield i

    def __iter__(self):
        buf = {int(f): [] for f in np.unique(self.flag)} or {0: []}
        for i in self._indices():
            b = buf[int(self.flag[i])]
            b.append(i)
            if len(b) == self.batch_size:
                yield list(b)
                del b[:]

    def __len__(self):
        return self.size

    def set_epoch(self, epoch):
        pass


def _dist_info(distributed):
    if distributed and dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def build_dataloader(dataset, samples_per_gpu=1, workers_per_gpu=0, num_gpus=1, dist=False, shuffle=True, seed=None,
                     drop_last=False, pin_memory=True, persistent_workers=False, round_up=True, runner_type='EpochBasedRunner',
                     sampler_cfg=None, **kwargs):
    rank, world = _dist_info(dist)
    seed = 0 if seed is None else seed
    common = dict(num_workers=workers_per_gpu, collate_fn=collate, pin_memory=pin_memory and torch.cuda.is_available(),
                  persistent_workers=persistent_workers and workers_per_gpu > 0,
                  worker_init_fn=(lambda w: worker_init_fn(w, workers_per_gpu, rank, seed)))
    if runner_type == 'IterBasedRunner' and dataset.task == 'det':
        bs = InfiniteGroupBatchSampler(dataset, samples_per_gpu, world, rank, seed, shuffle)
        loader = DataLoader(dataset, batch_sampler=bs, **common)
        loader.infinite = True
        return loader
    if dist or not shuffle:
        sampler = DistributedSampler(dataset, world, rank, shuffle, seed, round_up)
    else:
        sampler = None
    g = torch.Generator().manual_seed(seed)
    return DataLoader(dataset, batch_size=samples_per_gpu if dist else samples_per_gpu * max(num_gpus, 1), sampler=sampler,
                      shuffle=shuffle if sampler is None else False, drop_last=drop_last, generator=g, **common)
