"""Seeding helpers with the semantics the reference's tools/train.py:210-216 relies on (mmdet.apis
`init_random_seed` / `set_random_seed`): one seed agreed across ranks (rank 0's draw is broadcast; every rank sharing
the numpy seed is what keeps the random iteration strategies rank-consistent), optional per-rank offset
(`--diff-seed`), optional deterministic cuDNN."""
import random

import numpy as np
import torch
import torch.distributed as dist


def init_random_seed(seed=None, device='cuda'):
    if seed is not None:
        return int(seed)
    seed = int(np.random.randint(2 ** 31))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return seed
    t = torch.tensor(seed if dist.get_rank() == 0 else 0, dtype=torch.int32,
                     device=device if dist.get_backend() == 'nccl' else 'cpu')
    dist.broadcast(t, src=0)
    return int(t.item())


def set_random_seed(seed, deterministic=False):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    if deterministic:
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False
