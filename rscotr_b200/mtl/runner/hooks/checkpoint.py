"""CheckpointHook with mmcv's iteration-based behaviour (reference cfg
`checkpoint_config=dict(interval=100000)`, mtl/apis/train.py:80): every `interval` iterations rank 0
writes `iter_N.pth` plus `latest.pth`, and prunes to `max_keep_ckpts`."""
import os

import torch.distributed as dist


class CheckpointHook:
    def __init__(self, interval=-1, by_epoch=False, save_optimizer=True, out_dir=None, max_keep_ckpts=-1,
                 save_last=True, meta=None, **kwargs):
        self.interval, self.by_epoch, self.save_optimizer = interval, by_epoch, save_optimizer
        self.out_dir, self.max_keep_ckpts, self.save_last = out_dir, max_keep_ckpts, save_last
        self.meta = dict(meta) if meta else None         # (tools/train.py:229-235 stores library versions and CLASSES here)

    def after_train_iter(self, runner):
        if self.by_epoch or self.interval <= 0:
            return
        n = runner.iter
        if n % self.interval == 0 or (self.save_last and n == runner.max_iters):
            self._save(runner, n)

    def _save(self, runner, n):
        if dist.is_available() and dist.is_initialized() and dist.get_rank() != 0:
            return
        out_dir = self.out_dir or runner.work_dir
        if out_dir is None:
            return
        path = runner.save_checkpoint(out_dir, 'iter_%d.pth' % n, meta=self.meta, save_optimizer=self.save_optimizer)
        runner.meta.setdefault('hook_msgs', {})['last_ckpt'] = path
        if self.max_keep_ckpts > 0:
            for old in range(n - self.max_keep_ckpts * self.interval, 0, -self.interval):
                f = os.path.join(out_dir, 'iter_%d.pth' % old)
                if not os.path.exists(f):
                    break
                os.remove(f)
