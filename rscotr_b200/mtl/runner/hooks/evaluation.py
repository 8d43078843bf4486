"""MultiDatasetsEvalHook (reference mtl/runner/hooks/evaluation.py:29-148): every
`interval` iterations run all val loaders, call each dataset's evaluate(), log
"{dataset}.{metric}", keep the best weighted mean of the configured keys."""
import torch.distributed as dist

from ...engine.test import multi_gpu_test, single_gpu_test


class KeyIndicator:
    """the `save_best` keys and weights (reference evaluation.py:9-26); repr() = the best-checkpoint file stem."""

    def __init__(self, **kwargs):
        self.key_indicator = dict(**kwargs)

    def __getitem__(self, item):
        return self.key_indicator[item]

    def __repr__(self):
        return '_'.join(k.replace('.', '_') for k in self.key_indicator)

    def __len__(self):
        return len(self.key_indicator)

    def items(self):
        return self.key_indicator.items()


class MultiDatasetsEvalHook:
    def __init__(self, dataloaders, start=None, interval=1, by_epoch=False, save_best=None, test_fn=None,
                 **eval_kwargs):
        self.dataloaders, self.start, self.interval, self.by_epoch = dataloaders, start, interval, by_epoch
        # reference evaluation.py: a str / list save_best is normalised to {key: 1} weights
        if isinstance(save_best, str):
            save_best = {save_best: 1}
        elif isinstance(save_best, (list, tuple)):
            save_best = {k: 1 for k in save_best}
        self.save_best = dict(save_best) if isinstance(save_best, dict) else save_best
        self.test_fn = test_fn          # default: single process -> single_gpu_test; distributed -> multi_gpu_test (rank 0 evaluates)
        self.eval_kwargs = eval_kwargs
        self.best_score = None
        self.best_ckpt_path = None

    def after_train_iter(self, runner):
        # (runner.iter already counts the iteration that just finished: mmcv's `iter + 1`)
        if self.by_epoch or runner.iter % self.interval != 0:
            return
        if self.start is not None and runner.iter < self.start:
            return
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        test_fn = self.test_fn or (multi_gpu_test if distributed else single_gpu_test)
        results = test_fn(runner.model, self.dataloaders)
        if any(v is None for v in results.values()):       # (gathered on rank 0: the other ranks only helped to compute)
            runner.model.train()
            return
        score = self.evaluate(runner, results)
        if score is not None and (self.best_score is None or score > self.best_score):
            self.best_score = score
            runner.meta['best_score'] = score
            self._save_best(runner)
        runner.model.train()

    def _save_best(self, runner):
        """EvalHook._save_ckpt: keep ONE `best_<keys>_iter_N.pth` next to the periodic checkpoints
        (reference evaluation.py:124-128; file stem = KeyIndicator.__repr__, evaluation.py:18-21)."""
        import os
        work_dir = getattr(runner, 'work_dir', None)
        if not work_dir or not hasattr(runner, 'save_checkpoint') or not isinstance(self.save_best, dict):
            return
        if dist.is_available() and dist.is_initialized() and dist.get_rank() != 0:
            return
        stem = repr(KeyIndicator(**self.save_best))
        if self.best_ckpt_path and os.path.isfile(self.best_ckpt_path):
            os.remove(self.best_ckpt_path)
        name = 'best_%s_iter_%d.pth' % (stem, runner.iter)
        self.best_ckpt_path = runner.save_checkpoint(work_dir, name, create_symlink=False)
        runner.meta.setdefault('hook_msgs', {}).update(best_score=self.best_score, best_ckpt=self.best_ckpt_path)

    def evaluate(self, runner, results):
        logs = {}
        for name, loader in self.dataloaders.items():
            ds = loader.dataset
            if not hasattr(ds, 'evaluate'):
                continue
            res = ds.evaluate(results[name], **self.eval_kwargs.get(ds.task, {}))
            for k, v in res.items():
                logs['%s.%s' % (name, k)] = v
        runner.log_buffer.update(logs)
        if isinstance(self.save_best, dict) and logs:
            keys = [k for k in self.save_best if k in logs]
            if keys:
                return sum(logs[k] * self.save_best[k] for k in keys) / len(self.save_best)
        return None
