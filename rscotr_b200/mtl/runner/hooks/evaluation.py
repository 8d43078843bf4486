"""MultiDatasetsEvalHook (reference mtl/runner/hooks/evaluation.py:29-148): every
`interval` iterations run all val loaders, call each dataset's evaluate(), log
"{dataset}.{metric}", keep the best weighted mean of the configured keys."""
from ...engine.test import single_gpu_test


class MultiDatasetsEvalHook:
    def __init__(self, dataloaders, start=None, interval=1, by_epoch=False, save_best=None, test_fn=None,
                 **eval_kwargs):
        self.dataloaders, self.start, self.interval, self.by_epoch = dataloaders, start, interval, by_epoch
        self.save_best = dict(save_best) if isinstance(save_best, dict) else save_best
        self.test_fn = test_fn or single_gpu_test
        self.eval_kwargs = eval_kwargs
        self.best_score = None

    def after_train_iter(self, runner):
        if self.by_epoch or (runner.iter + 1) % self.interval != 0:
            return
        if self.start is not None and runner.iter + 1 < self.start:
            return
        results = self.test_fn(runner.model, self.dataloaders)
        score = self.evaluate(runner, results)
        if score is not None and (self.best_score is None or score > self.best_score):
            self.best_score = score
            runner.meta['best_score'] = score
        runner.model.train()

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
