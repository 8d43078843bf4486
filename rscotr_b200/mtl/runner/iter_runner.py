"""Minimal IterBasedRunner with mmcv semantics: loop over an IterLoader of
the MultiDataLoader until max_iters, call the step engine, then the registered hooks."""
import time


class _LogBuffer(dict):
    def update(self, vars, count=1):
        for k, v in vars.items():
            self[k] = v


class IterBasedRunner:
    def __init__(self, engine, max_iters, work_dir=None, logger=None, meta=None, log_interval=50):
        self.engine, self.model, self.optimizer = engine, engine.model, engine.optimizer
        self.max_iters, self.work_dir, self.logger, self.meta = max_iters, work_dir, logger, meta or {}
        self.log_interval = log_interval
        self.hooks = []
        self.log_buffer = _LogBuffer()
        self.outputs = None

    iter = property(lambda self: self.engine.iter)

    def register_hook(self, hook, priority='NORMAL'):
        self.hooks.append(hook)

    def run(self, data_loaders, workflow=(('train', 1),), **kwargs):
        loader = data_loaders[0]
        it = iter(loader)
        self.model.train()
        t0 = time.time()
        while self.engine.iter < self.max_iters:
            try:
                batch = next(it)
            except StopIteration:
                it = iter(loader)
                batch = next(it)
            self.outputs = self.engine.train_iter(batch)
            if self.log_interval and self.engine.iter % self.log_interval == 0:
                self.log_buffer.update(dict(self.outputs['log_vars'].items()), self.outputs['num_samples'])
                if self.logger:
                    self.logger.info('iter %d  %.3fs/iter  %s', self.engine.iter,
                                     (time.time() - t0) / self.log_interval, dict(self.log_buffer))
                t0 = time.time()
            for h in self.hooks:
                if hasattr(h, 'after_train_iter'):
                    h.after_train_iter(self)
