"""Minimal IterBasedRunner with mmcv semantics: loop over an IterLoader of
the MultiDataLoader until max_iters, call the step engine, then the registered hooks."""
import time


class _LogBuffer(dict):
    """mmcv LogBuffer semantics (values logged every `interval` iterations are the MEANS over the iterations of the
    window in which the key appeared) without a host sync per iteration: the packed device tensors of the steps'
    log vars are summed on the device per key set, and read back once when the window is logged."""

    def __init__(self):
        super().__init__()
        self._acc = {}                         # tuple(keys) -> [device sum, count, weight]

    def update(self, vars, count=1):
        for k, v in vars.items():
            self[k] = v

    def accumulate(self, log_vars):
        keys, packed = getattr(log_vars, '_keys', None), getattr(log_vars, '_packed', None)
        if keys is None:                       # a plain dict of floats
            keys, packed = list(log_vars.keys()), None
        slot = self._acc.get(tuple(keys))
        vals = packed.detach() if packed is not None else None
        if slot is None:
            self._acc[tuple(keys)] = [vals.clone() if vals is not None else [float(v) for v in log_vars.values()], 1,
                                      getattr(log_vars, '_weight', 1)]
        else:
            if vals is not None:
                slot[0] += vals
            else:
                slot[0] = [a + float(b) for a, b in zip(slot[0], log_vars.values())]
            slot[1] += 1

    def average(self):
        """-> {key: mean over the window}; clears the window."""
        out = {}
        for keys, (total, n, weight) in self._acc.items():
            vals = total.tolist() if hasattr(total, 'tolist') else list(total)
            vals = vals[len(vals) - len(keys):]                # (distributed: element 0 is the log-var count check)
            for k, v in zip(keys, vals):
                out[k] = v * weight / n
        self._acc.clear()
        self.update(out)
        return out


def dist_rank():
    import torch.distributed as dist
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


class _GradNorm:
    """the step's gradient norm in the (keys, packed device tensor, weight) form `_LogBuffer.accumulate` sums."""

    def __init__(self, t):
        self._keys, self._packed, self._weight = ['grad_norm'], t.detach().reshape(1).float(), 1


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

    # -- checkpoints (mmcv BaseRunner.save_checkpoint / load_checkpoint / IterBasedRunner.resume) ------------
    def save_checkpoint(self, out_dir, filename_tmpl='iter_{}.pth', meta=None, save_optimizer=True, create_symlink=True):
        import os
        import shutil
        from ..utils.checkpoint import save_checkpoint
        m = dict(self.meta or {})
        m.update(meta or {})
        name = filename_tmpl.format(self.iter) if '{}' in filename_tmpl else filename_tmpl
        path = save_checkpoint(self.engine, os.path.join(out_dir, name), meta=m, save_optimizer=save_optimizer)
        if create_symlink:
            shutil.copyfile(path, os.path.join(out_dir, 'latest.pth'))      # (mmcv links; a copy also works on any fs)
        return path

    def load_checkpoint(self, filename, map_location='cpu', strict=False, revise_keys=((r'^module\.', ''),)):
        from ..utils.checkpoint import load_checkpoint
        return load_checkpoint(self.model, filename, strict=strict, revise_keys=revise_keys, engine=self.engine)[0]

    def resume(self, checkpoint, resume_optimizer=True, map_location='default'):
        from ..utils.checkpoint import resume
        meta = resume(self.engine, checkpoint, resume_optimizer)
        self.meta.update({k: v for k, v in meta.items() if k in ('hook_msgs',)})
        if self.logger:
            self.logger.info('resumed from %s, iter %d', checkpoint, self.iter)
        return meta

    def _write_json_log(self, sec_per_iter):
        """mmcv TextLoggerHook's `<timestamp>.log.json`: one JSON object per log interval (mode, iter, lr, memory in MB,
        time per iteration, the window means of the log vars) -- the format the mm* analysis tools read."""
        if not self.work_dir:
            return
        if dist_rank() != 0:
            return
        import json
        import os
        import torch
        os.makedirs(self.work_dir, exist_ok=True)
        lrs = sorted({round(float(g['lr']), 12) for g in self.optimizer.param_groups})
        rec = dict(mode='train', epoch=1, iter=int(self.engine.iter), lr=lrs[-1] if lrs else None,
                   memory=int(torch.cuda.max_memory_allocated() // (1 << 20)) if torch.cuda.is_available() else 0,
                   time=round(float(sec_per_iter), 5))
        rec.update({k: (round(v, 5) if isinstance(v, float) else v) for k, v in self.log_buffer.items()
                    if isinstance(v, (int, float))})
        name = '%s.log.json' % (getattr(self, 'timestamp', None) or 'train')
        with open(os.path.join(self.work_dir, name), 'a') as f:
            f.write(json.dumps(rec) + '\n')

    def run(self, data_loaders, workflow=(('train', 1),), **kwargs):
        loader = data_loaders[0]
        state = dict(it=iter(loader))

        def next_batch():
            try:
                return next(state['it'])
            except StopIteration:                       # mmcv IterLoader: start the next epoch
                state['it'] = iter(loader)
                return next(state['it'])
        self.model.train()
        t0 = time.time()
        batch = next_batch() if self.engine.iter < self.max_iters else None
        while self.engine.iter < self.max_iters:
            self.outputs = self.engine.train_iter(batch)        # enqueues the step (asynchronous on the GPU)
            # fetch the next batch while the GPU works and start its host->device copy on the copy stream
            batch = next_batch() if self.engine.iter < self.max_iters else None
            if batch is not None:
                self.engine.prefetch(batch)
            self.log_buffer.accumulate(self.outputs['log_vars'])
            gn = self.engine.last_grad_norm                 # mmcv OptimizerHook logs the pre-clip gradient norm
            if gn is not None:
                self.log_buffer.accumulate(_GradNorm(gn))
            if self.log_interval and self.engine.iter % self.log_interval == 0:
                self.log_buffer.average()
                self._write_json_log((time.time() - t0) / self.log_interval)
                if self.logger:
                    self.logger.info('iter %d  %.3fs/iter  %s', self.engine.iter,
                                     (time.time() - t0) / self.log_interval, dict(self.log_buffer))
                t0 = time.time()
            for h in self.hooks:
                if hasattr(h, 'after_train_iter'):
                    h.after_train_iter(self)
