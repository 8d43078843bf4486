"""AdamW over the step engine's flat parameter / gradient buffers: one fused kernel launch
(rsc_adamw_step) per contiguous run of parameters that share (lr multiplier, weight decay),
torch.optim.AdamW arithmetic, global-norm clip scale folded in, lr / step count on the device so
the launches are CUDA-graph replayable.  Param-group rules = mtl/utils/optimizer.py (reference
mtl/utils/optimizer.py:25-55 + mmcv custom_keys)."""
import torch

from ... import _lib
from ..utils.optimizer import param_settings


class FlatAdamW:
    def __init__(self, engine, cfg):
        cfg = dict(cfg)
        cfg.pop('type', None)
        cfg.pop('constructor', None)
        paramwise = cfg.pop('paramwise_cfg', None)
        self.base_lr = float(cfg['lr'])
        self.betas = tuple(cfg.get('betas', (0.9, 0.999)))
        self.eps = float(cfg.get('eps', 1e-8))
        base_wd = float(cfg.get('weight_decay', 1e-2))
        self.engine = engine
        dev = engine.device
        settings = {n: (lr, wd) for n, _, lr, wd in param_settings(engine.model, dict(lr=self.base_lr, weight_decay=base_wd),
                                                                    paramwise)}
        # contiguous runs of identical (lr_mult, wd) over the flat layout (alignment gaps included)
        self.runs = []
        total = engine.flat_param.numel()
        spans = engine._spans
        for k, (n, s0, e0) in enumerate(spans):
            lr, wd = settings[n]
            key = (lr / self.base_lr, wd)
            nxt = spans[k + 1][1] if k + 1 < len(spans) else total
            if self.runs and self.runs[-1][2] == key and self.runs[-1][1] == s0:
                self.runs[-1] = (self.runs[-1][0], nxt, key)
            else:
                self.runs.append((s0, nxt, key))
        self.exp_avg = torch.zeros_like(engine.flat_param)
        self.exp_avg_sq = torch.zeros_like(engine.flat_param)
        self.step_t = torch.zeros(1, dtype=torch.float32, device=dev)
        self.lr_t = torch.full((1,), self.base_lr, dtype=torch.float32, device=dev)
        groups = {}
        for (n, s0, e0), p in zip(spans, engine._params):
            groups.setdefault(settings[n], []).append(p)
        self.param_groups = [dict(params=ps, lr=lr, weight_decay=wd, betas=self.betas, eps=self.eps)
                             for (lr, wd), ps in groups.items()]

    def set_lr_scale(self, scale):
        self.lr_t.fill_(self.base_lr * scale)

    def zero_grad(self, set_to_none=False):
        self.engine.flat_grad.zero_()

    def step_flat(self, clip_coef=None):
        eng = self.engine
        self.step_t += 1
        st = torch.cuda.current_stream().cuda_stream
        clip_ptr = None if clip_coef is None else clip_coef.data_ptr()
        lp = eng.flat_param_lp.data_ptr() if eng.flat_param_lp is not None else None
        with torch.cuda.device(eng.device):
            for s0, e0, (lr_mult, wd) in self.runs:
                _lib.call('rsc_adamw_step', eng.flat_param.data_ptr() + 4 * s0, eng.flat_grad.data_ptr() + 4 * s0,
                          self.exp_avg.data_ptr() + 4 * s0, self.exp_avg_sq.data_ptr() + 4 * s0, e0 - s0,
                          self.lr_t.data_ptr(), lr_mult, self.betas[0], self.betas[1], self.eps, wd,
                          self.step_t.data_ptr(), clip_ptr, None if lp is None else lp + 2 * s0, st,
                          alg_bytes=(28 if lp is None else 30) * (e0 - s0))

    def step(self):
        self.step_flat(None)

    def state_dict(self):
        return dict(exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq, step=self.step_t, lr=self.lr_t)

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd['exp_avg'])
        self.exp_avg_sq.copy_(sd['exp_avg_sq'])
        self.step_t.copy_(sd['step'])
        self.lr_t.copy_(sd['lr'])
