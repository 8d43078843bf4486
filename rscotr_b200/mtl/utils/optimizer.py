"""build_optimizer with the reference's paramwise rules (mtl/utils/optimizer.py:25-55 +
mmcv DefaultOptimizerConstructor custom_keys): for each parameter the
first custom key (sorted by length desc, then alphabetically) that is a SUBSTRING of
the full parameter name sets lr_mult / decay_mult.  Parameters with identical
(lr, weight_decay) share one param group (mmcv makes one group per parameter; the
update is identical, the launch count is not)."""
import copy

import torch


def param_settings(model, optimizer_cfg, paramwise_cfg):
    """(name, param, lr, weight_decay) per trainable parameter with mmcv's DefaultOptimizerConstructor rules:
    custom_keys first (longest key that is a substring of the name wins and ends the search), otherwise
    bias_lr_mult / bias_decay_mult for biases outside norm layers, norm_decay_mult for norm layers,
    dwconv_decay_mult for depth-wise convolutions."""
    base_lr = optimizer_cfg['lr']
    base_wd = optimizer_cfg.get('weight_decay', None)
    pw = paramwise_cfg or {}
    custom_keys = pw.get('custom_keys', {})
    sorted_keys = sorted(sorted(custom_keys.keys()), key=len, reverse=True)
    bias_lr_mult, bias_decay_mult = pw.get('bias_lr_mult', 1.), pw.get('bias_decay_mult', 1.)
    norm_decay_mult, dwconv_decay_mult = pw.get('norm_decay_mult', 1.), pw.get('dwconv_decay_mult', 1.)
    norms = (torch.nn.modules.batchnorm._BatchNorm, torch.nn.modules.instancenorm._InstanceNorm, torch.nn.GroupNorm,
             torch.nn.LayerNorm)
    kind = {}
    for module in model.modules():
        is_norm = isinstance(module, norms)
        is_dw = isinstance(module, torch.nn.Conv2d) and module.in_channels == module.groups and module.groups > 1
        for local, p in module.named_parameters(recurse=False):
            kind[id(p)] = (is_norm, is_dw, local)
    out = []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        lr, wd = base_lr, base_wd
        for key in sorted_keys:
            if key in name:
                lr = base_lr * custom_keys[key].get('lr_mult', 1.)
                if base_wd is not None:
                    wd = base_wd * custom_keys[key].get('decay_mult', 1.)
                break
        else:
            is_norm, is_dw, local = kind.get(id(p), (False, False, name.rsplit('.', 1)[-1]))
            if local == 'bias' and not is_norm:
                lr = base_lr * bias_lr_mult
            if base_wd is not None:
                if is_norm:
                    wd = base_wd * norm_decay_mult
                elif is_dw:
                    wd = base_wd * dwconv_decay_mult
                elif local == 'bias':
                    wd = base_wd * bias_decay_mult
        out.append((name, p, lr, wd))
    return out


def build_optimizer(model, cfg):
    optimizer_cfg = copy.deepcopy(dict(cfg))
    constructor_type = optimizer_cfg.pop('constructor', 'MTLOptimizerConstructor')
    if constructor_type not in ('MTLOptimizerConstructor', 'DefaultOptimizerConstructor'):
        raise KeyError('%s is not registered in the optimizer builder registry.' % constructor_type)
    paramwise_cfg = optimizer_cfg.pop('paramwise_cfg', None)
    if hasattr(model, 'module'):
        model = model.module
    opt_type = optimizer_cfg.pop('type')
    groups = {}
    for name, p, lr, wd in param_settings(model, optimizer_cfg, paramwise_cfg):
        groups.setdefault((lr, wd), []).append(p)
    params = []
    for (lr, wd), ps in groups.items():
        g = dict(params=ps, lr=lr)
        if wd is not None:
            g['weight_decay'] = wd
        params.append(g)
    cls = getattr(torch.optim, opt_type)
    kwargs = dict(optimizer_cfg)
    if opt_type in ('AdamW', 'Adam') and all(p.is_cuda for g in params for p in g['params']):
        kwargs.setdefault('fused', True)
    return cls(params, **kwargs)


class MTLOptimizerConstructor:
    """reference mtl/utils/optimizer.py:40-55 (an mmcv DefaultOptimizerConstructor): `constructor(model)` -> optimizer."""

    def __init__(self, optimizer_cfg, paramwise_cfg=None):
        self.optimizer_cfg, self.paramwise_cfg = dict(optimizer_cfg), paramwise_cfg

    def __call__(self, model):
        cfg = dict(self.optimizer_cfg)
        if self.paramwise_cfg:
            cfg['paramwise_cfg'] = self.paramwise_cfg
        return build_optimizer(model, cfg)


def build_optimizer_constructor(cfg):
    cfg = dict(cfg)
    t = cfg.pop('type')
    if t not in ('MTLOptimizerConstructor', 'DefaultOptimizerConstructor'):
        raise KeyError('%s is not registered in the optimizer builder registry.' % t)
    return MTLOptimizerConstructor(**cfg)
