"""Checkpoint compatibility (SURVEY 8f rank 4 / section 5 "Checkpoint / resume").

* `swin_converter` / `load_swin_pretrained`: the `init_cfg=dict(type='Pretrained', ...)` +
  `convert_weights=True` branch of the backbone (reference cfg: configs/multi/MTL_slvlcls_swin-t-p4-w7_1x1_
  resisc&dior&potsdam.py:24-25; algorithm: mmdet 2.25 `swin_converter` + `SwinTransformer.init_weights`,
  SURVEY D.1): official Swin key names -> the `stages.{i}.blocks.{j}.attn.w_msa.*` layout, PatchMerging
  reduction / norm re-ordered from the official (x0,x1,x2,x3) concat to nn.Unfold's c*4+kh*2+kw order,
  relative-position tables bicubically resized when the window differs.
* `save_checkpoint` / `load_checkpoint` / `resume`: the mmcv runner's file layout
  (`meta`, `state_dict`, `optimizer`; tools/train.py:228-235) on top of the step engine's flat
  buffers.  The optimizer entry is a torch.optim.AdamW state dict with ONE param group per parameter in
  named_parameters() order (what mmcv's DefaultOptimizerConstructor builds with paramwise_cfg), so a file
  written here resumes in the reference and vice versa.
"""
import os
import time
from collections import OrderedDict

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------- pretrained backbone
def _unfold_reduction_order(w):
    out_c, in_c = w.shape
    return w.reshape(out_c, 4, in_c // 4)[:, [0, 2, 1, 3], :].transpose(1, 2).reshape(out_c, in_c)


def _unfold_norm_order(v):
    c = v.shape[0]
    return v.reshape(4, c // 4)[[0, 2, 1, 3], :].transpose(0, 1).reshape(c)


def swin_converter(ckpt):
    """official microsoft/Swin-Transformer state dict -> `backbone.`-prefixed mmdet layout."""
    out = OrderedDict()
    for k, v in ckpt.items():
        if k.startswith('head'):
            continue
        nv, nk = v, k
        if k.startswith('layers'):
            if 'attn.' in k:
                nk = k.replace('attn.', 'attn.w_msa.')
            elif 'mlp.' in k:
                if 'mlp.fc1.' in k:
                    nk = k.replace('mlp.fc1.', 'ffn.layers.0.0.')
                elif 'mlp.fc2.' in k:
                    nk = k.replace('mlp.fc2.', 'ffn.layers.1.')
                else:
                    nk = k.replace('mlp.', 'ffn.')
            elif 'downsample' in k:
                if 'reduction.' in k:
                    nv = _unfold_reduction_order(v)
                elif 'norm.' in k:
                    nv = _unfold_norm_order(v)
            nk = nk.replace('layers', 'stages', 1)
        elif k.startswith('patch_embed'):
            if 'proj' in k:
                nk = k.replace('proj', 'projection')
        out['backbone.' + nk] = nv
    return out


def _read(path_or_dict):
    if isinstance(path_or_dict, (dict, OrderedDict)):
        return path_or_dict
    return torch.load(path_or_dict, map_location='cpu', weights_only=False)


def load_swin_pretrained(backbone, checkpoint, convert_weights=False, logger=None):
    """SwinTransformer.init_weights, Pretrained branch.  Returns load_state_dict's result."""
    ckpt = _read(checkpoint)
    sd = ckpt.get('state_dict', ckpt.get('model', ckpt))
    if convert_weights:
        sd = swin_converter(sd)
    state = OrderedDict((k[9:], v) for k, v in sd.items() if k.startswith('backbone.'))
    if state and next(iter(state)).startswith('module.'):
        state = OrderedDict((k[7:], v) for k, v in state.items())
    own = backbone.state_dict()
    if state.get('absolute_pos_embed') is not None and 'absolute_pos_embed' in own:
        ape = state['absolute_pos_embed']
        n1, l, c1 = ape.shape
        n2, c2, h, w = own['absolute_pos_embed'].shape
        if n1 == n2 and c1 == c2 and l == h * w:
            state['absolute_pos_embed'] = ape.view(n2, h, w, c2).permute(0, 3, 1, 2).contiguous()
        elif logger:
            logger.warning('Error in loading absolute_pos_embed, pass')
    for key in [k for k in state if 'relative_position_bias_table' in k]:
        if key not in own:
            continue
        pre, cur = state[key], own[key]
        (l1, h1), (l2, h2) = pre.shape, cur.shape
        if h1 != h2:
            if logger:
                logger.warning('Error in loading %s, pass' % key)
        elif l1 != l2:
            s1, s2 = int(l1 ** 0.5), int(l2 ** 0.5)
            r = F.interpolate(pre.permute(1, 0).reshape(1, h1, s1, s1).float(), size=(s2, s2), mode='bicubic')
            state[key] = r.view(h2, l2).permute(1, 0).contiguous().to(pre.dtype)
    return backbone.load_state_dict(state, strict=False)


# ------------------------------------------------------------------------------------- runner checkpoints
def _named_params(model):
    return list(model.named_parameters())


def optimizer_state_dict(engine):
    """the step engine's optimizer state as a torch.optim.AdamW-style state dict, one group per parameter
    (flat fused AdamW on CUDA, or the stock torch optimizer of the CPU / non-AdamW configurations)."""
    opt = engine.optimizer
    flat = hasattr(opt, 'exp_avg')
    group_of = {id(p): g for g in opt.param_groups for p in g['params']}
    base_of = {id(p): b for g, b in zip(opt.param_groups, engine._base_lrs) for p in g['params']}
    span = {n: (s, e) for n, s, e in engine._spans}
    step = float(opt.step_t.item()) if flat else None
    state, groups = {}, []
    for i, (n, p) in enumerate(_named_params(engine.model)):
        g = group_of.get(id(p))
        entry = {k: v for k, v in (g or opt.param_groups[0]).items() if k != 'params'}
        entry['initial_lr'] = base_of.get(id(p), entry['lr'])
        if flat:
            entry['lr'] = entry['initial_lr'] * float(opt.lr_t.item()) / opt.base_lr
            entry.update(amsgrad=False, maximize=False, foreach=None, capturable=False, differentiable=False, fused=None)
        entry['params'] = [i]
        groups.append(entry)
        if flat:
            if n in span and step > 0:
                s, e = span[n]
                state[i] = dict(step=torch.tensor(step), exp_avg=opt.exp_avg[s:e].view_as(p).detach().cpu().clone(),
                                exp_avg_sq=opt.exp_avg_sq[s:e].view_as(p).detach().cpu().clone())
        elif p in opt.state and opt.state[p]:
            state[i] = {k: (v.detach().cpu().clone() if torch.is_tensor(v) else v) for k, v in opt.state[p].items()}
    # mmcv's DefaultOptimizerConstructor builds ONE group over model.parameters() when the config has no paramwise_cfg
    # and one group per parameter otherwise (mtl/utils/optimizer.py:40-55): identical hyper-parameters everywhere means the
    # former, so the checkpoint resumes in the reference's optimizer as well
    hyper = [{k: v for k, v in g.items() if k != 'params'} for g in groups]
    if len(groups) > 1 and all(h == hyper[0] for h in hyper[1:]):
        groups = [dict(hyper[0], params=[i for g in groups for i in g['params']])]
    return dict(state=state, param_groups=groups)


def load_optimizer_state_dict(engine, sd):
    opt = engine.optimizer
    flat = hasattr(opt, 'exp_avg')
    if flat and 'exp_avg' in sd:                         # FlatAdamW's own flat layout
        opt.load_state_dict(sd)
        return
    named = _named_params(engine.model)
    saved = [(i, g) for g in sd['param_groups'] for i in g['params']]
    assert len(saved) == len(named), 'optimizer state has %d parameters, the model %d' % (len(saved), len(named))
    span = {n: (s, e) for n, s, e in engine._spans}
    group_of = {id(p): g for g in opt.param_groups for p in g['params']}
    steps = []
    if flat:
        opt.exp_avg.zero_()
        opt.exp_avg_sq.zero_()
    for (n, p), (i, g) in zip(named, saved):
        st = sd['state'].get(i)
        if id(p) in group_of and not flat:
            group_of[id(p)]['lr'] = g['lr']
        if st is None:
            continue
        if flat:
            if n not in span:
                continue
            s, e = span[n]
            opt.exp_avg[s:e].copy_(st['exp_avg'].reshape(-1))
            opt.exp_avg_sq[s:e].copy_(st['exp_avg_sq'].reshape(-1))
            steps.append(float(st['step']))
        else:
            opt.state[p] = {k: (v.to(p.device).clone() if torch.is_tensor(v) and v.dim() else
                                (v.clone() if torch.is_tensor(v) else v)) for k, v in st.items()}
    if flat:
        # the reference steps every parameter on every iteration once it has a gradient (zero-filled afterwards),
        # so the per-parameter counts agree up to the first round-robin cycle; the flat kernel keeps one count
        opt.step_t.fill_(max(steps) if steps else 0.0)
        g0 = saved[0][1]
        if g0.get('initial_lr'):
            opt.lr_t.fill_(opt.base_lr * g0['lr'] / g0['initial_lr'])


def weights_to_cpu(state_dict):
    out = OrderedDict((k, v.detach().cpu().clone()) for k, v in state_dict.items())
    out._metadata = getattr(state_dict, '_metadata', OrderedDict())
    return out


def save_checkpoint(engine, filename, meta=None, save_optimizer=True):
    """mmcv.runner.save_checkpoint layout; written to a temp name and renamed (a crash never leaves a torn file)."""
    meta = dict(meta or {})
    meta.update(iter=int(engine.iter), epoch=int(meta.get('epoch', 0)), time=time.asctime())
    model = engine.model
    if getattr(model, 'CLASSES', None) is not None:
        meta.setdefault('CLASSES', model.CLASSES)
    ckpt = dict(meta=meta, state_dict=weights_to_cpu(model.state_dict()))
    if save_optimizer:
        ckpt['optimizer'] = optimizer_state_dict(engine)
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    tmp = '%s.tmp.%d' % (filename, os.getpid())
    torch.save(ckpt, tmp)
    os.replace(tmp, filename)
    return filename


def load_checkpoint(model, filename, strict=False, revise_keys=((r'^module\.', ''),), engine=None):
    """mmcv.runner.load_checkpoint: weights only (`load_from`).  With `engine`, the bf16 shadow is refreshed."""
    import re
    ckpt = _read(filename)
    sd = ckpt.get('state_dict', ckpt)
    meta = getattr(sd, '_metadata', OrderedDict())
    for pat, rep in revise_keys:
        sd = OrderedDict((re.sub(pat, rep, k), v) for k, v in sd.items())
    sd._metadata = meta
    result = model.load_state_dict(sd, strict=strict)
    if engine is not None:
        engine.sync_lp()
        engine._graphs.clear()
    return ckpt, result


def resume(engine, filename, resume_optimizer=True):
    """IterBasedRunner.resume: weights + iteration counter + optimizer state."""
    ckpt, _ = load_checkpoint(engine.model, filename, strict=False, engine=engine)
    engine.iter = int(ckpt['meta']['iter'])
    if resume_optimizer and 'optimizer' in ckpt:
        load_optimizer_state_dict(engine, ckpt['optimizer'])
    engine._lr_scale_now = None                          # the schedule is re-derived from engine.iter
    return ckpt['meta']


def find_latest_checkpoint(work_dir, suffix='pth'):
    """mmdet.utils.find_latest_checkpoint: `latest.pth` if present, else the highest iter_N / epoch_N."""
    if not os.path.isdir(work_dir):
        return None
    latest = os.path.join(work_dir, 'latest.' + suffix)
    if os.path.exists(latest):
        return latest
    best, best_n = None, -1
    for f in os.listdir(work_dir):
        stem, _, ext = f.rpartition('.')
        if ext != suffix or '_' not in stem:
            continue
        try:
            n = int(stem.split('_')[-1])
        except ValueError:
            continue
        if n > best_n:
            best, best_n = os.path.join(work_dir, f), n
    return best
