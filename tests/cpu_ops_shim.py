"""TEST INFRASTRUCTURE: run the host-side model logic on CPU by substituting the oracle
for the CUDA ops (monkeypatch of rscotr_b200.ops).  The product never does this -- its
ops raise on CPU tensors; this shim exists so `-m "not gpu"` tests can exercise module
wiring, losses, matching and the step engine without a GPU."""
import contextlib

import torch
import torch.nn.functional as F

from oracle import swin as osw
from oracle import transformer as otr


def _focal(pred, target, gamma=2.0, alpha=0.25):
    p = pred.float().sigmoid()
    t = F.one_hot(target, pred.shape[1] + 1)[:, :pred.shape[1]].float()
    pt = (1 - p) * t + p * (1 - t)
    fw = (alpha * t + (1 - alpha) * (1 - t)) * pt.pow(gamma)
    return F.binary_cross_entropy_with_logits(pred.float(), t, reduction='none') * fw


def _msda(value, shapes, starts, loc, w, im2col_step=64):
    return otr.ms_deform_attn_core(value, [(int(h), int(ww)) for h, ww in shapes.tolist()], loc, w)


def _gap(x, channels_last=False):
    return x.mean(1) if channels_last else x.flatten(2).mean(-1)


@contextlib.contextmanager
def cpu_ops():
    from rscotr_b200 import ops
    saved = {k: getattr(ops, k) for k in ('wmsa', 'patch_merge_ln', 'ms_deform_attn', 'global_avg_pool',
                                          'bilinear_resize', 'sigmoid_focal_loss', 'layer_norm', 'linear')}
    ops.wmsa = lambda qkv, b, t, hw, heads, ws=7, shift=0, scale=None: osw.wmsa_core(qkv, b, t, hw, heads, ws, shift, scale)
    ops.patch_merge_ln = lambda x, hw, g, b, eps=1e-5: osw.patch_merge_ln(x, hw, g, b, eps)
    ops.ms_deform_attn = _msda
    ops.global_avg_pool = _gap
    ops.bilinear_resize = lambda x, size: F.interpolate(x, size=tuple(size), mode='bilinear', align_corners=False)
    ops.sigmoid_focal_loss = _focal
    ops.linear = lambda x, w, b=None, rows=None: F.linear(x, w if rows is None else w[rows[0]:rows[1]],
                                                          b if rows is None or b is None else b[rows[0]:rows[1]])
    ops.layer_norm = lambda x, g, b, eps=1e-5, out_dtype=None: F.layer_norm(x, (x.shape[-1],), g, b, eps)
    try:
        yield
    finally:
        for k, v in saved.items():
            setattr(ops, k, v)
