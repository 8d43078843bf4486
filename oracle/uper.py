"""CPU oracle of the UPerNet decode head.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED against the reference (row a20 is a north-star extension: the reference repo contains no
UPerNet; mmseg 0.28 `decode_heads/{uper_head,psp_head,fcn_head}.py` is absent from this image).  The functions
restate the published layer order on an mmseg-layout state dict, one op per line, and are cross-checked in
tests/test_uper_head.py against the independent implementation importable here
(`transformers.models.upernet.modeling_upernet.UperNetHead` / `UperNetFCNHead`)."""
import torch
import torch.nn.functional as F


def conv_module(sd, prefix, x, padding=0, training=False, eps=1e-5):
    """mmcv ConvModule(conv -> BN -> ReLU); batch statistics when training."""
    x = F.conv2d(x, sd[prefix + 'conv.weight'], sd.get(prefix + 'conv.bias'), padding=padding)
    if prefix + 'bn.weight' in sd:
        if training:
            mean = x.mean(dim=(0, 2, 3), keepdim=True)
            var = x.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
        else:
            mean = sd[prefix + 'bn.running_mean'].view(1, -1, 1, 1)
            var = sd[prefix + 'bn.running_var'].view(1, -1, 1, 1)
        x = (x - mean) / torch.sqrt(var + eps) * sd[prefix + 'bn.weight'].view(1, -1, 1, 1) + sd[prefix + 'bn.bias'].view(1, -1, 1, 1)
    return x.clamp(min=0)


def _up(x, size):
    return F.interpolate(x, size=tuple(size), mode='bilinear', align_corners=False)


def uper_head(sd, prefix, feats, pool_scales=(1, 2, 3, 6), training=False):
    """feats: list of 4 maps, fine -> coarse.  Returns the (B, num_classes, H/4, W/4) logits (no dropout)."""
    top = feats[-1]
    psp = [top]
    for k, s in enumerate(pool_scales):
        p = F.adaptive_avg_pool2d(top, s)
        psp.append(_up(conv_module(sd, '%spsp_modules.%d.1.' % (prefix, k), p, training=training), top.shape[2:]))
    lat = [conv_module(sd, '%slateral_convs.%d.' % (prefix, i), feats[i], training=training) for i in range(len(feats) - 1)]
    lat.append(conv_module(sd, prefix + 'bottleneck.', torch.cat(psp, 1), padding=1, training=training))
    for i in range(len(lat) - 1, 0, -1):
        lat[i - 1] = lat[i - 1] + _up(lat[i], lat[i - 1].shape[2:])
    outs = [conv_module(sd, '%sfpn_convs.%d.' % (prefix, i), lat[i], padding=1, training=training) for i in range(len(lat) - 1)]
    outs.append(lat[-1])
    outs = [outs[0]] + [_up(o, outs[0].shape[2:]) for o in outs[1:]]
    x = conv_module(sd, prefix + 'fpn_bottleneck.', torch.cat(outs, 1), padding=1, training=training)
    return F.conv2d(x, sd[prefix + 'conv_seg.weight'], sd[prefix + 'conv_seg.bias'])


def fcn_head(sd, prefix, feats, in_index=2, num_convs=1, training=False):
    x = feats[in_index]
    for i in range(num_convs):
        x = conv_module(sd, '%sconvs.%d.' % (prefix, i), x, padding=1, training=training)
    return F.conv2d(x, sd[prefix + 'conv_seg.weight'], sd[prefix + 'conv_seg.bias'])


def seg_losses(logit, label, ignore_index, loss_weight=1.0):
    """mmseg BaseDecodeHead.losses."""
    up = _up(logit, label.shape[2:])
    lab = label.squeeze(1)
    ce = F.cross_entropy(up, lab, reduction='none', ignore_index=ignore_index).mean() * loss_weight
    valid = lab != ignore_index
    acc = ((up.argmax(1) == lab) & valid).sum().float() * 100.0 / valid.sum().clamp(min=1)
    return dict(loss_ce=ce, acc_seg=acc)
