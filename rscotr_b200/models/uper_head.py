"""UPerNet decode head (PPM + FPN) and the FCN auxiliary head for the segmentation-only configuration
(BASELINE config #5: Swin-B, 512x512 Potsdam tiles).  The reference repo has no UPerNet of
its own; the row follows mmseg 0.28 `UPerHead` / `PPM` / `FCNHead` / `EncoderDecoder`, with
mmseg's parameter names so its checkpoints load:
`psp_modules.{k}.1.{conv,bn}`, `bottleneck`, `lateral_convs.{i}`, `fpn_convs.{i}`, `fpn_bottleneck`, `conv_seg`.

The convolutions are im2col GEMMs on this repo's kernels (rsc_im2col_* gather + the tcgen05 rsc_linear_* GEMMs; 1x1
convolutions are the GEMM alone), the PPM pooling is rsc_adaptive_avgpool_*; the loss is
the fused bilinear-upsample + cross-entropy kernel shared with the Mask2Former head (rsc_upsample_ce_*), so
the (B, C, 512, 512) up-sampled logits are never written.  BatchNorm is per rank (SyncBN would add a
forward-time collective; the north star exchanges gradients only -- SURVEY D.5)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..config import MODELS, build_from_cfg
from .mtl import SingleTaskModel, add_prefix
from .bricks import BatchNorm2d, Conv2d
from .seg_head import resize


class AdaptiveAvgPool2d(nn.AdaptiveAvgPool2d):
    """PPM pooling on rsc_adaptive_avgpool_* (channels-last) for CUDA fp32 / bf16 maps with C % 8 == 0"""

    def forward(self, x):
        s = self.output_size if isinstance(self.output_size, int) else (
            self.output_size[0] if self.output_size[0] == self.output_size[1] else None)
        if s is not None and x.is_cuda and x.dim() == 4 and x.shape[1] % 8 == 0 and x.dtype in (torch.float32, torch.bfloat16):
            return ops.adaptive_avg_pool2d(x, s)
        return super().forward(x)


class ConvModule(nn.Module):
    """conv (no bias when a norm follows) -> BatchNorm -> ReLU, mmcv's sub-module names `conv` / `bn`."""

    def __init__(self, cin, cout, kernel_size, padding=0, dilation=1, norm=True, act=True):
        super().__init__()
        self.conv = Conv2d(cin, cout, kernel_size, padding=padding, dilation=dilation, bias=not norm)
        self.bn = BatchNorm2d(cout) if norm else None      # training forward: rsc_groupnorm_* with the ReLU fused
        self.act = act
        nn.init.kaiming_normal_(self.conv.weight, mode='fan_out', nonlinearity='relu')
        if self.conv.bias is not None:
            nn.init.zeros_(self.conv.bias)

    def forward(self, x):
        x = self.conv(x)
        if self.bn is not None:
            return self.bn(x, relu=self.act)
        return F.relu(x) if self.act else x


def seg_ce_losses(seg_logit, seg_label, ignore_index, loss_weight=1.0):
    """mmseg BaseDecodeHead.losses: bilinear resize to the label size, CE mean over ALL pixels (ignored ones
    contribute 0), top-1 accuracy over the non-ignored pixels."""
    if ops.upsample_ce_supported(seg_logit, seg_label):
        stats = ops.upsample_ce(seg_logit, seg_label.squeeze(1), ignore_index)
        return dict(loss_ce=stats[0] * (loss_weight / seg_label.numel()),
                    acc_seg=stats[1].detach() * (100.0 / stats[2].detach().clamp(min=1)))
    seg_logit = resize(seg_logit, size=seg_label.shape[2:])
    seg_label = seg_label.squeeze(1)
    ce = F.cross_entropy(seg_logit.float(), seg_label, reduction='none', ignore_index=ignore_index)
    with torch.no_grad():
        valid = seg_label != ignore_index
        correct = (seg_logit.argmax(1) == seg_label) & valid
        acc = correct.sum().float() * (100.0 / valid.sum().clamp(min=1).float())
    return dict(loss_ce=loss_weight * ce.mean(), acc_seg=acc)


class _DecodeHead(nn.Module):
    def __init__(self, in_channels, channels, num_classes, in_index, dropout_ratio=0.1, norm_cfg=None, align_corners=False,
                 ignore_index=255, loss_decode=None, **kwargs):
        super().__init__()
        self.in_channels, self.channels, self.num_classes, self.in_index = in_channels, channels, num_classes, in_index
        self.align_corners, self.ignore_index = align_corners, ignore_index
        self.loss_weight = float((loss_decode or {}).get('loss_weight', 1.0))
        self.norm = norm_cfg is not None
        if norm_cfg is not None and norm_cfg.get('type') not in ('BN', 'SyncBN'):
            raise KeyError('norm %s is not supported by the UPerNet heads' % norm_cfg.get('type'))
        self.conv_seg = Conv2d(channels, num_classes, kernel_size=1)
        self.dropout = nn.Dropout2d(dropout_ratio) if dropout_ratio > 0 else None
        nn.init.normal_(self.conv_seg.weight, mean=0, std=0.01)
        nn.init.zeros_(self.conv_seg.bias)

    def cls_seg(self, feat):
        if self.dropout is not None:
            feat = self.dropout(feat)
        return self.conv_seg(feat)

    def losses(self, seg_logit, seg_label):
        return seg_ce_losses(seg_logit, seg_label, self.ignore_index, self.loss_weight)

    def forward_train(self, inputs, img_metas, gt_semantic_seg, train_cfg=None):
        return self.losses(self(inputs), gt_semantic_seg)

    def forward_test(self, inputs, img_metas=None, test_cfg=None):
        return self(inputs)


@MODELS.register_module()
class UPerHead(_DecodeHead):
    def __init__(self, in_channels, channels, num_classes, in_index=(0, 1, 2, 3), pool_scales=(1, 2, 3, 6), **kwargs):
        super().__init__(list(in_channels), channels, num_classes, list(in_index), **kwargs)
        top = self.in_channels[-1]
        # pyramid pooling on the coarsest map: pooled to s x s, 1x1 conv, bilinearly back up
        self.psp_modules = nn.ModuleList(
            nn.Sequential(AdaptiveAvgPool2d(s), ConvModule(top, channels, 1, norm=self.norm)) for s in pool_scales)
        self.bottleneck = ConvModule(top + len(pool_scales) * channels, channels, 3, padding=1, norm=self.norm)
        self.lateral_convs = nn.ModuleList(ConvModule(c, channels, 1, norm=self.norm) for c in self.in_channels[:-1])
        self.fpn_convs = nn.ModuleList(ConvModule(channels, channels, 3, padding=1, norm=self.norm) for _ in self.in_channels[:-1])
        self.fpn_bottleneck = ConvModule(len(self.in_channels) * channels, channels, 3, padding=1, norm=self.norm)

    def psp_forward(self, x):
        size = x.shape[2:]
        outs = [x] + [resize(m(x), size=size, mode='bilinear', align_corners=self.align_corners) for m in self.psp_modules]
        return self.bottleneck(torch.cat(outs, dim=1))

    def forward(self, inputs):
        inputs = [inputs[i] for i in self.in_index]
        lat = [conv(x) for conv, x in zip(self.lateral_convs, inputs)]
        lat.append(self.psp_forward(inputs[-1]))
        for i in range(len(lat) - 1, 0, -1):                       # top-down pathway
            lat[i - 1] = lat[i - 1] + resize(lat[i], size=lat[i - 1].shape[2:], mode='bilinear', align_corners=self.align_corners)
        outs = [conv(x) for conv, x in zip(self.fpn_convs, lat[:-1])] + [lat[-1]]
        size = outs[0].shape[2:]
        outs = [outs[0]] + [resize(o, size=size, mode='bilinear', align_corners=self.align_corners) for o in outs[1:]]
        return self.cls_seg(self.fpn_bottleneck(torch.cat(outs, dim=1)))


@MODELS.register_module()
class FCNHead(_DecodeHead):
    def __init__(self, in_channels, channels, num_classes, in_index=-1, num_convs=2, kernel_size=3, concat_input=True,
                 dilation=1, **kwargs):
        super().__init__(in_channels, channels, num_classes, in_index, **kwargs)
        pad = (kernel_size // 2) * dilation
        convs = [ConvModule(in_channels if i == 0 else channels, channels, kernel_size, padding=pad, dilation=dilation,
                            norm=self.norm) for i in range(num_convs)]
        self.convs = nn.Sequential(*convs) if num_convs > 0 else nn.Identity()
        self.concat_input = concat_input
        if concat_input:
            self.conv_cat = ConvModule(in_channels + channels, channels, kernel_size, padding=kernel_size // 2, norm=self.norm)

    def forward(self, inputs):
        x = inputs[self.in_index]
        out = self.convs(x)
        if self.concat_input:
            out = self.conv_cat(torch.cat([x, out], dim=1))
        return self.cls_seg(out)


@MODELS.register_module()
class EncoderDecoder(SingleTaskModel):
    """mmseg EncoderDecoder(backbone, decode_head, auxiliary_head): single-task segmentation with the step engine's
    model interface (train_step / train_step_begin / _host / _finish, forward(return_loss=...))."""
    default_task = 'seg'

    def __init__(self, backbone, decode_head, neck=None, auxiliary_head=None, train_cfg=None, test_cfg=None, pretrained=None,
                 init_cfg=None):
        super().__init__()
        assert neck is None, 'no reference / baseline configuration puts a neck in front of UPerNet'
        self.backbone = build_from_cfg(backbone, MODELS)
        self.decode_head = build_from_cfg(decode_head, MODELS)
        aux = auxiliary_head if isinstance(auxiliary_head, (list, tuple)) else ([auxiliary_head] if auxiliary_head else [])
        self.auxiliary_head = nn.ModuleList(build_from_cfg(a, MODELS) for a in aux) if len(aux) != 1 else build_from_cfg(aux[0], MODELS)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg or dict(mode='whole')
        self.num_classes, self.align_corners = self.decode_head.num_classes, self.decode_head.align_corners

    def init_weights(self):
        if hasattr(self.backbone, 'init_weights'):
            self.backbone.init_weights()

    def extract_feat(self, img):
        return self.backbone(img)

    def _aux_heads(self):
        a = self.auxiliary_head
        return list(a) if isinstance(a, nn.ModuleList) else [a]

    def forward_train(self, img, img_metas, gt_semantic_seg, **kwargs):
        x = self.extract_feat(img)
        losses = add_prefix(self.decode_head.forward_train(x, img_metas, gt_semantic_seg, self.train_cfg), 'decode')
        heads = self._aux_heads()
        for k, head in enumerate(heads):
            losses.update(add_prefix(head.forward_train(x, img_metas, gt_semantic_seg, self.train_cfg),
                                     'aux' if len(heads) == 1 else 'aux_%d' % k))
        return losses

    def encode_decode(self, img, img_metas):
        out = self.decode_head.forward_test(self.extract_feat(img), img_metas, self.test_cfg)
        return resize(out, size=img.shape[2:], mode='bilinear', align_corners=self.align_corners)

    def simple_test(self, img, img_meta, rescale=True):
        assert self.test_cfg.get('mode', 'whole') == 'whole'
        logit = self.encode_decode(img, img_meta)
        if rescale:
            h, w = img_meta[0]['img_shape'][:2]
            logit = resize(logit[:, :, :h, :w].contiguous(), size=img_meta[0]['ori_shape'][:2], mode='bilinear',
                           align_corners=self.align_corners)
        out = F.softmax(logit.float(), dim=1)
        if img_meta[0].get('flip', False):
            out = out.flip(dims=(3,) if img_meta[0]['flip_direction'] == 'horizontal' else (2,))
        return list(out.argmax(dim=1).cpu().numpy())

    def forward(self, img, img_metas, return_loss=True, task=None, dataset_name=None, **kwargs):
        from .mtl import normalize_on_device
        if return_loss:
            return self.forward_train(normalize_on_device(img, img_metas), img_metas, **kwargs)
        if isinstance(img, list):
            img, img_metas = img[0], img_metas[0]
        return self.simple_test(normalize_on_device(img, img_metas), img_metas, **kwargs)
