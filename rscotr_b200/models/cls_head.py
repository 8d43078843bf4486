"""Single-level classification head (reference models/multi/cls_head/slvl_cls_head.py,
mmcls LinearClsHead / GlobalAveragePooling / LabelSmoothLoss / Augments)."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..config import MODELS


class GlobalAveragePooling(nn.Module):
    """mmcls GlobalAveragePooling(dim=2) on the rsc_gap kernel.  Accepts logical
    NCHW tensors; channels-last strides (what SwinTransformer returns) take the
    coalesced (B,HW,C) path without a layout copy."""

    def _pool(self, x):
        B, C = x.shape[:2]
        if x.dim() == 4 and x.permute(0, 2, 3, 1).is_contiguous():
            return ops.global_avg_pool(x.permute(0, 2, 3, 1).reshape(B, -1, C), channels_last=True)
        return ops.global_avg_pool(x.contiguous(), channels_last=False)

    def forward(self, inputs):
        if isinstance(inputs, (tuple, list)):
            return tuple(self._pool(x) for x in inputs)
        return self._pool(inputs)


class LabelSmoothLoss(nn.Module):
    """mmcls LabelSmoothLoss(mode='original'): soft-target CE, sum / avg_factor."""

    def __init__(self, label_smooth_val, num_classes=None, mode='original', reduction='mean', loss_weight=1.0):
        super().__init__()
        assert mode == 'original' and reduction == 'mean'
        self.label_smooth_val, self.num_classes, self.loss_weight = label_smooth_val, num_classes, loss_weight

    def forward(self, cls_score, label, avg_factor=None):
        num_classes = self.num_classes or cls_score.shape[1]
        if label.dim() == 1 or (label.dim() == 2 and label.shape[1] == 1):
            one_hot = F.one_hot(label.view(-1), num_classes).float()
        else:
            one_hot = label.float()
        smooth = one_hot * (1 - self.label_smooth_val) + self.label_smooth_val / num_classes
        loss = -(smooth * F.log_softmax(cls_score.float(), dim=-1)).sum(-1)
        loss = loss.sum() / avg_factor if avg_factor is not None else loss.mean()
        return self.loss_weight * loss


class CrossEntropyLoss(nn.Module):
    def __init__(self, loss_weight=1.0, **kwargs):
        super().__init__()
        self.loss_weight = loss_weight

    def forward(self, cls_score, label, avg_factor=None):
        if label.dim() > 1:
            loss = -(label * F.log_softmax(cls_score.float(), -1)).sum(-1)
        else:
            loss = F.cross_entropy(cls_score.float(), label, reduction='none')
        loss = loss.sum() / avg_factor if avg_factor is not None else loss.mean()
        return self.loss_weight * loss


def one_hot_encoding(gt, num_classes):
    return F.one_hot(gt.view(-1).long(), num_classes).float()


class BatchMixup:
    def __init__(self, alpha, num_classes, prob=1.0):
        self.alpha, self.num_classes, self.prob = alpha, num_classes, prob

    def __call__(self, img, gt_label, lam=None, index=None):
        one_hot = one_hot_encoding(gt_label, self.num_classes)
        lam = float(np.random.beta(self.alpha, self.alpha)) if lam is None else lam
        index = torch.randperm(img.size(0), device=img.device) if index is None else index
        return lam * img + (1 - lam) * img[index], lam * one_hot + (1 - lam) * one_hot[index]


class BatchCutMix:
    def __init__(self, alpha, num_classes, prob=1.0, cutmix_minmax=None, correct_lam=True):
        self.alpha, self.num_classes, self.prob, self.correct_lam = alpha, num_classes, prob, correct_lam

    def rand_bbox(self, img_shape, lam, margin=0., count=None):
        ratio = np.sqrt(1 - lam)
        H, W = img_shape[-2:]
        cut_h, cut_w = int(H * ratio), int(W * ratio)
        margin_y, margin_x = int(margin * cut_h), int(margin * cut_w)
        cy = np.random.randint(0 + margin_y, H - margin_y, size=count)
        cx = np.random.randint(0 + margin_x, W - margin_x, size=count)
        yl, yh = np.clip(cy - cut_h // 2, 0, H), np.clip(cy + cut_h // 2, 0, H)
        xl, xh = np.clip(cx - cut_w // 2, 0, W), np.clip(cx + cut_w // 2, 0, W)
        return yl, yh, xl, xh

    def __call__(self, img, gt_label):
        one_hot = one_hot_encoding(gt_label, self.num_classes)
        lam = float(np.random.beta(self.alpha, self.alpha))
        index = torch.randperm(img.size(0), device=img.device)
        yl, yh, xl, xh = self.rand_bbox(img.shape, lam)
        if self.correct_lam:
            lam = 1. - (yh - yl) * (xh - xl) / float(img.shape[-2] * img.shape[-1])
        img = img.clone()
        img[:, :, yl:yh, xl:xh] = img[index, :, yl:yh, xl:xh]
        return img, lam * one_hot + (1 - lam) * one_hot[index]


class Identity:
    def __init__(self, num_classes, prob=1.0):
        self.num_classes, self.prob = num_classes, prob

    def __call__(self, img, gt_label):
        return img, one_hot_encoding(gt_label, self.num_classes)


class Augments:
    """mmcls Augments: pick ONE batch augment per call with the configured probabilities."""
    _TYPES = {'BatchMixup': BatchMixup, 'BatchCutMix': BatchCutMix, 'Identity': Identity}

    def __init__(self, augments_cfg):
        cfgs = [augments_cfg] if isinstance(augments_cfg, dict) else list(augments_cfg)
        self.augments = []
        for c in cfgs:
            c = dict(c)
            self.augments.append(self._TYPES[c.pop('type')](**c))
        self.aug_probs = [a.prob for a in self.augments]
        has_identity = any(isinstance(a, Identity) for a in self.augments)
        if has_identity:
            assert sum(self.aug_probs) == 1.0
        else:
            assert sum(self.aug_probs) <= 1.0
            identity_prob = 1 - sum(self.aug_probs)
            if identity_prob > 0:
                self.augments.append(Identity(self.augments[0].num_classes, identity_prob))
                self.aug_probs.append(identity_prob)

    def __call__(self, img, gt_label):
        if not self.augments:
            return img, gt_label
        if img.is_cuda or self.device_rng:
            return self.call_on_device(img, gt_label)
        aug = self.augments[int(np.random.choice(len(self.augments), p=self.aug_probs))]
        return aug(img, gt_label)

    device_rng = False       # CPU tensors: the host-RNG path above unless a test asks for the device formulation

    def call_on_device(self, img, gt_label):
        """The same draw (which augment, Beta(alpha, alpha) mixing factor, permutation, CutMix box) with the DEVICE
        generator and static shapes, so the call can sit inside a captured CUDA graph: a host-side np.random draw
        would be frozen into the graph and replayed forever.  One blend for all cases:
            img' = w * img + (1 - w) * img[perm],   label' = lam * onehot + (1 - lam) * onehot[perm]
        with w = lam (Mixup), w = 1 outside / 0 inside the box and lam = 1 - box area / image area (CutMix), w = lam = 1
        (no augment)."""
        from .bricks import const_tensor
        dev, B, H, W = img.device, img.shape[0], img.shape[-2], img.shape[-1]
        nc = self.augments[0].num_classes
        one_hot = one_hot_encoding(gt_label, nc)
        perm = torch.rand(B, device=dev).argsort()
        u = torch.rand((), device=dev)
        w = torch.ones((1, 1), device=dev)
        lam = torch.ones((), device=dev)
        lo = 0.0
        for aug, p in zip(self.augments, self.aug_probs):
            sel = (u >= lo) & (u < lo + p)
            lo += p
            if isinstance(aug, Identity):
                continue
            g = torch._standard_gamma(const_tensor([float(aug.alpha), float(aug.alpha)], torch.float32, dev))
            lam_a = g[0] / (g[0] + g[1])                                   # Beta(alpha, alpha)
            if isinstance(aug, BatchMixup):
                w_a = lam_a.view(1, 1)
            else:
                ratio = (1 - lam_a).sqrt()
                half_h, half_w = (H * ratio).floor().long() // 2, (W * ratio).floor().long() // 2
                c = torch.rand(2, device=dev)
                cy, cx = (c[0] * H).long().clamp(max=H - 1), (c[1] * W).long().clamp(max=W - 1)
                yl, yh = (cy - half_h).clamp(0, H), (cy + half_h).clamp(0, H)
                xl, xh = (cx - half_w).clamp(0, W), (cx + half_w).clamp(0, W)
                ys, xs = torch.arange(H, device=dev).view(H, 1), torch.arange(W, device=dev).view(1, W)
                inside = (ys >= yl) & (ys < yh) & (xs >= xl) & (xs < xh)
                w_a = 1.0 - inside.float()
                if aug.correct_lam:
                    lam_a = 1.0 - ((yh - yl) * (xh - xl)).float() / float(H * W)
            w = torch.where(sel, w_a, w)
            lam = torch.where(sel, lam_a, lam)
        w = w.to(img.dtype)
        img = w * img + (1 - w) * img[perm]
        return img, lam * one_hot + (1 - lam) * one_hot[perm]


_CLS_LOSSES = {'LabelSmoothLoss': LabelSmoothLoss, 'CrossEntropyLoss': CrossEntropyLoss}


@MODELS.register_module()
class SlvlClsHead(nn.Module):
    uses_neck = False          # reads the last backbone stage only (lets MTL skip the neck on cls iterations)

    def __init__(self, num_classes, in_channels, loss=dict(type='CrossEntropyLoss', loss_weight=1.0), topk=(1,),
                 cal_acc=False, init_cfg=dict(type='Normal', layer='Linear', std=0.01)):
        super().__init__()
        if num_classes <= 0:
            raise ValueError('num_classes=%d must be a positive integer' % num_classes)
        self.in_channels, self.num_classes, self.cal_acc, self.topk = in_channels, num_classes, cal_acc, topk
        loss = dict(loss)
        self.compute_loss = _CLS_LOSSES[loss.pop('type')](**loss)
        self.fc = nn.Linear(in_channels, num_classes)
        self.avg_pool = GlobalAveragePooling()
        nn.init.normal_(self.fc.weight, 0, 0.01)
        nn.init.constant_(self.fc.bias, 0)

    def pre_logits(self, x):
        cls_token = self.avg_pool((x[-1],))          # the reference pools every map and keeps the last
        if isinstance(cls_token, (tuple, list)):
            cls_token = cls_token[-1]
        return cls_token

    def loss(self, cls_score, gt_label):
        num_samples = len(cls_score)
        losses = dict()
        losses['loss'] = self.compute_loss(cls_score, gt_label, avg_factor=num_samples)
        if self.cal_acc:
            pred = cls_score.argmax(-1)
            losses['accuracy'] = {'top-1': (pred == gt_label).float().mean() * 100}
        return losses

    def forward_train(self, neck_feature, backbone_feature, gt_label, shared_encoder, **kwargs):
        cls_score = self.fc(self.pre_logits(backbone_feature))
        return self.loss(cls_score, gt_label)

    def simple_test(self, neck_feature, backbone_feature, shared_encoder, softmax=True, post_process=True):
        cls_score = self.fc(self.pre_logits(backbone_feature))
        pred = F.softmax(cls_score.float(), dim=1) if softmax else cls_score
        if post_process:
            return list(pred.detach().cpu().numpy())
        return pred


@MODELS.register_module()
class MlvlClsPixelDecoder(nn.Module):
    """reference models/multi/cls_head/pixel_decoder.py:14-116: the neck levels go through the SHARED encoder
    (low -> high resolution, sine + level encodings, normalised grid reference points) and come back as maps."""

    def __init__(self, num_encoder_levels=4, strides=[4, 8, 16, 32], feat_channels=256, num_outs=4,
                 positional_encoding=dict(type='SinePositionalEncoding', num_feats=128, normalize=True), init_cfg=None):
        super().__init__()
        from .bricks import build_positional_encoding
        self.strides, self.num_encoder_levels, self.num_outs = strides, num_encoder_levels, num_outs
        self.postional_encoding = build_positional_encoding(positional_encoding)
        self.level_encoding = nn.Embedding(num_encoder_levels, feat_channels)

    def init_weights(self):
        nn.init.normal_(self.level_encoding.weight, mean=0, std=1)

    def forward(self, encoder, neck_feats):
        from .seg_head import encode_neck_levels
        return encode_neck_levels(self, encoder, neck_feats, neck_feats[0].shape[0], len(neck_feats))


@MODELS.register_module()
class MlvlClsHead(SlvlClsHead):
    """reference models/multi/cls_head/mlvl_cls_head.py:13-126: classification token from the shared encoder's
    multi-level memory, by one of eight pooling `scheme`s (1/2: GAP of level 0/1; 3: mean over all tokens;
    4: mean of the per-level GAPs; 5/6: learned token weights on level 0/1; 7: learned weights over all tokens;
    8: learned weights over the per-level GAPs).  Levels are ordered low -> high resolution."""
    uses_neck = True
    _FEAT_LEN = {5: (4,), 6: (7,), 7: (4, 7, 14, 28)}

    def __init__(self, *args, pixel_decoder=None, scheme=5, **kwargs):
        init_cfg = kwargs.pop('init_cfg', None)
        super().__init__(*args, **kwargs)
        assert scheme in range(1, 9), 'scheme must be 1..8 (0 is the reference\'s self-test mode)'
        self.scheme = scheme
        self.pixel_decoder = pixel_decoder if isinstance(pixel_decoder, nn.Module) else MODELS.build(pixel_decoder)
        if scheme in (5, 6, 7):
            n = sum(x ** 2 for x in self._FEAT_LEN[scheme])
            self.out_proj = nn.Linear(n, 1)
        elif scheme == 8:
            n = self.pixel_decoder.num_encoder_levels
            self.out_proj = nn.Linear(n, 1)
        if hasattr(self, 'out_proj'):
            nn.init.constant_(self.out_proj.weight, 1.0 / n)        # starts as the plain mean
            nn.init.constant_(self.out_proj.bias, 0)
        self.pixel_decoder.init_weights()
        # mmcv BaseModule.init_weights applies the head's init_cfg to EVERY matching layer below it, so with the
        # reference's `TruncNormal(layer='Linear')` entry (cfg MTL_swin-t...:66-69) fc and out_proj are both re-drawn
        from .bricks import apply_init_cfg
        apply_init_cfg(self, init_cfg)

    def pre_logits(self, mlvl_feats):
        s = self.scheme
        if s in (1, 2):
            return self.avg_pool(mlvl_feats[s - 1])
        if s in (5, 6):
            return self.out_proj(mlvl_feats[s - 5].flatten(2)).squeeze(-1)
        if s in (3, 7):
            seq = torch.cat([f.flatten(2) for f in mlvl_feats], dim=2)
            return seq.mean(dim=2) if s == 3 else self.out_proj(seq).squeeze(-1)
        tokens = self.avg_pool(tuple(mlvl_feats))
        if s == 4:
            return sum(tokens) / len(tokens)
        return self.out_proj(torch.stack(tokens, -1)).squeeze(-1)

    def forward_encoder(self, encoder, neck_feature, backbone_feature):
        return self.pixel_decoder(encoder, neck_feature)

    def forward_train(self, neck_feature, backbone_feature, gt_label, shared_encoder, **kwargs):
        x = self.forward_encoder(shared_encoder, neck_feature, backbone_feature)
        return self.loss(self.fc(self.pre_logits(x)), gt_label)

    def simple_test(self, neck_feature, backbone_feature, shared_encoder, softmax=True, post_process=True):
        x = self.forward_encoder(shared_encoder, neck_feature, backbone_feature)
        cls_score = self.fc(self.pre_logits(x))
        pred = F.softmax(cls_score.float(), dim=1) if softmax else cls_score
        return list(pred.detach().cpu().numpy()) if post_process else pred


MODELS.register_module()(GlobalAveragePooling)


@MODELS.register_module()
class LinearClsHead(SlvlClsHead):
    """mmcls LinearClsHead (reference configs/_base_/cls/swin-tiny.py:12-19): fc on the last element of the neck's
    output tuple (already pooled to (B, C))."""
    uses_neck = True

    def __init__(self, *args, init_cfg=None, **kwargs):
        super().__init__(*args, **kwargs)

    def pre_logits(self, x):
        return x[-1] if isinstance(x, (tuple, list)) else x

    def forward_train(self, x, gt_label, **kwargs):
        return self.loss(self.fc(self.pre_logits(x)), gt_label)

    def simple_test(self, x, softmax=True, post_process=True):
        cls_score = self.fc(self.pre_logits(x))
        pred = F.softmax(cls_score.float(), dim=1) if softmax else cls_score
        return list(pred.detach().cpu().numpy()) if post_process else pred
