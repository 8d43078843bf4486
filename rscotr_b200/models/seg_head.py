"""Mask2Former-style semantic head on the shared deformable encoder
(reference models/multi/seg_head/{mask2former_head,pixel_decoder}.py, mmseg 0.28
BaseDecodeHead.losses)."""
import copy

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..config import MODELS, build_from_cfg
from .bricks import (GeomCache, Linear, Conv2d, ConvModule, build_positional_encoding, build_transformer_layer_sequence,
                     const_tensor)


def resize(input, size=None, mode='bilinear', align_corners=False):
    """mmseg.ops.resize on the rsc_bilinear kernel (bilinear, align_corners=False)."""
    assert mode == 'bilinear' and not align_corners
    if tuple(input.shape[-2:]) == tuple(size):
        return input
    return ops.bilinear_resize(input, size)


def caffe2_xavier_init(conv, bias=0):
    nn.init.kaiming_uniform_(conv.weight, a=1, mode='fan_in', nonlinearity='leaky_relu')
    if conv.bias is not None:
        nn.init.constant_(conv.bias, bias)


def encode_neck_levels(self, encoder, neck_feats, batch_size, num_input_levels):
    """The part the seg and the multi-level cls pixel decoders share (seg_head/pixel_decoder.py:80-160,
    cls_head/pixel_decoder.py:40-116): neck levels from low to high resolution -> flattened tokens + sine / level
    positional encodings + normalised reference points -> shared encoder -> memory split back into (B,C,h,w) maps."""
    dev = neck_feats[0].device
    levels = [num_input_levels - i - 1 for i in range(self.num_encoder_levels)]
    shapes_py = [tuple(neck_feats[l].shape[-2:]) for l in levels]

    def make():
        """shape-only constants: all-false padding masks, sine encodings, reference points, level index"""
        padding_mask_list, pos_list, reference_points_list = [], [], []
        for i, level_idx in enumerate(levels):
            h, w = shapes_py[i]
            padding_mask_resized = torch.zeros((batch_size, h, w), dtype=torch.bool, device=dev)
            pos_list.append(self.postional_encoding(padding_mask_resized).flatten(2).permute(2, 0, 1))
            # MlvlPointGenerator.single_level_grid_priors / (w*stride, h*stride)
            sx = (torch.arange(w, device=dev, dtype=torch.float32) + 0.5) * self.strides[level_idx]
            sy = (torch.arange(h, device=dev, dtype=torch.float32) + 0.5) * self.strides[level_idx]
            yy, xx = torch.meshgrid(sy, sx, indexing='ij')
            reference_points = torch.stack([xx.reshape(-1), yy.reshape(-1)], -1)
            reference_points_list.append(reference_points / const_tensor(
                [[float(w * self.strides[level_idx]), float(h * self.strides[level_idx])]], torch.float32, dev))
            padding_mask_list.append(padding_mask_resized.flatten(1))
        spatial_shapes = const_tensor(shapes_py, torch.long, dev)
        reference_points = torch.cat(reference_points_list, dim=0)
        reference_points = reference_points[None, :, None].repeat(batch_size, 1, self.num_encoder_levels, 1)
        return dict(padding_masks=torch.cat(padding_mask_list, dim=1), pos=[p.contiguous() for p in pos_list],
                    spatial_shapes=spatial_shapes,
                    level_start_index=torch.cat((spatial_shapes.new_zeros((1,)),
                                                 spatial_shapes.prod(1).cumsum(0)[:-1])),
                    reference_points=reference_points,
                    valid_radios=reference_points.new_ones((batch_size, self.num_encoder_levels, 2)))
    if not hasattr(self, '_geom'):
        self._geom = GeomCache()
    geo = self._geom.get((batch_size, tuple(shapes_py), str(dev)), make)
    encoder_inputs = torch.cat([neck_feats[l].flatten(2).permute(2, 0, 1) for l in levels], dim=0)
    level_positional_encodings = torch.cat([p + self.level_encoding.weight[i].view(1, 1, -1)
                                            for i, p in enumerate(geo['pos'])], dim=0)
    padding_masks, spatial_shapes = geo['padding_masks'], geo['spatial_shapes']
    level_start_index, reference_points, valid_radios = geo['level_start_index'], geo['reference_points'], \
        geo['valid_radios']
    memory = encoder(query=encoder_inputs, key=None, value=None, query_pos=level_positional_encodings,
                     key_pos=None, attn_masks=None, key_padding_mask=None,
                     query_key_padding_mask=None,      # (the reference passes an all-False mask: same result)
                     spatial_shapes=spatial_shapes, reference_points=reference_points,
                     level_start_index=level_start_index, valid_radios=valid_radios)
    memory = memory.permute(1, 2, 0)
    num_query_per_level = [h * w for h, w in shapes_py]
    outs = torch.split(memory, num_query_per_level, dim=-1)
    outs = [x.reshape(batch_size, -1, shapes_py[i][0], shapes_py[i][1]) for i, x in enumerate(outs)]
    return outs


_MASK_BITS = __import__('os').environ.get('RSC_MASK_BITS', '1') != '0'       # rsc_m2f_mask_bits + rsc_attn_* (0: boolean mask + library SDPA)
_COMPACT_ATTN_MASK = __import__('os').environ.get('RSC_COMPACT_ATTN_MASK', '1') != '0'   # A/B'd on B200 in round 2 (+1.1 % it/s): promoted


@MODELS.register_module()
class MlvlSegPixelDecoder(nn.Module):
    def __init__(self, num_encoder_levels=4, in_channels=[256, 512, 1024, 2048], strides=[4, 8, 16, 32],
                 feat_channels=256, out_channels=256, num_outs=3, norm_cfg=dict(type='GN', num_groups=32),
                 act_cfg=dict(type='ReLU'), positional_encoding=dict(type='SinePositionalEncoding', num_feats=128,
                                                                     normalize=True), init_cfg=None):
        super().__init__()
        self.strides = strides
        self.num_input_levels = len(in_channels)
        self.num_encoder_levels = num_encoder_levels
        self.postional_encoding = build_positional_encoding(positional_encoding)
        self.level_encoding = nn.Embedding(self.num_encoder_levels, feat_channels)
        self.lateral_convs = nn.ModuleList()
        self.output_convs = nn.ModuleList()
        self.use_bias = norm_cfg is None
        for i in range(self.num_input_levels - self.num_encoder_levels - 1, -1, -1):
            self.lateral_convs.append(ConvModule(in_channels[i], feat_channels, 1, bias=self.use_bias,
                                                 norm_cfg=norm_cfg, act_cfg=None))
            self.output_convs.append(ConvModule(feat_channels, feat_channels, 3, stride=1, padding=1,
                                                bias=self.use_bias, norm_cfg=norm_cfg, act_cfg=act_cfg))
        self.mask_feature = Conv2d(feat_channels, out_channels, kernel_size=1, stride=1, padding=0)
        self.num_outs = num_outs

    def init_weights(self):
        for i in range(0, self.num_input_levels - self.num_encoder_levels):
            caffe2_xavier_init(self.lateral_convs[i].conv, bias=0)
            caffe2_xavier_init(self.output_convs[i].conv, bias=0)
        caffe2_xavier_init(self.mask_feature, bias=0)
        nn.init.normal_(self.level_encoding.weight, mean=0, std=1)

    def forward(self, encoder, neck_feats, backbone_feats):
        batch_size = backbone_feats[0].shape[0]
        outs = encode_neck_levels(self, encoder, neck_feats, batch_size, self.num_input_levels)
        for i in range(self.num_input_levels - self.num_encoder_levels - 1, -1, -1):
            x = backbone_feats[i]
            cur_feat = self.lateral_convs[i](x)
            y = cur_feat + resize(outs[-1], size=cur_feat.shape[-2:])
            outs.append(self.output_convs[i](y))
        multi_scale_features = outs[:self.num_outs]
        mask_feature = self.mask_feature(outs[-1])
        return mask_feature, multi_scale_features


@MODELS.register_module()
class Mask2FormerHead(nn.Module):
    def __init__(self, in_channels, feat_channels, out_channels, num_classes=5, num_queries=100,
                 num_transformer_feat_level=4, scheme=1, pixel_decoder=None, enforce_decoder_input_project=False,
                 transformer_decoder=None, positional_encoding=None, ignore_index=255,
                 loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0), align_corners=False,
                 init_cfg=None):
        super().__init__()
        self.scheme, self.num_classes, self.num_queries = scheme, num_classes, num_queries
        self.align_corners, self.ignore_index = align_corners, ignore_index
        self.num_transformer_feat_level = num_transformer_feat_level
        self.num_heads = transformer_decoder['transformerlayers']['attn_cfgs']['num_heads']
        self.num_transformer_decoder_layers = transformer_decoder['num_layers']
        pixel_decoder_ = copy.deepcopy(dict(pixel_decoder))
        pixel_decoder_.update(in_channels=in_channels, feat_channels=feat_channels, out_channels=out_channels)
        self.pixel_decoder = build_from_cfg(pixel_decoder_, MODELS)
        self.transformer_decoder = build_transformer_layer_sequence(transformer_decoder)
        self.decoder_embed_dims = self.transformer_decoder.embed_dims
        self.decoder_input_projs = nn.ModuleList()
        for _ in range(num_transformer_feat_level):
            if self.decoder_embed_dims != feat_channels or enforce_decoder_input_project:
                self.decoder_input_projs.append(Conv2d(feat_channels, self.decoder_embed_dims, kernel_size=1))
            else:
                self.decoder_input_projs.append(nn.Identity())
        self.decoder_positional_encoding = build_positional_encoding(positional_encoding)
        self.query_embed = nn.Embedding(self.num_queries, feat_channels)
        self.query_feat = nn.Embedding(self.num_queries, feat_channels)
        self.level_embed = nn.Embedding(self.num_transformer_feat_level, feat_channels)
        if self.scheme == 1:
            self.cls_embed = Linear(feat_channels, self.num_classes + 1)
        self.mask_embed = nn.Sequential(Linear(feat_channels, feat_channels), nn.ReLU(),
                                        Linear(feat_channels, feat_channels), nn.ReLU(),
                                        Linear(feat_channels, out_channels))
        assert loss_decode['type'] == 'CrossEntropyLoss' and not loss_decode.get('use_sigmoid', False)
        self.loss_weight = loss_decode.get('loss_weight', 1.0)
        self.init_weights()

    def init_weights(self):
        for m in self.decoder_input_projs:
            if isinstance(m, nn.Conv2d):
                caffe2_xavier_init(m, bias=0)
        self.pixel_decoder.init_weights()
        for p in self.transformer_decoder.parameters():
            if p.dim() > 1:
                nn.init.xavier_normal_(p)

    def forward_head(self, decoder_out, mask_feature, attn_mask_target_size, need_seg=True, need_mask=True):
        """mask2former_head.py:111-139.  -> (seg_mask, attn_mask).  On the bf16 CUDA path attn_mask is an ops.MaskBits
        (rsc_m2f_mask_bits: resize + threshold + the all-masked-row rule of :177-178 in one kernel, one bit per
        (image, query, key)); otherwise the reference's boolean tensor.  need_seg / need_mask: the decoder loop uses only
        the mask of the intermediate calls and only the segmentation of the last one (the reference computes and
        discards the rest)."""
        decoder_out = self.transformer_decoder.post_norm(decoder_out)
        decoder_out = decoder_out.transpose(0, 1)
        mask_embed = self.mask_embed(decoder_out)
        mask_pred = torch.einsum('bqd,bdhw->bqhw', mask_embed, mask_feature)
        seg_mask = None
        if not need_seg:
            pass
        elif self.scheme == 1:
            cls_pred = self.cls_embed(decoder_out)
            seg_mask = torch.einsum('bqc,bqhw->bchw', cls_pred, mask_pred)
        elif self.scheme == 2:
            seg_mask = mask_pred
        else:
            raise NotImplementedError
        if not need_mask:
            return seg_mask, None
        with torch.no_grad():
            if _MASK_BITS and ops.attention_supported(mask_pred, getattr(self, "decoder_embed_dims", 0), self.num_heads):
                return seg_mask, ops.m2f_attn_mask(mask_pred, attn_mask_target_size)
            attn_mask = resize(mask_pred.detach(), attn_mask_target_size)
            if _COMPACT_ATTN_MASK:
                # (RSC_COMPACT_ATTN_MASK=0 restores the reference's layout) the reference repeats the resized logits over the heads BEFORE the
                # sigmoid / compare (8x redundant element-wise work and an 8x larger boolean mask); the decision is the same
                # for every head, so keep one (B, 1, Q, K) mask and let the attention call broadcast it
                attn_mask = (attn_mask.flatten(2).float().sigmoid() < 0.5).unsqueeze(1)
            else:
                attn_mask = attn_mask.flatten(2).unsqueeze(1).repeat((1, self.num_heads, 1, 1)).flatten(0, 1)
                attn_mask = attn_mask.float().sigmoid() < 0.5
        return seg_mask, attn_mask

    def forward(self, encoder, neck_feats, backbone_feats, img_metas):
        batch_size = len(img_metas)
        mask_features, multi_scale_memorys = self.pixel_decoder(encoder, neck_feats, backbone_feats)
        decoder_inputs = []
        shapes = tuple(tuple(m.shape[-2:]) for m in multi_scale_memorys[:self.num_transformer_feat_level])
        dev = mask_features.device

        def make():
            return [self.decoder_positional_encoding(torch.zeros((batch_size,) + hw, dtype=torch.bool, device=dev))
                    .flatten(2).permute(2, 0, 1) for hw in shapes]
        if not hasattr(self, '_geom'):
            self._geom = GeomCache()
        decoder_positional_encodings = self._geom.get((batch_size, shapes, str(dev)), make)
        for i in range(self.num_transformer_feat_level):
            decoder_input = self.decoder_input_projs[i](multi_scale_memorys[i])
            decoder_input = decoder_input.flatten(2).permute(2, 0, 1)
            decoder_inputs.append(decoder_input + self.level_embed.weight[i].view(1, 1, -1))
        # (activations are in the compute dtype; fp32 embeddings / sine encodings would promote every q + pos add of
        # the 9 decoder layers to fp32 and force a cast in front of each GEMM: align the dtype once)
        adt = mask_features.dtype
        decoder_positional_encodings = [p.to(adt) for p in decoder_positional_encodings]
        query_feat = self.query_feat.weight.to(adt).unsqueeze(1).repeat((1, batch_size, 1))
        query_embed = self.query_embed.weight.to(adt).unsqueeze(1).repeat((1, batch_size, 1))
        with torch.no_grad():
            _, attn_mask = self.forward_head(query_feat, mask_features, multi_scale_memorys[0].shape[-2:], need_seg=False)
        last = self.num_transformer_decoder_layers - 1
        for i in range(self.num_transformer_decoder_layers):
            level_idx = i % self.num_transformer_feat_level
            if not isinstance(attn_mask, ops.MaskBits):
                # if a mask is all True (all background), set it all False (sync-free form of the reference's
                # attn_mask[torch.where(attn_mask.sum(-1) == attn_mask.shape[-1])] = False); the bit-mask kernel has
                # already applied the rule
                attn_mask = attn_mask & ~attn_mask.all(-1, keepdim=True)
            layer = self.transformer_decoder.layers[i]
            query_feat = layer(query=query_feat, key=decoder_inputs[level_idx], value=decoder_inputs[level_idx],
                               query_pos=query_embed, key_pos=decoder_positional_encodings[level_idx],
                               attn_masks=[attn_mask, None], query_key_padding_mask=None, key_padding_mask=None)
            if i < last:      # only the mask feeds the next layer (it is detached: no gradient through these calls)
                with torch.no_grad():
                    _, attn_mask = self.forward_head(
                        query_feat, mask_features,
                        multi_scale_memorys[(i + 1) % self.num_transformer_feat_level].shape[-2:], need_seg=False)
            else:
                mask_pred, _ = self.forward_head(query_feat, mask_features, None, need_mask=False)
        return mask_pred

    def losses(self, seg_logit, seg_label):
        """mmseg BaseDecodeHead.losses: bilinear resize to the label size, CE mean over all
        pixels (ignore_index contributes 0), top-1 accuracy over non-ignored pixels."""
        loss = dict()
        if ops.upsample_ce_supported(seg_logit, seg_label):
            # fused: the (B,C,H,W) up-sampled logits are never materialised (rsc_upsample_ce_{fwd,bwd})
            stats = ops.upsample_ce(seg_logit, seg_label.squeeze(1), self.ignore_index)
            loss['loss_ce'] = stats[0] * (self.loss_weight / seg_label.numel())
            loss['acc_seg'] = stats[1].detach() * (100.0 / stats[2].detach().clamp(min=1))
            return loss
        seg_logit = resize(seg_logit, size=seg_label.shape[2:])
        seg_label = seg_label.squeeze(1)
        ce = F.cross_entropy(seg_logit.float(), seg_label, reduction='none', ignore_index=self.ignore_index)
        loss['loss_ce'] = self.loss_weight * ce.mean()
        with torch.no_grad():
            valid = seg_label != self.ignore_index
            correct = (seg_logit.argmax(1) == seg_label) & valid
            loss['acc_seg'] = correct.sum().float() * (100.0 / valid.sum().clamp(min=1).float())
        return loss

    def forward_train(self, neck_feats, backbone_feats, img_metas, gt_semantic_seg, shared_encoder):
        seg_logits = self.forward(shared_encoder, neck_feats, backbone_feats, img_metas)
        return self.losses(seg_logits, gt_semantic_seg)

    def forward_test(self, neck_feats, backbone_feats, img_metas, shared_encoder):
        return self.forward(shared_encoder, neck_feats, backbone_feats, img_metas)
