"""DINO detection head that threads the SHARED deformable encoder through itself.

Mirrors (names, arguments, outputs, loss keys) the reference's
models/multi/bbox_head/{dino_head,transformer,query_denoising}.py and the
vendored mmdet heads in models/multi/bbox_head/mmdet_detr_head/ (SURVEY 8a rows
a13-a16), with these structural changes for a GPU-resident step:
  * on CUDA the 7 x B Hungarian problems (cost matrices + linear sum assignment) run in ONE kernel launch
    (rsc_det_match) and the 13 x 3 loss terms in fused kernels (rsc_det_loss_*): the det step has no host phase
    and is one CUDA graph (reference: scipy behind a device->host sync per (layer, image), detr_head.py:513).
    `fused_loss = False` keeps the reference's structure (ATen losses, scipy on the host with one batched
    device->host copy) -- used by the tests as the independent implementation the kernels are checked against;
  * no .cuda() hard-coding (query_denoising.py:125-180), no per-loss .item().
"""
import copy
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..config import MODELS, build_from_cfg
from .bricks import (GeomCache, LayerNorm, Linear, PackedLosses, TransformerLayerSequence, const_tensor, MultiScaleDeformableAttention, build_positional_encoding,
                     build_transformer_layer_sequence, inverse_sigmoid)


# ------------------------------------------------------------------ box utils
def bbox_cxcywh_to_xyxy(b):
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def bbox_xyxy_to_cxcywh(b):
    x1, y1, x2, y2 = b.unbind(-1)
    return torch.stack([(x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1], -1)


def giou(b1, b2, aligned, eps=1e-6):
    """mmdet bbox_overlaps(mode='giou'); b1 (..., N, 4), b2 (..., M, 4) xyxy."""
    a1 = (b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1])
    a2 = (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1])
    if aligned:
        lt = torch.max(b1[..., :2], b2[..., :2])
        rb = torch.min(b1[..., 2:], b2[..., 2:])
        union_base = a1 + a2
        elt = torch.min(b1[..., :2], b2[..., :2])
        erb = torch.max(b1[..., 2:], b2[..., 2:])
    else:
        lt = torch.max(b1[..., :, None, :2], b2[..., None, :, :2])
        rb = torch.min(b1[..., :, None, 2:], b2[..., None, :, 2:])
        union_base = a1[..., None] + a2[..., None, :]
        elt = torch.min(b1[..., :, None, :2], b2[..., None, :, :2])
        erb = torch.max(b1[..., :, None, 2:], b2[..., None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    overlap = wh[..., 0] * wh[..., 1]
    union = (union_base - overlap).clamp(min=eps)
    ious = overlap / union
    ewh = (erb - elt).clamp(min=0)
    earea = (ewh[..., 0] * ewh[..., 1]).clamp(min=eps)
    return ious - (earea - union) / earea


def build_MLP(input_dim, hidden_dim, output_dim, num_layers):
    assert num_layers > 1, 'num_layers should be greater than 1 but got %d' % num_layers
    h = [hidden_dim] * (num_layers - 1)
    layers = []
    for n, k in zip([input_dim] + h[:-1], h):
        layers.extend((Linear(n, k), nn.ReLU()))
    layers.append(Linear(hidden_dim, output_dim))
    return nn.Sequential(*layers)


# ------------------------------------------------------------------ decoder
@MODELS.register_module()
class DinoTransformerDecoder(TransformerLayerSequence):
    """reference: models/multi/bbox_head/transformer.py:31-131"""

    def __init__(self, *args, return_intermediate=False, **kwargs):
        super().__init__(*args, **kwargs)
        self.return_intermediate = return_intermediate
        self.ref_point_head = build_MLP(self.embed_dims * 2, self.embed_dims, self.embed_dims, 2)
        self.norm = LayerNorm(self.embed_dims)

    @staticmethod
    def gen_sineembed_for_position(pos_tensor):
        """transformer.py:43-76: per coordinate a 128-d embedding sin / cos (even / odd index) of
        2*pi*v / 10000^(2*(j//2)/128), concatenated in the order (y, x[, w, h]).  Three kernels instead of ~25:
        one gather for the order, one fused multiply-add with cached constants, one sin (cos(a) = sin(a + pi/2))."""
        n = pos_tensor.size(-1)
        if n not in (2, 4):
            raise ValueError('Unknown pos_tensor shape(-1):{}'.format(n))
        dev = pos_tensor.device
        j = torch.arange(128, dtype=torch.float64)
        inv = const_tensor((2 * math.pi / (10000 ** (2 * (j // 2) / 128))).tolist(), torch.float32, dev)
        phase = const_tensor(((j % 2) * (math.pi / 2)).tolist(), torch.float32, dev)
        order = const_tensor([1, 0] if n == 2 else [1, 0, 2, 3], torch.long, dev)
        p = pos_tensor.float().index_select(-1, order).unsqueeze(-1)           # (B, N, n, 1) in output order
        return torch.addcmul(phase, p, inv).sin().flatten(2)

    def forward(self, query, *args, reference_points=None, valid_ratios=None, reg_branches=None, **kwargs):
        output = query
        intermediate = []
        intermediate_reference_points = [reference_points]
        for lid, layer in enumerate(self.layers):
            if reference_points.shape[-1] == 4:
                reference_points_input = reference_points[:, :, None] * \
                    torch.cat([valid_ratios, valid_ratios], -1)[:, None]
            else:
                assert reference_points.shape[-1] == 2
                reference_points_input = reference_points[:, :, None] * valid_ratios[:, None]
            query_sine_embed = self.gen_sineembed_for_position(reference_points_input[:, :, 0, :])
            query_pos = self.ref_point_head(query_sine_embed).permute(1, 0, 2)
            output = layer(output, *args, query_pos=query_pos, reference_points=reference_points_input, **kwargs)
            output = output.permute(1, 0, 2)
            if reg_branches is not None:
                tmp = reg_branches[lid](output)
                assert reference_points.shape[-1] == 4
                if tmp.is_cuda and tmp.dtype in (torch.float32, torch.bfloat16):
                    new_reference_points = ops.box_refine(tmp, reference_points, 1e-3)     # one kernel (rsc_box_refine_*)
                else:
                    new_reference_points = (tmp.float() + inverse_sigmoid(reference_points, eps=1e-3)).sigmoid()
                reference_points = new_reference_points.detach()
            output = output.permute(1, 0, 2)
            if self.return_intermediate:
                intermediate.append(self.norm(output))
                intermediate_reference_points.append(new_reference_points)   # look forward twice
        if self.return_intermediate:
            return torch.stack(intermediate), torch.stack(intermediate_reference_points)
        return output, reference_points


@MODELS.register_module()
class DinoTransformer(nn.Module):
    """reference: models/multi/bbox_head/transformer.py:134-272 (encoder passed in)."""

    def __init__(self, decoder=None, as_two_stage=False, num_feature_levels=4, two_stage_num_proposals=300,
                 init_cfg=None):
        super().__init__()
        self.decoder = build_transformer_layer_sequence(decoder)
        self.as_two_stage = as_two_stage
        self.num_feature_levels = num_feature_levels
        self.two_stage_num_proposals = two_stage_num_proposals
        self.embed_dims = self.decoder.embed_dims
        self.level_embeds = nn.Parameter(torch.Tensor(self.num_feature_levels, self.embed_dims))
        self.enc_output = Linear(self.embed_dims, self.embed_dims)
        self.enc_output_norm = LayerNorm(self.embed_dims)
        self.query_embed = nn.Embedding(self.two_stage_num_proposals, self.embed_dims)

    def init_weights(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MultiScaleDeformableAttention):
                m.init_weights()
        nn.init.normal_(self.level_embeds)
        nn.init.normal_(self.query_embed.weight.data)

    @staticmethod
    def get_valid_ratio(mask):
        _, H, W = mask.shape
        valid_H = torch.sum(~mask[:, :, 0], 1)
        valid_W = torch.sum(~mask[:, 0, :], 1)
        return torch.stack([valid_W.float() / W, valid_H.float() / H], -1)

    @staticmethod
    def get_reference_points(spatial_shapes, valid_ratios, device):
        ref_list = []
        for lvl, (H, W) in enumerate(spatial_shapes):
            ref_y, ref_x = torch.meshgrid(torch.linspace(0.5, H - 0.5, H, dtype=torch.float32, device=device),
                                          torch.linspace(0.5, W - 0.5, W, dtype=torch.float32, device=device),
                                          indexing='ij')
            ref_y = ref_y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H)
            ref_x = ref_x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W)
            ref_list.append(torch.stack((ref_x, ref_y), -1))
        reference_points = torch.cat(ref_list, 1)
        return reference_points[:, :, None] * valid_ratios[:, None]

    @staticmethod
    def proposal_grid(memory_padding_mask, spatial_shapes):
        """the memory-independent half of gen_encoder_output_proposals: (output_proposals with inf at padded /
        invalid positions, drop mask (N,S,1) = padded | invalid)."""
        N = memory_padding_mask.shape[0]
        dev = memory_padding_mask.device
        proposals = []
        _cur = 0
        for lvl, (H, W) in enumerate(spatial_shapes):
            mask_flatten_ = memory_padding_mask[:, _cur:(_cur + H * W)].view(N, H, W, 1)
            valid_H = torch.sum(~mask_flatten_[:, :, 0, 0], 1)
            valid_W = torch.sum(~mask_flatten_[:, 0, :, 0], 1)
            grid_y, grid_x = torch.meshgrid(
                torch.linspace(0, H - 1, H, dtype=torch.float32, device=dev),
                torch.linspace(0, W - 1, W, dtype=torch.float32, device=dev), indexing='ij')
            grid = torch.cat([grid_x.unsqueeze(-1), grid_y.unsqueeze(-1)], -1)
            scale = torch.cat([valid_W.unsqueeze(-1), valid_H.unsqueeze(-1)], 1).view(N, 1, 1, 2)
            grid = (grid.unsqueeze(0).expand(N, -1, -1, -1) + 0.5) / scale
            wh = torch.ones_like(grid) * 0.05 * (2.0 ** lvl)
            proposals.append(torch.cat((grid, wh), -1).view(N, -1, 4))
            _cur += H * W
        output_proposals = torch.cat(proposals, 1)
        output_proposals_valid = ((output_proposals > 0.01) & (output_proposals < 0.99)).all(-1, keepdim=True)
        output_proposals = torch.log(output_proposals / (1 - output_proposals))
        output_proposals = output_proposals.masked_fill(memory_padding_mask.unsqueeze(-1), float('inf'))
        output_proposals = output_proposals.masked_fill(~output_proposals_valid, float('inf'))
        return output_proposals, memory_padding_mask.unsqueeze(-1) | ~output_proposals_valid

    def gen_encoder_output_proposals(self, memory, memory_padding_mask, spatial_shapes, grid=None):
        output_proposals, drop = grid if grid is not None else self.proposal_grid(memory_padding_mask, spatial_shapes)
        output_memory = memory.masked_fill(drop, float(0))
        output_memory = self.enc_output_norm(self.enc_output(output_memory))
        return output_memory, output_proposals

    def _geometry(self, mlvl_masks, mlvl_pos_embeds, shapes_py, device):
        """everything of forward() that depends only on the masks / shapes (cached by GeomCache)."""
        mask_flatten = torch.cat([m.flatten(1) for m in mlvl_masks], 1)
        pos_levels = [p.flatten(2).transpose(1, 2).contiguous() for p in mlvl_pos_embeds]
        spatial_shapes = const_tensor(shapes_py, torch.long, device)
        level_start_index = torch.cat((spatial_shapes.new_zeros((1,)), spatial_shapes.prod(1).cumsum(0)[:-1]))
        valid_ratios = torch.stack([self.get_valid_ratio(m) for m in mlvl_masks], 1)
        reference_points = self.get_reference_points(shapes_py, valid_ratios, device=device)
        grid = self.proposal_grid(mask_flatten, shapes_py)
        return dict(masks=list(mlvl_masks), mask_flatten=mask_flatten, pos_levels=pos_levels,
                    spatial_shapes=spatial_shapes, level_start_index=level_start_index, valid_ratios=valid_ratios,
                    reference_points=reference_points, grid=grid)

    def forward(self, mlvl_feats, mlvl_masks, query_embed, mlvl_pos_embeds, dn_label_query, dn_bbox_query, attn_mask,
                encoder, reg_branches=None, cls_branches=None, no_padding=False, **kwargs):
        """no_padding: the caller knows on the host that no image of the batch is padded (every mask is all-False):
        the key-padding masks are then not handed to the attention modules at all (mmcv zero-fills `value` under the
        mask with a clone + masked_fill per layer, forward and backward, even when the mask is empty)."""
        assert self.as_two_stage and query_embed is None, 'as_two_stage must be True for DINO'
        shapes_py = [tuple(f.shape[-2:]) for f in mlvl_feats]
        dev = mlvl_feats[0].device
        # the masks come out of the head's own cache, so their identity is a valid key (the entry keeps them alive)
        key = (tuple(id(m) for m in mlvl_masks), tuple(id(p) for p in mlvl_pos_embeds), tuple(shapes_py))
        if not hasattr(self, '_geom'):
            self._geom = GeomCache()
        geo = self._geom.get(key, lambda: dict(self._geometry(mlvl_masks, mlvl_pos_embeds, shapes_py, dev),
                                               pos=list(mlvl_pos_embeds)))
        feat_flatten = torch.cat([feat.flatten(2).transpose(1, 2) for feat in mlvl_feats], 1)
        mask_flatten, spatial_shapes = geo['mask_flatten'], geo['spatial_shapes']
        # (per-level add + cat: the backward is four small reductions; an index gather's backward is a slow
        # sorted index_put over all 13 294 tokens)
        lvl_pos_embed_flatten = torch.cat([p + self.level_embeds[l].view(1, 1, -1)
                                           for l, p in enumerate(geo['pos_levels'])], 1)
        level_start_index, valid_ratios = geo['level_start_index'], geo['valid_ratios']
        reference_points = geo['reference_points']

        feat_flatten = feat_flatten.permute(1, 0, 2)
        lvl_pos_embed_flatten = lvl_pos_embed_flatten.permute(1, 0, 2)
        attn_kpm = None if no_padding else mask_flatten
        memory = encoder(query=feat_flatten, key=None, value=None, query_pos=lvl_pos_embed_flatten,
                         query_key_padding_mask=attn_kpm, spatial_shapes=spatial_shapes,
                         reference_points=reference_points, level_start_index=level_start_index,
                         valid_ratios=valid_ratios, **kwargs)
        memory = memory.permute(1, 0, 2)
        bs, _, c = memory.shape

        output_memory, output_proposals = self.gen_encoder_output_proposals(memory, mask_flatten, shapes_py,
                                                                            grid=geo['grid'])
        enc_outputs_class = cls_branches[self.decoder.num_layers](output_memory)
        enc_outputs_coord_unact = reg_branches[self.decoder.num_layers](output_memory).float() + output_proposals
        cls_out_features = cls_branches[self.decoder.num_layers].out_features
        topk = self.two_stage_num_proposals
        topk_indices = torch.topk(enc_outputs_class.max(-1)[0], topk, dim=1)[1]
        topk_score = torch.gather(enc_outputs_class, 1, topk_indices.unsqueeze(-1).repeat(1, 1, cls_out_features))
        topk_coords_unact = torch.gather(enc_outputs_coord_unact, 1, topk_indices.unsqueeze(-1).repeat(1, 1, 4))
        topk_anchor = topk_coords_unact.sigmoid()
        topk_coords_unact = topk_coords_unact.detach()

        query = self.query_embed.weight[:, None, :].repeat(1, bs, 1).transpose(0, 1)
        if dn_label_query is not None:
            query = torch.cat([dn_label_query.to(query.dtype), query], dim=1)
        if dn_bbox_query is not None:
            reference_points = torch.cat([dn_bbox_query, topk_coords_unact], dim=1)
        else:
            reference_points = topk_coords_unact
        reference_points = reference_points.sigmoid()

        query = query.permute(1, 0, 2)
        memory = memory.permute(1, 0, 2)
        inter_states, inter_references = self.decoder(
            query=query, key=None, value=memory, attn_masks=attn_mask, key_padding_mask=attn_kpm,
            reference_points=reference_points, spatial_shapes=spatial_shapes, level_start_index=level_start_index,
            valid_ratios=valid_ratios, reg_branches=reg_branches, **kwargs)
        return inter_states, inter_references, topk_score, topk_anchor


# ------------------------------------------------------------------ CDN
class CdnQueryGenerator:
    """reference: models/multi/bbox_head/query_denoising.py:8-201 (device agnostic).
    `forced_noise` (dict with keys p, new_label, rand_sign, rand_part) replaces the
    RNG draws so parity tests can feed the same noise to the oracle."""

    def __init__(self, num_queries, hidden_dim, num_classes, noise_scale=dict(label=0.5, box=0.4),
                 group_cfg=dict(dynamic=True, num_groups=None, num_dn_queries=None)):
        self.num_queries, self.hidden_dim, self.num_classes = num_queries, hidden_dim, num_classes
        self.label_noise_scale = noise_scale['label']
        self.box_noise_scale = noise_scale['box']
        self.dynamic_dn_groups = group_cfg.get('dynamic', False)
        if self.dynamic_dn_groups:
            assert 'num_dn_queries' in group_cfg, 'num_dn_queries should be set when using dynamic dn groups'
            self.num_dn = group_cfg['num_dn_queries']
        else:
            assert 'num_groups' in group_cfg, 'num_groups should be set when using static dn groups'
            self.num_dn = group_cfg['num_groups']
        assert isinstance(self.num_dn, int) and self.num_dn >= 1, \
            'Expected the num in group_cfg to have type int. Found %s ' % type(self.num_dn)
        self.forced_noise = None

    def get_num_groups(self, group_queries=None):
        if self.dynamic_dn_groups:
            assert group_queries is not None, 'group_queries should be provided when using dynamic dn groups'
            num_groups = 1 if group_queries == 0 else self.num_dn // group_queries
        else:
            num_groups = self.num_dn
        return int(max(num_groups, 1))

    def __call__(self, gt_bboxes, gt_labels=None, label_enc=None, img_metas=None):
        if gt_labels is not None:
            assert len(gt_bboxes) == len(gt_labels), \
                'the length of provided gt_labels %d should be equal to that of gt_bboxes %d' % (
                    len(gt_labels), len(gt_bboxes))
        assert gt_labels is not None and label_enc is not None and img_metas is not None
        batch_size = len(gt_bboxes)
        device = gt_bboxes[0].device
        boxes_n = []
        for img_meta, bboxes in zip(img_metas, gt_bboxes):
            img_h, img_w, _ = img_meta['img_shape']
            factor = const_tensor([[img_w, img_h, img_w, img_h]], bboxes.dtype, bboxes.device)
            boxes_n.append(bbox_xyxy_to_cxcywh(bboxes) / factor)
        known_num = [int(l.numel()) for l in gt_labels]
        single_pad = int(max(known_num))
        num_groups = self.get_num_groups(single_pad)
        labels = torch.cat(gt_labels)
        boxes = torch.cat(boxes_n)
        nbox = len(boxes)
        pad_size = int(single_pad * 2 * num_groups)
        tgt_size = pad_size + self.num_queries

        def make():
            """everything that depends only on the numbers of boxes per image (index maps, the block attention mask)"""
            batch_idx = torch.cat([torch.full((n,), i, dtype=torch.long, device=device) for i, n in enumerate(known_num)])
            known_bid = batch_idx.repeat(2 * num_groups, 1).view(-1)
            positive_idx = torch.arange(nbox, device=device).unsqueeze(0).repeat(num_groups, 1)
            positive_idx = positive_idx + (torch.arange(num_groups, device=device) * nbox * 2).unsqueeze(1)
            positive_idx = positive_idx.flatten()
            negative_idx = positive_idx + nbox
            neg_add = torch.zeros(2 * num_groups * nbox, 1, device=device)
            neg_add[negative_idx] = 1.0
            map_known_indice = None
            if nbox:
                map_known_indice = torch.cat([torch.arange(num, device=device) for num in known_num])
                map_known_indice = torch.cat([map_known_indice + single_pad * i for i in range(2 * num_groups)]).long()
            attn_mask = torch.zeros(tgt_size, tgt_size, dtype=torch.bool, device=device)
            attn_mask[pad_size:, :pad_size] = True
            for i in range(num_groups):
                lo, hi = single_pad * 2 * i, single_pad * 2 * (i + 1)
                attn_mask[lo:hi, hi:pad_size] = True
                attn_mask[lo:hi, :lo] = True
            attn_mask._rsc_add = {}        # (marks a shape-only constant: bricks._attend may cache its additive form)
            return dict(known_bid=known_bid, neg_add=neg_add, map_known_indice=map_known_indice, attn_mask=attn_mask)
        if not hasattr(self, '_geom'):
            self._geom = GeomCache()
        st = self._geom.get((tuple(known_num), str(device)), make)
        known_labels = labels.repeat(2 * num_groups, 1).view(-1)
        known_bboxs = boxes.repeat(2 * num_groups, 1)
        known_labels_expand = known_labels
        known_bbox_expand = known_bboxs
        fn = self.forced_noise or {}
        if self.label_noise_scale > 0:
            p = fn['p'].to(device) if 'p' in fn else torch.rand_like(known_labels_expand.float())
            new_label = fn['new_label'].to(device) if 'new_label' in fn else \
                torch.randint_like(known_labels_expand, 0, self.num_classes)
            # (reference draws new labels only for the chosen indices; drawing one per
            # slot and selecting is the same distribution and needs no host sync)
            known_labels_expand = torch.where(p < (self.label_noise_scale * 0.5), new_label, known_labels_expand)
        if self.box_noise_scale > 0:
            half = known_bboxs[:, 2:] / 2
            known_bbox_ = torch.cat([known_bboxs[:, :2] - half, known_bboxs[:, :2] + half], 1)     # xyxy
            diff = torch.cat([half, half], 1)
            rand_sign = fn['rand_sign'].to(device) if 'rand_sign' in fn else \
                torch.randint_like(known_bboxs, low=0, high=2, dtype=torch.float32)
            rand_sign = rand_sign * 2.0 - 1.0
            rand_part = (fn['rand_part'].to(device) if 'rand_part' in fn else torch.rand_like(known_bboxs))
            rand_part = (rand_part + st['neg_add']) * rand_sign            # negatives: noise in [1, 2)
            known_bbox_ = known_bbox_ + torch.mul(rand_part, diff) * self.box_noise_scale
            known_bbox_ = known_bbox_.clamp(min=0.0, max=1.0)
            known_bbox_expand = torch.cat([(known_bbox_[:, :2] + known_bbox_[:, 2:]) / 2,
                                           known_bbox_[:, 2:] - known_bbox_[:, :2]], 1)
        input_label_embed = label_enc(known_labels_expand.long())
        input_bbox_embed = inverse_sigmoid(known_bbox_expand, eps=1e-3)
        input_query_label = input_label_embed.new_zeros(batch_size, pad_size, self.hidden_dim)
        input_query_bbox = input_bbox_embed.new_zeros(batch_size, pad_size, 4)
        if nbox:
            input_query_label[(st['known_bid'], st['map_known_indice'])] = input_label_embed
            input_query_bbox[(st['known_bid'], st['map_known_indice'])] = input_bbox_embed
        attn_mask = st['attn_mask']
        dn_meta = {'pad_size': pad_size, 'num_dn_group': num_groups}
        return input_query_label, input_query_bbox, attn_mask, dn_meta


def build_dn_generator(dn_args):
    if dn_args is None:
        return None
    dn_args = dict(dn_args)
    t = dn_args.pop('type')
    if t == 'CdnQueryGenerator':
        return CdnQueryGenerator(**dn_args)
    raise NotImplementedError('%s is not supported yet' % t)


# ------------------------------------------------------------------ assigner / losses
class HungarianAssigner:
    """mmdet HungarianAssigner (FocalLossCost + BBoxL1Cost(xywh) + IoUCost(giou)), batched
    over the leading (layer) dimension; scipy solves on the host."""

    def __init__(self, cls_cost=dict(type='FocalLossCost', weight=2.0),
                 reg_cost=dict(type='BBoxL1Cost', weight=5.0, box_format='xywh'),
                 iou_cost=dict(type='IoUCost', iou_mode='giou', weight=2.0), **kwargs):
        assert cls_cost['type'] == 'FocalLossCost' and reg_cost['type'] == 'BBoxL1Cost' and iou_cost['type'] == 'IoUCost'
        assert reg_cost.get('box_format', 'xyxy') == 'xywh' and iou_cost.get('iou_mode', 'giou') == 'giou'
        self.w_cls, self.w_reg, self.w_iou = cls_cost.get('weight', 1.), reg_cost.get('weight', 1.), iou_cost.get('weight', 1.)
        self.alpha, self.gamma, self.eps = cls_cost.get('alpha', 0.25), cls_cost.get('gamma', 2), cls_cost.get('eps', 1e-12)

    def cost(self, bbox_pred, cls_pred, gt_bboxes, gt_labels, img_shape):
        """bbox_pred (..., Nq, 4) cxcywh normalised, cls_pred (..., Nq, C) logits -> (..., Nq, n_gt)."""
        img_h, img_w = img_shape[:2]
        factor = const_tensor([[img_w, img_h, img_w, img_h]], gt_bboxes.dtype, gt_bboxes.device)
        p = cls_pred.float().sigmoid()
        neg = -(1 - p + self.eps).log() * (1 - self.alpha) * p.pow(self.gamma)
        pos = -(p + self.eps).log() * self.alpha * (1 - p).pow(self.gamma)
        cls_cost = (pos[..., gt_labels] - neg[..., gt_labels]) * self.w_cls
        gt_n = bbox_xyxy_to_cxcywh(gt_bboxes / factor)
        bp = bbox_pred.float()
        reg_cost = (bp[..., :, None, :] - gt_n).abs().sum(-1) * self.w_reg
        iou_cost = -giou(bbox_cxcywh_to_xyxy(bp) * factor, gt_bboxes, aligned=False) * self.w_iou
        return cls_cost + reg_cost + iou_cost

    @staticmethod
    def solve(cost_np):
        from scipy.optimize import linear_sum_assignment
        return linear_sum_assignment(cost_np)


class FocalLoss(nn.Module):
    def __init__(self, use_sigmoid=True, gamma=2.0, alpha=0.25, reduction='mean', loss_weight=1.0, activated=False):
        super().__init__()
        assert use_sigmoid and not activated and reduction == 'mean'
        self.use_sigmoid, self.gamma, self.alpha, self.loss_weight = use_sigmoid, gamma, alpha, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None):
        loss = ops.sigmoid_focal_loss(pred.contiguous(), target.contiguous(), self.gamma, self.alpha)
        if weight is not None:
            loss = loss * weight.view(-1, 1)
        loss = loss.sum() / avg_factor if avg_factor is not None else loss.mean()
        return self.loss_weight * loss


class L1Loss(nn.Module):
    def __init__(self, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None):
        if target.numel() == 0:
            return pred.sum() * 0
        loss = (pred - target).abs()
        if weight is not None:
            loss = loss * weight
        loss = loss.sum() / avg_factor if avg_factor is not None else loss.mean()
        return self.loss_weight * loss


class GIoULoss(nn.Module):
    def __init__(self, eps=1e-6, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.eps, self.loss_weight = eps, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, any_positive=True):
        if weight is not None and not any_positive:
            return (pred * weight).sum()
        if weight is not None and weight.dim() > 1:
            weight = weight.mean(-1)
        loss = 1 - giou(pred, target, aligned=True, eps=self.eps)
        if weight is not None:
            loss = loss * weight
        loss = loss.sum() / avg_factor if avg_factor is not None else loss.mean()
        return self.loss_weight * loss


_LOSSES = {'FocalLoss': FocalLoss, 'L1Loss': L1Loss, 'GIoULoss': GIoULoss}


def build_loss(cfg):
    cfg = dict(cfg)
    return _LOSSES[cfg.pop('type')](**cfg)


def reduce_mean(t):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return t
    t = t.clone()
    dist.all_reduce(t.div_(dist.get_world_size()), op=dist.ReduceOp.SUM)
    return t


# ------------------------------------------------------------------ head
@MODELS.register_module()
class DINOHead(nn.Module):
    """reference: models/multi/bbox_head/dino_head.py (+ DETRHead / DeformableDETRHead
    in models/multi/bbox_head/mmdet_detr_head/)."""

    def __init__(self, num_classes, in_channels, num_query=100, num_reg_fcs=2, transformer=None,
                 sync_cls_avg_factor=False, positional_encoding=dict(type='SinePositionalEncoding', num_feats=128,
                                                                     normalize=True),
                 loss_cls=None, loss_bbox=dict(type='L1Loss', loss_weight=5.0),
                 loss_iou=dict(type='GIoULoss', loss_weight=2.0), train_cfg=None, test_cfg=dict(max_per_img=100),
                 with_box_refine=False, as_two_stage=False, dn_cfg=None, num_feature_levels=4, init_cfg=None,
                 **kwargs):
        super().__init__()
        transformer = copy.deepcopy(dict(transformer))
        if 'two_stage_num_proposals' in transformer:
            assert transformer['two_stage_num_proposals'] == num_query, \
                'two_stage_num_proposals must be equal to num_query for DINO'
        else:
            transformer['two_stage_num_proposals'] = num_query
        self.with_box_refine, self.as_two_stage = with_box_refine, as_two_stage
        if as_two_stage:
            transformer['as_two_stage'] = as_two_stage
        transformer.setdefault('num_feature_levels', num_feature_levels)
        self.bg_cls_weight = 0
        self.sync_cls_avg_factor = sync_cls_avg_factor
        if train_cfg:
            assert 'assigner' in train_cfg, 'assigner should be provided when train_cfg is set.'
            a = dict(train_cfg['assigner'])
            assert a.pop('type') == 'HungarianAssigner'
            self.assigner = HungarianAssigner(**a)
        self.num_query, self.num_classes, self.in_channels, self.num_reg_fcs = num_query, num_classes, in_channels, num_reg_fcs
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.loss_cls, self.loss_bbox, self.loss_iou = build_loss(loss_cls), build_loss(loss_bbox), build_loss(loss_iou)
        self.cls_out_channels = num_classes if self.loss_cls.use_sigmoid else num_classes + 1
        self.positional_encoding = build_positional_encoding(positional_encoding)
        self.transformer = build_from_cfg(transformer, MODELS)
        self.embed_dims = self.transformer.embed_dims
        num_feats = positional_encoding['num_feats']
        assert num_feats * 2 == self.embed_dims, \
            'embed_dims should be exactly 2 times of num_feats. Found %d and %d.' % (self.embed_dims, num_feats)
        assert self.as_two_stage, 'as_two_stage must be True for DINO'
        assert self.with_box_refine, 'with_box_refine must be True for DINO'
        self._init_layers()
        if dn_cfg is not None:
            dn_cfg = dict(dn_cfg)
            dn_cfg['num_classes'], dn_cfg['num_queries'], dn_cfg['hidden_dim'] = num_classes, num_query, self.embed_dims
        self.dn_generator = build_dn_generator(dn_cfg)
        self.fused_loss = True        # CUDA: rsc_det_match + rsc_det_loss_{fwd,bwd} (no host phase); False = ATen + scipy
        self.init_weights()

    def _init_layers(self):
        fc_cls = Linear(self.embed_dims, self.cls_out_channels)
        reg_branch = []
        for _ in range(self.num_reg_fcs):
            reg_branch.append(Linear(self.embed_dims, self.embed_dims))
            reg_branch.append(nn.ReLU())
        reg_branch.append(Linear(self.embed_dims, 4))
        reg_branch = nn.Sequential(*reg_branch)
        num_pred = self.transformer.decoder.num_layers + 1
        self.cls_branches = nn.ModuleList([copy.deepcopy(fc_cls) for _ in range(num_pred)])
        self.reg_branches = nn.ModuleList([copy.deepcopy(reg_branch) for _ in range(num_pred)])
        self.label_embedding = nn.Embedding(self.num_classes, self.embed_dims)

    def init_weights(self):
        self.transformer.init_weights()
        bias_init = float(-math.log((1 - 0.01) / 0.01))
        for m in self.cls_branches:
            nn.init.constant_(m.bias, bias_init)
        for m in self.reg_branches:
            nn.init.constant_(m[-1].weight, 0)
            nn.init.constant_(m[-1].bias, 0)
        nn.init.constant_(self.reg_branches[0][-1].bias.data[2:], -2.0)
        for m in self.reg_branches:
            nn.init.constant_(m[-1].bias.data[2:], 0.0)

    # -- forward ---------------------------------------------------------
    def forward_train(self, mlvl_feats, img_metas, gt_bboxes, gt_labels=None, gt_bboxes_ignore=None,
                      shared_encoder=None, proposal_cfg=None, **kwargs):
        assert proposal_cfg is None, '"proposal_cfg" must be None'
        assert self.dn_generator is not None, '"dn_cfg" must be set'
        dn_label_query, dn_bbox_query, attn_mask, dn_meta = self.dn_generator(
            gt_bboxes, gt_labels, self.label_embedding, img_metas)
        outs = self(shared_encoder, mlvl_feats, img_metas, dn_label_query, dn_bbox_query, attn_mask)
        return self.loss(*outs, gt_bboxes, gt_labels, img_metas, dn_meta, gt_bboxes_ignore=gt_bboxes_ignore)

    def forward_train_begin(self, mlvl_feats, img_metas, gt_bboxes, gt_labels=None, gt_bboxes_ignore=None,
                            shared_encoder=None):
        """forward_train up to (not including) the host-side matching; finish with
        loss_assign(pend) + loss_finish(pend)."""
        dn_label_query, dn_bbox_query, attn_mask, dn_meta = self.dn_generator(
            gt_bboxes, gt_labels, self.label_embedding, img_metas)
        outs = self(shared_encoder, mlvl_feats, img_metas, dn_label_query, dn_bbox_query, attn_mask)
        return self.loss_prepare(*outs, gt_bboxes, gt_labels, img_metas, dn_meta, gt_bboxes_ignore)

    def forward(self, encoder, mlvl_feats, img_metas, dn_label_query=None, dn_bbox_query=None, attn_mask=None):
        batch_size = mlvl_feats[0].size(0)
        input_img_h, input_img_w = img_metas[0]['batch_input_shape']

        def make():
            img_masks = mlvl_feats[0].new_ones((batch_size, input_img_h, input_img_w), dtype=torch.float32)
            for img_id in range(batch_size):
                img_h, img_w, _ = img_metas[img_id]['img_shape']
                img_masks[img_id, :img_h, :img_w] = 0
            masks, pes = [], []
            for feat in mlvl_feats:
                masks.append(F.interpolate(img_masks[None], size=feat.shape[-2:]).to(torch.bool).squeeze(0))
                pes.append(self.positional_encoding(masks[-1]))
            no_pad = all(tuple(m['img_shape'][:2]) == (input_img_h, input_img_w) for m in img_metas)
            return masks, pes, no_pad
        key = (tuple(tuple(m['img_shape'][:2]) for m in img_metas), (input_img_h, input_img_w),
               tuple(tuple(f.shape[-2:]) for f in mlvl_feats), str(mlvl_feats[0].device))
        if not hasattr(self, '_geom'):
            self._geom = GeomCache()
        mlvl_masks, mlvl_positional_encodings, no_pad = self._geom.get(key, make)
        hs, inter_references, topk_score, topk_anchor = self.transformer(
            mlvl_feats, mlvl_masks, None, mlvl_positional_encodings, dn_label_query, dn_bbox_query, attn_mask, encoder,
            reg_branches=self.reg_branches if self.with_box_refine else None,
            cls_branches=self.cls_branches if self.as_two_stage else None, no_padding=no_pad)
        hs = hs.permute(0, 2, 1, 3)
        if dn_label_query is not None and dn_label_query.size(1) == 0:
            hs[0] += self.label_embedding.weight[0, 0] * 0.0
        outputs_classes, outputs_coords = [], []
        # (unbind once: indexing per level would put one zero-fill + copy + add per level into the backward)
        for lvl, (hs_l, ref_l) in enumerate(zip(hs.unbind(0), inter_references.unbind(0)[:hs.shape[0]])):
            outputs_class = self.cls_branches[lvl](hs_l)
            tmp = self.reg_branches[lvl](hs_l)
            outputs_classes.append(outputs_class)
            if ref_l.shape[-1] == 4 and tmp.is_cuda and tmp.dtype in (torch.float32, torch.bfloat16):
                outputs_coords.append(ops.box_refine(tmp, ref_l, 1e-3))      # sigmoid(tmp + inverse_sigmoid(ref)): one kernel
                continue
            reference = inverse_sigmoid(ref_l, eps=1e-3)
            tmp = tmp.float()
            if reference.shape[-1] == 4:
                tmp = tmp + reference
            else:
                assert reference.shape[-1] == 2
                tmp = torch.cat([tmp[..., :2] + reference, tmp[..., 2:]], -1)
            outputs_coords.append(tmp.sigmoid())
        return torch.stack(outputs_classes), torch.stack(outputs_coords), topk_score, topk_anchor

    # -- losses ----------------------------------------------------------
    @staticmethod
    def extract_dn_outputs(all_cls_scores, all_bbox_preds, dn_meta):
        if dn_meta is not None:
            ps = dn_meta['pad_size']
            return (all_cls_scores[:, :, ps:, :], all_bbox_preds[:, :, ps:, :], all_cls_scores[:, :, :ps, :],
                    all_bbox_preds[:, :, :ps, :])
        return all_cls_scores, all_bbox_preds, None, None

    def loss(self, all_cls_scores, all_bbox_preds, enc_topk_scores, enc_topk_anchors, gt_bboxes_list, gt_labels_list,
             img_metas, dn_meta=None, gt_bboxes_ignore=None):
        """dino_head.py:152-234.  On CUDA: Hungarian matching and all 13 x 3 loss terms in a handful of
        kernels (loss_fused).  Otherwise in three phases so that a step engine can put the device work
        of phases 1 and 3 into CUDA graphs around the host-side matching of phase 2."""
        if self.fused_loss and all_cls_scores.is_cuda:
            return self.loss_fused(all_cls_scores, all_bbox_preds, enc_topk_scores, enc_topk_anchors, gt_bboxes_list,
                                   gt_labels_list, img_metas, dn_meta, gt_bboxes_ignore)
        pend = self.loss_prepare(all_cls_scores, all_bbox_preds, enc_topk_scores, enc_topk_anchors, gt_bboxes_list,
                                 gt_labels_list, img_metas, dn_meta, gt_bboxes_ignore)
        self.loss_assign(pend)
        return self.loss_finish(pend)

    def _avg_factors_dev(self, pos_counts, neg_counts, dev):
        """device version of _avg_factors: (2, n) tensor [cls_avg_factor; num_total_pos], both >= 1; the
        cross-rank means are ONE packed all-reduce on the device (no .item(), CUDA-graph capturable)."""
        pos = [float(x) for x in pos_counts]
        cls_avg = [p * 1.0 + float(n) * self.bg_cls_weight for p, n in zip(pos, neg_counts)]
        t = const_tensor([cls_avg, pos], torch.float32, dev)
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            r = t.clone()
            dist.all_reduce(r)
            r = r / dist.get_world_size()
            t = r if self.sync_cls_avg_factor else torch.stack([t[0], r[1]])
        return t.clamp(min=1.0)

    @staticmethod
    def dn_assign_table(sizes, num_groups, pad_size):
        """fixed targets of the denoising queries (dino_head.py:323-365 _get_dn_target_single) as a (B, pad_size)
        table: entry = global index (into the concatenated gt lists) of the gt box the query reconstructs, -1 for
        the negative / padding queries.  Group g occupies [g*single, (g+1)*single) with single = pad_size // groups;
        its first n_b slots are the positives of image b's n_b boxes."""
        single = pad_size // num_groups if num_groups else 0
        starts = [0]
        for n in sizes:
            starts.append(starts[-1] + n)
        table = [[-1] * pad_size for _ in sizes]
        for b, n in enumerate(sizes):
            for g in range(num_groups):
                for t in range(n):
                    table[b][g * single + t] = starts[b] + t
        return table

    def loss_fused(self, all_cls_scores, all_bbox_preds, enc_topk_scores, enc_topk_anchors, gt_bboxes_list,
                   gt_labels_list, img_metas, dn_meta=None, gt_bboxes_ignore=None):
        """Same loss terms / keys as loss_prepare + loss_assign + loss_finish, on the GPU end to end:
        rsc_det_match (cost matrices + linear sum assignment of all 7 x B problems), then one fused
        focal + L1 + GIoU kernel per query segment (encoder proposals, decoder layers, denoising part)."""
        assert gt_bboxes_ignore is None, \
            '%s only supports for gt_bboxes_ignore setting to None.' % self.__class__.__name__
        L, B, NqTot, C = all_cls_scores.shape
        dev = all_cls_scores.device
        ps = int(dn_meta['pad_size']) if dn_meta is not None else 0
        Nq = NqTot - ps
        sizes = [int(l.numel()) for l in gt_labels_list]
        max_gt, total_gt = (max(sizes) if sizes else 0), sum(sizes)
        starts = [0]
        for n in sizes:
            starts.append(starts[-1] + n)
        gt_start = const_tensor(starts, torch.int32, dev)
        img_wh = const_tensor([[float(m['img_shape'][1]), float(m['img_shape'][0])] for m in img_metas],
                              torch.float32, dev)
        gt_labels = torch.cat(gt_labels_list) if len(gt_labels_list) > 1 else gt_labels_list[0]
        gt_boxes = torch.cat(gt_bboxes_list) if len(gt_bboxes_list) > 1 else gt_bboxes_list[0]
        all_cls = all_cls_scores.contiguous()
        all_box = all_bbox_preds.float().contiguous()
        a = self.assigner
        has_enc = enc_topk_scores is not None
        with torch.no_grad():
            assign_dec, _, gt_norm = ops.det_match(all_cls, all_box, ps, Nq, gt_labels, gt_boxes, gt_start, img_wh,
                                                   max_gt, a.w_cls, a.w_reg, a.w_iou, a.alpha, a.gamma, a.eps)
            if has_enc:
                enc_cls, enc_box = enc_topk_scores.contiguous(), enc_topk_anchors.float().contiguous()
                assign_enc, _, _ = ops.det_match(enc_cls, enc_box, 0, enc_cls.shape[-2], gt_labels, gt_boxes, gt_start,
                                                 img_wh, max_gt, a.w_cls, a.w_reg, a.w_iou, a.alpha, a.gamma, a.eps,
                                                 want_gt_norm=False)
        # averaging factors: every gt is matched (n_gt <= Nq), so the positive counts are known on the host
        rows_match = (1 if has_enc else 0) + L
        pos_counts, neg_counts = [total_gt] * rows_match, [B * Nq - total_gt] * rows_match
        keys, segs, tensors, row = [], [], [all_cls, all_box], 0
        terms = ('loss_cls', 'loss_bbox', 'loss_iou')
        if has_enc:
            tensors += [enc_cls, enc_box]
            segs.append(dict(t=1, L=1, B=B, NqTot=enc_cls.shape[-2], q0=0, Nq=enc_cls.shape[-2], C=C, assign=assign_enc,
                             a_ls=B * enc_cls.shape[-2], row0=0, last_first=False, f0=0))
            keys += ['interm_' + t for t in terms]
            row = 1
        segs.append(dict(t=0, L=L, B=B, NqTot=NqTot, q0=ps, Nq=Nq, C=C, assign=assign_dec, a_ls=B * Nq, row0=row,
                         last_first=True, f0=row))
        keys += list(terms) + ['d%d.%s' % (n, t) for n in range(L - 1) for t in terms]
        row += L
        if dn_meta is not None:
            num_groups = int(dn_meta['num_dn_group'])
            dn_assign = self.dn_assign_table(sizes, num_groups, ps)
            assign_dn = const_tensor(dn_assign, torch.int32, dev) if ps else assign_dec
            npos = total_gt * num_groups
            pos_counts += [npos] * L
            neg_counts += [npos] * L
            segs.append(dict(t=0, L=L, B=B, NqTot=NqTot, q0=0, Nq=ps, C=C, assign=assign_dn, a_ls=0, row0=row,
                             last_first=True, f0=row))
            keys += ['dn_' + t for t in terms] + ['d%d.dn_%s' % (n, t) for n in range(L - 1) for t in terms]
            row += L
        factors = self._avg_factors_dev(pos_counts, neg_counts, dev)
        lc = self.loss_cls
        out = ops.det_loss(segs, tensors, row, gt_labels, gt_norm, img_wh, factors[0], factors[1], lc.gamma, lc.alpha,
                           lc.loss_weight, self.loss_bbox.loss_weight, self.loss_iou.loss_weight, self.loss_iou.eps)
        return PackedLosses(keys, out.view(-1))

    def loss_prepare(self, all_cls_scores, all_bbox_preds, enc_topk_scores, enc_topk_anchors, gt_bboxes_list,
                     gt_labels_list, img_metas, dn_meta=None, gt_bboxes_ignore=None):
        """phase 1 (device, sync free): split dn / matching parts, Hungarian cost matrices of all
        (layer, image) pairs, target buffers initialised to 'background', fixed dn targets."""
        assert gt_bboxes_ignore is None, \
            '%s only supports for gt_bboxes_ignore setting to None.' % self.__class__.__name__
        all_cls_scores, all_bbox_preds = all_cls_scores.float(), all_bbox_preds.float()   # @force_fp32
        all_cls_scores, all_bbox_preds, dn_cls_scores, dn_bbox_preds = self.extract_dn_outputs(
            all_cls_scores, all_bbox_preds, dn_meta)
        # matching part: stack [encoder proposals, decoder layers] -> (1+L, B, Nq, .)
        stacks_cls, stacks_box = all_cls_scores, all_bbox_preds
        if enc_topk_scores is not None:
            stacks_cls = torch.cat([enc_topk_scores.float()[None], all_cls_scores], 0)
            stacks_box = torch.cat([enc_topk_anchors.float()[None], all_bbox_preds], 0)
        nl, B, Nq, _ = stacks_cls.shape
        dev = stacks_cls.device
        sizes = [int(l.numel()) for l in gt_labels_list]
        with torch.no_grad():
            costs = [self.assigner.cost(stacks_box[:, b], stacks_cls[:, b], gt_bboxes_list[b], gt_labels_list[b],
                                        img_metas[b]['img_shape']).reshape(-1) for b in range(B) if sizes[b]]
            cost_flat = torch.cat(costs) if costs else None
            labels = torch.full((nl * B * Nq,), self.num_classes, dtype=torch.long, device=dev)
            bbox_targets = torch.zeros(nl * B * Nq, 4, device=dev)
            bbox_weights = torch.zeros(nl * B * Nq, 4, device=dev)
            gt_cat_labels = torch.cat(gt_labels_list)
            factors = torch.cat([const_tensor([[img_metas[b]['img_shape'][1], img_metas[b]['img_shape'][0]] * 2],
                                              gt_bboxes_list[b].dtype, dev).expand(sizes[b], 4) for b in range(B)])
            gt_cat_boxes = bbox_xyxy_to_cxcywh(torch.cat(gt_bboxes_list) / factors)
            dn_t = None
            if dn_cls_scores is not None:
                dn_t = self.get_dn_target(dn_bbox_preds[0], gt_bboxes_list, gt_labels_list, img_metas, dn_meta)
        return dict(stacks_cls=stacks_cls, stacks_box=stacks_box, dn_cls=dn_cls_scores, dn_box=dn_bbox_preds,
                    has_enc=enc_topk_scores is not None, img_metas=img_metas, sizes=sizes, cost_flat=cost_flat,
                    labels=labels, bbox_targets=bbox_targets, bbox_weights=bbox_weights, gt_cat_labels=gt_cat_labels,
                    gt_cat_boxes=gt_cat_boxes, dn_t=dn_t)

    @torch.no_grad()
    def loss_assign(self, pend):
        """phase 2 (host): ONE device->host copy of all cost matrices, scipy per (layer, image)
        problem, ONE host->device copy of the index lists written into the target buffers
        (detr_head.py:475-543); then the averaging factors of every loss term."""
        nl, B, Nq, _ = pend['stacks_cls'].shape
        dev, sizes = pend['stacks_cls'].device, pend['sizes']
        flat = pend['cost_flat'].cpu().numpy() if pend['cost_flat'] is not None else np.zeros(0, np.float32)
        rows, gts, off = [], [], 0
        num_pos = np.zeros(nl, dtype=np.int64)
        gt_off = np.concatenate([[0], np.cumsum(sizes)[:-1]]) if sizes else np.zeros(0, np.int64)
        for b in range(B):
            n = sizes[b]
            if not n:
                continue
            cb = flat[off:off + nl * Nq * n].reshape(nl, Nq, n)
            off += nl * Nq * n
            for l in range(nl):
                r, c = HungarianAssigner.solve(cb[l])
                rows.append(((l * B + b) * Nq + r).astype(np.int64))      # flat index into (nl, B, Nq)
                gts.append((gt_off[b] + c).astype(np.int64))
                num_pos[l] += len(r)
        if rows:
            idx = torch.from_numpy(np.stack([np.concatenate(rows), np.concatenate(gts)]))
            idx = idx.pin_memory().to(dev, non_blocking=True) if dev.type == 'cuda' else idx
            rows_t, gidx = idx[0], idx[1]
            pend['labels'][rows_t] = pend['gt_cat_labels'][gidx]
            pend['bbox_targets'][rows_t] = pend['gt_cat_boxes'][gidx]
            pend['bbox_weights'][rows_t] = 1.0
        pos_counts = [int(x) for x in num_pos]
        neg_counts = [B * Nq - x for x in pos_counts]
        if pend['dn_t'] is not None:
            pos_counts.append(pend['dn_t']['num_pos'])
            neg_counts.append(pend['dn_t']['num_neg'])
        pend['num_pos'] = num_pos
        pend['cls_factors'], pend['pos_factors'] = self._avg_factors(pos_counts, neg_counts, pend['stacks_cls'])

    def loss_finish(self, pend):
        """phase 3 (device, sync free): the 7 matching losses and 6 denoising losses."""
        nl, B, Nq, _ = pend['stacks_cls'].shape
        img_metas = pend['img_metas']
        targets = dict(labels=pend['labels'].view(nl, B * Nq), bbox_targets=pend['bbox_targets'].view(nl, B * Nq, 4),
                       bbox_weights=pend['bbox_weights'].view(nl, B * Nq, 4), num_pos=pend['num_pos'])
        cls_factors, pos_factors = pend['cls_factors'], pend['pos_factors']
        loss_dict = dict()
        losses = [self.loss_single(pend['stacks_cls'][i], pend['stacks_box'][i], targets, i, img_metas, cls_factors[i],
                                   pos_factors[i]) for i in range(nl)]
        if pend['has_enc']:
            (loss_dict['interm_loss_cls'], loss_dict['interm_loss_bbox'], loss_dict['interm_loss_iou']) = losses[0]
            losses = losses[1:]
        loss_dict['loss_cls'], loss_dict['loss_bbox'], loss_dict['loss_iou'] = losses[-1]
        for n, (lc, lb, li) in enumerate(losses[:-1]):
            loss_dict['d%d.loss_cls' % n], loss_dict['d%d.loss_bbox' % n], loss_dict['d%d.loss_iou' % n] = lc, lb, li
        if pend['dn_cls'] is not None:
            dn_cls_scores, dn_bbox_preds, dn_t = pend['dn_cls'], pend['dn_box'], pend['dn_t']
            dn_losses = [self.loss_dn_single(dn_cls_scores[i], dn_bbox_preds[i], dn_t, img_metas, cls_factors[-1],
                                             pos_factors[-1]) for i in range(len(dn_cls_scores))]
            loss_dict['dn_loss_cls'], loss_dict['dn_loss_bbox'], loss_dict['dn_loss_iou'] = dn_losses[-1]
            for n, (lc, lb, li) in enumerate(dn_losses[:-1]):
                loss_dict['d%d.dn_loss_cls' % n], loss_dict['d%d.dn_loss_bbox' % n], loss_dict['d%d.dn_loss_iou' % n] = lc, lb, li
        return loss_dict

    def _avg_factors(self, pos_counts, neg_counts, like):
        """cls_avg_factor / num_total_pos of every loss_single call of this step
        (detr_head.py:372-391, dino_head.py:262-284) with ONE packed all-reduce instead of
        two reduce_mean + .item() per call."""
        pos = torch.tensor([float(x) for x in pos_counts], dtype=torch.float64)
        neg = torch.tensor([float(x) for x in neg_counts], dtype=torch.float64)
        cls_avg = pos * 1.0 + neg * self.bg_cls_weight
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            packed = torch.cat([cls_avg, pos]).to(like.device, torch.float32)
            dist.all_reduce(packed)
            packed = (packed / dist.get_world_size()).cpu().double()
            mean_cls_avg, pos = packed[:len(pos)], packed[len(pos):]
            if self.sync_cls_avg_factor:
                cls_avg = mean_cls_avg
        return [max(float(c), 1) for c in cls_avg], [max(float(p), 1.0) for p in pos]

    def _box_losses(self, bbox_preds, bbox_targets, bbox_weights, img_metas, num_total_pos, any_pos):
        factors = torch.cat([const_tensor([[m['img_shape'][1], m['img_shape'][0]] * 2], bbox_preds.dtype,
                                          bbox_preds.device).expand(bbox_preds.size(1), 4) for m in img_metas], 0)
        bp = bbox_preds.reshape(-1, 4)
        bboxes = bbox_cxcywh_to_xyxy(bp) * factors
        bboxes_gt = bbox_cxcywh_to_xyxy(bbox_targets) * factors
        loss_iou = self.loss_iou(bboxes, bboxes_gt, bbox_weights, avg_factor=num_total_pos, any_positive=any_pos)
        loss_bbox = self.loss_bbox(bp, bbox_targets, bbox_weights, avg_factor=num_total_pos)
        return loss_bbox, loss_iou

    def loss_single(self, cls_scores, bbox_preds, targets, i, img_metas, cls_avg_factor, num_total_pos):
        """detr_head.py:333-416 with the targets / averaging factors of layer i precomputed."""
        npos = int(targets['num_pos'][i])
        labels = targets['labels'][i]
        loss_cls = self.loss_cls(cls_scores.reshape(-1, self.cls_out_channels), labels, labels.new_ones(
            labels.shape, dtype=torch.float32), avg_factor=cls_avg_factor)
        loss_bbox, loss_iou = self._box_losses(bbox_preds, targets['bbox_targets'][i], targets['bbox_weights'][i],
                                               img_metas, num_total_pos, npos > 0)
        return loss_cls, loss_bbox, loss_iou

    @torch.no_grad()
    def get_dn_target(self, dn_bbox_pred, gt_bboxes_list, gt_labels_list, img_metas, dn_meta):
        """dino_head.py:311-365: fixed (non-Hungarian) targets of the denoising queries."""
        num_groups, pad_size = dn_meta['num_dn_group'], dn_meta['pad_size']
        assert pad_size % num_groups == 0
        single_pad = pad_size // num_groups
        B, num_bboxes, _ = dn_bbox_pred.shape
        dev = dn_bbox_pred.device
        labels = torch.full((B, num_bboxes), self.num_classes, dtype=torch.long, device=dev)
        bbox_targets = torch.zeros(B, num_bboxes, 4, device=dev)
        bbox_weights = torch.zeros(B, num_bboxes, 4, device=dev)
        npos = nneg = 0
        for b in range(B):
            n = int(gt_labels_list[b].numel())
            if n == 0:
                continue
            t = torch.arange(n, device=dev).unsqueeze(0).repeat(num_groups, 1)
            pos_inds = ((torch.arange(num_groups, device=dev) * single_pad).unsqueeze(1) + t).flatten()
            labels[b, pos_inds] = gt_labels_list[b][t.flatten()]
            bbox_weights[b].index_fill_(0, pos_inds, 1.0)       # (scalar index_put would stage a host tensor)
            img_h, img_w, _ = img_metas[b]['img_shape']
            factor = const_tensor([[img_w, img_h, img_w, img_h]], dn_bbox_pred.dtype, dn_bbox_pred.device)
            bbox_targets[b, pos_inds] = bbox_xyxy_to_cxcywh(gt_bboxes_list[b] / factor).repeat([num_groups, 1])
            npos += n * num_groups
            nneg += n * num_groups
        return dict(labels=labels.view(-1), bbox_targets=bbox_targets.view(-1, 4), bbox_weights=bbox_weights.view(-1, 4),
                    num_pos=npos, num_neg=nneg)

    def loss_dn_single(self, dn_cls_scores, dn_bbox_preds, t, img_metas, cls_avg_factor, num_total_pos):
        cls_scores = dn_cls_scores.reshape(-1, self.cls_out_channels)
        if len(cls_scores) > 0:
            loss_cls = self.loss_cls(cls_scores, t['labels'], t['labels'].new_ones(t['labels'].shape, dtype=torch.float32),
                                     avg_factor=cls_avg_factor)
        else:
            loss_cls = torch.zeros(1, dtype=cls_scores.dtype, device=cls_scores.device)
        loss_bbox, loss_iou = self._box_losses(dn_bbox_preds, t['bbox_targets'], t['bbox_weights'], img_metas,
                                               num_total_pos, t['num_pos'] > 0)
        return loss_cls, loss_bbox, loss_iou

    # -- test ------------------------------------------------------------
    def simple_test(self, feats, img_metas, shared_encoder=None, rescale=False):
        outs = self.forward(shared_encoder, feats, img_metas)
        return self.get_bboxes(*outs, img_metas, rescale=rescale)

    def get_bboxes(self, all_cls_scores, all_bbox_preds, enc_cls_scores, enc_bbox_preds, img_metas, rescale=False):
        """detr_head.py:_get_bboxes_single (sigmoid branch): top max_per_img over Nq*C."""
        cls_scores, bbox_preds = all_cls_scores[-1].float(), all_bbox_preds[-1].float()
        results = []
        max_per_img = self.test_cfg.get('max_per_img', self.num_query)
        for img_id in range(len(img_metas)):
            cls_score, bbox_pred = cls_scores[img_id].sigmoid(), bbox_preds[img_id]
            scores, indexes = cls_score.view(-1).topk(max_per_img)
            det_labels = indexes % self.num_classes
            bbox_pred = bbox_pred[indexes // self.num_classes]
            img_shape = img_metas[img_id]['img_shape']
            det_bboxes = bbox_cxcywh_to_xyxy(bbox_pred)
            det_bboxes[:, 0::2] = (det_bboxes[:, 0::2] * img_shape[1]).clamp(min=0, max=img_shape[1])
            det_bboxes[:, 1::2] = (det_bboxes[:, 1::2] * img_shape[0]).clamp(min=0, max=img_shape[0])
            if rescale:
                det_bboxes = det_bboxes / det_bboxes.new_tensor(img_metas[img_id]['scale_factor'])
            results.append((torch.cat((det_bboxes, scores.unsqueeze(1)), -1), det_labels))
        return results
