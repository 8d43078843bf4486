"""Oracle: neck, the three task heads, their losses and the MTL train step (TEST INFRA).

CPU / fp32 / eager restatement, keyed on the reference state-dict layout, of
  models/multi/multitask_learner.py:81-147, 229-306      (MTL dispatch, _parse_losses)
  models/multi/cls_head/slvl_cls_head.py:9-28             (+ mmcls LinearClsHead / LabelSmoothLoss)
  models/multi/bbox_head/dino_head.py:56-365              (DINOHead forward / loss / loss_dn)
  models/multi/bbox_head/transformer.py:78-272            (DinoTransformer(+Decoder))
  models/multi/bbox_head/query_denoising.py:55-201        (CdnQueryGenerator; RNG draws are injected)
  models/multi/bbox_head/mmdet_detr_head/detr_head.py:333-543  (loss_single, Hungarian targets)
  models/multi/seg_head/{mask2former_head.py:111-205, pixel_decoder.py:80-171}
and of the mmdet / mmcls / mmseg pieces they call.  It
follows the reference's control flow literally (per-image scipy matching, per-layer
loops); PARITY UNPINNED (see oracle/__init__.py).
"""
import math

import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment

from . import swin as osw
from . import transformer as otr


def lin(sd, pre, x):
    return F.linear(x, sd[pre + '.weight'], sd.get(pre + '.bias'))


def ln(sd, pre, x):
    return F.layer_norm(x, (x.shape[-1],), sd[pre + '.weight'], sd[pre + '.bias'])


# ---------------------------------------------------------------- neck (a8)
def channel_mapper(sd, pre, feats, num_groups=32):
    """mmdet ChannelMapper(kernel 1, GN, no act, num_outs=4): 1x1 convs on each input,
    extra 3x3 stride-2 conv on the LAST INPUT map."""
    outs = []
    for i, x in enumerate(feats):
        y = F.conv2d(x, sd[f'{pre}convs.{i}.conv.weight'])
        outs.append(F.group_norm(y, num_groups, sd[f'{pre}convs.{i}.gn.weight'], sd[f'{pre}convs.{i}.gn.bias']))
    y = F.conv2d(feats[-1], sd[f'{pre}extra_convs.0.conv.weight'], stride=2, padding=1)
    outs.append(F.group_norm(y, num_groups, sd[f'{pre}extra_convs.0.gn.weight'], sd[f'{pre}extra_convs.0.gn.bias']))
    return outs


# ---------------------------------------------------------------- cls (a12)
def cls_forward_train(sd, backbone_feats, gt_label, label_smooth=0.1, pre='cls_head.'):
    x = backbone_feats[-1]
    tok = F.adaptive_avg_pool2d(x, (1, 1)).view(x.size(0), -1)
    score = lin(sd, pre + 'fc', tok)
    nc = score.shape[1]
    one_hot = F.one_hot(gt_label, nc).float() if gt_label.dim() == 1 else gt_label.float()
    smooth = one_hot * (1 - label_smooth) + label_smooth / nc
    loss = (-(smooth * F.log_softmax(score, -1)).sum(-1)).sum() / len(score)
    return {'loss': loss}


# ---------------------------------------------------------------- det (a13-a16)
def bbox_cxcywh_to_xyxy(b):
    cx, cy, w, h = b.split((1, 1, 1, 1), dim=-1)
    return torch.cat([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


def bbox_xyxy_to_cxcywh(b):
    x1, y1, x2, y2 = b.split((1, 1, 1, 1), dim=-1)
    return torch.cat([(x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1], dim=-1)


def bbox_overlaps_giou(b1, b2, is_aligned, eps=1e-6):
    """mmdet bbox_overlaps(mode='giou')."""
    area1 = (b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1])
    area2 = (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1])
    if is_aligned:
        lt = torch.max(b1[..., :2], b2[..., :2])
        rb = torch.min(b1[..., 2:], b2[..., 2:])
        wh = (rb - lt).clamp(min=0)
        overlap = wh[..., 0] * wh[..., 1]
        union = area1 + area2 - overlap
        elt = torch.min(b1[..., :2], b2[..., :2])
        erb = torch.max(b1[..., 2:], b2[..., 2:])
    else:
        lt = torch.max(b1[..., :, None, :2], b2[..., None, :, :2])
        rb = torch.min(b1[..., :, None, 2:], b2[..., None, :, 2:])
        wh = (rb - lt).clamp(min=0)
        overlap = wh[..., 0] * wh[..., 1]
        union = area1[..., None] + area2[..., None, :] - overlap
        elt = torch.min(b1[..., :, None, :2], b2[..., None, :, :2])
        erb = torch.max(b1[..., :, None, 2:], b2[..., None, :, 2:])
    eps = union.new_tensor([eps])
    union = torch.max(union, eps)
    ious = overlap / union
    ewh = (erb - elt).clamp(min=0)
    earea = torch.max(ewh[..., 0] * ewh[..., 1], eps)
    return ious - (earea - union) / earea


def py_sigmoid_focal_loss(pred, target, weight, gamma, alpha, avg_factor):
    """mmdet py_sigmoid_focal_loss (its CPU branch), reduction mean with avg_factor."""
    p = pred.sigmoid()
    t = F.one_hot(target, pred.shape[1] + 1)[:, :pred.shape[1]].type_as(pred)
    pt = (1 - p) * t + p * (1 - t)
    fw = (alpha * t + (1 - alpha) * (1 - t)) * pt.pow(gamma)
    loss = F.binary_cross_entropy_with_logits(pred, t, reduction='none') * fw
    loss = loss * weight.view(-1, 1)
    return loss.sum() / avg_factor


def sineembed(pos_tensor):
    """DinoTransformerDecoder.gen_sineembed_for_position (transformer.py:43-76)."""
    scale = 2 * math.pi
    dim_t = torch.arange(128, dtype=torch.float32)
    dim_t = 10000 ** (2 * (dim_t // 2) / 128)
    embs = []
    for c in range(pos_tensor.size(-1)):
        e = pos_tensor[:, :, c] * scale
        p = e[:, :, None] / dim_t
        embs.append(torch.stack((p[:, :, 0::2].sin(), p[:, :, 1::2].cos()), dim=3).flatten(2))
    order = [1, 0] + list(range(2, len(embs)))            # (y, x, w, h)
    return torch.cat([embs[i] for i in order], dim=2)


def mlp(sd, pre, x, idxs):
    for n, i in enumerate(idxs):
        x = lin(sd, f'{pre}{i}', x)
        if n < len(idxs) - 1:
            x = F.relu(x)
    return x


def cdn_queries(sd, gt_bboxes, gt_labels, img_metas, noise, num_queries=600, num_classes=20, num_dn=100,
                label_noise_scale=0.5, box_noise_scale=1.0, pre='bbox_head.'):
    """CdnQueryGenerator.__call__ with the RNG draws passed in `noise`
    (p, new_label per slot, rand_sign, rand_part)."""
    boxes_n = []
    for m, b in zip(img_metas, gt_bboxes):
        h, w, _ = m['img_shape']
        boxes_n.append(bbox_xyxy_to_cxcywh(b) / b.new_tensor([w, h, w, h]).unsqueeze(0))
    known_num = [len(l) for l in gt_labels]
    single_pad = int(max(known_num))
    num_groups = max(1, num_dn // single_pad) if single_pad > 0 else 1
    labels, boxes = torch.cat(gt_labels), torch.cat(boxes_n)
    batch_idx = torch.cat([torch.full_like(t.long(), i) for i, t in enumerate(gt_labels)])
    known_labels = labels.repeat(2 * num_groups, 1).view(-1)
    known_bid = batch_idx.repeat(2 * num_groups, 1).view(-1)
    known_bboxs = boxes.repeat(2 * num_groups, 1)
    lab = known_labels.clone()
    chosen = noise['p'] < label_noise_scale * 0.5
    lab[chosen] = noise['new_label'][chosen]
    nb = len(boxes)
    positive_idx = torch.arange(nb).unsqueeze(0).repeat(num_groups, 1) + (torch.arange(num_groups) * nb * 2).unsqueeze(1)
    negative_idx = positive_idx.flatten() + nb
    xyxy = torch.zeros_like(known_bboxs)
    xyxy[:, :2] = known_bboxs[:, :2] - known_bboxs[:, 2:] / 2
    xyxy[:, 2:] = known_bboxs[:, :2] + known_bboxs[:, 2:] / 2
    diff = torch.cat([known_bboxs[:, 2:] / 2, known_bboxs[:, 2:] / 2], 1)
    rand_sign = noise['rand_sign'] * 2.0 - 1.0
    rand_part = noise['rand_part'].clone()
    rand_part[negative_idx] += 1.0
    rand_part *= rand_sign
    xyxy = (xyxy + rand_part * diff * box_noise_scale).clamp(min=0.0, max=1.0)
    bexp = torch.cat([(xyxy[:, :2] + xyxy[:, 2:]) / 2, xyxy[:, 2:] - xyxy[:, :2]], 1)
    label_embed = sd[pre + 'label_embedding.weight'][lab]
    bbox_embed = otr.inverse_sigmoid(bexp, eps=1e-3)
    pad_size = single_pad * 2 * num_groups
    B = len(gt_bboxes)
    q_label = torch.zeros(B, pad_size, label_embed.shape[1])
    q_bbox = torch.zeros(B, pad_size, 4)
    mk = torch.cat([torch.arange(n) for n in known_num])
    mk = torch.cat([mk + single_pad * i for i in range(2 * num_groups)]).long()
    q_label = q_label.index_put((known_bid.long(), mk), label_embed)
    q_bbox = q_bbox.index_put((known_bid.long(), mk), bbox_embed)
    tgt = pad_size + num_queries
    attn_mask = torch.zeros(tgt, tgt, dtype=torch.bool)
    attn_mask[pad_size:, :pad_size] = True
    for i in range(num_groups):
        if i == 0:
            attn_mask[single_pad * 2 * i:single_pad * 2 * (i + 1), single_pad * 2 * (i + 1):pad_size] = True
        if i == num_groups - 1:
            attn_mask[single_pad * 2 * i:single_pad * 2 * (i + 1), :single_pad * i * 2] = True
        else:
            attn_mask[single_pad * 2 * i:single_pad * 2 * (i + 1), single_pad * 2 * (i + 1):pad_size] = True
            attn_mask[single_pad * 2 * i:single_pad * 2 * (i + 1), :single_pad * 2 * i] = True
    return q_label, q_bbox, attn_mask, {'pad_size': pad_size, 'num_dn_group': num_groups}


def dino_forward(sd, neck_feats, img_metas, dn_label_query, dn_bbox_query, attn_mask, *, enc_layers=6, dec_layers=6,
                 num_query=600, pre='bbox_head.', enc_pre='shared_encoder.'):
    """DINOHead.forward (dino_head.py:84-150) + DinoTransformer.forward (transformer.py:164-272)."""
    tp = pre + 'transformer.'
    B = neck_feats[0].size(0)
    ih, iw = img_metas[0]['batch_input_shape']
    img_masks = neck_feats[0].new_ones((B, ih, iw))
    for i in range(B):
        h, w, _ = img_metas[i]['img_shape']
        img_masks[i, :h, :w] = 0
    masks, poss = [], []
    for f in neck_feats:
        masks.append(F.interpolate(img_masks[None], size=f.shape[-2:]).to(torch.bool).squeeze(0))
        poss.append(otr.sine_positional_encoding(masks[-1], 128, temperature=20, normalize=True))
    feat_f, mask_f, pos_f, shapes = [], [], [], []
    for lvl, (f, m, p) in enumerate(zip(neck_feats, masks, poss)):
        shapes.append(tuple(f.shape[-2:]))
        feat_f.append(f.flatten(2).transpose(1, 2))
        mask_f.append(m.flatten(1))
        pos_f.append(p.flatten(2).transpose(1, 2) + sd[tp + 'level_embeds'][lvl].view(1, 1, -1))
    feat_f, mask_f, pos_f = torch.cat(feat_f, 1), torch.cat(mask_f, 1), torch.cat(pos_f, 1)
    starts = [0]
    for h, w in shapes[:-1]:
        starts.append(starts[-1] + h * w)
    valid_ratios = torch.stack([otr.get_valid_ratio(m) for m in masks], 1)
    ref2 = otr.get_reference_points(shapes, valid_ratios)
    memory = otr.detr_encoder(sd, enc_pre, feat_f.permute(1, 0, 2), num_layers=enc_layers,
                              query_pos=pos_f.permute(1, 0, 2), query_key_padding_mask=mask_f, spatial_shapes=shapes,
                              reference_points=ref2, level_start_index=starts)
    memory = memory.permute(1, 0, 2)
    out_mem, out_prop = otr.gen_encoder_output_proposals(sd, tp, memory, mask_f, shapes)
    enc_cls = lin(sd, f'{pre}cls_branches.{dec_layers}', out_mem)
    enc_coord = mlp(sd, f'{pre}reg_branches.{dec_layers}.', out_mem, (0, 2, 4)) + out_prop
    topk_idx = torch.topk(enc_cls.max(-1)[0], num_query, dim=1)[1]
    topk_score = torch.gather(enc_cls, 1, topk_idx.unsqueeze(-1).repeat(1, 1, enc_cls.shape[-1]))
    topk_coords_unact = torch.gather(enc_coord, 1, topk_idx.unsqueeze(-1).repeat(1, 1, 4))
    topk_anchor = topk_coords_unact.sigmoid()
    topk_coords_unact = topk_coords_unact.detach()
    query = sd[tp + 'query_embed.weight'][:, None, :].repeat(1, B, 1).transpose(0, 1)
    if dn_label_query is not None:
        query = torch.cat([dn_label_query, query], 1)
        reference_points = torch.cat([dn_bbox_query, topk_coords_unact], 1)
    else:
        reference_points = topk_coords_unact
    reference_points = reference_points.sigmoid()
    # --- DinoTransformerDecoder.forward (transformer.py:78-131)
    output = query.permute(1, 0, 2)
    mem_sf = memory.permute(1, 0, 2)
    inter, inter_ref = [], [reference_points]
    dp = tp + 'decoder.'
    for lid in range(dec_layers):
        rp_in = reference_points[:, :, None] * torch.cat([valid_ratios, valid_ratios], -1)[:, None]
        qpos = mlp(sd, dp + 'ref_point_head.', sineembed(rp_in[:, :, 0, :]), (0, 2)).permute(1, 0, 2)
        output = otr.base_transformer_layer(
            sd, f'{dp}layers.{lid}.', ('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'), ['mha', 'msda'],
            output, None, mem_sf, query_pos=qpos, attn_masks=attn_mask, key_padding_mask=mask_f,
            reference_points=rp_in, spatial_shapes=shapes, level_start_index=starts)
        out_bf = output.permute(1, 0, 2)
        tmp = mlp(sd, f'{pre}reg_branches.{lid}.', out_bf, (0, 2, 4))
        new_ref = (tmp + otr.inverse_sigmoid(reference_points, eps=1e-3)).sigmoid()
        reference_points = new_ref.detach()
        inter.append(ln(sd, dp + 'norm', output))
        inter_ref.append(new_ref)
    hs = torch.stack(inter).permute(0, 2, 1, 3)
    inter_ref = torch.stack(inter_ref)
    classes, coords = [], []
    for lvl in range(hs.shape[0]):
        reference = otr.inverse_sigmoid(inter_ref[lvl], eps=1e-3)
        classes.append(lin(sd, f'{pre}cls_branches.{lvl}', hs[lvl]))
        coords.append((mlp(sd, f'{pre}reg_branches.{lvl}.', hs[lvl], (0, 2, 4)) + reference).sigmoid())
    return torch.stack(classes), torch.stack(coords), topk_score, topk_anchor


def hungarian_assign(bbox_pred, cls_pred, gt_bboxes, gt_labels, img_shape):
    """mmdet HungarianAssigner.assign with FocalLossCost(2) + BBoxL1Cost(5,xywh) + IoUCost(giou,2)."""
    num_gts, num_bboxes = gt_bboxes.size(0), bbox_pred.size(0)
    assigned_gt_inds = bbox_pred.new_full((num_bboxes,), 0, dtype=torch.long)
    if num_gts == 0 or num_bboxes == 0:
        return assigned_gt_inds
    h, w, _ = img_shape
    factor = gt_bboxes.new_tensor([w, h, w, h]).unsqueeze(0)
    p = cls_pred.sigmoid()
    neg = -(1 - p + 1e-12).log() * (1 - 0.25) * p.pow(2)
    pos = -(p + 1e-12).log() * 0.25 * (1 - p).pow(2)
    cls_cost = (pos[:, gt_labels] - neg[:, gt_labels]) * 2.0
    reg_cost = torch.cdist(bbox_pred, bbox_xyxy_to_cxcywh(gt_bboxes / factor), p=1) * 5.0
    iou_cost = -bbox_overlaps_giou(bbox_cxcywh_to_xyxy(bbox_pred) * factor, gt_bboxes, False) * 2.0
    cost = (cls_cost + reg_cost + iou_cost).detach().cpu()
    r, c = linear_sum_assignment(cost)
    assigned_gt_inds[torch.from_numpy(r)] = torch.from_numpy(c) + 1
    return assigned_gt_inds


def det_loss_single(cls_scores, bbox_preds, gt_bboxes_list, gt_labels_list, img_metas, num_classes=20):
    """DETRHead.loss_single (detr_head.py:333-416), single process (reduce_mean = identity)."""
    B = cls_scores.size(0)
    labels_l, bt_l, bw_l, npos, nneg = [], [], [], 0, 0
    for i in range(B):
        nq = bbox_preds[i].size(0)
        assigned = hungarian_assign(bbox_preds[i], cls_scores[i], gt_bboxes_list[i], gt_labels_list[i],
                                    img_metas[i]['img_shape'])
        pos_inds = torch.nonzero(assigned > 0, as_tuple=False).squeeze(-1).unique()
        neg_inds = torch.nonzero(assigned == 0, as_tuple=False).squeeze(-1).unique()
        pos_gt = assigned[pos_inds] - 1
        labels = gt_bboxes_list[i].new_full((nq,), num_classes, dtype=torch.long)
        labels[pos_inds] = gt_labels_list[i][pos_gt]
        bt, bw = torch.zeros_like(bbox_preds[i]), torch.zeros_like(bbox_preds[i])
        bw[pos_inds] = 1.0
        h, w, _ = img_metas[i]['img_shape']
        factor = bbox_preds[i].new_tensor([w, h, w, h]).unsqueeze(0)
        bt[pos_inds] = bbox_xyxy_to_cxcywh(gt_bboxes_list[i][pos_gt] / factor)
        labels_l.append(labels), bt_l.append(bt), bw_l.append(bw)
        npos += pos_inds.numel()
        nneg += neg_inds.numel()
    return _det_losses(cls_scores, bbox_preds, torch.cat(labels_l), torch.cat(bt_l), torch.cat(bw_l), npos, nneg,
                       img_metas)


def _det_losses(cls_scores, bbox_preds, labels, bbox_targets, bbox_weights, npos, nneg, img_metas, bg_cls_weight=0):
    C = cls_scores.shape[-1]
    cls_avg_factor = max(npos * 1.0 + nneg * bg_cls_weight, 1)
    loss_cls = 1.0 * py_sigmoid_focal_loss(cls_scores.reshape(-1, C), labels, torch.ones_like(labels, dtype=torch.float32),
                                           2.0, 0.25, cls_avg_factor)
    num_total_pos = max(float(npos), 1.0)
    factors = []
    for m, bp in zip(img_metas, bbox_preds):
        h, w, _ = m['img_shape']
        factors.append(bp.new_tensor([w, h, w, h]).unsqueeze(0).repeat(bp.size(0), 1))
    factors = torch.cat(factors, 0)
    bp = bbox_preds.reshape(-1, 4)
    bboxes, bboxes_gt = bbox_cxcywh_to_xyxy(bp) * factors, bbox_cxcywh_to_xyxy(bbox_targets) * factors
    wmean = bbox_weights.mean(-1)
    if not torch.any(bbox_weights > 0):
        loss_iou = (bboxes * bbox_weights).sum()
    else:
        loss_iou = 2.0 * ((1 - bbox_overlaps_giou(bboxes, bboxes_gt, True)) * wmean).sum() / num_total_pos
    loss_bbox = 5.0 * ((bp - bbox_targets).abs() * bbox_weights).sum() / num_total_pos
    return loss_cls, loss_bbox, loss_iou


def det_loss_dn_single(dn_cls, dn_box, gt_bboxes_list, gt_labels_list, img_metas, dn_meta, num_classes=20):
    """DINOHead.loss_dn_single / _get_dn_target_single (dino_head.py:247-365)."""
    G, pad = dn_meta['num_dn_group'], dn_meta['pad_size']
    single_pad = pad // G
    labels_l, bt_l, bw_l, npos, nneg = [], [], [], 0, 0
    for i in range(dn_cls.size(0)):
        n = len(gt_labels_list[i])
        nb = dn_box[i].size(0)
        if n > 0:
            t = torch.arange(0, n).long().unsqueeze(0).repeat(G, 1)
            pos_gt = t.flatten()
            pos_inds = ((torch.arange(G) * single_pad).long().unsqueeze(1) + t).flatten()
        else:
            pos_inds = pos_gt = torch.tensor([]).long()
        neg_inds = pos_inds + single_pad // 2
        labels = gt_bboxes_list[i].new_full((nb,), num_classes, dtype=torch.long)
        labels[pos_inds] = gt_labels_list[i][pos_gt]
        bt, bw = torch.zeros_like(dn_box[i]), torch.zeros_like(dn_box[i])
        bw[pos_inds] = 1.0
        h, w, _ = img_metas[i]['img_shape']
        factor = dn_box[i].new_tensor([w, h, w, h]).unsqueeze(0)
        bt[pos_inds] = bbox_xyxy_to_cxcywh(gt_bboxes_list[i] / factor).repeat([G, 1])
        labels_l.append(labels), bt_l.append(bt), bw_l.append(bw)
        npos += pos_inds.numel()
        nneg += neg_inds.numel()
    return _det_losses(dn_cls, dn_box, torch.cat(labels_l), torch.cat(bt_l), torch.cat(bw_l), npos, nneg, img_metas)


def det_loss(all_cls, all_box, enc_cls, enc_box, gt_bboxes, gt_labels, img_metas, dn_meta):
    """DINOHead.loss (dino_head.py:152-234)."""
    ps = dn_meta['pad_size']
    dn_cls, dn_box = all_cls[:, :, :ps], all_box[:, :, :ps]
    all_cls, all_box = all_cls[:, :, ps:], all_box[:, :, ps:]
    d = {}
    d['interm_loss_cls'], d['interm_loss_bbox'], d['interm_loss_iou'] = det_loss_single(
        enc_cls, enc_box, gt_bboxes, gt_labels, img_metas)
    per = [det_loss_single(all_cls[l], all_box[l], gt_bboxes, gt_labels, img_metas) for l in range(len(all_cls))]
    d['loss_cls'], d['loss_bbox'], d['loss_iou'] = per[-1]
    for n, (a, b, c) in enumerate(per[:-1]):
        d[f'd{n}.loss_cls'], d[f'd{n}.loss_bbox'], d[f'd{n}.loss_iou'] = a, b, c
    per = [det_loss_dn_single(dn_cls[l], dn_box[l], gt_bboxes, gt_labels, img_metas, dn_meta)
           for l in range(len(dn_cls))]
    d['dn_loss_cls'], d['dn_loss_bbox'], d['dn_loss_iou'] = per[-1]
    for n, (a, b, c) in enumerate(per[:-1]):
        d[f'd{n}.dn_loss_cls'], d[f'd{n}.dn_loss_bbox'], d[f'd{n}.dn_loss_iou'] = a, b, c
    return d


# ---------------------------------------------------------------- seg (a17-a19)
def seg_pixel_decoder(sd, neck_feats, *, enc_layers=6, pre='seg_head.pixel_decoder.', enc_pre='shared_encoder.',
                      strides=(4, 8, 16, 32)):
    """MlvlSegPixelDecoder.forward with 4 input = 4 encoder levels (pixel_decoder.py:80-171)."""
    B = neck_feats[0].shape[0]
    n = len(neck_feats)
    inputs, poss, shapes, refs = [], [], [], []
    for i in range(n):
        li = n - i - 1
        f = neck_feats[li]
        h, w = f.shape[-2:]
        pe = otr.sine_positional_encoding(torch.zeros(B, h, w, dtype=torch.bool), 128, temperature=10000, normalize=True)
        lpe = sd[pre + 'level_encoding.weight'][i].view(1, -1, 1, 1) + pe
        xs = (torch.arange(0, w) + 0.5) * strides[li]
        ys = (torch.arange(0, h) + 0.5) * strides[li]
        yy, xx = torch.meshgrid(ys, xs, indexing='ij')
        rp = torch.stack([xx.reshape(-1), yy.reshape(-1)], -1) / (f.new_tensor([[w, h]]) * strides[li])
        inputs.append(f.flatten(2).permute(2, 0, 1))
        poss.append(lpe.flatten(2).permute(2, 0, 1))
        shapes.append((h, w))
        refs.append(rp)
    starts = [0]
    for h, w in shapes[:-1]:
        starts.append(starts[-1] + h * w)
    ref = torch.cat(refs, 0)[None, :, None].repeat(B, 1, n, 1)
    total = sum(h * w for h, w in shapes)
    memory = otr.detr_encoder(sd, enc_pre, torch.cat(inputs, 0), num_layers=enc_layers, query_pos=torch.cat(poss, 0),
                              query_key_padding_mask=torch.zeros(B, total, dtype=torch.bool), spatial_shapes=shapes,
                              reference_points=ref, level_start_index=starts)
    memory = memory.permute(1, 2, 0)
    outs = torch.split(memory, [h * w for h, w in shapes], dim=-1)
    outs = [x.reshape(B, -1, shapes[i][0], shapes[i][1]) for i, x in enumerate(outs)]
    mask_feature = F.conv2d(outs[-1], sd[pre + 'mask_feature.weight'], sd[pre + 'mask_feature.bias'])
    return mask_feature, outs


def seg_forward(sd, neck_feats, *, enc_layers=6, dec_layers=9, num_heads=8, pre='seg_head.'):
    """Mask2FormerHead.forward, scheme 2 (mask2former_head.py:111-199)."""
    mask_feature, mem = seg_pixel_decoder(sd, neck_feats, enc_layers=enc_layers, pre=pre + 'pixel_decoder.')
    B = mask_feature.shape[0]
    nl = len(mem)
    dec_in, dec_pe = [], []
    for i in range(nl):
        x = mem[i].flatten(2).permute(2, 0, 1) + sd[pre + 'level_embed.weight'][i].view(1, 1, -1)
        pe = otr.sine_positional_encoding(torch.zeros((B,) + mem[i].shape[-2:], dtype=torch.bool), 128,
                                          temperature=10000, normalize=True)
        dec_in.append(x)
        dec_pe.append(pe.flatten(2).permute(2, 0, 1))
    qf = sd[pre + 'query_feat.weight'].unsqueeze(1).repeat((1, B, 1))
    qe = sd[pre + 'query_embed.weight'].unsqueeze(1).repeat((1, B, 1))

    def head(dec_out, size):
        d = ln(sd, pre + 'transformer_decoder.post_norm', dec_out).transpose(0, 1)
        me = mlp(sd, pre + 'mask_embed.', d, (0, 2, 4))
        mp = torch.einsum('bqd,bdhw->bqhw', me, mask_feature)
        am = F.interpolate(mp, size, mode='bilinear', align_corners=False)
        am = am.flatten(2).unsqueeze(1).repeat((1, num_heads, 1, 1)).flatten(0, 1)
        return mp, (am.sigmoid() < 0.5).detach()

    mask_pred, attn_mask = head(qf, mem[0].shape[-2:])
    for i in range(dec_layers):
        li = i % nl
        attn_mask[torch.where(attn_mask.sum(-1) == attn_mask.shape[-1])] = False
        qf = otr.base_transformer_layer(
            sd, f'{pre}transformer_decoder.layers.{i}.', ('cross_attn', 'norm', 'self_attn', 'norm', 'ffn', 'norm'),
            ['mha', 'mha'], qf, dec_in[li], dec_in[li], query_pos=qe, key_pos=dec_pe[li], attn_masks=[attn_mask, None])
        mask_pred, attn_mask = head(qf, mem[(i + 1) % nl].shape[-2:])
    return mask_pred


def seg_losses(seg_logit, seg_label, ignore_index=255):
    """mmseg BaseDecodeHead.losses."""
    seg_logit = F.interpolate(seg_logit, size=seg_label.shape[2:], mode='bilinear', align_corners=False)
    seg_label = seg_label.squeeze(1)
    loss = F.cross_entropy(seg_logit, seg_label, reduction='none', ignore_index=ignore_index).mean()
    valid = seg_label != ignore_index
    acc = ((seg_logit.argmax(1) == seg_label) & valid).sum().float() * 100.0 / valid.sum().clamp(min=1)
    return {'loss_ce': loss, 'acc_seg': acc}


# ---------------------------------------------------------------- MTL (a21)
def mtl_losses(sd, task, batch, *, cfg=None, noise=None):
    """MTL.forward(return_loss=True) for one task; returns the raw loss dict."""
    cfg = cfg or {}
    depths = cfg.get('depths', (2, 2, 6, 2))
    feats = osw.swin_transformer(sd, batch['img'], depths=depths, num_heads=cfg.get('num_heads', (3, 6, 12, 24)),
                                 drop_path_masks=cfg.get('drop_path_masks'))
    neck = channel_mapper(sd, 'neck.', feats[-3:])
    enc_layers = cfg.get('enc_layers', 6)
    if task == 'cls':
        return cls_forward_train(sd, feats, batch['gt_label'])
    if task == 'det':
        metas = batch['img_metas']
        for m in metas:
            m['batch_input_shape'] = tuple(batch['img'].shape[-2:])
        nq, nd = cfg.get('num_query', 600), cfg.get('num_dn', 100)
        ql, qb, am, dn_meta = cdn_queries(sd, batch['gt_bboxes'], batch['gt_labels'], metas, noise, num_queries=nq,
                                          num_dn=nd)
        outs = dino_forward(sd, neck, metas, ql, qb, am, enc_layers=enc_layers, dec_layers=cfg.get('det_dec_layers', 6),
                            num_query=nq)
        return det_loss(*outs, batch['gt_bboxes'], batch['gt_labels'], metas, dn_meta)
    if task == 'seg':
        logits = seg_forward(sd, neck, enc_layers=enc_layers, dec_layers=cfg.get('seg_dec_layers', 9))
        return {'seg.' + k: v for k, v in seg_losses(logits, batch['gt_semantic_seg']).items()}
    raise AssertionError(task)


def parse_losses(losses, task_weight=1.0):
    """MTL._parse_losses + train_step weighting (multitask_learner.py:229-306), single process."""
    log_vars = {}
    for k, v in losses.items():
        if isinstance(v, torch.Tensor):
            log_vars[k] = v.mean()
        elif isinstance(v, list):                     # multitask_learner.py:279-281
            log_vars[k] = sum(x.mean() for x in v)
        else:
            raise TypeError('%s is not a tensor or list of tensors' % k)
    loss = sum(v for k, v in log_vars.items() if 'loss' in k)
    log_vars['loss'] = loss
    return loss * task_weight, {k: float(v.detach()) * task_weight for k, v in log_vars.items()}


def cdn_noise(gt_labels, num_dn=100, num_classes=20, generator=None):
    """Draw the CDN generator's random numbers once so product and oracle share them."""
    n = sum(len(l) for l in gt_labels)
    single_pad = max(len(l) for l in gt_labels)
    G = max(1, num_dn // single_pad) if single_pad else 1
    m = n * 2 * G
    return dict(p=torch.rand(m, generator=generator), new_label=torch.randint(0, num_classes, (m,), generator=generator),
                rand_sign=torch.randint(0, 2, (m, 4), generator=generator).float(),
                rand_part=torch.rand(m, 4, generator=generator))
