"""Oracle: mmcv 1.6.1 transformer bricks + mmdet 2.25.1 DETR helpers (TEST INFRA).

Reference call sites: the shared encoder is built at
``models/multi/multitask_learner.py:51`` and invoked from
``models/multi/bbox_head/transformer.py:211-221``,
``models/multi/seg_head/pixel_decoder.py:134-146`` and
``models/multi/cls_head/pixel_decoder.py:95-107``.  The arithmetic is
mmcv ``cnn/bricks/transformer.py`` / ``ops/multi_scale_deform_attn.py`` and
mmdet ``models/utils/{transformer,positional_encoding}.py`` (not vendored;
restated per SURVEY.md Appendix D.2-D.4).
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# ms_deform_attn core == mmcv multi_scale_deformable_attn_pytorch (CPU branch
# of the op the reference calls)
# ----------------------------------------------------------------------------
def ms_deform_attn_core(value, spatial_shapes, sampling_locations, attention_weights):
    """value (B,Nv,H,D); spatial_shapes [(h,w)..]; loc (B,Nq,H,L,P,2) in [0,1]
    (x,y); w (B,Nq,H,L,P)  ->  (B,Nq,H*D)."""
    B, _, H, D = value.shape
    _, Nq, _, L, P, _ = sampling_locations.shape
    shapes = [(int(h), int(w)) for h, w in spatial_shapes]
    value_list = value.split([h * w for h, w in shapes], dim=1)
    grids = 2 * sampling_locations - 1
    sampled = []
    for lvl, (h, w) in enumerate(shapes):
        v = value_list[lvl].flatten(2).transpose(1, 2).reshape(B * H, D, h, w)
        g = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)
        sampled.append(F.grid_sample(v, g, mode='bilinear', padding_mode='zeros',
                                     align_corners=False))
    aw = attention_weights.transpose(1, 2).reshape(B * H, 1, Nq, L * P)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * aw).sum(-1).view(B, H * D, Nq)
    return out.transpose(1, 2).contiguous()


def ms_deform_attn_loops(value, spatial_shapes, level_start_index, loc, w):
    """Scalar-loop restatement of the CUDA kernel's per-sample arithmetic
    (ms_deformable_im2col_gpu_kernel: x_im = loc_x*W - 0.5, valid iff
    -1 < x_im < W, 4-corner bilinear with per-corner bounds).  Pure python:
    small cases only; used to pin ms_deform_attn_core's conventions."""
    B, Nv, H, D = value.shape
    _, Nq, _, L, P, _ = loc.shape
    out = torch.zeros(B, Nq, H, D, dtype=value.dtype)
    for b in range(B):
        for q in range(Nq):
            for hh in range(H):
                for l in range(L):
                    Hl, Wl = int(spatial_shapes[l][0]), int(spatial_shapes[l][1])
                    start = int(level_start_index[l])
                    for p in range(P):
                        x = float(loc[b, q, hh, l, p, 0]) * Wl - 0.5
                        y = float(loc[b, q, hh, l, p, 1]) * Hl - 0.5
                        if not (y > -1 and x > -1 and y < Hl and x < Wl):
                            continue
                        y0, x0 = math.floor(y), math.floor(x)
                        ly, lx = y - y0, x - x0
                        acc = torch.zeros(D, dtype=value.dtype)
                        for (yy, xx, ww) in ((y0, x0, (1 - ly) * (1 - lx)), (y0, x0 + 1, (1 - ly) * lx),
                                             (y0 + 1, x0, ly * (1 - lx)), (y0 + 1, x0 + 1, ly * lx)):
                            if 0 <= yy < Hl and 0 <= xx < Wl:
                                acc += ww * value[b, start + yy * Wl + xx, hh]
                        out[b, q, hh] += float(w[b, q, hh, l, p]) * acc
    return out.view(B, Nq, H * D)


def msda_init_state(pre, embed_dims=256, num_heads=8, num_levels=4, num_points=4, generator=None):
    """MultiScaleDeformableAttention.init_weights."""
    sd = {}
    sd[pre + 'sampling_offsets.weight'] = torch.zeros(num_heads * num_levels * num_points * 2, embed_dims)
    thetas = torch.arange(num_heads, dtype=torch.float32) * (2.0 * math.pi / num_heads)
    g = torch.stack([thetas.cos(), thetas.sin()], -1)
    g = (g / g.abs().max(-1, keepdim=True)[0]).view(num_heads, 1, 1, 2).repeat(1, num_levels, num_points, 1)
    for i in range(num_points):
        g[:, :, i, :] *= i + 1
    sd[pre + 'sampling_offsets.bias'] = g.view(-1)
    sd[pre + 'attention_weights.weight'] = torch.zeros(num_heads * num_levels * num_points, embed_dims)
    sd[pre + 'attention_weights.bias'] = torch.zeros(num_heads * num_levels * num_points)
    for n in ('value_proj', 'output_proj'):
        wt = torch.empty(embed_dims, embed_dims)
        torch.nn.init.xavier_uniform_(wt, generator=generator)
        sd[pre + n + '.weight'] = wt
        sd[pre + n + '.bias'] = torch.zeros(embed_dims)
    return sd


def msda(sd, pre, query, key=None, value=None, identity=None, query_pos=None, key_padding_mask=None,
         reference_points=None, spatial_shapes=None, level_start_index=None,
         num_heads=8, num_levels=4, num_points=4, **_ignored):
    """mmcv MultiScaleDeformableAttention.forward, batch_first=False.
    query (Nq,B,E) seq-first; returns (Nq,B,E)."""
    if value is None:
        value = query
    if identity is None:
        identity = query
    if query_pos is not None:
        query = query + query_pos
    query = query.permute(1, 0, 2)
    value = value.permute(1, 0, 2)
    B, Nq, E = query.shape
    Nv = value.shape[1]
    shapes = [(int(h), int(w)) for h, w in spatial_shapes]
    assert sum(h * w for h, w in shapes) == Nv
    value = F.linear(value, sd[pre + 'value_proj.weight'], sd[pre + 'value_proj.bias'])
    if key_padding_mask is not None:
        value = value.masked_fill(key_padding_mask[..., None], 0.0)
    value = value.view(B, Nv, num_heads, -1)
    off = F.linear(query, sd[pre + 'sampling_offsets.weight'], sd[pre + 'sampling_offsets.bias'])
    off = off.view(B, Nq, num_heads, num_levels, num_points, 2)
    aw = F.linear(query, sd[pre + 'attention_weights.weight'], sd[pre + 'attention_weights.bias'])
    aw = aw.view(B, Nq, num_heads, num_levels * num_points).softmax(-1)
    aw = aw.view(B, Nq, num_heads, num_levels, num_points)
    ss = torch.as_tensor(shapes, dtype=query.dtype)
    if reference_points.shape[-1] == 2:
        norm = torch.stack([ss[..., 1], ss[..., 0]], -1)
        loc = reference_points[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    elif reference_points.shape[-1] == 4:
        loc = reference_points[:, :, None, :, None, :2] + \
            off / num_points * reference_points[:, :, None, :, None, 2:] * 0.5
    else:
        raise ValueError('last dim of reference_points must be 2 or 4')
    out = ms_deform_attn_core(value, shapes, loc, aw)
    out = F.linear(out, sd[pre + 'output_proj.weight'], sd[pre + 'output_proj.bias'])
    return out.permute(1, 0, 2) + identity        # dropout p=0 in every config


# ----------------------------------------------------------------------------
# FFN / MultiheadAttention / BaseTransformerLayer           SURVEY D.2
# ----------------------------------------------------------------------------
def ffn(sd, pre, x, identity=None, act='relu'):
    y = F.linear(x, sd[pre + 'layers.0.0.weight'], sd[pre + 'layers.0.0.bias'])
    y = F.relu(y) if act == 'relu' else F.gelu(y)
    y = F.linear(y, sd[pre + 'layers.1.weight'], sd[pre + 'layers.1.bias'])
    return (x if identity is None else identity) + y


def mha(sd, pre, query, key=None, value=None, identity=None, query_pos=None, key_pos=None,
        attn_mask=None, key_padding_mask=None, num_heads=8, **_ignored):
    """mmcv MultiheadAttention (wraps nn.MultiheadAttention, seq-first)."""
    if key is None:
        key = query
    if value is None:
        value = key
    if identity is None:
        identity = query
    if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
        key_pos = query_pos
    if query_pos is not None:
        query = query + query_pos
    if key_pos is not None:
        key = key + key_pos
    E = query.shape[-1]
    out = F.multi_head_attention_forward(
        query, key, value, E, num_heads,
        sd[pre + 'attn.in_proj_weight'], sd[pre + 'attn.in_proj_bias'],
        None, None, False, 0.0,
        sd[pre + 'attn.out_proj.weight'], sd[pre + 'attn.out_proj.bias'],
        training=False, key_padding_mask=key_padding_mask, need_weights=False,
        attn_mask=attn_mask)[0]
    return identity + out


def base_transformer_layer(sd, pre, operation_order, attn_types, query, key=None, value=None,
                           query_pos=None, key_pos=None, attn_masks=None,
                           query_key_padding_mask=None, key_padding_mask=None, **kw):
    """mmcv BaseTransformerLayer.forward, post-norm (pre_norm False in all cfgs).
    attn_types: list of 'msda' | 'mha' per attention in the layer."""
    num_attn = len(attn_types)
    if attn_masks is None:
        attn_masks = [None] * num_attn
    elif isinstance(attn_masks, torch.Tensor):
        attn_masks = [attn_masks] * num_attn
    ai = ni = fi = 0
    fn = {'msda': msda, 'mha': mha}
    for op in operation_order:
        if op == 'self_attn':
            query = fn[attn_types[ai]](
                sd, f'{pre}attentions.{ai}.', query, query, query, None,
                query_pos=query_pos, key_pos=query_pos, attn_mask=attn_masks[ai],
                key_padding_mask=query_key_padding_mask, **kw)
            ai += 1
        elif op == 'norm':
            E = query.shape[-1]
            query = F.layer_norm(query, (E,), sd[f'{pre}norms.{ni}.weight'], sd[f'{pre}norms.{ni}.bias'])
            ni += 1
        elif op == 'cross_attn':
            query = fn[attn_types[ai]](
                sd, f'{pre}attentions.{ai}.', query, key, value, None,
                query_pos=query_pos, key_pos=key_pos, attn_mask=attn_masks[ai],
                key_padding_mask=key_padding_mask, **kw)
            ai += 1
        elif op == 'ffn':
            query = ffn(sd, f'{pre}ffns.{fi}.', query, None)
            fi += 1
    return query


def detr_encoder(sd, pre, query, num_layers=6, **kw):
    """mmdet DetrTransformerEncoder of BaseTransformerLayer(self_attn=MSDA)
    (cfg main:34-50); post_norm is None because the layers are post-norm."""
    for l in range(num_layers):
        query = base_transformer_layer(
            sd, f'{pre}layers.{l}.', ('self_attn', 'norm', 'ffn', 'norm'), ['msda'],
            query, None, None, **kw)
    return query


def transformer_layer_init_state(pre, attn_types, embed_dims=256, ffn_ch=2048, n_norm=2, generator=None):
    """PyTorch-default init for Linear/LN + MSDA.init_weights (multitask_learner.py:73-79)."""
    sd = {}

    def lin(name, out_f, in_f):
        l = torch.nn.Linear(in_f, out_f)
        if generator is not None:
            bound = 1 / math.sqrt(in_f)
            l.weight.data.uniform_(-bound, bound, generator=generator)
            l.bias.data.uniform_(-bound, bound, generator=generator)
        sd[name + '.weight'] = l.weight.data.clone()
        sd[name + '.bias'] = l.bias.data.clone()

    for i, t in enumerate(attn_types):
        a = f'{pre}attentions.{i}.'
        if t == 'msda':
            sd.update(msda_init_state(a, embed_dims, generator=generator))
        else:
            w = torch.empty(3 * embed_dims, embed_dims)
            torch.nn.init.xavier_uniform_(w, generator=generator)
            sd[a + 'attn.in_proj_weight'] = w
            sd[a + 'attn.in_proj_bias'] = torch.zeros(3 * embed_dims)
            lin(a + 'attn.out_proj', embed_dims, embed_dims)
            sd[a + 'attn.out_proj.bias'].zero_()
    lin(f'{pre}ffns.0.layers.0.0', ffn_ch, embed_dims)
    lin(f'{pre}ffns.0.layers.1', embed_dims, ffn_ch)
    for i in range(n_norm):
        sd[f'{pre}norms.{i}.weight'] = torch.ones(embed_dims)
        sd[f'{pre}norms.{i}.bias'] = torch.zeros(embed_dims)
    return sd


# ----------------------------------------------------------------------------
# mmdet helpers                                             SURVEY D.4 / App. A
# ----------------------------------------------------------------------------
def sine_positional_encoding(mask, num_feats=128, temperature=10000, normalize=True,
                             scale=2 * math.pi, eps=1e-6, offset=0.0):
    """mask (B,H,W) bool/int, True = padded.  -> (B, 2*num_feats, H, W)"""
    mask = mask.to(torch.int)
    not_mask = 1 - mask
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    x_embed = not_mask.cumsum(2, dtype=torch.float32)
    if normalize:
        y_embed = (y_embed + offset) / (y_embed[:, -1:, :] + eps) * scale
        x_embed = (x_embed + offset) / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    B, H, W = mask.shape
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).view(B, H, W, -1)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).view(B, H, W, -1)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


def inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(min=0, max=1)
    x1 = x.clamp(min=eps)
    x2 = (1 - x).clamp(min=eps)
    return torch.log(x1 / x2)


def get_valid_ratio(mask):
    _, H, W = mask.shape
    valid_H = torch.sum(~mask[:, :, 0], 1)
    valid_W = torch.sum(~mask[:, 0, :], 1)
    return torch.stack([valid_W.float() / W, valid_H.float() / H], -1)


def get_reference_points(spatial_shapes, valid_ratios):
    """DeformableDetrTransformer.get_reference_points -> (B, N, L, 2)"""
    ref_list = []
    for lvl, (H, W) in enumerate(spatial_shapes):
        ref_y, ref_x = torch.meshgrid(torch.linspace(0.5, H - 0.5, H), torch.linspace(0.5, W - 0.5, W),
                                      indexing='ij')
        ref_y = ref_y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H)
        ref_x = ref_x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W)
        ref_list.append(torch.stack((ref_x, ref_y), -1))
    reference_points = torch.cat(ref_list, 1)
    return reference_points[:, :, None] * valid_ratios[:, None]


def gen_encoder_output_proposals(sd, pre, memory, memory_padding_mask, spatial_shapes):
    """DeformableDetrTransformer.gen_encoder_output_proposals (pre = '...transformer.')."""
    N, S, C = memory.shape
    proposals = []
    _cur = 0
    for lvl, (H, W) in enumerate(spatial_shapes):
        mask_flatten_ = memory_padding_mask[:, _cur:(_cur + H * W)].view(N, H, W, 1)
        valid_H = torch.sum(~mask_flatten_[:, :, 0, 0], 1)
        valid_W = torch.sum(~mask_flatten_[:, 0, :, 0], 1)
        grid_y, grid_x = torch.meshgrid(torch.linspace(0, H - 1, H), torch.linspace(0, W - 1, W), indexing='ij')
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
    output_memory = memory.masked_fill(memory_padding_mask.unsqueeze(-1), float(0))
    output_memory = output_memory.masked_fill(~output_proposals_valid, float(0))
    output_memory = F.linear(output_memory, sd[pre + 'enc_output.weight'], sd[pre + 'enc_output.bias'])
    output_memory = F.layer_norm(output_memory, (C,), sd[pre + 'enc_output_norm.weight'],
                                 sd[pre + 'enc_output_norm.bias'])
    return output_memory, output_proposals
