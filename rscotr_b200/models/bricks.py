"""Transformer bricks with the registry names / state-dict layout the reference
configs expect (mmcv 1.6.1 `cnn/bricks/transformer.py`, `ops/multi_scale_deform_attn.py`,
mmdet 2.25.1 `models/utils/{transformer,positional_encoding}.py`,
`models/necks/channel_mapper.py`).

The ms_deform_attn core runs through the C-ABI kernel (ops.ms_deform_attn);
Linear / LayerNorm / Conv are library GEMMs.
"""
import copy
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..config import MODELS, build_from_cfg


_CONST_CACHE = {}
_DEFER = __import__('os').environ.get('RSC_NO_DEFER') is None     # A/B switch for the fused post-norm pairs
_OWN_ATTN = __import__('os').environ.get('RSC_OWN_ATTN', '1') != '0'       # rsc_attn_* core (0: library SDPA)
_CONST_ATTN_BIAS = __import__('os').environ.get('RSC_CONST_ATTN_BIAS', '1') != '0'    # A/B'd on B200 in round 2: promoted


def const_tensor(values, dtype, device):
    """Small host-defined constant (image factors, level shapes, ...) as a cached device tensor:
    created once (one pageable H2D copy), reused afterwards -- no per-step H2D copies and
    nothing illegal under CUDA-graph capture."""
    def freeze(v):
        return tuple(freeze(x) for x in v) if isinstance(v, (list, tuple)) else v
    key = (freeze(values), dtype, str(device))
    t = _CONST_CACHE.get(key)
    if t is None:
        t = torch.tensor(values, dtype=dtype, device=device)
        _CONST_CACHE[key] = t
    return t


def apply_init_cfg(module, init_cfg):
    """mmcv BaseModule.init_weights for the layer-wise entries the reference configs use (`TruncNormal` / `Normal` /
    `Constant` with a `layer` filter): every sub-module whose class name is listed is re-initialised."""
    for c in ([init_cfg] if isinstance(init_cfg, dict) else (init_cfg or [])):
        layers = c.get('layer')
        layers = [layers] if isinstance(layers, str) else list(layers or [])
        for m in module.modules():
            if type(m).__name__ not in layers or getattr(m, 'weight', None) is None:
                continue
            if c['type'] == 'TruncNormal':
                nn.init.trunc_normal_(m.weight, mean=c.get('mean', 0.), std=c.get('std', 1.), a=c.get('a', -2.), b=c.get('b', 2.))
            elif c['type'] == 'Normal':
                nn.init.normal_(m.weight, c.get('mean', 0.), c.get('std', 1.))
            elif c['type'] == 'Constant':
                nn.init.constant_(m.weight, c['val'])
            else:
                raise KeyError('init_cfg type %s is not supported' % c['type'])
            if getattr(m, 'bias', None) is not None:
                nn.init.constant_(m.bias, c.get('bias', 0.))


class GeomCache:
    """Small shape-keyed cache of no-grad geometry constants of a head (padding masks, sine positional
    encodings, reference points, proposal grids ...).  The reference recomputes them with ~25-80 tiny
    kernels each on every iteration although they only depend on the image / feature-map shapes."""

    def __init__(self, capacity=8):
        self.capacity, self.entries = capacity, {}

    def get(self, key, make):
        v = self.entries.get(key)
        if v is None:
            with torch.no_grad():
                v = make()
            if len(self.entries) >= self.capacity:
                self.entries.pop(next(iter(self.entries)))
            self.entries[key] = v
        return v


class PackedLosses(dict):
    """A loss dict (name -> scalar tensor, the reference's contract) whose values are the elements of ONE
    packed 1-D tensor: a fused loss kernel returns all its terms at once, MTL._parse_losses sums `packed`
    directly (2 kernels instead of ~3 per term), and per-key access still works for callers that want it."""

    def __init__(self, keys, packed):
        super().__init__()
        assert packed.dim() == 1 and packed.numel() == len(keys)
        self.key_list, self.packed = list(keys), packed
        for k in self.key_list:
            dict.__setitem__(self, k, None)

    def _fill(self):
        if self.key_list and dict.__getitem__(self, self.key_list[0]) is None:
            for k, v in zip(self.key_list, self.packed.unbind(0)):
                dict.__setitem__(self, k, v)

    def __getitem__(self, k):
        self._fill()
        return dict.__getitem__(self, k)

    def get(self, k, default=None):
        self._fill()
        return dict.get(self, k, default)

    def items(self):
        self._fill()
        return dict.items(self)

    def values(self):
        self._fill()
        return dict.values(self)


def build_activation(cfg):
    cfg = dict(cfg or dict(type='ReLU'))
    t = cfg.pop('type')
    if t == 'ReLU':
        return nn.ReLU(inplace=False)    # (outputs of the custom Linear function must not be modified in place)
    if t == 'GELU':
        return nn.GELU()
    raise KeyError('unsupported activation %s' % t)


class Linear(nn.Linear):
    """nn.Linear parameters / state-dict keys; forward through ops.linear (library GEMMs, bias
    gradient on rsc_colsum)."""

    def forward(self, x):
        return ops.linear(x, self.weight, self.bias)


class LayerNorm(nn.LayerNorm):
    """nn.LayerNorm parameters / state-dict keys, forward on rsc_layernorm_{fwd,bwd}: one pass,
    fp32 statistics, output already in the compute dtype of the GEMM that follows."""

    def forward(self, x):
        return ops.layer_norm(x, self.weight, self.bias, self.eps)


class GroupNorm(nn.GroupNorm):
    """nn.GroupNorm parameters / state-dict keys, forward on rsc_groupnorm_{fwd,bwd} (channels-last, optional fused ReLU)"""

    def forward(self, x, relu=False):
        if self.affine and ops.norm_supported(x, self.num_channels):
            return ops.group_norm(x, self.num_groups, self.weight, self.bias, self.eps, relu)
        y = super().forward(x)
        return F.relu(y) if relu else y


class BatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d (per rank, as the reference's single-GPU configs); the training-mode forward runs on
    rsc_groupnorm_{fwd,bwd} (one row, groups of one channel), evaluation uses the running statistics through torch"""

    def forward(self, x, relu=False):
        if (self.training and self.affine and self.track_running_stats and self.momentum is not None and
                ops.norm_supported(x, self.num_features)):
            if self.num_batches_tracked is not None:
                self.num_batches_tracked.add_(1)
            return ops.batch_norm_train(x, self.weight, self.bias, self.running_mean, self.running_var, self.momentum,
                                        self.eps, relu)
        y = super().forward(x)
        return F.relu(y) if relu else y


def build_norm(cfg, num_features):
    cfg = dict(cfg)
    t = cfg.pop('type')
    if t == 'LN':
        return LayerNorm(num_features, **cfg)
    if t == 'GN':
        return GroupNorm(cfg.pop('num_groups'), num_features, **cfg)
    if t in ('BN', 'SyncBN'):      # per-rank BN (no forward-time collectives)
        return BatchNorm2d(num_features, **{k: v for k, v in cfg.items() if k != 'requires_grad'})
    raise KeyError('unsupported norm %s' % t)


def norm_abbr(cfg):
    return {'LN': 'ln', 'GN': 'gn', 'BN': 'bn', 'SyncBN': 'bn'}[cfg['type']]


class DropPath(nn.Module):
    """Stochastic depth per sample (mmcv DropPath: x / keep * floor(keep + U[0,1))).  `forced_mask`
    lets parity tests inject the (already 1/keep-scaled) keep mask.  `add_to` fuses the residual add:
    identity + drop_path(x) as ONE addcmul pass."""

    def __init__(self, drop_prob=0.):
        super().__init__()
        self.drop_prob = drop_prob
        self.forced_mask = None
        self.drawn = None          # set by draw_drop_paths: this call's factor, drawn in one batched launch

    def _scale(self, x):
        """per-sample factor (B,1,...,1) or None for identity."""
        shape = (x.shape[0],) + (1,) * (x.dim() - 1)
        if self.forced_mask is not None:
            return self.forced_mask.to(x.dtype).view(shape)
        if self.drawn is not None:
            s, self.drawn = self.drawn, None
            return s.to(x.dtype).view(shape)
        if self.drop_prob == 0. or not self.training:
            return None
        keep = 1 - self.drop_prob
        mask = keep + torch.rand(shape, dtype=torch.float32, device=x.device)
        return (mask.floor() / keep).to(x.dtype)

    def scale_vec(self, x):
        """this call's per-sample factor as an fp32 (B,) tensor, or None for identity (fused add_ln path)."""
        if self.forced_mask is not None:
            return self.forced_mask.reshape(-1).float()
        if self.drawn is not None:
            s, self.drawn = self.drawn, None
            return s.reshape(-1).float()
        if self.drop_prob == 0. or not self.training:
            return None
        keep = 1 - self.drop_prob
        return (keep + torch.rand(x.shape[0], dtype=torch.float32, device=x.device)).floor() / keep

    def forward(self, x):
        s = self._scale(x)
        return x if s is None else x * s

    def add_to(self, identity, x):
        s = self._scale(x)
        if s is None:
            return identity + x
        return torch.addcmul(identity, x, s)


def draw_drop_paths(drop_paths, batch, device, dtype):
    """Draw the stochastic-depth factors of a whole list of DropPath modules with ONE rand launch
    (instead of rand / add / floor / div / cast per module): row k = floor(keep_k + U[0,1)) / keep_k."""
    active = [m for m in drop_paths if m.training and m.drop_prob > 0. and m.forced_mask is None]
    if not active:
        return
    keep = const_tensor([[1.0 - m.drop_prob] for m in active], torch.float32, device)
    s = ((keep + torch.rand(len(active), batch, dtype=torch.float32, device=device)).floor() / keep).to(dtype)
    for m, row in zip(active, s.unbind(0)):
        m.drawn = row


def build_dropout(cfg):
    if cfg is None:
        return nn.Identity()
    cfg = dict(cfg)
    t = cfg.pop('type')
    if t == 'DropPath':
        return DropPath(cfg.get('drop_prob', 0.))
    if t == 'Dropout':
        return nn.Dropout(cfg.get('drop_prob', cfg.get('p', 0.5)))
    raise KeyError(t)


@MODELS.register_module()
class FFN(nn.Module):
    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2, act_cfg=dict(type='ReLU', inplace=True),
                 ffn_drop=0., dropout_layer=None, add_identity=True, init_cfg=None, **kwargs):
        super().__init__()
        assert num_fcs >= 2
        self.embed_dims = embed_dims
        layers = []
        in_ch = embed_dims
        for _ in range(num_fcs - 1):
            layers.append(nn.Sequential(Linear(in_ch, feedforward_channels), build_activation(act_cfg),
                                        nn.Dropout(ffn_drop)))
            in_ch = feedforward_channels
        layers.append(Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = nn.Sequential(*layers)
        self.dropout_layer = build_dropout(dropout_layer)
        self.add_identity = add_identity

    def _fused(self, x, with_last_bias):
        """the two-Linear FFN as ops.mlp (activation and its gradient in the GEMM epilogues) when it applies, else None"""
        if len(self.layers) != 3:
            return None
        (lin, act, drop), last = self.layers[0], self.layers[1]
        code = ops.ACT_RELU if isinstance(act, nn.ReLU) else (
            ops.ACT_GELU if isinstance(act, nn.GELU) and act.approximate == 'none' else None)
        if code is None or (self.training and drop.p > 0.) or not ops.mlp_supported(x, lin.weight, lin.bias, last.weight):
            return None
        return ops.mlp(x, lin.weight, lin.bias, last.weight, last.bias if with_last_bias else None, code)

    def fused_add_ln(self, x, identity, norm):
        """(r, n) = (identity + FFN(x), norm(r)) through ops.mlp_add_ln (the residual add and the LayerNorm in the second
        GEMM's epilogue), or None when that path does not apply (dropout active, widths, dtype, short sequences)."""
        if len(self.layers) != 3 or not self.add_identity:
            return None
        (lin, act, drop), last = self.layers[0], self.layers[1]
        code = ops.ACT_RELU if isinstance(act, nn.ReLU) else (
            ops.ACT_GELU if isinstance(act, nn.GELU) and act.approximate == 'none' else None)
        active = self.training and (drop.p > 0. or self.layers[2].p > 0. or (
            isinstance(self.dropout_layer, nn.Dropout) and self.dropout_layer.p > 0.) or (
            isinstance(self.dropout_layer, DropPath) and self.dropout_layer.drop_prob > 0.))
        if code is None or active or last.bias is None or not ops.mlp_supported(x, lin.weight, lin.bias, last.weight) or \
                not ops.linear_add_ln_supported(x, last.weight, identity):
            return None
        return ops.mlp_add_ln(x, lin.weight, lin.bias, last.weight, last.bias, identity, None, norm.weight, norm.bias, code,
                              norm.eps)

    def _mlp(self, x):
        """self.layers(x) with Linear + bias + activation of the hidden layers as GEMM + ONE fused pass
        (ops.bias_relu / ops.bias_gelu: the bias gradient comes out of the activation's backward pass)."""
        z = self._fused(x, True)
        if z is not None:
            return self.layers[-1](z)
        for layer in list(self.layers)[:-2]:
            lin, act = layer[0], layer[1]
            code = ops.ACT_RELU if isinstance(act, nn.ReLU) else (
                ops.ACT_GELU if isinstance(act, nn.GELU) and act.approximate == 'none' else None)
            if code is not None and lin.bias is not None and ops.bias_act_supported(x, lin.out_features):
                x = layer[2](ops.bias_act(ops.linear(x, lin.weight, None), lin.bias, code))
            else:
                x = layer(x)
        return self.layers[-1](self.layers[-2](x))

    def _mlp_deferred(self, x):
        """(last Linear's output WITHOUT bias, that bias) when the trailing dropouts are inactive, else None."""
        last, drop = self.layers[-2], self.layers[-1]
        active = self.training and (drop.p > 0. or (isinstance(self.dropout_layer, nn.Dropout) and self.dropout_layer.p > 0.)
                                    or (isinstance(self.dropout_layer, DropPath) and self.dropout_layer.drop_prob > 0.))
        if active or not self.add_identity or last.bias is None:
            return None
        z = self._fused(x, False)
        if z is not None:
            return z, last.bias
        for layer in list(self.layers)[:-2]:
            lin, act = layer[0], layer[1]
            code = ops.ACT_RELU if isinstance(act, nn.ReLU) else (
                ops.ACT_GELU if isinstance(act, nn.GELU) and act.approximate == 'none' else None)
            if code is not None and lin.bias is not None and ops.bias_act_supported(x, lin.out_features):
                x = layer[2](ops.bias_act(ops.linear(x, lin.weight, None), lin.bias, code))
            else:
                x = layer(x)
        return ops.linear(x, last.weight, None), last.bias

    def forward(self, x, identity=None, _defer=False):
        if _defer:
            d = self._mlp_deferred(x)
            if d is not None:
                return d[0], d[1], (x if identity is None else identity)
        out = self._mlp(x)
        if not self.add_identity:
            return self.dropout_layer(out)
        if identity is None:
            identity = x
        if isinstance(self.dropout_layer, DropPath):
            return self.dropout_layer.add_to(identity, out)
        return identity + self.dropout_layer(out)


@MODELS.register_module()
class MultiheadAttention(nn.Module):
    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0., dropout_layer=dict(type='Dropout', drop_prob=0.),
                 init_cfg=None, batch_first=False, **kwargs):
        super().__init__()
        if 'dropout' in kwargs:
            attn_drop = kwargs['dropout']
            dropout_layer = dict(dropout_layer or dict(type='Dropout'))
            dropout_layer['drop_prob'] = kwargs.pop('dropout')
        self.embed_dims = embed_dims
        self.num_heads = num_heads
        self.batch_first = batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop, **kwargs)
        self.proj_drop = nn.Dropout(proj_drop)
        self.dropout_layer = build_dropout(dropout_layer)

    def _attend(self, query, key, value, attn_mask, key_padding_mask, same_qk=False, defer=False):
        """nn.MultiheadAttention.forward(need_weights=False) on seq-first (L,B,E) tensors with the packed
        in_proj applied through ops.linear (bf16 shadow weights, gradients accumulated into the flat
        buffer; q and k share one GEMM when they read the same tensor) and the rsc_attn_{fwd,bwd} core (bf16, head
        dim 32; anything else -- the fp32 parity path, attention dropout, key padding masks -- takes the library SDPA)."""
        m = self.attn
        E, H = self.embed_dims, self.num_heads
        L, B, _ = query.shape
        S = key.shape[0]
        W, bias = m.in_proj_weight, m.in_proj_bias
        qk = None
        if same_qk or key is query:
            qk = ops.linear(query, W, bias, rows=(0, 2 * E))
            q, k = qk[..., :E], qk[..., E:]
        else:
            q, k = ops.linear(query, W, bias, rows=(0, E)), ops.linear(key, W, bias, rows=(E, 2 * E))
        v = ops.linear(value, W, bias, rows=(2 * E, 3 * E))
        own = (_OWN_ATTN and ops.attention_supported(q, E, H) and key_padding_mask is None and
               not (self.training and m.dropout > 0.) and
               (attn_mask is None or isinstance(attn_mask, ops.MaskBits) or
                (torch.is_tensor(attn_mask) and attn_mask.dtype == torch.bool and
                 (attn_mask.dim() == 2 or (attn_mask.dim() == 4 and attn_mask.shape[1] == 1)))))
        if own:
            # rsc_attn_{fwd,bwd}: masks as bit words (one bit per (image | 1, query, key), shared by the heads)
            bits = attn_mask
            if torch.is_tensor(attn_mask):
                cache = getattr(attn_mask, '_rsc_add', None)       # shape-only constant (DINO denoising mask): pack once
                bits = cache.get('bits') if cache is not None else None
                if bits is None:
                    bits = ops.pack_mask_bits(attn_mask if attn_mask.dim() == 2 else attn_mask[:, 0])
                    if cache is not None:
                        cache['bits'] = bits
            out = ops.attention(q, k, v.to(q.dtype), H, bits, packed_qk=qk if qk is not None and qk.dtype == v.dtype else None)
            if defer:
                return ops.linear(out, m.out_proj.weight, None)
            return ops.linear(out, m.out_proj.weight, m.out_proj.bias)
        if isinstance(attn_mask, ops.MaskBits):
            raise RuntimeError('bit-packed attention masks need the rsc_attn kernels (bf16, head dim 32, no attention dropout)')
        q = q.reshape(L, B, H, E // H).permute(1, 2, 0, 3)
        k = k.reshape(S, B, H, E // H).permute(1, 2, 0, 3)
        v = v.reshape(S, B, H, E // H).permute(1, 2, 0, 3)
        mask = None
        if attn_mask is not None and _CONST_ATTN_BIAS and key_padding_mask is None and hasattr(attn_mask, '_rsc_add'):
            # (RSC_CONST_ATTN_BIAS=0 turns it off) a shape-only constant mask (the denoising mask of the DINO decoder, the same
            # object for all 6 layers of every step) is turned into the additive 0 / -inf bias ONCE per dtype instead of
            # an invert + masked_fill per layer call inside the attention op
            mask = attn_mask._rsc_add.get(q.dtype)
            if mask is None:
                mask = torch.zeros(attn_mask.shape, dtype=q.dtype, device=attn_mask.device).masked_fill_(attn_mask, float('-inf'))
                attn_mask._rsc_add[q.dtype] = mask
            mask = mask.view(B, H, L, S) if mask.dim() == 3 else mask.view(1, 1, L, S)
        elif attn_mask is not None:          # bool, True = masked out (torch MHA convention)
            assert attn_mask.dtype == torch.bool, 'only boolean attention masks are used by the reference heads'
            mask = ~attn_mask
            if mask.dim() == 4:
                pass                       # (B, 1, L, S): one mask per image, broadcast over the heads by the SDPA call
            else:
                mask = mask.view(B, H, L, S) if mask.dim() == 3 else mask.view(1, 1, L, S)
        if key_padding_mask is not None:
            kp = ~key_padding_mask.view(B, 1, 1, S)
            mask = kp if mask is None else mask & kp
        out = F.scaled_dot_product_attention(q, k, v.to(q.dtype), attn_mask=mask,
                                             dropout_p=m.dropout if self.training else 0.0)
        out = out.permute(2, 0, 1, 3).reshape(L, B, E)
        if defer:
            return ops.linear(out, m.out_proj.weight, None)
        return ops.linear(out, m.out_proj.weight, m.out_proj.bias)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None, attn_mask=None,
                key_padding_mask=None, **kwargs):
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
            key_pos = query_pos
        same_qk = key is query
        if query_pos is not None:
            query = query + query_pos.to(query.dtype)
        same_qk = same_qk and key_pos is query_pos
        if same_qk:
            key = query
        elif key_pos is not None:
            key = key + key_pos.to(key.dtype)
        if self.batch_first:
            query, key, value = (t.transpose(0, 1) for t in (query, key, value))
        drop_active = self.training and (self.proj_drop.p > 0. or (
            isinstance(self.dropout_layer, nn.Dropout) and self.dropout_layer.p > 0.) or (
            isinstance(self.dropout_layer, DropPath) and self.dropout_layer.drop_prob > 0.))
        defer = kwargs.get('_defer', False) and not drop_active and self.attn.out_proj.bias is not None
        out = self._attend(query, key, value, attn_mask, key_padding_mask, same_qk, defer)
        if self.batch_first:
            out = out.transpose(0, 1)
        if defer:       # the caller fuses bias + residual + the following LayerNorm (ops.add_ln)
            return out, self.attn.out_proj.bias, identity
        return identity + self.dropout_layer(self.proj_drop(out))


@MODELS.register_module()
class MultiScaleDeformableAttention(nn.Module):
    """mmcv MultiScaleDeformableAttention; the sampling core is rsc_msda_{fwd,bwd}."""

    # both read `query`: the step engine lays their parameters out back to back, one GEMM serves both (ops.linear_pair)
    _rsc_linear_pairs = (('sampling_offsets', 'attention_weights'),)

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, im2col_step=64, dropout=0.1,
                 batch_first=False, norm_cfg=None, init_cfg=None):
        super().__init__()
        if embed_dims % num_heads != 0:
            raise ValueError('embed_dims must be divisible by num_heads, but got %d and %d' % (embed_dims, num_heads))
        self.norm_cfg = norm_cfg
        self.dropout = nn.Dropout(dropout)
        self.batch_first = batch_first
        self.im2col_step = im2col_step
        self.embed_dims = embed_dims
        self.num_levels = num_levels
        self.num_heads = num_heads
        self.num_points = num_points
        self.sampling_offsets = Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = Linear(embed_dims, embed_dims)
        self.output_proj = Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        nn.init.constant_(self.sampling_offsets.weight, 0.)
        thetas = torch.arange(self.num_heads, dtype=torch.float32) * (2.0 * math.pi / self.num_heads)
        grid_init = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid_init = (grid_init / grid_init.abs().max(-1, keepdim=True)[0]).view(self.num_heads, 1, 1, 2).repeat(
            1, self.num_levels, self.num_points, 1)
        for i in range(self.num_points):
            grid_init[:, :, i, :] *= i + 1
        self.sampling_offsets.bias.data = grid_init.view(-1)
        nn.init.constant_(self.attention_weights.weight, 0.)
        nn.init.constant_(self.attention_weights.bias, 0.)
        nn.init.xavier_uniform_(self.value_proj.weight)
        nn.init.constant_(self.value_proj.bias, 0.)
        nn.init.xavier_uniform_(self.output_proj.weight)
        nn.init.constant_(self.output_proj.bias, 0.)

    def _sample_eager(self, value, sampling_offsets, attention_weights, reference_points, spatial_shapes,
                      level_start_index):
        """mmcv's op-by-op tail: softmax, sampling locations, then the sampling kernel."""
        bs, num_query = sampling_offsets.shape[:2]
        attention_weights = attention_weights.float().softmax(-1).view(
            bs, num_query, self.num_heads, self.num_levels, self.num_points)
        sampling_offsets = sampling_offsets.float()
        reference_points = reference_points.float()
        if reference_points.shape[-1] == 2:
            offset_normalizer = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1).float()
            sampling_locations = reference_points[:, :, None, :, None, :] + \
                sampling_offsets / offset_normalizer[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            sampling_locations = reference_points[:, :, None, :, None, :2] + \
                sampling_offsets / self.num_points * reference_points[:, :, None, :, None, 2:] * 0.5
        else:
            raise ValueError('Last dim of reference_points must be 2 or 4, but get %d instead.'
                             % reference_points.shape[-1])
        return ops.ms_deform_attn(value, spatial_shapes, level_start_index, sampling_locations, attention_weights,
                                  self.im2col_step)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, level_start_index=None, **kwargs):
        if value is None:
            value = query
        if identity is None:
            identity = query
        if query_pos is not None:
            query = query + query_pos.to(query.dtype)
        batch_first = self.batch_first or kwargs.get('_batch_first', False)
        if not batch_first:
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)
        bs, num_query, _ = query.shape
        bs, num_value, _ = value.shape
        value = self.value_proj(value)
        if key_padding_mask is not None:
            value = value.masked_fill(key_padding_mask[..., None], 0.0)
        value = value.view(bs, num_value, self.num_heads, -1)
        pv = self._pair_views()
        if (ops.linear_pair_supported(query, pv) and reference_points is not None and
                self.num_levels * self.num_points == 16 and self.embed_dims // self.num_heads == 32):
            both = ops.linear_pair(query, self.sampling_offsets, self.attention_weights, pv)   # [offsets | logits]
            if ops.msda_packed_supported(value, both, reference_points, self.num_levels, self.num_points):
                output = ops.ms_deform_attn_fused_packed(value, spatial_shapes, level_start_index, both,
                                                         reference_points, self.num_levels, self.num_points)
                return self._project_out(output, identity, batch_first, kwargs)
            n_off = self.sampling_offsets.out_features
            sampling_offsets, attention_weights = both[..., :n_off], both[..., n_off:]
            sampling_offsets = sampling_offsets.reshape(bs, num_query, self.num_heads, self.num_levels, self.num_points, 2)
            attention_weights = attention_weights.reshape(bs, num_query, self.num_heads, self.num_levels * self.num_points)
            output = self._sample_eager(value, sampling_offsets, attention_weights, reference_points, spatial_shapes,
                                        level_start_index)
            return self._project_out(output, identity, batch_first, kwargs)
        sampling_offsets = self.sampling_offsets(query).view(
            bs, num_query, self.num_heads, self.num_levels, self.num_points, 2)
        attention_weights = self.attention_weights(query).view(
            bs, num_query, self.num_heads, self.num_levels * self.num_points)
        if ops.msda_fused_supported(value, sampling_offsets, reference_points):
            # softmax, sampling-location arithmetic and the sampling op in one kernel (no fp32 (B,Nq,8,4,4,2) temporaries)
            output = ops.ms_deform_attn_fused(value, spatial_shapes, level_start_index, sampling_offsets,
                                              attention_weights, reference_points)
        else:
            output = self._sample_eager(value, sampling_offsets, attention_weights, reference_points, spatial_shapes,
                                        level_start_index)
        return self._project_out(output, identity, batch_first, kwargs)

    def _pair_views(self):
        """stacked parameter views for ops.linear_pair (None without the step engine's flat layout)"""
        g = getattr(self.sampling_offsets.weight, '_rsc_g', None)
        if g is None:
            return None
        key = (g.data_ptr(), self.sampling_offsets.weight.data_ptr())
        cached = getattr(self, '_pair_cache', None)
        if cached is None or cached[0] != key:
            cached = self._pair_cache = (key, ops.linear_pair_views(self.sampling_offsets, self.attention_weights))
        return cached[1]

    def _project_out(self, output, identity, batch_first, kwargs):
        defer = kwargs.get('_defer', False) and not (self.training and self.dropout.p > 0.)
        output = ops.linear(output, self.output_proj.weight, None) if defer else self.output_proj(output)
        if not batch_first:
            output = output.permute(1, 0, 2)
        if defer:       # the caller fuses bias + residual + the following LayerNorm (ops.add_ln)
            return output, self.output_proj.bias, identity
        return self.dropout(output) + identity


@MODELS.register_module()
class BaseTransformerLayer(nn.Module):
    def __init__(self, attn_cfgs=None, ffn_cfgs=dict(type='FFN', embed_dims=256, feedforward_channels=1024, num_fcs=2,
                                                     ffn_drop=0., act_cfg=dict(type='ReLU', inplace=True)),
                 operation_order=None, norm_cfg=dict(type='LN'), init_cfg=None, batch_first=False, **kwargs):
        super().__init__()
        assert set(operation_order) <= {'self_attn', 'norm', 'ffn', 'cross_attn'}
        num_attn = operation_order.count('self_attn') + operation_order.count('cross_attn')
        if isinstance(attn_cfgs, dict):
            attn_cfgs = [copy.deepcopy(attn_cfgs) for _ in range(num_attn)]
        else:
            assert num_attn == len(attn_cfgs)
        self.num_attn = num_attn
        self.operation_order = tuple(operation_order)
        self.norm_cfg = norm_cfg
        self.pre_norm = operation_order[0] == 'norm'
        self.batch_first = batch_first
        self.attentions = nn.ModuleList()
        for cfg in attn_cfgs:
            cfg = dict(cfg)
            cfg.setdefault('batch_first', batch_first)
            self.attentions.append(build_from_cfg(cfg, MODELS))
        self.embed_dims = self.attentions[0].embed_dims
        num_ffns = operation_order.count('ffn')
        if isinstance(ffn_cfgs, dict):
            ffn_cfgs = [copy.deepcopy(ffn_cfgs) for _ in range(num_ffns)]
        self.ffns = nn.ModuleList()
        for cfg in ffn_cfgs:
            cfg = dict(cfg)
            cfg.setdefault('type', 'FFN')
            cfg['embed_dims'] = cfg.get('embed_dims', self.embed_dims) if 'embed_dims' in cfg else self.embed_dims
            self.ffns.append(build_from_cfg(cfg, MODELS))
        self.norms = nn.ModuleList([build_norm(norm_cfg, self.embed_dims) for _ in range(operation_order.count('norm'))])

    def forward(self, query, key=None, value=None, query_pos=None, key_pos=None, attn_masks=None,
                query_key_padding_mask=None, key_padding_mask=None, **kwargs):
        norm_index = attn_index = ffn_index = 0
        identity = query
        if attn_masks is None:
            attn_masks = [None for _ in range(self.num_attn)]
        elif isinstance(attn_masks, torch.Tensor):
            attn_masks = [copy.copy(attn_masks) for _ in range(self.num_attn)]
        else:
            assert len(attn_masks) == self.num_attn
        ops_ = self.operation_order
        skip_norm = False
        for pos_, layer in enumerate(ops_):
            # post-norm pairs (sub-layer, LayerNorm): bias + residual add + LayerNorm in ONE kernel (ops.add_ln);
            # the sub-layer then returns (output without bias, bias, identity) instead of adding itself
            nxt_norm = (_DEFER and not self.pre_norm and pos_ + 1 < len(ops_) and ops_[pos_ + 1] == 'norm' and
                        isinstance(self.norms[norm_index], LayerNorm) and ops.add_ln_supported(query))
            if layer == 'self_attn':
                temp_key = temp_value = query
                out = self.attentions[attn_index](
                    query, temp_key, temp_value, identity if self.pre_norm else None, query_pos=query_pos,
                    key_pos=query_pos, attn_mask=attn_masks[attn_index], key_padding_mask=query_key_padding_mask,
                    _defer=nxt_norm, **kwargs)
                attn_index += 1
            elif layer == 'norm':
                if skip_norm:
                    skip_norm = False
                else:
                    query = self.norms[norm_index](query)
                norm_index += 1
                continue
            elif layer == 'cross_attn':
                out = self.attentions[attn_index](
                    query, key, value, identity if self.pre_norm else None, query_pos=query_pos, key_pos=key_pos,
                    attn_mask=attn_masks[attn_index], key_padding_mask=key_padding_mask, _defer=nxt_norm, **kwargs)
                attn_index += 1
            elif layer == 'ffn':
                out = None
                if nxt_norm:      # the whole sub-layer + residual + LayerNorm on the tcgen05 GEMMs when it applies
                    fused = self.ffns[ffn_index].fused_add_ln(query, query, self.norms[norm_index])
                    if fused is not None:
                        query, skip_norm, out = fused[1], True, fused[1]
                if out is None:
                    out = self.ffns[ffn_index](query, identity if self.pre_norm else None, _defer=nxt_norm)
                ffn_index += 1
            if isinstance(out, tuple):
                x_, bias_, id_ = out
                norm = self.norms[norm_index]
                _, query = ops.add_ln(id_.to(x_.dtype), x_, bias_, None, norm.weight, norm.bias, norm.eps)
                skip_norm = True
            else:
                query = out
            if layer != 'ffn':
                identity = query
        return query


class TransformerLayerSequence(nn.Module):
    def __init__(self, transformerlayers=None, num_layers=None, init_cfg=None):
        super().__init__()
        if isinstance(transformerlayers, dict):
            transformerlayers = [copy.deepcopy(transformerlayers) for _ in range(num_layers)]
        else:
            assert isinstance(transformerlayers, list) and len(transformerlayers) == num_layers
        self.num_layers = num_layers
        self.layers = nn.ModuleList([build_from_cfg(dict(c), MODELS) for c in transformerlayers])
        self.embed_dims = self.layers[0].embed_dims
        self.pre_norm = self.layers[0].pre_norm

    def forward(self, query, key, value, query_pos=None, key_pos=None, attn_masks=None, query_key_padding_mask=None,
                key_padding_mask=None, **kwargs):
        for layer in self.layers:
            query = layer(query, key, value, query_pos=query_pos, key_pos=key_pos, attn_masks=attn_masks,
                          query_key_padding_mask=query_key_padding_mask, key_padding_mask=key_padding_mask, **kwargs)
        return query


@MODELS.register_module()
class DetrTransformerEncoder(TransformerLayerSequence):
    def __init__(self, *args, post_norm_cfg=dict(type='LN'), **kwargs):
        super().__init__(*args, **kwargs)
        if post_norm_cfg is not None:
            self.post_norm = build_norm(post_norm_cfg, self.embed_dims) if self.pre_norm else None
        else:
            assert not self.pre_norm
            self.post_norm = None

    def forward(self, query, key=None, value=None, query_pos=None, **kwargs):
        """mmdet DetrTransformerEncoder.forward on seq-first (N, B, C) tensors.  When every attention of the stack is
        a MultiScaleDeformableAttention (the shared deformable encoder) and B > 1, the layers run BATCH-first inside:
        mmcv permutes query / value / output in every layer, which for B > 1 is three strided copies plus batched
        (instead of plain) GEMMs per layer; here the stack is transposed once on entry and once on exit."""
        msda_only = all(isinstance(a, MultiScaleDeformableAttention) and not a.batch_first
                        for layer in self.layers for a in layer.attentions)
        if msda_only and query.shape[1] > 1 and key is None and value is None:
            q = query.transpose(0, 1).contiguous()
            pos = None if query_pos is None else query_pos.transpose(0, 1).to(q.dtype).contiguous()
            x = super().forward(q, None, None, query_pos=pos, _batch_first=True, **kwargs)
            if self.post_norm is not None:
                x = self.post_norm(x)
            return x.transpose(0, 1)
        if query_pos is not None:
            query_pos = query_pos.to(query.dtype)        # (fp32 sine encodings would promote every q + pos to fp32)
        x = super().forward(query, key, value, query_pos=query_pos, **kwargs)
        if self.post_norm is not None:
            x = self.post_norm(x)
        return x


@MODELS.register_module()
class DetrTransformerDecoder(TransformerLayerSequence):
    def __init__(self, *args, post_norm_cfg=dict(type='LN'), return_intermediate=False, **kwargs):
        super().__init__(*args, **kwargs)
        self.return_intermediate = return_intermediate
        self.post_norm = build_norm(post_norm_cfg, self.embed_dims) if post_norm_cfg is not None else None

    def forward(self, query, *args, **kwargs):
        if not self.return_intermediate:
            x = super().forward(query, *args, **kwargs)
            if self.post_norm:
                x = self.post_norm(x)[None]
            return x
        intermediate = []
        for layer in self.layers:
            query = layer(query, *args, **kwargs)
            intermediate.append(self.post_norm(query) if self.post_norm is not None else query)
        return torch.stack(intermediate)


def build_transformer_layer_sequence(cfg, default_args=None):
    if isinstance(cfg, nn.Module):
        return cfg
    return build_from_cfg(dict(cfg), MODELS, default_args)


@MODELS.register_module()
class SinePositionalEncoding(nn.Module):
    def __init__(self, num_feats, temperature=10000, normalize=False, scale=2 * math.pi, eps=1e-6, offset=0.,
                 init_cfg=None):
        super().__init__()
        if normalize:
            assert isinstance(scale, (float, int))
        self.num_feats, self.temperature, self.normalize = num_feats, temperature, normalize
        self.scale, self.eps, self.offset = scale, eps, offset

    def forward(self, mask):
        mask = mask.to(torch.int)
        not_mask = 1 - mask
        y_embed = not_mask.cumsum(1, dtype=torch.float32)
        x_embed = not_mask.cumsum(2, dtype=torch.float32)
        if self.normalize:
            y_embed = (y_embed + self.offset) / (y_embed[:, -1:, :] + self.eps) * self.scale
            x_embed = (x_embed + self.offset) / (x_embed[:, :, -1:] + self.eps) * self.scale
        dim_t = torch.arange(self.num_feats, dtype=torch.float32, device=mask.device)
        dim_t = self.temperature ** (2 * (dim_t // 2) / self.num_feats)
        pos_x = x_embed[:, :, :, None] / dim_t
        pos_y = y_embed[:, :, :, None] / dim_t
        B, H, W = mask.size()
        pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).view(B, H, W, -1)
        pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).view(B, H, W, -1)
        return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


def build_positional_encoding(cfg):
    return build_from_cfg(dict(cfg), MODELS)


def inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(min=0, max=1)
    x1 = x.clamp(min=eps)
    x2 = (1 - x).clamp(min=eps)
    return torch.log(x1 / x2)


class Conv2d(nn.Conv2d):
    """nn.Conv2d whose CUDA forward / backward is an im2col GEMM on this repo's kernels (ops.conv2d: rsc_im2col_* +
    rsc_linear_*); same parameters / state-dict keys.  CPU tensors and the cases ops.conv2d does not cover (groups, dilation,
    odd channel counts like the 3-channel image stem) use the stock implementation."""

    def forward(self, x):
        if self.padding_mode == 'zeros' and ops.conv2d_supported(x, self.weight, self.stride, self.padding, self.dilation,
                                                                  self.groups):
            return ops.conv2d(x, self.weight, self.bias, self.stride, self.padding)
        return super().forward(x)


class ConvModule(nn.Module):
    """mmcv ConvModule (conv -> norm -> act); conv bias only without norm."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias='auto', conv_cfg=None,
                 norm_cfg=None, act_cfg=dict(type='ReLU'), inplace=True):
        super().__init__()
        if bias == 'auto':
            bias = norm_cfg is None
        self.conv = Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=bias)
        self.norm_name = None
        if norm_cfg is not None:
            self.norm_name = norm_abbr(norm_cfg)
            self.add_module(self.norm_name, build_norm(norm_cfg, out_channels))
        self.activate = build_activation(act_cfg) if act_cfg is not None else None

    def forward(self, x):
        x = self.conv(x)
        if self.norm_name is not None:
            norm = getattr(self, self.norm_name)
            if isinstance(self.activate, nn.ReLU) and isinstance(norm, (GroupNorm, BatchNorm2d)):
                return norm(x, relu=True)          # conv -> (norm + ReLU in one pass)
            x = norm(x)
        if self.activate is not None:
            x = self.activate(x)
        return x


@MODELS.register_module()
class ChannelMapper(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, conv_cfg=None, norm_cfg=None,
                 act_cfg=dict(type='ReLU'), num_outs=None, init_cfg=None):
        super().__init__()
        assert isinstance(in_channels, (list, tuple))
        self.extra_convs = None
        if num_outs is None:
            num_outs = len(in_channels)
        self.convs = nn.ModuleList([
            ConvModule(c, out_channels, kernel_size, padding=(kernel_size - 1) // 2, norm_cfg=norm_cfg, act_cfg=act_cfg)
            for c in in_channels])
        if num_outs > len(in_channels):
            self.extra_convs = nn.ModuleList()
            for i in range(len(in_channels), num_outs):
                c = in_channels[-1] if i == len(in_channels) else out_channels
                self.extra_convs.append(ConvModule(c, out_channels, 3, stride=2, padding=1, norm_cfg=norm_cfg,
                                                   act_cfg=act_cfg))
        self.init_weights()

    def init_weights(self):   # mmdet init_cfg: Xavier uniform on Conv2d
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0.)

    def forward(self, inputs):
        assert len(inputs) == len(self.convs)
        outs = [self.convs[i](inputs[i]) for i in range(len(inputs))]
        if self.extra_convs:
            for i in range(len(self.extra_convs)):
                outs.append(self.extra_convs[i](inputs[-1] if i == 0 else outs[-1]))
        return tuple(outs)
