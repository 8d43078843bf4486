"""Swin Transformer backbone with the mmdet 2.25.1 `SwinTransformer` signature
and state-dict layout (reference call site
models/multi/multitask_learner.py:83, config
configs/multi/MTL_slvlcls_swin-t-p4-w7_1x1_resisc&dior&potsdam.py:9-25).

Differences from the eager mmdet module are purely structural:
  * ShiftWindowMSA's pad / roll / partition / attention / reverse / unroll /
    crop chain is ONE kernel (ops.wmsa); the qkv Linear runs on the un-padded
    tokens and padded rows are synthesised from the qkv bias inside the kernel.
  * PatchMerging's unfold + LayerNorm is one kernel (ops.patch_merge_ln).
  * feature maps are returned as logical (B,C,H,W) tensors with channels-last
    strides (a view of the token tensor; no transpose copy).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..config import MODELS
from .bricks import FFN, DropPath, LayerNorm, Linear, draw_drop_paths


class WindowMSA(nn.Module):
    def __init__(self, embed_dims, num_heads, window_size, qkv_bias=True, qk_scale=None, attn_drop_rate=0.,
                 proj_drop_rate=0.):
        super().__init__()
        assert attn_drop_rate == 0., 'attention dropout is not used by any reference config'
        self.embed_dims = embed_dims
        self.window_size = window_size  # (Wh, Ww)
        self.num_heads = num_heads
        head_embed_dims = embed_dims // num_heads
        self.scale = qk_scale or head_embed_dims ** -0.5
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * window_size[0] - 1) * (2 * window_size[1] - 1), num_heads))
        Wh, Ww = window_size
        rel_index_coords = self.double_step_seq(2 * Ww - 1, Wh, 1, Ww)
        rel_position_index = rel_index_coords + rel_index_coords.T
        self.register_buffer('relative_position_index', rel_position_index.flip(1).contiguous())
        self.qkv = Linear(embed_dims, embed_dims * 3, bias=qkv_bias)
        self.proj = Linear(embed_dims, embed_dims)
        self.proj_drop = nn.Dropout(proj_drop_rate)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)

    @staticmethod
    def double_step_seq(step1, len1, step2, len2):
        seq1 = torch.arange(0, step1 * len1, step1)
        seq2 = torch.arange(0, step2 * len2, step2)
        return (seq1[:, None] + seq2[None, :]).reshape(1, -1)


class ShiftWindowMSA(nn.Module):
    def __init__(self, embed_dims, num_heads, window_size, shift_size=0, qkv_bias=True, qk_scale=None,
                 attn_drop_rate=0, proj_drop_rate=0, dropout_layer=dict(type='DropPath', drop_prob=0.)):
        super().__init__()
        self.window_size = window_size
        self.shift_size = shift_size
        assert 0 <= self.shift_size < self.window_size
        self.w_msa = WindowMSA(embed_dims, num_heads, (window_size, window_size), qkv_bias, qk_scale, attn_drop_rate,
                               proj_drop_rate)
        self.drop = DropPath(dropout_layer.get('drop_prob', 0.))

    def forward(self, query, hw_shape):
        B, L, C = query.shape
        H, W = hw_shape
        assert L == H * W, 'input feature has wrong size'
        m = self.w_msa
        qkv = m.qkv(query)                                   # GEMM on the L real tokens
        out = ops.wmsa(qkv, m.qkv.bias, m.relative_position_bias_table, (H, W), m.num_heads, self.window_size,
                       self.shift_size, m.scale)             # fused attention core
        return m.proj_drop(m.proj(out))          # DropPath is applied by the block, fused with the residual add


class SwinBlock(nn.Module):
    def __init__(self, embed_dims, num_heads, feedforward_channels, window_size=7, shift=False, qkv_bias=True,
                 qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0., act_cfg=dict(type='GELU'),
                 with_cp=False):
        super().__init__()
        self.with_cp = with_cp
        self.norm1 = LayerNorm(embed_dims)
        self.attn = ShiftWindowMSA(embed_dims, num_heads, window_size, window_size // 2 if shift else 0, qkv_bias,
                                   qk_scale, attn_drop_rate, drop_rate, dict(type='DropPath', drop_prob=drop_path_rate))
        self.norm2 = LayerNorm(embed_dims)
        self.ffn = FFN(embed_dims, feedforward_channels, 2, act_cfg, drop_rate,
                       dict(type='DropPath', drop_prob=drop_path_rate), add_identity=True)

    def forward(self, x, hw_shape):
        identity = x
        x = self.norm1(x)
        x = self.attn.drop.add_to(identity, self.attn(x, hw_shape))
        identity = x
        x = self.norm2(x)
        return self.ffn(x, identity=identity)

    def fusable(self, x):
        """the fused residual-stream kernels cover GELU FFNs without dropout on CUDA bf16 / fp32 tokens"""
        ffn = self.ffn
        return (ops.add_ln_supported(x) and len(ffn.layers) == 3 and isinstance(ffn.layers[0][1], nn.GELU) and
                ffn.layers[0][1].approximate == 'none' and ffn.layers[0][2].p == 0. and ffn.layers[2].p == 0. and
                self.attn.w_msa.proj_drop.p == 0. and ffn.add_identity and isinstance(ffn.dropout_layer, DropPath) and
                ffn.layers[0][0].bias is not None and ffn.layers[1].bias is not None and
                self.attn.w_msa.proj.bias is not None)

    def forward_fused(self, x, hw_shape, n1, next_norm):
        """Same arithmetic as forward() with the element-wise passes fused (ops.add_ln / ops.bias_gelu):
        x is the residual stream, n1 = norm1(x) if the previous fused pass already produced it, next_norm the
        LayerNorm that consumes this block's output (the next block's norm1 or the stage's output norm) or None.
        Returns (x_out, next_norm(x_out) or None)."""
        if n1 is None:
            n1 = self.norm1(x)
        m = self.attn.w_msa
        H, W = hw_shape
        qkv = m.qkv(n1)
        a = ops.wmsa(qkv, m.qkv.bias, m.relative_position_bias_table, (H, W), m.num_heads, self.attn.window_size,
                     self.attn.shift_size, m.scale)
        s1 = self.attn.drop.scale_vec(x)
        if ops.linear_add_ln_supported(a, m.proj.weight, x):
            # proj + bias + DropPath + residual + norm2 in ONE tcgen05 GEMM (the LayerNorm lives in its epilogue)
            x1, n2 = ops.linear_add_ln(a, m.proj.weight, m.proj.bias, x, s1, self.norm2.weight, self.norm2.bias, self.norm2.eps)
        else:
            a = ops.linear(a, m.proj.weight, None)             # proj bias is added inside add_ln
            x1, n2 = ops.add_ln(x, a, m.proj.bias, s1, self.norm2.weight, self.norm2.bias, self.norm2.eps)
        fc1, fc2 = self.ffn.layers[0][0], self.ffn.layers[1]
        s2 = self.ffn.dropout_layer.scale_vec(x)
        fused_mlp = ops.mlp_supported(n2, fc1.weight, fc1.bias, fc2.weight)
        if fused_mlp and next_norm is not None and ops.linear_add_ln_supported(n2, fc2.weight, x1):
            # fc1 + GELU | fc2 + bias + DropPath + residual + the NEXT LayerNorm: two GEMM launches for the whole MLP tail
            return ops.mlp_add_ln(n2, fc1.weight, fc1.bias, fc2.weight, fc2.bias, x1, s2, next_norm.weight, next_norm.bias,
                                  ops.ACT_GELU, next_norm.eps)
        if fused_mlp:
            # both Linears on the tcgen05 GEMMs: bias + GELU in fc1's epilogue, GELU' in the epilogue of fc2's dX
            f = ops.mlp(n2, fc1.weight, fc1.bias, fc2.weight, None, ops.ACT_GELU)
        else:
            g = ops.bias_gelu(ops.linear(n2, fc1.weight, None), fc1.bias)
            f = ops.linear(g, fc2.weight, None)
        if next_norm is not None:
            return ops.add_ln(x1, f, fc2.bias, s2, next_norm.weight, next_norm.bias, next_norm.eps)
        f = f + fc2.bias.to(f.dtype)
        return (x1 + f if s2 is None else torch.addcmul(x1, f, s2.to(f.dtype).view(-1, 1, 1))), None


class PatchMerging(nn.Module):
    """mmdet PatchMerging (kernel 2, stride 2, 'corner' padding, nn.Unfold channel order)."""

    def __init__(self, in_channels, out_channels, stride=2):
        super().__init__()
        assert stride == 2
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm = nn.LayerNorm(4 * in_channels)
        self.reduction = Linear(4 * in_channels, out_channels, bias=False)

    def forward(self, x, input_size):
        H, W = input_size
        x = ops.patch_merge_ln(x, (H, W), self.norm.weight, self.norm.bias, self.norm.eps)
        return self.reduction(x), ((H + 1) // 2, (W + 1) // 2)


class SwinBlockSequence(nn.Module):
    def __init__(self, embed_dims, num_heads, feedforward_channels, depth, window_size=7, qkv_bias=True, qk_scale=None,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0., downsample=None, act_cfg=dict(type='GELU'),
                 with_cp=False):
        super().__init__()
        dprs = drop_path_rate if isinstance(drop_path_rate, (list, tuple)) else [drop_path_rate] * depth
        self.blocks = nn.ModuleList([
            SwinBlock(embed_dims, num_heads, feedforward_channels, window_size, i % 2 == 1, qkv_bias, qk_scale,
                      drop_rate, attn_drop_rate, dprs[i], act_cfg, with_cp) for i in range(depth)])
        self.downsample = downsample

    def forward(self, x, hw_shape, out_norm=None):
        """-> (x_down, down_hw, out, out_hw); with `out_norm` (the stage's output LayerNorm) `out` is already
        normalised -- on the fused path the last block's residual update and that norm are one kernel."""
        n = None
        for k, block in enumerate(self.blocks):
            if block.fusable(x):
                nxt = self.blocks[k + 1].norm1 if k + 1 < len(self.blocks) else out_norm
                x, n = block.forward_fused(x, hw_shape, n, nxt)
            else:
                x, n = block(x, hw_shape), None
        out = x
        if out_norm is not None:
            out = n if n is not None else out_norm(x)
        if self.downsample:
            x_down, down_hw_shape = self.downsample(x, hw_shape)
            return x_down, down_hw_shape, out, hw_shape
        return x, hw_shape, out, hw_shape


class PatchEmbed(nn.Module):
    def __init__(self, in_channels=3, embed_dims=96, kernel_size=4, stride=4, norm=True):
        super().__init__()
        self.kernel = kernel_size
        self.projection = nn.Conv2d(in_channels, embed_dims, kernel_size, stride)
        self.norm = LayerNorm(embed_dims) if norm else None

    def forward(self, x):
        H, W = x.shape[-2:]
        ph, pw = (self.kernel - H % self.kernel) % self.kernel, (self.kernel - W % self.kernel) % self.kernel
        if ph or pw:
            x = F.pad(x, (0, pw, 0, ph))     # AdaptivePadding('corner')
        conv = self.projection
        if (x.is_cuda and self.kernel == 4 and conv.stride == (4, 4) and conv.padding == (0, 0) and
                x.dtype in (torch.float32, torch.bfloat16) and not x.requires_grad):
            # non-overlapping patches: one gather (rsc_patchify4) + ONE plain GEMM with K = 48 instead of a
            # channels-last conversion of the whole batch + an implicit-GEMM convolution
            B, hw = x.shape[0], (x.shape[2] // 4, x.shape[3] // 4)
            x = ops.linear(ops.patchify4(x), conv.weight, conv.bias).view(B, hw[0] * hw[1], conv.out_channels)
        else:
            x = conv(x.contiguous(memory_format=torch.channels_last))
            hw = (x.shape[2], x.shape[3])
            x = x.permute(0, 2, 3, 1).reshape(x.shape[0], hw[0] * hw[1], x.shape[1])
        if self.norm is not None:
            x = self.norm(x)
        return x, hw


@MODELS.register_module()
class SwinTransformer(nn.Module):
    ARCH = {'tiny': (96, (2, 2, 6, 2), (3, 6, 12, 24)), 'small': (96, (2, 2, 18, 2), (3, 6, 12, 24)),
            'base': (128, (2, 2, 18, 2), (4, 8, 16, 32))}

    def __init__(self, pretrain_img_size=224, in_channels=3, embed_dims=96, patch_size=4, window_size=7, mlp_ratio=4,
                 depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), strides=(4, 2, 2, 2), out_indices=(0, 1, 2, 3),
                 qkv_bias=True, qk_scale=None, patch_norm=True, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.1,
                 use_abs_pos_embed=False, act_cfg=dict(type='GELU'), norm_cfg=dict(type='LN'), with_cp=False,
                 pretrained=None, convert_weights=False, frozen_stages=-1, init_cfg=None, arch=None, img_size=None):
        super().__init__()
        if arch is not None:       # mmcls-style signature (configs/cls/*)
            embed_dims, depths, num_heads = self.ARCH[arch]
            out_indices = (3,) if out_indices == (0, 1, 2, 3) else out_indices
        assert not use_abs_pos_embed, 'no reference config uses the absolute position embedding'
        assert strides[0] == patch_size
        self.out_indices = tuple(out_indices)
        self.convert_weights = convert_weights
        self.init_cfg = init_cfg
        self.pretrained = pretrained
        self.patch_embed = PatchEmbed(in_channels, embed_dims, patch_size, strides[0], patch_norm)
        self.drop_after_pos = nn.Dropout(p=drop_rate)
        total_depth = sum(depths)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, total_depth)]
        self.stages = nn.ModuleList()
        in_ch = embed_dims
        for i, depth in enumerate(depths):
            downsample = PatchMerging(in_ch, 2 * in_ch, strides[i + 1]) if i < len(depths) - 1 else None
            self.stages.append(SwinBlockSequence(
                in_ch, num_heads[i], mlp_ratio * in_ch, depth, window_size, qkv_bias, qk_scale, drop_rate,
                attn_drop_rate, dpr[sum(depths[:i]):sum(depths[:i + 1])], downsample, act_cfg, with_cp))
            if downsample:
                in_ch = downsample.out_channels
        self.num_features = [int(embed_dims * 2 ** i) for i in range(len(depths))]
        for i in self.out_indices:
            self.add_module('norm%d' % i, LayerNorm(self.num_features[i]))
        self._random_init()

    def init_weights(self):
        """mmdet SwinTransformer.init_weights: init_cfg=None -> trunc_normal .02 linears, unit LayerNorm;
        init_cfg=dict(type='Pretrained', checkpoint=...) -> load (and, with convert_weights, convert) the
        checkpoint (mtl/utils/checkpoint.py).  The reference cfg names a URL; there is no network here, so a
        checkpoint that is not a local file leaves the random init in place and says so."""
        self._random_init()
        ckpt = self.pretrained
        if isinstance(self.init_cfg, dict) and self.init_cfg.get('type') == 'Pretrained':
            ckpt = self.init_cfg.get('checkpoint', ckpt)
        if ckpt:
            import os
            if os.path.isfile(str(ckpt)):
                from ..mtl.utils.checkpoint import load_swin_pretrained
                load_swin_pretrained(self, ckpt, self.convert_weights)
            else:
                import warnings
                warnings.warn('backbone checkpoint %r is not a local file (no network): keeping the random init' % (ckpt,))

    def _random_init(self):
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

    def forward(self, x):
        x, hw_shape = self.patch_embed(x)
        x = self.drop_after_pos(x)
        if self.training:
            if not hasattr(self, '_drop_paths'):
                self._drop_paths = [m for m in self.modules() if isinstance(m, DropPath)]
            draw_drop_paths(self._drop_paths, x.shape[0], x.device, torch.float32)
        outs = []
        # data-parallel step engine (mtl/engine/step.py::_on_grad_ready): told from inside backward when the
        # gradients of stage i and everything after it are final, so their all-reduce overlaps the earlier stages
        ready = getattr(self, '_grad_ready_cb', None) if torch.is_grad_enabled() else None
        for i, stage in enumerate(self.stages):
            if ready is not None and i > 0 and x.requires_grad:
                x.register_hook(lambda g, i=i: ready(i))
            x, hw_shape, out, out_hw_shape = stage(x, hw_shape,
                                                   getattr(self, 'norm%d' % i) if i in self.out_indices else None)
            if i in self.out_indices:
                # logical NCHW, channels-last strides: a view, no transpose copy
                out = out.view(-1, *out_hw_shape, self.num_features[i]).permute(0, 3, 1, 2)
                if ready is not None and out.requires_grad:
                    out.register_hook(lambda g: ready('outs'))
                outs.append(out)
        return outs
