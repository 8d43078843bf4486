"""Oracle: mmdet 2.25.1 ``SwinTransformer`` restated functionally (TEST INFRA).

Reference call site: ``models/multi/multitask_learner.py:83`` (``self.backbone(img)``)
with the backbone built by ``mtl/model/build.py:7-16`` from
``configs/multi/MTL_slvlcls_swin-t-p4-w7_1x1_resisc&dior&potsdam.py:9-25``.
The arithmetic is mmdet's ``models/backbones/swin.py`` + ``models/utils/transformer.py``
(not vendored; restated per SURVEY.md Appendix A / D.1).

All functions take a flat ``sd`` (state dict: key -> fp32 tensor) and a key
prefix, so the key layout of SURVEY.md section 8b is exercised as well.
"""
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# window index maps (integer; must be bit exact)           SURVEY 8a row a5
# ----------------------------------------------------------------------------
def window_partition(x, ws):
    """(B,H,W,C) -> (B*nW, ws, ws, C); H, W multiples of ws (mmdet swin.py)."""
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def window_reverse(windows, H, W, ws):
    """(B*nW, ws, ws, C) -> (B,H,W,C)."""
    B = int(windows.shape[0] / (H * W / ws / ws))
    x = windows.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def relative_position_index(ws):
    """mmdet WindowMSA.double_step_seq construction."""
    seq1 = torch.arange(0, (2 * ws - 1) * ws, 2 * ws - 1)
    seq2 = torch.arange(0, ws, 1)
    coords = (seq1[:, None] + seq2[None, :]).reshape(1, -1)
    idx = coords + coords.T
    return idx.flip(1).contiguous()


def shift_attn_mask(Hp, Wp, ws, shift, dtype=torch.float32):
    """(nW, ws*ws, ws*ws) additive mask of ShiftWindowMSA (0 / -100.0)."""
    img_mask = torch.zeros(1, Hp, Wp, 1, dtype=dtype)
    cnt = 0
    for h in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for w in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img_mask[:, h, w, :] = cnt
            cnt += 1
    mw = window_partition(img_mask, ws).view(-1, ws * ws)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    am = am.masked_fill(am != 0, -100.0).masked_fill(am == 0, 0.0)
    return am


def window_token_index(B, H, W, ws, shift):
    """int64 map (B*nW*ws*ws,) : source flat token index into (B,H,W) for every
    slot of the padded+shifted+partitioned window tensor, or -1 for a padded
    slot.  This is the composition pad -> roll(-shift) -> window_partition that
    ShiftWindowMSA applies, expressed on indices (exactness oracle for
    rsc_window_index_partition)."""
    pad_r = (ws - W % ws) % ws
    pad_b = (ws - H % ws) % ws
    idx = torch.arange(B * H * W, dtype=torch.int64).view(B, H, W, 1)
    idx = F.pad(idx + 1, (0, 0, 0, pad_r, 0, pad_b)) - 1   # padded slots -> -1
    if shift > 0:
        idx = torch.roll(idx, shifts=(-shift, -shift), dims=(1, 2))
    return window_partition(idx, ws).reshape(-1)


# ----------------------------------------------------------------------------
# WindowMSA / ShiftWindowMSA                               SURVEY 8a rows a3, a4
# ----------------------------------------------------------------------------
def window_msa(sd, pre, x, num_heads, ws, mask=None):
    """x: (B_, N, C) windows. pre: '...attn.w_msa.'"""
    B_, N, C = x.shape
    hd = C // num_heads
    scale = hd ** -0.5
    qkv = F.linear(x, sd[pre + 'qkv.weight'], sd.get(pre + 'qkv.bias'))
    qkv = qkv.reshape(B_, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * scale
    attn = q @ k.transpose(-2, -1)
    index = relative_position_index(ws)
    bias = sd[pre + 'relative_position_bias_table'][index.view(-1)].view(
        N, N, -1).permute(2, 0, 1).contiguous()
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(B_ // nW, nW, num_heads, N, N) + \
            mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, num_heads, N, N)
    attn = attn.softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B_, N, C)
    return F.linear(x, sd[pre + 'proj.weight'], sd[pre + 'proj.bias'])


def shift_window_msa(sd, pre, query, hw, num_heads, ws, shift):
    """query: (B, L, C). pre: '...attn.'  (mmdet ShiftWindowMSA.forward)"""
    B, L, C = query.shape
    H, W = hw
    assert L == H * W
    q = query.view(B, H, W, C)
    pad_r = (ws - W % ws) % ws
    pad_b = (ws - H % ws) % ws
    q = F.pad(q, (0, 0, 0, pad_r, 0, pad_b))
    Hp, Wp = q.shape[1], q.shape[2]
    if shift > 0:
        q = torch.roll(q, shifts=(-shift, -shift), dims=(1, 2))
        mask = shift_attn_mask(Hp, Wp, ws, shift, q.dtype)
    else:
        mask = None
    win = window_partition(q, ws).view(-1, ws * ws, C)
    out = window_msa(sd, pre + 'w_msa.', win, num_heads, ws, mask)
    out = out.view(-1, ws, ws, C)
    x = window_reverse(out, Hp, Wp, ws)
    if shift > 0:
        x = torch.roll(x, shifts=(shift, shift), dims=(1, 2))
    if pad_r > 0 or pad_b:
        x = x[:, :H, :W, :].contiguous()
    return x.view(B, H * W, C)


def drop_path(x, keep_mask):
    """DropPath with an explicit per-sample keep mask (B,) scaled by 1/keep_prob
    already (parity runs inject the mask; None = identity / eval)."""
    if keep_mask is None:
        return x
    return x * keep_mask.view(-1, *([1] * (x.dim() - 1)))


def swin_block(sd, pre, x, hw, num_heads, ws, shift, dp=None):
    """mmdet SwinBlock.forward; dp = optional (mask_attn, mask_ffn)."""
    C = x.shape[-1]
    identity = x
    y = F.layer_norm(x, (C,), sd[pre + 'norm1.weight'], sd[pre + 'norm1.bias'])
    y = shift_window_msa(sd, pre + 'attn.', y, hw, num_heads, ws, shift)
    x = identity + drop_path(y, None if dp is None else dp[0])
    identity = x
    y = F.layer_norm(x, (C,), sd[pre + 'norm2.weight'], sd[pre + 'norm2.bias'])
    y = F.linear(y, sd[pre + 'ffn.layers.0.0.weight'], sd[pre + 'ffn.layers.0.0.bias'])
    y = F.gelu(y)
    y = F.linear(y, sd[pre + 'ffn.layers.1.weight'], sd[pre + 'ffn.layers.1.bias'])
    return identity + drop_path(y, None if dp is None else dp[1])


# ----------------------------------------------------------------------------
# PatchEmbed / PatchMerging                                SURVEY 8a rows a1, a6
# ----------------------------------------------------------------------------
def patch_embed(sd, pre, img, patch=4):
    """Conv2d(k=s=patch) with 'corner' adaptive padding, flatten, LayerNorm."""
    H, W = img.shape[-2:]
    pad_h = (patch - H % patch) % patch
    pad_w = (patch - W % patch) % patch
    if pad_h or pad_w:
        img = F.pad(img, (0, pad_w, 0, pad_h))
    x = F.conv2d(img, sd[pre + 'projection.weight'], sd[pre + 'projection.bias'],
                 stride=patch)
    hw = (x.shape[2], x.shape[3])
    x = x.flatten(2).transpose(1, 2)
    if pre + 'norm.weight' in sd:
        x = F.layer_norm(x, (x.shape[-1],), sd[pre + 'norm.weight'], sd[pre + 'norm.bias'])
    return x, hw


def patch_merging(sd, pre, x, hw):
    """mmdet PatchMerging: nn.Unfold(2, stride 2) channel order c*4+kh*2+kw."""
    B, L, C = x.shape
    H, W = hw
    x = x.view(B, H, W, C).permute(0, 3, 1, 2)
    if H % 2 or W % 2:
        x = F.pad(x, (0, W % 2, 0, H % 2))
        H, W = x.shape[-2:]
    x = F.unfold(x, kernel_size=2, stride=2)            # (B, 4C, H/2*W/2)
    x = x.transpose(1, 2)
    x = F.layer_norm(x, (4 * C,), sd[pre + 'norm.weight'], sd[pre + 'norm.bias'])
    x = F.linear(x, sd[pre + 'reduction.weight'])
    return x, (H // 2, W // 2)


# ----------------------------------------------------------------------------
# whole backbone                                           SURVEY 8a rows a1-a7
# ----------------------------------------------------------------------------
def swin_transformer(sd, img, *, pre='backbone.', depths=(2, 2, 6, 2),
                     num_heads=(3, 6, 12, 24), window_size=7, patch_size=4,
                     out_indices=(0, 1, 2, 3), drop_path_masks=None):
    """Returns list of (B, C_i, H_i, W_i) maps (mmdet SwinTransformer.forward).

    drop_path_masks: optional list (one per block, network order) of
    (mask_attn, mask_ffn) each (B,) already divided by keep-prob."""
    x, hw = patch_embed(sd, pre + 'patch_embed.', img, patch_size)
    outs = []
    blk = 0
    for i, depth in enumerate(depths):
        for j in range(depth):
            shift = window_size // 2 if j % 2 == 1 else 0
            dp = None if drop_path_masks is None else drop_path_masks[blk]
            x = swin_block(sd, f'{pre}stages.{i}.blocks.{j}.', x, hw,
                           num_heads[i], window_size, shift, dp)
            blk += 1
        out, out_hw = x, hw
        if i < len(depths) - 1:
            x, hw = patch_merging(sd, f'{pre}stages.{i}.downsample.', x, hw)
        if i in out_indices:
            C = out.shape[-1]
            o = F.layer_norm(out, (C,), sd[f'{pre}norm{i}.weight'], sd[f'{pre}norm{i}.bias'])
            outs.append(o.view(-1, *out_hw, C).permute(0, 3, 1, 2).contiguous())
    return outs


def swin_init_state(embed_dims=96, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24),
                    window_size=7, patch_size=4, mlp_ratio=4, in_channels=3,
                    pre='backbone.', generator=None, out_indices=(0, 1, 2, 3)):
    """Random-init state dict with the mmdet key layout (trunc_normal .02 for
    linears / tables, zero biases, unit norms), used for fixtures/benches."""
    g = generator
    sd = {}

    def tn(*shape, std=0.02):
        t = torch.empty(*shape)
        torch.nn.init.trunc_normal_(t, std=std, generator=g)
        return t

    sd[pre + 'patch_embed.projection.weight'] = tn(embed_dims, in_channels, patch_size, patch_size)
    sd[pre + 'patch_embed.projection.bias'] = torch.zeros(embed_dims)
    sd[pre + 'patch_embed.norm.weight'] = torch.ones(embed_dims)
    sd[pre + 'patch_embed.norm.bias'] = torch.zeros(embed_dims)
    for i, depth in enumerate(depths):
        C = embed_dims * 2 ** i
        for j in range(depth):
            p = f'{pre}stages.{i}.blocks.{j}.'
            for n in ('norm1', 'norm2'):
                sd[p + n + '.weight'] = torch.ones(C)
                sd[p + n + '.bias'] = torch.zeros(C)
            w = p + 'attn.w_msa.'
            sd[w + 'relative_position_bias_table'] = tn((2 * window_size - 1) ** 2, num_heads[i])
            sd[w + 'relative_position_index'] = relative_position_index(window_size)
            sd[w + 'qkv.weight'] = tn(3 * C, C)
            sd[w + 'qkv.bias'] = tn(3 * C)       # non-zero so padded rows matter
            sd[w + 'proj.weight'] = tn(C, C)
            sd[w + 'proj.bias'] = torch.zeros(C)
            sd[p + 'ffn.layers.0.0.weight'] = tn(mlp_ratio * C, C)
            sd[p + 'ffn.layers.0.0.bias'] = torch.zeros(mlp_ratio * C)
            sd[p + 'ffn.layers.1.weight'] = tn(C, mlp_ratio * C)
            sd[p + 'ffn.layers.1.bias'] = torch.zeros(C)
        if i < len(depths) - 1:
            p = f'{pre}stages.{i}.downsample.'
            sd[p + 'norm.weight'] = torch.ones(4 * C)
            sd[p + 'norm.bias'] = torch.zeros(4 * C)
            sd[p + 'reduction.weight'] = tn(2 * C, 4 * C)
        if i in out_indices:
            sd[f'{pre}norm{i}.weight'] = torch.ones(C)
            sd[f'{pre}norm{i}.bias'] = torch.zeros(C)
    return sd


# ----------------------------------------------------------------------------
# attention core on an externally computed qkv tensor (oracle for rsc_wmsa_*)
# ----------------------------------------------------------------------------
def wmsa_core(qkv, qkv_bias, table, hw, num_heads, ws=7, shift=0, scale=None):
    """qkv (B, H*W, 3C) of the UN-padded tokens -> (B, H*W, C).

    Same chain as shift_window_msa between the qkv and proj Linears: a padded
    token is a zero row after norm1, so its qkv row equals the qkv bias (zeros
    if the Linear has no bias)."""
    B, L, C3 = qkv.shape
    C = C3 // 3
    H, W = hw
    hd = C // num_heads
    if scale is None:
        scale = hd ** -0.5
    x = qkv.view(B, H, W, C3)
    pad_r = (ws - W % ws) % ws
    pad_b = (ws - H % ws) % ws
    if pad_r or pad_b:
        fill = qkv_bias if qkv_bias is not None else qkv.new_zeros(C3)
        xp = fill.to(qkv.dtype).view(1, 1, 1, C3).expand(B, H + pad_b, W + pad_r, C3).clone()
        xp[:, :H, :W] = x
        x = xp
    Hp, Wp = x.shape[1], x.shape[2]
    mask = None
    if shift > 0:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
        mask = shift_attn_mask(Hp, Wp, ws, shift, x.dtype)
    win = window_partition(x, ws).view(-1, ws * ws, C3)
    B_, N = win.shape[0], ws * ws
    q, k, v = win.reshape(B_, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    attn = (q * scale) @ k.transpose(-2, -1)
    index = relative_position_index(ws)
    attn = attn + table[index.view(-1)].view(N, N, -1).permute(2, 0, 1).unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(B_ // nW, nW, num_heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, num_heads, N, N)
    out = (attn.softmax(-1) @ v).transpose(1, 2).reshape(B_, ws, ws, C)
    y = window_reverse(out, Hp, Wp, ws)
    if shift > 0:
        y = torch.roll(y, shifts=(shift, shift), dims=(1, 2))
    return y[:, :H, :W, :].reshape(B, H * W, C)


def patch_merge_ln(x, hw, gamma, beta, eps=1e-5):
    """unfold(2,2) in nn.Unfold channel order + LayerNorm(4C) (oracle for rsc_patch_merge_ln_*)."""
    B, L, C = x.shape
    H, W = hw
    x = x.view(B, H, W, C).permute(0, 3, 1, 2)
    if H % 2 or W % 2:
        x = F.pad(x, (0, W % 2, 0, H % 2))
    x = F.unfold(x, kernel_size=2, stride=2).transpose(1, 2)
    return F.layer_norm(x, (4 * C,), gamma, beta, eps)
