"""Pin the Swin oracle against torchvision's independent implementation and
against the closed forms of SURVEY.md Appendix A (CPU only)."""
import pytest
import torch

from oracle import swin as osw


def tv_to_mmdet(tv_sd, depths=(2, 2, 6, 2)):
    """torchvision swin_t state dict -> mmdet key layout (incl. the PatchMerging
    channel permutation [x0,x1,x2,x3]-concat -> Unfold order c*4+kh*2+kw)."""
    sd = {}
    sd['backbone.patch_embed.projection.weight'] = tv_sd['features.0.0.weight']
    sd['backbone.patch_embed.projection.bias'] = tv_sd['features.0.0.bias']
    sd['backbone.patch_embed.norm.weight'] = tv_sd['features.0.2.weight']
    sd['backbone.patch_embed.norm.bias'] = tv_sd['features.0.2.bias']
    for i, depth in enumerate(depths):
        f = 1 + 2 * i
        for j in range(depth):
            s, d = f'features.{f}.{j}.', f'backbone.stages.{i}.blocks.{j}.'
            for n in ('norm1', 'norm2'):
                sd[d + n + '.weight'] = tv_sd[s + n + '.weight']
                sd[d + n + '.bias'] = tv_sd[s + n + '.bias']
            sd[d + 'attn.w_msa.relative_position_bias_table'] = tv_sd[s + 'attn.relative_position_bias_table']
            sd[d + 'attn.w_msa.qkv.weight'] = tv_sd[s + 'attn.qkv.weight']
            sd[d + 'attn.w_msa.qkv.bias'] = tv_sd[s + 'attn.qkv.bias']
            sd[d + 'attn.w_msa.proj.weight'] = tv_sd[s + 'attn.proj.weight']
            sd[d + 'attn.w_msa.proj.bias'] = tv_sd[s + 'attn.proj.bias']
            sd[d + 'ffn.layers.0.0.weight'] = tv_sd[s + 'mlp.0.weight']
            sd[d + 'ffn.layers.0.0.bias'] = tv_sd[s + 'mlp.0.bias']
            sd[d + 'ffn.layers.1.weight'] = tv_sd[s + 'mlp.3.weight']
            sd[d + 'ffn.layers.1.bias'] = tv_sd[s + 'mlp.3.bias']
        if i < len(depths) - 1:
            s, d = f'features.{f + 1}.', f'backbone.stages.{i}.downsample.'
            C = tv_sd[s + 'norm.weight'].numel() // 4
            # tv channel = k*C + c with k = kh + 2*kw ; mmdet channel = c*4 + kh*2 + kw
            perm = torch.empty(4 * C, dtype=torch.long)
            for c in range(C):
                for kh in range(2):
                    for kw in range(2):
                        perm[c * 4 + kh * 2 + kw] = (kh + 2 * kw) * C + c
            sd[d + 'norm.weight'] = tv_sd[s + 'norm.weight'][perm]
            sd[d + 'norm.bias'] = tv_sd[s + 'norm.bias'][perm]
            sd[d + 'reduction.weight'] = tv_sd[s + 'reduction.weight'][:, perm]
    sd['backbone.norm3.weight'] = tv_sd['norm.weight']
    sd['backbone.norm3.bias'] = tv_sd['norm.bias']
    return sd


@pytest.mark.parametrize('size', [(256, 256), (288, 352), (260, 300)])
def test_swin_oracle_matches_torchvision(size):
    import torchvision
    torch.manual_seed(0)
    tv = torchvision.models.swin_t(weights=None).eval()
    with torch.no_grad():   # make biases / tables non trivial
        for n, p in tv.named_parameters():
            if p.dim() == 1 and 'norm' not in n:
                p.normal_(0, 0.05)
            if 'relative_position_bias_table' in n:
                p.normal_(0, 0.5)
    sd = tv_to_mmdet(tv.state_dict())
    x = torch.randn(2, 3, *size)
    with torch.no_grad():
        ref = tv.norm(tv.features(x)).permute(0, 3, 1, 2)
        out = osw.swin_transformer(sd, x, out_indices=(3,))[0]
    assert out.shape == ref.shape
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)


def test_relative_position_index_closed_form():
    idx = osw.relative_position_index(7)
    for p in range(49):
        for q in range(49):
            ip, jp, iq, jq = p // 7, p % 7, q // 7, q % 7
            assert idx[p, q].item() == (ip - iq + 6) * 13 + (jp - jq + 6)


def test_unfold_channel_order():
    x = torch.arange(2 * 3 * 4 * 4, dtype=torch.float32).view(2, 3, 4, 4)
    u = torch.nn.functional.unfold(x, 2, stride=2)   # (2, 12, 4)
    for c in range(3):
        for kh in range(2):
            for kw in range(2):
                assert torch.equal(u[:, c * 4 + kh * 2 + kw, :].view(2, 2, 2),
                                   x[:, c, kh::2, kw::2])


@pytest.mark.parametrize('H,W,shift', [(8, 8, 0), (8, 8, 3), (25, 25, 3), (16, 21, 3), (7, 7, 0), (50, 50, 0)])
def test_window_token_index_composition(H, W, shift):
    """The index-level map equals pad -> roll -> partition on real data and
    window_reverse -> roll -> crop inverts it."""
    B, C, ws = 2, 5, 7
    x = torch.randn(B, H, W, C)
    idx = osw.window_token_index(B, H, W, ws, shift)
    flat = torch.cat([x.reshape(-1, C), torch.zeros(1, C)])       # -1 -> zero row
    got = flat[idx]
    pad_r, pad_b = (ws - W % ws) % ws, (ws - H % ws) % ws
    q = torch.nn.functional.pad(x, (0, 0, 0, pad_r, 0, pad_b))
    if shift:
        q = torch.roll(q, (-shift, -shift), (1, 2))
    want = osw.window_partition(q, ws).reshape(-1, C)
    assert torch.equal(got, want)
    # inverse
    Hp, Wp = H + pad_b, W + pad_r
    back = osw.window_reverse(want.view(-1, ws, ws, C), Hp, Wp, ws)
    if shift:
        back = torch.roll(back, (shift, shift), (1, 2))
    assert torch.equal(back[:, :H, :W], x)


def test_shift_mask_uses_padded_grid():
    m = osw.shift_attn_mask(28, 28, 7, 3)
    assert m.shape == (16, 49, 49)
    assert set(m.unique().tolist()) == {-100.0, 0.0}
    assert (m[0] == 0).all()            # interior window: single region
    assert (m[-1] != 0).any()
