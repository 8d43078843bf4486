"""Pin the ms_deform_attn oracle: grid_sample branch == scalar restatement of the
CUDA kernel's arithmetic == HF transformers' independent module (CPU only)."""
import pytest
import torch

from oracle import transformer as otr


def _rand_inputs(B, Nq, H, D, shapes, P, seed, spread=0.3):
    g = torch.Generator().manual_seed(seed)
    Nv = sum(h * w for h, w in shapes)
    L = len(shapes)
    value = torch.randn(B, Nv, H, D, generator=g)
    # include out-of-range locations (negative / > 1) to exercise zero padding
    loc = torch.rand(B, Nq, H, L, P, 2, generator=g) * (1 + 2 * spread) - spread
    w = torch.rand(B, Nq, H, L, P, generator=g).flatten(-2).softmax(-1).view(B, Nq, H, L, P)
    return value, loc, w


@pytest.mark.parametrize('shapes', [[(5, 7), (3, 4)], [(6, 6), (3, 3), (2, 2), (1, 1)]])
def test_core_matches_kernel_arithmetic(shapes):
    value, loc, w = _rand_inputs(2, 9, 2, 4, shapes, 3, seed=1)
    starts = [0]
    for h, ww in shapes[:-1]:
        starts.append(starts[-1] + h * ww)
    a = otr.ms_deform_attn_core(value, shapes, loc, w)
    b = otr.ms_deform_attn_loops(value, shapes, starts, loc, w)
    torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('ref_dim', [2, 4])
def test_msda_module_matches_hf(ref_dim):
    from transformers.models.deformable_detr.configuration_deformable_detr import DeformableDetrConfig
    from transformers.models.deformable_detr.modeling_deformable_detr import (
        DeformableDetrMultiscaleDeformableAttention as HFMSDA)
    torch.manual_seed(0)
    cfg = DeformableDetrConfig(d_model=64, num_feature_levels=3, encoder_n_points=4, decoder_n_points=4,
                               encoder_attention_heads=4, decoder_attention_heads=4)
    hf = HFMSDA(cfg, num_heads=4, n_points=4).eval()
    with torch.no_grad():
        hf.sampling_offsets.weight.normal_(0, 0.05)
        hf.attention_weights.weight.normal_(0, 0.5)
        hf.attention_weights.bias.normal_(0, 0.5)
        hf.value_proj.bias.normal_(0, 0.1)
        hf.output_proj.bias.normal_(0, 0.1)
    sd = {'a.' + k: v.detach() for k, v in hf.state_dict().items()}
    shapes = [(8, 9), (4, 5), (2, 3)]
    Nv = sum(h * w for h, w in shapes)
    B, E = 2, 64
    Nq = Nv if ref_dim == 2 else 11
    q = torch.randn(B, Nq, E)
    pos = torch.randn(B, Nq, E)
    mem = q if ref_dim == 2 else torch.randn(B, Nv, E)
    ref = torch.rand(B, Nq, 3, ref_dim)
    pad = torch.zeros(B, Nv, dtype=torch.bool)
    pad[1, -7:] = True
    starts = torch.tensor([0, 72, 92])
    with torch.no_grad():
        want = hf(hidden_states=q, attention_mask=~pad, encoder_hidden_states=mem,
                  position_embeddings=pos, reference_points=ref,
                  spatial_shapes=torch.tensor(shapes), spatial_shapes_list=shapes,
                  level_start_index=starts)[0]
        got = otr.msda(sd, 'a.', q.transpose(0, 1), value=mem.transpose(0, 1),
                       identity=torch.zeros(Nq, B, E), query_pos=pos.transpose(0, 1),
                       key_padding_mask=pad, reference_points=ref, spatial_shapes=shapes,
                       level_start_index=starts, num_heads=4, num_levels=3, num_points=4).transpose(0, 1)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)


def test_sine_positional_encoding_matches_hf():
    from transformers.models.deformable_detr.modeling_deformable_detr import DeformableDetrSinePositionEmbedding as S
    m = torch.zeros(2, 6, 9, dtype=torch.bool)
    m[1, 4:, :] = True
    m[1, :, 7:] = True
    hf = S(128, temperature=20, normalize=True)
    try:
        want = hf(torch.zeros(2, 3, 6, 9), (~m).long())
    except TypeError:
        want = hf(shape=(2, 3, 6, 9), device='cpu', dtype=torch.float32, mask=(~m).long())
    got = otr.sine_positional_encoding(m, 128, 20, True)
    # HF deformable-detr uses (embed - 0.5) / (last + eps); mmdet uses offset=0 -> compare via offset arg
    got_hf_conv = otr.sine_positional_encoding(m, 128, 20, True, offset=-0.5)
    if want.dim() == 3:      # newer HF returns (B, H*W, C)
        want = want.view(2, 6, 9, -1).permute(0, 3, 1, 2)
    assert got.shape == want.shape
    torch.testing.assert_close(got_hf_conv, want, rtol=1e-5, atol=1e-5)
