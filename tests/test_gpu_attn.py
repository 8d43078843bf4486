"""GPU parity of the decoder attention kernels (csrc/attn.cu) through the C ABI / rscotr_b200.ops:
rsc_attn_fwd / rsc_attn_bwd against fp32 softmax attention on the same bf16-rounded operands, at the shapes of the
DINO decoder self-attention (about 1100 x 1100, constant denoising mask) and of the Mask2Former-style seg decoder
(100 queries x 100^2 / 50^2 / 25^2 keys, mask from mask_pred), plus ragged sizes; rsc_m2f_mask_bits against
F.interpolate + sigmoid < 0.5 + the all-masked-row rule (mask2former_head.py:134-139,177-178); the module
(bricks.MultiheadAttention, bf16 CUDA) against the oracle's mmcv MultiheadAttention restatement (fp32 CPU).
Tolerances: outputs / gradients are bf16 (relative rounding 2^-9) and P is rounded to bf16 before the PV / dV / dK / dQ
MMAs, hence 2e-2 relative norm (the same bar as the bf16 window-attention tests)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def ref_attention(q, k, v, heads, masked=None):
    """fp32: q (Lq,B,E), k / v (Lk,B,E); masked (B|1, Lq, Lk) bool, True = not attended"""
    Lq, B, E = q.shape
    Lk, d = k.shape[0], E // heads
    qh = q.float().view(Lq, B, heads, d).permute(1, 2, 0, 3)
    kh = k.float().view(Lk, B, heads, d).permute(1, 2, 0, 3)
    vh = v.float().view(Lk, B, heads, d).permute(1, 2, 0, 3)
    s = qh @ kh.transpose(-1, -2) * d ** -0.5
    if masked is not None:
        s = s.masked_fill(masked[:, None], float('-inf'))
    p = torch.softmax(s, -1)
    p = torch.nan_to_num(p, nan=0.0)             # fully masked rows: the kernel returns 0
    return (p @ vh).permute(2, 0, 1, 3).reshape(Lq, B, E)


def unpack(bits, Lk):
    b = bits.bits.cpu().to(torch.int64) & 0xffffffff
    sh = torch.arange(32)
    return ((b[..., None] >> sh) & 1).flatten(-2)[..., :Lk].bool()


def dn_mask(Lq, pad, groups):
    m = torch.zeros(Lq, Lq, dtype=torch.bool)
    m[pad:, :pad] = True
    single = pad // groups
    for i in range(groups):
        lo, hi = single * i, single * (i + 1)
        m[lo:hi, hi:pad] = True
        m[lo:hi, :lo] = True
    return m


CASES = [
    # (Lq, Lk, B, heads, mask kind)
    (1100, 1100, 1, 8, 'dn'),          # DINO decoder self-attention (900 queries + 200 denoising)
    (100, 10000, 2, 8, 'rand'),        # seg decoder cross-attention, stride-8 level (keys split over CTAs)
    (100, 2500, 2, 8, 'rand'),
    (100, 625, 2, 8, 'rand'),
    (100, 100, 2, 8, None),            # seg decoder self-attention
    (37, 70, 3, 4, 'rand'),            # ragged
    (64, 64, 1, 1, None),
    (65, 129, 1, 2, 'deadrow'),        # a fully masked query row -> zeros, no NaN
    (900, 900, 2, 8, None),
]


def _inputs(Lq, Lk, B, heads, kind, seed=0):
    g = torch.Generator().manual_seed(seed)
    E = heads * 32
    q = torch.randn(Lq, B, E, generator=g).bfloat16()
    k = torch.randn(Lk, B, E, generator=g).bfloat16()
    v = torch.randn(Lk, B, E, generator=g).bfloat16()
    masked = None
    if kind == 'dn':
        masked = dn_mask(Lq, 200, 10)[None]
    elif kind == 'rand':
        masked = torch.rand(B, Lq, Lk, generator=g) < 0.6
        masked[:, :, 0] = False
    elif kind == 'deadrow':
        masked = torch.rand(B, Lq, Lk, generator=g) < 0.3
        masked[:, 5] = True
        masked[:, 64] = True
    return q, k, v, masked


@pytest.mark.parametrize('Lq,Lk,B,heads,kind', CASES)
def test_attention_forward_backward_vs_fp32(Lq, Lk, B, heads, kind):
    from rscotr_b200 import ops
    q, k, v, masked = _inputs(Lq, Lk, B, heads, kind)
    qr, kr, vr = (t.float().requires_grad_() for t in (q, k, v))
    want = ref_attention(qr, kr, vr, heads, masked)
    g = torch.Generator().manual_seed(1)
    dout = torch.randn(want.shape, generator=g).bfloat16()
    want.backward(dout.float())
    qc, kc, vc = (t.cuda().requires_grad_() for t in (q, k, v))
    bits = None
    if masked is not None:
        bits = ops.pack_mask_bits(masked.cuda() if masked.shape[0] > 1 else masked[0].cuda())
        assert torch.equal(unpack(bits, Lk).view(masked.shape), masked)
    got = ops.attention(qc, kc, vc, heads, bits)
    got.backward(dout.cuda())
    assert torch.isfinite(got).all()
    assert rel(got, want) < 2e-2, rel(got, want)
    for name, a, b in (('dq', qc.grad, qr.grad), ('dk', kc.grad, kr.grad), ('dv', vc.grad, vr.grad)):
        assert torch.isfinite(a).all(), name
        assert rel(a, b) < 2e-2, (name, rel(a, b))
    if kind == 'deadrow':
        assert float(got[5].abs().max()) == 0. and float(got[64].abs().max()) == 0.


def test_attention_packed_qk_and_column_views():
    """self-attention layout of bricks.MultiheadAttention: q | k are the halves of ONE (L, B, 2E) projection output
    (strided views); the gradient comes back as one packed tensor"""
    from rscotr_b200 import ops
    L, B, heads = 300, 2, 8
    E = heads * 32
    g = torch.Generator().manual_seed(3)
    qk = torch.randn(L, B, 2 * E, generator=g).bfloat16()
    v = torch.randn(L, B, E, generator=g).bfloat16()
    qkr, vr = qk.float().requires_grad_(), v.float().requires_grad_()
    want = ref_attention(qkr[..., :E], qkr[..., E:], vr, heads)
    dout = torch.randn(want.shape, generator=g).bfloat16()
    want.backward(dout.float())
    qkc, vc = qk.cuda().requires_grad_(), v.cuda().requires_grad_()
    got = ops.attention(None, None, vc, heads, None, packed_qk=qkc)
    got.backward(dout.cuda())
    assert rel(got, want) < 2e-2
    assert qkc.grad.shape == qk.shape and rel(qkc.grad, qkr.grad) < 2e-2 and rel(vc.grad, vr.grad) < 2e-2


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32])
@pytest.mark.parametrize('size', [(100, 100), (50, 50), (25, 25), (13, 17)])
def test_m2f_mask_bits_vs_reference_ops(size, dtype):
    """the reference's mask path (mask2former_head.py:134-139 + the loop's :177-178 rule), op for op, against the one
    fused kernel; entries whose resized logit is within 1e-3 of the threshold may differ (bf16 rounding of the
    interpolation in the reference path)"""
    from rscotr_b200 import ops
    B, Q, Hi, Wi = 2, 100, 100, 100
    g = torch.Generator().manual_seed(5)
    mask_pred = torch.randn(B, Q, Hi, Wi, generator=g)
    mask_pred[0, 3] = -mask_pred[0, 3].abs() - 0.1           # everything masked -> un-masked by the rule
    mask_pred[1, 7] = mask_pred[1, 7].abs() + 0.1            # nothing masked
    mask_pred = mask_pred.to(dtype)
    r = F.interpolate(mask_pred.float(), size, mode='bilinear', align_corners=False).flatten(2)
    want = r.sigmoid() < 0.5
    want = want & ~want.all(-1, keepdim=True)
    bits = ops.m2f_attn_mask(mask_pred.cuda(), size)
    got = unpack(bits, size[0] * size[1])
    assert got.shape == want.shape
    sure = r.abs() > 1e-3
    assert torch.equal(got[sure], want[sure])
    assert not got[0, 3].any() and not got[1, 7].any()
    assert float((got != want).float().mean()) < 1e-3


def test_multihead_attention_module_vs_oracle():
    """bricks.MultiheadAttention on the bf16 CUDA path (own in-proj GEMMs + rsc_attn core + out-proj) against the
    oracle's mmcv MultiheadAttention (torch F.multi_head_attention_forward, fp32 CPU) with the DINO denoising mask"""
    from oracle import transformer as otr
    from rscotr_b200.models import bricks
    torch.manual_seed(0)
    L, B, E, H = 420, 2, 256, 8
    m = bricks.MultiheadAttention(E, H, attn_drop=0.0, proj_drop=0.0).eval()
    sd = {'x.' + k: v.detach().clone() for k, v in m.state_dict().items()}
    q = torch.randn(L, B, E)
    pos = torch.randn(L, B, E)
    mask = dn_mask(L, 120, 6)
    want = otr.mha(sd, 'x.', q, query_pos=pos, attn_mask=mask, num_heads=H)
    mc = m.cuda()
    with torch.autocast('cuda', dtype=torch.bfloat16):
        got = mc(q.cuda().bfloat16(), query_pos=pos.cuda().bfloat16(), attn_mask=mask.cuda())
    assert rel(got, want) < 2e-2, rel(got, want)
