"""GPU parity: every C-ABI kernel against the CPU oracle on seeded inputs.

Tolerances: integer index maps bit exact; fp32 kernels within 1e-3 relative of
the oracle (north_star); bf16 kernels against the oracle evaluated on the
bf16-rounded inputs, within bf16 rounding of the outputs.
"""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import swin as osw
from oracle import transformer as otr

pytestmark = pytest.mark.gpu


def _ops():
    from rscotr_b200 import ops
    return ops


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def assert_rel(a, b, tol, what=''):
    e = rel_err(a, b)
    assert e <= tol, '%s: relative error %.3e > %.1e' % (what, e, tol)


GEOMS = [(1, 7, 7), (2, 8, 8), (2, 16, 16), (1, 25, 25), (1, 50, 50), (2, 10, 17), (1, 14, 21), (1, 4, 5)]


# ---------------------------------------------------------------------------
# a5: window index maps, bit exact
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('B,H,W', GEOMS + [(1, 200, 200), (3, 100, 100)])
@pytest.mark.parametrize('shift', [0, 3])
def test_window_index_bit_exact(B, H, W, shift):
    ops = _ops()
    want = osw.window_token_index(B, H, W, 7, shift)
    got = ops.window_index_partition(B, H, W, 7, shift).cpu()
    assert torch.equal(got, want)
    # reverse map is the inverse on valid slots
    rev = ops.window_index_reverse(B, H, W, 7, shift).cpu()
    assert torch.equal(want[rev], torch.arange(B * H * W))


@pytest.mark.parametrize('B,H,W', GEOMS)
@pytest.mark.parametrize('shift', [0, 3])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_window_partition_reverse_exact(B, H, W, shift, dtype):
    ops = _ops()
    g = torch.Generator().manual_seed(H * 100 + W)
    x = torch.randn(B, H, W, 32, generator=g).to(dtype)
    ws = 7
    pad_r, pad_b = (ws - W % ws) % ws, (ws - H % ws) % ws
    xp = F.pad(x, (0, 0, 0, pad_r, 0, pad_b))
    if shift:
        xp = torch.roll(xp, (-shift, -shift), (1, 2))
    want = osw.window_partition(xp, ws).view(-1, ws * ws, 32)
    got = ops.window_partition(x.cuda(), ws, shift)
    assert torch.equal(got.cpu(), want)          # pure data movement: bit exact
    # reverse(partition(x)) == x
    back = ops.window_reverse(got, B, H, W, ws, shift)
    assert torch.equal(back.cpu(), x)
    # oracle reverse chain
    y = osw.window_reverse(want.view(-1, ws, ws, 32), xp.shape[1], xp.shape[2], ws)
    if shift:
        y = torch.roll(y, (shift, shift), (1, 2))
    assert torch.equal(y[:, :H, :W].contiguous(), back.cpu())


# ---------------------------------------------------------------------------
# a3/a4: fused window attention vs oracle ShiftWindowMSA (qkv / proj GEMMs are
# plain F.linear on both sides)
# ---------------------------------------------------------------------------
def _msa_state(C, heads, seed):
    g = torch.Generator().manual_seed(seed)
    sd = {
        'attn.w_msa.qkv.weight': torch.randn(3 * C, C, generator=g) * C ** -0.5,
        'attn.w_msa.qkv.bias': torch.randn(3 * C, generator=g) * 0.5,
        'attn.w_msa.proj.weight': torch.randn(C, C, generator=g) * C ** -0.5,
        'attn.w_msa.proj.bias': torch.randn(C, generator=g) * 0.1,
        'attn.w_msa.relative_position_bias_table': torch.randn(169, heads, generator=g),
    }
    return sd


def _run_msa_gpu(sd, x, hw, heads, shift, dtype):
    ops = _ops()
    p = {k: v.clone().cuda().requires_grad_(True) for k, v in sd.items()}
    xg = x.clone().cuda().to(dtype).requires_grad_(True)
    cast = (lambda t: t.to(dtype))
    qkv = F.linear(xg, cast(p['attn.w_msa.qkv.weight']), cast(p['attn.w_msa.qkv.bias']))
    o = ops.wmsa(qkv, p['attn.w_msa.qkv.bias'], p['attn.w_msa.relative_position_bias_table'], hw, heads, 7, shift)
    y = F.linear(o, cast(p['attn.w_msa.proj.weight']), cast(p['attn.w_msa.proj.bias']))
    return xg, p, y


@pytest.mark.parametrize('B,H,W', GEOMS)
@pytest.mark.parametrize('shift', [0, 3])
@pytest.mark.parametrize('heads', [3, 4])
def test_wmsa_fp32_fwd_bwd(B, H, W, shift, heads):
    C = heads * 32
    sd = _msa_state(C, heads, seed=H + W + shift)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, H * W, C, generator=g)
    gy = torch.randn(B, H * W, C, generator=g)
    # oracle
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xo = x.clone().requires_grad_(True)
    yo = osw.shift_window_msa(sdo, 'attn.', xo, (H, W), heads, 7, shift)
    yo.backward(gy)
    # kernel
    xg, p, y = _run_msa_gpu(sd, x, (H, W), heads, shift, torch.float32)
    y.backward(gy.cuda())
    assert_rel(y, yo, 1e-3, 'out')
    assert_rel(xg.grad, xo.grad, 1e-3, 'dx')
    for k in sd:
        assert_rel(p[k].grad, sdo[k].grad, 1e-3, 'd' + k)


@pytest.mark.parametrize('B,H,W', GEOMS)
@pytest.mark.parametrize('shift', [0, 3])
@pytest.mark.parametrize('heads', [3, 4])
def test_wmsa_bf16_fwd_bwd(B, H, W, shift, heads):
    """bf16 = the TMA-staged tcgen05 kernels (wmsa_tma.cu); includes odd window counts, padded windows
    (also with pad + shift > 7: padding in the second-to-last window row) and both shift masks."""
    C = heads * 32
    sd = {k: v.bfloat16().float() for k, v in _msa_state(C, heads, seed=7).items()}
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, H * W, C, generator=g).bfloat16().float()
    gy = torch.randn(B, H * W, C, generator=g).bfloat16().float()
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xo = x.clone().requires_grad_(True)
    yo = osw.shift_window_msa(sdo, 'attn.', xo, (H, W), heads, 7, shift)
    yo.backward(gy)
    xg, p, y = _run_msa_gpu(sd, x, (H, W), heads, shift, torch.bfloat16)
    y.backward(gy.cuda().bfloat16())
    assert_rel(y, yo, 2e-2, 'out')
    assert_rel(xg.grad, xo.grad, 3e-2, 'dx')
    for k in sd:       # incl. the qkv-bias gradient that flows through the padded rows
        assert_rel(p[k].grad, sdo[k].grad, 3e-2, 'd' + k)


# Geometries at which every CTA of the tcgen05 kernels walks MANY (window, head) units (the persistent loop with its
# prefetch, buffer flip, phase carry and the cross-unit bias-gradient accumulation): the T800 stage-0..3 maps
# (200->203, 100->105, 50->56, 25->28 padded) at their head counts.  40 368 units at (16,200,200,3) is the bench
# launch; (1,200,200,3) = 2 523 units on <= 888 CTAs already loops, (2,100,100,3) = 1 350, (2,50,50,12) = 1 536,
# (4,25,25,24) = 1 536.
PERSISTENT_GEOMS = [(2, 100, 100, 3), (1, 200, 200, 3), (2, 50, 50, 12), (4, 25, 25, 24), (3, 100, 100, 6)]


@pytest.mark.parametrize('B,H,W,heads', PERSISTENT_GEOMS)
@pytest.mark.parametrize('shift', [0, 3])
def test_wmsa_bf16_persistent_loop_vs_oracle(B, H, W, heads, shift):
    """The production (bf16 tensor-core) window attention at BASELINE stage geometries: forward and ALL gradients
    (dx through dqkv, d(qkv weight / bias), d(bias table), d(proj)) against the CPU oracle."""
    C = heads * 32
    sd = {k: v.bfloat16().float() for k, v in _msa_state(C, heads, seed=11 + heads).items()}
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, H * W, C, generator=g).bfloat16().float()
    gy = torch.randn(B, H * W, C, generator=g).bfloat16().float()
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xo = x.clone().requires_grad_(True)
    yo = osw.shift_window_msa(sdo, 'attn.', xo, (H, W), heads, 7, shift)
    yo.backward(gy)
    xg, p, y = _run_msa_gpu(sd, x, (H, W), heads, shift, torch.bfloat16)
    y.backward(gy.cuda().bfloat16())
    assert_rel(y, yo, 2e-2, 'out')
    assert_rel(xg.grad, xo.grad, 3e-2, 'dx')
    for k in sd:
        assert_rel(p[k].grad, sdo[k].grad, 3e-2, 'd' + k)


@pytest.mark.parametrize('B,H,W,heads', [(1, 200, 200, 3), (2, 50, 50, 12)])
@pytest.mark.parametrize('shift', [0, 3])
def test_wmsa_bf16_core_tight(B, H, W, heads, shift):
    """The attention core alone (no GEMMs around it) against the oracle core evaluated in fp32 on the same
    bf16-rounded q|k|v: only the kernel's own roundings (P and the outputs to bf16) remain -> 1e-2 of the output
    norm and elementwise |err| <= 3 bf16 ulps of the largest output."""
    ops = _ops()
    C = heads * 32
    g = torch.Generator().manual_seed(13)
    qkv = torch.randn(B, H * W, 3 * C, generator=g).bfloat16()
    bias = (torch.randn(3 * C, generator=g) * 0.5)
    table = torch.randn(169, heads, generator=g)
    gy = torch.randn(B, H * W, C, generator=g).bfloat16()
    qo = qkv.float().clone().requires_grad_(True)
    bo = bias.clone().requires_grad_(True)
    to = table.clone().requires_grad_(True)
    want = osw.wmsa_core(qo, bo, to, (H, W), heads, 7, shift)
    want.backward(gy.float())
    qg = qkv.cuda().requires_grad_(True)
    bg = bias.cuda().requires_grad_(True)
    tg = table.cuda().requires_grad_(True)
    got = ops.wmsa(qg, bg, tg, (H, W), heads, 7, shift)
    got.backward(gy.cuda())
    assert_rel(got, want, 1e-2, 'out')
    assert (got.float().cpu() - want).abs().max() <= 3 * 2 ** -8 * want.abs().max()
    assert_rel(qg.grad, qo.grad, 1.5e-2, 'dqkv')
    assert_rel(tg.grad, to.grad, 1.5e-2, 'd table')
    if H % 7 or W % 7:
        assert_rel(bg.grad, bo.grad, 2e-2, 'd qkv bias (padded rows)')


def test_wmsa_large_property():
    """BASELINE size (stage 0 of 800^2: 200x200, C=96): shifting the INPUT by one
    whole window (7 tokens) along W on a window-aligned grid permutes windows, so
    the un-shifted attention output must be the same permutation (size
    independent property; no oracle needed)."""
    ops = _ops()
    torch.manual_seed(0)
    B, H, W, C, heads = 2, 196, 196, 96, 3
    qkv = torch.randn(B, H, W, 3 * C, device='cuda')
    table = torch.randn(169, heads, device='cuda')
    o1 = ops.wmsa(qkv.view(B, H * W, -1), None, table, (H, W), heads, 7, 0).view(B, H, W, C)
    o2 = ops.wmsa(torch.roll(qkv, 7, 2).contiguous().view(B, H * W, -1), None, table, (H, W), heads, 7, 0)
    assert torch.equal(torch.roll(o1, 7, 2), o2.view(B, H, W, C))
    # rows of softmax sum to one: with v == 1 the output is exactly ~1
    qkv[..., 2 * C:] = 1.0
    o3 = ops.wmsa(qkv.view(B, H * W, -1), None, table, (H, W), heads, 7, 3)
    torch.testing.assert_close(o3, torch.ones_like(o3), rtol=1e-5, atol=1e-5)


# ---------------------------------------------------------------------------
# a6: PatchMerging gather + LN
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('B,H,W,C', [(2, 8, 8, 96), (1, 9, 7, 96), (2, 10, 6, 192), (1, 5, 5, 384), (1, 4, 4, 512),
                                     (1, 6, 6, 128)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('variant', [1, 2])
def test_patch_merge_ln(B, H, W, C, dtype, variant, pm_variant):
    ops = _ops()
    pm_variant(variant)
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(B, H * W, C, generator=g)
    gamma = 1 + 0.1 * torch.randn(4 * C, generator=g)
    beta = 0.1 * torch.randn(4 * C, generator=g)
    red = torch.randn(2 * C, 4 * C, generator=g) * (4 * C) ** -0.5
    if dtype == torch.bfloat16:
        x, red = x.bfloat16().float(), red.bfloat16().float()
    sd = {'d.norm.weight': gamma.clone().requires_grad_(True), 'd.norm.bias': beta.clone().requires_grad_(True),
          'd.reduction.weight': red.clone().requires_grad_(True)}
    xo = x.clone().requires_grad_(True)
    yo, hw = osw.patch_merging(sd, 'd.', xo, (H, W))
    gy = torch.randn(yo.shape, generator=g)
    yo.backward(gy)
    xg = x.clone().cuda().to(dtype).requires_grad_(True)
    gg, bg = gamma.clone().cuda().requires_grad_(True), beta.clone().cuda().requires_grad_(True)
    y = F.linear(ops.patch_merge_ln(xg, (H, W), gg, bg), red.cuda().to(dtype))
    y.backward(gy.cuda().to(dtype))
    tol = 1e-3 if dtype == torch.float32 else 3e-2
    assert hw == ((H + 1) // 2, (W + 1) // 2)
    assert_rel(y, yo, tol, 'y')
    assert_rel(xg.grad, xo.grad, tol, 'dx')
    assert_rel(gg.grad, sd['d.norm.weight'].grad, tol, 'dgamma')
    assert_rel(bg.grad, sd['d.norm.bias'].grad, tol, 'dbeta')


@pytest.fixture
def pm_variant():
    """rsc_set_patch_merge_variant for one test, restored afterwards"""
    from rscotr_b200 import _lib

    def choose(v):
        _lib.call('rsc_set_patch_merge_variant', v)
    yield choose
    _lib.call('rsc_set_patch_merge_variant', 0)


@pytest.mark.parametrize('B,H,W,C', [(2, 151, 151, 96), (1, 151, 149, 192), (1, 100, 100, 384)])
@pytest.mark.parametrize('variant', [1, 2])
def test_patch_merge_ln_persistent_loop(B, H, W, C, variant, pm_variant):
    """more tokens than one pass of the resident warps covers (several steps per warp, a ragged last step) and odd
    H / W (the zero-padded last row / column of mmdet's PatchMerging), bf16, against the fp32 oracle; both kernel
    variants (include/rscotr.h: rsc_set_patch_merge_variant)."""
    ops = _ops()
    pm_variant(variant)
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(B, H * W, C, generator=g).bfloat16().float()
    gamma = 1 + 0.1 * torch.randn(4 * C, generator=g)
    beta = 0.1 * torch.randn(4 * C, generator=g)
    sd = {'d.norm.weight': gamma.clone().requires_grad_(True), 'd.norm.bias': beta.clone().requires_grad_(True),
          'd.reduction.weight': torch.eye(4 * C).requires_grad_(True)}          # (identity: the LayerNorm output itself)
    xo = x.clone().requires_grad_(True)
    yo, hw = osw.patch_merging(sd, 'd.', xo, (H, W))
    gy = torch.randn(yo.shape, generator=g).bfloat16().float()
    yo.backward(gy)
    xg = x.clone().cuda().bfloat16().requires_grad_(True)
    gg, bg = gamma.clone().cuda().requires_grad_(True), beta.clone().cuda().requires_grad_(True)
    y = ops.patch_merge_ln(xg, (H, W), gg, bg)
    y.backward(gy.cuda().bfloat16())
    assert tuple(y.shape) == tuple(yo.shape)
    assert_rel(y, yo, 4e-3, 'y')
    assert (y.float().cpu() - yo).abs().max() <= 2 ** -7 * yo.abs().max()     # every token was written (bf16 rounding only)
    assert_rel(xg.grad, xo.grad, 6e-3, 'dx')
    assert_rel(gg.grad, sd['d.norm.weight'].grad, 2e-3, 'dgamma')
    assert_rel(bg.grad, sd['d.norm.bias'].grad, 2e-3, 'dbeta')


# ---------------------------------------------------------------------------
# a11: ms_deform_attn
# ---------------------------------------------------------------------------
def _msda_inputs(B, Nq, heads, shapes, P, seed, spread=0.25):
    g = torch.Generator().manual_seed(seed)
    Nv = sum(h * w for h, w in shapes)
    L = len(shapes)
    value = torch.randn(B, Nv, heads, 32, generator=g)
    loc = torch.rand(B, Nq, heads, L, P, 2, generator=g) * (1 + 2 * spread) - spread
    w = torch.rand(B, Nq, heads, L, P, generator=g).flatten(-2).softmax(-1).view(B, Nq, heads, L, P)
    starts = [0]
    for h, ww in shapes[:-1]:
        starts.append(starts[-1] + h * ww)
    return value, loc, w, torch.tensor(shapes), torch.tensor(starts)


MSDA_CASES = [
    (2, 37, 8, [(12, 9), (6, 5), (3, 3), (2, 2)], 4),
    (1, 5, 4, [(5, 7)], 2),
    (2, 300, 8, [(25, 25), (13, 13), (7, 7), (4, 4)], 4),
    (1, 64, 8, [(8, 8), (4, 4)], 8),
]


@pytest.mark.parametrize('B,Nq,heads,shapes,P', MSDA_CASES)
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_msda_fwd_bwd(B, Nq, heads, shapes, P, dtype):
    ops = _ops()
    value, loc, w, ss, st = _msda_inputs(B, Nq, heads, shapes, P, seed=Nq)
    if dtype == torch.bfloat16:
        value = value.bfloat16().float()
    vo, lo, wo = (t.clone().requires_grad_(True) for t in (value, loc, w))
    yo = otr.ms_deform_attn_core(vo, shapes, lo, wo)
    g = torch.Generator().manual_seed(3)
    gy = torch.randn(yo.shape, generator=g)
    if dtype == torch.bfloat16:
        gy = gy.bfloat16().float()
    yo.backward(gy)
    vg = value.clone().cuda().to(dtype).requires_grad_(True)
    lg, wg = loc.clone().cuda().requires_grad_(True), w.clone().cuda().requires_grad_(True)
    y = ops.ms_deform_attn(vg, ss.cuda(), st.cuda(), lg, wg, 64)
    y.backward(gy.cuda().to(dtype))
    tol = 1e-3 if dtype == torch.float32 else 1.5e-2
    assert_rel(y, yo, tol, 'out')
    assert_rel(vg.grad, vo.grad, tol, 'dvalue')
    assert_rel(lg.grad, lo.grad, tol, 'dloc')
    assert_rel(wg.grad, wo.grad, tol, 'dweight')


MSDA_T800 = [(100, 100), (50, 50), (25, 25), (13, 13)]


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_msda_encoder_shape_vs_oracle(dtype):
    """mmcv-signature op at the BASELINE encoder shape (B=2, Nq = Nv = 13 294, 8 heads, 4 levels x 4 points):
    values and all three gradients against the oracle's grid_sample restatement."""
    ops = _ops()
    value, loc, w, ss, st = _msda_inputs(2, 13294, 8, MSDA_T800, 4, seed=21, spread=0.1)
    if dtype == torch.bfloat16:
        value = value.bfloat16().float()
    vo, lo, wo = (t.clone().requires_grad_(True) for t in (value, loc, w))
    yo = otr.ms_deform_attn_core(vo, MSDA_T800, lo, wo)
    g = torch.Generator().manual_seed(3)
    gy = torch.randn(yo.shape, generator=g)
    if dtype == torch.bfloat16:
        gy = gy.bfloat16().float()
    yo.backward(gy)
    vg = value.clone().cuda().to(dtype).requires_grad_(True)
    lg, wg = loc.clone().cuda().requires_grad_(True), w.clone().cuda().requires_grad_(True)
    y = ops.ms_deform_attn(vg, ss.cuda(), st.cuda(), lg, wg, 64)
    y.backward(gy.cuda().to(dtype))
    tol = 1e-3 if dtype == torch.float32 else 1.5e-2
    assert_rel(y, yo, tol, 'out')
    assert_rel(vg.grad, vo.grad, tol, 'dvalue')
    assert_rel(lg.grad, lo.grad, tol, 'dloc')
    assert_rel(wg.grad, wo.grad, tol, 'dweight')


@pytest.mark.parametrize('Nq,ref_dim', [(13294, 2), (792, 4)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_msda_fused_t800_shapes_vs_oracle(Nq, ref_dim, dtype):
    """The fused module tail (softmax + sampling locations + op) at the BASELINE encoder shape (13 294 queries,
    2-d reference points) and decoder shape (792 queries = 600 + 192 denoising, 4-d reference boxes), B = 2."""
    ops = _ops()
    g = torch.Generator().manual_seed(17 + ref_dim)
    shapes = MSDA_T800
    B, heads, L, P = 2, 8, 4, 4
    Nv = sum(h * w for h, w in shapes)
    value = torch.randn(B, Nv, heads, 32, generator=g).to(dtype)
    offsets = (torch.randn(B, Nq, heads, L, P, 2, generator=g) * 3).to(dtype)
    logits = torch.randn(B, Nq, heads, L * P, generator=g).to(dtype)
    ref = torch.rand(B, Nq, L, ref_dim, generator=g)
    if ref_dim == 4:
        ref[..., 2:] = ref[..., 2:] * 0.5 + 0.05
    wout = torch.randn(B, Nq, heads * 32, generator=g)
    rv, ro, rl = (t.float().clone().requires_grad_(True) for t in (value, offsets, logits))
    aw = rl.softmax(-1).view(B, Nq, heads, L, P)
    if ref_dim == 2:
        norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32)
        loc = ref[:, :, None, :, None, :] + ro / norm[None, None, None, :, None, :]
    else:
        loc = ref[:, :, None, :, None, :2] + ro / P * ref[:, :, None, :, None, 2:] * 0.5
    want = otr.ms_deform_attn_core(rv, shapes, loc, aw)
    (want * wout).sum().backward()
    gv, go, gl = (t.detach().cuda().requires_grad_(True) for t in (value, offsets, logits))
    ss = torch.tensor(shapes).cuda()
    st = torch.tensor([0] + list(torch.tensor([h * w for h, w in shapes]).cumsum(0)[:-1])).cuda()
    assert ops.msda_fused_supported(gv, go, ref.cuda())
    got = ops.ms_deform_attn_fused(gv, ss, st, go, gl, ref.cuda())
    (got.float() * wout.cuda()).sum().backward()
    f32 = dtype == torch.float32
    assert_rel(got, want, 1e-4 if f32 else 6e-3, 'out')
    assert_rel(gv.grad, rv.grad, 1e-3 if f32 else 1e-2, 'd value')
    assert_rel(go.grad, ro.grad, 1e-3 if f32 else 1.5e-2, 'd offsets')
    assert_rel(gl.grad, rl.grad, 1e-3 if f32 else 1.5e-2, 'd logits')


def test_msda_full_size_linearity():
    """BASELINE size (N = 13294 tokens, 4 levels of an 800^2 image): the op is
    linear in value and in the attention weights; out-of-range samples give 0."""
    ops = _ops()
    shapes = [(100, 100), (50, 50), (25, 25), (13, 13)]
    value, loc, w, ss, st = _msda_inputs(1, 13294, 8, shapes, 4, seed=11)
    value, loc, w, ss, st = (t.cuda() for t in (value, loc, w, ss, st))
    y1 = ops.ms_deform_attn(value, ss, st, loc, w)
    y2 = ops.ms_deform_attn(2 * value, ss, st, loc, 0.5 * w)
    torch.testing.assert_close(y1, y2, rtol=1e-5, atol=1e-5)
    y0 = ops.ms_deform_attn(value, ss, st, loc + 5.0, w)
    assert torch.count_nonzero(y0) == 0
    # constant value field + in-range samples + weights summing to 1 -> constant out
    inr = loc.clamp(0.2, 0.8)
    yc = ops.ms_deform_attn(torch.ones_like(value), ss, st, inr, w)
    torch.testing.assert_close(yc, torch.ones_like(yc), rtol=1e-5, atol=1e-5)


# ---------------------------------------------------------------------------
# a12: GAP,  a18/a19: bilinear,  a16: focal
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_gap(dtype):
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 768, 5, 7, generator=g).to(dtype)
    for cl in (False, True):
        xi = x.flatten(2).transpose(1, 2).contiguous() if cl else x
        xo = xi.float().clone().requires_grad_(True)
        yo = xo.mean(1) if cl else xo.mean((2, 3))
        gy = torch.randn(yo.shape, generator=g)
        yo.backward(gy)
        xg = xi.clone().cuda().requires_grad_(True)
        y = ops.global_avg_pool(xg, cl)
        y.backward(gy.cuda().to(dtype))
        tol = 1e-5 if dtype == torch.float32 else 1e-2
        assert_rel(y, yo, tol, 'gap')
        assert_rel(xg.grad, xo.grad, tol, 'dgap')


@pytest.mark.parametrize('src,dst', [((13, 13), (25, 25)), ((25, 25), (50, 50)), ((100, 100), (13, 13)),
                                     ((10, 12), (80, 96)), ((50, 50), (25, 25)), ((7, 9), (7, 9)), ((1, 1), (4, 4))])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_bilinear(src, dst, dtype):
    ops = _ops()
    g = torch.Generator().manual_seed(src[0] + dst[0])
    x = torch.randn(2, 5, *src, generator=g).to(dtype)
    xo = x.float().clone().requires_grad_(True)
    yo = F.interpolate(xo, size=dst, mode='bilinear', align_corners=False)
    gy = torch.randn(yo.shape, generator=g).to(dtype)
    yo.backward(gy.float())
    xg = x.clone().cuda().requires_grad_(True)
    y = ops.bilinear_resize(xg, dst)
    y.backward(gy.cuda())
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert_rel(y, yo, tol, 'bilinear')
    assert_rel(xg.grad, xo.grad, tol, 'dbilinear')


def _py_sigmoid_focal_loss(pred, target, gamma=2.0, alpha=0.25):
    """mmdet py_sigmoid_focal_loss (the reference's CPU branch), reduction none."""
    p = pred.sigmoid()
    t = F.one_hot(target, pred.shape[1] + 1)[:, :pred.shape[1]].type_as(pred)
    pt = (1 - p) * t + p * (1 - t)
    fw = (alpha * t + (1 - alpha) * (1 - t)) * pt.pow(gamma)
    return F.binary_cross_entropy_with_logits(pred, t, reduction='none') * fw


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_sigmoid_focal_loss(dtype):
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(600, 20, generator=g) * 3).to(dtype)
    t = torch.randint(0, 21, (600,), generator=g)
    xo = x.float().clone().requires_grad_(True)
    lo = _py_sigmoid_focal_loss(xo, t)
    gy = torch.rand(lo.shape, generator=g)
    lo.backward(gy)
    xg = x.clone().cuda().requires_grad_(True)
    l = ops.sigmoid_focal_loss(xg, t.cuda())
    l.backward(gy.cuda())
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    assert_rel(l, lo, tol, 'focal')
    assert_rel(xg.grad, xo.grad, tol, 'dfocal')


# ---------------------------------------------------------------------------
# LayerNorm (a1, a2, a7, a9)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('rows,C', [(5, 96), (1000, 96), (333, 192), (64, 256), (77, 384), (31, 768), (9, 1024)])
@pytest.mark.parametrize('in_dt,out_dt', [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                          (torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float32)])
def test_layernorm(rows, C, in_dt, out_dt):
    ops = _ops()
    g = torch.Generator().manual_seed(rows + C)
    x = (torch.randn(rows, C, generator=g) * 2 + 0.5).to(in_dt)
    gamma = 1 + 0.2 * torch.randn(C, generator=g)
    beta = 0.2 * torch.randn(C, generator=g)
    gy = torch.randn(rows, C, generator=g).to(out_dt)
    xo, go, bo = x.float().clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yo = F.layer_norm(xo, (C,), go, bo)
    yo.backward(gy.float())
    xg = x.clone().cuda().requires_grad_(True)
    gg, bg = gamma.clone().cuda().requires_grad_(True), beta.clone().cuda().requires_grad_(True)
    y = ops.layer_norm(xg, gg, bg, 1e-5, out_dt)
    assert y.dtype == out_dt
    y.backward(gy.cuda())
    tol_y = 1e-5 if out_dt == torch.float32 else 6e-3
    tol_g = 2e-5 if in_dt == torch.float32 else 1e-2
    assert_rel(y, yo, tol_y, 'y')
    assert_rel(xg.grad, xo.grad, tol_g, 'dx')
    assert_rel(gg.grad, go.grad, 1e-4, 'dgamma')
    assert_rel(bg.grad, bo.grad, 1e-4, 'dbeta')


# ---------------------------------------------------------------------------
# a16: GPU Hungarian matching + fused detection losses
# ---------------------------------------------------------------------------
def _det_case(seed, L, B, Nq, ps, C, sizes, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    cls = (torch.randn(L, B, ps + Nq, C, generator=g) * 2 - 2).to(dtype)
    cxcy = torch.rand(L, B, ps + Nq, 2, generator=g) * 0.8 + 0.1
    wh = torch.rand(L, B, ps + Nq, 2, generator=g) * 0.3 + 0.02
    box = torch.cat([cxcy, wh], -1)
    shapes = [(640 + 32 * b, 800 - 16 * b, 3) for b in range(B)]
    gtb, gtl = [], []
    for b, n in enumerate(sizes):
        h, w = shapes[b][:2]
        x1, y1 = torch.rand(n, generator=g) * (w - 120), torch.rand(n, generator=g) * (h - 120)
        bw, bh = torch.rand(n, generator=g) * 100 + 16, torch.rand(n, generator=g) * 100 + 16
        gtb.append(torch.stack([x1, y1, x1 + bw, y1 + bh], -1))
        gtl.append(torch.randint(0, C, (n,), generator=g))
    return cls, box, gtb, gtl, [dict(img_shape=s) for s in shapes]


@pytest.mark.parametrize('sizes', [(5, 3), (8,), (0, 4), (17, 1, 9)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_det_match_cost_and_assignment(sizes, dtype):
    """cost matrix == mmdet's FocalLossCost + BBoxL1Cost + IoUCost (host-side HungarianAssigner.cost) and the
    assignment == scipy.optimize.linear_sum_assignment on that matrix, for every (layer, image) problem."""
    from scipy.optimize import linear_sum_assignment
    from rscotr_b200.models.det_head import HungarianAssigner
    ops = _ops()
    L, B, Nq, ps, C = 3, len(sizes), 60, 12, 20
    cls, box, gtb, gtl, metas = _det_case(11 + len(sizes), L, B, Nq, ps, C, sizes, dtype)
    starts = [0]
    for n in sizes:
        starts.append(starts[-1] + n)
    dev = 'cuda'
    img_wh = torch.tensor([[m['img_shape'][1], m['img_shape'][0]] for m in metas], dtype=torch.float32, device=dev)
    assign, cost, gt_norm = ops.det_match(cls.to(dev), box.to(dev), ps, Nq, torch.cat(gtl).to(dev), torch.cat(gtb).to(dev),
                                          torch.tensor(starts, dtype=torch.int32, device=dev), img_wh, max(sizes))
    assign, cost, gt_norm = assign.cpu().view(L, B, Nq), cost.cpu().view(L, B, max(sizes), Nq), gt_norm.cpu()
    asg = HungarianAssigner()
    for b, n in enumerate(sizes):
        h, w = metas[b]['img_shape'][:2]
        f = torch.tensor([w, h, w, h], dtype=torch.float32)
        from rscotr_b200.models.det_head import bbox_xyxy_to_cxcywh
        if n:
            assert torch.allclose(gt_norm[starts[b]:starts[b + 1]], bbox_xyxy_to_cxcywh(gtb[b] / f), atol=1e-6)
        for l in range(L):
            want_cost = asg.cost(box[l, b, ps:], cls[l, b, ps:].float(), gtb[b], gtl[b], metas[b]['img_shape'])   # (Nq, n)
            got_cost = cost[l, b, :n].t()
            if n:
                assert torch.allclose(got_cost, want_cost, rtol=1e-4, atol=2e-4), (got_cost - want_cost).abs().max()
            r, c = linear_sum_assignment(got_cost.double().numpy()) if n else ([], [])
            want = torch.full((Nq,), -1, dtype=torch.int32)
            for ri, ci in zip(r, c):
                want[ri] = starts[b] + ci
            assert torch.equal(assign[l, b], want), (l, b)
            r2, c2 = linear_sum_assignment(want_cost.double().numpy()) if n else ([], [])
            assert float(got_cost[r, c].sum()) <= float(got_cost[r2, c2].sum()) + 1e-3 if n else True


def test_det_match_large_problem():
    """DIOR-scale problem: 600 queries x 120 boxes; the device assignment is optimal (same total cost as scipy)."""
    from scipy.optimize import linear_sum_assignment
    ops = _ops()
    cls, box, gtb, gtl, metas = _det_case(3, 2, 1, 600, 0, 20, (120,))
    dev = 'cuda'
    img_wh = torch.tensor([[metas[0]['img_shape'][1], metas[0]['img_shape'][0]]], dtype=torch.float32, device=dev)
    assign, cost, _ = ops.det_match(cls.to(dev), box.to(dev), 0, 600, gtl[0].to(dev), gtb[0].to(dev),
                                    torch.tensor([0, 120], dtype=torch.int32, device=dev), img_wh, 120)
    for l in range(2):
        cm = cost[l].cpu().double().numpy()          # (120, 600)
        r, c = linear_sum_assignment(cm)
        a = assign[l].cpu()
        cols = torch.nonzero(a >= 0).flatten()
        assert len(cols) == 120 and sorted(a[cols].tolist()) == list(range(120))
        got_total = sum(cm[int(a[j]), int(j)] for j in cols)
        assert abs(got_total - cm[r, c].sum()) < 1e-6 * max(1.0, abs(cm[r, c].sum()))
        want = torch.full((600,), -1, dtype=torch.int32)
        want[torch.from_numpy(c)] = torch.from_numpy(r).int()
        assert torch.equal(a, want)


@pytest.mark.parametrize('sizes', [(5, 3), (0, 4), (6,)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_det_loss_fused_matches_aten_path(sizes, dtype):
    """DINOHead.loss through rsc_det_match + rsc_det_loss_{fwd,bwd} == the ATen + scipy path (values and
    gradients w.r.t. class logits and boxes), incl. the denoising part and an image without boxes."""
    from rscotr_b200.models.det_head import DINOHead
    from tests.test_host_parity import OCFG  # noqa: F401  (registers nothing, keeps import order)
    import rscotr_b200.models  # noqa: F401
    from rscotr_b200.config import Config
    import os
    cfg = Config.fromfile(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'configs', 'multi',
                                       'cotrain_swin-t_800.py'))
    hc = dict(cfg.model.bbox_head)
    hc.pop('type')
    hc.update(train_cfg=cfg.model.train_cfg.get('det'), test_cfg=cfg.model.test_cfg.get('det'), num_query=40)
    torch.manual_seed(0)
    head = DINOHead(**hc).cuda()
    L, B, C = 6, len(sizes), head.num_classes
    num_groups = max(1, 10 // max(max(sizes), 1))
    ps = max(sizes) * 2 * num_groups
    Nq = 40
    cls, box, gtb, gtl, metas = _det_case(5, L, B, Nq, ps, C, sizes, dtype)
    ecls, ebox, _, _, _ = _det_case(6, 1, B, Nq, 0, C, sizes, dtype)
    dn_meta = dict(pad_size=ps, num_dn_group=num_groups)
    outs = {}
    for fused in (False, True):
        head.fused_loss = fused
        ins = [t.clone().cuda().requires_grad_(True) for t in (cls, box, ecls[0], ebox[0])]
        losses = head.loss(ins[0], ins[1], ins[2], ins[3], [t.cuda() for t in gtb], [t.cuda() for t in gtl], metas, dn_meta)
        keys = list(losses.keys())
        total = sum((i + 1) * 0.37 * losses[k] for i, k in enumerate(keys))     # distinct upstream gradient per term
        total.backward()
        outs[fused] = ({k: float(losses[k]) for k in keys}, [t.grad.float().cpu() for t in ins])
    assert list(outs[True][0].keys()) == list(outs[False][0].keys())
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    for k, v in outs[False][0].items():
        assert abs(outs[True][0][k] - v) <= tol * max(1.0, abs(v)), (k, outs[True][0][k], v)
    for i, (g, w) in enumerate(zip(outs[True][1], outs[False][1])):
        e = float((g - w).norm() / w.norm().clamp_min(1e-12))
        assert e <= (1e-4 if dtype == torch.float32 else 2e-2), (i, e)


# ---------------------------------------------------------------------------
# a19: fused bilinear upsample + cross-entropy
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('shape', [(2, 7, 13, 10, 104, 80), (1, 100, 25, 25, 200, 200), (2, 5, 10, 9, 37, 50),
                                   (1, 33, 12, 12, 12, 12), (1, 3, 1, 1, 8, 8)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_upsample_ce_matches_interpolate_cross_entropy(shape, dtype):
    ops = _ops()
    B, C, h, w, H, W = shape
    g = torch.Generator().manual_seed(H + C)
    logits = (torch.randn(B, C, h, w, generator=g) * 3).to(dtype)
    label = torch.randint(0, C, (B, H, W), generator=g)
    label[torch.rand(B, H, W, generator=g) < 0.1] = 255
    ref_in = logits.float().clone().requires_grad_(True)
    up = F.interpolate(ref_in, size=(H, W), mode='bilinear', align_corners=False)
    ce = F.cross_entropy(up, label, reduction='sum', ignore_index=255)
    (ce * 0.37).backward()
    valid = label != 255
    correct = ((up.argmax(1) == label) & valid).sum()
    x = logits.cuda().requires_grad_(True)
    assert ops.upsample_ce_supported(x, label.cuda())
    stats = ops.upsample_ce(x, label.cuda(), 255)
    (stats[0] * 0.37).backward()
    assert abs(float(stats[0]) - float(ce)) <= 2e-5 * abs(float(ce)) + 1e-3
    assert int(stats[2]) == int(valid.sum())
    assert abs(int(stats[1]) - int(correct)) <= (0 if dtype == torch.float32 else 2) + int(1e-4 * valid.sum())
    assert_rel(x.grad, ref_in.grad, 1e-4 if dtype == torch.float32 else 6e-3, 'dlogits')


# ---------------------------------------------------------------------------
# a2: fused residual-stream passes (add + DropPath + bias + LayerNorm, bias + GELU)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('C', [96, 192, 384, 768, 256])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('with_scale', [False, True])
def test_add_ln_matches_eager(C, dtype, with_scale):
    ops = _ops()
    g = torch.Generator().manual_seed(C)
    B, L = 3, 37
    ident, x = torch.randn(B, L, C, generator=g).to(dtype), torch.randn(B, L, C, generator=g).to(dtype)
    bias, gamma, beta = torch.randn(C, generator=g) * .1, torch.rand(C, generator=g) + .5, torch.randn(C, generator=g) * .1
    scale = torch.tensor([0.0, 1.25, 1.25]) if with_scale else None
    w_r, w_n = torch.randn(B, L, C, generator=g), torch.randn(B, L, C, generator=g)
    # eager reference in fp32 on the same (rounded) inputs
    ri, rx = ident.float().clone().requires_grad_(True), x.float().clone().requires_grad_(True)
    rb, rg, rbe = bias.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    r = ri + (rx + rb) * (scale.view(B, 1, 1) if with_scale else 1.0)
    if dtype == torch.bfloat16:
        r = r + (r.detach().bfloat16().float() - r.detach())      # the stored residual is bf16
    n = F.layer_norm(r, (C,), rg, rbe, 1e-5)
    ((r * w_r).sum() + (n * w_n).sum()).backward()
    ins = [t.detach().cuda().requires_grad_(True) for t in (ident, x, bias, gamma, beta)]
    gr, gn = ops.add_ln(ins[0], ins[1], ins[2], None if scale is None else scale.cuda(), ins[3], ins[4], 1e-5)
    ((gr.float() * w_r.cuda()).sum() + (gn.float() * w_n.cuda()).sum()).backward()
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert_rel(gr, r, tol, 'r')
    assert_rel(gn, n, tol, 'n')
    for got, want, what in zip(ins, (ri, rx, rb, rg, rbe), ('d_identity', 'dx', 'dbias', 'dgamma', 'dbeta')):
        assert_rel(got.grad, want.grad, 2e-4 if dtype == torch.float32 else 2e-2, what)


@pytest.mark.parametrize('C,L', [(96, 200003), (192, 35003), (256, 13294), (384, 20001)])
def test_add_ln_many_rows(C, L):
    """more rows than one pass of the resident warps covers (several steps of two row groups per warp, ragged tail),
    per-sample DropPath scale incl. a dropped sample, bf16 against the fp32 eager chain."""
    ops = _ops()
    g = torch.Generator().manual_seed(C)
    B = 2
    ident, x = torch.randn(B, L, C, generator=g).bfloat16(), torch.randn(B, L, C, generator=g).bfloat16()
    bias, gamma, beta = torch.randn(C, generator=g) * .1, torch.rand(C, generator=g) + .5, torch.randn(C, generator=g) * .1
    scale = torch.tensor([1.25, 0.0])
    w_r, w_n = torch.randn(B, L, C, generator=g).bfloat16().float(), torch.randn(B, L, C, generator=g).bfloat16().float()
    ri, rx = ident.float().clone().requires_grad_(True), x.float().clone().requires_grad_(True)
    rb, rg, rbe = bias.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    r = ri + (rx + rb) * scale.view(B, 1, 1)
    r = r + (r.detach().bfloat16().float() - r.detach())      # the stored residual is bf16
    n = F.layer_norm(r, (C,), rg, rbe, 1e-5)
    ((r * w_r).sum() + (n * w_n).sum()).backward()
    ins = [t.detach().cuda().requires_grad_(True) for t in (ident, x, bias, gamma, beta)]
    gr, gn = ops.add_ln(ins[0], ins[1], ins[2], scale.cuda(), ins[3], ins[4], 1e-5)
    ((gr.float() * w_r.cuda()).sum() + (gn.float() * w_n.cuda()).sum()).backward()
    assert_rel(gr, r, 1e-3, 'r')
    assert (gr.float().cpu() - r.detach()).abs().max() <= 2 ** -7 * float(r.detach().abs().max())
    assert_rel(gn, n, 4e-3, 'n')
    assert (gn.float().cpu() - n.detach()).abs().max() <= 2 ** -7 * float(n.detach().abs().max())    # every row was written
    for got, want, what in zip(ins, (ri, rx, rb, rg, rbe), ('d_identity', 'dx', 'dbias', 'dgamma', 'dbeta')):
        assert_rel(got.grad, want.grad, 6e-3, what)


@pytest.mark.parametrize('rows,C', [(50, 384), (1000, 768), (33, 3072), (7, 8)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_bias_gelu_matches_eager(rows, C, dtype):
    ops = _ops()
    g = torch.Generator().manual_seed(rows)
    h, bias, w = (torch.randn(rows, C, generator=g) * 2).to(dtype), torch.randn(C, generator=g), torch.randn(rows, C, generator=g)
    rh, rb = h.float().clone().requires_grad_(True), bias.clone().requires_grad_(True)
    y = F.gelu(rh + rb)
    (y * w).sum().backward()
    gh, gb = h.detach().cuda().requires_grad_(True), bias.detach().cuda().requires_grad_(True)
    gy = ops.bias_gelu(gh, gb)
    (gy.float() * w.cuda()).sum().backward()
    assert_rel(gy, y, 1e-5 if dtype == torch.float32 else 6e-3, 'y')
    assert_rel(gh.grad, rh.grad, 1e-4 if dtype == torch.float32 else 1e-2, 'dh')
    assert_rel(gb.grad, rb.grad, 1e-4 if dtype == torch.float32 else 2e-2, 'dbias')
    # ReLU variant (encoder / decoder FFNs)
    rh2, rb2 = h.float().clone().requires_grad_(True), bias.clone().requires_grad_(True)
    (F.relu(rh2 + rb2) * w).sum().backward()
    gh2, gb2 = h.detach().cuda().requires_grad_(True), bias.detach().cuda().requires_grad_(True)
    gy2 = ops.bias_relu(gh2, gb2)
    (gy2.float() * w.cuda()).sum().backward()
    assert_rel(gy2, F.relu(rh2 + rb2), 1e-6 if dtype == torch.float32 else 6e-3, 'relu y')
    assert_rel(gh2.grad, rh2.grad, 1e-6 if dtype == torch.float32 else 1e-2, 'relu dh')
    assert_rel(gb2.grad, rb2.grad, 1e-5 if dtype == torch.float32 else 2e-2, 'relu dbias')


@pytest.mark.parametrize('rows,C', [(50, 384), (1000, 768), (7, 8)])
def test_bias_gelu_logistic_fit_variant(rows, C):
    """act=2, the bf16 default since round 2 (A/B on B200: profiles/r02_ab_switches.log): same tolerances as the erf
    form; fp32 inputs are refused."""
    ops = _ops()
    g = torch.Generator().manual_seed(rows)
    h, bias, w = (torch.randn(rows, C, generator=g) * 3).bfloat16(), torch.randn(C, generator=g), torch.randn(rows, C, generator=g)
    h[0, :8] = torch.tensor([-60., -9., -7., -5., 5., 7., 9., 60.]).bfloat16()          # the clamp and the tails
    rh, rb = h.float().clone().requires_grad_(True), bias.clone().requires_grad_(True)
    y = F.gelu(rh + rb)
    (y * w).sum().backward()
    gh, gb = h.detach().cuda().requires_grad_(True), bias.detach().cuda().requires_grad_(True)
    gy = ops.bias_act(gh, gb, ops.ACT_GELU_SIG)
    (gy.float() * w.cuda()).sum().backward()
    assert torch.isfinite(gy).all()
    assert_rel(gy, y, 6e-3, 'y')
    assert_rel(gh.grad, rh.grad, 1e-2, 'dh')
    assert_rel(gb.grad, rb.grad, 2e-2, 'dbias')
    with pytest.raises(RuntimeError):
        ops.bias_act(h.float().cuda(), bias.cuda(), ops.ACT_GELU_SIG)


# ---------------------------------------------------------------------------
# a10/a11: fused softmax + sampling locations + ms_deform_attn
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('ref_dim', [2, 4])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_msda_fused_matches_oracle_chain(ref_dim, dtype):
    """rsc_msda_fused_{fwd,bwd} == softmax + location arithmetic (mmcv MultiScaleDeformableAttention.forward)
    + the oracle sampling op, values and gradients w.r.t. value / offsets / logits, incl. samples that leave
    the feature maps."""
    ops = _ops()
    g = torch.Generator().manual_seed(7 + ref_dim)
    shapes = [(12, 9), (6, 5), (3, 3), (2, 2)]
    B, heads, L, P = 2, 8, 4, 4
    Nv = sum(h * w for h, w in shapes)
    Nq = 37
    value = torch.randn(B, Nv, heads, 32, generator=g).to(dtype)
    offsets = (torch.randn(B, Nq, heads, L, P, 2, generator=g) * 3).to(dtype)
    logits = torch.randn(B, Nq, heads, L * P, generator=g).to(dtype)
    ref = torch.rand(B, Nq, L, ref_dim, generator=g)
    if ref_dim == 4:
        ref[..., 2:] = ref[..., 2:] * 0.5 + 0.05
    wout = torch.randn(B, Nq, heads * 32, generator=g)
    # oracle chain in fp32 on the same (rounded) inputs
    rv, ro, rl = (t.float().clone().requires_grad_(True) for t in (value, offsets, logits))
    aw = rl.softmax(-1).view(B, Nq, heads, L, P)
    if ref_dim == 2:
        norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32)
        loc = ref[:, :, None, :, None, :] + ro / norm[None, None, None, :, None, :]
    else:
        loc = ref[:, :, None, :, None, :2] + ro / P * ref[:, :, None, :, None, 2:] * 0.5
    want = otr.ms_deform_attn_core(rv, shapes, loc, aw)
    (want * wout).sum().backward()
    gv, go, gl = (t.detach().cuda().requires_grad_(True) for t in (value, offsets, logits))
    ss = torch.tensor(shapes).cuda()
    st = torch.tensor([0] + list(torch.tensor([h * w for h, w in shapes]).cumsum(0)[:-1])).cuda()
    assert ops.msda_fused_supported(gv, go, ref.cuda())
    got = ops.ms_deform_attn_fused(gv, ss, st, go, gl, ref.cuda())
    (got.float() * wout.cuda()).sum().backward()
    f32 = dtype == torch.float32
    assert_rel(got, want, 1e-5 if f32 else 6e-3, 'out')
    assert_rel(gv.grad, rv.grad, 1e-4 if f32 else 1e-2, 'd value')
    assert_rel(go.grad, ro.grad, 1e-4 if f32 else 1.5e-2, 'd offsets')
    assert_rel(gl.grad, rl.grad, 1e-4 if f32 else 1.5e-2, 'd logits')


@pytest.mark.parametrize('ref_dim', [2, 4])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_msda_fused_packed_rows_match_oracle_chain(ref_dim, dtype):
    """the same kernels with ROW STRIDES: offsets and logits as column ranges of one (B, Nq, 384) matrix (the output of
    the single GEMM over the stacked sampling_offsets / attention_weights weights), both gradients into one matrix.
    Compared with the oracle chain and, bit for bit, with the dense-row call."""
    ops = _ops()
    g = torch.Generator().manual_seed(17 + ref_dim)
    shapes = [(11, 13), (6, 7), (3, 4), (2, 2)]
    B, heads, L, P = 2, 8, 4, 4
    Nv = sum(h * w for h, w in shapes)
    Nq = 53
    value = torch.randn(B, Nv, heads, 32, generator=g).to(dtype)
    both = torch.randn(B, Nq, heads * L * P * 3, generator=g)
    both[..., :heads * L * P * 2] *= 3
    both = both.to(dtype)
    n_off = heads * L * P * 2
    ref = torch.rand(B, Nq, L, ref_dim, generator=g)
    if ref_dim == 4:
        ref[..., 2:] = ref[..., 2:] * 0.5 + 0.05
    wout = torch.randn(B, Nq, heads * 32, generator=g)
    rv, rb = (t.float().clone().requires_grad_(True) for t in (value, both))
    ro = rb[..., :n_off].reshape(B, Nq, heads, L, P, 2)
    aw = rb[..., n_off:].reshape(B, Nq, heads, L * P).softmax(-1).view(B, Nq, heads, L, P)
    if ref_dim == 2:
        norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32)
        loc = ref[:, :, None, :, None, :] + ro / norm[None, None, None, :, None, :]
    else:
        loc = ref[:, :, None, :, None, :2] + ro / P * ref[:, :, None, :, None, 2:] * 0.5
    want = otr.ms_deform_attn_core(rv, shapes, loc, aw)
    (want * wout).sum().backward()
    ss = torch.tensor(shapes).cuda()
    st = torch.tensor([0] + list(torch.tensor([h * w for h, w in shapes]).cumsum(0)[:-1])).cuda()
    gv, gb = (t.detach().cuda().requires_grad_(True) for t in (value, both))
    assert ops.msda_packed_supported(gv, gb, ref.cuda(), L, P)
    got = ops.ms_deform_attn_fused_packed(gv, ss, st, gb, ref.cuda(), L, P)
    (got.float() * wout.cuda()).sum().backward()
    f32 = dtype == torch.float32
    assert_rel(got, want, 1e-5 if f32 else 6e-3, 'out')
    assert_rel(gv.grad, rv.grad, 1e-4 if f32 else 1e-2, 'd value')
    assert_rel(gb.grad[..., :n_off], rb.grad[..., :n_off], 1e-4 if f32 else 1.5e-2, 'd offsets')
    assert_rel(gb.grad[..., n_off:], rb.grad[..., n_off:], 1e-4 if f32 else 1.5e-2, 'd logits')
    # dense rows: same arithmetic, so the same bits (the value gradient is an atomic sum: tolerance)
    dv = value.cuda().requires_grad_(True)
    do = both[..., :n_off].reshape(B, Nq, heads, L, P, 2).contiguous().cuda().requires_grad_(True)
    dl = both[..., n_off:].reshape(B, Nq, heads, L * P).contiguous().cuda().requires_grad_(True)
    dense = ops.ms_deform_attn_fused(dv, ss, st, do, dl, ref.cuda())
    (dense.float() * wout.cuda()).sum().backward()
    assert torch.equal(dense, got)
    assert torch.equal(do.grad.reshape(B, Nq, -1), gb.grad[..., :n_off])
    assert torch.equal(dl.grad.reshape(B, Nq, -1), gb.grad[..., n_off:])
    assert_rel(dv.grad, gv.grad, 1e-5 if f32 else 1e-2, 'd value, dense vs packed')


def test_msda_fused_rejects_short_row_strides():
    from rscotr_b200 import _lib
    z = torch.zeros(8, device='cuda')
    with pytest.raises(RuntimeError, match='row strides'):
        _lib.call('rsc_msda_fused_fwd', z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(),
                  z.data_ptr(), 1, 4, 1, 8, 4, 4, 2, 0, 0, 100, 0, 0)


# ---------------------------------------------------------------------------
# a1: patch-embedding gather
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('dtype,odt', [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                       (torch.bfloat16, torch.bfloat16)])
def test_patchify4_matches_unfold_and_conv(dtype, odt):
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 24, 36, generator=g).to(dtype)
    want = F.unfold(x.float(), kernel_size=4, stride=4).transpose(1, 2).reshape(-1, 48)      # (c, kh, kw) order
    got = ops.patchify4(x.cuda(), odt)
    assert got.dtype == odt and tuple(got.shape) == (2 * 6 * 9, 48)
    assert torch.equal(got.float().cpu(), want.to(odt).float())
    # the projection as a GEMM over the gathered rows == Conv2d(k4, s4)
    conv = torch.nn.Conv2d(3, 96, 4, 4)
    ref = conv(x.float()).flatten(2).transpose(1, 2).reshape(-1, 96)
    y = F.linear(want, conv.weight.view(96, 48), conv.bias)
    assert_rel(y, ref, 1e-5, 'conv as gemm')


@pytest.mark.parametrize('to_rgb', [True, False])
def test_normalize_u8_equals_host_normalise_then_pad(to_rgb):
    """rsc_normalize_u8 (device half of a deferred Normalize) == the libraries' host order: normalise each image at
    its own size, then pad right / bottom with zeros; bit-for-bit the same fp32 arithmetic ((x - mean) * (1 / std))"""
    from rscotr_b200.models.mtl import normalize_on_device
    g = torch.Generator().manual_seed(0)
    B, H, W = 3, 40, 52
    img = torch.randint(0, 256, (B, 3, H, W), generator=g, dtype=torch.uint8)
    shapes = [(40, 52), (33, 52), (40, 17)]
    cfg = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=to_rgb)
    metas = [dict(img_norm_cfg=cfg, img_shape=(h, w, 3)) for h, w in shapes]
    want = torch.zeros(B, 3, H, W)
    mean = torch.tensor(cfg['mean']).view(3, 1, 1)
    inv = (1.0 / torch.tensor(cfg['std'], dtype=torch.float64)).float().view(3, 1, 1)
    for b, (h, w) in enumerate(shapes):
        x = img[b].flip(0) if to_rgb else img[b]
        want[b, :, :h, :w] = ((x.float() - mean) * inv)[:, :h, :w]
    got = normalize_on_device(img.cuda(), metas)
    assert got.dtype == torch.float32 and torch.allclose(got.cpu(), want, rtol=1e-6, atol=1e-6)
    assert float(got[1, :, 33:].abs().max()) == 0. and float(got[2, :, :, 17:].abs().max()) == 0.


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('shape', [(2, 256, 25, 25, 50, 50), (1, 512, 16, 16, 128, 128), (2, 64, 13, 7, 20, 33), (1, 8, 6, 6, 6, 6),
                                   (2, 256, 50, 50, 100, 100)])
def test_bilinear_channels_last_matches_interpolate(shape, dtype):
    """rsc_bilinear_cl_{fwd,bwd}: the resize on channels-last maps (no transpose copies) == F.interpolate(bilinear,
    align_corners=False) forward and backward; the result keeps channels-last strides"""
    ops = _ops()
    B, C, Hi, Wi, Ho, Wo = shape
    g = torch.Generator().manual_seed(Hi * Wo + C)
    x = torch.randn(B, C, Hi, Wi, generator=g).to(dtype)
    dy = torch.randn(B, C, Ho, Wo, generator=g).to(dtype)
    xr = x.float().clone().requires_grad_(True)
    want = F.interpolate(xr, size=(Ho, Wo), mode='bilinear', align_corners=False)
    want.backward(dy.float())
    xc = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    got = ops.bilinear_resize(xc, (Ho, Wo))
    assert got.shape == want.shape
    if (Hi, Wi) != (1, 1) and C > 1:
        assert got.permute(0, 2, 3, 1).is_contiguous()
    got.backward(dy.cuda().contiguous(memory_format=torch.channels_last))
    tol = 1e-5 if dtype == torch.float32 else 6e-3
    assert_rel(got, want, tol, 'y')
    assert_rel(xc.grad, xr.grad, tol, 'dx')


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_box_refine_matches_eager_chain(dtype):
    """rsc_box_refine_{fwd,bwd} == (tmp.float() + inverse_sigmoid(ref, 1e-3)).sigmoid() of DinoTransformerDecoder.forward /
    DINOHead.forward, values and both gradients, including references on and beyond the clamp boundaries"""
    from rscotr_b200.models.bricks import inverse_sigmoid
    ops = _ops()
    g = torch.Generator().manual_seed(11)
    ref = torch.rand(2, 1100, 4, generator=g)
    ref[0, :6, 0] = torch.tensor([0.0, 1.0, 5e-4, 1 - 5e-4, 1e-3, 1 - 1e-3])
    ref[0, 6:8, 1] = torch.tensor([-0.2, 1.3])
    tmp = (torch.randn(2, 1100, 4, generator=g) * 2).to(dtype)
    dout = torch.randn(2, 1100, 4, generator=g)
    tr, rr = tmp.float().clone().requires_grad_(True), ref.clone().requires_grad_(True)
    want = (tr + inverse_sigmoid(rr, eps=1e-3)).sigmoid()
    want.backward(dout)
    tc, rc = tmp.cuda().requires_grad_(True), ref.cuda().requires_grad_(True)
    got = ops.box_refine(tc, rc, 1e-3)
    assert got.dtype == torch.float32
    got.backward(dout.cuda())
    assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-6)
    # (1 - s) cancels for saturated boxes: one ulp of s is 1e-3 of (1 - s) at s = 0.9999, so compare in norm and, element-wise,
    # with the tolerance that one ulp of the sigmoid implies
    assert_rel(rc.grad, rr.grad, 1e-4, 'dref')
    assert torch.allclose(rc.grad.cpu(), rr.grad, rtol=2e-2, atol=1e-4)
    assert_rel(tc.grad, tr.grad, 1e-5 if dtype == torch.float32 else 4e-3, 'dtmp')
