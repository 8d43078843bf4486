"""GPU parity of the tcgen05 Linear kernels (csrc/gemm_tc.cu) through the C ABI: rsc_linear_fwd / rsc_linear_dx /
rsc_linear_dw against fp32 torch on the same bf16-rounded operands.  Tolerances: the outputs are bf16 (relative
rounding 2^-9) and the GELU is the one-tanh fit of the erf form (|error| < 3e-4 absolute)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (M, N, K): Swin stage widths (96..768, x3 for qkv, x4 for the MLP), encoder FFN 256 <-> 2048, ragged token counts,
# N / K that are not multiples of the tile (288 = 3 x 96, 96 < 128, 48 = patch embedding)
SHAPES = [(300, 96, 96), (1000, 288, 96), (257, 384, 96), (640, 96, 384), (512, 192, 768), (130, 2048, 256),
          (1000, 256, 2048), (333, 768, 3072), (129, 3072, 768), (77, 96, 48), (4000, 576, 192), (128, 128, 64),
          (1, 256, 256), (5000, 1152, 384)]


def _call(name, *args):
    from rscotr_b200 import _lib
    _lib.call(name, *args)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _mk(M, N, K, seed):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(M, K, generator=g)).bfloat16()
    w = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16()
    b = torch.randn(N, generator=g)
    return x, w, b


@pytest.mark.parametrize('M,N,K', SHAPES)
@pytest.mark.parametrize('act', [0, 1, 2])
@pytest.mark.parametrize('with_bias', [True, False])
def test_linear_fwd(M, N, K, act, with_bias):
    x, w, b = _mk(M, N, K, M + N + K + act)
    h_ref = x.float() @ w.float().t() + (b if with_bias else 0)
    hb = h_ref.bfloat16().float()
    y_ref = F.gelu(hb) if act == 1 else (h_ref.relu() if act == 2 else h_ref)
    xg, wg, bg = x.cuda(), w.cuda(), b.cuda()
    y = torch.full((M, N), float('nan'), dtype=torch.bfloat16, device='cuda')
    h = torch.full((M, N), float('nan'), dtype=torch.bfloat16, device='cuda') if act == 1 else None
    _call('rsc_linear_fwd', xg.data_ptr(), wg.data_ptr(), bg.data_ptr() if with_bias else None, y.data_ptr(),
          h.data_ptr() if h is not None else None, M, N, K, K, K, N, act, _stream())
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()
    assert rel(y, y_ref) < (6e-3 if act == 1 else 4e-3)
    if act == 1:
        assert rel(h, h_ref) < 4e-3
        # |GELU error| bound of the one-tanh fit, evaluated on the kernel's own (rounded) pre-activation
        assert (y.float().cpu() - F.gelu(h.float().cpu())).abs().max() < 3e-4 + 2 ** -8 * y.float().abs().max().item()


@pytest.mark.parametrize('M,N,K', SHAPES)
@pytest.mark.parametrize('act', [0, 1, 2])
def test_linear_dx(M, N, K, act):
    """dX (M,K) = (dY (M,N) W (N,K)) * act'(aux)"""
    g = torch.Generator().manual_seed(M * 7 + N + K + act)
    dy = torch.randn(M, N, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) * N ** -0.5).bfloat16()
    aux = (torch.randn(M, K, generator=g) * 1.5).bfloat16()
    ref = dy.float() @ w.float()
    if act == 1:
        a = aux.float().clone().requires_grad_(True)
        F.gelu(a).sum().backward()
        ref = ref * a.grad
    elif act == 2:
        ref = ref * (aux.float() > 0)
    dx = torch.full((M, K), float('nan'), dtype=torch.bfloat16, device='cuda')
    dyg, wg, auxg = dy.cuda(), w.cuda(), aux.cuda()          # (named: a temporary would be freed before the launch)
    _call('rsc_linear_dx', dyg.data_ptr(), wg.data_ptr(), auxg.data_ptr() if act else None, dx.data_ptr(),
          M, N, K, N, K, K, act, _stream())
    torch.cuda.synchronize()
    assert torch.isfinite(dx).all()
    assert rel(dx, ref) < (8e-3 if act == 1 else 4e-3)


@pytest.mark.parametrize('M,N,K', SHAPES + [(40000, 384, 96), (26588, 2048, 256)])
@pytest.mark.parametrize('with_db', [True, False])
def test_linear_dw(M, N, K, with_db):
    """dW (N,K) += dY^T X, db (N) += colsum(dY): accumulated on top of existing values"""
    g = torch.Generator().manual_seed(M + 3 * N + K)
    dy = torch.randn(M, N, generator=g).bfloat16()
    x = torch.randn(M, K, generator=g).bfloat16()
    dw0 = torch.randn(N, K, generator=g)
    db0 = torch.randn(N, generator=g)
    dw_ref = dw0.double() + dy.double().t() @ x.double()
    db_ref = db0.double() + dy.double().sum(0)
    dw, db = dw0.cuda(), db0.cuda()
    dyg, xg = dy.cuda(), x.cuda()
    _call('rsc_linear_dw', dyg.data_ptr(), xg.data_ptr(), dw.data_ptr(), db.data_ptr() if with_db else None,
          M, N, K, N, K, K, _stream())
    torch.cuda.synchronize()
    assert rel(dw, dw_ref) < 1e-4
    if with_db:
        assert rel(db, db_ref) < 1e-4
    else:
        assert torch.equal(db.cpu(), db0)


def test_linear_leading_dimensions():
    """operands that are column slices of wider matrices (the q / k / v thirds of a packed projection)"""
    M, N, K = 500, 192, 96
    g = torch.Generator().manual_seed(5)
    xw = torch.randn(M, 3 * K, generator=g).bfloat16().cuda()
    w = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16().cuda()
    yw = torch.zeros(M, 2 * N, dtype=torch.bfloat16, device='cuda')
    x = xw[:, K:2 * K]
    y = yw[:, N:]
    _call('rsc_linear_fwd', x.data_ptr(), w.data_ptr(), None, y.data_ptr(), None, M, N, K, 3 * K, K, 2 * N, 0, _stream())
    torch.cuda.synchronize()
    assert rel(y, x.float() @ w.float().t()) < 4e-3
    assert torch.count_nonzero(yw[:, :N]) == 0


@pytest.mark.parametrize('B,L,N,K', [(2, 2500, 96, 96), (4, 1100, 96, 384), (2, 2100, 192, 768), (1, 13294, 256, 2048),
                                     (3, 1400, 256, 256), (2, 2049, 128, 64), (1, 4096, 128, 96)])
@pytest.mark.parametrize('with_scale', [True, False])
def test_linear_add_ln_fused(B, L, N, K, with_scale):
    """ops.linear_add_ln (Linear + residual add with the per-sample DropPath scale + LayerNorm in the GEMM epilogue) against
    the fp32 composition, values and every gradient (x, weight, bias, identity, gamma, beta)."""
    from rscotr_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + N + K)
    x = torch.randn(B, L, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, generator=g) * 0.3
    ident = (torch.randn(B, L, N, generator=g) * 2 + 0.5).bfloat16()
    scale = (torch.rand(B, generator=g) > 0.3).float() / 0.7 if with_scale else None
    gamma, beta = 1 + 0.2 * torch.randn(N, generator=g), 0.2 * torch.randn(N, generator=g)
    wr, wn = torch.randn(B, L, N, generator=g), torch.randn(B, L, N, generator=g)
    # fp32 composition on the same rounded operands
    xr, wrf, br, ir, gr, ber = (t.float().clone().requires_grad_(True) for t in (x, w, bias, ident, gamma, beta))
    y = F.linear(xr, wrf, br)
    r_ref = ir + (y * scale.view(B, 1, 1) if with_scale else y)
    n_ref = F.layer_norm(r_ref, (N,), gr, ber, 1e-5)
    ((r_ref * wr).sum() + (n_ref * wn).sum()).backward()
    xc, wc, bc, ic, gc, bec = (t.clone().cuda().requires_grad_(True) for t in (x, w, bias, ident, gamma, beta))
    # (the dispatch rule sends the 256-wide / long-K encoder FFN to the un-fused pair; the kernel itself is tested for it too)
    assert ops.linear_add_ln_supported(xc, wc, ic) or (N == 256 and K >= 512)
    r, n = ops.linear_add_ln(xc, wc, bc, ic, scale.cuda() if with_scale else None, gc, bec, 1e-5)
    ((r.float() * wr.cuda()).sum() + (n.float() * wn.cuda()).sum()).backward()
    torch.cuda.synchronize()
    assert rel(r, r_ref) < 5e-3 and rel(n, n_ref) < 8e-3
    assert rel(xc.grad, xr.grad) < 2e-2
    assert rel(wc.grad, wrf.grad) < 2e-2
    assert rel(bc.grad, br.grad) < 2e-2
    assert rel(ic.grad, ir.grad) < 2e-2
    assert rel(gc.grad, gr.grad) < 2e-2 and rel(bec.grad, ber.grad) < 2e-2


# ---------------------------------------------------------------------------
# convolutions as im2col GEMMs, PPM pooling (csrc/conv_ops.cu + gemm_tc.cu)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('B,Cin,Cout,H,W,k,stride,pad', [
    (2, 192, 256, 100, 100, 1, 1, 0),      # ChannelMapper lateral (stage-1 map of an 800^2 image)
    (2, 768, 256, 25, 25, 3, 2, 1),        # ChannelMapper extra level: 3x3 stride 2 -> 13x13
    (2, 512, 512, 64, 64, 3, 1, 1),        # UPerNet fpn_convs
    (1, 96, 64, 17, 23, 3, 1, 1), (1, 64, 128, 9, 9, 5, 2, 2), (3, 256, 256, 50, 50, 1, 1, 0)])
@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32])
def test_conv2d_as_im2col_gemm(B, Cin, Cout, H, W, k, stride, pad, dtype):
    """ops.conv2d == F.conv2d: values, input / weight / bias gradients; channels-last and NCHW-contiguous inputs"""
    from rscotr_b200 import ops
    g = torch.Generator().manual_seed(Cin + Cout + H + k)
    x = torch.randn(B, Cin, H, W, generator=g).to(dtype)
    w = (torch.randn(Cout, Cin, k, k, generator=g) * (Cin * k * k) ** -0.5).to(dtype)
    b = torch.randn(Cout, generator=g) * 0.2
    xr, wr, br = (t.float().clone().requires_grad_(True) for t in (x, w, b))
    yr = F.conv2d(xr, wr, br, stride=stride, padding=pad)
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy)
    for cl in (True, False):
        xc = x.cuda()
        if cl:
            xc = xc.contiguous(memory_format=torch.channels_last)
        xc = xc.requires_grad_(True)
        wc, bc = w.clone().cuda().requires_grad_(True), b.clone().cuda().requires_grad_(True)
        assert ops.conv2d_supported(xc, wc, stride, pad)
        y = ops.conv2d(xc, wc, bc, stride, pad)
        assert y.shape == yr.shape
        y.backward(gy.cuda().to(dtype))
        tol = 1e-2 if dtype == torch.bfloat16 else 1e-4
        assert rel(y, yr) < tol and rel(xc.grad, xr.grad) < 2 * tol
        assert rel(wc.grad, wr.grad) < 2 * tol and rel(bc.grad, br.grad) < 2 * tol


@pytest.mark.parametrize('B,C,H,W,S', [(2, 1024, 16, 16, 1), (2, 1024, 16, 16, 2), (2, 1024, 16, 16, 3), (2, 1024, 16, 16, 6),
                                       (1, 768, 25, 25, 6), (1, 64, 4, 5, 6), (3, 128, 13, 13, 3)])
@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32])
def test_adaptive_avg_pool(B, C, H, W, S, dtype):
    """PPM pooling == nn.AdaptiveAvgPool2d incl. overlapping bins (H % S != 0) and bins finer than the map (H < S)"""
    from rscotr_b200 import ops
    g = torch.Generator().manual_seed(C + H + S)
    x = torch.randn(B, C, H, W, generator=g).to(dtype)
    xr = x.float().clone().requires_grad_(True)
    yr = F.adaptive_avg_pool2d(xr, S)
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy)
    xc = x.cuda().requires_grad_(True)
    y = ops.adaptive_avg_pool2d(xc, S)
    y.backward(gy.cuda().to(dtype))
    tol = 6e-3 if dtype == torch.bfloat16 else 1e-5
    assert y.shape == yr.shape and rel(y, yr) < tol and rel(xc.grad, xr.grad) < tol


@pytest.mark.parametrize('M,N,K', [(1100, 256, 256), (1100, 512, 256), (100, 384, 256), (37, 8, 16), (2200, 2048, 256), (1100, 256, 2048),
                                   (200, 256, 256), (900, 96, 264), (4000, 80, 256), (1, 256, 256)])
@pytest.mark.parametrize('want', ['all', 'no_dx', 'no_db'])
def test_small_linear_backward_one_launch(M, N, K, want):
    """rsc_small_linear_bwd (dX, dW +=, db += of a small Linear layer in one launch) against fp32 torch on the same
    bf16-rounded operands; dW / db are ACCUMULATED into pre-filled fp32 buffers (row stride of dW > K: a view of a wider buffer)"""
    g = torch.Generator().manual_seed(M + N + K)
    dy = torch.randn(M, N, generator=g).bfloat16()
    x = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16()
    dw0 = torch.randn(N, K + 8, generator=g)
    db0 = torch.randn(N, generator=g)
    want_dx = dy.float() @ w.float()
    want_dw = dw0[:, :K] + dy.float().t() @ x.float()
    want_db = db0 + dy.float().sum(0)
    dyc, xc, wc = dy.cuda(), x.cuda(), w.cuda()
    dwc, dbc = dw0.cuda(), db0.cuda()
    dxc = torch.empty(M, K, dtype=torch.bfloat16, device='cuda') if want != 'no_dx' else None
    _call('rsc_small_linear_bwd', dyc.data_ptr(), xc.data_ptr(), wc.data_ptr(), None if dxc is None else dxc.data_ptr(),
          dwc.data_ptr(), None if want == 'no_db' else dbc.data_ptr(), M, N, K, N, K, K, K, K + 8, _stream())
    torch.cuda.synchronize()
    if dxc is not None:
        assert rel(dxc, want_dx) < 6e-3, rel(dxc, want_dx)
    assert rel(dwc[:, :K], want_dw) < 1e-3, rel(dwc[:, :K], want_dw)
    assert torch.equal(dwc[:, K:].cpu(), dw0[:, K:])                  # nothing written outside the (N, K) view
    if want != 'no_db':
        assert rel(dbc, want_db) < 1e-3
    else:
        assert torch.equal(dbc.cpu(), db0)


@pytest.mark.parametrize('M', [300, 5000, 13294])
@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32])
def test_linear_pair_one_gemm_for_two_sibling_layers(M, dtype):
    """ops.linear_pair: sampling_offsets (256 -> 256) and attention_weights (256 -> 128) of mmcv
    MultiScaleDeformableAttention as ONE GEMM over parameter views stacked in flat buffers (the layout
    StepEngine._pair_linears produces), forward and backward, against the two separate fp32 Linears.  M = 300 takes the
    library GEMMs, M >= 4096 the tcgen05 kernels."""
    from rscotr_b200 import ops
    g = torch.Generator().manual_seed(M)
    K, N1, N2 = 256, 256, 128
    sizes = [N1 * K, N2 * K, N1, N2]
    total = sum(sizes)
    flat = (torch.randn(total, generator=g) * 0.06).cuda()
    flat_lp = flat.bfloat16()
    flat_g = torch.zeros(total, device='cuda')
    m1, m2 = torch.nn.Linear(K, N1).cuda(), torch.nn.Linear(K, N2).cuda()
    offs = [0, N1 * K, (N1 + N2) * K, (N1 + N2) * K + N1]
    for p, o, n in zip((m1.weight, m2.weight, m1.bias, m2.bias), offs, sizes):
        p.data = flat[o:o + n].view_as(p)
        p._rsc_lp = flat_lp[o:o + n].view_as(p)
        p._rsc_g = flat_g[o:o + n].view_as(p)
    pv = ops.linear_pair_views(m1, m2)
    assert pv is not None and pv['w_lp'].shape == (N1 + N2, K) and pv['gb'].shape == (N1 + N2,)
    x = torch.randn(2, M // 2, K, generator=g).to(dtype).cuda().requires_grad_(True)
    assert ops.linear_pair_supported(x, pv)
    wy = torch.randn(2, M // 2, N1 + N2, generator=g).cuda()
    y = ops.linear_pair(x, m1, m2, pv)
    assert y.shape == (2, M // 2, N1 + N2) and y.dtype == dtype
    (y.float() * wy).sum().backward()
    torch.cuda.synchronize()
    # fp32 reference on the operands the GEMM read
    wsrc = flat_lp.float() if dtype == torch.bfloat16 else flat
    W = wsrc[:(N1 + N2) * K].view(N1 + N2, K).cpu()
    b = flat[(N1 + N2) * K:].cpu()
    xr = x.detach().float().cpu().requires_grad_(True)
    Wr, br = W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.linear(xr, Wr, br)
    dy = wy.cpu().to(dtype).float()        # the gradient arrives in the compute dtype
    (yr * dy).sum().backward()
    lo = dtype == torch.bfloat16
    assert rel(y, yr) < (4e-3 if lo else 1e-5)
    assert rel(x.grad, xr.grad) < (4e-3 if lo else 1e-5)
    assert rel(flat_g[:(N1 + N2) * K].view(N1 + N2, K), Wr.grad) < (4e-3 if lo else 1e-4)
    assert rel(flat_g[(N1 + N2) * K:], br.grad) < (4e-3 if lo else 1e-4)
    assert m1.weight.grad is None and m2.bias.grad is None           # autograd saw no parameter gradient
