"""torch.autograd wrappers over the C-ABI kernels (include/rscotr.h).

PyTorch is plumbing here: it owns device memory and the stream; all arithmetic
of these ops happens in librscotr_b200.so.  Every op raises on a non-CUDA
tensor -- there is deliberately no CPU / eager fallback.
"""
import torch

from . import _lib
from ._lib import RSC_BF16, RSC_F32, call


def _dt(t):
    if t.dtype == torch.float32:
        return RSC_F32
    if t.dtype == torch.bfloat16:
        return RSC_BF16
    raise TypeError('rscotr_b200 kernels take float32 or bfloat16 activations, got %s' % t.dtype)


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError('rscotr_b200 ops run on CUDA tensors only (no CPU fallback)')


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _f32(t):
    return None if t is None else t.detach().float().contiguous()


# ---------------------------------------------------------------------------
# window index maps / partition / reverse        
# ---------------------------------------------------------------------------
def _nwin(H, W, ws):
    return ((H + ws - 1) // ws) * ((W + ws - 1) // ws)


def _padded_tokens(B, H, W, ws):
    """B * Lp of SURVEY 8d: tokens of the window-padded grid."""
    return B * _nwin(H, W, ws) * ws * ws


def window_index_partition(B, H, W, ws, shift, device='cuda'):
    out = torch.empty(B * _nwin(H, W, ws) * ws * ws, dtype=torch.int64, device=device)
    _cuda(out)
    with torch.cuda.device(out.device):
        call('rsc_window_index_partition', out.data_ptr(), B, H, W, ws, shift, _stream())
    return out


def window_index_reverse(B, H, W, ws, shift, device='cuda'):
    out = torch.empty(B * H * W, dtype=torch.int64, device=device)
    _cuda(out)
    with torch.cuda.device(out.device):
        call('rsc_window_index_reverse', out.data_ptr(), B, H, W, ws, shift, _stream())
    return out


def _partition_raw(x, ws, shift):
    B, H, W, C = x.shape
    out = torch.empty(B * _nwin(H, W, ws), ws * ws, C, dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        call('rsc_window_partition', x.data_ptr(), out.data_ptr(), B, H, W, C, ws, shift, _dt(x), _stream())
    return out


def _reverse_raw(win, B, H, W, ws, shift):
    C = win.shape[-1]
    out = torch.empty(B, H, W, C, dtype=win.dtype, device=win.device)
    with torch.cuda.device(win.device):
        call('rsc_window_reverse', win.data_ptr(), out.data_ptr(), B, H, W, C, ws, shift, _dt(win), _stream())
    return out


class _WindowPartition(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ws, shift):
        _cuda(x)
        ctx.meta = (x.shape, ws, shift)
        return _partition_raw(x.contiguous(), ws, shift)

    @staticmethod
    def backward(ctx, g):
        (B, H, W, C), ws, shift = ctx.meta
        return _reverse_raw(g.contiguous(), B, H, W, ws, shift), None, None


class _WindowReverse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, win, B, H, W, ws, shift):
        _cuda(win)
        ctx.meta = (ws, shift)
        return _reverse_raw(win.contiguous(), B, H, W, ws, shift)

    @staticmethod
    def backward(ctx, g):
        ws, shift = ctx.meta
        return _partition_raw(g.contiguous(), ws, shift), None, None, None, None, None


def window_partition(x, ws, shift=0):
    """(B,H,W,C) -> (B*nW, ws*ws, C): pad + roll(-shift) + window_partition."""
    return _WindowPartition.apply(x, ws, shift)


def window_reverse(windows, B, H, W, ws, shift=0):
    """(B*nW, ws*ws, C) -> (B,H,W,C): window_reverse + roll(+shift) + crop."""
    return _WindowReverse.apply(windows, B, H, W, ws, shift)


# ---------------------------------------------------------------------------
# fused shifted-window attention core            
# ---------------------------------------------------------------------------
class _WMSA(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, qkv_bias, table, H, W, heads, ws, shift, scale, fg):
        _cuda(qkv, table)
        B = qkv.shape[0]
        C = qkv.shape[-1] // 3
        qkv = qkv.contiguous()
        bias32 = _f32(qkv_bias)
        table32 = _f32(table)
        out = torch.empty(B, H * W, C, dtype=qkv.dtype, device=qkv.device)
        with torch.cuda.device(qkv.device):
            call('rsc_wmsa_fwd', qkv.data_ptr(), _p(bias32), table32.data_ptr(), out.data_ptr(), B, H, W, C, heads,
                 ws, shift, scale, _dt(qkv), _stream(), alg_bytes=4 * _padded_tokens(B, H, W, ws) * C * qkv.element_size())
        ctx.save_for_backward(qkv, bias32, table32)
        ctx.meta = (B, H, W, C, heads, ws, shift, scale, qkv_bias is not None and qkv_bias.dtype, table.dtype, fg)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, bias32, table32 = ctx.saved_tensors
        B, H, W, C, heads, ws, shift, scale, bias_dtype, table_dtype, fg = ctx.meta
        dout = dout.contiguous()
        dqkv = torch.empty_like(qkv)
        # the kernel ACCUMULATES d(table) / d(qkv bias) with atomics: straight into the flat gradient buffer when
        # the step engine attached one
        direct = fg is not None and fg[1] is not None and (bias32 is None or fg[0] is not None)
        dtable = fg[1] if direct else torch.zeros_like(table32)
        dbias = None if bias32 is None else (fg[0] if direct else torch.zeros_like(bias32))
        with torch.cuda.device(qkv.device):
            call('rsc_wmsa_bwd', qkv.data_ptr(), _p(bias32), table32.data_ptr(), dout.data_ptr(), dqkv.data_ptr(),
                 dtable.data_ptr(), _p(dbias), B, H, W, C, heads, ws, shift, scale, _dt(qkv), _stream(),
                 alg_bytes=7 * _padded_tokens(B, H, W, ws) * C * qkv.element_size())
        if direct:
            return (dqkv,) + (None,) * 10
        if dbias is not None:
            dbias = dbias.to(bias_dtype)
        return dqkv, dbias, dtable.to(table_dtype), None, None, None, None, None, None, None


def wmsa(qkv, qkv_bias, table, hw, heads, ws=7, shift=0, scale=None):
    """qkv (B, H*W, 3C) un-padded tokens -> (B, H*W, C).

    qkv_bias: the qkv Linear's bias (value of a zero-padded token's row) or None;
    table: relative_position_bias_table ((2ws-1)^2, heads)."""
    H, W = hw
    C = qkv.shape[-1] // 3
    if scale is None:
        scale = (C // heads) ** -0.5
    fg = (None if qkv_bias is None else _flat_grad(qkv_bias), _flat_grad(table))
    return _WMSA.apply(qkv, qkv_bias, table, H, W, heads, ws, shift, float(scale), fg)


# ---------------------------------------------------------------------------
# PatchMerging gather + LayerNorm                 
# ---------------------------------------------------------------------------
class _PatchMergeLN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, H, W, eps):
        _cuda(x, gamma, beta)
        B, L, C = x.shape
        x = x.contiguous()
        g32, b32 = _f32(gamma), _f32(beta)
        Ho, Wo = (H + 1) // 2, (W + 1) // 2
        y = torch.empty(B, Ho * Wo, 4 * C, dtype=x.dtype, device=x.device)
        mean = torch.empty(B * Ho * Wo, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        with torch.cuda.device(x.device):
            call('rsc_patch_merge_ln_fwd', x.data_ptr(), g32.data_ptr(), b32.data_ptr(), y.data_ptr(),
                 mean.data_ptr(), rstd.data_ptr(), B, H, W, C, eps, _dt(x), _stream(),
                 alg_bytes=2 * x.numel() * x.element_size())
        ctx.save_for_backward(x, g32, mean, rstd)
        ctx.meta = (B, H, W, C, gamma.dtype, beta.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, g32, mean, rstd = ctx.saved_tensors
        B, H, W, C, gdt, bdt = ctx.meta
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dg = torch.zeros(4 * C, dtype=torch.float32, device=x.device)
        db = torch.zeros_like(dg)
        with torch.cuda.device(x.device):
            call('rsc_patch_merge_ln_bwd', x.data_ptr(), g32.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                 dy.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), B, H, W, C, _dt(x), _stream(),
                 alg_bytes=3 * x.numel() * x.element_size())
        return dx, dg.to(gdt), db.to(bdt), None, None, None


def patch_merge_ln(x, hw, gamma, beta, eps=1e-5):
    """x (B,H*W,C) -> LayerNorm(unfold2x2(x)) (B, ceil(H/2)*ceil(W/2), 4C)."""
    return _PatchMergeLN.apply(x, gamma, beta, hw[0], hw[1], float(eps))


# ---------------------------------------------------------------------------
# LayerNorm                                       
# ---------------------------------------------------------------------------
class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps, out_dtype, fg):
        _cuda(x, gamma, beta)
        C = x.shape[-1]
        xc = x.contiguous()
        rows = xc.numel() // C
        g32, b32 = _f32(gamma), _f32(beta)
        y = torch.empty(xc.shape, dtype=out_dtype, device=x.device)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        with torch.cuda.device(x.device):
            call('rsc_layernorm_fwd', xc.data_ptr(), g32.data_ptr(), b32.data_ptr(), y.data_ptr(), mean.data_ptr(),
                 rstd.data_ptr(), rows, C, eps, _dt(xc), _dt(y), _stream(),
                 alg_bytes=xc.numel() * (xc.element_size() + y.element_size()))
        ctx.save_for_backward(xc, g32, mean, rstd)
        ctx.meta = (rows, C, gamma.dtype, beta.dtype, fg[0], fg[1])
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, g32, mean, rstd = ctx.saved_tensors
        rows, C, gdt, bdt, gg, gb = ctx.meta
        dy = dy.contiguous()
        dx = torch.empty_like(xc)
        direct = gg is not None and gb is not None     # the kernel ACCUMULATES d(gamma) / d(beta) with atomics
        dg = gg if direct else torch.zeros(C, dtype=torch.float32, device=xc.device)
        db = gb if direct else torch.zeros_like(dg)
        with torch.cuda.device(xc.device):
            call('rsc_layernorm_bwd', xc.data_ptr(), g32.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dy.data_ptr(),
                 dx.data_ptr(), dg.data_ptr(), db.data_ptr(), rows, C, _dt(xc), _dt(dy), _stream(),
                 alg_bytes=xc.numel() * (2 * xc.element_size() + dy.element_size()))
        if direct:
            return dx, None, None, None, None, None
        return dx, dg.to(gdt), db.to(bdt), None, None, None


def _flat_grad(p):
    """the step engine's fp32 gradient view of parameter p (None outside the engine / without autograd)."""
    g = getattr(p, '_rsc_g', None)
    if g is None or not torch.is_grad_enabled() or not p.requires_grad or g.dtype != torch.float32:
        return None
    return g


# ---------------------------------------------------------------------------
# fused residual-stream passes of the Swin block   
# ---------------------------------------------------------------------------
def add_ln_supported(x):
    return x.is_cuda and x.dtype in (torch.float32, torch.bfloat16) and bool(_lib.lib().rsc_add_ln_supported(x.shape[-1]))


class _AddLN(torch.autograd.Function):
    """(r, n) = (identity + (x + bias) * scale[sample], LayerNorm(r)); see include/rscotr.h."""

    @staticmethod
    def forward(ctx, identity, x, bias, scale, gamma, beta, eps, fg):
        ctx.set_materialize_grads(False)      # an unused output (post-norm layers drop r) costs no zero fill / read
        _cuda(identity, x, gamma, beta)
        C = x.shape[-1]
        idc, xc = identity.contiguous(), x.contiguous()
        assert idc.dtype == xc.dtype and idc.shape == xc.shape
        rows = xc.numel() // C
        rps = rows // xc.shape[0]
        b32, s32, g32, be32 = _f32(bias), _f32(scale), _f32(gamma), _f32(beta)
        r = torch.empty(xc.shape, dtype=xc.dtype, device=xc.device)      # (canonical strides, not xc's)
        n = torch.empty(xc.shape, dtype=xc.dtype, device=xc.device)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        with torch.cuda.device(x.device):
            call('rsc_add_ln_fwd', idc.data_ptr(), xc.data_ptr(), _p(b32), _p(s32), g32.data_ptr(), be32.data_ptr(),
                 r.data_ptr(), n.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, rps, C, float(eps), _dt(xc), _stream(),
                 alg_bytes=4 * xc.numel() * xc.element_size())
        ctx.save_for_backward(r, g32, mean, rstd, s32)
        ctx.meta = (rows, rps, C, None if bias is None else bias.dtype, gamma.dtype, beta.dtype,
                    fg[0], fg[1], fg[2],
                    ctx.needs_input_grad[1])
        return r, n

    @staticmethod
    def backward(ctx, dr_ext, dn):
        r, g32, mean, rstd, s32 = ctx.saved_tensors
        rows, rps, C, bdt, gdt, bedt, gbias, gg, gb, need_dx = ctx.meta
        dev = r.device
        if dn is None:
            dn = torch.zeros_like(r)
        dn = dn.contiguous()
        dr_ext = None if dr_ext is None else dr_ext.contiguous()
        d_id = torch.empty_like(r)
        dx = torch.empty_like(r) if s32 is not None else None
        direct = gg is not None and gb is not None
        dg = gg if direct else torch.zeros(C, dtype=torch.float32, device=dev)
        db = gb if direct else torch.zeros(C, dtype=torch.float32, device=dev)
        dbias = None
        if bdt is not None:
            dbias = gbias if gbias is not None else torch.zeros(C, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            call('rsc_add_ln_bwd', r.data_ptr(), g32.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dn.data_ptr(),
                 _p(dr_ext), _p(s32), d_id.data_ptr(), _p(dx), dg.data_ptr(), db.data_ptr(), _p(dbias), rows, rps, C,
                 _dt(r), _stream(), alg_bytes=(4 + (dr_ext is not None) + (dx is not None)) * r.numel() * r.element_size())
        return (d_id, (dx if dx is not None else d_id) if need_dx else None,
                None if (bdt is None or gbias is not None) else dbias.to(bdt), None,
                None if direct else dg.to(gdt), None if direct else db.to(bedt), None, None)


def add_ln(identity, x, bias, scale, gamma, beta, eps=1e-5):
    """r = identity + (x + bias) * scale[b]; n = LayerNorm(r).  bias (C,) / scale (B,) may be None."""
    fg = (None if bias is None else _flat_grad(bias), _flat_grad(gamma), _flat_grad(beta))
    return _AddLN.apply(identity, x, bias, scale, gamma, beta, eps, fg)


ACT_GELU, ACT_RELU, ACT_GELU_SIG = 0, 1, 2
# bf16 GELU through the logistic fit of the normal CDF (|error| < 2.6e-5, half the ALU work; the same function the GEMM
# epilogue of csrc/gemm_tc.cu evaluates through one tanh).  A/B'd on B200 in round 2: promoted; RSC_GELU_SIG=0 restores erf.
_GELU_SIG = __import__('os').environ.get('RSC_GELU_SIG', '1') != '0'


class _BiasAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, bias, act, fg):
        _cuda(h, bias)
        hc = h.contiguous()
        C = hc.shape[-1]
        rows = hc.numel() // C
        b32 = _f32(bias)
        y = torch.empty(hc.shape, dtype=hc.dtype, device=hc.device)
        with torch.cuda.device(h.device):
            call('rsc_bias_act_fwd', hc.data_ptr(), b32.data_ptr(), y.data_ptr(), rows, C, act, _dt(hc), _stream(),
                 alg_bytes=2 * hc.numel() * hc.element_size())
        ctx.save_for_backward(hc, b32)
        ctx.meta = (rows, C, act, bias.dtype, fg)
        return y

    @staticmethod
    def backward(ctx, dy):
        hc, b32 = ctx.saved_tensors
        rows, C, act, bdt, gbias = ctx.meta
        dy = dy.contiguous()
        dh = torch.empty_like(hc)
        dbias = gbias if gbias is not None else torch.zeros(C, dtype=torch.float32, device=hc.device)
        with torch.cuda.device(hc.device):
            call('rsc_bias_act_bwd', hc.data_ptr(), b32.data_ptr(), dy.data_ptr(), dh.data_ptr(), dbias.data_ptr(), rows, C,
                 act, _dt(hc), _stream(), alg_bytes=3 * hc.numel() * hc.element_size())
        return dh, None if gbias is not None else dbias.to(bdt), None, None


def bias_gelu(h, bias):
    """gelu(h + bias) (erf form); the backward also produces the bias gradient (column sums) in the same pass."""
    act = ACT_GELU_SIG if (_GELU_SIG and h.dtype == torch.bfloat16) else ACT_GELU
    return _BiasAct.apply(h, bias, act, _flat_grad(bias))


def bias_relu(h, bias):
    """relu(h + bias) with the bias gradient out of the same backward pass."""
    return _BiasAct.apply(h, bias, ACT_RELU, _flat_grad(bias))


def bias_act(h, bias, act):
    return _BiasAct.apply(h, bias, act, _flat_grad(bias))


def bias_act_supported(x, out_features):
    return x.is_cuda and out_features % 8 == 0 and out_features <= 8192 and \
        (torch.get_autocast_dtype('cuda') if torch.is_autocast_enabled('cuda') else x.dtype) in (torch.float32, torch.bfloat16)


def layer_norm(x, gamma, beta, eps=1e-5, out_dtype=None):
    """LayerNorm over the last dim; output dtype defaults to the active autocast dtype
    (so the GEMM that follows reads bf16 directly) or x.dtype outside autocast."""
    if out_dtype is None:
        out_dtype = torch.get_autocast_dtype('cuda') if torch.is_autocast_enabled('cuda') else x.dtype
    # (the flat-gradient views are looked up HERE: inside Function.forward the parameters are re-wrapped and lose
    # the attributes the step engine attached)
    return _LayerNorm.apply(x, gamma, beta, float(eps), out_dtype, (_flat_grad(gamma), _flat_grad(beta)))


# ---------------------------------------------------------------------------
# multi-scale deformable attention                
# ---------------------------------------------------------------------------
class MultiScaleDeformableAttnFunction(torch.autograd.Function):
    """Same call signature as mmcv.ops.multi_scale_deform_attn.
    MultiScaleDeformableAttnFunction.apply(value, value_spatial_shapes,
    value_level_start_index, sampling_locations, attention_weights, im2col_step)."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step=64):
        _cuda(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights)
        B, Nv, heads, D = value.shape
        _, Nq, _, L, P, _ = sampling_locations.shape
        if D != 32:
            raise RuntimeError('rsc_msda: head dim must be 32, got %d' % D)
        if int(im2col_step) <= 0 or B % min(B, int(im2col_step)) != 0:
            raise RuntimeError('batch(%d) must divide im2col_step(%d)' % (B, im2col_step))
        value = value.contiguous()
        shapes = value_spatial_shapes.to(torch.int64).contiguous()
        starts = value_level_start_index.to(torch.int64).contiguous()
        loc = sampling_locations.float().contiguous()
        aw = attention_weights.float().contiguous()
        out = torch.empty(B, Nq, heads * D, dtype=value.dtype, device=value.device)
        with torch.cuda.device(value.device):
            call('rsc_msda_fwd', value.data_ptr(), shapes.data_ptr(), starts.data_ptr(), loc.data_ptr(),
                 aw.data_ptr(), out.data_ptr(), B, Nv, Nq, heads, L, P, int(im2col_step), _dt(value), _stream(),
                 alg_bytes=(value.numel() + out.numel()) * value.element_size() + (loc.numel() + aw.numel()) * 4)
        ctx.save_for_backward(value, shapes, starts, loc, aw)
        ctx.meta = (int(im2col_step), sampling_locations.dtype, attention_weights.dtype)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        value, shapes, starts, loc, aw = ctx.saved_tensors
        im2col_step, loc_dt, aw_dt = ctx.meta
        B, Nv, heads, D = value.shape
        _, Nq, _, L, P, _ = loc.shape
        grad_output = grad_output.contiguous()
        gv = torch.zeros(value.shape, dtype=torch.float32, device=value.device)
        gl = torch.empty_like(loc)
        ga = torch.empty_like(aw)
        with torch.cuda.device(value.device):
            call('rsc_msda_bwd', value.data_ptr(), shapes.data_ptr(), starts.data_ptr(), loc.data_ptr(),
                 aw.data_ptr(), grad_output.data_ptr(), gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), B, Nv, Nq, heads,
                 L, P, im2col_step, _dt(value), _stream(),
                 alg_bytes=(value.numel() + grad_output.numel()) * value.element_size() + 2 * value.numel() * 4 +
                 2 * (loc.numel() + aw.numel()) * 4)
        return gv.to(value.dtype), None, None, gl.to(loc_dt), ga.to(aw_dt), None


class _MSDAFused(torch.autograd.Function):
    """softmax(logits) + sampling locations from (reference points, raw offsets) + ms_deform_attn in one kernel
    (rsc_msda_fused_{fwd,bwd}); value (B,Nv,heads,32), offsets (B,Nq,heads,L,P,2), logits (B,Nq,heads,L*P),
    ref (B,Nq,L,2|4) fp32 without gradient."""

    @staticmethod
    def forward(ctx, value, shapes, starts, offsets, logits, ref):
        _cuda(value, shapes, starts, offsets, logits, ref)
        B, Nv, heads, D = value.shape
        _, Nq, _, L, P, _ = offsets.shape
        value = value.contiguous()
        offsets = offsets.contiguous()
        logits = logits.to(offsets.dtype).contiguous()
        ref = ref.float().contiguous()
        shapes = shapes.to(torch.int64).contiguous()
        starts = starts.to(torch.int64).contiguous()
        out = torch.empty(B, Nq, heads * D, dtype=value.dtype, device=value.device)
        with torch.cuda.device(value.device):
            call('rsc_msda_fused_fwd', value.data_ptr(), shapes.data_ptr(), starts.data_ptr(), offsets.data_ptr(),
                 logits.data_ptr(), ref.data_ptr(), out.data_ptr(), B, Nv, Nq, heads, L, P, ref.shape[-1], _dt(value),
                 _dt(offsets), 0, 0, _stream(),
                 alg_bytes=(value.numel() + out.numel()) * value.element_size() +
                 (offsets.numel() + logits.numel()) * offsets.element_size())
        ctx.save_for_backward(value, shapes, starts, offsets, logits, ref)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        value, shapes, starts, offsets, logits, ref = ctx.saved_tensors
        B, Nv, heads, D = value.shape
        _, Nq, _, L, P, _ = offsets.shape
        grad_output = grad_output.contiguous()
        gv = torch.zeros(value.shape, dtype=torch.float32, device=value.device)
        go = torch.empty_like(offsets)
        gl = torch.empty_like(logits)
        with torch.cuda.device(value.device):
            call('rsc_msda_fused_bwd', value.data_ptr(), shapes.data_ptr(), starts.data_ptr(), offsets.data_ptr(),
                 logits.data_ptr(), ref.data_ptr(), grad_output.data_ptr(), gv.data_ptr(), go.data_ptr(), gl.data_ptr(),
                 B, Nv, Nq, heads, L, P, ref.shape[-1], _dt(value), _dt(offsets), 0, 0, _stream(),
                 alg_bytes=(value.numel() + grad_output.numel()) * value.element_size() + 2 * value.numel() * 4 +
                 2 * (offsets.numel() + logits.numel()) * offsets.element_size())
        return gv.to(value.dtype), None, None, go, gl, None


class _MSDAFusedPacked(torch.autograd.Function):
    """_MSDAFused with the raw offsets and attention logits as column ranges of ONE matrix `both`
    (B, Nq, heads*L*P*3) = [offsets | logits] -- the output of a single GEMM over the stacked sampling_offsets /
    attention_weights weights (linear_pair); the kernels take row strides, and the backward writes both gradients
    into one matrix for that GEMM's backward."""

    @staticmethod
    def forward(ctx, value, shapes, starts, both, ref, L, P):
        _cuda(value, shapes, starts, both, ref)
        B, Nv, heads, D = value.shape
        Nq = both.shape[1]
        n_off = heads * L * P * 2
        W = both.shape[-1]
        assert W == n_off + heads * L * P and both.is_contiguous()
        value = value.contiguous()
        ref = ref.float().contiguous()
        shapes = shapes.to(torch.int64).contiguous()
        starts = starts.to(torch.int64).contiguous()
        out = torch.empty(B, Nq, heads * D, dtype=value.dtype, device=value.device)
        es = both.element_size()
        with torch.cuda.device(value.device):
            call('rsc_msda_fused_fwd', value.data_ptr(), shapes.data_ptr(), starts.data_ptr(), both.data_ptr(),
                 both.data_ptr() + n_off * es, ref.data_ptr(), out.data_ptr(), B, Nv, Nq, heads, L, P, ref.shape[-1],
                 _dt(value), _dt(both), W, W, _stream(),
                 alg_bytes=(value.numel() + out.numel()) * value.element_size() + both.numel() * es)
        ctx.save_for_backward(value, shapes, starts, both, ref)
        ctx.lp = (L, P)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        value, shapes, starts, both, ref = ctx.saved_tensors
        L, P = ctx.lp
        B, Nv, heads, D = value.shape
        Nq, W = both.shape[1], both.shape[-1]
        n_off = heads * L * P * 2
        grad_output = grad_output.contiguous()
        gv = torch.zeros(value.shape, dtype=torch.float32, device=value.device)
        gboth = torch.empty_like(both)
        es = both.element_size()
        with torch.cuda.device(value.device):
            call('rsc_msda_fused_bwd', value.data_ptr(), shapes.data_ptr(), starts.data_ptr(), both.data_ptr(),
                 both.data_ptr() + n_off * es, ref.data_ptr(), grad_output.data_ptr(), gv.data_ptr(), gboth.data_ptr(),
                 gboth.data_ptr() + n_off * es, B, Nv, Nq, heads, L, P, ref.shape[-1], _dt(value), _dt(both), W, W, _stream(),
                 alg_bytes=(value.numel() + grad_output.numel()) * value.element_size() + 2 * value.numel() * 4 +
                 2 * both.numel() * es)
        return gv.to(value.dtype), None, None, gboth, None, None, None


def msda_fused_supported(value, offsets, ref):
    return (value.is_cuda and value.shape[-1] == 32 and value.shape[-2] % 4 == 0 and
            offsets.shape[-3] * offsets.shape[-2] == 16 and ref.shape[-1] in (2, 4) and not ref.requires_grad and
            value.dtype in (torch.float32, torch.bfloat16) and offsets.dtype in (torch.float32, torch.bfloat16))


def ms_deform_attn_fused(value, spatial_shapes, level_start_index, offsets, logits, reference_points):
    return _MSDAFused.apply(value, spatial_shapes, level_start_index, offsets, logits, reference_points)


def ms_deform_attn_fused_packed(value, spatial_shapes, level_start_index, both, reference_points, num_levels, num_points):
    """`both` (B, Nq, heads*L*P*3) = [raw offsets | attention logits] of linear_pair"""
    return _MSDAFusedPacked.apply(value, spatial_shapes, level_start_index, both, reference_points, num_levels, num_points)


def msda_packed_supported(value, both, ref, num_levels, num_points):
    return (value.is_cuda and value.shape[-1] == 32 and value.shape[-2] % 4 == 0 and num_levels * num_points == 16 and
            ref.shape[-1] in (2, 4) and not ref.requires_grad and value.dtype in (torch.float32, torch.bfloat16) and
            both.dtype in (torch.float32, torch.bfloat16) and both.shape[-1] == value.shape[-2] * 48)


def ms_deform_attn(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, im2col_step=64):
    return MultiScaleDeformableAttnFunction.apply(value, spatial_shapes, level_start_index, sampling_locations,
                                                  attention_weights, im2col_step)


# ---------------------------------------------------------------------------
# global average pool                             
# ---------------------------------------------------------------------------
class _GAP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, channels_last):
        _cuda(x)
        x = x.contiguous()
        if channels_last:
            B, HW, C = x.shape
        else:
            B, C = x.shape[:2]
            HW = x[0, 0].numel()
        y = torch.empty(B, C, dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            call('rsc_gap_fwd', x.data_ptr(), y.data_ptr(), B, C, HW, int(channels_last), _dt(x), _stream(),
                 alg_bytes=x.numel() * x.element_size())
        ctx.meta = (x.shape, B, C, HW, channels_last)
        return y

    @staticmethod
    def backward(ctx, dy):
        shape, B, C, HW, channels_last = ctx.meta
        dy = dy.contiguous()
        dx = torch.empty(shape, dtype=dy.dtype, device=dy.device)
        with torch.cuda.device(dy.device):
            call('rsc_gap_bwd', dy.data_ptr(), dx.data_ptr(), B, C, HW, int(channels_last), _dt(dy), _stream(),
                 alg_bytes=dx.numel() * dx.element_size())
        return dx, None


def global_avg_pool(x, channels_last=False):
    """x (B,C,H,W) [channels_last=False] or (B,L,C) [True] -> (B,C)."""
    return _GAP.apply(x, channels_last)


# ---------------------------------------------------------------------------
# bilinear resize, align_corners=False            
# ---------------------------------------------------------------------------
class _Bilinear(torch.autograd.Function):
    """NCHW-contiguous planes -> rsc_bilinear_*; maps with channels-last strides (what the convolution / norm kernels
    produce) -> rsc_bilinear_cl_*, output channels-last as well: no transpose copy on either side of the resize."""

    @staticmethod
    def forward(ctx, x, Ho, Wo):
        _cuda(x)
        B, C, Hi, Wi = x.shape
        vec = 16 // x.element_size()
        cl = C % vec == 0 and C >= vec and not x.is_contiguous() and x.permute(0, 2, 3, 1).is_contiguous()
        with torch.cuda.device(x.device):
            if cl:
                y = torch.empty(B, Ho, Wo, C, dtype=x.dtype, device=x.device)
                call('rsc_bilinear_cl_fwd', x.data_ptr(), y.data_ptr(), B, C, Hi, Wi, Ho, Wo, _dt(x), _stream(),
                     alg_bytes=(x.numel() + y.numel()) * x.element_size())
                y = y.permute(0, 3, 1, 2)
            else:
                x = x.contiguous()
                y = torch.empty(B, C, Ho, Wo, dtype=x.dtype, device=x.device)
                call('rsc_bilinear_fwd', x.data_ptr(), y.data_ptr(), B * C, Hi, Wi, Ho, Wo, _dt(x), _stream(),
                     alg_bytes=(x.numel() + y.numel()) * x.element_size())
        ctx.meta = (B, C, Hi, Wi, Ho, Wo, cl)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, C, Hi, Wi, Ho, Wo, cl = ctx.meta
        with torch.cuda.device(dy.device):
            if cl:
                dyl = dy.permute(0, 2, 3, 1)
                if not dyl.is_contiguous():
                    dyl = dyl.contiguous()
                dx = torch.empty(B, Hi, Wi, C, dtype=dy.dtype, device=dy.device)
                call('rsc_bilinear_cl_bwd', dyl.data_ptr(), dx.data_ptr(), B, C, Hi, Wi, Ho, Wo, _dt(dy), _stream(),
                     alg_bytes=(dx.numel() + dy.numel()) * dy.element_size())
                return dx.permute(0, 3, 1, 2), None, None
            dy = dy.contiguous()
            dx = torch.empty(B, C, Hi, Wi, dtype=dy.dtype, device=dy.device)
            call('rsc_bilinear_bwd', dy.data_ptr(), dx.data_ptr(), B * C, Hi, Wi, Ho, Wo, _dt(dy), _stream(),
                 alg_bytes=(dx.numel() + dy.numel()) * dy.element_size())
        return dx, None, None


def bilinear_resize(x, size):
    """F.interpolate(x, size=size, mode='bilinear', align_corners=False) for NCHW x."""
    return _Bilinear.apply(x, int(size[0]), int(size[1]))


# ---------------------------------------------------------------------------
# sigmoid focal loss (element-wise)               
# ---------------------------------------------------------------------------
class SigmoidFocalLossFunction(torch.autograd.Function):
    """mmcv.ops.focal_loss.SigmoidFocalLossFunction with weight=None, reduction='none'."""

    @staticmethod
    def forward(ctx, input, target, gamma=2.0, alpha=0.25):
        _cuda(input, target)
        input = input.contiguous()
        target = target.to(torch.int64).contiguous()
        N, C = input.shape
        out = torch.empty(N, C, dtype=torch.float32, device=input.device)
        with torch.cuda.device(input.device):
            call('rsc_sigmoid_focal_loss_fwd', input.data_ptr(), target.data_ptr(), out.data_ptr(), N, C,
                 float(gamma), float(alpha), _dt(input), _stream(), alg_bytes=input.numel() * (input.element_size() + 4))
        ctx.save_for_backward(input, target)
        ctx.meta = (float(gamma), float(alpha))
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input, target = ctx.saved_tensors
        gamma, alpha = ctx.meta
        N, C = input.shape
        gi = torch.empty(N, C, dtype=torch.float32, device=input.device)
        with torch.cuda.device(input.device):
            call('rsc_sigmoid_focal_loss_bwd', input.data_ptr(), target.data_ptr(), gi.data_ptr(), N, C, gamma, alpha,
                 _dt(input), _stream(), alg_bytes=input.numel() * (input.element_size() + 4))
        return (gi * grad_output).to(input.dtype), None, None, None


def sigmoid_focal_loss(input, target, gamma=2.0, alpha=0.25):
    return SigmoidFocalLossFunction.apply(input, target, gamma, alpha)


# ---------------------------------------------------------------------------
# detection losses: GPU Hungarian matching + fused focal / L1 / GIoU   
# ---------------------------------------------------------------------------
def det_match(cls, box, q0, Nq, gt_labels, gt_boxes, gt_start, img_wh, max_gt, w_cls=2.0, w_reg=5.0, w_iou=2.0,
              alpha=0.25, gamma=2.0, eps=1e-12, want_gt_norm=True):
    """cls (..., B, NqTot, C) logits, box (..., B, NqTot, 4) fp32 cxcywh; matches queries q0:q0+Nq of every
    (leading index, image) problem against that image's gt boxes.  Returns (assign int32 (P, Nq) holding the
    global gt index or -1, cost (P, max_gt, Nq), gt_norm (G, 4))."""
    _cuda(cls, box, gt_boxes, gt_start, img_wh)
    assert cls.is_contiguous() and box.is_contiguous() and box.dtype == torch.float32
    B, NqTot, C = cls.shape[-3:]
    P = cls.numel() // (NqTot * C)
    dev = cls.device
    assign = torch.empty(P, Nq, dtype=torch.int32, device=dev)
    cost = torch.empty(P, max(max_gt, 1), Nq, dtype=torch.float32, device=dev)
    G = gt_boxes.shape[0]
    gt_norm = torch.empty(G, 4, dtype=torch.float32, device=dev) if want_gt_norm else None
    gt_boxes = gt_boxes.float().contiguous()
    with torch.cuda.device(dev):
        call('rsc_det_match', cls.data_ptr(), box.data_ptr(), gt_labels.data_ptr(), gt_boxes.data_ptr(),
             gt_start.data_ptr(), img_wh.data_ptr(), P, B, NqTot, q0, Nq, C, max_gt, w_cls, w_reg, w_iou, alpha, gamma,
             eps, cost.data_ptr(), assign.data_ptr(), _p(gt_norm), _dt(cls), _stream())
    return assign, cost, gt_norm


class _DetLoss(torch.autograd.Function):
    """segs: list of dicts(t=index of the (cls, box) tensor pair, L, B, NqTot, q0, Nq, C, assign, a_ls, row0,
    last_first); tensors = cls0, box0, cls1, box1, ...; returns (rows, 3) fp32 loss sums."""

    @staticmethod
    def forward(ctx, segs, common, rows, *tensors):
        gt_labels, gt_norm, img_wh, cls_factor, pos_factor, hyper = common
        out = torch.zeros(rows, 3, dtype=torch.float32, device=tensors[0].device)
        with torch.cuda.device(out.device):
            for sg in segs:
                cls, box = tensors[2 * sg['t']], tensors[2 * sg['t'] + 1]
                call('rsc_det_loss_fwd', cls.data_ptr(), box.data_ptr(), sg['assign'].data_ptr(), _p(gt_labels),
                     _p(gt_norm), img_wh.data_ptr(), cls_factor.data_ptr() + 4 * sg['f0'],
                     pos_factor.data_ptr() + 4 * sg['f0'], out.data_ptr(), sg['L'], sg['B'], sg['NqTot'], sg['q0'],
                     sg['Nq'], sg['C'], sg['a_ls'], sg['row0'], int(sg['last_first']), *hyper, _dt(cls), _stream())
        ctx.segs, ctx.common = segs, common
        ctx.save_for_backward(*tensors)
        return out

    @staticmethod
    def backward(ctx, dout):
        tensors = ctx.saved_tensors
        gt_labels, gt_norm, img_wh, cls_factor, pos_factor, hyper = ctx.common
        dout = dout.contiguous().float()
        covered = {}
        for sg in ctx.segs:
            covered[sg['t']] = covered.get(sg['t'], 0) + sg['Nq']
        grads = []
        for k in range(len(tensors) // 2):
            full = covered.get(k, 0) == tensors[2 * k].shape[-2]
            mk = torch.empty_like if full else torch.zeros_like
            grads += [mk(tensors[2 * k]), mk(tensors[2 * k + 1])]
        with torch.cuda.device(dout.device):
            for sg in ctx.segs:
                cls, box = tensors[2 * sg['t']], tensors[2 * sg['t'] + 1]
                call('rsc_det_loss_bwd', cls.data_ptr(), box.data_ptr(), sg['assign'].data_ptr(), _p(gt_labels),
                     _p(gt_norm), img_wh.data_ptr(), cls_factor.data_ptr() + 4 * sg['f0'],
                     pos_factor.data_ptr() + 4 * sg['f0'], dout.data_ptr(), grads[2 * sg['t']].data_ptr(),
                     grads[2 * sg['t'] + 1].data_ptr(), sg['L'], sg['B'], sg['NqTot'], sg['q0'], sg['Nq'], sg['C'],
                     sg['a_ls'], sg['row0'], int(sg['last_first']), *hyper, _dt(cls), _stream())
        return (None, None, None) + tuple(grads)


def det_loss(segs, tensors, rows, gt_labels, gt_norm, img_wh, cls_factor, pos_factor, gamma=2.0, alpha=0.25,
             w_cls=1.0, w_l1=5.0, w_iou=2.0, eps=1e-6):
    """Fused sigmoid-focal + L1 + GIoU losses of several query segments -> (rows, 3) tensor of
    (loss_cls, loss_bbox, loss_iou), weighted and averaged (see include/rscotr.h)."""
    _cuda(*tensors)
    for k in range(0, len(tensors), 2):
        assert tensors[k].is_contiguous() and tensors[k + 1].is_contiguous() and tensors[k + 1].dtype == torch.float32
    hyper = (float(gamma), float(alpha), float(w_cls), float(w_l1), float(w_iou), float(eps))
    return _DetLoss.apply(segs, (gt_labels, gt_norm, img_wh, cls_factor, pos_factor, hyper), rows, *tensors)


# ---------------------------------------------------------------------------
# fused bilinear upsample + cross-entropy        
# ---------------------------------------------------------------------------
class _UpsampleCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, label, ignore_index):
        _cuda(logits, label)
        logits, label = logits.contiguous(), label.contiguous()
        B, C, h, w = logits.shape
        H, W = label.shape[-2:]
        stats = torch.zeros(3, dtype=torch.float32, device=logits.device)
        lse = torch.empty(B, H, W, dtype=torch.float32, device=logits.device)
        with torch.cuda.device(logits.device):
            call('rsc_upsample_ce_fwd', logits.data_ptr(), label.data_ptr(), lse.data_ptr(), stats.data_ptr(), B, C, h,
                 w, H, W, int(ignore_index), _dt(logits), _stream(),
                 alg_bytes=logits.numel() * logits.element_size() + label.numel() * 8 + lse.numel() * 4)
        ctx.save_for_backward(logits, label, lse)
        return stats

    @staticmethod
    def backward(ctx, dstats):
        logits, label, lse = ctx.saved_tensors
        B, C, h, w = logits.shape
        H, W = label.shape[-2:]
        g = dstats[:1].float().contiguous()
        dlogits = torch.empty_like(logits)
        ws = torch.empty(B * h * w * 4 * C, dtype=torch.float32, device=logits.device)
        with torch.cuda.device(logits.device):
            call('rsc_upsample_ce_bwd', logits.data_ptr(), label.data_ptr(), lse.data_ptr(), g.data_ptr(),
                 dlogits.data_ptr(), ws.data_ptr(), B, C, h, w, H, W, _dt(logits), _stream(),
                 alg_bytes=2 * logits.numel() * logits.element_size() + label.numel() * 8 + lse.numel() * 4)
        return dlogits, None, None


def upsample_ce(logits, label, ignore_index=255):
    """logits (B,C,h,w), label (B,H,W) int64 -> stats (3,) = [sum of CE over non-ignored pixels of the
    bilinearly (align_corners=False) up-sampled logits, #correct (argmax == label), #non-ignored]."""
    return _UpsampleCE.apply(logits, label, ignore_index)


def upsample_ce_supported(logits, label):
    h, w = logits.shape[-2:]
    H, W = label.shape[-2:]
    return logits.is_cuda and H >= h and W >= w and 2.0 * H / h + 5 <= 24 and 2.0 * W / w + 5 <= 24 and \
        logits.dtype in (torch.float32, torch.bfloat16)


# ---------------------------------------------------------------------------
# patch-embedding gather                          
# ---------------------------------------------------------------------------
def patchify4(x, out_dtype=None):
    """(B,Cin,H,W) NCHW image (no gradient) -> (B*(H/4)*(W/4), Cin*16) rows in Conv2d weight order (c, kh, kw)."""
    _cuda(x)
    assert not x.requires_grad, 'patchify4 is a gather of the input image: no backward'
    x = x.contiguous()
    B, Cin, H, W = x.shape
    if out_dtype is None:
        out_dtype = torch.get_autocast_dtype('cuda') if torch.is_autocast_enabled('cuda') else x.dtype
    y = torch.empty(B * (H // 4) * (W // 4), Cin * 16, dtype=out_dtype, device=x.device)
    with torch.cuda.device(x.device):
        call('rsc_patchify4', x.data_ptr(), y.data_ptr(), B, Cin, H, W, _dt(x), _dt(y), _stream(),
             alg_bytes=x.numel() * x.element_size() + y.numel() * y.element_size())
    return y


# ---------------------------------------------------------------------------
# Linear with the bias gradient on rsc_colsum        (GEMMs stay in the library)
# ---------------------------------------------------------------------------
def colsum(x2d, out=None):
    """(rows, C) -> (C,) fp32 column sums; with `out` (fp32, contiguous) the sums are ADDED to it."""
    _cuda(x2d)
    x2d = x2d.contiguous()
    rows, C = x2d.shape
    y = torch.zeros(C, dtype=torch.float32, device=x2d.device) if out is None else out
    assert y.dtype == torch.float32 and y.is_contiguous() and y.numel() == C
    with torch.cuda.device(x2d.device):
        call('rsc_colsum', x2d.data_ptr(), y.data_ptr(), rows, C, _dt(x2d), _stream(),
             alg_bytes=x2d.numel() * x2d.element_size())
    return y


# ---- tcgen05 GEMM entry points (csrc/gemm_tc.cu) ----------------------------------------------------------
# RSC_OWN_GEMM=0 sends every Linear back to the library GEMMs (A/B switch for the benchmarks)
_OWN_GEMM = __import__('os').environ.get('RSC_OWN_GEMM', '1') != '0'


def _tc2d(t):
    """a bf16 matrix the TMA descriptors can address: last dim contiguous, 16-byte aligned rows"""
    return (t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 8 == 0 and
            t.shape[1] % 8 == 0 and t.data_ptr() % 16 == 0 and t.shape[0] > 0)


_TC_MIN_ROWS = 4096     # fewer token rows than this: the fixed cost of a persistent 200 KB-smem kernel (TMEM allocation,
                        # descriptor fetch, pipeline fill) outweighs the fused epilogue; the decoders' small GEMMs stay on
                        # the library


def _tc_linear_ok(x2, w):
    return _OWN_GEMM and _tc2d(x2) and _tc2d(w) and w.shape[0] % 8 == 0 and x2.shape[0] >= _TC_MIN_ROWS


# Measured on B200 (profiles/r02_gbench_*.jsonl): the single-CTA kernels match or beat the library where the GEMM is
# memory-bound and wherever an element-wise pass rides in the epilogue; a PLAIN GEMM with a long contraction
# (>= 512) is tensor-bound and the library's 2-CTA kernels are ~1.5x faster there.
_PLAIN_MAX_K = int(__import__('os').environ.get('RSC_GEMM_PLAIN_MAX_K', 512))


def _own_plain(contraction):
    return contraction < _PLAIN_MAX_K


def _own_dw(M, N, K):
    return not (M <= 16384 and N * K >= (1 << 20))


def gemm_fwd(x2, w, bias32, act=0):
    """act(x2 w^T + bias) -> (y, h); act: 0 none, 1 GELU (h = pre-activation, kept for the backward), 2 ReLU"""
    M, K = x2.shape
    N = w.shape[0]
    y = torch.empty(M, N, dtype=torch.bfloat16, device=x2.device)
    h = torch.empty_like(y) if act == 1 else None
    with torch.cuda.device(x2.device):
        call('rsc_linear_fwd', x2.data_ptr(), w.data_ptr(), _p(bias32), y.data_ptr(), _p(h), M, N, K, x2.stride(0),
             w.stride(0), N, act, _stream(), alg_bytes=2 * (M * K + N * K + M * N * (2 if act == 1 else 1)), alg_flops=2 * M * N * K)
    return y, h


def gemm_dx(dy2, w, aux=None, act=0):
    """(dy2 w) * act'(aux): act 1 -> aux = saved pre-activation, act 2 -> aux = saved ReLU output"""
    M, N = dy2.shape
    K = w.shape[1]
    dx = torch.empty(M, K, dtype=torch.bfloat16, device=dy2.device)
    with torch.cuda.device(dy2.device):
        call('rsc_linear_dx', dy2.data_ptr(), w.data_ptr(), _p(aux), dx.data_ptr(), M, N, K, dy2.stride(0), w.stride(0), K,
             act, _stream(), alg_bytes=2 * (M * N + N * K + M * K * (2 if act else 1)), alg_flops=2 * M * N * K)
    return dx


def gemm_dw(dy2, x2, dw32, db32):
    """dw32 (N,K fp32, row stride dw32.stride(0)) += dy2^T x2; db32 (N fp32, or None) += column sums of dy2"""
    M, N = dy2.shape
    K = x2.shape[1]
    with torch.cuda.device(dy2.device):
        call('rsc_linear_dw', dy2.data_ptr(), x2.data_ptr(), dw32.data_ptr(), _p(db32), M, N, K, dy2.stride(0), x2.stride(0),
             dw32.stride(0), _stream(), alg_bytes=2 * (M * N + M * K), alg_flops=2 * M * N * K)


def _accumulate_dw(dy2, x2, gw, gb, wdt, bdt, wshape, want_db):
    """weight / bias gradient of y = x W^T + b: into the flat fp32 views when present (returns None, None), else fresh"""
    M, N = dy2.shape
    K = x2.shape[1]
    dw = db = None
    tc = _OWN_GEMM and _tc2d(x2) and _tc2d(dy2) and M >= _TC_MIN_ROWS and _own_dw(M, N, K)
    if tc and (gw is None or (gw.dtype == torch.float32 and gw.stride(1) == 1 and gw.stride(0) % 4 == 0 and gw.data_ptr() % 16 == 0)):
        tw = gw if gw is not None else torch.zeros(N, K, dtype=torch.float32, device=dy2.device)
        tb = None
        if want_db:
            tb = gb if gb is not None else torch.zeros(N, dtype=torch.float32, device=dy2.device)
        gemm_dw(dy2, x2, tw, tb)
        if gw is None:
            dw = tw.to(wdt).view(wshape)
        if want_db and gb is None:
            db = tb.to(bdt)
        return dw, db
    if gw is not None:
        _addmm_into(gw, dy2, x2)
    else:
        dw = torch.mm(dy2.t(), x2).to(wdt).view(wshape)
    if want_db:
        if gb is not None:
            colsum(dy2, out=gb)
        else:
            db = colsum(dy2).to(bdt)
    return dw, db


def _addmm_into(gw, dy2, x2):
    if dy2.dtype == torch.float32:
        torch.addmm(gw, dy2.t(), x2, out=gw)
    else:
        torch.addmm(gw, dy2.t(), x2, out_dtype=torch.float32, out=gw)


# RSC_SMALL_BWD=1: the small layers' backward (dX, dW +=, db +=) as ONE launch of rsc_small_linear_bwd instead of the library's
# three.  A/B'd on B200 in round 2 (bench.py, same box): 81.5 vs 81.85 it/s with N, K <= 512 and 78.0 vs 82.1 without the limit
# (one CTA walks a whole contraction; the library's split-K kernels are faster than the launches they cost) -> NOT promoted.
_SMALL_BWD = __import__('os').environ.get('RSC_SMALL_BWD', '0') == '1'
_SMALL_BWD_MAX = int(__import__('os').environ.get('RSC_SMALL_BWD_MAX', 512))


def _small_bwd_ok(dy2, x2, w, gw, gb, want_db, want_dw):
    """engine mode (gradients accumulate into the flat fp32 buffer), bf16, fewer rows than the tcgen05 kernels want"""
    if not (_SMALL_BWD and want_dw and gw is not None and dy2.is_cuda and dy2.dtype == torch.bfloat16 and
            x2.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and gw.dtype == torch.float32):
        return False
    if want_db and gb is None:
        return False
    M, N = dy2.shape
    K = w.shape[1]
    # (one CTA walks a whole contraction: beyond 512 the library's split-K kernels win, measured on B200)
    return (M < _TC_MIN_ROWS and N % 8 == 0 and K % 8 == 0 and N <= _SMALL_BWD_MAX and K <= _SMALL_BWD_MAX and x2.dim() == 2 and x2.stride(1) == 1 and x2.stride(0) % 8 == 0 and
            w.dim() == 2 and w.stride(1) == 1 and w.stride(0) % 8 == 0 and gw.dim() == 2 and gw.stride(1) == 1 and
            gw.stride(0) % 2 == 0 and x2.data_ptr() % 16 == 0 and w.data_ptr() % 16 == 0 and dy2.data_ptr() % 16 == 0 and
            gw.data_ptr() % 8 == 0)


class _Linear(torch.autograd.Function):
    """y = x W^T + b.  bf16: the tcgen05 kernels of csrc/gemm_tc.cu (forward with the bias in the epilogue, dX with the
    weight read MN-major, dW + db in one kernel) where they are at least as fast as the library, see _own_plain;
    otherwise library GEMMs with db from rsc_colsum.  Runs in the dtype of x.  When the step engine has attached its
    flat buffers to the parameters (`_rsc_lp` = bf16 shadow, `_rsc_g` = fp32 gradient view) the weight is read from
    the shadow (no cast kernel) and dW / db are ACCUMULATED straight into the flat gradient buffer in fp32 --
    autograd then sees no gradient for them."""

    @staticmethod
    def forward(ctx, x, weight, bias, w, b, gw, gb):
        # weight / bias: the (master) parameters autograd tracks; w / b: what the GEMM reads.  The GEMM always sees a
        # plain 2-D row-major operand: tensors like (N, 1, C) with strides (C, N*C, 1) are "contiguous" for torch but
        # send F.linear's 3-D path to pathological cuBLAS kernels (a 512x8-tile GEMM, 30x slower, was observed)
        x2 = x.reshape(-1, x.shape[-1])
        if _tc_linear_ok(x2, w) and _own_plain(x2.shape[1]):
            b32 = None if bias is None else (bias.detach() if bias.dtype == torch.float32 else bias.detach().float())
            y = gemm_fwd(x2, w, b32, 0)[0].view(*x.shape[:-1], w.shape[0])
        else:
            y = torch.nn.functional.linear(x2, w, b).view(*x.shape[:-1], w.shape[0])
        ctx.save_for_backward(x, w)
        ctx.meta = (weight.dtype, None if bias is None else bias.dtype, gw, gb)
        ctx.wshape = weight.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        wdt, bdt, gw, gb = ctx.meta
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        x2 = x.reshape(-1, x.shape[-1])
        dx = dw = db = None
        if _small_bwd_ok(dy2, x2, w, gw, gb, bdt is not None and ctx.needs_input_grad[2], ctx.needs_input_grad[1]):
            # small layer in engine mode: dX, dW += and db += in ONE launch (rsc_small_linear_bwd)
            M, N = dy2.shape
            K = w.shape[1]
            if ctx.needs_input_grad[0]:
                dx = torch.empty(M, K, dtype=torch.bfloat16, device=dy2.device)
            with torch.cuda.device(dy2.device):
                call('rsc_small_linear_bwd', dy2.data_ptr(), x2.data_ptr(), w.data_ptr(), _p(dx), gw.data_ptr(), _p(gb), M, N, K,
                     dy2.stride(0), x2.stride(0), w.stride(0), K, gw.stride(0), _stream(),
                     alg_bytes=2 * (2 * M * N + 2 * M * K + N * K), alg_flops=4 * M * N * K)
            return (dx.view(x.shape) if dx is not None else None), None, None, None, None, None, None
        if ctx.needs_input_grad[0]:
            if _tc_linear_ok(dy2, w) and w.shape[1] % 8 == 0 and _own_plain(dy2.shape[1]):
                dx = gemm_dx(dy2, w).view(x.shape)
            else:
                dx = torch.mm(dy2, w).view(x.shape)
        want_db = bdt is not None and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1]:
            dw, db = _accumulate_dw(dy2, x2, gw, gb, wdt, bdt, ctx.wshape, want_db)
        elif want_db:
            if gb is not None:
                colsum(dy2, out=gb)
            else:
                db = colsum(dy2).to(bdt)
        return dx, dw, db, None, None, None, None


class _LinearPair(torch.autograd.Function):
    """[y1 | y2] = x [W1; W2]^T + [b1; b2] as ONE GEMM each way: two sibling Linear layers that read the same input
    (mmcv MultiScaleDeformableAttention.sampling_offsets / attention_weights, ops/multi_scale_deform_attn.py) whose
    parameters the step engine laid out back to back in its flat buffers (StepEngine._pair_linears), so the stacked
    weight / bias / gradient matrices are plain views.  weight1 .. bias2 only tie the node to the parameters;
    `pv` = dict(w, b: fp32 masters | w_lp, b_lp: bf16 shadows | gw, gb: fp32 gradient views), all stacked."""

    @staticmethod
    def forward(ctx, x, weight1, bias1, weight2, bias2, pv):
        x2 = x.reshape(-1, x.shape[-1])
        lp = x.dtype == torch.bfloat16
        w = pv['w_lp'] if lp else pv['w']
        if _tc_linear_ok(x2, w) and _own_plain(x2.shape[1]):
            y = gemm_fwd(x2, w, pv['b'], 0)[0]
        else:
            y = torch.nn.functional.linear(x2, w, pv['b_lp'] if lp else pv['b'])
        ctx.save_for_backward(x, w)
        ctx.pv = pv
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        pv = ctx.pv
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        x2 = x.reshape(-1, x.shape[-1])
        dx = None
        if ctx.needs_input_grad[0]:
            if _tc_linear_ok(dy2, w) and w.shape[1] % 8 == 0 and _own_plain(dy2.shape[1]):
                dx = gemm_dx(dy2, w).view(x.shape)
            else:
                dx = torch.mm(dy2, w).view(x.shape)
        _accumulate_dw(dy2, x2, pv['gw'], pv['gb'], None, None, None, True)
        return dx, None, None, None, None, None


def linear_pair_views(m1, m2):
    """the stacked views for _LinearPair if the step engine placed m1 / m2's parameters back to back, else None"""
    ps = (m1.weight, m2.weight, m1.bias, m2.bias)
    if any(p is None or not p.requires_grad for p in ps):
        return None

    def stacked(a, b, attr):
        ta, tb = (getattr(p, attr, None) if attr else p.data for p in (a, b))
        if ta is None or tb is None or not (ta.is_contiguous() and tb.is_contiguous()) or ta.dtype != tb.dtype:
            return None
        if tb.data_ptr() != ta.data_ptr() + ta.numel() * ta.element_size() or ta.shape[1:] != tb.shape[1:]:
            return None
        # one view over both: as_strided on the first tensor (same storage -- the engine's flat buffer)
        shape = (ta.shape[0] + tb.shape[0],) + tuple(ta.shape[1:])
        return ta.as_strided(shape, ta.stride(), ta.storage_offset())

    pv = dict(w=stacked(m1.weight, m2.weight, None), b=stacked(m1.bias, m2.bias, None),
              w_lp=stacked(m1.weight, m2.weight, '_rsc_lp'), b_lp=stacked(m1.bias, m2.bias, '_rsc_lp'),
              gw=stacked(m1.weight, m2.weight, '_rsc_g'), gb=stacked(m1.bias, m2.bias, '_rsc_g'))
    if pv['w'] is None or pv['b'] is None or pv['gw'] is None or pv['gb'] is None:
        return None
    return pv


# RSC_LINEAR_PAIR=0: sampling_offsets / attention_weights as two GEMMs again (A/B switch)
_LINEAR_PAIR = __import__('os').environ.get('RSC_LINEAR_PAIR', '1') != '0'


def linear_pair(x, m1, m2, pv):
    if torch.is_autocast_enabled('cuda') and x.is_cuda:
        x = x.to(torch.get_autocast_dtype('cuda'))
    return _LinearPair.apply(x, m1.weight, m1.bias, m2.weight, m2.bias, pv)


def linear_pair_supported(x, pv):
    if not (_LINEAR_PAIR and pv is not None and x.is_cuda and torch.is_grad_enabled()):
        return False
    dt = torch.get_autocast_dtype('cuda') if torch.is_autocast_enabled('cuda') else x.dtype
    if dt == torch.bfloat16:
        return pv['w_lp'] is not None and pv['b_lp'] is not None
    return dt == torch.float32


class _MLP(torch.autograd.Function):
    """z = act(x W1^T + b1) W2^T (+ b2): the two Linears of an FFN with the activation fused into the first GEMM's
    epilogue and its gradient into the epilogue of the second GEMM's dX (csrc/gemm_tc.cu).  Replaces
    Linear -> bias + activation kernel -> Linear of mmcv FFN (cfg :9-25 GELU, :34-50 ReLU)."""

    @staticmethod
    def forward(ctx, x, weight1, bias1, weight2, bias2, w1, w2, act, grads):
        x2 = x.reshape(-1, x.shape[-1])
        b1 = bias1.detach() if bias1.dtype == torch.float32 else bias1.detach().float()
        y, h = gemm_fwd(x2, w1, b1, act)
        b2 = None if bias2 is None else (bias2.detach() if bias2.dtype == torch.float32 else bias2.detach().float())
        if _own_plain(y.shape[1]):
            z = gemm_fwd(y, w2, b2, 0)[0]
        else:
            z = torch.nn.functional.linear(y, w2, None if b2 is None else b2.to(y.dtype))
        ctx.save_for_backward(x2, w1, w2, y, h)
        ctx.meta = (act, grads, weight1.dtype, bias1.dtype, weight2.dtype, None if bias2 is None else bias2.dtype,
                    weight1.shape, weight2.shape, x.shape)
        return z.view(*x.shape[:-1], w2.shape[0])

    @staticmethod
    def backward(ctx, dz):
        x2, w1, w2, y, h = ctx.saved_tensors
        act, (gw1, gb1, gw2, gb2), w1dt, b1dt, w2dt, b2dt, w1shape, w2shape, xshape = ctx.meta
        dz2 = dz.reshape(-1, dz.shape[-1])
        if not dz2.is_contiguous():
            dz2 = dz2.contiguous()
        # second Linear: weight / bias gradient, then its input gradient with act' in the epilogue
        dw2, db2 = _accumulate_dw(dz2, y, gw2, gb2, w2dt, b2dt, w2shape, b2dt is not None and ctx.needs_input_grad[4])
        dh = gemm_dx(dz2, w2, h if act == 1 else y, act)
        dw1, db1 = _accumulate_dw(dh, x2, gw1, gb1, w1dt, b1dt, w1shape, True)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = (gemm_dx(dh, w1) if _own_plain(dh.shape[1]) else torch.mm(dh, w1)).view(xshape)
        return dx, dw1, db1, dw2, db2, None, None, None, None


class _LinearAddLN(torch.autograd.Function):
    """(r, n) = (identity + (x W^T + bias) * scale[sample], LayerNorm(r)): ops.linear + ops.add_ln with the add / LayerNorm in
    the GEMM epilogue (rsc_linear_add_ln_fwd).  Backward = rsc_add_ln_bwd followed by the Linear's dX / dW kernels."""

    @staticmethod
    def forward(ctx, x, weight, bias, identity, scale, gamma, beta, w, eps, grads):
        ctx.set_materialize_grads(False)      # an unused output (post-norm layers drop r) costs no zero fill / read
        x2 = x.reshape(-1, x.shape[-1])
        idc = identity.contiguous()
        N = w.shape[0]
        rows = x2.shape[0]
        rps = rows // idc.shape[0]
        b32 = None if bias is None else (bias.detach() if bias.dtype == torch.float32 else bias.detach().float())
        s32, g32, be32 = _f32(scale), _f32(gamma), _f32(beta)
        r = torch.empty(idc.shape, dtype=torch.bfloat16, device=x.device)
        n = torch.empty_like(r)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        K = x2.shape[1]
        with torch.cuda.device(x.device):
            call('rsc_linear_add_ln_fwd', x2.data_ptr(), w.data_ptr(), _p(b32), idc.data_ptr(), _p(s32), g32.data_ptr(),
                 be32.data_ptr(), r.data_ptr(), n.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, N, K, x2.stride(0),
                 w.stride(0), rps, float(eps), _stream(), alg_bytes=2 * (rows * K + N * K + 3 * rows * N),
                 alg_flops=2 * rows * N * K)
        ctx.save_for_backward(x2, w, r, g32, mean, rstd, s32)
        ctx.meta = (rows, rps, N, weight.dtype, None if bias is None else bias.dtype, gamma.dtype, beta.dtype, grads,
                    weight.shape, x.shape)
        return r, n

    @staticmethod
    def backward(ctx, dr_ext, dn):
        x2, w, r, g32, mean, rstd, s32 = ctx.saved_tensors
        rows, rps, C, wdt, bdt, gdt, bedt, (gw, gbias, gg, gb), wshape, xshape = ctx.meta
        dev = r.device
        if dn is None:
            dn = torch.zeros_like(r)
        dn = dn.contiguous()
        dr_ext = None if dr_ext is None else dr_ext.contiguous()
        d_id = torch.empty_like(r)
        dy = torch.empty_like(r) if s32 is not None else None          # gradient of the Linear's output (scaled branch)
        direct = gg is not None and gb is not None
        dg = gg if direct else torch.zeros(C, dtype=torch.float32, device=dev)
        db = gb if direct else torch.zeros(C, dtype=torch.float32, device=dev)
        dbias = None
        if bdt is not None:
            dbias = gbias if gbias is not None else torch.zeros(C, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            call('rsc_add_ln_bwd', r.data_ptr(), g32.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dn.data_ptr(),
                 _p(dr_ext), _p(s32), d_id.data_ptr(), _p(dy), dg.data_ptr(), db.data_ptr(), _p(dbias), rows, rps, C,
                 _dt(r), _stream(), alg_bytes=(4 + (dr_ext is not None) + (dy is not None)) * r.numel() * r.element_size())
        dy2 = (dy if dy is not None else d_id).view(rows, C)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = (gemm_dx(dy2, w) if _own_plain(C) else torch.mm(dy2, w)).view(xshape)
        if ctx.needs_input_grad[1]:
            dw, _ = _accumulate_dw(dy2, x2, gw, None, wdt, None, wshape, False)       # (the bias gradient came out of add_ln_bwd)
        return (dx, dw, None if (bdt is None or gbias is not None) else dbias.to(bdt), d_id, None,
                None if direct else dg.to(gdt), None if direct else db.to(bedt), None, None, None)


class _MLPAddLN(torch.autograd.Function):
    """(r, n) = (identity + (act(x W1^T + b1) W2^T + b2) * scale[sample], LayerNorm(r)): a whole FFN sub-layer with its
    residual add and the following LayerNorm in three kernels (GEMM + activation epilogue, GEMM + add / LayerNorm epilogue).
    Backward: rsc_add_ln_bwd, then the kernels of _MLP.backward."""

    @staticmethod
    def forward(ctx, x, weight1, bias1, weight2, bias2, identity, scale, gamma, beta, w1, w2, act, eps, grads):
        ctx.set_materialize_grads(False)      # an unused output (post-norm layers drop r) costs no zero fill / read
        x2 = x.reshape(-1, x.shape[-1])
        idc = identity.contiguous()
        rows, C = x2.shape[0], w2.shape[0]
        rps = rows // idc.shape[0]
        b1 = bias1.detach() if bias1.dtype == torch.float32 else bias1.detach().float()
        y, h = gemm_fwd(x2, w1, b1, act)
        b2 = None if bias2 is None else (bias2.detach() if bias2.dtype == torch.float32 else bias2.detach().float())
        s32, g32, be32 = _f32(scale), _f32(gamma), _f32(beta)
        r = torch.empty(idc.shape, dtype=torch.bfloat16, device=x.device)
        n = torch.empty_like(r)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        Kh = y.shape[1]
        with torch.cuda.device(x.device):
            call('rsc_linear_add_ln_fwd', y.data_ptr(), w2.data_ptr(), _p(b2), idc.data_ptr(), _p(s32), g32.data_ptr(),
                 be32.data_ptr(), r.data_ptr(), n.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, C, Kh, y.stride(0),
                 w2.stride(0), rps, float(eps), _stream(), alg_bytes=2 * (rows * Kh + C * Kh + 3 * rows * C),
                 alg_flops=2 * rows * C * Kh)
        ctx.save_for_backward(x2, w1, w2, y, h, r, g32, mean, rstd, s32)
        ctx.meta = (act, rows, rps, C, grads, weight1.dtype, bias1.dtype, weight2.dtype, None if bias2 is None else bias2.dtype,
                    gamma.dtype, beta.dtype, weight1.shape, weight2.shape, x.shape)
        return r, n

    @staticmethod
    def backward(ctx, dr_ext, dn):
        x2, w1, w2, y, h, r, g32, mean, rstd, s32 = ctx.saved_tensors
        (act, rows, rps, C, (gw1, gb1, gw2, gb2, gg, gb), w1dt, b1dt, w2dt, b2dt, gdt, bedt, w1shape, w2shape,
         xshape) = ctx.meta
        dev = r.device
        if dn is None:
            dn = torch.zeros_like(r)
        dn = dn.contiguous()
        dr_ext = None if dr_ext is None else dr_ext.contiguous()
        d_id = torch.empty_like(r)
        dz = torch.empty_like(r) if s32 is not None else None
        direct = gg is not None and gb is not None
        dg = gg if direct else torch.zeros(C, dtype=torch.float32, device=dev)
        db = gb if direct else torch.zeros(C, dtype=torch.float32, device=dev)
        dbias2 = None
        if b2dt is not None:
            dbias2 = gb2 if gb2 is not None else torch.zeros(C, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            call('rsc_add_ln_bwd', r.data_ptr(), g32.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dn.data_ptr(),
                 _p(dr_ext), _p(s32), d_id.data_ptr(), _p(dz), dg.data_ptr(), db.data_ptr(), _p(dbias2), rows, rps, C,
                 _dt(r), _stream(), alg_bytes=(4 + (dr_ext is not None) + (dz is not None)) * r.numel() * r.element_size())
        dz2 = (dz if dz is not None else d_id).view(rows, C)
        dw2, _ = _accumulate_dw(dz2, y, gw2, None, w2dt, None, w2shape, False)
        dh = gemm_dx(dz2, w2, h if act == 1 else y, act)
        dw1, db1 = _accumulate_dw(dh, x2, gw1, gb1, w1dt, b1dt, w1shape, True)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = (gemm_dx(dh, w1) if _own_plain(dh.shape[1]) else torch.mm(dh, w1)).view(xshape)
        return (dx, dw1, db1, dw2, None if (b2dt is None or gb2 is not None) else dbias2.to(b2dt), d_id, None,
                None if direct else dg.to(gdt), None if direct else db.to(bedt), None, None, None, None, None)


def mlp_add_ln(x, weight1, bias1, weight2, bias2, identity, scale, gamma, beta, act, eps=1e-5):
    """FFN sub-layer + residual + LayerNorm -> (r, n).  Call only when mlp_supported(x, W1, b1, W2) and
    linear_add_ln_supported(x, W2, identity)."""
    if torch.is_autocast_enabled('cuda'):
        x = x.to(torch.get_autocast_dtype('cuda'))
    x = x.contiguous()
    grads = (_flat_grad(weight1), _flat_grad(bias1), _flat_grad(weight2), None if bias2 is None else _flat_grad(bias2),
             _flat_grad(gamma), _flat_grad(beta))
    if not torch.is_grad_enabled():
        grads = (None,) * 6
    w1, w2 = _compute_copy(weight1, None, torch.bfloat16), _compute_copy(weight2, None, torch.bfloat16)
    return _MLPAddLN.apply(x, weight1, bias1, weight2, bias2, identity, scale, gamma, beta, w1, w2,
                           {ACT_GELU: 1, ACT_RELU: 2, ACT_GELU_SIG: 1}[act], float(eps), grads)


def linear_add_ln_supported(x, weight, identity):
    """the fused Linear + residual + LayerNorm kernel: bf16, the normalised width fits one GEMM tile (<= 256)"""
    if not (_OWN_GEMM and x.is_cuda and weight.dim() == 2):
        return False
    dt = torch.get_autocast_dtype('cuda') if torch.is_autocast_enabled('cuda') else x.dtype
    N, K = weight.shape
    rows = x.numel() // x.shape[-1]
    # (a 256-wide tile leaves room for two pipeline stages only beside its two staging tiles: with a long contraction the
    # library GEMM + rsc_add_ln_fwd pair is faster -- measured on the encoder FFN, K = 2048)
    return (dt == torch.bfloat16 and identity.dtype == torch.bfloat16 and N % 32 == 0 and N <= 256 and K % 8 == 0 and
            (N <= 192 or K < 512) and rows >= _TC_MIN_ROWS and rows % identity.shape[0] == 0 and bool(_lib.lib().rsc_add_ln_supported(N)))


def linear_add_ln(x, weight, bias, identity, scale, gamma, beta, eps=1e-5):
    """r = identity + (x W^T + bias) * scale[b]; n = LayerNorm(r); -> (r, n).  Call only when linear_add_ln_supported(...)."""
    if torch.is_autocast_enabled('cuda'):
        x = x.to(torch.get_autocast_dtype('cuda'))
    x = x.contiguous()
    grads = (_flat_grad(weight), None if bias is None else _flat_grad(bias), _flat_grad(gamma), _flat_grad(beta))
    if not torch.is_grad_enabled():
        grads = (None, None, None, None)
    w = _compute_copy(weight, None, torch.bfloat16)
    return _LinearAddLN.apply(x, weight, bias, identity, scale, gamma, beta, w, float(eps), grads)


def mlp_supported(x, weight1, bias1, weight2):
    """the fused FFN runs on bf16 CUDA activations with feature counts that are multiples of 8"""
    if not (_OWN_GEMM and x.is_cuda and bias1 is not None and weight1.dim() == 2 and weight2.dim() == 2):
        return False
    dt = torch.get_autocast_dtype('cuda') if torch.is_autocast_enabled('cuda') else x.dtype
    rows = x.numel() // x.shape[-1]
    return dt == torch.bfloat16 and rows >= _TC_MIN_ROWS and all(d % 8 == 0 for d in (*weight1.shape, *weight2.shape))


def mlp(x, weight1, bias1, weight2, bias2, act):
    """act(x W1^T + b1) W2^T (+ b2); act = ACT_GELU (erf form) or ACT_RELU.  Call only when mlp_supported(...)."""
    if torch.is_autocast_enabled('cuda'):
        x = x.to(torch.get_autocast_dtype('cuda'))
    x = x.contiguous()
    grads = (_flat_grad(weight1), _flat_grad(bias1), _flat_grad(weight2), None if bias2 is None else _flat_grad(bias2))
    if not torch.is_grad_enabled():
        grads = (None, None, None, None)
    # (either both views of a Linear are attached or neither)
    w1, w2 = _compute_copy(weight1, None, torch.bfloat16), _compute_copy(weight2, None, torch.bfloat16)
    return _MLP.apply(x, weight1, bias1, weight2, bias2, w1, w2, {ACT_GELU: 1, ACT_RELU: 2, ACT_GELU_SIG: 1}[act], grads)


def _rows(t, rows):
    return t if t is None or rows is None else t[rows[0]:rows[1]]


def _linear_nd(x, weight, bias):
    n_out = weight.shape[0]
    if torch.is_autocast_enabled('cuda') and x.is_cuda:
        x = x.to(torch.get_autocast_dtype('cuda'))
    if not (x.is_cuda and x.dtype in (torch.float32, torch.bfloat16) and (bias is None or n_out % 4 == 0)):
        return torch.nn.functional.linear(x, weight.reshape(n_out, -1), bias)
    gw, gb = getattr(weight, '_rsc_g', None), getattr(bias, '_rsc_g', None)
    if not torch.is_grad_enabled():
        gw = gb = None
    w = _compute_copy(weight, None, x.dtype).reshape(n_out, -1)
    return _Linear.apply(x, weight, bias, w, _compute_copy(bias, None, x.dtype), None if gw is None else gw.view(n_out, -1), gb)


def _compute_copy(p, rows, dtype):
    """what the GEMM reads for parameter p: the engine's bf16 shadow if there is one, else a cast."""
    if p is None:
        return None
    lp = getattr(p, '_rsc_lp', None)
    t = _rows(lp if lp is not None and lp.dtype == dtype else p.detach(), rows)
    return t if t.dtype == dtype else t.to(dtype)


def linear(x, weight, bias=None, rows=None):
    """F.linear in the active compute dtype; on CUDA (out_features % 4 == 0) the bias gradient comes
    from rsc_colsum instead of a separate ATen reduction.  The GEMMs are library calls either way.
    `rows=(r0, r1)` applies only output rows r0:r1 of weight / bias (the q / k / v thirds of a packed
    in_proj) without materialising slices of the master weights."""
    n_out = weight.shape[0] if rows is None else rows[1] - rows[0]
    if weight.dim() != 2:      # a Conv2d kernel used as the (out, in*kh*kw) matrix of a patch-embedding GEMM
        assert rows is None
        return _linear_nd(x, weight, bias)
    if x.is_cuda and (bias is None or n_out % 4 == 0):
        if torch.is_autocast_enabled('cuda'):
            x = x.to(torch.get_autocast_dtype('cuda'))
        if x.dtype in (torch.float32, torch.bfloat16):
            gw, gb = getattr(weight, '_rsc_g', None), getattr(bias, '_rsc_g', None)
            if not torch.is_grad_enabled():
                gw = gb = None
            if rows is not None and (gw is None or (gb is None and bias is not None)):   # no engine: autograd slices
                weight, bias, rows, gw, gb = _rows(weight, rows), _rows(bias, rows), None, None, None
            return _Linear.apply(x, weight, bias, _compute_copy(weight, rows, x.dtype),
                                 _compute_copy(bias, rows, x.dtype), _rows(gw, rows), _rows(gb, rows))
    return torch.nn.functional.linear(x, _rows(weight, rows), _rows(bias, rows))


# ---------------------------------------------------------------------------
# convolutions as im2col GEMMs, PPM pooling          
# ---------------------------------------------------------------------------
class _Im2Col(torch.autograd.Function):
    """channels-last (B,H,W,C) -> (B*Ho*Wo, kh*kw*C), column = (tap, c); backward = the adjoint gather"""

    @staticmethod
    def forward(ctx, x, kh, kw, stride, pad):
        _cuda(x)
        B, H, W, C = x.shape
        Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
        col = torch.empty(B * Ho * Wo, kh * kw * C, dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            call('rsc_im2col_fwd', x.data_ptr(), col.data_ptr(), B, H, W, C, kh, kw, stride, pad, _dt(x), _stream(),
                 alg_bytes=(x.numel() + col.numel()) * x.element_size())
        ctx.meta = (B, H, W, C, kh, kw, stride, pad)
        return col

    @staticmethod
    def backward(ctx, dcol):
        B, H, W, C, kh, kw, stride, pad = ctx.meta
        dcol = dcol.contiguous()
        dx = torch.empty(B, H, W, C, dtype=dcol.dtype, device=dcol.device)
        with torch.cuda.device(dcol.device):
            call('rsc_im2col_bwd', dcol.data_ptr(), dx.data_ptr(), B, H, W, C, kh, kw, stride, pad, _dt(dcol), _stream(),
                 alg_bytes=(dx.numel() + dcol.numel()) * dcol.element_size())
        return dx, None, None, None, None


def conv2d_supported(x, weight, stride=1, padding=0, dilation=1, groups=1):
    """nn.Conv2d cases that run as (gather +) GEMM: CUDA, fp32 / bf16 compute, square stride / padding, no dilation / groups,
    input channels a multiple of 8"""
    if not (x.is_cuda and x.dim() == 4 and weight.dim() == 4 and groups == 1):
        return False
    one = lambda v: v if isinstance(v, int) else (v[0] if len(set(v)) == 1 else None)
    if one(stride) is None or one(padding) is None or one(dilation) != 1:
        return False
    dt = torch.get_autocast_dtype('cuda') if torch.is_autocast_enabled('cuda') else x.dtype
    return dt in (torch.float32, torch.bfloat16) and x.shape[1] % 8 == 0 and weight.shape[0] % 8 == 0


def conv2d(x, weight, bias=None, stride=1, padding=0):
    """F.conv2d(x, weight, bias, stride, padding) on a (B,C,H,W) tensor (any strides; channels-last is free) as an im2col
    GEMM: the k x k gather is rsc_im2col_fwd (none for 1x1 / stride 1), the contraction ops.linear (the tcgen05 kernels in
    bf16).  Returns a (B,Cout,Ho,Wo) tensor with channels-last strides."""
    one = lambda v: v if isinstance(v, int) else v[0]
    stride, padding = one(stride), one(padding)
    Cout, Cin, kh, kw = weight.shape
    if torch.is_autocast_enabled('cuda'):
        x = x.to(torch.get_autocast_dtype('cuda'))
    B, _, H, W = x.shape
    xl = x.permute(0, 2, 3, 1)                    # (B,H,W,C): a view for channels-last inputs
    if not xl.is_contiguous():
        xl = xl.contiguous()
    Ho, Wo = (H + 2 * padding - kh) // stride + 1, (W + 2 * padding - kw) // stride + 1
    if kh == 1 and kw == 1 and stride == 1 and padding == 0:
        # the (Cout, Cin, 1, 1) parameter IS the GEMM's weight matrix: the engine's bf16 shadow / flat gradient views apply
        y = _linear_nd(xl.reshape(B * H * W, Cin), weight, bias)
    else:
        a = _Im2Col.apply(xl, kh, kw, stride, padding)
        w2 = weight.permute(0, 2, 3, 1).reshape(Cout, kh * kw * Cin)      # (Cout, kh, kw, Cin): matches the gather's column order
        y = linear(a, w2, bias)
    return y.view(B, Ho, Wo, Cout).permute(0, 3, 1, 2)


class _AdaptiveAvgPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, S):
        _cuda(x)
        B, H, W, C = x.shape
        y = torch.empty(B, S, S, C, dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            call('rsc_adaptive_avgpool_fwd', x.data_ptr(), y.data_ptr(), B, H, W, C, S, _dt(x), _stream(),
                 alg_bytes=(x.numel() + y.numel()) * x.element_size())
        ctx.meta = (B, H, W, C, S)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, H, W, C, S = ctx.meta
        dy = dy.contiguous()
        dx = torch.empty(B, H, W, C, dtype=dy.dtype, device=dy.device)
        with torch.cuda.device(dy.device):
            call('rsc_adaptive_avgpool_bwd', dy.data_ptr(), dx.data_ptr(), B, H, W, C, S, _dt(dy), _stream(),
                 alg_bytes=(dx.numel() + dy.numel()) * dy.element_size())
        return dx, None


def adaptive_avg_pool2d(x, size):
    """nn.AdaptiveAvgPool2d(size) on a (B,C,H,W) CUDA tensor (fp32 / bf16, C % 8 == 0) -> (B,C,size,size), channels-last strides"""
    xl = x.permute(0, 2, 3, 1)
    if not xl.is_contiguous():
        xl = xl.contiguous()
    return _AdaptiveAvgPool.apply(xl, int(size)).permute(0, 3, 1, 2)


# ---------------------------------------------------------------------------
# decoder attention core + bit masks              (8f rank 2)
# ---------------------------------------------------------------------------
class MaskBits:
    """An attention mask as bit words (rsc_attn_*): `bits` int32 (Bm, Lq, ceil(Lk/32)), Bm = 1 (shared by the batch) or B;
    bit set = that key is masked for that query, the same for every head."""

    def __init__(self, bits, Lk):
        self.bits, self.Lk = bits, Lk


def pack_mask_bits(mask):
    """boolean mask (Lq, Lk) or (B, Lq, Lk), True = masked (the torch MHA convention) -> MaskBits"""
    _cuda(mask)
    assert mask.dtype == torch.bool and mask.dim() in (2, 3)
    m = mask.contiguous().view(torch.uint8)
    Lk = m.shape[-1]
    bits = torch.empty(m.shape[:-1] + ((Lk + 31) // 32,), dtype=torch.int32, device=m.device)
    with torch.cuda.device(m.device):
        call('rsc_pack_mask_bits', m.data_ptr(), bits.data_ptr(), m.numel() // Lk, Lk, _stream())
    return MaskBits(bits.view((1,) * (3 - bits.dim()) + bits.shape), Lk)


def m2f_attn_mask(mask_pred, size):
    """The Mask2Former cross-attention mask of the next decoder layer from mask_pred (B, Q, Hi, Wi) logits: bilinear
    resize to the key grid `size`, sigmoid < 0.5 = masked, rows with every key masked un-masked
    (mask2former_head.py:134-139 + :177-178) -> MaskBits (B, Q, ceil(h*w/32)); one kernel, nothing else is materialised."""
    _cuda(mask_pred)
    x = mask_pred.detach().contiguous()
    B, Q, Hi, Wi = x.shape
    h, w = int(size[0]), int(size[1])
    bits = torch.empty(B, Q, (h * w + 31) // 32, dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        call('rsc_m2f_mask_bits', x.data_ptr(), bits.data_ptr(), B * Q, Hi, Wi, h, w, _dt(x), _stream(),
             alg_bytes=x.numel() * x.element_size())
    return MaskBits(bits, h * w)


def attention_supported(x, embed_dims, heads):
    return x.is_cuda and x.dtype == torch.bfloat16 and embed_dims == heads * 32


def _sl_sb(t):
    """(sequence, batch) element strides of a (L, B, E) view whose last dimension is contiguous"""
    assert t.stride(2) == 1
    return t.stride(0), t.stride(1)


class _Attention(torch.autograd.Function):
    """q / k / v: (L, B, E) bf16 views (last dim contiguous) of `srcs` = the projection outputs they were sliced from;
    layout[i] = (index into srcs, column offset).  Gradients are written straight into tensors shaped like the
    sources, so a packed q|k projection gets ONE (L, B, 2E) gradient without slice-backward zero fills."""

    @staticmethod
    def forward(ctx, heads, mask, layout, *srcs):
        E = heads * 32
        q, k, v = (srcs[i][..., c:c + E] for i, c in layout)
        Lq, B, _ = q.shape
        Lk = k.shape[0]
        out = torch.empty(Lq, B, E, dtype=q.dtype, device=q.device)
        lse = torch.empty(B * heads, Lq, dtype=torch.float32, device=q.device)
        scale = 32 ** -0.5
        with torch.cuda.device(q.device):
            ns = _lib.lib().rsc_attn_nsplit(B, heads, Lq, Lk)
            ws = torch.empty(ns * B * heads * Lq * 34, dtype=torch.float32, device=q.device) if ns > 1 else None
            mb = mask.bits if mask is not None else None
            if mb is not None:
                assert mask.Lk == Lk and mb.shape[1] == Lq and mb.shape[0] in (1, B), 'mask bits do not match the attention shape'
            call('rsc_attn_fwd', q.data_ptr(), k.data_ptr(), v.data_ptr(), _p(mb), out.data_ptr(), lse.data_ptr(), _p(ws),
                 B, heads, Lq, Lk, 32, *_sl_sb(q), *_sl_sb(k), *_sl_sb(v), *_sl_sb(out),
                 0 if mb is None or mb.shape[0] == 1 else mb.stride(0), ns, scale, _stream(),
                 alg_bytes=(2 * q.numel() + 2 * k.numel()) * 2, alg_flops=4 * B * heads * Lq * Lk * 32)
        ctx.save_for_backward(out, lse, *srcs)
        ctx.meta = (heads, mask, layout, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        heads, mask, layout, scale = ctx.meta
        out, lse, *srcs = ctx.saved_tensors
        E = heads * 32
        q, k, v = (srcs[i][..., c:c + E] for i, c in layout)
        Lq, B, _ = q.shape
        Lk = k.shape[0]
        dout = dout.contiguous()
        grads = [torch.empty_like(s) for s in srcs]
        dk, dv = (grads[i][..., c:c + E] for i, c in layout[1:])
        dq32 = torch.empty(Lq, B, E, dtype=torch.float32, device=q.device)
        delta = torch.empty(B * heads * Lq, dtype=torch.float32, device=q.device)
        mb = mask.bits if mask is not None else None
        with torch.cuda.device(q.device):
            call('rsc_attn_bwd', q.data_ptr(), k.data_ptr(), v.data_ptr(), _p(mb), out.data_ptr(), dout.data_ptr(),
                 lse.data_ptr(), delta.data_ptr(), dq32.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, heads, Lq, Lk, 32,
                 *_sl_sb(q), *_sl_sb(k), *_sl_sb(v), *_sl_sb(out), *_sl_sb(dk), *_sl_sb(dv),
                 0 if mb is None or mb.shape[0] == 1 else mb.stride(0), scale, _stream(),
                 alg_bytes=(3 * q.numel() + 4 * k.numel()) * 2, alg_flops=10 * B * heads * Lq * Lk * 32)
        i, c = layout[0]
        grads[i][..., c:c + E].copy_(dq32)
        return (None, None, None) + tuple(grads)


def attention(q, k, v, heads, mask=None, packed_qk=None):
    """softmax(q k^T / sqrt(32) + mask) v per head.  q (Lq,B,E), k / v (Lk,B,E) bf16 CUDA tensors, E = heads*32.
    packed_qk (L,B,2E): the output of ONE q|k projection (self-attention); q and k are then its column halves and its
    gradient is produced in one piece (no slice-backward zero fills).  mask: MaskBits or None.  -> (Lq,B,E)."""
    E = heads * 32

    def ok(t):
        return t if t.stride(2) == 1 and t.stride(0) % 8 == 0 and t.stride(1) % 8 == 0 else t.contiguous()
    if packed_qk is not None:
        _cuda(packed_qk, v)
        return _Attention.apply(heads, mask, ((0, 0), (0, E), (1, 0)), ok(packed_qk), ok(v))
    _cuda(q, k, v)
    return _Attention.apply(heads, mask, ((0, 0), (1, 0), (2, 0)), ok(q), ok(k), ok(v))


# ---------------------------------------------------------------------------
# GroupNorm / BatchNorm (+ReLU) on channels-last maps   
# ---------------------------------------------------------------------------
def norm_supported(x, num_channels):
    return (x.is_cuda and x.dim() == 4 and x.dtype in (torch.float32, torch.bfloat16) and
            bool(_lib.lib().rsc_norm_supported(int(num_channels), _dt(x))))


class _GroupNorm(torch.autograd.Function):
    """x (R, P, C) channels-last; groups of Cg channels per row (see include/rscotr.h: GroupNorm and training-mode
    BatchNorm are the same reduction).  `running` = (running_mean, running_var, momentum) for BatchNorm."""

    @staticmethod
    def forward(ctx, x, gamma, beta, Cg, eps, relu, running, fg):
        R, P, C = x.shape
        g32, b32 = _f32(gamma), _f32(beta)
        y = torch.empty_like(x)
        stats = torch.empty(R, C // Cg, 2, dtype=torch.float32, device=x.device)
        ws = torch.empty(R * C * 2, dtype=torch.float32, device=x.device)
        rm, rv, mom = running if running is not None else (None, None, 0.0)
        with torch.cuda.device(x.device):
            call('rsc_groupnorm_fwd', x.data_ptr(), g32.data_ptr(), b32.data_ptr(), y.data_ptr(), stats.data_ptr(), ws.data_ptr(),
                 R, P, C, Cg, float(eps), int(relu), _p(rm), _p(rv), float(mom), _dt(x), _stream(),
                 alg_bytes=3 * x.numel() * x.element_size())
        ctx.save_for_backward(x, g32, b32, stats)
        ctx.meta = (Cg, relu, gamma.dtype, beta.dtype, fg[0], fg[1])
        return y

    @staticmethod
    def backward(ctx, dy):
        x, g32, b32, stats = ctx.saved_tensors
        Cg, relu, gdt, bdt, gg, gb = ctx.meta
        R, P, C = x.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        direct = gg is not None and gb is not None
        dg = gg if direct else torch.zeros(C, dtype=torch.float32, device=x.device)
        db = gb if direct else torch.zeros_like(dg)
        ws = torch.empty(R * C * 2 + R * (C // Cg) * 2, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            call('rsc_groupnorm_bwd', x.data_ptr(), dy.data_ptr(), g32.data_ptr(), b32.data_ptr(), stats.data_ptr(), dx.data_ptr(),
                 dg.data_ptr(), db.data_ptr(), ws.data_ptr(), R, P, C, Cg, int(relu), _dt(x), _stream(),
                 alg_bytes=5 * x.numel() * x.element_size())
        if direct:
            return dx, None, None, None, None, None, None, None
        return dx, dg.to(gdt), db.to(bdt), None, None, None, None, None


def _channels_last_rows(x):
    """(B,C,H,W) -> contiguous (B, H*W, C) (a view when x already has channels-last strides)"""
    B, C, H, W = x.shape
    xl = x.permute(0, 2, 3, 1)
    if not xl.is_contiguous():
        xl = xl.contiguous()
    return xl.view(B, H * W, C)


def group_norm(x, num_groups, gamma, beta, eps=1e-5, relu=False):
    """nn.GroupNorm(num_groups) [+ ReLU] on (B,C,H,W) -> (B,C,H,W) with channels-last strides"""
    _cuda(x, gamma, beta)
    B, C, H, W = x.shape
    y = _GroupNorm.apply(_channels_last_rows(x), gamma, beta, C // int(num_groups), eps, bool(relu), None,
                         (_flat_grad(gamma), _flat_grad(beta)))
    return y.view(B, H, W, C).permute(0, 3, 1, 2)


def batch_norm_train(x, gamma, beta, running_mean, running_var, momentum=0.1, eps=1e-5, relu=False):
    """nn.BatchNorm2d in training mode (batch statistics, running statistics updated in place) [+ ReLU]"""
    _cuda(x, gamma, beta)
    B, C, H, W = x.shape
    running = None if running_mean is None else (running_mean, running_var, momentum)
    y = _GroupNorm.apply(_channels_last_rows(x).view(1, B * H * W, C), gamma, beta, 1, eps, bool(relu), running,
                         (_flat_grad(gamma), _flat_grad(beta)))
    return y.view(B, H, W, C).permute(0, 3, 1, 2)


def normalize_u8(img, mean, inv_std, valid_hw=None, to_rgb=True, out_dtype=torch.float32):
    """uint8 (B,C,H,W) batch -> (img[channels flipped if to_rgb] - mean) * inv_std, zero outside valid_hw (B,2) int32"""
    _cuda(img, mean, inv_std, valid_hw)
    img = img.contiguous()
    B, C, H, W = img.shape
    out = torch.empty(B, C, H, W, dtype=out_dtype, device=img.device)
    with torch.cuda.device(img.device):
        call('rsc_normalize_u8', img.data_ptr(), out.data_ptr(), mean.data_ptr(), inv_std.data_ptr(), _p(valid_hw), B, C, H, W,
             int(to_rgb), _dt(out), _stream(), alg_bytes=img.numel() + out.numel() * out.element_size())
    return out


# ---------------------------------------------------------------------------
# iterative box refinement of the DINO decoder / head   
# ---------------------------------------------------------------------------
class _BoxRefine(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tmp, ref, eps):
        _cuda(tmp, ref)
        tmp = tmp.contiguous()
        ref32 = ref.detach().float().contiguous()
        out = torch.empty(tmp.shape, dtype=torch.float32, device=tmp.device)
        with torch.cuda.device(tmp.device):
            call('rsc_box_refine_fwd', tmp.data_ptr(), ref32.data_ptr(), out.data_ptr(), tmp.numel(), float(eps), _dt(tmp), _stream())
        ctx.save_for_backward(out, ref32)
        ctx.meta = (eps, tmp.dtype, ref.dtype, ctx.needs_input_grad[1])
        return out

    @staticmethod
    def backward(ctx, dout):
        out, ref32 = ctx.saved_tensors
        eps, tdt, rdt, need_ref = ctx.meta
        dout = dout.float().contiguous()
        dtmp = torch.empty(out.shape, dtype=tdt, device=out.device)
        dref = torch.empty_like(out) if need_ref else None
        with torch.cuda.device(out.device):
            call('rsc_box_refine_bwd', out.data_ptr(), ref32.data_ptr(), dout.data_ptr(), dtmp.data_ptr(), _p(dref), out.numel(),
                 float(eps), _dt(dtmp), _stream())
        return dtmp, (dref.to(rdt) if need_ref else None), None


def box_refine(tmp, ref, eps=1e-3):
    """sigmoid(tmp.float() + inverse_sigmoid(ref, eps)) -> fp32, one kernel each way (tmp fp32 / bf16, same shape as ref)"""
    assert tmp.shape == ref.shape
    return _BoxRefine.apply(tmp, ref, eps)


KernelTimer = _lib.KernelTimer


def launch_count():
    return _lib.launch_count()


def reset_launch_count():
    _lib.reset_launch_count()
