"""ctypes binding of librscotr_b200.so (the C ABI declared in include/rscotr.h).

No CPU fallback: if the shared library is missing the import of any op raises.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'librscotr_b200.so')
HEADER = os.path.join(os.path.dirname(HERE), 'include', 'rscotr.h')

RSC_F32, RSC_BF16 = 0, 1

_lib = None

_P = ctypes.c_void_p
_I = ctypes.c_int
_F = ctypes.c_float

_SIGS = {
    'rsc_window_index_partition': [_P, _I, _I, _I, _I, _I, _P],
    'rsc_window_index_reverse': [_P, _I, _I, _I, _I, _I, _P],
    'rsc_window_partition': [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    'rsc_window_reverse': [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    'rsc_wmsa_fwd': [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    'rsc_wmsa_bwd': [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    'rsc_patch_merge_ln_fwd': [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P],
    'rsc_patch_merge_ln_bwd': [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    'rsc_layernorm_fwd': [_P, _P, _P, _P, _P, _P, ctypes.c_int64, _I, _F, _I, _I, _P],
    'rsc_layernorm_bwd': [_P, _P, _P, _P, _P, _P, _P, _P, ctypes.c_int64, _I, _I, _I, _P],
    'rsc_adamw_step': [_P, _P, _P, _P, ctypes.c_int64, _P, _F, _F, _F, _F, _F, _P, _P, _P, _P],
    'rsc_det_match': [_P] * 6 + [_I] * 7 + [_F] * 6 + [_P, _P, _P, _I, _P],
    'rsc_det_loss_fwd': [_P] * 9 + [_I] * 9 + [_F] * 6 + [_I, _P],
    'rsc_det_loss_bwd': [_P] * 11 + [_I] * 9 + [_F] * 6 + [_I, _P],
    'rsc_upsample_ce_fwd': [_P] * 4 + [_I] * 8 + [_P],
    'rsc_upsample_ce_bwd': [_P] * 6 + [_I] * 7 + [_P],
    'rsc_add_ln_supported': [_I],
    'rsc_add_ln_fwd': [_P] * 10 + [ctypes.c_int64, ctypes.c_int64, _I, _F, _I, _P],
    'rsc_add_ln_bwd': [_P] * 12 + [ctypes.c_int64, ctypes.c_int64, _I, _I, _P],
    'rsc_bias_act_fwd': [_P, _P, _P, ctypes.c_int64, _I, _I, _I, _P],
    'rsc_bias_act_bwd': [_P, _P, _P, _P, _P, ctypes.c_int64, _I, _I, _I, _P],
    'rsc_patchify4': [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    'rsc_colsum': [_P, _P, ctypes.c_int64, _I, _I, _P],
    'rsc_msda_fwd': [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    'rsc_msda_bwd': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    'rsc_msda_fused_fwd': [_P] * 7 + [_I] * 11 + [_P],
    'rsc_msda_fused_bwd': [_P] * 10 + [_I] * 11 + [_P],
    'rsc_gap_fwd': [_P, _P, _I, _I, _I, _I, _I, _P],
    'rsc_gap_bwd': [_P, _P, _I, _I, _I, _I, _I, _P],
    'rsc_bilinear_fwd': [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    'rsc_bilinear_bwd': [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    'rsc_sigmoid_focal_loss_fwd': [_P, _P, _P, _I, _I, _F, _F, _I, _P],
    'rsc_sigmoid_focal_loss_bwd': [_P, _P, _P, _I, _I, _F, _F, _I, _P],
    'rsc_im2col_fwd': [_P, _P] + [_I] * 9 + [_P],
    'rsc_im2col_bwd': [_P, _P] + [_I] * 9 + [_P],
    'rsc_adaptive_avgpool_fwd': [_P, _P] + [_I] * 6 + [_P],
    'rsc_adaptive_avgpool_bwd': [_P, _P] + [_I] * 6 + [_P],
    'rsc_linear_fwd': [_P, _P, _P, _P, _P, ctypes.c_int64, _I, _I, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _I, _P],
    'rsc_linear_add_ln_fwd': [_P] * 11 + [ctypes.c_int64, _I, _I, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _F, _P],
    'rsc_linear_dx': [_P, _P, _P, _P, ctypes.c_int64, _I, _I, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _I, _P],
    'rsc_linear_dw': [_P, _P, _P, _P, ctypes.c_int64, _I, _I, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _P],
    'rsc_attn_nsplit': [_I, _I, _I, _I],
    'rsc_attn_fwd': [_P] * 7 + [_I] * 5 + [ctypes.c_int64] * 9 + [_I, _F, _P],
    'rsc_attn_bwd': [_P] * 11 + [_I] * 5 + [ctypes.c_int64] * 13 + [_F, _P],
    'rsc_m2f_mask_bits': [_P, _P] + [_I] * 6 + [_P],
    'rsc_pack_mask_bits': [_P, _P, ctypes.c_int64, _I, _P],
    'rsc_norm_supported': [_I, _I],
    'rsc_groupnorm_fwd': [_P] * 6 + [_I] * 4 + [_F, _I, _P, _P, _F, _I, _P],
    'rsc_groupnorm_bwd': [_P] * 9 + [_I] * 6 + [_P],
    'rsc_normalize_u8': [_P] * 5 + [_I] * 6 + [_P],
    'rsc_bilinear_cl_fwd': [_P, _P] + [_I] * 7 + [_P],
    'rsc_bilinear_cl_bwd': [_P, _P] + [_I] * 7 + [_P],
    'rsc_nvls_allreduce_mean': [_P, ctypes.c_int64, ctypes.c_int64, _I, _I, _F, _I, _P],
    'rsc_box_refine_fwd': [_P, _P, _P, ctypes.c_int64, _F, _I, _P],
    'rsc_box_refine_bwd': [_P, _P, _P, _P, _P, ctypes.c_int64, _F, _I, _P],
    'rsc_small_linear_bwd': [_P] * 6 + [_I] * 3 + [ctypes.c_int64] * 5 + [_P],
    'rsc_set_gemm_sms': [_I, ctypes.c_int64],
    'rsc_set_patch_merge_variant': [_I],
}


def declared_symbols():
    """Every function name declared in include/rscotr.h."""
    txt = open(HEADER).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(rsc_[a-z0-9_]+)\s*\(', txt)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'librscotr_b200.so not built (%s); run `python -c "import __graft_entry__ as g; g.build()"`. '
                'There is no CPU fallback.' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _I
        L.rsc_last_error.restype = ctypes.c_char_p
        L.rsc_version.restype = _I
        L.rsc_launch_count.restype = ctypes.c_int64
        L.rsc_reset_launch_count.restype = None
        _lib = L
    return _lib


def check(status, name):
    if status != 0:
        raise RuntimeError('%s failed (%d): %s' % (name, status, lib().rsc_last_error().decode()))


_timer = None      # set by KernelTimer: collects (name, alg_bytes, start_event, end_event)


def call(name, *args, alg_bytes=0, alg_flops=0):
    if _timer is None:
        check(getattr(lib(), name)(*args), name)
        return
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(getattr(lib(), name)(*args), name)
    e1.record()
    _timer.append((name, alg_bytes, e0, e1, alg_flops))


class KernelTimer:
    """with KernelTimer() as kt: ...   -> kt.summary(): per C-ABI entry point the number of
    launches, summed device time (CUDA events on the launching stream) and summed
    algorithmic bytes.  Used by bench.py for the roofline object."""

    def __enter__(self):
        global _timer
        self.records = []
        _timer = self.records
        return self

    def __exit__(self, *exc):
        global _timer
        _timer = None

    def summary(self, peak_gbs=None, peak_tflops=None):
        """peak_gbs / peak_tflops: when given, `big_roof_ms` sums per launch the slower of (bytes / HBM peak) and
        (flops / tensor peak) -- the time a launch would take AT its own binding limit."""
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, nbytes, e0, e1, nflops in self.records:
            d = out.setdefault(name, dict(launches=0, ms=0.0, bytes=0, big_launches=0, big_ms=0.0, big_bytes=0, big_flops=0,
                                          big_roof_ms=0.0))
            t = e0.elapsed_time(e1)
            d['launches'] += 1
            d['ms'] += t
            d['bytes'] += nbytes
            if nbytes >= self.BIG:        # launches large enough for bandwidth to be the question
                d['big_launches'] += 1
                d['big_ms'] += t
                d['big_bytes'] += nbytes
                d['big_flops'] += nflops
                if peak_gbs:
                    d['big_roof_ms'] += 1e3 * max(nbytes / (peak_gbs * 1e9), nflops / (peak_tflops * 1e12) if peak_tflops else 0.0)
        for d in out.values():
            d['gbs'] = d['bytes'] / d['ms'] / 1e6 if d['ms'] > 0 else 0.0
            d['big_gbs'] = d['big_bytes'] / d['big_ms'] / 1e6 if d['big_ms'] > 0 else 0.0
        return out

    BIG = 32 << 20


def launch_count():
    return int(lib().rsc_launch_count())


def reset_launch_count():
    lib().rsc_reset_launch_count()
