"""CPU-side checks of the C-ABI boundary: the shared library builds / loads and
exports every symbol include/rscotr.h declares; host-side argument validation
does not need a GPU."""
import ctypes
import os

import pytest

from rscotr_b200 import _lib, build


@pytest.fixture(scope='module')
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.lib()


def test_header_symbols_exported(lib):
    names = _lib.declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), 'symbol %s declared in include/rscotr.h but not exported' % n


def test_bindings_cover_header():
    declared = set(_lib.declared_symbols())
    bound = set(_lib._SIGS) | {'rsc_last_error', 'rsc_version', 'rsc_launch_count', 'rsc_reset_launch_count'}
    assert declared == bound


def test_version_and_counter(lib):
    assert lib.rsc_version() >= 100
    lib.rsc_reset_launch_count()
    assert lib.rsc_launch_count() == 0


def test_invalid_args_rejected_without_gpu(lib):
    # argument validation happens before any CUDA call
    st = lib.rsc_wmsa_fwd(None, None, None, None, 1, 7, 7, 96, 3, 7, 0, ctypes.c_float(1.0), 0, None)
    assert st == 1 and b'null pointer' in lib.rsc_last_error()
    st = lib.rsc_wmsa_fwd(None, None, None, None, 1, 7, 7, 100, 3, 7, 0, ctypes.c_float(1.0), 0, None)
    assert st == 1 and b'head_dim' in lib.rsc_last_error()
    st = lib.rsc_wmsa_fwd(None, None, None, None, 1, 7, 7, 96, 3, 8, 0, ctypes.c_float(1.0), 0, None)
    assert st == 1 and b'window_size' in lib.rsc_last_error()
    st = lib.rsc_msda_fwd(None, None, None, None, None, None, 1, 10, 10, 6, 4, 4, 64, 0, None)
    assert st == 1 and b'multiple of 4' in lib.rsc_last_error()
    st = lib.rsc_window_index_partition(None, 0, 7, 7, 7, 0, None)
    assert st == 1 and b'empty' in lib.rsc_last_error()


def test_ops_refuse_cpu_tensors():
    import torch
    from rscotr_b200 import ops
    with pytest.raises(RuntimeError, match='CUDA tensors only'):
        ops.wmsa(torch.zeros(1, 49, 288), None, torch.zeros(169, 3), (7, 7), 3)
    with pytest.raises(RuntimeError, match='CUDA tensors only'):
        ops.global_avg_pool(torch.zeros(1, 4, 2, 2))


def test_fused_entry_points_validate_arguments_without_gpu(lib):
    """the fused element-wise / loss entry points: argument errors come back as status 1 + message, before any CUDA call."""
    F32, BF16 = 0, 1
    st = lib.rsc_bias_act_fwd(None, None, None, 10, 100, 0, BF16, None)            # C % 8
    assert st == 1 and b'C % 8' in lib.rsc_last_error()
    st = lib.rsc_bias_act_fwd(None, None, None, 10, 96, 3, BF16, None)             # unknown activation
    assert st == 1 and b'act must be' in lib.rsc_last_error()
    st = lib.rsc_bias_act_fwd(None, None, None, 10, 96, 2, F32, None)              # the logistic-fit GELU is bf16 only
    assert st == 1 and b'bf16 only' in lib.rsc_last_error()
    st = lib.rsc_bias_act_fwd(None, None, None, 10, 96, 2, BF16, None)             # accepted combination: next check is the pointers
    assert st == 1 and b'null pointer' in lib.rsc_last_error()
    assert lib.rsc_add_ln_supported(96) == 1 and lib.rsc_add_ln_supported(100) == 0
