"""The kernels added in the second half of round 2, once each at small shapes, for compute-sanitizer:
    compute-sanitizer --tool initcheck python tools/sanitize_new_kernels.py     (uninitialised global reads)
    compute-sanitizer --tool racecheck python tools/sanitize_new_kernels.py     (shared-memory hazards)
    compute-sanitizer --tool memcheck  python tools/sanitize_new_kernels.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rscotr_b200 import _lib, ops  # noqa: E402


def main():
    dev = 'cuda'
    _lib.call('rsc_set_patch_merge_variant', 2)        # the round-2 PatchMerging kernels (opt-in)
    g = torch.Generator().manual_seed(0)
    for dtype in (torch.float32, torch.bfloat16):
        for B, H, W, C in [(2, 16, 16, 96), (2, 8, 8, 192), (2, 4, 4, 384), (1, 9, 7, 96), (1, 4, 4, 512), (3, 37, 41, 96)]:
            x = torch.randn(B, H * W, C, generator=g).to(dtype).to(dev).requires_grad_(True)
            ga = torch.randn(4 * C, generator=g).to(dev).requires_grad_(True)
            be = torch.randn(4 * C, generator=g).to(dev).requires_grad_(True)
            y = ops.patch_merge_ln(x, (H, W), ga, be)
            y.backward(torch.randn(y.shape, generator=g).to(dtype).to(dev))
            torch.cuda.synchronize()
            assert torch.isfinite(y).all() and torch.isfinite(x.grad).all()
        # fused ms_deform_attn with row strides
        shapes = [(6, 5), (3, 3), (2, 2), (1, 1)]
        Nv = sum(h * w for h, w in shapes)
        value = torch.randn(2, Nv, 8, 32, generator=g).to(dtype).to(dev).requires_grad_(True)
        both = torch.randn(2, 19, 384, generator=g).to(dtype).to(dev).requires_grad_(True)
        ref = torch.rand(2, 19, 4, 2, generator=g).to(dev)
        ss = torch.tensor(shapes).to(dev)
        st = torch.tensor([0, 30, 39, 43]).to(dev)
        out = ops.ms_deform_attn_fused_packed(value, ss, st, both, ref, 4, 4)
        out.backward(torch.randn(out.shape, generator=g).to(dtype).to(dev))
        torch.cuda.synchronize()
        assert torch.isfinite(both.grad).all() and torch.isfinite(value.grad).all()
    print('sanitize script done')


if __name__ == '__main__':
    main()
