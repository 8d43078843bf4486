"""Is rsc_patch_merge_ln_{fwd,bwd} (variant 2) a pure function of its inputs?  Same inputs, 60 calls per shape with poisoned
(NaN-filled, then freed) allocations in between so that recycled memory is never clean; y / mean / rstd / dx compared BIT FOR BIT
with the first call, d(gamma) / d(beta) (atomic sums) to 1e-5, and everything with variant 1.   python tools/pm_determinism.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rscotr_b200 import _lib, ops  # noqa: E402


def once(x, ga, be, dy, H, W):
    xx = x.clone().requires_grad_(True)
    g, b = ga.clone().requires_grad_(True), be.clone().requires_grad_(True)
    y = ops.patch_merge_ln(xx, (H, W), g, b)
    y.backward(dy)
    return y.detach(), xx.grad, g.grad, b.grad


def main():
    dev = 'cuda'
    gen = torch.Generator().manual_seed(0)
    bad = 0
    for dtype in (torch.float32, torch.bfloat16):
        for B, H, W, C in [(2, 16, 16, 96), (2, 8, 8, 192), (2, 4, 4, 384), (2, 32, 32, 32), (2, 16, 16, 64), (2, 8, 8, 128),
                           (16, 50, 50, 96)]:
            x = torch.randn(B, H * W, C, generator=gen).to(dtype).to(dev)
            ga, be = torch.randn(4 * C, generator=gen).to(dev), torch.randn(4 * C, generator=gen).to(dev)
            dy = torch.randn(B, ((H + 1) // 2) * ((W + 1) // 2), 4 * C, generator=gen).to(dtype).to(dev)
            _lib.call('rsc_set_patch_merge_variant', 1)
            r1 = once(x, ga, be, dy, H, W)
            _lib.call('rsc_set_patch_merge_variant', 2)
            r0 = once(x, ga, be, dy, H, W)
            mism = [0, 0, 0]
            for it in range(60):
                junk = [torch.full((int(torch.randint(1, 1 << 18, (1,))),), float('nan'), device=dev) for _ in range(4)]
                del junk
                r = once(x, ga, be, dy, H, W)
                mism[0] += int(not torch.equal(r[0], r0[0]))
                mism[1] += int(not torch.equal(r[1], r0[1]))
                mism[2] += int(not (torch.allclose(r[2], r0[2], rtol=1e-5, atol=1e-6) and torch.allclose(r[3], r0[3], rtol=1e-5, atol=1e-6)))
            v = [float((a.float() - b.float()).norm() / b.float().norm()) for a, b in zip(r0, r1)]
            bad += sum(mism)
            print('%s B%d %dx%d C%d: mismatching calls of 60 (y, dx, dgamma|dbeta) = %s; variant 2 vs 1 rel (y, dx, dg, db) = %s' % (
                str(dtype)[6:], B, H, W, C, mism, ' '.join('%.1e' % t for t in v)), flush=True)
    _lib.call('rsc_set_patch_merge_variant', 0)
    print('TOTAL mismatching calls', bad)


if __name__ == '__main__':
    main()
