"""Decoder-attention micro-benchmark: rsc_attn_{fwd,bwd} (+ rsc_m2f_mask_bits) against what they replace (library SDPA
with a materialised boolean mask, as bricks.MultiheadAttention did before).  CUDA-event medians, no L2 flush (the
operands of these launches are L2-resident in the step as well).   python tools/abench.py [tag]"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rscotr_b200 import ops  # noqa: E402
from tools.kbench import timeit  # noqa: E402

CASES = [('dino.self', 1100, 1100, 1, 'const'), ('m2f.cross.l0', 100, 10000, 2, 'pred'), ('m2f.cross.l1', 100, 2500, 2, 'pred'),
         ('m2f.cross.l2', 100, 625, 2, 'pred'), ('m2f.self', 100, 100, 2, None)]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'abench'
    dev, H, E = 'cuda', 8, 256
    rows = []
    for name, Lq, Lk, B, kind in CASES:
        q = torch.randn(Lq, B, E, device=dev, dtype=torch.bfloat16, requires_grad=True)
        k = torch.randn(Lk, B, E, device=dev, dtype=torch.bfloat16, requires_grad=True)
        v = torch.randn(Lk, B, E, device=dev, dtype=torch.bfloat16, requires_grad=True)
        dout = torch.randn(Lq, B, E, device=dev, dtype=torch.bfloat16)
        bits = boolmask = None
        t_mask_own = t_mask_lib = 0.0
        if kind == 'const':
            m = torch.rand(Lq, Lk, device=dev) < 0.3
            m[:, 0] = False
            bits, boolmask = ops.pack_mask_bits(m), (~m).view(1, 1, Lq, Lk)
        elif kind == 'pred':
            side = int(Lk ** 0.5)
            pred = torch.randn(B, Lq, 100, 100, device=dev, dtype=torch.bfloat16)
            bits = ops.m2f_attn_mask(pred, (side, side))
            t_mask_own = timeit(lambda: ops.m2f_attn_mask(pred, (side, side)))

            def lib_mask():
                a = ops.bilinear_resize(pred, (side, side))
                a = (a.flatten(2).float().sigmoid() < 0.5).unsqueeze(1)
                a = a & ~a.all(-1, keepdim=True)
                return ~a
            boolmask = lib_mask()
            t_mask_lib = timeit(lib_mask)

        def sd(t, L):
            return t.view(L, B, H, 32).permute(1, 2, 0, 3)
        own_f = lambda: ops.attention(q, k, v, H, bits)
        lib_f = lambda: F.scaled_dot_product_attention(sd(q, Lq), sd(k, Lk), sd(v, Lk), attn_mask=boolmask)
        t_of, t_lf = timeit(own_f), timeit(lib_f)
        o1, o2 = own_f(), lib_f()
        t_ob = timeit(lambda: torch.autograd.grad(o1, (q, k, v), dout, retain_graph=True))
        d2 = dout.view(Lq, B, H, 32).permute(1, 2, 0, 3)
        t_lb = timeit(lambda: torch.autograd.grad(o2, (q, k, v), d2, retain_graph=True))
        rows.append(dict(name=name, Lq=Lq, Lk=Lk, B=B, own_fwd_us=round(t_of * 1e3, 1), lib_fwd_us=round(t_lf * 1e3, 1),
                         own_bwd_us=round(t_ob * 1e3, 1), lib_bwd_us=round(t_lb * 1e3, 1), own_mask_us=round(t_mask_own * 1e3, 1),
                         lib_mask_us=round(t_mask_lib * 1e3, 1)))
        print(json.dumps(rows[-1]), flush=True)
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/%s.jsonl' % tag, 'w') as f:
        for r in rows:
            f.write(json.dumps(r) + '\n')


if __name__ == '__main__':
    main()
