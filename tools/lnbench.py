"""add + bias + LayerNorm passes (rsc_add_ln_{fwd,bwd}, rsc_layernorm_{fwd,bwd}) at the shapes of the BASELINE workload:
Swin stages 0-3 of the cls batch (16 x 200^2 / 100^2 / 50^2 / 25^2 tokens, C = 96..768), the seg batch (2 images) and the
shared encoder (2 x 13 294 tokens, C = 256).  CUDA-event medians, L2 flushed.   python tools/lnbench.py [tag]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rscotr_b200 import ops  # noqa: E402
from tools.kbench import timeit  # noqa: E402

SHAPES = [('cls.s0', 16, 40000, 96), ('cls.s1', 16, 10000, 192), ('cls.s2', 16, 2500, 384), ('cls.s3', 16, 625, 768),
          ('seg.s0', 2, 40000, 96), ('seg.s1', 2, 10000, 192), ('enc', 2, 13294, 256), ('det.enc', 1, 13294, 256)]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'lnbench'
    dev = 'cuda'
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for name, B, T, C in SHAPES:
        idn = torch.randn(B, T, C, device=dev, dtype=torch.bfloat16, requires_grad=True)
        x = torch.randn(B, T, C, device=dev, dtype=torch.bfloat16, requires_grad=True)
        gam, bet, bia = (torch.randn(C, device=dev).requires_grad_(True) for _ in range(3))
        scale = torch.ones(B, device=dev)
        n_el = idn.numel()
        t = timeit(lambda: ops.add_ln(idn, x, bia, scale, gam, bet, 1e-5), flush=flush)
        r, n = ops.add_ln(idn, x, bia, scale, gam, bet, 1e-5)
        dr, dn = torch.randn_like(r), torch.randn_like(n)
        tb = timeit(lambda: torch.autograd.grad((r, n), (idn, x, bia, gam, bet), (dr, dn), retain_graph=True), flush=flush)
        tl = timeit(lambda: ops.layer_norm(x, gam, bet, 1e-5), flush=flush)
        y = ops.layer_norm(x, gam, bet, 1e-5)
        tlb = timeit(lambda: torch.autograd.grad(y, (x, gam, bet), dn, retain_graph=True), flush=flush)
        rows.append(dict(name=name, B=B, T=T, C=C, MB=round(n_el * 2 / 1e6, 1),
                         add_ln_fwd_us=round(t * 1e3, 1), add_ln_fwd_gbs=round(4 * n_el * 2 / t / 1e6),
                         add_ln_bwd_us=round(tb * 1e3, 1), add_ln_bwd_gbs=round(6 * n_el * 2 / tb / 1e6),
                         ln_fwd_us=round(tl * 1e3, 1), ln_fwd_gbs=round(2 * n_el * 2 / tl / 1e6),
                         ln_bwd_us=round(tlb * 1e3, 1), ln_bwd_gbs=round(3 * n_el * 2 / tlb / 1e6)))
        print(json.dumps(rows[-1]), flush=True)
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/%s.jsonl' % tag, 'w') as f:
        for r_ in rows:
            f.write(json.dumps(r_) + '\n')


if __name__ == '__main__':
    main()
