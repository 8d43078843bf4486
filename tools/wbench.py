"""Window-attention micro-benchmark (bf16, BASELINE stage geometries): CUDA-event medians with an L2 flush between
iterations, achieved algorithmic GB/s (4 / 7 x B x Lp x C x 2 bytes) and the fraction of the measured HBM peak.
  python tools/wbench.py [tag]     env: KB_B (batch, default 16)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rscotr_b200 import ops  # noqa: E402
from tools.kbench import timeit  # noqa: E402


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'wbench'
    B = int(os.environ.get('KB_B', 16))
    peak = 6553.9
    try:
        peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
    except Exception:
        pass
    dev = 'cuda'
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for (S, C, heads) in [(200, 96, 3), (100, 192, 6), (50, 384, 12), (25, 768, 24)]:
        for shift in (0, 3):
            qkv = torch.randn(B, S * S, 3 * C, device=dev, dtype=torch.bfloat16)
            bias = torch.randn(3 * C, device=dev)
            table = torch.randn(169, heads, device=dev)
            dout = torch.randn(B, S * S, C, device=dev, dtype=torch.bfloat16)
            Sp = (S + 6) // 7 * 7
            alg = 4 * B * Sp * Sp * C * 2
            t = timeit(lambda: ops.wmsa(qkv, bias, table, (S, S), heads, 7, shift, 32 ** -0.5), flush=flush)
            rows.append(dict(k='wmsa_fwd', S=S, C=C, shift=shift, us=round(t * 1e3, 1), gbs=round(alg / t / 1e6),
                             frac=round(alg / t / 1e6 / peak, 3)))
            q2 = qkv.clone().requires_grad_(True)
            b2, t2 = bias.clone().requires_grad_(True), table.clone().requires_grad_(True)
            o = ops.wmsa(q2, b2, t2, (S, S), heads, 7, shift)
            tb = timeit(lambda: torch.autograd.grad(o, (q2, b2, t2), dout, retain_graph=True), flush=flush)
            rows.append(dict(k='wmsa_bwd', S=S, C=C, shift=shift, us=round(tb * 1e3, 1), gbs=round(7 * alg / 4 / tb / 1e6),
                             frac=round(7 * alg / 4 / tb / 1e6 / peak, 3)))
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/%s.jsonl' % tag, 'w') as f:
        for r in rows:
            print(json.dumps(r))
            f.write(json.dumps(r) + '\n')


if __name__ == '__main__':
    main()
