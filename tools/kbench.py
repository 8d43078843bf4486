"""Per-kernel micro-benchmark at BASELINE (Swin-T, 800^2) sizes: CUDA-event
timings and achieved algorithmic GB/s.  Run on the GPU box:  python tools/kbench.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rscotr_b200 import ops  # noqa: E402


def timeit(fn, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    B = int(os.environ.get('KB_B', 4))
    dev = 'cuda'
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2 (126 MB)
    rows = []
    stages = [(200, 96, 3), (100, 192, 6), (50, 384, 12), (25, 768, 24)]
    for dt in (torch.bfloat16, torch.float32):
        es = 2 if dt == torch.bfloat16 else 4
        for (S, C, heads) in stages:
            for shift in (0, 3):
                qkv = torch.randn(B, S * S, 3 * C, device=dev, dtype=dt)
                bias = torch.randn(3 * C, device=dev)
                table = torch.randn(169, heads, device=dev)
                dout = torch.randn(B, S * S, C, device=dev, dtype=dt)
                Sp = (S + 6) // 7 * 7
                alg = 4 * B * Sp * Sp * C * es           # read qkv + write out (padded tokens)
                t = timeit(lambda: ops.wmsa(qkv, bias, table, (S, S), heads, 7, shift, 32 ** -0.5), flush=flush)
                rows.append(dict(k='wmsa_fwd', dtype=str(dt), S=S, C=C, shift=shift, ms=t, gbs=alg / t / 1e6))
                q2 = qkv.clone().requires_grad_(True)
                o = ops.wmsa(q2, bias, table, (S, S), heads, 7, shift)
                tb = timeit(lambda: torch.autograd.grad(o, q2, dout, retain_graph=True), flush=flush)
                rows.append(dict(k='wmsa_bwd', dtype=str(dt), S=S, C=C, shift=shift, ms=tb, gbs=2 * alg / tb / 1e6))
        # msda encoder shape
        shapes = [(100, 100), (50, 50), (25, 25), (13, 13)]
        Nv = sum(h * w for h, w in shapes)
        value = torch.randn(B, Nv, 8, 32, device=dev, dtype=dt)
        loc = torch.rand(B, Nv, 8, 4, 4, 2, device=dev)
        w = torch.rand(B, Nv, 8, 4, 4, device=dev).flatten(-2).softmax(-1).view(B, Nv, 8, 4, 4)
        ss = torch.tensor(shapes, device=dev)
        st = torch.tensor([0, 10000, 12500, 13125], device=dev)
        alg = B * Nv * 256 * (2 * es + 12)
        t = timeit(lambda: ops.ms_deform_attn(value, ss, st, loc, w), flush=flush)
        rows.append(dict(k='msda_fwd', dtype=str(dt), ms=t, gbs=alg / t / 1e6))
        v2, l2, w2 = value.clone().requires_grad_(True), loc.clone().requires_grad_(True), w.clone().requires_grad_(True)
        o = ops.ms_deform_attn(v2, ss, st, l2, w2)
        go = torch.randn_like(o)
        tb = timeit(lambda: torch.autograd.grad(o, (v2, l2, w2), go, retain_graph=True), flush=flush)
        rows.append(dict(k='msda_bwd', dtype=str(dt), ms=tb, gbs=2 * alg / tb / 1e6))
        # patch merging stage 0
        x = torch.randn(B, 200 * 200, 96, device=dev, dtype=dt)
        g, b_ = torch.ones(384, device=dev), torch.zeros(384, device=dev)
        t = timeit(lambda: ops.patch_merge_ln(x, (200, 200), g, b_), flush=flush)
        rows.append(dict(k='patch_merge_ln_fwd', dtype=str(dt), ms=t, gbs=2 * x.numel() * es / t / 1e6))
        # bilinear seg-loss upsample (B,100,100,100)->800
        xs = torch.randn(1, 100, 100, 100, device=dev, dtype=dt)
        t = timeit(lambda: ops.bilinear_resize(xs, (800, 800)), flush=flush)
        rows.append(dict(k='bilinear_up8', dtype=str(dt), ms=t, gbs=xs.numel() * 65 * es / t / 1e6))
    # fused element-wise passes at the stage-0 sizes of the cls task (B x 40000 tokens): bias + GELU on the 4C hidden
    # activations (default erf form and the opt-in logistic fit, act=2), add + bias + LayerNorm on C
    Bc = int(os.environ.get('KB_BC', 16))
    for C, hid in ((96, 384), (192, 768)):
        ntok = Bc * (40000 if C == 96 else 10000)
        h = torch.randn(ntok, hid, device=dev, dtype=torch.bfloat16)
        hb = torch.randn(hid, device=dev)
        dy = torch.randn_like(h)
        for act, name in ((ops.ACT_GELU, 'gelu_erf'), (ops.ACT_GELU_SIG, 'gelu_logistic')):
            t = timeit(lambda: ops.bias_act(h, hb, act), flush=flush)
            rows.append(dict(k='bias_act_fwd', variant=name, C=hid, ms=t, gbs=2 * h.numel() * 2 / t / 1e6))
            h2, b2 = h.clone().requires_grad_(True), hb.clone().requires_grad_(True)
            y = ops.bias_act(h2, b2, act)
            tb = timeit(lambda: torch.autograd.grad(y, (h2, b2), dy, retain_graph=True), flush=flush)
            rows.append(dict(k='bias_act_bwd', variant=name, C=hid, ms=tb, gbs=3 * h.numel() * 2 / tb / 1e6))
        idn = torch.randn(Bc, ntok // Bc, C, device=dev, dtype=torch.bfloat16)
        xx = torch.randn_like(idn)
        gam, bet, bia = torch.ones(C, device=dev), torch.zeros(C, device=dev), torch.randn(C, device=dev)
        t = timeit(lambda: ops.add_ln(idn, xx, bia, None, gam, bet, 1e-5), flush=flush)
        rows.append(dict(k='add_ln_fwd', C=C, ms=t, gbs=4 * idn.numel() * 2 / t / 1e6))
    for r in rows:
        print(json.dumps(r))
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/kbench.jsonl', 'w') as f:
        for r in rows:
            f.write(json.dumps(r) + '\n')


if __name__ == '__main__':
    main()
