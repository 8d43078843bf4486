"""Launch the fused element-wise / deformable-attention kernels once each at the BASELINE workload's largest shapes
(cls batch 16, stage 0: 640 000 tokens x 96 / 384 channels; shared encoder at B=2) -- target for `ncu --set full`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rscotr_b200 import ops  # noqa: E402

dev, bf = 'cuda', torch.bfloat16
torch.manual_seed(0)
rows = 16 * 200 * 200
for it in range(2):
    # bias + GELU (fc1 of a stage-0 block)
    h = torch.randn(16, 40000, 384, device=dev, dtype=bf, requires_grad=True)
    b = torch.randn(384, device=dev, requires_grad=True)
    y = ops.bias_gelu(h, b)
    y.backward(torch.randn_like(y))
    # residual + DropPath + bias + LayerNorm
    ident = torch.randn(16, 40000, 96, device=dev, dtype=bf, requires_grad=True)
    x = torch.randn(16, 40000, 96, device=dev, dtype=bf, requires_grad=True)
    bias, g, be = (torch.randn(96, device=dev, requires_grad=True) for _ in range(3))
    scale = (torch.rand(16, device=dev) > 0.1).float() / 0.9
    r, n = ops.add_ln(ident, x, bias, scale, g, be)
    (r.float().sum() + (n.float() * 0.5).sum()).backward()
    # fused ms_deform_attn tail, shared-encoder shape
    shapes = [(100, 100), (50, 50), (25, 25), (13, 13)]
    Nv = sum(a * c for a, c in shapes)
    B = 2
    value = torch.randn(B, Nv, 8, 32, device=dev, dtype=bf, requires_grad=True)
    refs = []
    for hh, ww in shapes:
        ys, xs = torch.meshgrid((torch.arange(hh, device=dev) + 0.5) / hh, (torch.arange(ww, device=dev) + 0.5) / ww, indexing='ij')
        refs.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
    ref = torch.cat(refs)[None, :, None, :].repeat(B, 1, 4, 1).contiguous()
    off = (torch.randn(B, Nv, 8, 4, 4, 2, device=dev) * 2).to(bf).requires_grad_(True)
    lg = torch.randn(B, Nv, 8, 16, device=dev).to(bf).requires_grad_(True)
    out = ops.ms_deform_attn_fused(value, torch.tensor(shapes, device=dev), torch.tensor([0, 10000, 12500, 13125], device=dev),
                                   off, lg, ref)
    out.backward(torch.randn_like(out))
torch.cuda.synchronize()
print('done')
