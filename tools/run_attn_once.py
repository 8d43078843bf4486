"""Launch the decoder attention kernels once each at the BASELINE shapes (DINO self-attention 1100 x 1100 with the denoising
mask; seg decoder cross-attention 100 queries x 100^2 keys with the mask generated from mask_pred) -- target for
`ncu --set full`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rscotr_b200 import ops  # noqa: E402

dev, H, E = 'cuda', 8, 256
for Lq, Lk, B, kind in ((1100, 1100, 1, 'const'), (100, 10000, 2, 'pred')):
    q = torch.randn(Lq, B, E, device=dev, dtype=torch.bfloat16, requires_grad=True)
    k = torch.randn(Lk, B, E, device=dev, dtype=torch.bfloat16, requires_grad=True)
    v = torch.randn(Lk, B, E, device=dev, dtype=torch.bfloat16, requires_grad=True)
    if kind == 'const':
        m = torch.rand(Lq, Lk, device=dev) < 0.3
        m[:, 0] = False
        bits = ops.pack_mask_bits(m)
    else:
        bits = ops.m2f_attn_mask(torch.randn(B, Lq, 100, 100, device=dev, dtype=torch.bfloat16), (100, 100))
    for _ in range(2):
        o = ops.attention(q, k, v, H, bits)
        o.backward(torch.randn_like(o))
torch.cuda.synchronize()
print('done')
