"""Launch the bf16 window-attention forward/backward a few times at the stage-0 size of the
BASELINE workload (B=16, 200x200 tokens, C=96) -- target for `ncu --set full`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rscotr_b200 import ops  # noqa: E402

S, C, heads, B = int(os.environ.get('S', 200)), int(os.environ.get('C', 96)), int(os.environ.get('HEADS', 3)), 16
qkv = torch.randn(B, S * S, 3 * C, device='cuda', dtype=torch.bfloat16, requires_grad=True)
bias = torch.randn(3 * C, device='cuda')
table = torch.randn(169, heads, device='cuda')
dout = torch.randn(B, S * S, C, device='cuda', dtype=torch.bfloat16)
for i in range(4):
    o = ops.wmsa(qkv, bias, table, (S, S), heads, 7, 3 * (i % 2))
    o.backward(dout)
torch.cuda.synchronize()
print('done')
