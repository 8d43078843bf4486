"""Launch the tcgen05 Linear kernels once each at the Swin stage-0 / stage-1 shapes of the BASELINE workload
(cls batch 16: M = 640 000 / 160 000 tokens) and at the encoder FFN shape -- target for `ncu --set full`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rscotr_b200 import _lib  # noqa: E402

st = lambda: torch.cuda.current_stream().cuda_stream
for name, M, N, K, act in [('s0.fc1', 640000, 384, 96, 1), ('s0.fc2', 640000, 96, 384, 0), ('s1.fc1', 160000, 768, 192, 1),
                           ('enc.ffn1', 26588, 2048, 256, 2)]:
    x = torch.randn(M, K, device='cuda', dtype=torch.bfloat16)
    w = (torch.randn(N, K, device='cuda') * K ** -0.5).bfloat16()
    b = torch.randn(N, device='cuda')
    y, h = torch.empty(M, N, device='cuda', dtype=torch.bfloat16), torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
    dy = torch.randn(M, N, device='cuda', dtype=torch.bfloat16)
    dx = torch.empty(M, K, device='cuda', dtype=torch.bfloat16)
    dw, db = torch.zeros(N, K, device='cuda'), torch.zeros(N, device='cuda')
    for _ in range(2):
        _lib.call('rsc_linear_fwd', x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), h.data_ptr() if act == 1 else None,
                  M, N, K, K, K, N, act, st())
        _lib.call('rsc_linear_dx', dy.data_ptr(), w.data_ptr(), None, dx.data_ptr(), M, N, K, N, K, K, 0, st())
        _lib.call('rsc_linear_dw', dy.data_ptr(), x.data_ptr(), dw.data_ptr(), db.data_ptr(), M, N, K, N, K, K, st())
    torch.cuda.synchronize()
print('done')
