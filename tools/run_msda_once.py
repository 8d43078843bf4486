"""Launch ms_deform_attn forward/backward at the shared-encoder shape of the BASELINE workload
(800x800 -> levels 100^2, 50^2, 25^2, 13^2 = 13 294 tokens, 8 heads x 32, 4 points) -- target for `ncu`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rscotr_b200 import ops  # noqa: E402

B = int(os.environ.get('B', 2))
shapes = [(100, 100), (50, 50), (25, 25), (13, 13)]
Nv = sum(h * w for h, w in shapes)
dev = 'cuda'
torch.manual_seed(0)
value = torch.randn(B, Nv, 8, 32, device=dev, dtype=torch.bfloat16, requires_grad=True)
# reference points on each level's grid (as the encoder does) + small learned-offset-like jitter
refs = []
for h, w in shapes:
    ys, xs = torch.meshgrid((torch.arange(h, device=dev) + 0.5) / h, (torch.arange(w, device=dev) + 0.5) / w, indexing='ij')
    refs.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
ref = torch.cat(refs)[None, :, None, None, None, :]                          # (1, Nq, 1, 1, 1, 2)
loc = (ref + torch.randn(B, Nv, 8, 4, 4, 2, device=dev) * 0.03).requires_grad_(True)
w = torch.rand(B, Nv, 8, 16, device=dev).softmax(-1).view(B, Nv, 8, 4, 4).requires_grad_(True)
ss = torch.tensor(shapes, device=dev)
st = torch.tensor([0, 10000, 12500, 13125], device=dev)
for i in range(3):
    out = ops.ms_deform_attn(value, ss, st, loc, w)
    out.backward(torch.randn_like(out))
torch.cuda.synchronize()
print('done')
