"""Times the flat-gradient NCCL all-reduce sizes of the three tasks (cls 27.6 M, det 48 M, seg 52 M fp32 elements)
and a few small packed reductions.  torchrun --nproc-per-node N tools/allreduce_probe.py"""
import os

import torch
import torch.distributed as dist

local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
world = dist.get_world_size()
for n in (16, 27_600_000, 48_000_000, 52_000_000):
    x = torch.ones(n, device='cuda')
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        dist.all_reduce(x)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    if dist.get_rank() == 0:
        print('world %d  all_reduce %10d fp32: %.3f ms  (%.1f GB/s algbw)' % (world, n, ms, n * 4 / ms / 1e6), flush=True)
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
