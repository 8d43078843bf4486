"""torch.profiler breakdown of one co-training cycle (cls, det, seg) on the GPU box.
Writes gpurun_out/profile_<task>.txt (top CUDA kernels by total time)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rscotr_b200.mtl.engine.step import _to_device  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    os.environ['RSC_CUDA_GRAPHS'] = '0'      # eager: kernel names are what we are after
    cfg, model, engine, loader = bench.build(bench.CONFIG, 'bf16', dev)
    it = iter(loader)
    batches = [_to_device(next(it), dev) for _ in range(3)]
    for _ in range(2):
        for b in batches:
            engine.train_iter(b)
    torch.cuda.synchronize()
    os.makedirs('gpurun_out', exist_ok=True)
    for b in batches:
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            engine.train_iter(b)
            torch.cuda.synchronize()
        txt = prof.key_averages().table(sort_by='self_cuda_time_total', row_limit=40, max_name_column_width=90)
        open('gpurun_out/profile_%s.txt' % b['task'], 'w').write(txt)
        print('=====', b['task'])
        print(txt[-6000:])


if __name__ == '__main__':
    main()
