"""Which Python call sites launch the many tiny ATen kernels of a det / seg step?  Groups aten::fill_ / copy_ /
add / mul ... launches by the innermost rscotr_b200 source line (torch.profiler with_stack)."""
import collections
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
    os.environ['RSC_CUDA_GRAPHS'] = '0'
    cfg, model, engine, loader = bench.build(bench.CONFIG, 'bf16', dev)
    it = iter(loader)
    batches = [_to_device(next(it), dev) for _ in range(3)]
    for _ in range(2):
        for b in batches:
            engine.train_iter(b)
    torch.cuda.synchronize()
    want = sys.argv[1:] or ['det', 'seg']
    for b in batches:
        if b['task'] not in want:
            continue
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
            engine.train_iter(b)
            torch.cuda.synchronize()
        by_site = collections.Counter()
        by_site_us = collections.Counter()
        for ev in prof.events():
            if ev.device_type != torch.autograd.DeviceType.CPU or not ev.kernels:
                continue
            chain, par = [], ev.cpu_parent
            while par is not None and len(chain) < 4:
                chain.append(par.name.replace('autograd::engine::evaluate_function: ', 'bwd:'))
                par = par.cpu_parent
            site = ' < '.join(chain) if chain else '(top level)'
            key = '%s | %s' % (ev.name, site)
            by_site[key] += len(ev.kernels)
            by_site_us[key] += sum(k.duration for k in ev.kernels)
        print('=====', b['task'], 'kernels', sum(by_site.values()))
        for k, n in by_site.most_common(70):
            print('%5d %8.1f us  %s' % (n, by_site_us[k], k[:150]))


if __name__ == '__main__':
    main()
