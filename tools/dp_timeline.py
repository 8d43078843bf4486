"""Kernel timeline of replayed data-parallel steps on rank 0 (torchrun, N >= 1): per step, the busy time of the compute
stream, the idle gaps longer than 15 us (with the kernels on either side) and when the exchange kernels ran.
   torchrun --nproc-per-node 2 tools/dp_timeline.py"""
import os
import sys

import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rscotr_b200.mtl.engine.step import _to_device  # noqa: E402


def main():
    world, rank, local = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    cfg, model, engine, loader = bench.build(bench.CONFIG, 'bf16', device)
    it = iter(loader)
    dev = [_to_device(next(it), device) for _ in range(6)]
    for i in range(18):
        engine.train_iter(dev[i % 6])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(6):
            engine.train_iter(dev[i % 6])
        torch.cuda.synchronize()
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
        evs.sort(key=lambda e: e.time_range.start)
        t0 = evs[0].time_range.start
        comm = [e for e in evs if 'nccl' in e.name.lower() or 'nvls' in e.name.lower() or 'barrier' in e.name.lower()]
        comp = [e for e in evs if e not in comm]
        print('kernels %d (comm %d), span %.2f ms for 6 steps' % (len(evs), len(comm), (evs[-1].time_range.end - t0) / 1e3))
        busy = sum(e.time_range.end - e.time_range.start for e in comp) / 1e3
        print('compute-kernel busy time %.2f ms' % busy)
        gaps = []
        for a, b in zip(comp, comp[1:]):
            g = b.time_range.start - a.time_range.end
            if g > 15:
                gaps.append((g, a, b))
        print('idle gaps > 15 us between consecutive compute kernels: %d, total %.2f ms' % (len(gaps), sum(g for g, _, _ in gaps) / 1e3))
        for g, a, b in sorted(gaps, key=lambda x: -x[0])[:14]:
            print('  gap %7.1f us at t=%8.2f ms   after %-48s before %-48s' % (g, (a.time_range.end - t0) / 1e3, a.name[:48], b.name[:48]))
        for e in comm[:40]:
            print('  comm %-60s start %8.2f ms  dur %7.1f us' % (e.name[:60], (e.time_range.start - t0) / 1e3,
                                                                 e.time_range.end - e.time_range.start))
    # leave like bench.py does: tearing NCCL down while captured collectives are alive can block
    sys.stdout.flush()
    if world > 1:
        engine._graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
    os._exit(0)


if __name__ == '__main__':
    main()
