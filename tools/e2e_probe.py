"""Where does the end-to-end loop lose time against the device-resident loop?  Same engine, 30 steps each:
  A  device batches (bench timed region A)
  B  host batches + prefetch + loss read-back one step later (bench timed region B)
  C  B without the loss read-back
  D  B without the prefetch (train_iter copies from pinned host memory itself)
  E  device batches + loss read-back one step later
   python tools/e2e_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rscotr_b200.mtl.engine.step import _to_device  # noqa: E402


def main():
    device = torch.device('cuda', 0)
    cfg, model, engine, loader = bench.build(bench.CONFIG, 'bf16', device)
    it = iter(loader)
    NB, steps = 6, 30
    host = [next(it) for _ in range(NB)]
    dev = [_to_device(b, device) for b in host]
    for i in range(12):
        engine.train_iter(dev[i % NB])
    torch.cuda.synchronize()

    def run(batches, prefetch, readback):
        loss_host = torch.empty(steps, dtype=torch.float32).pin_memory()
        evs = [None] * steps
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if prefetch:
            engine.prefetch(batches[0])
        for i in range(steps):
            o = engine.train_iter(batches[i % NB])
            if readback:
                loss_host[i:i + 1].copy_(o['loss'].detach().reshape(1).float(), non_blocking=True)
                evs[i] = torch.cuda.Event()
                evs[i].record()
            if prefetch and i + 1 < steps:
                engine.prefetch(batches[(i + 1) % NB])
            if readback and i > 0:
                evs[i - 1].synchronize()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps

    for rep in range(3):
        sampler = None
        if rep == 2:      # third pass: with bench.py's nvidia-smi clock sampler running
            sampler = bench.ClockSampler(0)
            sampler.start()
            print('-- with the nvidia-smi sampler (-lms 100)')
        print('A device            %.3f ms/step' % run(dev, False, False))
        print('B host+pref+read    %.3f' % run(host, True, True))
        print('C host+pref         %.3f' % run(host, True, False))
        print('D host+read         %.3f' % run(host, False, True))
        print('E device+read       %.3f' % run(dev, False, True), flush=True)
        if sampler is not None:
            print(sampler.stop())
    # what bench.py does between its two timed regions: a sustained leg, then an eager per-kernel pass
    for i in range(240):
        engine.train_iter(dev[i % NB])
    from rscotr_b200 import ops
    with ops.KernelTimer() as kt:
        for i in range(NB):
            engine.train_iter(dev[i % NB])
        kt.summary()
    print('-- after a sustained leg + an eager KernelTimer pass')
    print('B host+pref+read    %.3f' % run(host, True, True))
    print('A device            %.3f' % run(dev, False, False))
    print('B host+pref+read    %.3f' % run(host, True, True))
    print('A device            %.3f' % run(dev, False, False), flush=True)


if __name__ == '__main__':
    main()
