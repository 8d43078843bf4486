"""Phase-level GPU time / kernel-count breakdown of one co-training cycle (eager, bf16).

For each task: wraps the top-level modules in record_function ranges, runs one iteration
under torch.profiler and attributes every CUDA kernel to the innermost enclosing range
(forward) or to its autograd node (backward).  Writes gpurun_out/phases_<task>.json and a
full per-kernel table gpurun_out/kernels_<task>.txt.
"""
import collections
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile, record_function

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rscotr_b200.mtl.engine.step import _to_device  # noqa: E402


def wrap(mod, name):
    stack = []

    def pre(m, a, k=None):
        r = record_function('phase:' + name)
        r.__enter__()
        stack.append(r)

    def post(m, a, o):
        stack.pop().__exit__(None, None, None)
    mod.register_forward_pre_hook(pre)
    mod.register_forward_hook(post)


def wrap_method(obj, meth, name):
    f = getattr(obj, meth)

    def g(*a, **k):
        with record_function('phase:' + name):
            return f(*a, **k)
    setattr(obj, meth, g)


def ksum(ev):
    n, t = 0, 0.0
    for k in ev.kernels:
        n += 1
        t += k.duration
    for c in ev.cpu_children:
        a, b = ksum(c)
        n += a
        t += b
    return n, t


def main():
    dev = torch.device('cuda', 0)
    os.environ['RSC_CUDA_GRAPHS'] = '0'
    cfg, model, engine, loader = bench.build(bench.CONFIG, 'bf16', dev)
    wrap(model.backbone, 'backbone')
    wrap(model.neck, 'neck')
    wrap(model.shared_encoder, 'shared_encoder')
    wrap(model.bbox_head.transformer.decoder, 'det_decoder')
    wrap_method(model.bbox_head, 'dn_generator', 'det_cdn')
    wrap_method(model.bbox_head, 'loss_prepare', 'det_loss_prepare')
    wrap_method(model.bbox_head, 'loss_assign', 'det_loss_assign')
    wrap_method(model.bbox_head, 'loss_finish', 'det_loss_finish')
    wrap(model.bbox_head, 'det_head_fwd_other')
    wrap(model.seg_head.pixel_decoder, 'seg_pixel_decoder_other')
    wrap_method(model.seg_head, 'losses', 'seg_losses')
    wrap_method(model.seg_head, 'forward', 'seg_head_fwd_other')
    wrap_method(model, '_finish', 'parse_losses')
    wrap_method(engine, '_collect_grads', 'collect_grads')
    wrap_method(engine.optimizer, 'step_flat', 'adamw')
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
        evs = prof.events()
        phases = collections.OrderedDict()
        claimed = set()

        def visit(ev, cur):
            name = cur
            if ev.name.startswith('phase:'):
                name = ev.name[6:]
            elif ev.name.startswith('autograd::engine::evaluate_function: '):
                name = 'bwd:' + ev.name.split(': ')[1]
            for k in ev.kernels:
                d = phases.setdefault(name, [0, 0.0])
                d[0] += 1
                d[1] += k.duration
            for c in ev.cpu_children:
                visit(c, name)
        for ev in evs:
            if ev.cpu_parent is None and ev.device_type == torch.autograd.DeviceType.CPU:
                visit(ev, 'other')
        tot_n = sum(v[0] for v in phases.values())
        tot_t = sum(v[1] for v in phases.values())
        fwd = {k: v for k, v in phases.items() if not k.startswith('bwd:')}
        bwd = {k: v for k, v in phases.items() if k.startswith('bwd:')}
        out = dict(task=b['task'], kernels=tot_n, cuda_ms=tot_t / 1000.0,
                   fwd={k: dict(n=v[0], ms=round(v[1] / 1000.0, 3)) for k, v in fwd.items()},
                   bwd_total=dict(n=sum(v[0] for v in bwd.values()), ms=round(sum(v[1] for v in bwd.values()) / 1000.0, 3)),
                   bwd={k: dict(n=v[0], ms=round(v[1] / 1000.0, 3))
                        for k, v in sorted(bwd.items(), key=lambda kv: -kv[1][1])[:40]})
        json.dump(out, open('gpurun_out/phases_%s.json' % b['task'], 'w'), indent=1)
        print(json.dumps(out))
        # full kernel table
        agg = collections.defaultdict(lambda: [0, 0.0])
        for ev in evs:
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                agg[ev.name][0] += 1
                agg[ev.name][1] += ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time
        with open('gpurun_out/kernels_%s.txt' % b['task'], 'w') as f:
            for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write('%9.1f us %6d  %s\n' % (t, n, k[:160]))


if __name__ == '__main__':
    main()
