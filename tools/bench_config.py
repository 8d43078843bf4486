"""Time the step engine on any model config with synthetic batches of its `synthetic` shape (one task per
dataset entry, round robin), e.g. the segmentation-only Swin-B + UPerNet configuration:

    python tools/bench_config.py configs/seg/upernet_swin-b_512_potsdam.py --steps 20

Prints one JSON line: iterations/s, images/s, ms per iteration (CUDA events, after warm-up + graph capture)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('config')
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--batch', type=int, default=None)
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    args = ap.parse_args()
    import rscotr_b200.models  # noqa: F401
    from rscotr_b200.config import Config, MODELS
    from rscotr_b200.mtl.data import build_datasets
    from rscotr_b200.mtl.engine import StepEngine
    cfg = Config.fromfile(args.config)
    torch.manual_seed(0)
    model = MODELS.build(cfg.model)
    model.init_weights()
    model.train()
    eng = StepEngine(model, dict(cfg.optimizer), grad_clip=(cfg.get('optimizer_config') or {}).get('grad_clip'), device='cuda',
                     compute_dtype=torch.bfloat16 if args.dtype == 'bf16' else torch.float32)
    entries = {k: dict(task=v['task']) for k, v in cfg.data.items()}
    sets = build_datasets(entries, synthetic=dict(cfg.get('synthetic') or {}))
    batches = []
    for i, (name, ds) in enumerate(sets.items()):
        bs = args.batch or cfg.data[name].get('data', {}).get('samples_per_gpu', 1)
        b = ds.make_batch(bs, torch.Generator().manual_seed(i), pin=True)
        b.update(task=ds.task, dataset_name=name)
        batches.append(b)
    n = len(batches)
    for i in range(args.warmup * n):
        eng.train_iter(batches[i % n])
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps * n):
        out = eng.train_iter(batches[i % n])
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / (args.steps * n)
    imgs = sum(b['img'].shape[0] for b in batches) / n
    print(json.dumps(dict(config=args.config, dtype=args.dtype, it_per_s=1000.0 / ms, img_per_s=imgs * 1000.0 / ms, ms_per_iter=ms,
                          batch=[int(b['img'].shape[0]) for b in batches], img=list(batches[0]['img'].shape[1:]),
                          loss=float(out['loss'].detach()), graphs=sum('gA' in s for s in eng._graphs.values()),
                          graph_failures=eng.graph_failures, peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)))


if __name__ == '__main__':
    main()
