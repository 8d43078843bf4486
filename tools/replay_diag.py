"""How far apart do two fp32 trainings of the small seg / det model end up after 5 SGD steps -- eager vs eager (the floor
set by the order of the fp32 atomic sums, amplified by thresholded masks / assignments) and eager vs CUDA-graph replay --
with the round's switches on and off?  Diagnostic for tests/test_gpu_model.py::test_cuda_graph_replay_matches_eager.
    python tools/replay_diag.py [task ...]        (RSC_PATCH_MERGE_V2=1 selects the round-2 PatchMerging kernels)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from oracle import heads as oh  # noqa: E402
from rscotr_b200 import ops  # noqa: E402
from rscotr_b200.mtl.engine import StepEngine  # noqa: E402
from tests.test_host_parity import _setup  # noqa: E402


def run(task, use_graphs, iters=5):
    model, batch = _setup(task, seed=7)
    if task == 'det':
        noise = oh.cdn_noise(batch['gt_labels'], num_dn=10, generator=torch.Generator().manual_seed(5))
        model.bbox_head.dn_generator.forced_noise = {k: v.cuda() for k, v in noise.items()}
    eng = StepEngine(model, dict(type='SGD', lr=1e-2, momentum=0.9), grad_clip=dict(max_norm=0.1, norm_type=2),
                     device='cuda', compute_dtype=torch.float32, use_graphs=use_graphs)
    grads = []
    for _ in range(iters):
        eng.train_iter(batch)
        torch.cuda.synchronize()
        grads.append(eng.flat_grad.clone())
    return {n: p.detach().clone() for n, p in model.named_parameters()}, grads


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12))


def main():
    tasks = sys.argv[1:] or ['seg']
    for task in tasks:
        for pair in (True, False):
            ops._LINEAR_PAIR = pair
            e0, g0 = run(task, False)
            e1, g1 = run(task, False)
            gr, g2 = run(task, True)
            for tag, (a, ga) in (('eager-eager', (e1, g1)), ('eager-graph', (gr, g2))):
                worst = max((rel(a[n], e0[n]), n) for n in e0)
                print('%s pm_v2=%s pair=%d %s: worst param rel %.2e (%s); flat grad rel per iter %s' % (
                    task, os.environ.get("RSC_PATCH_MERGE_V2", "0"), pair, tag, worst[0], worst[1],
                    ' '.join('%.1e' % rel(x, y) for x, y in zip(ga, g0))), flush=True)


if __name__ == '__main__':
    main()
