"""tests/test_uper_head.py::test_gpu_encoder_decoder_graph_replay_matches_eager, instrumented: how far apart are two EAGER
trainings, and eager vs CUDA-graph replay, in units of that test's tolerance (1e-3 * |p| + 1e-8)?
    python tools/replay_diag_uper.py         (RSC_PATCH_MERGE_V2=1 selects the round-2 PatchMerging kernels)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import rscotr_b200.models  # noqa: E402,F401
from rscotr_b200.config import MODELS, Config  # noqa: E402
from rscotr_b200.mtl.data import build_datasets  # noqa: E402
from rscotr_b200.mtl.engine import StepEngine  # noqa: E402


def run(use_graphs):
    cfg = Config.fromfile('configs/seg/upernet_swin-b_512_potsdam.py')
    m = cfg.model
    m.backbone.embed_dims, m.backbone.depths, m.backbone.num_heads, m.backbone.drop_path_rate = 32, [2, 2, 2, 2], [1, 2, 4, 8], 0.0
    m.decode_head.in_channels, m.decode_head.channels, m.decode_head.dropout_ratio = [32, 64, 128, 256], 32, 0.0
    m.auxiliary_head.in_channels, m.auxiliary_head.channels, m.auxiliary_head.dropout_ratio = 128, 16, 0.0
    torch.manual_seed(0)
    model = MODELS.build(m)
    model.init_weights()
    model.train()
    ds = build_datasets({'potsdam': dict(task='seg')}, synthetic=dict(img_size=(128, 128), seg=dict(num_classes=6)))['potsdam']
    batch = ds.make_batch(2, torch.Generator().manual_seed(0), pin=False)
    batch.update(task='seg', dataset_name='potsdam')
    eng = StepEngine(model, dict(type='SGD', lr=1e-2, momentum=0.9), device='cuda', compute_dtype=torch.float32,
                     use_graphs=use_graphs)
    g = []
    for _ in range(5):
        eng.train_iter(batch)
        torch.cuda.synchronize()
        g.append(eng.flat_grad.clone())
    return {n: p.detach().clone() for n, p in model.named_parameters()}, g


def main():
    e0, g0 = run(False)
    for tag, ug in (('eager', False), ('eager', False), ('graph', True), ('graph', True)):
        p, g = run(ug)
        worst = max((float((p[n] - e0[n]).norm()) / (1e-3 * float(e0[n].norm()) + 1e-8), n) for n in e0)
        gr = ' '.join('%.1e' % float((a - b).norm() / b.norm()) for a, b in zip(g, g0))
        print('pm_v2=%s %s vs eager#0: worst |dp| / tol = %.2f (%s); flat grad rel per iter %s' % (
            os.environ.get("RSC_PATCH_MERGE_V2", "0"), tag, worst[0], worst[1], gr), flush=True)


if __name__ == '__main__':
    main()
