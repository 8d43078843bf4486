"""How much does ONE gradient of the small UPerNet / Swin model (the one of tools/replay_diag_uper.py) move when the outputs of
PatchMerging / every Linear are perturbed in their last bits?  Runs on the CPU through the test shim (oracle ops): no GPU.
Result (round 2): 1.5e-7 of the whole gradient for 3e-8 .. 1e-7 relative perturbations -- the model is NOT ill-conditioned, so the
3e-4 .. 5e-4 run-to-run differences seen on the GPU with the round-2 PatchMerging kernels come from the GPU path, not from the
mathematics (only gradients that are zero in exact arithmetic, norm biases in front of a BatchNorm, and the PPM 1x1 branch in front
of a batch-of-2 BatchNorm move more).      python tools/cpu_sensitivity_uper.py"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rscotr_b200.models
from rscotr_b200.config import MODELS, Config
from rscotr_b200.mtl.data import build_datasets
from rscotr_b200.mtl.engine import StepEngine
from tests.cpu_ops_shim import cpu_ops
from rscotr_b200 import ops
torch.set_num_threads(8)

def build():
    cfg = Config.fromfile(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'configs/seg/upernet_swin-b_512_potsdam.py'))
    m = cfg.model
    m.backbone.embed_dims, m.backbone.depths, m.backbone.num_heads, m.backbone.drop_path_rate = 32, [2, 2, 2, 2], [1, 2, 4, 8], 0.0
    m.decode_head.in_channels, m.decode_head.channels, m.decode_head.dropout_ratio = [32, 64, 128, 256], 32, 0.0
    m.auxiliary_head.in_channels, m.auxiliary_head.channels, m.auxiliary_head.dropout_ratio = 128, 16, 0.0
    torch.manual_seed(0)
    model = MODELS.build(m); model.init_weights(); model.train()
    ds = build_datasets({'potsdam': dict(task='seg')}, synthetic=dict(img_size=(128, 128), seg=dict(num_classes=6)))['potsdam']
    batch = ds.make_batch(2, torch.Generator().manual_seed(0), pin=False)
    batch.update(task='seg', dataset_name='potsdam')
    return model, batch

def grads(perturb_pm=0.0, perturb_lin=0.0, seed=1):
    model, batch = build()
    eng = StepEngine(model, dict(type='SGD', lr=1e-2, momentum=0.9), device='cpu', compute_dtype=torch.float32, use_graphs=False)
    g = torch.Generator().manual_seed(seed)
    with cpu_ops():
        pm0, lin0 = ops.patch_merge_ln, ops.linear
        if perturb_pm:
            def pm(x, hw, ga, be, eps=1e-5):
                y = pm0(x, hw, ga, be, eps)
                return y * (1 + perturb_pm * torch.randn(y.shape, generator=g))
            ops.patch_merge_ln = pm
        if perturb_lin:
            def lin(x, w, b=None, rows=None):
                y = lin0(x, w, b, rows)
                return y * (1 + perturb_lin * torch.randn(y.shape, generator=g))
            ops.linear = lin
        eng.train_iter(batch)
    return eng.flat_grad.clone(), {n: (s, e) for n, s, e in eng._spans}

g0, spans = grads()
g0b, _ = grads()
print('repeat (no perturbation): rel', float((g0b - g0).norm() / g0.norm()))
for ppm, plin in [(3e-8, 0), (0, 3e-8), (1e-7, 1e-7)]:
    g1, _ = grads(ppm, plin)
    rel = float((g1 - g0).norm() / g0.norm())
    worst = sorted(((float((g1[s:e] - g0[s:e]).norm() / g0[s:e].norm().clamp_min(1e-20)), n) for n, (s, e) in spans.items()), reverse=True)[:4]
    print('perturb pm=%g lin=%g: flat grad rel %.2e; worst spans %s' % (ppm, plin, rel, [(round(a, 6), n) for a, n in worst]))
