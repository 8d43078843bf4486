"""CPU parity of the WHOLE co-training step's host logic: product modules (CUDA ops
substituted by the oracle shim) vs the independent functional oracle (oracle/heads.py),
same weights, same batch, same injected CDN noise.  Loss dicts and gradients must agree
to fp32 round-off; this pins module wiring, matching, target assignment and loss
bookkeeping (the kernels themselves are pinned by tests/test_gpu_*.py)."""
import os

import pytest
import torch

import rscotr_b200.models  # noqa: F401
from oracle import heads as oh
from rscotr_b200.config import MODELS
from rscotr_b200.mtl.data import build_datasets
from tests.cpu_ops_shim import cpu_ops
from tests.test_host_model import small_cfg

OCFG = dict(enc_layers=2, det_dec_layers=2, seg_dec_layers=5, num_query=30, num_dn=10)


def _setup(task, seed=0, img=64):
    torch.manual_seed(seed)
    cfg = small_cfg()
    model = MODELS.build(cfg.model)
    model.init_weights()
    # make the zero-initialised pieces non-trivial so every path carries signal
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if ('sampling_offsets.weight' in n or 'attention_weights' in n or 'reg_branches' in n and n.endswith('4.weight')):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
            if 'qkv.bias' in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.2)
    model.train()
    ds = build_datasets({'x': dict(task=task)}, synthetic=dict(img_size=(img, img), det=dict(num_boxes=3)))['x']
    batch = ds.make_batch(2, torch.Generator().manual_seed(seed + 2), pin=False)
    batch.update(task=task, dataset_name='x')
    return model, batch


@pytest.mark.parametrize('task', ['cls', 'det', 'seg'])
def test_step_matches_oracle_cpu(task):
    model, batch = _setup(task)
    noise = None
    if task == 'det':
        noise = oh.cdn_noise(batch['gt_labels'], num_dn=10, generator=torch.Generator().manual_seed(5))
        model.bbox_head.dn_generator.forced_noise = noise
    with cpu_ops():
        out = model.train_step(dict(batch), None)
        out['loss'].backward()
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    losses = oh.mtl_losses(sd, task, dict(batch), cfg=OCFG, noise=noise)
    loss, log_vars = oh.parse_losses(losses, model.task_weight[task])
    loss.backward()
    got = dict(out['log_vars'].items())
    for k, v in log_vars.items():
        key = '%s.x.%s' % (task, k)
        assert key in got, key
        assert abs(got[key] - v) <= 1e-4 * max(1.0, abs(v)), (key, got[key], v)
    assert len(got) == len(log_vars)
    checked = 0
    for n, p in model.named_parameters():
        go = sd[n].grad
        if p.grad is None:
            assert go is None or float(go.abs().max()) == 0.0, n
            continue
        assert go is not None, n
        err = (p.grad - go).norm() / go.norm().clamp_min(1e-9)
        assert float(err) < 2e-3 or float((p.grad - go).abs().max()) < 1e-6, (n, float(err))
        checked += 1
    assert checked > 50
