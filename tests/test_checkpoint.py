"""Checkpoint compatibility on CPU: official-Swin -> mmdet-layout conversion (keys, PatchMerging
permutation semantics, bias-table resize), the runner checkpoint layout, resume."""
import os
import warnings
from collections import OrderedDict

import pytest
import torch
import torch.nn.functional as F

import rscotr_b200.models  # noqa: F401
from rscotr_b200.config import MODELS
from rscotr_b200.mtl.engine import StepEngine
from rscotr_b200.mtl.utils import checkpoint as C
from rscotr_b200.models.swin import SwinTransformer, PatchMerging
from tests.cpu_ops_shim import cpu_ops
from tests.test_host_model import small_cfg


def _official_from(backbone):
    """inverse of the converter: an 'official' Swin state dict (microsoft/Swin-Transformer names and
    PatchMerging channel order) that must convert back to `backbone`'s weights exactly."""
    inv = [0, 2, 1, 3]                                      # the q permutation is an involution
    out = OrderedDict()
    for k, v in backbone.state_dict().items():
        k2 = k.replace('stages', 'layers', 1).replace('attn.w_msa.', 'attn.').replace('ffn.layers.0.0.', 'mlp.fc1.') \
              .replace('ffn.layers.1.', 'mlp.fc2.').replace('patch_embed.projection', 'patch_embed.proj')
        if 'downsample.reduction.' in k:
            o, i = v.shape
            v = v.reshape(o, i // 4, 4).transpose(1, 2)[:, inv, :].reshape(o, i)
        elif 'downsample.norm.' in k:
            c = v.shape[0]
            v = v.reshape(c // 4, 4).transpose(0, 1)[inv, :].reshape(c)
        out[k2] = v.clone()
    out['head.weight'] = torch.zeros(10, 8)                 # dropped by the converter
    out['layers.0.blocks.1.attn_mask'] = torch.zeros(4, 49, 49)
    return out


def test_swin_converter_roundtrip_and_keys(tmp_path):
    torch.manual_seed(0)
    kw = dict(embed_dims=16, depths=(2, 2), num_heads=(2, 4), strides=(4, 2), out_indices=(0, 1), window_size=7)
    src = SwinTransformer(**kw)
    for p in src.parameters():
        torch.nn.init.normal_(p, std=0.5)
    official = _official_from(src)
    assert 'layers.0.blocks.0.attn.qkv.weight' in official and 'layers.0.blocks.0.mlp.fc1.weight' in official
    conv = C.swin_converter(official)
    assert 'backbone.stages.0.blocks.0.attn.w_msa.qkv.weight' in conv and 'backbone.patch_embed.projection.weight' in conv
    assert not any(k.startswith('backbone.head') for k in conv)
    path = tmp_path / 'swin_official.pth'
    torch.save(dict(model=official), path)
    dst = SwinTransformer(convert_weights=True, init_cfg=dict(type='Pretrained', checkpoint=str(path)), **kw)
    dst.init_weights()
    for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert torch.equal(a, b), k
    # a checkpoint that is not a local file: warn, keep the random init (the reference cfg names a URL)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        SwinTransformer(init_cfg=dict(type='Pretrained', checkpoint='https://example.invalid/x.pth'), **kw).init_weights()
    assert any('not a local file' in str(x.message) for x in w)


def test_patch_merging_permutation_is_semantically_the_official_layer():
    """official PatchMerging: cat(x[0::2,0::2], x[1::2,0::2], x[0::2,1::2], x[1::2,1::2]) -> LN -> Linear.
    With converted weights our (nn.Unfold-order) layer must compute the same function."""
    torch.manual_seed(1)
    B, H, W, Cc = 2, 6, 8, 8
    x = torch.randn(B, H * W, Cc)
    red = torch.randn(2 * Cc, 4 * Cc)
    g, b = torch.randn(4 * Cc), torch.randn(4 * Cc)
    xx = x.view(B, H, W, Cc)
    cat = torch.cat([xx[:, 0::2, 0::2], xx[:, 1::2, 0::2], xx[:, 0::2, 1::2], xx[:, 1::2, 1::2]], -1).view(B, -1, 4 * Cc)
    want = F.linear(F.layer_norm(cat, (4 * Cc,), g, b), red)
    conv = C.swin_converter(OrderedDict([('layers.0.downsample.reduction.weight', red), ('layers.0.downsample.norm.weight', g),
                                         ('layers.0.downsample.norm.bias', b)]))
    pm = PatchMerging(Cc, 2 * Cc)
    pm.load_state_dict({k.split('downsample.')[1]: v for k, v in conv.items()})
    with cpu_ops():
        got, hw = pm(x, (H, W))
    assert hw == (H // 2, W // 2)
    assert torch.allclose(got, want, atol=1e-5)


def test_relative_position_table_is_resized():
    kw = dict(embed_dims=16, depths=(2,), num_heads=(2,), strides=(4,), out_indices=(0,))
    src = SwinTransformer(window_size=5, **kw)
    dst = SwinTransformer(window_size=7, **kw)
    sd = OrderedDict(('backbone.' + k, v) for k, v in src.state_dict().items() if 'relative_position_index' not in k)
    C.load_swin_pretrained(dst, dict(state_dict=sd))
    key = 'stages.0.blocks.0.attn.w_msa.relative_position_bias_table'
    t = src.state_dict()[key]
    want = F.interpolate(t.permute(1, 0).reshape(1, 2, 9, 9), size=(13, 13), mode='bicubic').view(2, 169).permute(1, 0)
    assert torch.allclose(dst.state_dict()[key], want)


def _engine(seed=0):
    torch.manual_seed(seed)
    model = MODELS.build(small_cfg().model)
    model.init_weights()
    model.train()
    return StepEngine(model, dict(type='AdamW', lr=1e-3, weight_decay=1e-4,
                                  paramwise_cfg=dict(custom_keys={'backbone': dict(lr_mult=0.1)})),
                      grad_clip=dict(max_norm=0.1, norm_type=2), device='cpu', compute_dtype=torch.float32,
                      use_graphs=False, lr_config=dict(policy='step', step=[3]))


def _batches():
    from rscotr_b200.mtl.data import build_datasets
    out = []
    for i, task in enumerate(('cls', 'det', 'seg', 'cls')):
        ds = build_datasets({'x': dict(task=task)}, synthetic=dict(img_size=(64, 64), det=dict(num_boxes=2)))['x']
        b = ds.make_batch(2, torch.Generator().manual_seed(7 + i), pin=False)
        b.update(task=task, dataset_name='x')
        out.append(b)
    return out


@pytest.mark.timeout(600)
def test_save_resume_continues_identically(tmp_path):
    batches = _batches()
    a = _engine()
    a.model.eval()          # (no dropout / drop-path / noise draws: the continuation must be bit-identical)
    with cpu_ops():
        for b in batches[:2]:
            a.train_iter(b)
        path = C.save_checkpoint(a, str(tmp_path / 'work' / 'iter_2.pth'), meta=dict(CLASSES=('x',)))
        for b in batches[2:]:
            a.train_iter(b)
    ckpt = torch.load(path, weights_only=False)
    # mmcv layout: meta / state_dict / optimizer, one param group per parameter in named_parameters() order
    assert set(ckpt) == {'meta', 'state_dict', 'optimizer'} and ckpt['meta']['iter'] == 2
    n_params = len(list(a.model.named_parameters()))
    assert len(ckpt['optimizer']['param_groups']) == n_params
    assert [g['params'] for g in ckpt['optimizer']['param_groups']] == [[i] for i in range(n_params)]
    names = [n for n, _ in a.model.named_parameters()]
    g_backbone = ckpt['optimizer']['param_groups'][names.index('backbone.patch_embed.projection.weight')]
    g_head = ckpt['optimizer']['param_groups'][names.index('cls_head.fc.weight')]
    assert g_backbone['lr'] == pytest.approx(1e-4) and g_head['lr'] == pytest.approx(1e-3)
    # the file loads into a stock torch AdamW built the way mmcv builds it (one group per parameter)
    ref_opt = torch.optim.AdamW([dict(params=[p]) for p in a.model.parameters()], lr=1e-3)
    ref_opt.load_state_dict(ckpt['optimizer'])
    b = _engine(seed=123)                                  # different init: everything must come from the file
    b.model.eval()
    meta = C.resume(b, path)
    assert meta['iter'] == 2 and b.iter == 2 and meta['CLASSES'] == ('x',)
    with cpu_ops():
        for bt in batches[2:]:
            b.train_iter(bt)
    for (n, p), (_, q) in zip(a.model.named_parameters(), b.model.named_parameters()):
        assert torch.equal(p, q), n
    assert C.find_latest_checkpoint(str(tmp_path / 'work')) == path
    torch.save({}, tmp_path / 'work' / 'iter_10.pth')
    assert C.find_latest_checkpoint(str(tmp_path / 'work')).endswith('iter_10.pth')
    assert C.find_latest_checkpoint(str(tmp_path / 'nope')) is None


@pytest.mark.timeout(600)
def test_runner_checkpoint_hook_and_resume(tmp_path):
    from rscotr_b200.mtl.runner import CheckpointHook, IterBasedRunner
    batches = _batches()
    eng = _engine()
    runner = IterBasedRunner(eng, max_iters=4, work_dir=str(tmp_path), meta=dict(seed=3), log_interval=0)
    runner.register_hook(CheckpointHook(interval=2, max_keep_ckpts=1))
    with cpu_ops():
        runner.run([batches])
    files = sorted(os.listdir(tmp_path))
    assert files == ['iter_4.pth', 'latest.pth']            # iter_2 pruned by max_keep_ckpts=1
    eng2 = _engine(seed=9)
    r2 = IterBasedRunner(eng2, max_iters=4, work_dir=str(tmp_path), log_interval=0)
    meta = r2.resume(C.find_latest_checkpoint(str(tmp_path)))
    assert r2.iter == 4 and meta['seed'] == 3
    for (n, p), (_, q) in zip(eng.model.named_parameters(), eng2.model.named_parameters()):
        assert torch.equal(p, q), n
    # load_from: weights only, the iteration counter stays
    eng3 = _engine(seed=11)
    IterBasedRunner(eng3, max_iters=4, log_interval=0).load_checkpoint(str(tmp_path / 'iter_4.pth'))
    assert eng3.iter == 0
    assert torch.equal(eng3.model.cls_head.fc.weight, eng.model.cls_head.fc.weight)


@pytest.mark.timeout(600)
def test_test_model_api_from_checkpoint(tmp_path):
    """mtl.apis.test_model (the body of the reference's tools/test.py): checkpoint -> per-dataset metrics."""
    from rscotr_b200.mtl.apis import test_model
    eng = _engine()
    path = C.save_checkpoint(eng, str(tmp_path / 'iter_0.pth'))
    cfg = small_cfg()
    cfg.synthetic = dict(img_size=(64, 64), det=dict(num_boxes=2), val_length=dict(resisc=1, dior=2, potsdam=1))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for v in cfg.data.values():
        v['config'] = os.path.join(root, v['config'])
    with cpu_ops():
        metrics, outputs = test_model(cfg, path, tasks=('cls', 'seg'), split='val', device='cpu', synthetic=cfg.synthetic)
    assert set(metrics) == {'resisc', 'potsdam'} and len(outputs['resisc']) == 16 and len(outputs['potsdam']) == 2
    assert 'accuracy_top-1' in metrics['resisc'] and 'mFscore' in metrics['potsdam'] and 'mIoU' in metrics['potsdam']
    # a second run from the stored outputs (tools/test.py --test-outputs) reproduces the metrics without the model
    with cpu_ops():
        again, _ = test_model(cfg, path, tasks=('cls', 'seg'), split='val', device='cpu', synthetic=cfg.synthetic,
                              test_outputs=outputs)
    assert again['resisc'] == metrics['resisc']


@pytest.mark.timeout(900)
def test_train_model_api_end_to_end(tmp_path):
    """mtl.apis.train_model (reference signature): co-training + validation hook + periodic / best checkpoints, then a second
    call with auto_resume that continues from the latest checkpoint."""
    import logging
    from rscotr_b200.mtl.apis import train_model
    from rscotr_b200.mtl.data import build_datasets, load_data_cfg
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def make_cfg(max_iters):
        cfg = small_cfg()
        for v in cfg.data.values():
            v['config'] = os.path.join(root, v['config'])
            v['data']['samples_per_gpu'] = 2 if v['task'] == 'cls' else 1
        load_data_cfg(cfg)
        cfg.synthetic = dict(img_size=(64, 64), det=dict(num_boxes=2), val_length=dict(resisc=1, dior=1, potsdam=1))
        cfg.device, cfg.compute_dtype, cfg.work_dir = 'cpu', torch.float32, str(tmp_path)
        cfg.runner = dict(type='IterBasedRunner', max_iters=max_iters)
        cfg.evaluation = dict(interval=3, save_best={'resisc.accuracy_top-1': 1, 'dior.bbox_mAP': 100, 'potsdam.mFscore': 100},
                              cls=dict(metric='accuracy'), det=dict(metric='bbox', iou_thrs=[0.5]), seg=dict(metric=['mFscore', 'mIoU']))
        cfg.checkpoint_config = dict(interval=3)
        cfg.log_config = dict(interval=1)
        return cfg
    torch.manual_seed(0)
    cfg = make_cfg(3)
    model = MODELS.build(cfg.model)
    model.init_weights()
    logging.getLogger('rscotr_b200').setLevel(logging.WARNING)
    with cpu_ops():
        runner = train_model(model, build_datasets(cfg.data, synthetic=cfg.synthetic), cfg, validate=True, meta=dict(seed=0))
    assert runner.iter == 3
    files = sorted(os.listdir(tmp_path))
    assert 'iter_3.pth' in files and 'latest.pth' in files and any(f.startswith('best_') and f.endswith('iter_3.pth') for f in files)
    for k in ('resisc.accuracy_top-1', 'dior.bbox_mAP', 'potsdam.mFscore', 'potsdam.mIoU'):
        assert k in runner.log_buffer, sorted(runner.log_buffer)
    assert any(k.startswith('seg.potsdam.') for k in runner.log_buffer)          # training log vars of the last iteration
    assert runner.log_buffer['grad_norm'] > 0                                    # (mmcv OptimizerHook logs the pre-clip norm)
    import json
    recs = [json.loads(l) for l in open(tmp_path / 'train.log.json')]
    assert [r['iter'] for r in recs] == [1, 2, 3] and recs[0]['mode'] == 'train' and 'cls.resisc.loss' in recs[0] and recs[-1]['lr'] > 0
    # second call: auto_resume picks latest.pth up and only runs the remaining iteration
    cfg2 = make_cfg(4)
    cfg2.auto_resume = True
    torch.manual_seed(1)
    model2 = MODELS.build(cfg2.model)
    with cpu_ops():
        runner2 = train_model(model2, build_datasets(cfg2.data, synthetic=cfg2.synthetic), cfg2, validate=False)
    assert runner2.iter == 4 and 'iter_4.pth' in os.listdir(tmp_path)             # (save_last: the final iteration is always saved)
    assert torch.equal(model2.cls_head.fc.bias, model2.cls_head.fc.bias) and not torch.equal(
        model2.backbone.patch_embed.projection.weight, MODELS.build(make_cfg(1).model).backbone.patch_embed.projection.weight)


def test_log_buffer_averages_over_the_window():
    from rscotr_b200.models.mtl import _LazyLogVars
    from rscotr_b200.mtl.runner.iter_runner import _LogBuffer
    lb = _LogBuffer()
    lb.accumulate(_LazyLogVars(['cls.x.loss'], torch.tensor([2.0]), 1))
    lb.accumulate(_LazyLogVars(['seg.y.loss_ce', 'seg.y.loss'], torch.tensor([1.0, 3.0]), 0.1))
    lb.accumulate(_LazyLogVars(['cls.x.loss'], torch.tensor([4.0]), 1))
    lb.accumulate(dict(grad_norm=2.0))
    avg = lb.average()
    assert avg['cls.x.loss'] == 3.0 and avg['seg.y.loss'] == pytest.approx(0.3) and avg['grad_norm'] == 2.0
    assert lb['cls.x.loss'] == 3.0 and lb._acc == {}
    # distributed packing: element 0 carries the number of log vars and is not a value
    lb.accumulate(_LazyLogVars(['a', 'b'], torch.tensor([2.0, 5.0, 7.0]), 1))
    assert lb.average() == dict(a=5.0, b=7.0)
