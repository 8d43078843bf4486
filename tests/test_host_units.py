"""CPU unit tests of the small host-side helpers the step relies on: shape-keyed geometry cache,
packed loss dicts, batched stochastic-depth draws, the static denoising-target table, flat buffer
layout of the step engine."""
import torch

import rscotr_b200.models  # noqa: F401
from rscotr_b200.models.bricks import DropPath, GeomCache, PackedLosses, draw_drop_paths
from rscotr_b200.models.det_head import DINOHead


def test_geom_cache_evicts_oldest_and_runs_without_grad():
    c = GeomCache(capacity=2)
    calls = []

    def make(k):
        def f():
            calls.append(k)
            assert not torch.is_grad_enabled()
            return torch.full((1,), float(k))
        return f

    a = c.get(('s', 1), make(1))
    assert c.get(('s', 1), make(1)) is a and calls == [1]
    c.get(('s', 2), make(2))
    c.get(('s', 3), make(3))                       # evicts key 1 (insertion order)
    assert set(c.entries) == {('s', 2), ('s', 3)}
    c.get(('s', 1), make(1))
    assert calls == [1, 2, 3, 1]


def test_packed_losses_behaves_like_the_reference_loss_dict():
    packed = torch.tensor([1., 2., 3.], requires_grad=True)
    d = PackedLosses(['loss_cls', 'loss_bbox', 'loss_iou'], packed * 2)
    assert list(d.keys()) == ['loss_cls', 'loss_bbox', 'loss_iou'] and len(d) == 3 and 'loss_bbox' in d
    assert float(d['loss_bbox']) == 4. and float(d.get('loss_iou')) == 6. and d.get('nope') is None
    assert [float(v) for v in d.values()] == [2., 4., 6.]
    assert {k: float(v) for k, v in d.items()} == dict(loss_cls=2., loss_bbox=4., loss_iou=6.)
    # per-key tensors are views of the packed tensor: gradients flow back to it
    sum(v for k, v in d.items() if 'loss' in k).backward()
    assert torch.equal(packed.grad, torch.full((3,), 2.))


def test_drop_path_forced_mask_and_batched_draws():
    x = torch.arange(12.).view(3, 2, 2)
    m = DropPath(0.5).train()
    m.forced_mask = torch.tensor([0., 2., 2.])
    assert torch.equal(m(x), x * m.forced_mask.view(3, 1, 1))
    assert torch.equal(m.scale_vec(x), m.forced_mask)
    idn = torch.ones_like(x)
    assert torch.equal(m.add_to(idn, x), idn + x * m.forced_mask.view(3, 1, 1))
    # eval / p = 0: identity, no tensor made
    assert DropPath(0.5).eval().scale_vec(x) is None and DropPath(0.).train().scale_vec(x) is None
    # one batched draw for a list of modules: factors are 0 or 1/keep, consumed exactly once
    mods = [DropPath(0.).train(), DropPath(0.25).train(), DropPath(0.5).train(), m]
    torch.manual_seed(0)
    draw_drop_paths(mods, 64, 'cpu', torch.float32)
    assert mods[0].drawn is None and mods[3].drawn is None        # p = 0 and forced masks are skipped
    for mod in mods[1:3]:
        keep = 1 - mod.drop_prob
        s = mod.scale_vec(torch.zeros(64, 1))
        assert s.shape == (64,) and bool(((s == 0) | ((s - 1 / keep).abs() < 1e-6)).all())
        assert mod.drawn is None
    # the draw keeps roughly `keep` of the samples
    big = [DropPath(0.3).train()]
    draw_drop_paths(big, 20000, 'cpu', torch.float32)
    assert abs(float((big[0].drawn > 0).float().mean()) - 0.7) < 0.02


def test_dn_assign_table_layout():
    # 2 images with 2 and 1 boxes, 3 groups, pad_size = 3 groups * 2 * max_gt(2) = 12 -> single = 4
    t = DINOHead.dn_assign_table([2, 1], 3, 12)
    assert t[0] == [0, 1, -1, -1] * 3
    assert t[1] == [2, -1, -1, -1] * 3
    assert DINOHead.dn_assign_table([0, 0], 2, 0) == [[], []]
    assert DINOHead.dn_assign_table([], 0, 0) == []


def test_step_engine_flat_layout_and_cpu_prefetch_noop():
    from rscotr_b200.config import MODELS
    from rscotr_b200.mtl.engine import StepEngine
    from tests.test_host_model import small_cfg
    torch.manual_seed(0)
    model = MODELS.build(small_cfg().model)
    model.init_weights()
    eng = StepEngine(model, dict(type='AdamW', lr=1e-3, weight_decay=1e-4), device='cpu',
                     compute_dtype=torch.float32, use_graphs=False)
    # every parameter is a view of the flat buffer, spans are 64-element aligned and do not overlap
    base = eng.flat_param.data_ptr()
    spans = sorted(((p.data_ptr() - base) // 4, p.numel()) for p in model.parameters())
    end = 0
    for off, n in spans:
        assert off % 64 == 0 and off >= end
        end = off + n
    assert end <= eng.flat_param.numel() == eng.flat_grad.numel()
    eng.prefetch(dict(task='cls'))                 # no copy stream on CPU: must be a no-op, not an error


def test_step_engine_pairs_sibling_linears():
    """StepEngine._pair_linears: sampling_offsets / attention_weights of every MultiScaleDeformableAttention sit back to
    back in the flat buffers ([w1 | w2 | b1 | b2]), so ops.linear_pair_views returns stacked VIEWS (one GEMM serves both
    layers); values and gradient aliasing checked here, the GEMM itself in the -m gpu tests."""
    from rscotr_b200 import ops
    from rscotr_b200.config import MODELS
    from rscotr_b200.models.bricks import MultiScaleDeformableAttention
    from rscotr_b200.mtl.engine import StepEngine
    from tests.test_host_model import small_cfg
    torch.manual_seed(0)
    model = MODELS.build(small_cfg().model)
    model.init_weights()
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    eng = StepEngine(model, dict(type='AdamW', lr=1e-3, weight_decay=1e-4), device='cpu',
                     compute_dtype=torch.float32, use_graphs=False)
    for n, p in model.named_parameters():              # the re-ordering moved no values
        assert torch.equal(p.detach(), before[n]), n
    mods = [m for m in model.modules() if isinstance(m, MultiScaleDeformableAttention)]
    assert mods
    for m in mods:
        pv = ops.linear_pair_views(m.sampling_offsets, m.attention_weights)
        assert pv is not None and pv['w_lp'] is None          # (the bf16 shadow exists on CUDA only)
        n1, n2 = m.sampling_offsets.out_features, m.attention_weights.out_features
        assert pv['w'].shape == (n1 + n2, m.embed_dims) and pv['b'].shape == (n1 + n2,)
        assert torch.equal(pv['w'], torch.cat([m.sampling_offsets.weight.detach(), m.attention_weights.weight.detach()]))
        assert torch.equal(pv['b'], torch.cat([m.sampling_offsets.bias.detach(), m.attention_weights.bias.detach()]))
        pv['gw'].fill_(1.0)
        pv['gb'].fill_(2.0)
        for lin in (m.sampling_offsets, m.attention_weights):
            assert float(eng.grad_view(lin.weight).min()) == 1.0 and float(eng.grad_view(lin.bias).min()) == 2.0
        assert m._pair_views() is not None
    eng.flat_grad.zero_()
    # parameters that are NOT adjacent (any other two Linears) give no views
    assert ops.linear_pair_views(mods[0].value_proj, mods[0].output_proj) is None


def test_eval_hook_on_synthetic_loaders():
    """MultiDatasetsEvalHook / single_gpu_test over the synthetic val loaders (what a GPU box without data runs)."""
    from rscotr_b200.config import MODELS
    from rscotr_b200.mtl.data import build_datasets
    from rscotr_b200.mtl.data.synthetic import _SyntheticLoader
    from rscotr_b200.mtl.engine.test import single_gpu_test
    from tests.cpu_ops_shim import cpu_ops
    from tests.test_host_model import small_cfg
    torch.manual_seed(0)
    model = MODELS.build(small_cfg().model)
    model.init_weights()
    sets = build_datasets({'c': dict(task='cls'), 'd': dict(task='det'), 's': dict(task='seg')},
                          split='val', synthetic=dict(img_size=(64, 64), det=dict(num_boxes=2)))
    loaders = {k: _SyntheticLoader(v, 2, 2, seed=3, pin=False) for k, v in sets.items()}
    with cpu_ops():
        results = single_gpu_test(model, loaders)
    assert len(results['c']) == 4 and results['c'][0].shape == (45,)
    assert len(results['d']) == 4 and len(results['d'][0]) == 20 and results['d'][0][0].shape[1] == 5
    assert len(results['s']) == 4 and results['s'][0].shape == (64, 64)
    acc = sets['c'].evaluate(results['c'], metric='accuracy')
    assert set(acc) == {'accuracy_top-1', 'accuracy_top-5'}
    ap = sets['d'].evaluate(results['d'], metric='bbox', iou_thrs=[0.5])
    assert -1 <= ap['bbox_mAP'] <= 1 and ap['bbox_mAP_75'] == -1.0
    seg = sets['s'].evaluate(results['s'], metric=['mFscore', 'mIoU'])
    assert 0 <= seg['mIoU'] <= 1 and 'mFscore' in seg


def test_mlvl_cls_head_in_the_cotraining_model():
    """the reference's second co-training config (MTL_swin-t-p4-w7_1x1_resisc&dior&potsdam.py:54-69): cls goes through
    neck + SHARED encoder + MlvlClsHead; gradients reach all of them and the engine's cls range grows accordingly."""
    import pytest
    from rscotr_b200.config import MODELS
    from rscotr_b200.mtl.data import build_datasets
    from rscotr_b200.mtl.engine import StepEngine
    from tests.cpu_ops_shim import cpu_ops
    from tests.test_host_model import small_cfg
    for scheme in (2, 8):
        cfg = small_cfg()
        cfg.model.cls_head = dict(
            type='MlvlClsHead', scheme=scheme, num_classes=45, in_channels=256,
            pixel_decoder=dict(type='MlvlClsPixelDecoder', num_encoder_levels=4, num_outs=4),
            loss=dict(type='LabelSmoothLoss', label_smooth_val=0.1, mode='original'), cal_acc=False,
            init_cfg=[dict(type='TruncNormal', layer='Linear', std=0.02, bias=0.), dict(type='Constant', layer='LayerNorm', val=1., bias=0.)])
        torch.manual_seed(0)
        model = MODELS.build(cfg.model)
        model.init_weights()
        model.train()
        head = model.cls_head
        assert abs(float(head.fc.weight.detach().std()) - 0.02) < 0.002      # the config's TruncNormal(std=.02), not LinearClsHead's .01
        if scheme == 8:
            assert head.out_proj.weight.shape == (1, 4)
        eng = StepEngine(model, dict(type='AdamW', lr=1e-3, weight_decay=1e-4), device='cpu', compute_dtype=torch.float32,
                         use_graphs=False)
        ds = build_datasets({'x': dict(task='cls')}, synthetic=dict(img_size=(64, 64)))['x']
        batch = ds.make_batch(2, torch.Generator().manual_seed(1), pin=False)
        batch.update(task='cls', dataset_name='x')
        with cpu_ops():
            out = model.train_step(dict(batch), None)
            out['loss'].backward()
            eng._collect_grads()
        def gnorm(p):
            return float(eng.grad_view(p).norm())
        # (level_encoding only feeds the offset / attention-weight projections, which start at zero weight: no grad yet)
        assert gnorm(model.neck.convs[0].conv.weight) > 0
        assert gnorm(model.shared_encoder.layers[0].ffns[0].layers[1].weight) > 0 and gnorm(head.fc.weight) > 0
        assert gnorm(model.bbox_head.fc_cls.weight if hasattr(model.bbox_head, 'fc_cls') else next(model.bbox_head.parameters())) == 0
        assert out['log_vars']['cls.x.loss'] == pytest.approx(float(out['loss']))
        with cpu_ops():
            model.eval()
            pred = model(task='cls', img=[batch['img']], img_metas=[batch['img_metas']], return_loss=False)
        assert len(pred) == 2 and pred[0].shape == (45,) and abs(float(pred[0].sum()) - 1) < 1e-4


def test_detection_only_configuration():
    """BASELINE configs[3]: the MTL schema with only the det head, Swin-S depths, constant strategy."""
    import os
    from rscotr_b200.config import Config, MODELS
    from rscotr_b200.mtl.data import build_datasets, build_multidataloader, load_data_cfg
    from rscotr_b200.mtl.engine import StepEngine
    from tests.cpu_ops_shim import cpu_ops
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = Config.fromfile(os.path.join(root, 'configs/multi/det_only_swin-s_800.py'))
    assert cfg.model.backbone.depths == [2, 2, 18, 2] and cfg.model.cls_head is None and list(cfg.data.keys()) == ['dior']
    m = cfg.model
    m.backbone.depths = [2, 2, 2, 2]                       # (CPU-sized; the graded shape runs on the GPU)
    m.backbone.drop_path_rate = 0.0
    m.bbox_head.num_query, m.bbox_head.dn_cfg.group_cfg.num_dn_queries = 30, 10
    m.shared_encoder.num_layers, m.bbox_head.transformer.decoder.num_layers = 2, 2
    torch.manual_seed(0)
    model = MODELS.build(m)
    model.init_weights()
    model.train()
    assert model.cls_head is None and model.seg_head is None
    for v in cfg.data.values():
        v['config'] = os.path.join(root, v['config'])
    load_data_cfg(cfg)
    cfg.synthetic = dict(img_size=(64, 64), det=dict(num_boxes=2), length=dict(dior=3))
    loader = build_multidataloader(cfg, False, build_datasets(cfg.data, synthetic=cfg.synthetic))
    eng = StepEngine(model, dict(cfg.optimizer), grad_clip=dict(cfg.optimizer_config.grad_clip), device='cpu',
                     compute_dtype=torch.float32, use_graphs=False)
    it = iter(loader)
    with cpu_ops():
        for _ in range(2):
            b = next(it)
            assert b['task'] == 'det' and b['dataset_name'] == 'dior' and b['img'].shape[0] == 2
            out = eng.train_iter(b)
    assert float(out['loss']) > 0 and any(k.startswith('det.dior.') for k in out['log_vars'])


def test_single_task_classifier():
    """BASELINE configs[0]: Swin-T single-task classification, 2 x 3 x 256 x 256, on CPU: the reference's
    configs/cls file builds unmodified (when mounted), the repo's own config describes the same model / optimizer,
    one engine step runs, the param groups follow mmcv's norm / bias / custom-key rules."""
    import os
    import pytest
    from rscotr_b200.config import Config, MODELS
    from rscotr_b200.mtl.data import build_datasets
    from rscotr_b200.mtl.engine import StepEngine
    from rscotr_b200.mtl.utils.optimizer import param_settings
    from tests.cpu_ops_shim import cpu_ops
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = Config.fromfile(os.path.join(root, 'configs/cls/swin-tiny_resisc.py'))
    ref_path = '/root/reference/configs/cls/swin-tiny_1xb16_resisc.py'
    if os.path.exists(ref_path):
        ref = Config.fromfile(ref_path)
        strip = lambda d: {k: v for k, v in dict(d).items() if k != 'init_cfg'}
        assert strip(ref.model.backbone) == strip(cfg.model.backbone) and dict(ref.model.head) == dict(cfg.model.head)
        assert dict(ref.optimizer) == dict(cfg.optimizer) and dict(ref.optimizer_config) == dict(cfg.optimizer_config)
        assert MODELS.build(ref.model).__class__.__name__ == 'ImageClassifier'
    torch.manual_seed(0)
    model = MODELS.build(cfg.model)
    model.init_weights()
    model.train()
    assert sum(p.numel() for p in model.parameters()) == 27_553_959          # Swin-T + Linear(768, 45)
    settings = {n: (lr, wd) for n, _, lr, wd in param_settings(model, dict(lr=2e-4, weight_decay=1e-4), cfg.optimizer.paramwise_cfg)}
    assert settings['backbone.stages.0.blocks.0.norm1.weight'][1] == 0.0                     # norm_decay_mult
    assert settings['backbone.stages.0.blocks.0.attn.w_msa.qkv.bias'][1] == 0.0              # bias_decay_mult
    assert settings['backbone.stages.0.blocks.0.attn.w_msa.relative_position_bias_table'][1] == 0.0   # custom key
    assert settings['backbone.stages.0.blocks.0.attn.w_msa.qkv.weight'] == (2e-4, 1e-4)
    eng = StepEngine(model, dict(cfg.optimizer), grad_clip=dict(cfg.optimizer_config.grad_clip), device='cpu',
                     compute_dtype=torch.float32, use_graphs=False)
    ds = build_datasets({'resisc': dict(task='cls')}, synthetic=dict(cfg.synthetic))['resisc']
    batch = ds.make_batch(2, torch.Generator().manual_seed(0), pin=False)
    assert batch['img'].shape == (2, 3, 256, 256)
    batch.update(task='cls', dataset_name='resisc')
    with cpu_ops():
        out = eng.train_iter(batch)
        logs = dict(out['log_vars'].items())
        assert set(logs) == {'cls.resisc.loss'} and logs['cls.resisc.loss'] == pytest.approx(float(out['loss'].detach()), rel=1e-5)
        model.eval()
        pred = model(img=batch['img'], img_metas=batch['img_metas'], return_loss=False)
    assert len(pred) == 2 and pred[0].shape == (45,)


def test_logistic_gelu_fit_is_within_bf16_resolution():
    """the opt-in act=2 of rsc_bias_act_* (csrc/fused_ew.cu::gelu_sig_cdf), restated in float32 numpy: error of the value
    and of the derivative against the exact erf form, behaviour for large |x| (the clamp), relative accuracy of the left tail."""
    import numpy as np
    from scipy.special import erf
    x = np.concatenate([np.linspace(-40, 40, 400001), [-1e4, 1e4]]).astype(np.float32)
    x2 = np.minimum(x * x, np.float32(49.0))
    p = x2 * np.float32(0.0010142630552444944) + np.float32(-0.10677572400272597)
    p = x2 * p + np.float32(-2.3011213394566354)
    with np.errstate(over='ignore'):
        cdf = (np.float32(1) / (np.float32(1) + np.exp2(x * p).astype(np.float32))).astype(np.float32)
    xd = x.astype(np.float64)
    phi = 0.5 * (1 + erf(xd / np.sqrt(2)))
    assert np.max(np.abs(x * cdf - xd * phi)) < 3e-5
    pdf = np.exp(-0.5 * xd * xd) / np.sqrt(2 * np.pi)
    assert np.max(np.abs((cdf + xd * pdf) - (phi + xd * pdf))) < 6e-5
    assert cdf[-1] == 1.0 and cdf[-2] == 0.0 and np.all(np.isfinite(cdf))
    tail = (xd > -5) & (xd < -3)
    assert np.max(np.abs(cdf[tail] - phi[tail]) / phi[tail]) < 0.05         # a few percent RELATIVE down to x = -5 (Phi = 3e-7)


def test_graph_cache_is_bounded():
    """StepEngine._graph_slot: at most `max_graphs` captured signatures; a new signature replaces a captured one only when
    that one has been idle for `graph_idle_iters` iterations, otherwise it runs eagerly (None)."""
    from rscotr_b200.mtl.engine import StepEngine
    eng = StepEngine(torch.nn.Linear(4, 2), dict(type='SGD', lr=0.1), device='cpu', compute_dtype=torch.float32, use_graphs=False)
    eng.max_graphs, eng.graph_idle_iters = 2, 10
    a = eng._graph_slot('a')
    assert a == dict(eager=0, last=0) and eng._graph_slot('a') is a
    a['gA'] = object()                                    # "captured"
    eng.iter = 3
    b = eng._graph_slot('b')
    b['gA'] = object()
    eng.iter = 5
    assert eng._graph_slot('c') is None and set(eng._graphs) == {'a', 'b'}      # full, nobody idle long enough
    assert eng._graph_slot('b')['last'] == 5
    eng.iter = 12
    c = eng._graph_slot('c')                               # 'a' idle since iteration 0 -> evicted
    assert c is not None and set(eng._graphs) == {'b', 'c'}
    eng.iter = 13
    assert eng._graph_slot('d') is not None               # 'c' is not captured yet: it does not count against the bound


def test_batch_augments_device_formulation():
    """Augments.call_on_device (graph-capturable Mixup / CutMix): same distributions and formulas as the host-RNG path."""
    from rscotr_b200.models.cls_head import Augments
    aug = Augments([dict(type='BatchMixup', alpha=0.8, num_classes=5, prob=0.5), dict(type='BatchCutMix', alpha=1.0, num_classes=5, prob=0.3)])
    assert len(aug.augments) == 3                                           # + Identity with the remaining 0.2
    torch.manual_seed(0)
    B, H, W = 6, 24, 32
    img = torch.arange(B, dtype=torch.float32).view(B, 1, 1, 1).expand(B, 3, H, W).contiguous() + 10     # image b is the constant 10 + b
    label = torch.arange(B) % 5
    kinds = dict(mixup=0, cutmix=0, none=0)
    lams = []
    for _ in range(600):
        out, soft = aug.call_on_device(img, label)
        assert out.shape == img.shape and soft.shape == (B, 5) and torch.allclose(soft.sum(1), torch.ones(B), atol=1e-6)
        # recover the permutation partner and the mixing weight of sample 0
        per_pixel = out[0, 0]
        vals = per_pixel.unique()
        if len(vals) == 1 and float(vals[0]) == 10.0 and float(soft[0].max()) == 1.0:
            kinds['none'] += 1
            continue
        if len(vals) == 2 or (len(vals) == 1 and float(vals[0]) == float(int(vals[0])) and float(soft[0].max()) < 1.0):
            # CutMix: every pixel is exactly one of the two source images; label weight = 1 - pasted area / image area
            kinds['cutmix'] += 1
            src = out[:, 0] - 10
            partner_pixels = (src != torch.arange(B, dtype=torch.float32).view(B, 1, 1)).float().mean((1, 2))
            lam = 1 - partner_pixels
            for b in range(B):
                if partner_pixels[b] > 0:      # (a sample whose partner is itself is unchanged)
                    assert abs(float(soft[b, label[b]]) - float(lam[b])) < 1e-5 or float(soft[b].max()) == 1.0
            rows = (src[0] != 0).any(1).nonzero().flatten()
            cols = (src[0] != 0).any(0).nonzero().flatten()
            if len(rows):                       # the pasted region is ONE axis-aligned rectangle
                box = src[0][rows[0]:rows[-1] + 1, cols[0]:cols[-1] + 1]
                assert (box != 0).all() and (src[0] != 0).sum() == box.numel()
            continue
        kinds['mixup'] += 1
        assert len(vals) == 1                                               # a constant blend of two constant images
        lams.append(float(soft[0].max()))
    n = sum(kinds.values())
    # (sample 0 is its own partner with probability 1/6, which hides an augmentation as 'none')
    assert abs(kinds['mixup'] / n - 0.5 * 5 / 6) < 0.07 and abs(kinds['cutmix'] / n - 0.3 * 5 / 6) < 0.07, kinds
    assert 0.5 <= min(lams) and max(lams) <= 1.0 and np_std(lams) > 0.05


def np_std(v):
    import numpy as np
    return float(np.std(v))


def test_lr_policies_match_mmcv_closed_forms():
    import math
    from rscotr_b200.mtl.engine import StepEngine
    def eng(cfg, max_iters=100):
        e = StepEngine(torch.nn.Linear(2, 2), dict(type='SGD', lr=0.5), device='cpu', compute_dtype=torch.float32, use_graphs=False,
                       lr_config=cfg)
        e.max_iters = max_iters
        return e
    e = eng(dict(policy='step', step=[10, 20], gamma=0.1))
    assert [e.lr_scale(i) for i in (0, 9, 10, 19, 20, 99)] == pytest_approx([1, 1, .1, .1, .01, .01])
    e = eng(dict(policy='poly', power=1.0, min_lr=0.0, warmup='linear', warmup_iters=10, warmup_ratio=1e-6))
    assert e.lr_scale(50) == 0.5 and e.lr_scale(0) == pytest_approx(1e-6) and e.lr_scale(5) == pytest_approx(0.95 * (1 - 0.5 * (1 - 1e-6)))
    e = eng(dict(policy='CosineAnnealing', min_lr_ratio=1e-2, warmup='exp', warmup_iters=4, warmup_ratio=0.1))
    assert e.lr_scale(100) == pytest_approx(1e-2) and e.lr_scale(50) == pytest_approx(0.01 + 0.5 * 0.99)
    assert e.lr_scale(2) == pytest_approx((0.01 + 0.495 * (math.cos(math.pi * 0.02) + 1)) * 0.1 ** 0.5)
    e = eng(dict(policy='fixed', warmup='constant', warmup_iters=3, warmup_ratio=0.25))
    assert [e.lr_scale(i) for i in range(5)] == [0.25, 0.25, 0.25, 1.0, 1.0]
    # the engine applies it to the optimizer's groups
    e = eng(dict(policy='step', step=[1]))
    e.iter = 1
    e._update_lr()
    assert e.optimizer.param_groups[0]['lr'] == pytest_approx(0.05)
    import pytest
    with pytest.raises(KeyError):
        eng(dict(policy='cyclic')).lr_scale(0)


def pytest_approx(v):
    import pytest
    return pytest.approx(v, rel=1e-9, abs=1e-15)


def test_compact_attention_mask_is_equivalent(monkeypatch):
    """RSC_COMPACT_ATTN_MASK=1 (one (B,1,Q,K) mask broadcast over the heads) gives the same seg losses and gradients as the
    reference's per-head (B*heads, Q, K) mask."""
    import rscotr_b200.models.seg_head as sh
    from rscotr_b200.config import MODELS
    from rscotr_b200.mtl.data import build_datasets
    from tests.cpu_ops_shim import cpu_ops
    from tests.test_host_model import small_cfg
    ds = build_datasets({'x': dict(task='seg')}, synthetic=dict(img_size=(64, 64)))['x']
    batch = ds.make_batch(2, torch.Generator().manual_seed(2), pin=False)
    batch.update(task='seg', dataset_name='x')
    outs = []
    for compact in (False, True):
        monkeypatch.setattr(sh, '_COMPACT_ATTN_MASK', compact)
        torch.manual_seed(0)
        model = MODELS.build(small_cfg().model)
        model.init_weights()
        model.train()
        with cpu_ops():
            out = model.train_step(dict(batch), None)
            out['loss'].backward()
        outs.append((float(out['loss']), model.seg_head.query_feat.weight.grad.clone(),
                     model.backbone.patch_embed.projection.weight.grad.clone()))
    assert outs[0][0] == outs[1][0]
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-5, atol=1e-8) and torch.allclose(outs[0][2], outs[1][2], rtol=1e-5, atol=1e-8)


def test_constant_attention_bias_is_equivalent(monkeypatch):
    """RSC_CONST_ATTN_BIAS=1: the DINO denoising mask as a cached additive bias gives the same det losses / gradients as the
    boolean mask, and the cache is actually hit (the marker survives the per-attention shallow copies)."""
    import rscotr_b200.models.bricks as br
    from rscotr_b200.config import MODELS
    from rscotr_b200.mtl.data import build_datasets
    from oracle import heads as oh
    from tests.cpu_ops_shim import cpu_ops
    from tests.test_host_model import small_cfg
    ds = build_datasets({'x': dict(task='det')}, synthetic=dict(img_size=(64, 64), det=dict(num_boxes=2)))['x']
    batch = ds.make_batch(2, torch.Generator().manual_seed(2), pin=False)
    batch.update(task='det', dataset_name='x')
    noise = oh.cdn_noise(batch['gt_labels'], num_dn=10, generator=torch.Generator().manual_seed(5))
    outs, caches = [], []
    for const in (False, True):
        monkeypatch.setattr(br, '_CONST_ATTN_BIAS', const)
        torch.manual_seed(0)
        model = MODELS.build(small_cfg().model)
        model.init_weights()
        model.train()
        model.bbox_head.dn_generator.forced_noise = noise
        with cpu_ops():
            out = model.train_step(dict(batch), None)
            out['loss'].backward()
        outs.append((float(out['loss']), model.bbox_head.transformer.decoder.layers[0].attentions[0].attn.in_proj_weight.grad.clone()))
        masks = [v['attn_mask'] for v in model.bbox_head.dn_generator._geom.entries.values()]
        caches.append(sum(len(m._rsc_add) for m in masks))
    assert outs[0][0] == __import__('pytest').approx(outs[1][0], rel=1e-6)
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-5, atol=1e-8)
    assert masks and caches[0] == 0 and caches[1] >= 1
