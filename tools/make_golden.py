"""Generates tests/golden/* by RUNNING THE REFERENCE'S OWN CODE (read-only, from /root/reference) for the two
pieces of the hot path that are importable in this container once their un-installable imports are stubbed:

  * mtl/data/iteration_strategies.py       -- needs `omegaconf`; a 20-line attribute-dict shim
    stands in for OmegaConf.structured / merge (configuration plumbing only, no arithmetic);
  * models/multi/bbox_head/query_denoising.py::CdnQueryGenerator (row a15) -- imports two mmdet helpers
    (`bbox_xyxy_to_cxcywh`, `inverse_sigmoid`), restated here from mmdet 2.25.1 (3 lines each), and hard-codes
    `.cuda()`; `Tensor.cuda` / `.to('cuda')` are no-ops for the duration of the call.  The RNG draws the reference
    makes (rand_like / randint_like) are recorded next to its outputs so that a test can feed the very same noise
    to this repo's generator and to the oracle.

Nothing under /root/reference is modified or copied; only inputs, recorded draws and outputs are stored.
The fixtures travel with the repo; the GPU box never needs /root/reference.

    python tools/make_golden.py            # rewrites tests/golden/
"""
import dataclasses
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('RSC_REFERENCE', '/root/reference')
OUT = os.path.join(ROOT, 'tests', 'golden')


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


# ----------------------------------------------------------------------------- omegaconf shim
class _AttrDict(dict):
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


def _structured(obj):
    if isinstance(obj, type):
        obj = obj.__new__(obj)
        fields = {f.name: (f.default if f.default is not dataclasses.MISSING else '???') for f in dataclasses.fields(obj)}
        return _AttrDict(fields)
    return _AttrDict(dataclasses.asdict(obj))


def install_omegaconf_shim():
    m = types.ModuleType('omegaconf')
    m.MISSING = '???'

    class OmegaConf:
        structured = staticmethod(_structured)

        @staticmethod
        def merge(a, b):
            out = _AttrDict(a)
            out.update(dict(b) if b is not None else {})
            return out

        @staticmethod
        def create(d=None):
            return _AttrDict(d or {})
    m.OmegaConf = OmegaConf
    sys.modules['omegaconf'] = m


def golden_iteration_strategies():
    install_omegaconf_shim()
    ref = load('mtl/data/iteration_strategies.py', 'ref_iteration_strategies')
    lens = dict(resisc=7, dior=3, potsdam=5)

    sizes = dict(resisc=70, dior=9, potsdam=20)      # dataset sizes, deliberately not proportional to the loader lengths

    class FakeLoader:          # the strategies only use len(dataloader) and, for size-proportional, len(dataloader.dataset)
        def __init__(self, n, m):
            self.n, self.dataset = n, list(range(m))

        def __len__(self):
            return self.n
    loaders = {k: FakeLoader(v, sizes[k]) for k, v in lens.items()}
    cases = [('ConstantIterationStrategy', dict(idx=1)), ('RoundRobinIterationStrategy', dict(start_idx=0)),
             ('RoundRobinIterationStrategy', dict(start_idx=2)),
             ('RepeatedSequenceIterationStrategy', dict(sequence=[0, 0, 2, 1])),
             ('RandomIterationStrategy', dict()), ('WeightedRandomIterationStrategy', dict(p=[6, 1, 3])),
             ('SizeProportionalIterationStrategy', dict())]
    out = []
    for name, kw in cases:
        cls = getattr(ref, name)
        np.random.seed(123)
        try:
            if 'sequence' in kw:           # positional constructor arguments in the reference
                strat = cls(loaders, kw['sequence'])
            elif 'p' in kw:
                strat = cls(loaders, kw['p'])
            else:
                strat = cls.from_params(loaders, **kw)
        except Exception as e:      # a strategy whose constructor needs more of omegaconf than the shim offers
            out.append(dict(strategy=name, kwargs=kw, error='%s: %s' % (type(e).__name__, e)))
            continue
        seq = [int(strat()) for _ in range(40)]
        out.append(dict(strategy=name, kwargs=kw, numpy_seed=123, loader_lengths=lens, dataset_sizes=sizes,
                        should_exhaust_all_iterators=bool(strat.should_exhaust_all_iterators), sequence=seq))
    json.dump(dict(source='mtl/data/iteration_strategies.py (reference, run in place)', cases=out),
              open(os.path.join(OUT, 'reference_iteration_strategies.json'), 'w'), indent=1)
    return out


# ----------------------------------------------------------------------------- CDN query generator
def install_mmdet_shim():
    def bbox_xyxy_to_cxcywh(bbox):        # mmdet 2.25.1 core/bbox/transforms.py
        x1, y1, x2, y2 = bbox.split((1, 1, 1, 1), dim=-1)
        return torch.cat([(x1 + x2) / 2, (y1 + y2) / 2, (x2 - x1), (y2 - y1)], dim=-1)

    def inverse_sigmoid(x, eps=1e-5):     # mmdet 2.25.1 models/utils/transformer.py
        x = x.clamp(min=0, max=1)
        x1 = x.clamp(min=eps)
        x2 = (1 - x).clamp(min=eps)
        return torch.log(x1 / x2)
    for name in ('mmdet', 'mmdet.core', 'mmdet.models', 'mmdet.models.utils', 'mmdet.models.utils.transformer'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['mmdet.core'].bbox_xyxy_to_cxcywh = bbox_xyxy_to_cxcywh
    sys.modules['mmdet.models.utils.transformer'].inverse_sigmoid = inverse_sigmoid


def golden_cdn():
    install_mmdet_shim()
    ref = load('models/multi/bbox_head/query_denoising.py', 'ref_query_denoising')
    fixtures = []
    orig_cuda, orig_to = torch.Tensor.cuda, torch.Tensor.to
    orig_rand_like, orig_randint_like = torch.rand_like, torch.randint_like

    def to_cpu(self, *a, **k):
        a = tuple(x for x in a if not (isinstance(x, str) and x.startswith('cuda')))
        k = {kk: v for kk, v in k.items() if not (kk == 'device' and str(v).startswith('cuda'))}
        return orig_to(self, *a, **k) if (a or k) else self
    for case, (sizes, num_dn, label_scale, box_scale, seed) in enumerate(
            [((3, 5), 10, 0.5, 1.0, 0), ((8,), 100, 0.5, 0.4, 1), ((1, 4, 2), 6, 0.5, 1.0, 2)]):
        g = torch.Generator().manual_seed(100 + seed)
        shapes = [(640 + 32 * b, 800 - 16 * b, 3) for b in range(len(sizes))]
        gt_bboxes, gt_labels = [], []
        for b, n in enumerate(sizes):
            h, w = shapes[b][:2]
            x1, y1 = torch.rand(n, generator=g) * (w - 150), torch.rand(n, generator=g) * (h - 150)
            bw, bh = torch.rand(n, generator=g) * 120 + 16, torch.rand(n, generator=g) * 120 + 16
            gt_bboxes.append(torch.stack([x1, y1, x1 + bw, y1 + bh], -1))
            gt_labels.append(torch.randint(0, 20, (n,), generator=g))
        hidden = 16
        label_enc = torch.nn.Embedding(20, hidden)
        with torch.no_grad():
            label_enc.weight.copy_(torch.randn(20, hidden, generator=g))
        gen = ref.CdnQueryGenerator(num_queries=30, hidden_dim=hidden, num_classes=20,
                                    noise_scale=dict(label=label_scale, box=box_scale),
                                    group_cfg=dict(dynamic=True, num_groups=None, num_dn_queries=num_dn))
        draws = []

        def rec_rand_like(t, *a, **k):
            r = orig_rand_like(t, *a, **k)
            draws.append(('rand_like', r.clone()))
            return r

        def rec_randint_like(t, *a, **k):
            r = orig_randint_like(t, *a, **k)
            draws.append(('randint_like', r.clone()))
            return r
        torch.manual_seed(seed)
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.Tensor.to = to_cpu
        torch.rand_like, torch.randint_like = rec_rand_like, rec_randint_like
        try:
            with torch.no_grad():
                q_label, q_bbox, attn_mask, dn_meta = gen(gt_bboxes, gt_labels, label_enc,
                                                          [dict(img_shape=s) for s in shapes])
        finally:
            torch.Tensor.cuda, torch.Tensor.to = orig_cuda, orig_to
            torch.rand_like, torch.randint_like = orig_rand_like, orig_randint_like
        assert [k for k, _ in draws] == ['rand_like', 'randint_like', 'randint_like', 'rand_like'], [k for k, _ in draws]
        fixtures.append(dict(sizes=list(sizes), num_dn_queries=num_dn, noise_scale=dict(label=label_scale, box=box_scale),
                             img_shapes=shapes, gt_bboxes=gt_bboxes, gt_labels=gt_labels,
                             label_embedding=label_enc.weight.detach().clone(), num_queries=30, hidden_dim=hidden,
                             draw_p=draws[0][1], draw_new_label_chosen=draws[1][1], draw_rand_sign=draws[2][1],
                             draw_rand_part=draws[3][1], out_query_label=q_label, out_query_bbox=q_bbox,
                             out_attn_mask=attn_mask, out_dn_meta={k: int(v) for k, v in dn_meta.items()}))
    torch.save(dict(source='models/multi/bbox_head/query_denoising.py::CdnQueryGenerator.__call__ (reference, run in '
                           'place on CPU; mmdet helpers restated, .cuda() neutralised)', cases=fixtures),
               os.path.join(OUT, 'reference_cdn.pt'))
    return fixtures


# ----------------------------------------------------------------------------- DINO decoder sine embedding
def golden_sineembed():
    """models/multi/bbox_head/transformer.py::DinoTransformerDecoder.gen_sineembed_for_position (a static, pure-torch
    method; row a14).  The module's mmcv / mmdet imports (registries and base classes it subclasses) are satisfied
    with empty stand-ins -- none of them takes part in this function."""
    install_mmdet_shim()

    class _Registry:
        def register_module(self, *a, **k):
            return lambda cls: cls
    for name in ('mmcv', 'mmcv.cnn', 'mmcv.cnn.bricks', 'mmcv.cnn.bricks.registry', 'mmdet.models.utils.builder'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['mmcv.cnn.bricks.registry'].TRANSFORMER_LAYER_SEQUENCE = _Registry()
    sys.modules['mmdet.models.utils.builder'].TRANSFORMER = _Registry()
    t = sys.modules['mmdet.models.utils.transformer']
    for cls in ('DeformableDetrTransformerDecoder', 'DeformableDetrTransformer', 'Transformer'):
        setattr(t, cls, type(cls, (torch.nn.Module,), {}))
    t.build_transformer_layer_sequence = lambda *a, **k: None
    ref = load('models/multi/bbox_head/transformer.py', 'ref_dino_transformer')
    g = torch.Generator().manual_seed(7)
    cases = []
    for n in (2, 4):
        pos = torch.rand(2, 37, n, generator=g)
        cases.append(dict(pos=pos, out=ref.DinoTransformerDecoder.gen_sineembed_for_position(pos)))
    torch.save(dict(source='models/multi/bbox_head/transformer.py::DinoTransformerDecoder.gen_sineembed_for_position '
                           '(reference, run in place)', cases=cases), os.path.join(OUT, 'reference_sineembed.pt'))
    return cases


# ----------------------------------------------------------------------------- DINO head: denoising targets
class _AutoStubModule(types.ModuleType):
    """a module whose every attribute is an inert stand-in: usable as a base class, as a registry
    (`X.register_module()`), as a decorator factory (`@force_fp32(apply_to=...)`) -- enough to IMPORT the
    reference's head modules; none of the stand-ins is executed by the functions the goldens call."""

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)

        class _Stub(torch.nn.Module):
            def __init__(self, *a, **k):
                super().__init__()

            def __call__(self, f=None, *a, **k):      # decorator use: @stub(...)(fn) -> fn
                return f

            @classmethod
            def register_module(cls, *a, **k):
                return lambda c: c
        _Stub.__name__ = name
        setattr(self, name, _Stub)
        return _Stub


def golden_dn_targets():
    """models/multi/bbox_head/dino_head.py::DINOHead.get_dn_target / _get_dn_target_single (rows a16, dn part):
    the fixed (non-Hungarian) targets of the denoising queries.  Called UNBOUND on a bare namespace carrying
    num_classes: the class itself cannot be constructed without mmdet."""
    install_mmdet_shim()

    def multi_apply(func, *args, **kwargs):        # mmdet 2.25.1 core/utils/misc.py
        from functools import partial
        pfunc = partial(func, **kwargs) if kwargs else func
        return tuple(map(list, zip(*map(pfunc, *args))))
    for name in ('mmcv', 'mmcv.cnn', 'mmcv.cnn.bricks', 'mmcv.cnn.bricks.transformer', 'mmcv.runner', 'mmdet.models.builder',
                 'mmdet.models.utils', 'mmdet.models.dense_heads', 'mmdet.models.dense_heads.anchor_free_head'):
        if not isinstance(sys.modules.get(name), _AutoStubModule):
            old = sys.modules.get(name)
            new = _AutoStubModule(name)
            if old is not None:
                new.__dict__.update({k: v for k, v in old.__dict__.items() if not k.startswith('__')})
            sys.modules[name] = new
    core = _AutoStubModule('mmdet.core')
    core.__dict__.update({k: v for k, v in sys.modules['mmdet.core'].__dict__.items() if not k.startswith('__')})
    core.multi_apply = multi_apply
    sys.modules['mmdet.core'] = core
    pkg = types.ModuleType('ref_bbox_head')
    pkg.__path__ = [os.path.join(REF, 'models', 'multi', 'bbox_head')]
    sys.modules['ref_bbox_head'] = pkg
    sub = types.ModuleType('ref_bbox_head.mmdet_detr_head')
    sub.__path__ = [os.path.join(REF, 'models', 'multi', 'bbox_head', 'mmdet_detr_head')]
    sys.modules['ref_bbox_head.mmdet_detr_head'] = sub
    detr = load('models/multi/bbox_head/mmdet_detr_head/detr_head.py', 'ref_bbox_head.mmdet_detr_head.detr_head')
    ddetr = load('models/multi/bbox_head/mmdet_detr_head/deformable_detr_head.py',
                 'ref_bbox_head.mmdet_detr_head.deformable_detr_head')
    sub.DETRHead, sub.DeformableDETRHead = detr.DETRHead, ddetr.DeformableDETRHead
    load('models/multi/bbox_head/query_denoising.py', 'ref_bbox_head.query_denoising')
    dino = load('models/multi/bbox_head/dino_head.py', 'ref_bbox_head.dino_head')
    H = dino.DINOHead
    fake = types.SimpleNamespace(num_classes=20)
    fake._get_dn_target_single = lambda *a: H._get_dn_target_single(fake, *a)
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    cases = []
    try:
        for sizes, num_groups in [((3, 5), 2), ((8,), 12), ((0, 4), 3)]:
            g = torch.Generator().manual_seed(sum(sizes) + num_groups)
            single = max(sizes) * 2
            dn_meta = dict(pad_size=single * num_groups, num_dn_group=num_groups)
            shapes = [(640 + 32 * b, 800 - 16 * b, 3) for b in range(len(sizes))]
            gtb, gtl, preds = [], [], []
            for b, n in enumerate(sizes):
                h, w = shapes[b][:2]
                x1, y1 = torch.rand(n, generator=g) * (w - 150), torch.rand(n, generator=g) * (h - 150)
                gtb.append(torch.stack([x1, y1, x1 + 40 + 80 * torch.rand(n, generator=g), y1 + 30 + 90 * torch.rand(n, generator=g)], -1))
                gtl.append(torch.randint(0, 20, (n,), generator=g))
                preds.append(torch.rand(dn_meta['pad_size'], 4, generator=g))
            metas = [dict(img_shape=s) for s in shapes]
            out = H.get_dn_target(fake, preds, gtb, gtl, metas, [dn_meta for _ in sizes])
            cases.append(dict(sizes=list(sizes), dn_meta=dn_meta, img_shapes=shapes, gt_bboxes=gtb, gt_labels=gtl,
                              dn_bbox_preds=preds, labels=out[0], label_weights=out[1], bbox_targets=out[2],
                              bbox_weights=out[3], num_total_pos=int(out[4]), num_total_neg=int(out[5])))
    finally:
        torch.Tensor.cuda = orig_cuda
    torch.save(dict(source='models/multi/bbox_head/dino_head.py::DINOHead.get_dn_target (reference, run in place, unbound)',
                    cases=cases), os.path.join(OUT, 'reference_dn_targets.pt'))
    return cases


# ----------------------------------------------------------------------------- MTL.train_step / _parse_losses, seg forward_head
def _stub_packages(names):
    for name in names:
        parts = name.split('.')
        for k in range(1, len(parts) + 1):
            n = '.'.join(parts[:k])
            if not isinstance(sys.modules.get(n), _AutoStubModule):
                old, new = sys.modules.get(n), _AutoStubModule(n)
                if old is not None:
                    new.__dict__.update({kk: v for kk, v in old.__dict__.items() if not kk.startswith('__')})
                new.__path__ = []
                sys.modules[n] = new


def golden_train_step():
    """models/multi/multitask_learner.py::MTL.train_step + _parse_losses (row a21), called UNBOUND on a small
    callable object that returns preset loss dicts: total = sum of the '*loss*' means, log_vars prefixed with
    '<task>.<dataset>' and multiplied by the task weight."""
    _stub_packages(['matplotlib.font_manager', 'matplotlib.pyplot', 'matplotlib.collections', 'matplotlib.patches',
                    'mmcv.runner', 'mmcv.cnn.bricks.transformer', 'mmcv.cnn', 'mmcls.models.utils.augment', 'mmdet.core',
                    'mmdet.core.visualization', 'mmseg.core', 'mmseg.ops', 'mtl.model.build', 'mmdet.models.utils.transformer'])
    sys.modules['mmseg.core'].add_prefix = lambda inputs, prefix: {'%s.%s' % (prefix, k): v for k, v in inputs.items()}  # mmseg 0.28 core/utils/misc.py
    ref = load('models/multi/multitask_learner.py', 'ref_multitask_learner')
    MTL = ref.MTL

    class Fake:
        task_weight = dict(cls=1, det=1, seg=0.1)

        def __init__(self, losses):
            self.losses = losses

        def __call__(self, **data):
            return self.losses

        def _parse_losses(self, losses):
            return MTL._parse_losses(self, losses)
    g = torch.Generator().manual_seed(11)
    cases = []
    for task, ds, keys in [('cls', 'resisc', ['loss']), ('seg', 'potsdam', ['seg.loss_ce', 'seg.acc_seg']),
                           ('det', 'dior', ['interm_loss_cls', 'loss_cls', 'loss_bbox', 'd0.loss_iou', 'dn_loss_cls', 'list_loss'])]:
        losses = {}
        for k in keys:
            losses[k] = [torch.rand(3, generator=g), torch.rand((), generator=g)] if k == 'list_loss' else \
                torch.rand((2,) if k == 'loss_bbox' else (), generator=g) * 3
        data = dict(task=task, dataset_name=ds, img_metas=[{}] * (2 if task != 'det' else 1))
        out = MTL.train_step(Fake({k: (v if isinstance(v, list) else v.clone()) for k, v in losses.items()}), data, None)
        cases.append(dict(task=task, dataset_name=ds, losses=losses, loss=out['loss'].clone(),
                          log_vars={k: float(v) for k, v in out['log_vars'].items()}, num_samples=out['num_samples']))
    torch.save(dict(source='models/multi/multitask_learner.py::MTL.train_step/_parse_losses (reference, run in place, unbound)',
                    cases=cases), os.path.join(OUT, 'reference_train_step.pt'))
    return cases


def golden_seg_forward_head():
    """models/multi/seg_head/mask2former_head.py::Mask2FormerHead.forward_head (row a18), called UNBOUND on a namespace
    holding plain torch modules (post_norm LayerNorm, mask_embed MLP): mask prediction einsum, bilinear resize to
    the next level and the sigmoid < 0.5 attention mask repeated over heads."""
    _stub_packages(['mmcv.cnn', 'mmcv.cnn.bricks.transformer', 'mmcv.runner', 'mmseg.models.builder',
                    'mmseg.models.decode_heads.decode_head', 'mmdet.models.utils.transformer'])
    pkg = types.ModuleType('ref_seg_head')
    pkg.__path__ = [os.path.join(REF, 'models', 'multi', 'seg_head')]
    sys.modules['ref_seg_head'] = pkg
    try:
        ref = load('models/multi/seg_head/mask2former_head.py', 'ref_seg_head.mask2former_head')
    except Exception:
        _stub_packages(['ref_seg_head.pixel_decoder'])
        ref = load('models/multi/seg_head/mask2former_head.py', 'ref_seg_head.mask2former_head')
    H = ref.Mask2FormerHead
    torch.manual_seed(3)
    C = 32
    fake = types.SimpleNamespace(scheme=2, num_heads=4,
                                 transformer_decoder=types.SimpleNamespace(post_norm=torch.nn.LayerNorm(C)),
                                 mask_embed=torch.nn.Sequential(torch.nn.Linear(C, C), torch.nn.ReLU(), torch.nn.Linear(C, C),
                                                                torch.nn.ReLU(), torch.nn.Linear(C, C)))
    decoder_out = torch.randn(7, 2, C)                # (queries, batch, C)
    mask_feature = torch.randn(2, C, 12, 10)
    with torch.no_grad():
        seg_mask, attn_mask = H.forward_head(fake, decoder_out, mask_feature, (5, 6))
    sd = {'post_norm.' + k: v for k, v in fake.transformer_decoder.post_norm.state_dict().items()}
    sd.update({'mask_embed.' + k: v for k, v in fake.mask_embed.state_dict().items()})
    torch.save(dict(source='models/multi/seg_head/mask2former_head.py::Mask2FormerHead.forward_head (reference, run in place, unbound)',
                    state=sd, decoder_out=decoder_out, mask_feature=mask_feature, target_size=(5, 6), num_heads=4,
                    seg_mask=seg_mask, attn_mask=attn_mask), os.path.join(OUT, 'reference_seg_forward_head.pt'))
    return seg_mask.shape, attn_mask.shape


# ----------------------------------------------------------------------------- MultiDataLoader
class _ToyDataset(torch.utils.data.Dataset):
    """n samples {'idx': i}; `task` is what MultiDataLoader tags batches with"""

    def __init__(self, n, task):
        self.n, self.task = n, task

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return dict(idx=torch.tensor(i))


def toy_loaders():
    spec = dict(resisc=(7, 'cls', 2), dior=(3, 'det', 1), potsdam=(5, 'seg', 2))     # (samples, task, batch size)
    return {k: torch.utils.data.DataLoader(_ToyDataset(n, t), batch_size=b, shuffle=False) for k, (n, t, b) in spec.items()}


def golden_multi_data_loader():
    """mtl/data/multi_data_loader.py::MultiDataLoader (row a22), run in place as a synthetic package so that its
    relative imports (.iteration_strategies, .sample) resolve without executing mtl/data/__init__.py (mmcv)."""
    install_omegaconf_shim()
    pkg = types.ModuleType('ref_mtl_data')
    pkg.__path__ = [os.path.join(REF, 'mtl', 'data')]
    sys.modules['ref_mtl_data'] = pkg
    strat_mod = load('mtl/data/iteration_strategies.py', 'ref_mtl_data.iteration_strategies')
    load('mtl/data/sample.py', 'ref_mtl_data.sample')
    mdl = load('mtl/data/multi_data_loader.py', 'ref_mtl_data.multi_data_loader')
    out = []
    for name, kw in [('RoundRobinIterationStrategy', {}), ('SizeProportionalIterationStrategy', {}),
                     ('RandomIterationStrategy', {})]:
        loaders = toy_loaders()
        np.random.seed(5)
        strat = getattr(strat_mod, name).from_params(loaders, **kw)
        loader = mdl.MultiDataLoader(loaders, strat)
        seq, it = [], iter(loader)
        for _ in range(30):
            try:
                b = next(it)
            except StopIteration:
                seq.append('StopIteration')
                it = iter(loader)
                continue
            seq.append([b['dataset_name'], b['task'], [int(v) for v in b['idx']]])
        out.append(dict(strategy=name, numpy_seed=5, length=len(loader), sequence=seq))
    json.dump(dict(source='mtl/data/multi_data_loader.py::MultiDataLoader (reference, run in place)',
                   loaders='resisc: 7 samples / batch 2 / cls; dior: 3 / 1 / det; potsdam: 5 / 2 / seg; shuffle off',
                   cases=out), open(os.path.join(OUT, 'reference_multi_data_loader.json'), 'w'), indent=1)
    return out


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    t = golden_train_step()
    print('train_step:', [(c['task'], float(c['loss']), len(c['log_vars'])) for c in t])
    print('seg forward_head:', golden_seg_forward_head())
    d = golden_dn_targets()
    print('dn targets:', [(c['sizes'], c['num_total_pos'], c['num_total_neg']) for c in d])
    m = golden_multi_data_loader()
    print('multi data loader:', [(c['strategy'], c['sequence'][:4]) for c in m])
    e = golden_sineembed()
    print('sine embed cases:', [tuple(c['out'].shape) for c in e])
    s = golden_iteration_strategies()
    print('iteration strategies:', [(c['strategy'], c.get('error', 'ok')) for c in s])
    f = golden_cdn()
    print('cdn cases:', [(c['sizes'], c['out_dn_meta']) for c in f])
