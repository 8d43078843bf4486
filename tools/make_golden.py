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
import torch.nn.functional as F

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


# ----------------------------------------------------------------------------- control flow of the heads with toy sub-modules
class ToyDecoderLayer(torch.nn.Module):
    """Deterministic stand-in for an mmcv BaseTransformerLayer (third-party): mixes the query with a masked mean of
    value (+ key_pos) and with query_pos, so the output depends on which level / mask / positions it is handed.
    Used IDENTICALLY by the reference run (tools/make_golden.py) and by the tests of this repo."""

    def forward(self, query, key=None, value=None, query_pos=None, key_pos=None, attn_masks=None,
                query_key_padding_mask=None, key_padding_mask=None, reference_points=None, **kwargs):
        out = query
        if value is not None:
            v = value.float() + (key_pos.float() if key_pos is not None and key_pos.shape == value.shape else 0)
            mask = attn_masks[0] if isinstance(attn_masks, (list, tuple)) else attn_masks
            if mask is not None and mask.dim() == 3:           # (B*heads, Q, K): use head 0 of every image
                B = value.shape[1]
                w = (~mask.view(B, -1, mask.shape[-2], mask.shape[-1])[:, 0]).float()      # (B, Q, K)
                w = w / w.sum(-1, keepdim=True).clamp(min=1.0)
                out = out + 0.5 * torch.einsum('bqk,kbc->qbc', w, v).to(out.dtype)
            else:
                out = out + 0.5 * v.mean(0, keepdim=True).to(out.dtype)
        if query_pos is not None:
            out = out + 0.1 * query_pos.to(out.dtype)
        if reference_points is not None:                        # (B, Q, L, 2|4) -> (Q, B, 1)
            out = out + 0.05 * reference_points.float().mean((2, 3)).transpose(0, 1).unsqueeze(-1).to(out.dtype)
        return torch.tanh(out)


def toy_seg_parts(C=16, Q=6, B=2):
    g = torch.Generator().manual_seed(21)
    shapes = [(3, 3), (4, 5), (7, 6), (9, 8)]                  # low -> high resolution
    memories = [torch.randn(B, C, h, w, generator=g) for h, w in shapes]
    mask_features = torch.randn(B, C, 9, 8, generator=g)
    pos = {hw: torch.randn(B, C, hw[0], hw[1], generator=g) for hw in shapes}
    emb = lambda n: torch.randn(n, C, generator=g)
    return dict(C=C, Q=Q, B=B, shapes=shapes, memories=memories, mask_features=mask_features, pos=pos,
                level_embed=emb(4), query_feat=emb(Q), query_embed=emb(Q), post_norm_w=torch.rand(C, generator=g) + 0.5,
                post_norm_b=torch.randn(C, generator=g) * 0.1,
                mlp=[torch.randn(C, C, generator=g) * 0.3 for _ in range(3)], mlp_b=[torch.randn(C, generator=g) * 0.1 for _ in range(3)])


def golden_seg_forward():
    """models/multi/seg_head/mask2former_head.py::Mask2FormerHead.forward (row a18: level cycling, all-masked-row reset,
    forward_head after every layer) with toy decoder layers / pixel decoder / positional encoding."""
    ref = sys.modules['ref_seg_head.mask2former_head']
    H = ref.Mask2FormerHead
    t = toy_seg_parts()
    C = t['C']
    post_norm = torch.nn.LayerNorm(C)
    mask_embed = torch.nn.Sequential(torch.nn.Linear(C, C), torch.nn.ReLU(), torch.nn.Linear(C, C), torch.nn.ReLU(),
                                     torch.nn.Linear(C, C))
    with torch.no_grad():
        post_norm.weight.copy_(t['post_norm_w']), post_norm.bias.copy_(t['post_norm_b'])
        for k, i in enumerate((0, 2, 4)):
            mask_embed[i].weight.copy_(t['mlp'][k]), mask_embed[i].bias.copy_(t['mlp_b'][k])
    emb = lambda w: torch.nn.Embedding.from_pretrained(w.clone(), freeze=False)
    fake = types.SimpleNamespace(
        scheme=2, num_heads=2, num_transformer_feat_level=4, num_transformer_decoder_layers=9,
        pixel_decoder=lambda enc, neck, bb: (t['mask_features'], t['memories']),
        decoder_input_projs=[torch.nn.Identity() for _ in range(4)], level_embed=emb(t['level_embed']),
        decoder_positional_encoding=lambda mask: t['pos'][tuple(mask.shape[-2:])],
        query_feat=emb(t['query_feat']), query_embed=emb(t['query_embed']), mask_embed=mask_embed,
        transformer_decoder=types.SimpleNamespace(post_norm=post_norm, layers=[ToyDecoderLayer() for _ in range(9)]))
    fake.forward_head = lambda *a: H.forward_head(fake, *a)
    with torch.no_grad():
        out = H.forward(fake, None, None, None, [{}] * t['B'])
    torch.save(dict(source='models/multi/seg_head/mask2former_head.py::Mask2FormerHead.forward (reference, run in place, toy layers)',
                    out=out), os.path.join(OUT, 'reference_seg_forward.pt'))
    return tuple(out.shape)


def toy_det_parts(C=16, B=2, Q=9, pad=4, L=3, classes=5):
    g = torch.Generator().manual_seed(33)
    lin = lambda o, i: (torch.randn(o, i, generator=g) * 0.3, torch.randn(o, generator=g) * 0.1)
    return dict(C=C, B=B, Q=Q, pad=pad, L=L, classes=classes,
                query=torch.randn(pad + Q, B, C, generator=g), memory=torch.randn(20, B, C, generator=g),
                reference_points=torch.rand(B, pad + Q, 4, generator=g) * 0.8 + 0.1,
                valid_ratios=torch.rand(B, 4, 2, generator=g) * 0.3 + 0.7,
                ref_point_head=[lin(C, 32 * 4 * 4), lin(C, C)], norm=(torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1),
                reg=[[lin(C, C), lin(C, C), lin(4, C)] for _ in range(L + 1)], cls=[lin(classes, C) for _ in range(L + 1)],
                topk_score=torch.randn(B, Q, classes, generator=g), topk_anchor=torch.rand(B, Q, 4, generator=g),
                feats=[torch.randn(B, C, 8, 8, generator=g), torch.randn(B, C, 4, 4, generator=g)])


def _mlp_from(ws, cls=torch.nn.Linear):
    layers = []
    for k, (w, b) in enumerate(ws):
        l = cls(w.shape[1], w.shape[0])
        with torch.no_grad():
            l.weight.copy_(w), l.bias.copy_(b)
        layers.append(l)
        if k + 1 < len(ws):
            layers.append(torch.nn.ReLU())
    return torch.nn.Sequential(*layers) if len(layers) > 1 else layers[0]


def golden_dino_decoder_and_head():
    """models/multi/bbox_head/transformer.py::DinoTransformerDecoder.forward (row a14: reference-point scaling, sine
    embedding -> ref_point_head, box refinement with detach, look-forward-twice outputs) and
    models/multi/bbox_head/dino_head.py::DINOHead.forward (row a13: padding masks, per-level class / box branches on
    inverse_sigmoid(reference)), both run in place with toy decoder layers / a toy transformer."""
    tr = sys.modules['ref_dino_transformer']
    dino = sys.modules['ref_bbox_head.dino_head']
    t = toy_det_parts()
    C, L = t['C'], t['L']
    # NB: the sine embedding is 4 x 128 = 512 wide by construction
    rph = [(torch.randn(C, 512, generator=torch.Generator().manual_seed(1)) * 0.05, t['ref_point_head'][0][1]), t['ref_point_head'][1]]
    norm = torch.nn.LayerNorm(C)
    with torch.no_grad():
        norm.weight.copy_(t['norm'][0]), norm.bias.copy_(t['norm'][1])
    reg = [_mlp_from(ws) for ws in t['reg']]
    cls = [_mlp_from([w]) for w in t['cls']]
    dec = types.SimpleNamespace(layers=[ToyDecoderLayer() for _ in range(L)], ref_point_head=_mlp_from(rph), norm=norm,
                                return_intermediate=True,
                                gen_sineembed_for_position=tr.DinoTransformerDecoder.gen_sineembed_for_position)
    with torch.no_grad():
        hs, refs = tr.DinoTransformerDecoder.forward(dec, t['query'], None, t['memory'], reference_points=t['reference_points'],
                                                     valid_ratios=t['valid_ratios'], reg_branches=reg)
    # head: toy transformer returns the decoder outputs above
    metas = [dict(img_shape=(28, 32, 3), batch_input_shape=(32, 32)), dict(img_shape=(32, 30, 3), batch_input_shape=(32, 32))]
    seen = {}

    def toy_transformer(mlvl_feats, mlvl_masks, query_embeds, mlvl_pos, dn_label_query, dn_bbox_query, attn_mask, encoder,
                        reg_branches=None, cls_branches=None, **kw):
        seen['masks'] = [m.clone() for m in mlvl_masks]
        return hs, refs, t['topk_score'], t['topk_anchor']
    head = types.SimpleNamespace(transformer=toy_transformer, positional_encoding=lambda m: m.float().unsqueeze(1),
                                 with_box_refine=True, as_two_stage=True, reg_branches=reg, cls_branches=cls,
                                 label_embedding=torch.nn.Embedding(t['classes'], C))
    with torch.no_grad():
        oc, ob, ts, ta = dino.DINOHead.forward(head, None, t['feats'], metas, torch.zeros(t['B'], t['pad'], C),
                                               torch.zeros(t['B'], t['pad'], 4), None)
    torch.save(dict(source='DinoTransformerDecoder.forward + DINOHead.forward (reference, run in place, toy layers)',
                    rph0=rph[0][0], hs=hs, refs=refs, masks=seen['masks'], outputs_classes=oc, outputs_coords=ob, metas=metas),
               os.path.join(OUT, 'reference_dino_decoder_head.pt'))
    return tuple(hs.shape), tuple(refs.shape), tuple(oc.shape), tuple(ob.shape)


def toy_loss_parts(L=3, B=2, Q=12, pad=8, classes=20, sizes=(3, 2), groups=2):
    g = torch.Generator().manual_seed(66)
    shapes = [(640, 800, 3), (608, 736, 3)]
    gtb, gtl = [], []
    for b, n in enumerate(sizes):
        h, w = shapes[b][:2]
        x1, y1 = torch.rand(n, generator=g) * (w - 200), torch.rand(n, generator=g) * (h - 200)
        gtb.append(torch.stack([x1, y1, x1 + 30 + 150 * torch.rand(n, generator=g), y1 + 30 + 150 * torch.rand(n, generator=g)], -1))
        gtl.append(torch.randint(0, classes, (n,), generator=g))
    box = lambda *s_: torch.cat([torch.rand(*s_, 2, generator=g) * 0.8 + 0.1, torch.rand(*s_, 2, generator=g) * 0.3 + 0.05], -1)
    return dict(L=L, B=B, Q=Q, pad=pad, classes=classes, sizes=sizes, img_shapes=shapes, gt_bboxes=gtb, gt_labels=gtl,
                all_cls=torch.randn(L, B, pad + Q, classes, generator=g) - 2.0, all_box=box(L, B, pad + Q),
                enc_cls=torch.randn(B, Q, classes, generator=g) - 2.0, enc_box=box(B, Q),
                dn_meta=dict(pad_size=pad, num_dn_group=groups))


def golden_dino_loss():
    """models/multi/bbox_head/dino_head.py::DINOHead.loss (+ loss_dn, loss_dn_single, get_dn_target) and
    mmdet_detr_head/detr_head.py::DETRHead.loss_single / get_targets / _get_target_single (row a16), run in place,
    UNBOUND on a namespace.  The mmdet objects those methods call (HungarianAssigner, PseudoSampler, FocalLoss,
    L1Loss, GIoULoss, reduce_mean, box converters) are third-party and absent: they are restated from the oracle
    (oracle/heads.py) and therefore NOT pinned by this fixture -- the in-tree flow is: target construction,
    cls_avg_factor / num_total_pos, the per-image rescale factors, loss weights, dn handling and the 39 keys."""
    sys.path.insert(0, ROOT)
    from oracle import heads as oh
    core = sys.modules['mmdet.core']
    core.bbox_cxcywh_to_xyxy, core.bbox_xyxy_to_cxcywh = oh.bbox_cxcywh_to_xyxy, oh.bbox_xyxy_to_cxcywh
    core.reduce_mean = lambda t: t                       # single process
    detr = sys.modules['ref_bbox_head.mmdet_detr_head.detr_head']
    dino = sys.modules['ref_bbox_head.dino_head']
    for mod in (detr, dino):                             # names the modules imported before the shim was complete
        mod.reduce_mean, mod.bbox_cxcywh_to_xyxy, mod.bbox_xyxy_to_cxcywh = core.reduce_mean, oh.bbox_cxcywh_to_xyxy, oh.bbox_xyxy_to_cxcywh
        mod.multi_apply = core.multi_apply

    class ToyAssigner:       # mmdet HungarianAssigner.assign -> AssignResult(num_gts, gt_inds, max_overlaps, labels)
        def assign(self, bbox_pred, cls_pred, gt_bboxes, gt_labels, img_meta, gt_bboxes_ignore=None, eps=1e-7):
            gt_inds = oh.hungarian_assign(bbox_pred, cls_pred, gt_bboxes, gt_labels, img_meta['img_shape'])
            return types.SimpleNamespace(gt_inds=gt_inds, num_gts=gt_bboxes.size(0))

    class ToySampler:        # mmdet PseudoSampler.sample -> SamplingResult
        def sample(self, assign_result, bboxes, gt_bboxes, **kw):
            pos_inds = torch.nonzero(assign_result.gt_inds > 0, as_tuple=False).squeeze(-1).unique()
            neg_inds = torch.nonzero(assign_result.gt_inds == 0, as_tuple=False).squeeze(-1).unique()
            pos_gt = assign_result.gt_inds[pos_inds] - 1
            pos_gt_bboxes = gt_bboxes[pos_gt.long(), :] if gt_bboxes.numel() else torch.empty_like(gt_bboxes).view(-1, 4)
            return types.SimpleNamespace(pos_inds=pos_inds, neg_inds=neg_inds, pos_assigned_gt_inds=pos_gt,
                                         pos_gt_bboxes=pos_gt_bboxes)

    def loss_cls(pred, target, weight=None, avg_factor=None):              # mmdet FocalLoss(gamma 2, alpha .25, w 1)
        return 1.0 * oh.py_sigmoid_focal_loss(pred, target, weight, 2.0, 0.25, avg_factor)

    def loss_bbox(pred, target, weight=None, avg_factor=None):             # mmdet L1Loss(loss_weight 5)
        if target.numel() == 0:
            return pred.sum() * 0
        return 5.0 * ((pred - target).abs() * weight).sum() / avg_factor

    def loss_iou(pred, target, weight=None, avg_factor=None):              # mmdet GIoULoss(loss_weight 2)
        if weight is not None and not torch.any(weight > 0):
            return (pred * weight).sum()
        w = weight.mean(-1)
        return 2.0 * ((1 - oh.bbox_overlaps_giou(pred, target, True)) * w).sum() / avg_factor
    t = toy_loss_parts()
    H, D = dino.DINOHead, detr.DETRHead
    class FakeHead(types.SimpleNamespace):        # (the reference formats self.__class__.__name__ into an assert message)
        pass
    fake = FakeHead(num_classes=t['classes'], cls_out_channels=t['classes'], bg_cls_weight=0, sync_cls_avg_factor=True,
                                 assigner=ToyAssigner(), sampler=ToySampler(), loss_cls=loss_cls, loss_bbox=loss_bbox,
                                 loss_iou=loss_iou, extract_dn_outputs=H.extract_dn_outputs)
    for name, owner in [('loss_single', D), ('get_targets', D), ('_get_target_single', D), ('loss_dn', H),
                        ('loss_dn_single', H), ('get_dn_target', H), ('_get_dn_target_single', H)]:
        setattr(fake, name, (lambda f: (lambda *a, **k: f(fake, *a, **k)))(getattr(owner, name)))
    metas = [dict(img_shape=s) for s in t['img_shapes']]
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        with torch.no_grad():
            losses = H.loss(fake, t['all_cls'], t['all_box'], t['enc_cls'], t['enc_box'], t['gt_bboxes'], t['gt_labels'], metas,
                            t['dn_meta'])
    finally:
        torch.Tensor.cuda = orig_cuda
    out = {k: float(v) for k, v in losses.items()}
    json.dump(dict(source='DINOHead.loss / DETRHead.loss_single (reference, run in place; third-party assigner / sampler / '
                          'loss modules restated from the oracle)', keys=list(out.keys()), losses=out),
              open(os.path.join(OUT, 'reference_dino_loss.json'), 'w'), indent=1)
    return len(out), sum(out.values())


def toy_two_stage_parts(C=16, B=2, K=7, classes=5, pad=3):
    g = torch.Generator().manual_seed(44)
    lin = lambda o, i: (torch.randn(o, i, generator=g) * 0.3, torch.randn(o, generator=g) * 0.1)
    shapes = [(6, 5), (3, 3)]
    masks = [torch.zeros(B, h, w, dtype=torch.bool) for h, w in shapes]
    masks[0][1, :, 4:] = True                      # image 1 is padded on the right
    masks[1][1, :, 2:] = True
    return dict(C=C, B=B, K=K, classes=classes, pad=pad, shapes=shapes, masks=masks,
                feats=[torch.randn(B, C, h, w, generator=g) for h, w in shapes],
                pos=[torch.randn(B, C, h, w, generator=g) for h, w in shapes], level_embeds=torch.randn(2, C, generator=g),
                enc_output=lin(C, C), enc_norm=(torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1),
                cls=lin(classes, C), reg=[lin(C, C), lin(C, C), lin(4, C)], query_embed=torch.randn(K, C, generator=g),
                dn_label=torch.randn(B, pad, C, generator=g), dn_bbox=torch.randn(B, pad, 4, generator=g))


def toy_encoder(query=None, query_pos=None, **kw):
    return torch.tanh(query + 0.1 * query_pos.to(query.dtype))


class ToyDecoder:
    num_layers = 1

    def __call__(self, query=None, value=None, reference_points=None, **kw):
        self.seen = dict(query=query.clone(), reference_points=reference_points.clone(), value=value.clone(),
                         key_padding_mask=kw.get('key_padding_mask'), attn_masks=kw.get('attn_masks'))
        return torch.stack([query, query * 0.5]), torch.stack([reference_points, reference_points * 0.9])


def golden_dino_transformer():
    """models/multi/bbox_head/transformer.py::DinoTransformer.forward (row a13: flattening, level embeddings, the
    two-stage top-k proposal selection, dn / matching query concatenation), run in place with a toy encoder and
    decoder.  Its three mmdet BASE-CLASS methods (get_valid_ratio, get_reference_points,
    gen_encoder_output_proposals; third-party, not in /root/reference) are taken from this repo's restatement, so
    THEY are not pinned by this fixture -- the in-tree control flow around them is."""
    sys.path.insert(0, ROOT)
    from rscotr_b200.models.det_head import DinoTransformer as Mine
    tr = sys.modules['ref_dino_transformer']
    t = toy_two_stage_parts()
    C = t['C']
    enc_out = _mlp_from([t['enc_output']])
    enc_norm = torch.nn.LayerNorm(C)
    with torch.no_grad():
        enc_norm.weight.copy_(t['enc_norm'][0]), enc_norm.bias.copy_(t['enc_norm'][1])
    dec = ToyDecoder()
    fake = types.SimpleNamespace(as_two_stage=True, level_embeds=t['level_embeds'], decoder=dec, two_stage_num_proposals=t['K'],
                                 query_embed=torch.nn.Embedding.from_pretrained(t['query_embed'].clone()), enc_output=enc_out,
                                 enc_output_norm=enc_norm, get_valid_ratio=Mine.get_valid_ratio,
                                 get_reference_points=lambda ss, vr, device=None: Mine.get_reference_points(
                                     [tuple(x) for x in ss.tolist()], vr, device))
    fake.gen_encoder_output_proposals = lambda mem, mask, ss: Mine.gen_encoder_output_proposals(
        fake, mem, mask, [tuple(x) for x in ss.tolist()])
    fake.proposal_grid = Mine.proposal_grid
    cls = [None, _mlp_from([t['cls']])]
    reg = [None, _mlp_from(t['reg'])]
    with torch.no_grad():
        out = tr.DinoTransformer.forward(fake, t['feats'], t['masks'], None, t['pos'], t['dn_label'], t['dn_bbox'], None,
                                         toy_encoder, reg_branches=reg, cls_branches=cls)
    torch.save(dict(source='models/multi/bbox_head/transformer.py::DinoTransformer.forward (reference, run in place, toy encoder/decoder)',
                    out=[o.clone() for o in out], decoder_saw={k: v for k, v in dec.seen.items()}),
               os.path.join(OUT, 'reference_dino_transformer.pt'))
    return [tuple(o.shape) for o in out]


def toy_pixel_decoder_parts(C=8, B=2):
    g = torch.Generator().manual_seed(55)
    sizes = [(16, 12), (8, 6), (4, 3), (2, 2)]                 # neck levels, strides 8 / 16 / 32 / 64 (high -> low resolution)
    return dict(C=C, B=B, sizes=sizes, strides=[8, 16, 32, 64], neck=[torch.randn(B, C, h, w, generator=g) for h, w in sizes],
                pos={hw: torch.randn(B, C, hw[0], hw[1], generator=g) for hw in sizes},
                level_encoding=torch.randn(4, C, generator=g), mask_w=torch.randn(C, C, 1, 1, generator=g) * 0.3,
                mask_b=torch.randn(C, generator=g) * 0.1, backbone=[torch.randn(B, C, 4, 4, generator=g)])


class ToyPointGenerator:
    """mmdet 2.25.1 core/anchor/point_generator.py MlvlPointGenerator.single_level_grid_priors, offset 0.5, restated
    (third-party; NOT pinned by the fixture)."""

    def __init__(self, strides):
        self.strides = strides

    def single_level_grid_priors(self, featmap_size, level_idx, device='cpu'):
        h, w = featmap_size
        s = self.strides[level_idx]
        sx = (torch.arange(0, w, device=device) + 0.5) * s
        sy = (torch.arange(0, h, device=device) + 0.5) * s
        xx = sx.repeat(len(sy))
        yy = sy.view(-1, 1).repeat(1, len(sx)).view(-1)
        return torch.stack([xx, yy], dim=-1).float()


class ToyEncoder:
    def __call__(self, query=None, query_pos=None, **kw):
        self.seen = dict(query=query.clone(), query_pos=query_pos.clone(), reference_points=kw['reference_points'].clone(),
                         spatial_shapes=kw['spatial_shapes'].clone(), level_start_index=kw['level_start_index'].clone())
        return torch.tanh(query + 0.1 * query_pos.to(query.dtype))


def golden_seg_pixel_decoder():
    """models/multi/seg_head/pixel_decoder.py::MlvlSegPixelDecoder.forward (row a17: level order low -> high
    resolution, level encodings, normalised reference points, what the shared encoder is called with, memory split,
    mask_feature conv), run in place with a toy encoder / positional encoding / point generator."""
    _stub_packages(['mmcv.cnn', 'mmcv.cnn.bricks.transformer', 'mmcv.runner', 'mmdet.core.anchor'])
    ref = load('models/multi/seg_head/pixel_decoder.py', 'ref_seg_head.pixel_decoder')
    t = toy_pixel_decoder_parts()
    C = t['C']
    mask_feature = torch.nn.Conv2d(C, C, 1)
    with torch.no_grad():
        mask_feature.weight.copy_(t['mask_w']), mask_feature.bias.copy_(t['mask_b'])
    enc = ToyEncoder()
    fake = types.SimpleNamespace(num_encoder_levels=4, num_input_levels=4, strides=t['strides'], num_outs=4,
                                 postional_encoding=lambda m: t['pos'][tuple(m.shape[-2:])],
                                 level_encoding=torch.nn.Embedding.from_pretrained(t['level_encoding'].clone()),
                                 point_generator=ToyPointGenerator(t['strides']), lateral_convs=[], output_convs=[],
                                 mask_feature=mask_feature)
    with torch.no_grad():
        mf, feats = ref.MlvlSegPixelDecoder.forward(fake, enc, t['neck'], t['backbone'])
    torch.save(dict(source='models/multi/seg_head/pixel_decoder.py::MlvlSegPixelDecoder.forward (reference, run in place, toy parts)',
                    mask_feature=mf, feats=list(feats), encoder_saw=enc.seen), os.path.join(OUT, 'reference_seg_pixel_decoder.pt'))
    return tuple(mf.shape), [tuple(f.shape) for f in feats]


def toy_mlvl_cls_parts(C=8, B=3):
    g = torch.Generator().manual_seed(77)
    sizes = (4, 7, 14, 28)                                     # low -> high resolution, what the cls pixel decoder returns
    feats = [torch.randn(B, C, x, x, generator=g) for x in sizes]
    n = {5: 16, 6: 49, 7: sum(x * x for x in sizes), 8: 4}
    proj = {k: (torch.randn(1, v, generator=g) * 0.2, torch.randn(1, generator=g) * 0.1) for k, v in n.items()}
    return dict(C=C, B=B, feats=feats, proj=proj)


def golden_cls_mlvl():
    """models/multi/cls_head/pixel_decoder.py::MlvlClsPixelDecoder.forward and mlvl_cls_head.py::MlvlClsHead.pre_logits_1..8
    (reference, run in place, unbound on a stand-in `self`): what the shared encoder is called with for the cls task and the
    eight pooling schemes.  mmcls GlobalAveragePooling (third-party) is restated as mean over (H, W)."""
    _stub_packages(['mmcv.cnn', 'mmcv.cnn.bricks.transformer', 'mmcv.runner', 'mmdet.core.anchor', 'mmcls.models.builder',
                    'mmcls.models.heads.linear_head', 'mmcls.models.necks.gap'])
    dec = load('models/multi/cls_head/pixel_decoder.py', 'ref_cls_head.pixel_decoder')
    head = load('models/multi/cls_head/mlvl_cls_head.py', 'ref_cls_head.mlvl_cls_head')
    t = toy_pixel_decoder_parts()
    enc = ToyEncoder()
    fake = types.SimpleNamespace(num_encoder_levels=4, strides=t['strides'], num_outs=4,
                                 postional_encoding=lambda m: t['pos'][tuple(m.shape[-2:])],
                                 level_encoding=torch.nn.Embedding.from_pretrained(t['level_encoding'].clone()),
                                 point_generator=ToyPointGenerator(t['strides']))
    with torch.no_grad():
        outs = dec.MlvlClsPixelDecoder.forward(fake, enc, t['neck'])
    m = toy_mlvl_cls_parts()

    def gap(x):
        return tuple(f.mean(dim=(2, 3)) for f in x) if isinstance(x, tuple) else x.mean(dim=(2, 3))
    tokens = {}
    for k in range(1, 9):
        self_ = types.SimpleNamespace(avg_pool=gap)
        if k in m['proj']:
            w, b = m['proj'][k]
            lin = torch.nn.Linear(w.shape[1], 1)
            with torch.no_grad():
                lin.weight.copy_(w), lin.bias.copy_(b)
            self_.out_proj = lin
        with torch.no_grad():
            tokens[k] = getattr(head.MlvlClsHead, 'pre_logits_%d' % k)(self_, m['feats'])
    torch.save(dict(source='models/multi/cls_head/{pixel_decoder,mlvl_cls_head}.py (reference, run in place, toy parts)',
                    outs=list(outs), encoder_saw=enc.seen, tokens=tokens), os.path.join(OUT, 'reference_cls_mlvl.pt'))
    return [tuple(o.shape) for o in outs], {k: tuple(v.shape) for k, v in tokens.items()}


def toy_inference_parts():
    g = torch.Generator().manual_seed(91)
    L, B, Q, C = 2, 2, 9, 5
    return dict(all_cls=torch.randn(L, B, Q, C, generator=g) * 2, all_box=torch.rand(L, B, Q, 4, generator=g) * 0.5 + 0.2,
                metas=[dict(img_shape=(60, 80, 3), scale_factor=np.array([1.25, 1.2, 1.25, 1.2], dtype=np.float32)),
                       dict(img_shape=(48, 64, 3), scale_factor=np.array([0.5, 0.5, 0.5, 0.5], dtype=np.float32))],
                num_classes=C, num_query=Q, max_per_img=7,
                seg_logit=torch.randn(2, 4, 12, 16, generator=g), img=torch.zeros(2, 3, 48, 64),
                seg_metas=[dict(img_shape=(40, 60, 3), ori_shape=(80, 120, 3), flip=f, flip_direction='horizontal') for f in (False, True)])


def golden_inference():
    """the evaluation path (8f rank 4): mmdet_detr_head/deformable_detr_head.py::get_bboxes + detr_head.py::_get_bboxes_single
    (sigmoid top-k over queries x classes, cxcywh -> xyxy, clamp, rescale) and multitask_learner.py::simple_test_seg /
    inference_seg / whole_inference_seg / simple_test_det / forward_test, run in place (unbound) on stand-in objects."""
    install_mmdet_shim()
    _stub_packages(['mmcv.cnn', 'mmcv.cnn.bricks.transformer', 'mmcv.runner', 'mmdet.core', 'mmdet.models.utils', 'mmdet.models.builder',
                    'mmdet.models.dense_heads.anchor_free_head', 'mmdet.models.utils.transformer'])
    core = sys.modules['mmdet.core']

    def bbox_cxcywh_to_xyxy(bbox):          # mmdet 2.25.1 core/bbox/transforms.py
        cx, cy, w, h = bbox.split((1, 1, 1, 1), dim=-1)
        return torch.cat([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)
    core.bbox_cxcywh_to_xyxy = bbox_cxcywh_to_xyxy
    detr = load('models/multi/bbox_head/mmdet_detr_head/detr_head.py', 'ref_bbox_head.mmdet_detr_head.detr_head')
    sys.modules['ref_bbox_head.mmdet_detr_head'] = types.ModuleType('ref_bbox_head.mmdet_detr_head')
    ddetr = load('models/multi/bbox_head/mmdet_detr_head/deformable_detr_head.py', 'ref_bbox_head.mmdet_detr_head.deformable_detr_head')
    t = toy_inference_parts()
    out = dict(source='reference get_bboxes / simple_test_* (run in place, unbound)')
    for rescale in (False, True):
        fake = types.SimpleNamespace(test_cfg=dict(max_per_img=t['max_per_img']), num_query=t['num_query'], num_classes=t['num_classes'],
                                     loss_cls=types.SimpleNamespace(use_sigmoid=True))
        fake._get_bboxes_single = lambda *a, **k: detr.DETRHead._get_bboxes_single(fake, *a, **k)
        with torch.no_grad():
            res = ddetr.DeformableDETRHead.get_bboxes(fake, t['all_cls'].clone(), t['all_box'].clone(), None, None, t['metas'], rescale=rescale)
        out['det_rescale_%s' % rescale] = [(b.clone(), l.clone()) for b, l in res]
    # ---- MTL inference methods
    _stub_packages(['matplotlib.font_manager', 'matplotlib.pyplot', 'matplotlib.collections', 'matplotlib.patches',
                    'mmcv.runner', 'mmcv.cnn.bricks.transformer', 'mmcv.cnn', 'mmcls.models.utils.augment', 'mmdet.core',
                    'mmdet.core.visualization', 'mmseg.core', 'mmseg.ops', 'mtl.model.build', 'mmdet.models.utils.transformer'])

    def resize(input, size=None, scale_factor=None, mode='nearest', align_corners=None, warning=True):   # mmseg 0.28 ops/wrappers.py
        return F.interpolate(input, size, scale_factor, mode, align_corners)

    def bbox2result(bboxes, labels, num_classes):                                                  # mmdet 2.25.1 core/bbox/transforms.py
        if bboxes.shape[0] == 0:
            return [np.zeros((0, 5), dtype=np.float32) for _ in range(num_classes)]
        bboxes, labels = bboxes.detach().cpu().numpy(), labels.detach().cpu().numpy()
        return [bboxes[labels == i, :] for i in range(num_classes)]
    sys.modules['mmseg.ops'].resize = resize
    sys.modules['mmdet.core'].bbox2result = bbox2result
    ref = load('models/multi/multitask_learner.py', 'ref_multitask_learner_inf')
    ref.resize, ref.bbox2result = resize, bbox2result
    MTL = ref.MTL

    class Seg:
        align_corners = False

        def forward_test(self, neck, backbone, img_meta, enc):
            return t['seg_logit']

    class Box:
        num_classes = t['num_classes']

        def simple_test(self, feat, img_metas, rescale=False, shared_encoder=None):
            self.saw = dict(rescale=rescale, batch_input_shape=[m['batch_input_shape'] for m in img_metas])
            return out['det_rescale_%s' % rescale]

    class Fake:
        test_cfg = dict(seg=types.SimpleNamespace(mode='whole'))
        shared_encoder = None
        seg_head, bbox_head = Seg(), Box()

        def extract_feat(self, img):
            return ([img], [img])
        whole_inference_seg = MTL.whole_inference_seg
        inference_seg = MTL.inference_seg
        simple_test_seg = MTL.simple_test_seg
        simple_test_det = MTL.simple_test_det
        simple_test = MTL.simple_test
        forward_test = MTL.forward_test
    fk = Fake()
    with torch.no_grad():
        for k, meta in enumerate(t['seg_metas']):
            for rescale in (True, False):
                out['seg_%d_rescale_%s' % (k, rescale)] = fk.simple_test_seg(t['img'], [dict(meta), dict(meta)], rescale)
                out['seg_prob_%d_rescale_%s' % (k, rescale)] = fk.inference_seg(t['img'], [dict(meta), dict(meta)], rescale).clone()
        metas = [dict(m) for m in t['metas']]
        out['det_results'] = fk.forward_test('det', [t['img']], [metas], rescale=True)
        out['det_saw'] = fk.bbox_head.saw
        out['seg_via_forward_test'] = fk.forward_test(['seg', 'seg'], [t['img']], [[dict(t['seg_metas'][0])] * 2])
    torch.save(out, os.path.join(OUT, 'reference_inference.pt'))
    return sorted(out.keys())


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
    print('seg forward:', golden_seg_forward())
    print('seg pixel decoder:', golden_seg_pixel_decoder())
    print('cls mlvl:', golden_cls_mlvl())
    print('inference:', golden_inference())
    d = golden_dn_targets()
    print('dn targets:', [(c['sizes'], c['num_total_pos'], c['num_total_neg']) for c in d])
    e0 = golden_sineembed()
    print('dino decoder + head:', golden_dino_decoder_and_head())
    print('dino transformer:', golden_dino_transformer())
    print('dino loss:', golden_dino_loss())
    m = golden_multi_data_loader()
    print('multi data loader:', [(c['strategy'], c['sequence'][:4]) for c in m])
    e = golden_sineembed()
    print('sine embed cases:', [tuple(c['out'].shape) for c in e])
    s = golden_iteration_strategies()
    print('iteration strategies:', [(c['strategy'], c.get('error', 'ok')) for c in s])
    f = golden_cdn()
    print('cdn cases:', [(c['sizes'], c['out_dn_meta']) for c in f])
