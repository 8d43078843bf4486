"""Pins against fixtures produced by RUNNING THE REFERENCE'S OWN CODE (tools/make_golden.py, which executes
/root/reference/mtl/data/iteration_strategies.py and
/root/reference/models/multi/bbox_head/query_denoising.py in place; see that script for the two import shims).
Rows a22 (iteration strategies) and a15 (contrastive-denoising query generator) of SURVEY section 8a: both this
repo's host-side implementation and the oracle must reproduce the reference bit for bit on the recorded draws."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import heads as oh
from rscotr_b200.mtl.data import iteration_strategies as strategies
from rscotr_b200.models.det_head import CdnQueryGenerator

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


class _Loader:
    def __init__(self, n, m):
        self.n, self.dataset = n, list(range(m))

    def __len__(self):
        return self.n


def _cases():
    return json.load(open(os.path.join(GOLDEN, 'reference_iteration_strategies.json')))['cases']


@pytest.mark.parametrize('case', _cases(), ids=lambda c: '%s-%s' % (c['strategy'], '-'.join(map(str, c['kwargs'].values()))))
def test_iteration_strategy_matches_reference_run(case):
    assert 'error' not in case, case
    loaders = {k: _Loader(n, case['dataset_sizes'][k]) for k, n in case['loader_lengths'].items()}
    np.random.seed(case['numpy_seed'])
    strat = getattr(strategies, case['strategy'])(loaders, **case['kwargs'])
    assert [int(strat()) for _ in range(len(case['sequence']))] == case['sequence']
    assert bool(strat.should_exhaust_all_iterators) == case['should_exhaust_all_iterators']


def _cdn_cases():
    return torch.load(os.path.join(GOLDEN, 'reference_cdn.pt'), weights_only=False)['cases']


def _noise(c):
    """the reference's recorded draws in the (p, new_label per slot, rand_sign, rand_part) form both
    implementations take: it draws new labels only for the chosen slots, in order"""
    p = c['draw_p']
    chosen = torch.nonzero(p < c['noise_scale']['label'] * 0.5).view(-1)
    new_label = torch.zeros_like(p, dtype=torch.long)
    new_label[chosen] = c['draw_new_label_chosen'].long()
    return dict(p=p, new_label=new_label, rand_sign=c['draw_rand_sign'], rand_part=c['draw_rand_part'])


@pytest.mark.parametrize('idx', range(3))
def test_cdn_generator_matches_reference_run(idx):
    c = _cdn_cases()[idx]
    gen = CdnQueryGenerator(num_queries=c['num_queries'], hidden_dim=c['hidden_dim'], num_classes=20,
                            noise_scale=c['noise_scale'],
                            group_cfg=dict(dynamic=True, num_groups=None, num_dn_queries=c['num_dn_queries']))
    gen.forced_noise = _noise(c)
    emb = torch.nn.Embedding(20, c['hidden_dim'])
    with torch.no_grad():
        emb.weight.copy_(c['label_embedding'])
        q_label, q_bbox, attn_mask, dn_meta = gen(c['gt_bboxes'], c['gt_labels'], emb,
                                                  [dict(img_shape=tuple(s)) for s in c['img_shapes']])
    assert {k: int(v) for k, v in dn_meta.items()} == c['out_dn_meta']
    assert torch.equal(attn_mask, c['out_attn_mask'])
    assert torch.equal(q_label, c['out_query_label'])
    assert torch.allclose(q_bbox, c['out_query_bbox'], rtol=0, atol=1e-6)


@pytest.mark.parametrize('idx', range(3))
def test_oracle_cdn_matches_reference_run(idx):
    c = _cdn_cases()[idx]
    sd = {'bbox_head.label_embedding.weight': c['label_embedding']}
    q_label, q_bbox, attn_mask, dn_meta = oh.cdn_queries(
        sd, c['gt_bboxes'], c['gt_labels'], [dict(img_shape=tuple(s)) for s in c['img_shapes']], _noise(c),
        num_queries=c['num_queries'], num_classes=20, num_dn=c['num_dn_queries'],
        label_noise_scale=c['noise_scale']['label'], box_noise_scale=c['noise_scale']['box'])
    assert {k: int(v) for k, v in dn_meta.items()} == c['out_dn_meta']
    assert torch.equal(attn_mask, c['out_attn_mask'])
    assert torch.equal(q_label, c['out_query_label'])
    assert torch.allclose(q_bbox, c['out_query_bbox'], rtol=0, atol=1e-6)


@pytest.mark.parametrize('idx', range(2))
def test_sine_embedding_matches_reference_run(idx):
    """row a14: DinoTransformerDecoder.gen_sineembed_for_position of the reference, run in place."""
    from rscotr_b200.models.det_head import DinoTransformerDecoder
    c = torch.load(os.path.join(GOLDEN, 'reference_sineembed.pt'), weights_only=False)['cases'][idx]
    assert torch.equal(oh.sineembed(c['pos']), c['out'])                                   # oracle: same op sequence
    got = DinoTransformerDecoder.gen_sineembed_for_position(c['pos'])                      # product: cos(a) = sin(a + pi/2)
    assert got.shape == c['out'].shape and torch.allclose(got, c['out'], rtol=0, atol=2e-6)


def test_multi_data_loader_matches_reference_run():
    """row a22: batch order, dataset / task tags, re-ignition of exhausted loaders and the StopIteration of the
    exhaust-all strategies, against the reference's MultiDataLoader run in place on the same toy loaders."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), '..', 'tools'))
    from make_golden import toy_loaders
    from rscotr_b200.mtl.data.multi_data_loader import MultiDataLoader
    for c in json.load(open(os.path.join(GOLDEN, 'reference_multi_data_loader.json')))['cases']:
        loaders = toy_loaders()
        np.random.seed(c['numpy_seed'])
        loader = MultiDataLoader(loaders, getattr(strategies, c['strategy'])(loaders))
        assert len(loader) == c['length']
        seq, it = [], iter(loader)
        for _ in range(len(c['sequence'])):
            try:
                b = next(it)
            except StopIteration:
                seq.append('StopIteration')
                it = iter(loader)
                continue
            seq.append([b['dataset_name'], b['task'], [int(v) for v in b['idx']]])
        assert seq == c['sequence'], c['strategy']


@pytest.mark.parametrize('idx', range(3))
def test_dn_targets_match_reference_run(idx):
    """row a16 (denoising part): DINOHead.get_dn_target of the reference, run in place, against (a) this repo's
    ATen-path get_dn_target and (b) the assignment table the fused loss kernels consume."""
    import types
    from rscotr_b200.models.det_head import DINOHead, bbox_xyxy_to_cxcywh
    c = torch.load(os.path.join(GOLDEN, 'reference_dn_targets.pt'), weights_only=False)['cases'][idx]
    metas = [dict(img_shape=tuple(s)) for s in c['img_shapes']]
    want_labels, want_t, want_w = torch.cat(c['labels']), torch.cat(c['bbox_targets']), torch.cat(c['bbox_weights'])
    assert all(bool((w == 1).all()) for w in c['label_weights'])
    fake = types.SimpleNamespace(num_classes=20)
    got = DINOHead.get_dn_target(fake, torch.stack(c['dn_bbox_preds']), c['gt_bboxes'], c['gt_labels'], metas, c['dn_meta'])
    assert torch.equal(got['labels'], want_labels)
    assert torch.equal(got['bbox_weights'], want_w)
    assert torch.allclose(got['bbox_targets'], want_t, rtol=0, atol=1e-6)
    assert (got['num_pos'], got['num_neg']) == (c['num_total_pos'], c['num_total_neg'])
    # (b) the table of the fused path: label / target / weight of query q = those of gt table[b][q]
    table = torch.tensor(DINOHead.dn_assign_table(c['sizes'], c['dn_meta']['num_dn_group'], c['dn_meta']['pad_size']))
    gl = torch.cat(c['gt_labels'])
    gb = torch.cat([bbox_xyxy_to_cxcywh(b / torch.tensor([s[1], s[0], s[1], s[0]], dtype=torch.float32))
                    for b, s in zip(c['gt_bboxes'], c['img_shapes'])]) if len(gl) else torch.zeros(0, 4)
    flat = table.view(-1)
    lab = torch.where(flat >= 0, gl[flat.clamp(min=0)] if len(gl) else flat, torch.full_like(flat, 20))
    assert torch.equal(lab, want_labels)
    assert torch.equal((flat >= 0).float()[:, None].expand(-1, 4), want_w)
    pos = flat >= 0
    assert torch.allclose(gb[flat[pos]], want_t[pos], rtol=0, atol=1e-6)


@pytest.mark.parametrize('idx', range(3))
def test_train_step_postprocessing_matches_reference_run(idx):
    """row a21: MTL.train_step / _parse_losses of the reference (run in place, unbound) against this repo's
    MTL._finish (generic dict path and the packed path the fused loss kernels use) and the oracle."""
    import types
    from rscotr_b200.models.bricks import PackedLosses
    from rscotr_b200.models.mtl import MTL
    c = torch.load(os.path.join(GOLDEN, 'reference_train_step.pt'), weights_only=False)['cases'][idx]
    fake = types.SimpleNamespace(task_weight=dict(cls=1, det=1, seg=0.1), _reduce_log_vars=MTL._reduce_log_vars)
    fake._parse_losses = lambda losses: MTL._parse_losses(fake, losses)
    n = c['num_samples']
    data = dict(task=c['task'], dataset_name=c['dataset_name'], img_metas=[{}] * n)
    out = MTL._finish(fake, {k: (list(v) if isinstance(v, list) else v.clone()) for k, v in c['losses'].items()}, data)
    assert out['num_samples'] == n
    assert abs(float(out['loss']) - float(c['loss'])) <= 1e-6 * max(1.0, abs(float(c['loss'])))
    got = dict(out['log_vars'].items())
    assert list(got.keys()) == list(c['log_vars'].keys())
    for k, v in c['log_vars'].items():
        assert abs(got[k] - v) <= 1e-6 * max(1.0, abs(v)), (k, got[k], v)
    # oracle
    loss_o, log_o = oh.parse_losses({k: (list(v) if isinstance(v, list) else v.clone()) for k, v in c['losses'].items()},
                                    fake.task_weight[c['task']])
    assert abs(float(loss_o) - float(c['loss'])) <= 1e-6 * max(1.0, abs(float(c['loss'])))
    # packed path (scalar terms only)
    if all(torch.is_tensor(v) and v.dim() == 0 for v in c['losses'].values()):
        keys = list(c['losses'].keys())
        out2 = MTL._finish(fake, PackedLosses(keys, torch.stack([c['losses'][k] for k in keys])), data)
        assert abs(float(out2['loss']) - float(c['loss'])) <= 1e-6 * max(1.0, abs(float(c['loss'])))
        assert {k: round(v, 6) for k, v in out2['log_vars'].items()} == {k: round(v, 6) for k, v in got.items()}


def test_seg_forward_head_matches_reference_run():
    """row a18: Mask2FormerHead.forward_head of the reference (run in place, unbound) against this repo's method
    (CPU through the test shim of the CUDA ops): mask prediction and the boolean attention mask."""
    import types
    from rscotr_b200.models import bricks
    from rscotr_b200.models.seg_head import Mask2FormerHead
    from tests.cpu_ops_shim import cpu_ops
    c = torch.load(os.path.join(GOLDEN, 'reference_seg_forward_head.pt'), weights_only=False)
    C = c['decoder_out'].shape[-1]
    post_norm = bricks.LayerNorm(C)
    mask_embed = torch.nn.Sequential(bricks.Linear(C, C), torch.nn.ReLU(), bricks.Linear(C, C), torch.nn.ReLU(),
                                     bricks.Linear(C, C))
    post_norm.load_state_dict({k[len('post_norm.'):]: v for k, v in c['state'].items() if k.startswith('post_norm.')})
    mask_embed.load_state_dict({k[len('mask_embed.'):]: v for k, v in c['state'].items() if k.startswith('mask_embed.')})
    fake = types.SimpleNamespace(scheme=2, num_heads=c['num_heads'], mask_embed=mask_embed,
                                 transformer_decoder=types.SimpleNamespace(post_norm=post_norm))
    with cpu_ops(), torch.no_grad():
        seg_mask, attn_mask = Mask2FormerHead.forward_head(fake, c['decoder_out'], c['mask_feature'], tuple(c['target_size']))
    assert torch.allclose(seg_mask, c['seg_mask'], rtol=1e-5, atol=1e-5)
    assert attn_mask.shape == c['attn_mask'].shape and attn_mask.dtype == torch.bool
    assert float((attn_mask != c['attn_mask']).float().mean()) <= 1e-3      # (values within 1e-6 of the 0.5 threshold)
