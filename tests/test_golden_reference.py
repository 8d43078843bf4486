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
