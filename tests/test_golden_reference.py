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
# the CUDA variants compare fp32 kernels with CPU-generated fixtures: no TF32 in the library GEMMs / convs
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEVICES = ['cpu', pytest.param('cuda', marks=pytest.mark.gpu)]
DEVICES_NEW = DEVICES       # (the CUDA variants of the inference goldens were validated on a B200 in round 2)


def _to(obj, device):
    if torch.is_tensor(obj) or isinstance(obj, torch.nn.Module):
        return obj.to(device)
    if isinstance(obj, dict):
        return {k: _to(v, device) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to(v, device) for v in obj)
    return obj


def _ctx(device):
    """CPU: the CUDA ops are substituted by the oracle (test shim); CUDA: the real kernels through the C ABI."""
    import contextlib
    from tests.cpu_ops_shim import cpu_ops
    return cpu_ops() if device == 'cpu' else contextlib.nullcontext()


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


@pytest.mark.parametrize('device', DEVICES)
def test_seg_forward_head_matches_reference_run(device):
    """row a18: Mask2FormerHead.forward_head of the reference (run in place, unbound) against this repo's method
    (CPU through the test shim of the CUDA ops): mask prediction and the boolean attention mask."""
    import types
    from rscotr_b200.models import bricks
    from rscotr_b200.models.seg_head import Mask2FormerHead
    from tests.cpu_ops_shim import cpu_ops
    c = _to(torch.load(os.path.join(GOLDEN, 'reference_seg_forward_head.pt'), weights_only=False), device)
    C = c['decoder_out'].shape[-1]
    post_norm = bricks.LayerNorm(C)
    mask_embed = torch.nn.Sequential(bricks.Linear(C, C), torch.nn.ReLU(), bricks.Linear(C, C), torch.nn.ReLU(),
                                     bricks.Linear(C, C))
    post_norm.load_state_dict({k[len('post_norm.'):]: v for k, v in c['state'].items() if k.startswith('post_norm.')})
    mask_embed.load_state_dict({k[len('mask_embed.'):]: v for k, v in c['state'].items() if k.startswith('mask_embed.')})
    fake = types.SimpleNamespace(scheme=2, num_heads=c['num_heads'], mask_embed=mask_embed.to(device),
                                 transformer_decoder=types.SimpleNamespace(post_norm=post_norm.to(device)))
    with _ctx(device), torch.no_grad():
        seg_mask, attn_mask = Mask2FormerHead.forward_head(fake, c['decoder_out'], c['mask_feature'], tuple(c['target_size']))
    assert torch.allclose(seg_mask, c['seg_mask'], rtol=1e-5, atol=1e-5)
    assert attn_mask.dtype == torch.bool
    if attn_mask.dim() == 4:    # the compact (B, 1, Q, K) mask (default since round 2): the reference's layout repeats it
        attn_mask = attn_mask.repeat(1, c['num_heads'], 1, 1).flatten(0, 1)      # over the heads -> (B * heads, Q, K)
    assert attn_mask.shape == c['attn_mask'].shape
    assert float((attn_mask != c['attn_mask']).float().mean()) <= 1e-3      # (values within 1e-6 of the 0.5 threshold)


def _tools():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), '..', 'tools'))
    import make_golden
    return make_golden


@pytest.mark.parametrize('device', DEVICES)
def test_seg_head_forward_control_flow_matches_reference_run(device, monkeypatch):
    """row a18: Mask2FormerHead.forward of the reference (level cycling, reset of fully masked rows, forward_head after
    every layer), run in place with toy decoder layers, against this repo's forward with the SAME toy parts.  The toy
    layers consume the reference's (B * heads, Q, K) mask layout, so the compact-mask default is switched off here
    (its equality with that layout is test_host_units.py::test_compact_attention_mask_is_equivalent)."""
    import types
    from rscotr_b200.models import bricks, seg_head
    from rscotr_b200.models.seg_head import Mask2FormerHead
    monkeypatch.setattr(seg_head, '_COMPACT_ATTN_MASK', False)
    from tests.cpu_ops_shim import cpu_ops
    mg = _tools()
    t = _to(mg.toy_seg_parts(), device)
    want = torch.load(os.path.join(GOLDEN, 'reference_seg_forward.pt'), weights_only=False)['out'].to(device)
    C = t['C']
    post_norm = bricks.LayerNorm(C)
    mask_embed = torch.nn.Sequential(bricks.Linear(C, C), torch.nn.ReLU(), bricks.Linear(C, C), torch.nn.ReLU(),
                                     bricks.Linear(C, C))
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
        query_feat=emb(t['query_feat']), query_embed=emb(t['query_embed']), mask_embed=mask_embed.to(device),
        transformer_decoder=types.SimpleNamespace(post_norm=post_norm.to(device),
                                                  layers=[mg.ToyDecoderLayer() for _ in range(9)]))
    fake.forward_head = lambda *a, **k: Mask2FormerHead.forward_head(fake, *a, **k)
    with _ctx(device), torch.no_grad():
        got = Mask2FormerHead.forward(fake, None, None, None, [{}] * t['B'])
    assert got.shape == want.shape
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-4), float((got - want).abs().max())


@pytest.mark.parametrize('device', DEVICES)
def test_dino_decoder_and_head_control_flow_match_reference_run(device):
    """rows a14 / a13: DinoTransformerDecoder.forward and DINOHead.forward of the reference, run in place with toy
    decoder layers / a toy transformer, against this repo's methods with the SAME toy parts."""
    import types
    from rscotr_b200.models import bricks
    from rscotr_b200.models.det_head import DINOHead, DinoTransformerDecoder
    from tests.cpu_ops_shim import cpu_ops
    mg = _tools()
    t = _to(mg.toy_det_parts(), device)
    c = _to(torch.load(os.path.join(GOLDEN, 'reference_dino_decoder_head.pt'), weights_only=False), device)
    C, L = t['C'], t['L']
    rph = [(c['rph0'], t['ref_point_head'][0][1]), t['ref_point_head'][1]]
    norm = bricks.LayerNorm(C)
    with torch.no_grad():
        norm.weight.copy_(t['norm'][0]), norm.bias.copy_(t['norm'][1])
    reg = [mg._mlp_from(ws).to(device) for ws in t['reg']]
    cls = [mg._mlp_from([w]).to(device) for w in t['cls']]
    dec = types.SimpleNamespace(layers=[mg.ToyDecoderLayer() for _ in range(L)], ref_point_head=mg._mlp_from(rph).to(device),
                                norm=norm.to(device), return_intermediate=True,
                                gen_sineembed_for_position=DinoTransformerDecoder.gen_sineembed_for_position)
    with _ctx(device), torch.no_grad():
        hs, refs = DinoTransformerDecoder.forward(dec, t['query'], None, t['memory'], reference_points=t['reference_points'],
                                                  valid_ratios=t['valid_ratios'], reg_branches=reg)
    assert torch.allclose(hs, c['hs'], rtol=1e-4, atol=1e-5), float((hs - c['hs']).abs().max())
    assert torch.allclose(refs, c['refs'], rtol=1e-4, atol=1e-5)
    seen = {}

    def toy_transformer(mlvl_feats, mlvl_masks, query_embeds, mlvl_pos, dn_label_query, dn_bbox_query, attn_mask, encoder,
                        reg_branches=None, cls_branches=None, **kw):
        seen['masks'] = [m.clone() for m in mlvl_masks]
        return c['hs'], c['refs'], t['topk_score'], t['topk_anchor']
    head = types.SimpleNamespace(transformer=toy_transformer, positional_encoding=lambda m: m.float().unsqueeze(1),
                                 with_box_refine=True, as_two_stage=True, reg_branches=reg, cls_branches=cls,
                                 label_embedding=torch.nn.Embedding(t['classes'], C).to(device))
    with _ctx(device), torch.no_grad():
        oc, ob, ts, ta = DINOHead.forward(head, None, t['feats'], c['metas'], torch.zeros(t['B'], t['pad'], C, device=device),
                                          torch.zeros(t['B'], t['pad'], 4, device=device), None)
    assert all(torch.equal(a, b) for a, b in zip(seen['masks'], c['masks']))
    assert torch.allclose(oc, c['outputs_classes'], rtol=1e-5, atol=1e-6)
    assert torch.allclose(ob, c['outputs_coords'], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('device', DEVICES)
def test_dino_transformer_control_flow_matches_reference_run(device):
    """row a13: DinoTransformer.forward of the reference (flattening, level embeddings, two-stage top-k proposals,
    dn / matching query concatenation), run in place with a toy encoder / decoder, against this repo's forward with
    the SAME toy parts -- incl. what is handed to the decoder (queries, reference points, memory, padding mask)."""
    import types
    from rscotr_b200.models import bricks
    from rscotr_b200.models.det_head import DinoTransformer
    from tests.cpu_ops_shim import cpu_ops
    mg = _tools()
    t = _to(mg.toy_two_stage_parts(), device)
    c = _to(torch.load(os.path.join(GOLDEN, 'reference_dino_transformer.pt'), weights_only=False), device)
    C = t['C']
    enc_out = mg._mlp_from([t['enc_output']], cls=bricks.Linear).to(device)
    enc_norm = bricks.LayerNorm(C).to(device)
    with torch.no_grad():
        enc_norm.weight.copy_(t['enc_norm'][0]), enc_norm.bias.copy_(t['enc_norm'][1])
    dec = mg.ToyDecoder()
    fake = types.SimpleNamespace(as_two_stage=True, level_embeds=t['level_embeds'], decoder=dec, two_stage_num_proposals=t['K'],
                                 query_embed=torch.nn.Embedding.from_pretrained(t['query_embed'].clone()).to(device),
                                 enc_output=enc_out, enc_output_norm=enc_norm, get_valid_ratio=DinoTransformer.get_valid_ratio,
                                 get_reference_points=DinoTransformer.get_reference_points,
                                 proposal_grid=DinoTransformer.proposal_grid)
    fake.gen_encoder_output_proposals = lambda *a, **k: DinoTransformer.gen_encoder_output_proposals(fake, *a, **k)
    fake._geometry = lambda *a: DinoTransformer._geometry(fake, *a)
    cls = [None, mg._mlp_from([t['cls']]).to(device)]
    reg = [None, mg._mlp_from(t['reg']).to(device)]
    with _ctx(device), torch.no_grad():
        out = DinoTransformer.forward(fake, t['feats'], t['masks'], None, t['pos'], t['dn_label'], t['dn_bbox'], None,
                                      mg.toy_encoder, reg_branches=reg, cls_branches=cls)
    for got, want in zip(out, c['out']):
        assert got.shape == want.shape and torch.allclose(got, want, rtol=1e-5, atol=1e-5), float((got - want).abs().max())
    for k in ('query', 'reference_points', 'value'):
        assert torch.allclose(dec.seen[k], c['decoder_saw'][k], rtol=1e-5, atol=1e-5), k
    assert torch.equal(dec.seen['key_padding_mask'], c['decoder_saw']['key_padding_mask'])


@pytest.mark.parametrize('device', DEVICES)
def test_seg_pixel_decoder_control_flow_matches_reference_run(device):
    """row a17: MlvlSegPixelDecoder.forward of the reference, run in place with a toy encoder / positional encoding /
    point generator, against this repo's forward with the SAME toy parts -- incl. what the shared encoder is
    called with (inputs, level positional encodings, reference points, level shapes / start indices)."""
    import types
    from rscotr_b200.models.seg_head import MlvlSegPixelDecoder
    from tests.cpu_ops_shim import cpu_ops
    mg = _tools()
    t = _to(mg.toy_pixel_decoder_parts(), device)
    c = _to(torch.load(os.path.join(GOLDEN, 'reference_seg_pixel_decoder.pt'), weights_only=False), device)
    C = t['C']
    mask_feature = torch.nn.Conv2d(C, C, 1).to(device)
    with torch.no_grad():
        mask_feature.weight.copy_(t['mask_w']), mask_feature.bias.copy_(t['mask_b'])
    enc = mg.ToyEncoder()
    fake = types.SimpleNamespace(num_encoder_levels=4, num_input_levels=4, strides=t['strides'], num_outs=4,
                                 postional_encoding=lambda m: t['pos'][tuple(m.shape[-2:])],
                                 level_encoding=torch.nn.Embedding.from_pretrained(t['level_encoding'].clone()).to(device),
                                 lateral_convs=[], output_convs=[], mask_feature=mask_feature)
    with _ctx(device), torch.no_grad():
        mf, feats = MlvlSegPixelDecoder.forward(fake, enc, t['neck'], t['backbone'])
    assert torch.allclose(mf, c['mask_feature'], rtol=1e-5, atol=1e-6)
    assert len(feats) == len(c['feats']) and all(torch.allclose(a, b, rtol=1e-5, atol=1e-6) for a, b in zip(feats, c['feats']))
    for k in ('query', 'query_pos', 'reference_points'):
        assert torch.allclose(enc.seen[k], c['encoder_saw'][k], rtol=1e-6, atol=1e-6), k
    assert torch.equal(enc.seen['spatial_shapes'], c['encoder_saw']['spatial_shapes'])
    assert torch.equal(enc.seen['level_start_index'], c['encoder_saw']['level_start_index'])


@pytest.mark.parametrize('device', DEVICES)
def test_cls_mlvl_head_matches_reference_run(device):
    """8f rank 4: MlvlClsPixelDecoder.forward and MlvlClsHead.pre_logits_1..8 of the reference, run in place
    (tools/make_golden.py::golden_cls_mlvl), against this repo's classes with the SAME toy parts."""
    import types
    from rscotr_b200.models.cls_head import GlobalAveragePooling, MlvlClsHead, MlvlClsPixelDecoder
    mg = _tools()
    t = _to(mg.toy_pixel_decoder_parts(), device)
    c = _to(torch.load(os.path.join(GOLDEN, 'reference_cls_mlvl.pt'), weights_only=False), device)
    enc = mg.ToyEncoder()
    fake = types.SimpleNamespace(num_encoder_levels=4, strides=t['strides'], num_outs=4,
                                 postional_encoding=lambda m: t['pos'][tuple(m.shape[-2:])],
                                 level_encoding=torch.nn.Embedding.from_pretrained(t['level_encoding'].clone()).to(device))
    with _ctx(device), torch.no_grad():
        outs = MlvlClsPixelDecoder.forward(fake, enc, t['neck'])
    assert len(outs) == len(c['outs']) and all(torch.allclose(a, b, rtol=1e-5, atol=1e-6) for a, b in zip(outs, c['outs']))
    for k in ('query', 'query_pos', 'reference_points'):
        assert torch.allclose(enc.seen[k], c['encoder_saw'][k], rtol=1e-6, atol=1e-6), k
    assert torch.equal(enc.seen['spatial_shapes'], c['encoder_saw']['spatial_shapes'])
    assert torch.equal(enc.seen['level_start_index'], c['encoder_saw']['level_start_index'])
    m = _to(mg.toy_mlvl_cls_parts(), device)
    for k in range(1, 9):
        self_ = types.SimpleNamespace(scheme=k, avg_pool=GlobalAveragePooling())
        if k in m['proj']:
            w, b = m['proj'][k]
            lin = torch.nn.Linear(w.shape[1], 1).to(device)
            with torch.no_grad():
                lin.weight.copy_(w), lin.bias.copy_(b)
            self_.out_proj = lin
        with _ctx(device), torch.no_grad():
            tok = MlvlClsHead.pre_logits(self_, m['feats'])
        assert tok.shape == c['tokens'][k].shape and torch.allclose(tok, c['tokens'][k], rtol=1e-5, atol=1e-5), k


@pytest.mark.parametrize('device', DEVICES_NEW)
def test_inference_path_matches_reference_run(device):
    """8f rank 4 (evaluation path): the reference's get_bboxes / _get_bboxes_single and MTL.simple_test_seg / inference_seg /
    whole_inference_seg / simple_test_det / forward_test, run in place (tools/make_golden.py::golden_inference), against
    this repo's DINOHead.get_bboxes and MTL methods on the same inputs."""
    import types
    import numpy as np
    from rscotr_b200.models.det_head import DINOHead
    from rscotr_b200.models.mtl import MTL
    mg = _tools()
    t = mg.toy_inference_parts()
    c = torch.load(os.path.join(GOLDEN, 'reference_inference.pt'), weights_only=False)
    for rescale in (False, True):
        fake = types.SimpleNamespace(test_cfg=dict(max_per_img=t['max_per_img']), num_query=t['num_query'], num_classes=t['num_classes'])
        with torch.no_grad():
            res = DINOHead.get_bboxes(fake, t['all_cls'].to(device), t['all_box'].to(device), None, None, t['metas'], rescale=rescale)
        for (b, l), (wb, wl) in zip(res, c['det_rescale_%s' % rescale]):
            assert torch.equal(l.cpu(), wl) and torch.allclose(b.cpu(), wb, rtol=1e-5, atol=1e-5)

    class Seg:
        align_corners = False

        def forward_test(self, neck, backbone, img_meta, enc):
            return t['seg_logit'].to(device)

    class Box:
        num_classes = t['num_classes']

        def simple_test(self, feat, img_metas, rescale=False, shared_encoder=None):
            self.saw = dict(rescale=rescale, batch_input_shape=[m['batch_input_shape'] for m in img_metas])
            return c['det_rescale_%s' % rescale]

    class Fake:
        test_cfg = dict(seg=dict(mode='whole'))
        shared_encoder = None
        seg_head, bbox_head = Seg(), Box()

        def extract_feat(self, img):
            return ([img], [img])
        whole_inference_seg, inference_seg, simple_test_seg = MTL.whole_inference_seg, MTL.inference_seg, MTL.simple_test_seg
        simple_test_det, simple_test, forward_test = MTL.simple_test_det, MTL.simple_test, MTL.forward_test
    fk = Fake()
    img = t['img'].to(device)
    with _ctx(device), torch.no_grad():
        for k, meta in enumerate(t['seg_metas']):
            for rescale in (True, False):
                prob = fk.inference_seg(img, [dict(meta), dict(meta)], rescale)
                assert torch.allclose(prob.cpu(), c['seg_prob_%d_rescale_%s' % (k, rescale)], rtol=1e-4, atol=1e-5), (k, rescale)
                pred = fk.simple_test_seg(img, [dict(meta), dict(meta)], rescale)
                want = c['seg_%d_rescale_%s' % (k, rescale)]
                assert len(pred) == len(want) == 2
                for a, b in zip(pred, want):
                    assert a.shape == b.shape and (a != b).mean() < 0.002          # (argmax ties at fp32 rounding level)
        det = fk.forward_test('det', [img], [[dict(m) for m in t['metas']]], rescale=True)
        assert fk.bbox_head.saw == c['det_saw']
        for a, b in zip(det, c['det_results']):
            assert len(a) == len(b) == t['num_classes'] and all(np.allclose(x, y) for x, y in zip(a, b))
        seg = fk.forward_test(['seg', 'seg'], [img], [[dict(t['seg_metas'][0])] * 2])
        assert all((a != b).mean() < 0.002 for a, b in zip(seg, c['seg_via_forward_test']))


def _dino_head():
    import rscotr_b200.models  # noqa: F401
    from rscotr_b200.config import Config
    from rscotr_b200.models.det_head import DINOHead
    cfg = Config.fromfile(os.path.join(os.path.dirname(GOLDEN), '..', 'configs', 'multi', 'cotrain_swin-t_800.py'))
    hc = dict(cfg.model.bbox_head)
    hc.pop('type')
    hc.update(train_cfg=cfg.model.train_cfg.get('det'), test_cfg=cfg.model.test_cfg.get('det'), num_query=12)
    return DINOHead(**hc)


def _loss_case(device='cpu'):
    mg = _tools()
    t = mg.toy_loss_parts()
    want = json.load(open(os.path.join(GOLDEN, 'reference_dino_loss.json')))
    metas = [dict(img_shape=tuple(s)) for s in t['img_shapes']]
    args = [t['all_cls'].to(device), t['all_box'].to(device), t['enc_cls'].to(device), t['enc_box'].to(device),
            [b.to(device) for b in t['gt_bboxes']], [l.to(device) for l in t['gt_labels']], metas, t['dn_meta']]
    return args, want


def test_dino_loss_flow_matches_reference_run():
    """row a16: DINOHead.loss / DETRHead.loss_single of the reference, run in place (third-party assigner / sampler /
    loss modules restated from the oracle), against this repo's ATen + scipy path (CPU) and the oracle: same 21
    keys in the same order, same values."""
    from tests.cpu_ops_shim import cpu_ops
    args, want = _loss_case()
    head = _dino_head()
    head.fused_loss = False
    with cpu_ops(), torch.no_grad():
        got = head.loss(*args)
    assert list(got.keys()) == want['keys']
    for k, v in want['losses'].items():
        assert abs(float(got[k]) - v) <= 2e-5 * max(1.0, abs(v)), (k, float(got[k]), v)
    ora = oh.det_loss(*args[:4], args[4], args[5], args[6], args[7])
    for k, v in want['losses'].items():
        assert abs(float(ora[k]) - v) <= 1e-6 * max(1.0, abs(v)), (k, float(ora[k]), v)


@pytest.mark.gpu
def test_dino_loss_cuda_kernels_match_reference_run_golden():
    """the CUDA path (rsc_det_match + rsc_det_loss_fwd through the C ABI) against the SAME committed fixture."""
    args, want = _loss_case('cuda')
    head = _dino_head().cuda()
    assert head.fused_loss
    with torch.no_grad():
        got = head.loss(*args)
    assert list(got.keys()) == want['keys']
    for k, v in want['losses'].items():
        assert abs(float(got[k]) - v) <= 5e-5 * max(1.0, abs(v)), (k, float(got[k]), v)
