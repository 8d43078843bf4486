"""UPerNet / FCN heads and the single-task EncoderDecoder on CPU: against the functional
oracle (oracle/uper.py), against the independent HF implementation, and one engine step."""
import pytest
import torch

import rscotr_b200.models  # noqa: F401
from oracle import uper as O
from rscotr_b200.config import Config, MODELS
from rscotr_b200.models.uper_head import FCNHead, UPerHead
from tests.cpu_ops_shim import cpu_ops

CH = [8, 16, 32, 64]


def _feats(B=2, base=24, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(B, c, base >> i, base >> i, generator=g) for i, c in enumerate(CH)]


def _heads(seed=0):
    torch.manual_seed(seed)
    kw = dict(num_classes=6, norm_cfg=dict(type='BN'), ignore_index=5, dropout_ratio=0.1)
    up = UPerHead(CH, 16, loss_decode=dict(loss_weight=1.0), **kw)
    aux = FCNHead(CH[2], 12, in_index=2, num_convs=1, concat_input=False, loss_decode=dict(loss_weight=0.4), **kw)
    for h in (up, aux):                                    # non-trivial BN statistics / affine
        for m in h.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.5)
                m.running_var.uniform_(0.5, 2.0)
                torch.nn.init.normal_(m.weight, 1.0, 0.2)
                torch.nn.init.normal_(m.bias, 0.0, 0.2)
    return up, aux


@pytest.mark.parametrize('training', [False, True])
def test_heads_match_the_oracle(training):
    up, aux = _heads()
    feats = _feats()
    for h in (up, aux):
        h.train(training)
        if h.dropout is not None:
            h.dropout.p = 0.0                                # (the oracle has no dropout)
    with cpu_ops():
        got_u, got_a = up(feats), aux(feats)
    want_u = O.uper_head(up.state_dict(), '', feats, training=training)
    want_a = O.fcn_head(aux.state_dict(), '', feats, in_index=2, num_convs=1, training=training)
    assert got_u.shape == (2, 6, 24, 24) and got_a.shape == (2, 6, 6, 6)
    assert torch.allclose(got_u, want_u, atol=2e-5, rtol=1e-4)
    assert torch.allclose(got_a, want_a, atol=2e-5, rtol=1e-4)
    # losses (mmseg BaseDecodeHead.losses): weight, ignore index, accuracy over the non-ignored pixels
    label = torch.randint(0, 6, (2, 1, 96, 96), generator=torch.Generator().manual_seed(1))
    with cpu_ops():
        lu, la = up.losses(got_u, label), aux.losses(got_a, label)
    wu, wa = O.seg_losses(want_u, label, 5, 1.0), O.seg_losses(want_a, label, 5, 0.4)
    for a, b in ((lu, wu), (la, wa)):
        assert torch.allclose(a['loss_ce'], b['loss_ce'], rtol=1e-4) and torch.allclose(a['acc_seg'], b['acc_seg'], rtol=1e-4)


def test_uper_head_matches_the_hf_implementation():
    hf = pytest.importorskip('transformers.models.upernet.modeling_upernet')
    from transformers import UperNetConfig
    cfg = UperNetConfig(hidden_size=16, pool_scales=[1, 2, 3, 6], num_labels=6, auxiliary_channels=12, auxiliary_num_convs=1,
                        auxiliary_concat_input=False, auxiliary_in_channels=CH[2])
    up, aux = _heads()
    up.eval(), aux.eval()
    ref_u, ref_a = hf.UperNetHead(cfg, CH).eval(), hf.UperNetFCNHead(cfg, CH, in_index=2).eval()

    def port(sd):        # mmseg names -> HF names
        out = {}
        for k, v in sd.items():
            k = k.replace('.bn.', '.batch_norm.').replace('conv_seg.', 'classifier.')
            out[k] = v
        return out
    missing = ref_u.load_state_dict(port(up.state_dict()), strict=False)
    # (HF registers every pooling block twice; the `blocks.*` aliases share the tensors loaded through `psp_modules.{i}.1`)
    assert all('.blocks.' in k for k in missing.missing_keys) and not missing.unexpected_keys, missing
    res = ref_a.load_state_dict(port(aux.state_dict()), strict=False)
    assert not res.missing_keys and not res.unexpected_keys, res
    feats = _feats(seed=3)
    with cpu_ops():
        got_u, got_a = up(feats), aux(feats)
    assert torch.allclose(got_u, ref_u(feats), atol=2e-5, rtol=1e-4)
    assert torch.allclose(got_a, ref_a(feats), atol=2e-5, rtol=1e-4)


@pytest.mark.timeout(600)
def test_encoder_decoder_trains_through_the_step_engine():
    from rscotr_b200.mtl.data import build_datasets
    from rscotr_b200.mtl.engine import StepEngine
    cfg = Config.fromfile('configs/seg/upernet_swin-b_512_potsdam.py')
    m = cfg.model
    m.backbone.embed_dims, m.backbone.depths, m.backbone.num_heads, m.backbone.drop_path_rate = 16, [2, 2, 2, 2], [1, 2, 4, 8], 0.0
    m.decode_head.in_channels, m.decode_head.channels = [16, 32, 64, 128], 32
    m.auxiliary_head.in_channels, m.auxiliary_head.channels = 64, 16
    torch.manual_seed(0)
    model = MODELS.build(m)
    model.init_weights()
    keys = list(model.state_dict().keys())
    assert 'decode_head.psp_modules.3.1.bn.running_var' in keys and 'auxiliary_head.convs.0.conv.weight' in keys
    eng = StepEngine(model, dict(cfg.optimizer), device='cpu', compute_dtype=torch.float32, use_graphs=False)
    ds = build_datasets({'potsdam': dict(task='seg')}, synthetic=dict(img_size=(64, 64), seg=dict(num_classes=6)))['potsdam']
    batch = ds.make_batch(2, torch.Generator().manual_seed(0), pin=False)
    batch.update(task='seg', dataset_name='potsdam')
    model.train()
    with cpu_ops():
        before = model.decode_head.conv_seg.weight.detach().clone()
        out = eng.train_iter(batch)
        logs = dict(out['log_vars'].items())
        assert set(logs) == {'seg.potsdam.decode.loss_ce', 'seg.potsdam.decode.acc_seg', 'seg.potsdam.aux.loss_ce',
                             'seg.potsdam.aux.acc_seg', 'seg.potsdam.loss'}
        assert abs(logs['seg.potsdam.loss'] - logs['seg.potsdam.decode.loss_ce'] - logs['seg.potsdam.aux.loss_ce']) < 1e-5
        assert not torch.equal(before, model.decode_head.conv_seg.weight)
        model.eval()
        pred = model(img=[batch['img']], img_metas=[batch['img_metas']], return_loss=False)
    assert len(pred) == 2 and pred[0].shape == (64, 64) and pred[0].max() < 6


# ------------------------------------------------------------------------------------------ GPU (C-ABI kernels)
def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


@pytest.mark.gpu
def test_gpu_heads_match_the_oracle_fwd_bwd():
    """UPerHead / FCNHead on CUDA (rsc_bilinear_* resizes, rsc_upsample_ce_* loss) against the CPU oracle, fp32,
    tolerance 1e-3 relative (the north star's fp32 bound)."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    up, aux = _heads()
    for h in (up, aux):
        h.train()
        h.dropout.p = 0.0
    feats = _feats()
    label = torch.randint(0, 6, (2, 1, 96, 96), generator=torch.Generator().manual_seed(1))
    sd_u = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in up.state_dict().items()}
    sd_a = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in aux.state_dict().items()}
    wu = O.seg_losses(O.uper_head(sd_u, '', feats, training=True), label, 5, 1.0)
    wa = O.seg_losses(O.fcn_head(sd_a, '', feats, in_index=2, num_convs=1, training=True), label, 5, 0.4)
    (wu['loss_ce'] + wa['loss_ce']).backward()
    up.cuda(), aux.cuda()
    cf = [f.cuda() for f in feats]
    gu, ga = up.forward_train(cf, None, label.cuda()), aux.forward_train(cf, None, label.cuda())
    (gu['loss_ce'] + ga['loss_ce']).backward()
    for got, want in ((gu, wu), (ga, wa)):
        assert abs(float(got['loss_ce']) - float(want['loss_ce'])) < 1e-3 * abs(float(want['loss_ce']))
        assert abs(float(got['acc_seg']) - float(want['acc_seg'])) < 1e-2
    for head, sd in ((up, sd_u), (aux, sd_a)):
        for n, p in head.named_parameters():
            assert _rel(p.grad, sd[n].grad) < 2e-3, (n, _rel(p.grad, sd[n].grad))


@pytest.mark.gpu
def test_gpu_encoder_decoder_graph_replay_matches_eager():
    from rscotr_b200.mtl.data import build_datasets
    from rscotr_b200.mtl.engine import StepEngine
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    finals = []
    for use_graphs in (False, True):
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
        losses = [float(eng.train_iter(batch)['loss'].detach()) for _ in range(5)]
        if use_graphs:
            assert any('gA' in st for st in eng._graphs.values()) and eng.graph_failures == 0
        finals.append((losses, {n: p.detach().clone() for n, p in model.named_parameters()}))
    (l0, p0), (l1, p1) = finals
    assert l0[-1] < l0[0]                                      # it learns the fixed batch
    for a, b in zip(l0, l1):
        assert abs(a - b) <= 2e-3 * max(1.0, abs(a)), (l0, l1)
    for n in p0:      # (absolute floor: biases in front of a norm / the k-bias of softmax attention stay ~1e-13 noise)
        assert float((p1[n] - p0[n]).norm()) <= 1e-3 * float(p0[n].norm()) + 1e-8, n
