"""world_size-2 (gloo, CPU) test of the data-parallel logic of the step engine: flat-range
gradient all-reduce per task, packed loss-factor / log-var reductions, identical replicas.
The CUDA ops are substituted by the oracle shim (tests/cpu_ops_shim.py); NCCL replaces gloo on
the GPU box, the code path is the same."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    import rscotr_b200.models  # noqa: F401
    from rscotr_b200.config import MODELS
    from rscotr_b200.mtl.data import build_datasets
    from rscotr_b200.mtl.engine import StepEngine
    from tests.cpu_ops_shim import cpu_ops
    from tests.test_host_model import small_cfg
    torch.manual_seed(0)                                   # same weights on every rank
    model = MODELS.build(small_cfg().model)
    model.init_weights()
    model.train()
    os.environ['RSC_MIN_BUCKET'] = '1'                     # (tiny model: let every in-backward bucket fire)
    eng = StepEngine(model, dict(type='AdamW', lr=1e-3, weight_decay=1e-4), grad_clip=dict(max_norm=0.1, norm_type=2),
                     device='cpu', compute_dtype=torch.float32, use_graphs=False)
    assert eng.world == world
    res = {}
    with cpu_ops():
        for task in ('cls', 'det', 'seg'):
            ds = build_datasets({'x': dict(task=task)}, synthetic=dict(img_size=(64, 64), det=dict(num_boxes=2 + rank)))['x']
            batch = ds.make_batch(2, torch.Generator().manual_seed(100 + rank), pin=False)     # different shard per rank
            batch.update(task=task, dataset_name='x')
            # local (un-reduced) gradient of this rank, for the check below
            model.zero_grad(set_to_none=False)
            local = model.train_step(dict(batch), None)
            local['loss'].backward()
            eng._collect_grads()
            g_local = eng.flat_grad.clone()
            out = eng.train_iter(batch)            # first iteration of the task: ranges found, plain exchange
            res[task] = dict(g_local=g_local, loss=float(out['loss'].detach()), log=dict(out['log_vars'].items()),
                             ranges=list(eng._task_ranges[task]))
            # second iteration: the overlapped exchange (buckets all-reduced from inside backward).  Its result must
            # be the cross-rank mean of the local gradients at the same weights.
            eng.flat_grad.zero_()
            torch.manual_seed(7)
            model.train_step(dict(batch), None)['loss'].backward()
            eng._collect_grads()
            res[task]['g_local2'] = eng.flat_grad.clone()
            torch.manual_seed(7)
            eng.train_iter(batch)
            res[task]['g_mean2'] = eng.flat_grad.clone()
            res[task]['buckets'] = eng.last_buckets
    # packed device-side averaging factors of the fused det loss (one all-reduce, no .item())
    res['factors'] = model.bbox_head._avg_factors_dev([2 + rank, 0], [10, 5], 'cpu').clone()
    res['sync'] = bool(model.bbox_head.sync_cls_avg_factor)
    res['params'] = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    torch.save(res, os.path.join(out_dir, 'rank%d.pt' % rank))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_data_parallel_step(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / 'rank0.pt')
    r1 = torch.load(tmp_path / 'rank1.pt')
    # cls_avg_factor / num_total_pos: cross-rank means (pos 2 and 3 -> 2.5), clamped to >= 1 (pos 0 -> 1)
    for r in (r0, r1):
        assert torch.allclose(r['factors'][1], torch.tensor([2.5, 1.0]))
        if r['sync']:
            assert torch.allclose(r['factors'][0], torch.tensor([2.5, 1.0]))
    # replicas stay identical
    assert torch.equal(r0['params'], r1['params'])
    for task in ('cls', 'det', 'seg'):
        # the logged values are the cross-rank means (packed all-reduce), identical on both ranks
        assert r0[task]['log'].keys() == r1[task]['log'].keys()
        for k in r0[task]['log']:
            assert abs(r0[task]['log'][k] - r1[task]['log'][k]) < 1e-6, (task, k)
        assert r0[task]['ranges'] == r1[task]['ranges'] and len(r0[task]['ranges']) >= 1
        # overlapped exchange == mean of the local gradients, on both ranks, from >= 3 buckets
        want = (r0[task]['g_local2'] + r1[task]['g_local2']) / 2
        want = want * torch.clamp(0.1 / (want.norm() + 1e-6), max=1.0)        # (the CPU optimizer path clips in place)
        for r in (r0, r1):
            assert torch.allclose(r[task]['g_mean2'], want, rtol=1e-5, atol=1e-7), task
            assert r[task]['buckets'] >= 3, (task, r[task]['buckets'])
    # the cls task only touches backbone + cls_head: one contiguous range starting at 0
    assert r0['cls']['ranges'][0][0] == 0 and len(r0['cls']['ranges']) == 1
    assert len(r0['seg']['ranges']) == 2          # backbone..shared_encoder | seg_head (bbox_head skipped)


def _train_model_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    import rscotr_b200.models  # noqa: F401
    from rscotr_b200.config import MODELS
    from rscotr_b200.mtl.apis import train_model
    from rscotr_b200.mtl.data import build_datasets, load_data_cfg
    from tests.cpu_ops_shim import cpu_ops
    from tests.test_host_model import small_cfg
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = small_cfg()
    for v in cfg.data.values():
        v['config'] = os.path.join(root, v['config'])
        v['data']['samples_per_gpu'] = 2 if v['task'] == 'cls' else 1
    load_data_cfg(cfg)
    cfg.synthetic = dict(img_size=(64, 64), det=dict(num_boxes=2))
    cfg.device, cfg.compute_dtype, cfg.work_dir = 'cpu', torch.float32, os.path.join(out_dir, 'work')
    cfg.runner = dict(type='IterBasedRunner', max_iters=3)
    cfg.checkpoint_config = dict(interval=3)
    cfg.log_config = dict(interval=3)
    torch.manual_seed(0)
    model = MODELS.build(cfg.model)
    model.init_weights()
    with cpu_ops():
        runner = train_model(model, build_datasets(cfg.data, synthetic=cfg.synthetic), cfg, distributed=True, validate=False)
    torch.save(dict(params=torch.cat([p.detach().reshape(-1) for p in model.parameters()]), logs=dict(runner.log_buffer),
                    iter=runner.iter), os.path.join(out_dir, 'tm_rank%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_train_model(tmp_path):
    """mtl.apis.train_model(distributed=True) on 2 gloo ranks: different data shards, identical replicas and logged
    values afterwards, the checkpoint written once (rank 0)."""
    port = _free_port()
    mp.spawn(_train_model_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / 'tm_rank0.pt'), torch.load(tmp_path / 'tm_rank1.pt')
    assert r0['iter'] == r1['iter'] == 3 and torch.equal(r0['params'], r1['params'])
    assert r0['logs'].keys() == r1['logs'].keys() and any(k.startswith('seg.') for k in r0['logs'])
    for k in r0['logs']:
        assert abs(r0['logs'][k] - r1['logs'][k]) < 1e-6, k
    assert sorted(os.listdir(tmp_path / 'work')) == ['iter_3.pth', 'latest.pth', 'train.log.json']          # written by rank 0 only


def _eval_worker(rank, world, port, out_dir, resisc, potsdam):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    if world > 1:
        dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    import rscotr_b200.models  # noqa: F401
    from rscotr_b200.config import Config, MODELS
    from rscotr_b200.mtl.data.datasets import build_dataset
    from rscotr_b200.mtl.data.loader import build_dataloader
    from rscotr_b200.mtl.engine.test import multi_gpu_test
    from tests.cpu_ops_shim import cpu_ops
    from tests.test_host_model import small_cfg
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def shrink(pipeline, **over):
        out = []
        for t in pipeline:
            t = dict(t)
            if t['type'] in over:
                t.update(over[t['type']])
            if 'transforms' in t:
                t['transforms'] = shrink(t['transforms'], **over)
            out.append(t)
        return out
    c = dict(Config.fromfile(os.path.join(root, 'configs/datasets/resisc45.py'))._cfg_dict['data']['val'])
    c.update(data_prefix=os.path.join(resisc, 'val'), pipeline=shrink(c['pipeline'], Resize=dict(size=(64, 64))))
    s = dict(Config.fromfile(os.path.join(root, 'configs/datasets/potsdam.py'))._cfg_dict['data']['val'])
    s.update(data_root=potsdam, pipeline=shrink(s['pipeline'], MultiScaleFlipAug=dict(img_scale=(64, 64))))
    sets = dict(resisc=build_dataset(c, 'cls', dict(test_mode=True)), potsdam=build_dataset(s, 'seg', dict(test_mode=True)))
    loaders = {k: build_dataloader(v, samples_per_gpu=2, workers_per_gpu=0, dist=world > 1, shuffle=False) for k, v in sets.items()}
    torch.manual_seed(0)
    model = MODELS.build(small_cfg().model)
    model.init_weights()
    with cpu_ops():
        res = multi_gpu_test(model, loaders)
    if rank == 0:
        metrics = dict(resisc=sets['resisc'].evaluate(res['resisc'], metric='accuracy'),
                       potsdam=sets['potsdam'].evaluate(res['potsdam'], metric=['mIoU', 'mFscore']))
        torch.save(dict(n={k: len(v) for k, v in res.items()}, scores=[r.tolist() for r in res['resisc']], metrics=metrics),
                   os.path.join(out_dir, 'eval_world%d.pt' % world))
    else:
        assert all(v is None for v in res.values())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_evaluation_equals_single_process(tmp_path):
    """multi_gpu_test: shards by rank, gathers and re-interleaves the per-sample results on rank 0 -- same results and
    metrics as one process over the whole validation sets (12 images / 3 tiles: the padded duplicate is trimmed)."""
    from tests import data_fixtures as FX
    resisc, potsdam = FX.make_resisc(str(tmp_path / 'resisc')), FX.make_potsdam(str(tmp_path / 'potsdam'))
    _eval_worker(0, 1, 0, str(tmp_path), resisc, potsdam)
    mp.spawn(_eval_worker, args=(2, _free_port(), str(tmp_path), resisc, potsdam), nprocs=2, join=True)
    one, two = torch.load(tmp_path / 'eval_world1.pt', weights_only=False), torch.load(tmp_path / 'eval_world2.pt', weights_only=False)
    assert one['n'] == two['n'] == dict(resisc=12, potsdam=3)
    assert torch.allclose(torch.tensor(one['scores']), torch.tensor(two['scores']), atol=1e-6)
    assert one['metrics']['resisc'] == two['metrics']['resisc']
    for k, v in one['metrics']['potsdam'].items():
        assert v == pytest.approx(two['metrics']['potsdam'][k], nan_ok=True), k
