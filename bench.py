#!/usr/bin/env python
"""Co-training iterations/sec (Swin-T, synthetic 3x800x800, cls 16 / det 1 / seg 2 per GPU,
round robin) -- BASELINE.json's metric on configs[1] (N=1) / configs[2] (N=8).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One *step* = one co-training iteration: H2D (e2e leg only) -> shared Swin backbone ->
neck -> shared deformable encoder -> one task head -> losses -> backward -> flat NCCL
all-reduce of the gradients -> global-norm clip -> AdamW.  Steps cycle cls, det, seg.
Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle (the stand-in
for the reference's own CPU path, which cannot be installed here: mmcv-full 1.6.1 /
mmdet / mmcls / mmseg are absent and there is no network) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
CONFIG = os.path.join(ROOT, 'configs', 'multi', 'cotrain_swin-t_800.py')
METRIC = 'co-training iters/sec (Swin-T, 3x800x800)'
TASK_ORDER = ('resisc', 'dior', 'potsdam')
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
# (profiles/r02_ncu_wmsa_tma_stage0_B16.txt: stage-0 launch, B=16; the bench's average launch is smaller)
NCU_TRAFFIC = {   # per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the profiled shape
    'rsc_wmsa_bwd': dict(stage0_B16_bytes=820.2e6, stage0_B16_alg_bytes=860.2e6,
                         source='profiles/r02_ncu_wmsa_tma_stage0_B16.txt'),
    'rsc_wmsa_fwd': dict(stage0_B16_bytes=468.6e6, stage0_B16_alg_bytes=491.5e6,
                         source='profiles/r02_ncu_wmsa_tma_stage0_B16.txt'),
    'rsc_msda_fused_bwd': dict(encoder_B2_bytes=129.9e6, note='L2-resident gather / atomics: DRAM traffic is 4 % of '
                               'peak (ncu capture of the un-fused kernel at the same shape, gpurun prof_msda)'),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=6)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default=CONFIG)
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-budget-s', type=float, default=300.0)
    ap.add_argument('--sustained-s', type=float, default=5.0,
                    help='length of the extra sustained leg (>= 30 round-robin cycles and >= this many seconds; 0 = skip)')
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """samples nvidia-smi during the timed region (B200_PROFILING.md 'clocks line')."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])), mx.append(float(r[1])), pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(sm), reasons=sorted(reasons))


# ----------------------------------------------------------------------------- helpers
def build(cfg_path, dtype, device):
    import torch
    import rscotr_b200.models  # noqa: F401
    from rscotr_b200.config import Config, MODELS
    from rscotr_b200.mtl.data import build_datasets, build_multidataloader, load_data_cfg
    from rscotr_b200.mtl.engine import StepEngine
    cfg = Config.fromfile(cfg_path)
    load_data_cfg(cfg, config_root=ROOT)
    torch.manual_seed(0)
    model = MODELS.build(cfg.model)
    model.init_weights()
    model.train()
    engine = StepEngine(model, cfg.optimizer, grad_clip=cfg.optimizer_config.get('grad_clip'), device=device,
                        compute_dtype=torch.bfloat16 if dtype == 'bf16' else torch.float32, lr_config=cfg.lr_config)
    datasets = build_datasets(cfg.data, synthetic=cfg.get('synthetic'))
    loader = build_multidataloader(cfg, int(os.environ.get('WORLD_SIZE', 1)) > 1, datasets)
    return cfg, model, engine, loader


def peaks():
    """(HBM GB/s, dense bf16 TFLOP/s sustained, source): the kernels are timed inside a long step -> sustained figure"""
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), float(d.get('bf16_tflops_sustained', d.get('bf16_tflops', 1400.0))), \
            'measured (MEASURED_PEAKS.json hbm_gbs / bf16_tflops_sustained)'
    return 6650.0, 1400.0, 'fallback (B200_PROFILING.md 6.65 TB/s, ~1.4 PFLOP/s sustained)'


# ----------------------------------------------------------------------------- CPU oracle arm
def oracle_step_fn(cfg, img_hw):
    """CPU fp32 oracle train step (fwd + bwd + clip + AdamW) on the REAL per-GPU batch of each task.

    step(name, images=None) runs one co-training iteration of task `name`: the batch is walked one image at a time with
    the gradients accumulated (the losses are batch means, so this is the same step; eager CPU code gains nothing from
    batching and a 16-image Swin autograd graph at 800^2 does not fit host memory), then ONE clip + AdamW update.
    `images` < batch size bounds the sample: the step then processes that many images and reports how many it did."""
    import torch
    import rscotr_b200.models  # noqa: F401
    from oracle import heads as oh
    from rscotr_b200.config import MODELS
    from rscotr_b200.mtl.data.synthetic import SyntheticDataset
    from rscotr_b200.mtl.utils.optimizer import param_settings
    torch.manual_seed(0)
    model = MODELS.build(cfg.model)          # only as a container of identically initialised weights
    model.init_weights()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    params = {n: sd[n].requires_grad_(True) for n, _ in model.named_parameters()}
    opt_cfg = dict(cfg.optimizer)
    pw = opt_cfg.pop('paramwise_cfg', None)
    opt_cfg.pop('type')
    groups = {}
    for n, _, lr, wd in param_settings(model, opt_cfg, pw):
        groups.setdefault((lr, wd), []).append(params[n])
    optim = torch.optim.AdamW([dict(params=ps, lr=lr, weight_decay=wd) for (lr, wd), ps in groups.items()])
    del model
    g = torch.Generator().manual_seed(3)
    bs = per_gpu_batch(cfg)
    batches = {}
    for name in TASK_ORDER:
        task = cfg.data[name]['task']
        ds = SyntheticDataset(task, img_size=img_hw, **dict(cfg.get('synthetic', {}).get(task, {})))
        micro = []
        for _ in range(bs[name]):
            b = ds.make_batch(1, g, pin=False)
            b.update(task=task, dataset_name=name)
            micro.append(b)
        batches[name] = micro
    tw = dict(cls=1, det=1, seg=1)
    tw.update(cfg.model.get('task_weight') or {})
    max_norm = cfg.optimizer_config.get('grad_clip', {}).get('max_norm', 0.1)

    def step(name, images=None):
        micro = batches[name]
        n = len(micro) if images is None else max(1, min(len(micro), images))
        optim.zero_grad(set_to_none=True)
        total = 0.0
        for b in micro[:n]:
            b = dict(b)
            task = b['task']
            noise = oh.cdn_noise(b['gt_labels'], generator=g) if task == 'det' else None
            losses = oh.mtl_losses(sd, task, b, noise=noise)
            loss, _ = oh.parse_losses(losses, tw[task])
            (loss / n).backward()
            total += float(loss.detach()) / n
        torch.nn.utils.clip_grad_norm_([p for p in params.values() if p.grad is not None], max_norm)
        optim.step()
        return total, n

    return step


def per_gpu_batch(cfg):
    out = {}
    for n, d in cfg.data.items():
        sub = d.get('config')
        bs = (d.get('data') or {}).get('samples_per_gpu')
        if bs is None and hasattr(sub, 'get'):
            bs = (sub.get('data') or {}).get('samples_per_gpu')
        out[n] = int(bs or 1)
    return out


def _nvtx_range(name):
    """push an NVTX range, return the function that pops it (a no-op pair if NVTX is unavailable)."""
    try:
        import torch
        torch.cuda.nvtx.range_push(name)
        return torch.cuda.nvtx.range_pop
    except Exception:
        return lambda: None


def _time_oracle_steps(step, names, images_per_step):
    """-> {task name: [seconds per FULL step]}: each call processes images_per_step[name] images (the whole per-GPU
    batch unless the sample had to be bounded) and the time is scaled to the full batch."""
    bs_done = {}
    times = {}
    for name in names:
        t0 = time.time()
        _, n = step(name, images_per_step.get(name))
        dt = time.time() - t0
        bs_done[name] = n
        times.setdefault(name, []).append(dt)
    return times, bs_done


def run_reference(args):
    """Reference arm: the CPU oracle on all host cores, at the arm's own config: 3x800x800, per-GPU batches 16 / 1 / 2,
    round robin.  Every timed step is a REAL co-training iteration on the full batch (micro-batches of one image,
    gradients accumulated, one clip + AdamW update).  Only when --steps / --warmup would exceed --cpu-budget-s is the
    cls batch sampled (fewer of its 16 images per step, time scaled back up; the JSON then says `extrapolated`)."""
    if int(os.environ.get('RANK', 0)) != 0:
        return
    import torch
    from rscotr_b200.config import Config
    from rscotr_b200.mtl.data import load_data_cfg
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    cfg = Config.fromfile(args.config)
    load_data_cfg(cfg, config_root=ROOT)
    hw = tuple(cfg.synthetic['img_size'])
    bs = per_gpu_batch(cfg)
    step = oracle_step_fn(cfg, hw)
    step(TASK_ORDER[0], 1)                   # allocator / thread-pool warm-up
    t0 = time.time()
    step(TASK_ORDER[0], 1)                   # one cls image: probe of the per-image cost
    probe = time.time() - t0
    n_total = args.steps + args.warmup
    # projected cost of the full run: a cycle = 16 cls images + 1 det image (~2.5x a cls image) + 2 seg images (~2x)
    cycle = probe * (bs[TASK_ORDER[0]] + 2.5 * bs[TASK_ORDER[1]] + 2.0 * bs[TASK_ORDER[2]])
    images = {}
    if cycle * n_total / 3.0 > args.cpu_budget_s:
        spare = args.cpu_budget_s * 3.0 / n_total - probe * (2.5 * bs[TASK_ORDER[1]] + 2.0 * bs[TASK_ORDER[2]])
        images[TASK_ORDER[0]] = int(max(1, min(bs[TASK_ORDER[0]], spare / probe)))
    times = {n: [] for n in TASK_ORDER}
    done = {}
    for i in range(n_total):
        name = TASK_ORDER[i % 3]
        t0 = time.time()
        _, n = step(name, images.get(name))
        dt = (time.time() - t0) * bs[name] / n
        done[name] = n
        if i >= args.warmup:
            times[name].append(dt)
    for name in TASK_ORDER:                  # fewer than 3 timed steps: every task still needs one sample for the cycle time
        if not times[name]:
            t0 = time.time()
            _, n = step(name, images.get(name))
            times[name].append((time.time() - t0) * bs[name] / n)
            done[name] = n
    per_step = {n: sum(v) / len(v) for n, v in times.items()}
    # mean over the timed steps in the order they ran (cls, det, seg, cls, ...)
    seq = [per_step[TASK_ORDER[(args.warmup + i) % 3]] for i in range(args.steps)] if args.steps else list(per_step.values())
    value = len(seq) / sum(seq)
    extrap = any(done[n] < bs[n] for n in TASK_ORDER)
    sample = ('oracle (CPU fp32 restatement) co-training steps at 3x%dx%d on the per-GPU batches %s, micro-batches of one '
              'image with accumulated gradients + one clip / AdamW update per step; images processed per step %s%s; '
              'seconds per full step %s') % (
        hw[0], hw[1], bs, done, ' (cls batch SAMPLED and scaled: extrapolated)' if extrap else ' (full batches: measured)',
        {k: round(v, 2) for k, v in per_step.items()})
    out = dict(metric=METRIC, value=value, unit='iters/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
               ms_per_step=1000.0 / value, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
               data='synthetic', impl='reference', extrapolated=extrap,
               config=dict(workload='Swin-T MTL co-training (cls+seg+det round-robin) 3x800x800, per-GPU batch 16/1/2',
                           per_gpu_batch=bs,
                           note='reference itself is not installable offline (mmcv-full/mmdet/mmcls/mmseg absent); '
                                'this arm runs the CPU oracle port'),
               cpu_baseline=dict(value=value, unit='iters/s', cores=cores, kind='port', sample=sample),
               e2e=dict(value=value, unit='iters/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out))


def cpu_baseline(cfg, budget_s=40.0):
    """Bounded CPU sample inside the default run, SAME method as the reference arm: one real co-training cycle
    (cls step on its 16 images, det step, seg step) of the oracle at 3x800x800 on all host cores; the cls batch is
    sampled (4 of its 16 images, time scaled) to keep the sample within ~20-30 s."""
    import torch
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    hw = tuple(cfg.synthetic['img_size'])
    bs = per_gpu_batch(cfg)
    step = oracle_step_fn(cfg, hw)
    step(TASK_ORDER[0], 1)                   # warm-up: one cls image
    images = {TASK_ORDER[0]: min(4, bs[TASK_ORDER[0]])}
    per_step, done = {}, {}
    for name in TASK_ORDER:
        t0 = time.time()
        _, n = step(name, images.get(name))
        per_step[name] = (time.time() - t0) * bs[name] / n
        done[name] = n
    return dict(value=3.0 / sum(per_step.values()), unit='iters/s', cores=cores, kind='port',
                sample=('one oracle (CPU fp32) co-training cycle at 3x%dx%d on the per-GPU batches %s (micro-batches of one '
                        'image, accumulated gradients, one clip + AdamW per step); images processed per step %s, cls '
                        'time scaled to its 16 images; seconds per full step %s') % (
                    hw[0], hw[1], bs, done, {k: round(v, 2) for k, v in per_step.items()}))


# ----------------------------------------------------------------------------- main arm
def main():
    args = parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    import torch
    import torch.distributed as dist
    from rscotr_b200 import ops
    from rscotr_b200.mtl.engine.step import _to_device, h2d_bytes

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    # stdout carries exactly ONE line (the JSON): libraries that print there (NCCL's version banner) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    assert world == args.gpus, 'launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)' % (args.gpus, world)

    T0 = time.time()
    cfg, model, engine, loader = build(args.config, args.dtype, device)
    it = iter(loader)
    names = list(cfg.data.keys())                                # round-robin order of the datasets (one task each)
    NB = 2 * len(names)
    host_batches = [next(it) for _ in range(NB)]                 # 2 distinct batches per dataset, pinned host memory
    dev_batches = [_to_device(b, device) for b in host_batches]
    torch.cuda.synchronize()

    def trace(msg):
        if os.environ.get('RSC_BENCH_TRACE'):
            print('[bench rank %d %.1fs] %s' % (rank, time.time() - T0, msg), file=sys.stderr, flush=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up (also fixes the per-task all-reduce ranges, cuBLAS heuristics, allocator)
    # (with CUDA graphs every task needs 2 eager iterations + the capturing one before the timed region)
    n_warm = max(args.warmup, 3 * len(names) if engine.use_graphs else 3)
    for i in range(n_warm):
        trace('warm-up step %d (%s)' % (i, dev_batches[i % NB]['task']))
        engine.train_iter(dev_batches[i % NB])
        if os.environ.get('RSC_BENCH_TRACE'):
            torch.cuda.synchronize()
    # the end-to-end path (timed region B) has its own one-off costs: the first prefetch of a (task, shapes) signature
    # allocates its device staging buffers -- one untimed step per distinct host batch
    for i in range(NB):
        engine.prefetch(host_batches[i])
        engine.train_iter(host_batches[i])
    trace('warm-up done')
    barrier()

    loss_host = torch.empty(args.steps, dtype=torch.float32).pin_memory()       # read-back buffer of region B (page-locking synchronises)

    # ---- timed region A: inputs resident in HBM
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.reset_launch_count()
    engine.replayed_launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per_task = {}
    barrier()
    nvtx = _nvtx_range('timed_region')               # lets `ncu --nvtx --nvtx-include "timed_region/"` profile only these steps
    e0.record()
    evs = []
    for i in range(args.steps):
        a = torch.cuda.Event(enable_timing=True)
        a.record()
        out = engine.train_iter(dev_batches[i % NB])
        b = torch.cuda.Event(enable_timing=True)
        b.record()
        evs.append((dev_batches[i % NB]['task'], a, b))
    e1.record()
    nvtx()
    trace('timed region A enqueued')
    barrier()
    trace('timed region A done')
    ms = e0.elapsed_time(e1)
    launches = ops.launch_count() + engine.replayed_launches      # eager launches + launches inside graph replays
    for t, a, b in evs:
        per_task.setdefault(t, []).append(a.elapsed_time(b))
    final_loss = float(out['loss'].detach())
    # ---- timed region B: end to end through the public API, pinned host inputs, loss read back
    # It follows region A directly (tools/e2e_probe.py: for ~50 steps after the eager, host-bound per-kernel pass further
    # down BOTH loops run 3 % slower); its own path was warmed up before region A.
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    hb = db = 0
    # The loss of every step is read back into pinned host memory (4 bytes, async D2H + event) and consumed LAG steps
    # later, after the next step has been enqueued -- the logging lag of any real training loop -- so the host never
    # leaves the GPU idle between steps; the next batch's H2D copy runs on the copy stream meanwhile.
    read_ev = [None] * args.steps
    losses_read = []
    LAG = 2      # steps between enqueueing a step and reading its loss: with 1 the host is at most one step ahead of the GPU and
                 # every host-side hiccup (the nvidia-smi clock sampler takes driver locks every 100 ms) idles the device

    def consume(k):
        read_ev[k].synchronize()
        losses_read.append(float(loss_host[k]))

    engine.prefetch(host_batches[0])
    for i in range(args.steps):
        batch = host_batches[i % NB]
        hb += h2d_bytes(batch)
        o = engine.train_iter(batch)
        loss_host[i:i + 1].copy_(o['loss'].detach().reshape(1).float(), non_blocking=True)   # D2H of the step's result
        read_ev[i] = torch.cuda.Event()
        read_ev[i].record()
        db += 4
        if i + 1 < args.steps:
            engine.prefetch(host_batches[(i + 1) % NB])                    # next step's H2D overlaps this step's compute
        if i >= LAG:
            consume(i - LAG)
    for k in range(max(args.steps - LAG, 0), args.steps):
        consume(k)
    assert len(losses_read) == args.steps and all(v == v for v in losses_read), 'e2e: missing / NaN loss read-back'
    f1.record()
    barrier()
    trace('timed region B done')
    ms_e2e = f0.elapsed_time(f1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- sustained leg: >= 30 full round-robin cycles and >= --sustained-s seconds, with its own clock record (the
    # timed region above is what --steps asks for; at 20-30 steps it lasts 0.3-0.4 s, i.e. it is a burst number)
    sustained = None
    if args.sustained_s > 0:
        nn_ = len(names)
        n_sus = max(30 * nn_, int(1.1 * args.sustained_s / max(ms / args.steps / 1000.0, 1e-4)) // nn_ * nn_ + nn_)
        s_sampler = ClockSampler(local)
        if rank == 0:
            s_sampler.start()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(n_sus):
            engine.train_iter(dev_batches[i % NB])
        s1.record()
        barrier()
        sus_ms = s0.elapsed_time(s1)
        if world > 1:
            t = torch.tensor([sus_ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sus_ms = float(t[0])
        sustained = dict(value=world * n_sus / (sus_ms / 1000.0), unit='iters/s', steps=n_sus, cycles=n_sus // len(names),
                         seconds=sus_ms / 1000.0, ms_per_step=sus_ms / n_sus,
                         clocks=s_sampler.stop() if rank == 0 else None)
        trace('sustained leg done')
    # per-kernel CUDA-event timing of the same steps (two cycles).  The step engine runs eagerly while a
    # KernelTimer is active: events cannot be recorded per kernel inside a CUDA-graph replay.
    with ops.KernelTimer() as kt:
        for i in range(NB):
            engine.train_iter(dev_batches[i % NB])
        ksum = kt.summary(*peaks()[:2])
    trace('kernel-timer pass done')
    kscale = args.steps / float(NB)                                       # normalise kernel ms to the timed region's steps
    for d in ksum.values():
        for k in ('ms', 'bytes', 'big_ms', 'big_bytes', 'big_flops', 'big_roof_ms'):
            d[k] *= kscale
        d['launches'] = int(round(d['launches'] * kscale))
        d['big_launches'] = int(round(d['big_launches'] * kscale))

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    def finish():
        """Leave together: tearing NCCL down while captured collectives are alive (or while a peer has already
        gone) can block or abort, so drop the graphs, meet at a barrier AFTER rank 0 has printed, give
        destroy_process_group a bounded chance and exit."""
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            import gc
            engine._graphs.clear()
            gc.collect()
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            th = threading.Thread(target=dist.destroy_process_group, daemon=True)
            th.start()
            th.join(15.0)
            os._exit(0)

    if rank != 0:
        return finish()

    value = world * args.steps / (ms / 1000.0)                   # whole-job iterations (one per rank per step)
    e2e = world * args.steps / (ms_e2e / 1000.0)
    peak, peak_tf, peak_src = peaks()
    mine_ms = sum(d['ms'] for d in ksum.values())

    def roof(name, d):
        """roofline object of one entry point over its launches with >= 32 MB of algorithmic bytes: the binding limit is
        the slower of (bytes / HBM peak) and (flops / tensor peak) -- the GEMMs switch sides with the shape"""
        if not d['big_launches']:
            return None
        t_s = d['big_ms'] / 1000.0
        t_hbm, t_tc = d['big_bytes'] / (peak * 1e9), d['big_flops'] / (peak_tf * 1e12)
        tensor = t_tc > t_hbm
        achieved = d['big_flops'] / t_s / 1e12 if tensor else d['big_bytes'] / t_s / 1e9
        r = dict(kernel=name, bound='tensor' if tensor else 'hbm', achieved=achieved, peak=peak_tf if tensor else peak,
                 unit='TFLOP/s' if tensor else 'GB/s', frac=achieved / (peak_tf if tensor else peak),
                 traffic=NCU_TRAFFIC.get(name), launches=d['big_launches'], avg_us=1000.0 * d['big_ms'] / d['big_launches'],
                 alg_bytes_per_launch=d['big_bytes'] / d['big_launches'], share_of_step=d['big_ms'] / ms)
        if d['big_flops']:
            # the GEMMs straddle the ridge point (HBM-bound at the 96 / 192-wide Swin stages, tensor-bound at 384 / 768
            # and in the encoder FFN): frac_per_launch_bound = sum over launches of the time AT that launch's own
            # binding limit, max(bytes / HBM peak, flops / tensor peak), over the measured time
            r.update(alg_flops_per_launch=d['big_flops'] / d['big_launches'], hbm_gbs=d['big_bytes'] / t_s / 1e9,
                     tflops=d['big_flops'] / t_s / 1e12, frac_per_launch_bound=d.get('big_roof_ms', 0.0) / d['big_ms'])
        return r

    # dominant kernel = most device time over launches that move >= 32 MB (for the many tiny launches of the
    # decoders the question is launch latency, not bandwidth); its roofline numbers are over those launches
    top = max(ksum.items(), key=lambda kv: kv[1]['big_ms']) if ksum else (None, None)
    roofline = roof(*top) if top[0] else None
    if roofline:
        roofline.update(scope='launches with >= 32 MB of algorithmic bytes',
                        timed='CUDA events around every launch during an eager pass of the same steps (the timed '
                              'region itself replays CUDA graphs)',
                        peak_source=peak_src, own_kernels_share_of_step=mine_ms / ms)
    # the window-attention kernels the previous round's review named, and the other Linear kernels, beside it
    roofline_also = [r for r in (roof(k, ksum[k]) for k in ('rsc_wmsa_bwd', 'rsc_wmsa_fwd', 'rsc_linear_fwd', 'rsc_linear_dx',
                                                            'rsc_linear_dw') if k in ksum and (not top[0] or k != top[0])) if r]
    bs = per_gpu_batch(cfg)
    out = dict(metric=METRIC, value=value, unit='iters/s', n_gpus=world, steps=args.steps, warmup=n_warm,
               ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
               dtype=args.dtype if args.dtype != 'fp32' else 'f32', data='synthetic',
               config=dict(workload=('Swin-T MTL co-training (cls+seg+det round-robin) 3x800x800 (BASELINE configs[%d])'
                                     % (1 if world == 1 else 2)) if os.path.abspath(args.config) == os.path.abspath(CONFIG) else
                           '%s: %s, synthetic %s' % (os.path.relpath(args.config, ROOT), cfg.model.get('type'),
                                                     'x'.join(str(v) for v in dev_batches[0]['img'].shape[1:])),
                           per_gpu_batch=bs, global_batch={k: v * world for k, v in bs.items()},
                           parallelism='dp%d' % world, strategy='round_robin', config_file=os.path.relpath(args.config, ROOT),
                           l2='no flush: each step streams >1 GB of activations (inputs alone 123 MB fp32 for the cls '
                              'batch), far above the 126 MB L2',
                           weights='random init, %.1f M params' % (sum(p.numel() for p in model.parameters()) / 1e6),
                           final_loss=final_loss),
               clocks=clocks,
               e2e=dict(value=e2e, unit='iters/s', h2d_bytes_per_step=hb // args.steps, d2h_bytes_per_step=db // args.steps,
                        ms_per_step=ms_e2e / args.steps,
                        pipeline='pinned host inputs, H2D of step i+1 on a copy stream during step i; loss of step i '
                                 'copied to pinned host memory and read after step i+2 is enqueued'),
               gpu_launches=launches, cuda_graphs=bool(engine.use_graphs), cuda_graph_capture_failures=engine.graph_failures,
               # which of the A/B'd variants this run used (defaults unless the environment says otherwise; DESIGN.md sections 4 / 9)
               switches=dict(patch_merge_kernels=2 if os.environ.get('RSC_PATCH_MERGE_V2') == '1' else 1,
                             linear_pair=os.environ.get('RSC_LINEAR_PAIR', '1') != '0',
                             msda_bwd_branch_free=os.environ.get('RSC_MSDA_BWD_BF', '1') != '0',
                             own_gemm=os.environ.get('RSC_OWN_GEMM', '1') != '0'),
               ms_per_task={k: sum(v) / len(v) for k, v in per_task.items()},
               roofline=roofline, roofline_also=roofline_also, sustained=sustained,
               kernels={k: dict(launches=d['launches'], ms=round(d['ms'], 3), gbs=round(d['gbs'], 1),
                                big_launches=d['big_launches'], big_ms=round(d['big_ms'], 3),
                                big_gbs=round(d['big_gbs'], 1), big_frac=round(d['big_gbs'] / peak, 4))
                        for k, d in sorted(ksum.items())})
    if world == 1 and not args.no_cpu_baseline and os.path.abspath(args.config) == os.path.abspath(CONFIG):
        del engine, model, dev_batches
        torch.cuda.empty_cache()
        out['cpu_baseline'] = cpu_baseline(cfg)
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(out) + '\n').encode())
    finish()


if __name__ == '__main__':
    main()
