"""Co-training step engine: one task per iteration through the shared backbone and one
head (reference hot loop: mmcv IterBasedRunner.train + OptimizerHook, SURVEY 3.1 / D.6;
rows a21-a23 and 8e).

B200-first structure:
  * every parameter's .grad is a VIEW into one flat fp32 buffer ordered
    [backbone | cls_head | neck | shared_encoder | bbox_head | seg_head]: zero_grad is
    one memset (zero-FILL, which is also what the reference's torch-1.11 zero_grad did,
    SURVEY 3.4), the global-norm clip is one norm + one scale, and the data-parallel
    exchange is a flat NCCL all-reduce over the contiguous range(s) the current task
    touches -- gradients only, nothing in forward except the packed loss factors.
  * bf16 autocast for the GEMMs / kernels, fp32 master weights, fused AdamW.
  * inputs arrive in pinned host memory and are copied with non_blocking H2D.
"""
import bisect
import math
import os
import re

import torch
import torch.distributed as dist

from ... import _lib
from ..utils.optimizer import build_optimizer

_ORDER = ('backbone', 'cls_head', 'neck', 'shared_encoder', 'bbox_head', 'seg_head')
_STAGE_RE = re.compile(r'backbone\.(?:stages\.(\d+)\.|norm(\d+)\.)')


def _to_device(obj, device):
    if torch.is_tensor(obj):
        return obj.to(device, non_blocking=True)
    if isinstance(obj, list):
        return [_to_device(o, device) for o in obj]
    if isinstance(obj, tuple):
        return tuple(_to_device(o, device) for o in obj)
    if isinstance(obj, dict):
        return {k: _to_device(v, device) for k, v in obj.items()}
    return obj


def _static_copy(obj, device):
    """device copy with private storage (never aliases the caller's tensors)."""
    if torch.is_tensor(obj):
        return obj.to(device, non_blocking=True) if not obj.is_cuda else obj.clone()
    if isinstance(obj, list):
        return [_static_copy(o, device) for o in obj]
    if isinstance(obj, tuple):
        return tuple(_static_copy(o, device) for o in obj)
    if isinstance(obj, dict):
        return {k: _static_copy(v, device) for k, v in obj.items()}
    return obj


def h2d_bytes(obj):
    if torch.is_tensor(obj):
        return 0 if obj.is_cuda else obj.numel() * obj.element_size()
    if isinstance(obj, (list, tuple)):
        return sum(h2d_bytes(o) for o in obj)
    if isinstance(obj, dict):
        return sum(h2d_bytes(v) for v in obj.values())
    return 0


_SKIP_EXCHANGE = os.environ.get('RSC_SKIP_EXCHANGE', '0') == '1'


class _NoWork:
    def wait(self):
        pass


class _StreamWork:
    """work handle of an exchange that runs on the engine's own communication stream"""
    def __init__(self, stream, device):
        self.stream, self.device = stream, device

    def wait(self):
        torch.cuda.current_stream(self.device).wait_stream(self.stream)


class _DivAfter:
    """async SUM all-reduce whose wait() finishes the mean (back ends without an averaging reduction)."""
    def __init__(self, work, seg, world):
        self.work, self.seg, self.world = work, seg, world

    def wait(self):
        self.work.wait()
        self.seg.div_(self.world)


class StepEngine:
    def __init__(self, model, optimizer_cfg, grad_clip=None, device='cuda', compute_dtype=torch.bfloat16,
                 lr_config=None, use_graphs=None):
        self.device = torch.device(device)
        self.model = model.to(self.device)
        self.compute_dtype = compute_dtype
        self.grad_clip = dict(grad_clip) if grad_clip else None
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self._build_flat_grads()
        if use_graphs is None:
            use_graphs = os.environ.get('RSC_CUDA_GRAPHS', '1') != '0'
        self.use_graphs = bool(use_graphs) and self.device.type == 'cuda'
        optimizer_cfg = dict(optimizer_cfg)
        if optimizer_cfg.get('type') == 'AdamW' and self.device.type == 'cuda' and not optimizer_cfg.get('amsgrad'):
            from .flat_adamw import FlatAdamW
            self.optimizer = FlatAdamW(self, optimizer_cfg)      # one fused kernel per (lr, wd) run of the flat buffer
        else:
            if self.use_graphs and optimizer_cfg.get('type') in ('Adam', 'AdamW'):
                optimizer_cfg['capturable'] = True
            self.optimizer = build_optimizer(self.model, optimizer_cfg)
        self._base_lrs = [float(g['lr']) for g in self.optimizer.param_groups]
        self._graphs = {}
        self.graph_warmup = 2            # eager iterations per (task, shapes) before capture
        self.max_graphs = int(os.environ.get('RSC_MAX_GRAPHS', 12))
        self.graph_idle_iters = 512
        self.replayed_launches = 0       # rscotr kernels executed through graph replays
        self.graph_failures = 0
        self.lr_config = dict(lr_config) if lr_config else None
        self.max_iters = (self.lr_config or {}).get('max_iters', 1)     # (poly / cosine horizons; train_model sets it)
        self.iter = 0
        self._task_ranges = {}
        self.last_grad_norm = None
        self._copy_stream = torch.cuda.Stream(self.device) if self.device.type == 'cuda' else None

    # -- flat parameter / gradient buffers ---------------------------------
    def _build_flat_grads(self):
        """Parameters become views into ONE flat fp32 buffer (and their gradients land in a second
        one with the same layout), ordered by _ORDER; every tensor starts on a 64-element boundary."""
        named = list(self.model.named_parameters())
        def rank(n):
            top = n.split('.')[0]
            return _ORDER.index(top) if top in _ORDER else len(_ORDER)

        def sub(n):
            # backbone parameters in the order their gradients complete, last first: stage K's output norm sits
            # with stage K, so "everything from stage K on" is ONE contiguous tail of the backbone range (the
            # buckets of the overlapped exchange, _on_grad_ready)
            m = _STAGE_RE.match(n)
            return int(m.group(1) or m.group(2)) + 1 if m else 0
        order = [i for i in sorted(range(len(named)), key=lambda i: (rank(named[i][0]), sub(named[i][0]), i))
                 if named[i][1].requires_grad]
        order = self._pair_linears(named, order)
        self._spans = []          # (name, start, end)
        off = 0
        for i in order:
            n, p = named[i]
            assert p.dtype == torch.float32, 'master weights are fp32'
            self._spans.append((n, off, off + p.numel()))
            off = (off + p.numel() + 63) // 64 * 64      # 128-byte aligned bf16 shadow views (TMA-friendly GEMM operands)
        total = off
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=self.device)
        # In-switch gradient exchange (csrc/nvls.cu): the gradient buffer lives in symmetric memory with an NVLS
        # multicast mapping, every rank reduces 1/world of a range with multimem.ld_reduce / multimem.st.
        # RSC_NVLS=1 opts in; without multicast support (or on any set-up error, on every rank alike) NCCL is used.
        self._nvls = None
        if self.world > 1 and self.device.type == 'cuda' and os.environ.get('RSC_NVLS', '0') == '1':
            try:
                import torch.distributed._symmetric_memory as symm
                fg = symm.empty(total, dtype=torch.float32, device=self.device)
                hdl = symm.rendezvous(fg, dist.group.WORLD)
                if not hdl.multicast_ptr:
                    raise RuntimeError('no NVLS multicast mapping')
                fg.zero_()
                self.flat_grad = fg
                self._nvls = dict(hdl=hdl, mc=int(hdl.multicast_ptr), stream=torch.cuda.Stream(self.device),
                                  ctas=int(os.environ.get('RSC_NVLS_CTAS', 16)), rank=dist.get_rank())
            except Exception as e:      # noqa: BLE001  (symmetric memory is an experimental torch API)
                import warnings
                warnings.warn('in-switch gradient exchange unavailable (%s: %s); using NCCL' % (type(e).__name__, e))
        # bf16 shadow of the master weights, refreshed by the AdamW kernel itself: the GEMMs read it
        # directly (ops.linear), so no fp32->bf16 weight cast kernel runs in the step
        lp = self.device.type == 'cuda' and self.compute_dtype == torch.bfloat16
        self.flat_param_lp = torch.zeros(total, dtype=torch.bfloat16, device=self.device) if lp else None
        self._params, self._grad_views = [], []
        for (n, s0, e0), i in zip(self._spans, order):
            p = named[i][1]
            self.flat_param[s0:e0].copy_(p.data.reshape(-1))
            p.data = self.flat_param[s0:e0].view_as(p)
            p.grad = None
            self._params.append(p)
            self._grad_views.append(self.flat_grad[s0:e0].view_as(p))
            # ops.linear & co. accumulate this parameter's gradient straight into the flat buffer (fp32
            # GEMM output, beta = 1) instead of bf16 dW -> cast -> AccumulateGrad -> copy
            p._rsc_g = self._grad_views[-1]
            p._rsc_lp = self.flat_param_lp[s0:e0].view_as(p) if lp else None
        self._grad_of = {id(p): v for p, v in zip(self._params, self._grad_views)}
        self._span_starts = [s0 for _, s0, _ in self._spans]
        # gradient-completion frontiers of the overlapped exchange: key -> flat offset from which every gradient is
        # final once the backbone reports `key` during backward ('outs' = the whole backbone is still pending)
        self._frontier_of = {}
        tops = [n.split('.')[0] for n, _, _ in self._spans]
        if 'backbone' in tops:
            self._frontier_of['outs'] = max(e for (n, s, e), t in zip(self._spans, tops) if t == 'backbone')
            for n, s0, _ in self._spans:
                m = _STAGE_RE.match(n)
                if m and m.group(1) is not None:
                    self._frontier_of.setdefault(int(m.group(1)), s0)
        self._min_bucket = int(os.environ.get('RSC_MIN_BUCKET', 1 << 20))     # elements; smaller pieces ride with the next
        self._pending = None
        bb = getattr(self.model, 'backbone', None)
        if self.world > 1 and bb is not None and os.environ.get('RSC_OVERLAP_EXCHANGE', '1') != '0':
            bb._grad_ready_cb = self._on_grad_ready
        # data parallel, opt-in (RSC_DP_GEMM_SMS=120): size the persistent GEMMs of the small-batch (det / seg) steps for fewer SMs
        # so that the exchange kernels running next to them do not push their last CTAs into a second wave
        # (csrc/gemm_tc.cu::gemm_sms).  A/B'd on 2 x B200 in round 2: within the box-to-box spread (det 10.2 -> 9.7 ms on one box,
        # 9.74 -> 9.77 ms on the next; the cls step loses 0.4-0.7 ms when it is included) -> off by default.
        self._dp_gemm_sms = int(os.environ.get('RSC_DP_GEMM_SMS', 0)) if self.world > 1 and self.device.type == 'cuda' else 0
        self._dp_gemm_max_batch = int(os.environ.get('RSC_DP_GEMM_SMS_MAX_BATCH', 4))
        if self.world > 1 and self.device.type == 'cuda' and os.environ.get('RSC_ASYNC_LOG_REDUCE', '1') != '0':
            from ...models import mtl as _mtl
            _mtl.ASYNC_LOG_WORKS = []
        if self.world > 1:
            # what the reference's DDP wrapper does at construction (mtl/apis/train.py:37-46): rank 0's weights
            # (and buffers) win, so per-rank init differences (--diff-seed, nondeterministic init) cannot leave the
            # replicas silently different -- only gradients are exchanged afterwards
            dist.broadcast(self.flat_param, src=0)
            for b in self.model.buffers():
                dist.broadcast(b.data, src=0)
        self.sync_lp()

    def _pair_linears(self, named, order):
        """Sibling Linear layers that read the same input (a module lists them in `_rsc_linear_pairs`, e.g.
        MultiScaleDeformableAttention: sampling_offsets / attention_weights) are laid out as
        [weight1 | weight2 | bias1 | bias2], so that the stacked matrices are plain views of the flat buffers and ONE
        GEMM serves both layers (ops.linear_pair).  Needs sizes that keep the 64-element alignment gap-free; anything
        else is left as it was (the module then runs two GEMMs)."""
        pos = {named[i][0]: k for k, i in enumerate(order)}
        index = {n: i for i, (n, _) in enumerate(named)}
        order = list(order)
        for mod_name, mod in self.model.named_modules():
            for a, b in getattr(mod, '_rsc_linear_pairs', ()):
                pre = mod_name + '.' if mod_name else ''
                names = [pre + a + '.weight', pre + b + '.weight', pre + a + '.bias', pre + b + '.bias']
                if not all(n in pos for n in names):
                    continue
                ks = sorted(pos[n] for n in names)
                if ks != list(range(ks[0], ks[0] + 4)):
                    continue
                if named[index[names[0]]][1].numel() % 64 or named[index[names[2]]][1].numel() % 64:
                    continue
                for k, n in zip(ks, names):
                    order[k] = index[n]
        return order

    def sync_lp(self):
        """refresh the bf16 shadow from the fp32 master weights (after init / load_state_dict / an optimizer
        other than the flat AdamW kernel)."""
        if self.flat_param_lp is not None:
            self.flat_param_lp.copy_(self.flat_param)

    def grad_view(self, param):
        """the slice of the flat gradient buffer that belongs to `param`."""
        return self._grad_of[id(param)]

    def _collect_grads(self, lo=0, hi=None):
        """autograd's per-parameter gradients -> flat buffer (one multi-tensor copy instead of one
        accumulate kernel per parameter); parameters the task did not touch keep their zero fill.
        `lo`, `hi`: only the parameters whose span starts inside that flat range."""
        dst, src = [], []
        i0 = bisect.bisect_left(self._span_starts, lo)
        i1 = len(self._params) if hi is None else bisect.bisect_left(self._span_starts, hi)
        for p, v in zip(self._params[i0:i1], self._grad_views[i0:i1]):
            if p.grad is not None:
                dst.append(v)
                src.append(p.grad)
                p.grad = None
        if dst:
            torch._foreach_add_(dst, src)      # (add, not copy: ops.linear may have accumulated into the view already)

    def _active_ranges(self, task):
        """contiguous flat ranges that received a gradient for `task` (found once per task)."""
        if task not in self._task_ranges:
            flags = torch.stack([(self.flat_grad[s:e] != 0).any() for _, s, e in self._spans]).tolist()
            if self.world > 1:    # every rank must agree on the ranges
                t = torch.tensor(flags, device=self.device, dtype=torch.int32)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                flags = t.bool().tolist()
            tops = {}
            for (n, s, e), f in zip(self._spans, flags):
                top = n.split('.')[0]
                lo, hi, any_f = tops.get(top, (s, e, False))
                tops[top] = (min(lo, s), max(hi, e), any_f or f)
            ranges = []
            for top, (lo, hi, f) in sorted(tops.items(), key=lambda kv: kv[1][0]):
                if not f:
                    continue
                if ranges and lo - ranges[-1][1] <= (1 << 18):     # bridge gaps <= 1 MB (e.g. the 35 k-param cls_head):
                    ranges[-1] = (ranges[-1][0], hi)                # reducing a few zeros beats one more collective
                else:
                    ranges.append((lo, hi))
            self._task_ranges[task] = ranges
        return self._task_ranges[task]

    # -- lr schedule (mmcv LrUpdaterHook family, by_epoch=False) ------------
    def lr_scale(self, it):
        """lr(it) / base_lr for the policies the reference's configs use (mmcv 1.6 runner/hooks/lr_updater.py):
        'step' (gamma, step list), 'CosineAnnealing' (min_lr_ratio), 'poly' (power, min_lr = 0), 'fixed'; optional
        'constant' / 'linear' / 'exp' warm-up over `warmup_iters` iterations starting at `warmup_ratio`.  Every
        param group scales by the same factor, so the flat AdamW kernel needs ONE device scalar."""
        c = self.lr_config
        if not c:
            return 1.0
        policy = c.get('policy', 'fixed')
        max_iters = c.get('max_iters', self.max_iters)
        if policy == 'step':
            steps = c['step']
            steps = [steps] if isinstance(steps, int) else steps
            scale = c.get('gamma', 0.1) ** sum(it >= s for s in steps)
        elif policy in ('CosineAnnealing', 'cosine'):
            if c.get('min_lr'):
                raise NotImplementedError('CosineAnnealing with an absolute min_lr is group dependent; use min_lr_ratio')
            r = c.get('min_lr_ratio', 0.0)
            scale = r + 0.5 * (1 - r) * (math.cos(math.pi * it / max_iters) + 1)
        elif policy == 'poly':
            if c.get('min_lr'):
                raise NotImplementedError('poly with a non-zero absolute min_lr is group dependent')
            scale = (1 - it / max_iters) ** c.get('power', 1.0)
        elif policy == 'fixed':
            scale = 1.0
        else:
            raise KeyError('lr policy %r is not supported' % policy)
        warm = c.get('warmup')
        if warm and it < c.get('warmup_iters', 0):
            ratio, n = c.get('warmup_ratio', 0.1), c['warmup_iters']
            if warm == 'constant':
                scale *= ratio
            elif warm == 'linear':
                scale *= 1 - (1 - it / n) * (1 - ratio)
            elif warm == 'exp':
                scale *= ratio ** (1 - it / n)
            else:
                raise KeyError('warmup %r is not supported' % warm)
        return scale

    def _update_lr(self):
        if not self.lr_config:
            return
        scale = self.lr_scale(self.iter)
        if scale == getattr(self, '_lr_scale_now', 1.0):
            return
        self._lr_scale_now = scale
        for g, base in zip(self.optimizer.param_groups, self._base_lrs):
            g['lr'] = base * scale
        if hasattr(self.optimizer, 'set_lr_scale'):
            self.optimizer.set_lr_scale(scale)                   # device scalar: graphs stay valid
        else:
            self._graphs.clear()                                 # lr is baked into captured launches: re-capture

    # -- gradient exchange (row 8e; reference: the DDP wrapper of mtl/apis/train.py:37-46) --------------------
    def _all_reduce_mean(self, lo, hi, async_op):
        """mean over ranks of flat_grad[lo:hi], in place.  NCCL averages inside the collective (ncclAvg: no separate
        1/world pass over the gradients); gloo (CPU tests) sums and divides."""
        seg = self.flat_grad[lo:hi]
        if _SKIP_EXCHANGE:            # (diagnostic only: measures what a multi-GPU step costs WITHOUT its gradient exchange)
            return _NoWork()
        if self._nvls is not None:
            nv = self._nvls
            cur, comm = torch.cuda.current_stream(self.device), nv['stream']
            comm.wait_stream(cur)                      # the range is final on this rank ...
            with torch.cuda.stream(comm):
                nv['hdl'].barrier(channel=0)           # ... and on every other rank
                _lib.call('rsc_nvls_allreduce_mean', nv['mc'], lo, (hi + 63) // 64 * 64, nv['rank'], self.world,
                          1.0 / self.world, nv['ctas'], comm.cuda_stream, alg_bytes=8 * (hi - lo))
                nv['hdl'].barrier(channel=0)           # every rank's share is stored everywhere
            work = _StreamWork(comm, self.device)
            if not async_op:
                work.wait()
            return work
        if self.device.type == 'cuda':
            return dist.all_reduce(seg, op=dist.ReduceOp.AVG, async_op=async_op)
        work = dist.all_reduce(seg, async_op=async_op)
        return _DivAfter(work, seg, self.world) if async_op else seg.div_(self.world)

    def _on_grad_ready(self, key):
        """Called by the backbone from inside backward (tensor hooks on its outputs / stage inputs): the engine runs
        nodes in reverse creation order, so when the gradient of stage K's input is about to be consumed every
        parameter used after it in forward -- flat offsets >= frontier(K) -- has its final gradient.  That tail of the
        task's active range(s) is all-reduced NOW, asynchronously (NCCL's own stream; captured as a parallel branch
        of the step's CUDA graph), while backward continues through the earlier stages."""
        st = self._pending
        if st is None or key not in self._frontier_of:
            return
        new = self._frontier_of[key]
        if new >= st['frontier']:
            return
        pieces = [(max(lo, new), min(hi, st['frontier'])) for lo, hi in st['ranges']]
        pieces = [(a, b) for a, b in pieces if b > a]
        if sum(b - a for a, b in pieces) < self._min_bucket:
            return                                    # too small to pay for a collective: rides with the next bucket
        self._collect_grads(new, st['frontier'])
        for a, b in pieces:
            st['works'].append(self._all_reduce_mean(a, b, True))
        st['frontier'] = new

    def _exchange_grads(self, task):
        """after backward: reduce what the in-backward buckets have not covered, then join them."""
        st, self._pending = self._pending, None
        if st is None:                                # first iteration of this task: the ranges are found now
            self._collect_grads()
            for lo, hi in self._active_ranges(task):
                self._all_reduce_mean(lo, hi, False)
            self._join_log_works()
            return
        self._collect_grads(0, st['frontier'])
        for lo, hi in st['ranges']:
            if lo < st['frontier']:
                st['works'].append(self._all_reduce_mean(lo, min(hi, st['frontier']), True))
        self.last_buckets = len(st['works'])
        for w in st['works']:
            w.wait()
        self._join_log_works()

    def _join_log_works(self):
        from ...models import mtl as _mtl
        if _mtl.ASYNC_LOG_WORKS:
            for t in _mtl.ASYNC_LOG_WORKS:
                t._rsc_work.wait()
                t._rsc_work = None
            del _mtl.ASYNC_LOG_WORKS[:]

    # -- one co-training iteration ------------------------------------------
    def _backward_and_step(self, outputs, task):
        """OptimizerHook.after_train_iter: backward, gradient exchange, global-norm clip, AdamW."""
        if self.world > 1:
            ranges = self._task_ranges.get(task)
            self._pending = dict(ranges=ranges, frontier=self.flat_grad.numel(), works=[]) if ranges else None
            try:
                outputs['loss'].backward()
            except BaseException:
                self._pending = None
                raise
            self._exchange_grads(task)
        else:
            outputs['loss'].backward()
            self._collect_grads()
        clip_coef = None
        if self.grad_clip:
            max_norm = float(self.grad_clip['max_norm'])
            total_norm = torch.linalg.vector_norm(self.flat_grad, float(self.grad_clip.get('norm_type', 2)))
            clip_coef = torch.clamp(max_norm / (total_norm + 1e-6), max=1.0)
            self.last_grad_norm = total_norm
        if hasattr(self.optimizer, 'step_flat'):
            self.optimizer.step_flat(clip_coef)                  # clip scale folded into the AdamW kernel
        else:
            if clip_coef is not None:
                self.flat_grad.mul_(clip_coef)
            for p, v in zip(self._params, self._grad_views):
                p.grad = v
            self.optimizer.step()
            for p in self._params:
                p.grad = None
            self.sync_lp()

    def _autocast(self):
        return torch.autocast('cuda', dtype=self.compute_dtype, enabled=self.compute_dtype != torch.float32)

    def _set_gemm_sms(self, batch):
        """per step (the grid is baked into a captured graph, so this runs before eager steps and captures only)"""
        if self._dp_gemm_sms:
            img = batch.get('img') if isinstance(batch, dict) else None
            n = img.shape[0] if torch.is_tensor(img) else (len(img) if isinstance(img, (list, tuple)) else 1 << 30)
            _lib.call('rsc_set_gemm_sms', self._dp_gemm_sms if n <= self._dp_gemm_max_batch else 0, 0)

    def _train_iter_eager(self, data):
        self._set_gemm_sms(data)
        self.flat_grad.zero_()
        with self._autocast():
            outputs = self.model.train_step(data, self.optimizer)
        self._backward_and_step(outputs, data['task'])
        return outputs

    @staticmethod
    def _signature(batch):
        def sig(o):
            if torch.is_tensor(o):
                return (tuple(o.shape), str(o.dtype))
            if isinstance(o, (list, tuple)):
                return tuple(sig(x) for x in o)
            if isinstance(o, dict):
                return tuple((k, sig(v)) for k, v in sorted(o.items()) if k == 'img_shape' or
                             torch.is_tensor(v) or isinstance(v, (list, tuple, dict)))
            return o if isinstance(o, (int, float, str, bool, type(None))) else None
        return sig(batch)

    @staticmethod
    def _copy_in(static, batch):
        if torch.is_tensor(static):
            static.copy_(batch, non_blocking=True)
        elif isinstance(static, (list, tuple)):
            for s_, b_ in zip(static, batch):
                StepEngine._copy_in(s_, b_)
        elif isinstance(static, dict):
            for k in static:
                if k in batch:                # (the model may have annotated its static img_metas)
                    StepEngine._copy_in(static[k], batch[k])

    def _capture(self, st, batch):
        """Record one iteration of this (task, shapes) into CUDA graphs: graph A = zero grads +
        forward (+ loss/backward/clip/AdamW when the task has no host phase); for det, graph B =
        losses + backward + clip + AdamW after the host-side Hungarian matching."""
        static = _static_copy(batch, self.device)
        task = static['task']
        self._set_gemm_sms(static)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        gA, gB = torch.cuda.CUDAGraph(), None
        with torch.cuda.graph(gA):
            self.flat_grad.zero_()
            with self._autocast():
                ctx = self.model.train_step_begin(static)
                outputs = self.model.train_step_finish(ctx) if ctx['pending'] is None else None
            if outputs is not None:
                self._backward_and_step(outputs, task)
        nA = _lib.launch_count() - n0
        nB = 0
        if outputs is None:
            gA.replay()                                   # real forward so the matching sees real costs
            self.model.train_step_host(ctx)
            n1 = _lib.launch_count()
            gB = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gB, pool=gA.pool()):
                with self._autocast():
                    outputs = self.model.train_step_finish(ctx)
                self._backward_and_step(outputs, task)
            nB = _lib.launch_count() - n1
            gB.replay()
        else:
            gA.replay()
        st.update(gA=gA, gB=gB, ctx=ctx, static=static, outputs=outputs, launches=nA + nB, grad_norm=self.last_grad_norm)

    def prefetch(self, data_batch):
        """Start the host->device copy of a (pinned) host batch on the copy stream while the current step is still
        running; the train_iter call that receives the same batch object then only does a device-to-device copy
        into the captured graph's input buffers.  (H2D prefetch on a copy stream.)  A no-op
        until the batch's (task, shapes) signature has been captured."""
        if self._copy_stream is None:
            return
        st = self._graphs.get(self._signature(data_batch))
        if not st or 'static' not in st:
            return
        cs = self._copy_stream
        if 'stage' not in st:
            st['stage'] = _static_copy(st['static'], self.device)
            cs.wait_stream(torch.cuda.current_stream(self.device))     # (the clones above ran on the main stream)
        if st.get('stage_free') is not None:
            cs.wait_event(st['stage_free'])              # the previous consumer of the staging buffers is done
        with torch.cuda.stream(cs):
            self._copy_in(st['stage'], data_batch)
            ev = torch.cuda.Event()
            ev.record(cs)
        st['staged'] = (id(data_batch), ev)

    def _graph_slot(self, key):
        """Bookkeeping entry of a (task, shapes) signature, or None when it may not get a CUDA graph.  Every captured
        graph pins its activations, so their number is bounded (`max_graphs`, RSC_MAX_GRAPHS): real detection batches
        have a different signature per ground-truth count.  A signature is captured only after `graph_warmup` eager
        runs (rare ones never are); when the bound is reached a new signature takes the slot of a captured one that
        has not been replayed for `graph_idle_iters` iterations, and runs eagerly otherwise."""
        st = self._graphs.get(key)
        if st is not None:
            st['last'] = self.iter
            return st
        captured = [(v.get('last', 0), k) for k, v in self._graphs.items() if 'gA' in v]
        if len(captured) >= self.max_graphs:
            last, victim = min(captured)
            if self.iter - last < self.graph_idle_iters:
                return None
            del self._graphs[victim]
        if len(self._graphs) > 4096:                      # forget the oldest never-captured signatures
            for k in sorted((k for k, v in self._graphs.items() if 'gA' not in v), key=lambda k: self._graphs[k].get('last', 0))[:2048]:
                del self._graphs[k]
        st = self._graphs[key] = dict(eager=0, last=self.iter)
        return st

    def train_iter(self, data_batch):
        """model.train_step + OptimizerHook.after_train_iter.  Returns train_step's outputs.
        After `graph_warmup` eager iterations of a (task, shapes) signature the iteration is
        replayed from CUDA graphs (launch-bound otherwise: ~4000 small kernels per det step)."""
        self._update_lr()
        if not self.use_graphs or _lib._timer is not None:
            outputs = self._train_iter_eager(_to_device(data_batch, self.device))
            self.iter += 1
            return outputs
        key = self._signature(data_batch)
        st = self._graph_slot(key)
        if st is None:                                    # too many distinct (task, shapes) signatures: no graph for this one
            outputs = self._train_iter_eager(_to_device(data_batch, self.device))
            self.iter += 1
            return outputs
        if 'gA' not in st:
            if st['eager'] < self.graph_warmup:
                st['eager'] += 1
                outputs = self._train_iter_eager(_to_device(data_batch, self.device))
                self.iter += 1
                return outputs
            try:
                self._capture(st, data_batch)             # captures AND performs this iteration
            except Exception as e:                        # keep training eagerly, but say so loudly
                import sys
                import traceback
                traceback.print_exc()
                print('[rscotr_b200] CUDA-graph capture failed for task %r (%s: %s); this signature runs eagerly'
                      % (data_batch.get('task'), type(e).__name__, e), file=sys.stderr)
                st.clear()
                st.update(eager=-(1 << 60))               # never try again
                self.graph_failures += 1
                torch.cuda.synchronize()
                outputs = self._train_iter_eager(_to_device(data_batch, self.device))
                self.iter += 1
                return outputs
        else:
            staged = st.pop('staged', None)
            if staged is not None and staged[0] == id(data_batch):
                cur = torch.cuda.current_stream(self.device)
                cur.wait_event(staged[1])
                self._copy_in(st['static'], st['stage'])          # device-to-device
                st['stage_free'] = torch.cuda.Event()
                st['stage_free'].record(cur)
            else:
                self._copy_in(st['static'], data_batch)
            st['gA'].replay()
            if st['gB'] is not None:
                self.model.train_step_host(st['ctx'])
                st['gB'].replay()
        self.replayed_launches += st['launches']
        self.last_grad_norm = st.get('grad_norm')         # (the norm tensor this signature's graph writes)
        self.iter += 1
        out = st['outputs']
        return dict(loss=out['loss'], log_vars=out['log_vars'].rebind(), num_samples=out['num_samples'])
