"""MTL: shared Swin backbone + neck + shared deformable encoder + three task heads,
one task per iteration.  Same constructor arguments, method names, train_step
contract and loss/log keys as the reference's models/multi/multitask_learner.py:34-353."""
from collections import OrderedDict

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from ..config import MODELS, build_from_cfg
from .bricks import MultiScaleDeformableAttention, PackedLosses, build_transformer_layer_sequence, const_tensor
from .cls_head import Augments
from .seg_head import resize

supported_tasks = ('cls', 'det', 'seg')


def add_prefix(inputs, prefix):
    return {'%s.%s' % (prefix, k): v for k, v in inputs.items()}


def _build(cfg):
    if isinstance(cfg, nn.Module) or cfg is None:
        return cfg
    return build_from_cfg(dict(cfg), MODELS)


# set to a list by the step engine (data-parallel CUDA runs): pending asynchronous log-value reductions of the current
# iteration, joined by StepEngine._backward_and_step
ASYNC_LOG_WORKS = None


def normalize_on_device(img, img_metas):
    """Device half of a deferred `Normalize` (mtl/data/transforms.py, defer=True): the loader ships the batch
    as uint8 BGR (4x fewer H2D bytes than fp32); here (x[RGB] - mean) / std is applied on the device and the
    right/bottom padding is re-zeroed (the reference pads AFTER normalising, with 0).  A float batch passes
    through untouched."""
    if not torch.is_tensor(img) or img.dtype != torch.uint8:
        return img
    cfg = img_metas[0]['img_norm_cfg']
    H, W = img.shape[-2:]
    shapes = [tuple(m.get('img_shape', (H, W))[:2]) for m in img_metas]
    if img.is_cuda and W % 4 == 0:            # one kernel: flip + normalise + padding re-zeroed (rsc_normalize_u8)
        from .. import ops
        valid = const_tensor([list(s) for s in shapes], torch.int32, img.device) if any(s != (H, W) for s in shapes) else None
        return ops.normalize_u8(img, const_tensor([float(v) for v in cfg['mean']], torch.float32, img.device),
                                const_tensor([1.0 / float(v) for v in cfg['std']], torch.float32, img.device), valid,
                                cfg.get('to_rgb', True))
    mean = const_tensor([float(v) for v in cfg['mean']], torch.float32, img.device).view(1, -1, 1, 1)
    inv = const_tensor([1.0 / float(v) for v in cfg['std']], torch.float32, img.device).view(1, -1, 1, 1)
    x = img.flip(1) if cfg.get('to_rgb', True) else img
    x = (x.float() - mean) * inv
    if any(s != (H, W) for s in shapes):
        hs = const_tensor([s[0] for s in shapes], torch.int64, img.device).view(-1, 1, 1, 1)
        ws = const_tensor([s[1] for s in shapes], torch.int64, img.device).view(-1, 1, 1, 1)
        ys = torch.arange(H, device=img.device).view(1, 1, H, 1)
        xs = torch.arange(W, device=img.device).view(1, 1, 1, W)
        x = x * ((ys < hs) & (xs < ws))
    return x


def bbox2result(bboxes, labels, num_classes):
    if bboxes.shape[0] == 0:
        import numpy as np
        return [np.zeros((0, 5), dtype=np.float32) for _ in range(num_classes)]
    bboxes, labels = bboxes.detach().cpu().numpy(), labels.detach().cpu().numpy()
    return [bboxes[labels == i, :] for i in range(num_classes)]


@MODELS.register_module()
class MTL(nn.Module):
    PALETTE = None

    def __init__(self, backbone, neck, shared_encoder, cls_head=None, bbox_head=None, seg_head=None, task_weight=None,
                 train_cfg=None, test_cfg=None, init_cfg=None):
        super().__init__()
        self.backbone = _build(backbone)
        self.neck = _build(neck)
        self.shared_encoder = build_transformer_layer_sequence(shared_encoder)
        self.task_weight = dict(cls=1, det=1, seg=1)
        if task_weight is not None:
            assert isinstance(task_weight, dict)
            self.task_weight.update(task_weight)
        train_cfg = train_cfg or dict(cls=dict(), det=None, seg=dict())
        test_cfg = test_cfg or dict(cls=dict(), det=dict(max_per_img=100), seg=dict(mode='whole'))
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.cls_augments = None
        cls_augments_cfg = (train_cfg.get('cls') or {}).get('augments', None)
        if cls_augments_cfg is not None:
            self.cls_augments = Augments(cls_augments_cfg)
        if bbox_head is not None and not isinstance(bbox_head, nn.Module):
            bbox_head = dict(bbox_head)
            bbox_head.update(train_cfg=train_cfg.get('det'))
            bbox_head.update(test_cfg=test_cfg.get('det'))
        self.task_pretrain = train_cfg.get('task_pretrain', None)
        self.cls_head = _build(cls_head)
        self.bbox_head = _build(bbox_head)
        self.seg_head = _build(seg_head)
        self.CLASSES = None

    def init_weights(self):
        # (every sub-module initialises itself on construction; what mmcv's recursive init_weights adds on top is
        # the backbone's `Pretrained` checkpoint)
        bb = self.backbone
        if getattr(bb, 'pretrained', None) or (isinstance(getattr(bb, 'init_cfg', None), dict)
                                               and bb.init_cfg.get('type') == 'Pretrained'):
            bb.init_weights()
        for layer in self.shared_encoder.layers:
            for attn in layer.attentions:
                if isinstance(attn, MultiScaleDeformableAttention):
                    attn.init_weights()

    def extract_feat(self, img):
        backbone_feature = self.backbone(img)
        neck_feature = self.neck(backbone_feature[-3:])
        return neck_feature, backbone_feature

    def forward_train(self, task, *args, **kwargs):
        assert task in supported_tasks
        return getattr(self, 'forward_train_%s' % task)(*args, **kwargs)

    def forward_test(self, task, img, img_metas, *args, **kwargs):
        if isinstance(task, list):
            task = list(set(task))
            if len(task) == 1:
                task = task[0]
            else:
                raise NotImplementedError('The current implementation only support same task in a batch')
        if isinstance(img, list):
            if len(img) != 1:
                raise NotImplementedError('The current implementation does not support TTA ')
            img = img[0]
        if isinstance(img_metas[0], list):
            img_metas = img_metas[0]
        return self.simple_test(task, img, img_metas, *args, **kwargs)

    def simple_test(self, task, *args, **kwargs):
        assert task in supported_tasks
        return getattr(self, 'simple_test_%s' % task)(*args, **kwargs)

    def _extract_feat_cls(self, img):
        """The reference runs the neck for every task and the single-level cls head then discards its output
        (multitask_learner.py:122 -> :84).  The neck has no buffers, so skipping it when the head does
        not read it changes no result and no gradient."""
        if getattr(self.cls_head, 'uses_neck', True):
            return self.extract_feat(img)
        return None, self.backbone(img)

    def forward_train_cls(self, img, gt_label, **kwargs):
        if self.cls_augments is not None:
            img, gt_label = self.cls_augments(img, gt_label)
        neck_feature, backbone_feature = self._extract_feat_cls(img)
        losses = dict()
        losses.update(self.cls_head.forward_train(neck_feature, backbone_feature, gt_label, self.shared_encoder))
        return losses

    def forward_train_det(self, img, img_metas, gt_bboxes, gt_labels, gt_bboxes_ignore=None):
        batch_input_shape = tuple(img[0].size()[-2:])
        for img_meta in img_metas:
            img_meta['batch_input_shape'] = batch_input_shape
        x = self.extract_feat(img)[0]
        return self.bbox_head.forward_train(x, img_metas, gt_bboxes, gt_labels, gt_bboxes_ignore, self.shared_encoder)

    def forward_train_seg(self, img, img_metas, gt_semantic_seg):
        neck_feature, backbone_feature = self.extract_feat(img)
        losses = dict()
        loss_decode = self.seg_head.forward_train(neck_feature, backbone_feature, img_metas, gt_semantic_seg,
                                                  self.shared_encoder)
        losses.update(add_prefix(loss_decode, 'seg'))
        return losses

    def simple_test_cls(self, img, img_metas=None, **kwargs):
        neck_feature, backbone_feature = self._extract_feat_cls(img)
        return self.cls_head.simple_test(neck_feature, backbone_feature, shared_encoder=self.shared_encoder, **kwargs)

    def simple_test_det(self, img, img_metas, rescale=False):
        for m in img_metas:
            m['batch_input_shape'] = tuple(img.size()[-2:])
        feat = self.extract_feat(img)[0]
        results_list = self.bbox_head.simple_test(feat, img_metas, rescale=rescale, shared_encoder=self.shared_encoder)
        return [bbox2result(b, l, self.bbox_head.num_classes) for b, l in results_list]

    def whole_inference_seg(self, img, img_meta, rescale):
        neck_feature, backbone_feature = self.extract_feat(img)
        seg_logit = self.seg_head.forward_test(neck_feature, backbone_feature, img_meta, self.shared_encoder)
        seg_logit = resize(seg_logit, size=img.shape[2:])
        if rescale:
            resize_shape = img_meta[0]['img_shape'][:2]
            seg_logit = seg_logit[:, :, :resize_shape[0], :resize_shape[1]]
            seg_logit = resize(seg_logit.contiguous(), size=img_meta[0]['ori_shape'][:2])
        return seg_logit

    def inference_seg(self, img, img_meta, rescale):
        assert self.test_cfg['seg']['mode'] in ['whole']
        ori_shape = img_meta[0]['ori_shape']
        assert all(_['ori_shape'] == ori_shape for _ in img_meta)
        seg_logit = self.whole_inference_seg(img, img_meta, rescale)
        output = F.softmax(seg_logit.float(), dim=1)
        if img_meta[0].get('flip', False):
            flip_direction = img_meta[0]['flip_direction']
            assert flip_direction in ['horizontal', 'vertical']
            output = output.flip(dims=(3,)) if flip_direction == 'horizontal' else output.flip(dims=(2,))
        return output

    def simple_test_seg(self, img, img_meta, rescale=True):
        seg_pred = self.inference_seg(img, img_meta, rescale).argmax(dim=1)
        return list(seg_pred.cpu().numpy())

    # train_step in three phases (device / host / device) so the step engine can capture the device
    # phases in CUDA graphs; only the det task has host work (Hungarian matching) in the middle.
    def train_step_begin(self, data):
        task = data.get('task', None)
        if torch.is_tensor(data['img']) and data['img'].dtype == torch.uint8:
            data = dict(data, img=normalize_on_device(data['img'], data['img_metas']))
        if task == 'det' and hasattr(self.bbox_head, 'forward_train_begin') and not (
                getattr(self.bbox_head, 'fused_loss', False) and data['img'].is_cuda):
            # (with the GPU matching + fused loss kernels the det step has no host phase at all)
            img, img_metas = data['img'], data['img_metas']
            batch_input_shape = tuple(img[0].size()[-2:])
            for img_meta in img_metas:
                img_meta['batch_input_shape'] = batch_input_shape
            x = self.extract_feat(img)[0]
            pend = self.bbox_head.forward_train_begin(x, img_metas, data['gt_bboxes'], data['gt_labels'],
                                                      data.get('gt_bboxes_ignore'), self.shared_encoder)
            return dict(data=data, pending=pend, losses=None)
        return dict(data=data, pending=None, losses=self(**data))

    def train_step_host(self, ctx):
        if ctx['pending'] is not None:
            self.bbox_head.loss_assign(ctx['pending'])

    def train_step_finish(self, ctx):
        data = ctx['data']
        losses = ctx['losses'] if ctx['pending'] is None else self.bbox_head.loss_finish(ctx['pending'])
        return self._finish(losses, data)

    def train_step(self, data, optimizer):
        losses = self(**data)
        return self._finish(losses, data)

    def _finish(self, losses, data):
        loss, keys, packed = self._parse_losses(losses)
        task = data.get('task', None)
        dataset_name = data.get('dataset_name', None)
        weight = self.task_weight[task] if hasattr(self, 'task_weight') else 1
        loss = loss * weight
        log_vars = _LazyLogVars(['%s.%s.%s' % (task, dataset_name, k) for k in keys], packed, weight)
        return dict(loss=loss, log_vars=log_vars, num_samples=len(data['img_metas']))

    def val_step(self, data, optimizer=None):
        losses = self(**data)
        loss, keys, packed = self._parse_losses(losses)
        log_vars = _LazyLogVars(['%s.%s.%s' % (data.get('task', None), data.get('dataset_name', None), k)
                                 for k in keys], packed, 1)
        return dict(loss=loss, log_vars=log_vars, num_samples=len(data['img_metas']))

    def forward(self, task, img, img_metas, return_loss=True, dataset_name=None, **kwargs):
        if isinstance(img, list):
            img = [normalize_on_device(i, m) for i, m in zip(img, img_metas)]
        else:
            img = normalize_on_device(img, img_metas)
        if return_loss:
            return self.forward_train(task=task, img=img, img_metas=img_metas, **kwargs)
        return self.forward_test(task=task, img=img, img_metas=img_metas, **kwargs)

    def _parse_losses(self, losses):
        """Same totals / log keys as multitask_learner.py:274-306, but ONE device->host
        transfer and (distributed) ONE packed all-reduce instead of one per log var."""
        if isinstance(losses, PackedLosses):        # all terms already in one tensor (fused loss kernels)
            keys, packed = list(losses.key_list), losses.packed.float()
            sel = [i for i, k in enumerate(keys) if 'loss' in k]
            loss = packed.sum() if len(sel) == len(keys) else packed[const_tensor(sel, torch.long, packed.device)].sum()
            keys.append('loss')
            packed = torch.cat([packed.detach(), loss.detach().reshape(1)])
            return loss, keys, self._reduce_log_vars(keys, packed)
        log_vars = OrderedDict()
        for loss_name, loss_value in losses.items():
            if isinstance(loss_value, torch.Tensor):
                log_vars[loss_name] = loss_value.mean()
            elif isinstance(loss_value, list):
                log_vars[loss_name] = sum(_loss.mean() for _loss in loss_value)
            else:
                raise TypeError('%s is not a tensor or list of tensors' % loss_name)
        loss = sum(_value for _key, _value in log_vars.items() if 'loss' in _key)
        log_vars['loss'] = loss
        packed = torch.stack([v.detach().float().reshape(()) for v in log_vars.values()])
        return loss, list(log_vars.keys()), self._reduce_log_vars(list(log_vars.keys()), packed)

    @staticmethod
    def _reduce_log_vars(keys, packed):
        """cross-rank mean of the log values with ONE packed all-reduce (reference: one all_reduce per log
        var, multitask_learner.py:289-304).  The reference's "same number of log vars on every rank" assertion
        rides along as element 0 and is checked when the values are read on the host (_LazyLogVars), so the
        step stays free of host syncs and CUDA-graph capturable."""
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size()
            n = torch.cat([const_tensor([float(len(keys))], torch.float32, packed.device), packed])
            if ASYNC_LOG_WORKS is not None and n.is_cuda:
                # nothing in the step reads the logged values: average on NCCL's stream (ncclAvg) and let the step engine
                # join at the end of the iteration -- no mid-step rendezvous of the ranks for a logging collective
                n._rsc_work = dist.all_reduce(n, op=dist.ReduceOp.AVG, async_op=True)
                ASYNC_LOG_WORKS.append(n)
                return n
            dist.all_reduce(n)
            return n / world
        return packed

    def load_task_pretrain(self):
        if self.task_pretrain is None:
            print('You did not set task_pretrain, hence it is skipped.')
            return
        rule = self.task_pretrain.get('rule', None)
        sd = torch.load(self.task_pretrain['pretrained'], map_location='cpu')
        if 'state_dict' in sd:
            sd = sd['state_dict']
        if rule == 'dino_mmdet':
            out = OrderedDict()
            for name, param in sd.items():
                if name.startswith('neck') and name.endswith('conv.bias'):
                    continue
                new_name = name.replace('bbox_head.transformer.encoder', 'shared_encoder', 1) \
                    if name.startswith('bbox_head.transformer.encoder') else name
                assert new_name not in out, '%s-->%s' % (name, new_name)
                out[new_name] = param
            sd = out
        incompatible = self.load_state_dict(sd, strict=False)
        print('load task pretrain of rule:%s\nincompatiblekeys: %s' % (rule, incompatible))


class _LazyLogVars(OrderedDict):
    """log_vars (name -> float, the reference's contract) whose values live in ONE packed device
    tensor and are materialised (one D2H copy) on first read: a training loop that only logs every
    N iterations never synchronises in between, and the step stays CUDA-graph capturable."""

    def __init__(self, keys, packed, weight=1):
        super().__init__()
        self._keys, self._packed, self._weight, self._done = list(keys), packed, weight, False
        for k in self._keys:
            OrderedDict.__setitem__(self, k, None)

    def _materialise(self):
        if not self._done:
            w = getattr(self._packed, '_rsc_work', None)      # an asynchronous reduction nobody has joined (a step outside the
            if w is not None:                                   # engine, e.g. validation): join before reading
                w.wait()
                self._packed._rsc_work = None
            vals = self._packed.tolist()
            if len(vals) == len(self._keys) + 1:      # distributed: element 0 = mean number of log vars
                assert int(round(vals[0])) == len(self._keys), \
                    'loss log variables are different across GPUs!\nlen(log_vars): %d keys: %s' % (
                        len(self._keys), ','.join(self._keys))
                vals = vals[1:]
            for k, v in zip(self._keys, vals):
                OrderedDict.__setitem__(self, k, v * self._weight)
            self._done = True

    def __getitem__(self, k):
        self._materialise()
        return OrderedDict.__getitem__(self, k)

    def get(self, k, default=None):
        self._materialise()
        return OrderedDict.get(self, k, default)

    def items(self):
        self._materialise()
        return OrderedDict.items(self)

    def values(self):
        self._materialise()
        return OrderedDict.values(self)

    def rebind(self):
        """a fresh, un-materialised view of the same packed tensor (CUDA-graph replays)."""
        return _LazyLogVars(self._keys, self._packed, self._weight)


class SingleTaskModel(nn.Module):
    """step-engine interface (train_step / train_step_begin / _host / _finish, same contract as MTL) for the
    single-task models of the reference's configs/cls and the UPerNet configuration: `forward(return_loss=True)`
    returns a loss dict, terms whose key contains 'loss' are summed, everything is logged."""
    default_task = None

    def _parse_losses(self, losses):
        keys = list(losses.keys())
        vals = torch.stack([losses[k].mean().float() for k in keys])
        mask = const_tensor([1.0 if 'loss' in k else 0.0 for k in keys], torch.float32, vals.device)
        loss = (vals * mask).sum()
        packed = torch.cat([vals.detach(), loss.detach().view(1)])
        return loss, keys + ['loss'], packed

    def train_step(self, data, optimizer=None):
        data = dict(data)
        task, name = data.pop('task', self.default_task), data.pop('dataset_name', None)
        losses = self(**data)
        loss, keys, packed = self._parse_losses(losses)
        packed = MTL._reduce_log_vars(keys, packed)
        prefix = '%s.%s.' % (task, name) if name is not None else ''
        return dict(loss=loss, log_vars=_LazyLogVars([prefix + k for k in keys], packed, 1), num_samples=len(data['img_metas']))

    def train_step_begin(self, data):
        return dict(data=data, pending=None, outputs=self.train_step(data))

    def train_step_host(self, ctx):
        pass

    def train_step_finish(self, ctx):
        return ctx['outputs']
