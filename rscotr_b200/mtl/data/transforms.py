"""The data pipelines the reference's three dataset configs name, DataContainer-free.

Reference sites: configs/_base_/cls/resisc_swin_224.py:8-56 (+ rand_aug.py), configs/_base_/det/dior.py:11-40,
configs/_base_/seg/potsdam_IRRG_all.py:8-56.  The transform classes themselves are third-party there
(mmcls 0.23 / mmdet 2.25 / mmseg 0.28 `datasets/pipelines`, image ops of mmcv 1.6 `mmcv/image`); they are
restated here with the same arguments, the same `results` dict keys, and the same order of random draws
from `numpy.random` / `random` (so the libraries' worker seeding carries over).  One registry serves the
three tasks; where the libraries disagree on an argument name (flip_ratio / prob / flip_prob,
img_scale / size) every spelling is accepted.

Images stay uint8 HWC BGR until `Normalize`.  With `defer=True` Normalize only records `img_norm_cfg`
(+ `norm_deferred`) and the batch travels to the GPU as uint8 (4x fewer H2D bytes); the model-side
`normalize_on_device` then does the subtraction/scale (and re-zeroes the padding) on the device.
"""
import inspect
import math
import os
import random

import cv2
import numpy as np
import torch

PIPELINES = {}


def register(cls):
    PIPELINES[cls.__name__] = cls
    return cls


def build_transform(cfg, task=None):
    cfg = dict(cfg)
    t = cfg.pop('type')
    if t not in PIPELINES:
        raise KeyError('%s is not in the pipeline registry' % t)
    cls = PIPELINES[t]
    if 'task' in inspect.signature(cls.__init__).parameters:
        cfg.setdefault('task', task)
    return cls(**cfg)


class Compose:
    def __init__(self, transforms, task=None):
        self.transforms = [build_transform(t, task) if isinstance(t, dict) else t for t in transforms]

    def __call__(self, data):
        for t in self.transforms:
            data = t(data)
            if data is None:
                return None
        return data


# ------------------------------------------------------------------------------------------ mmcv.image
_CV2_INTERP = dict(nearest=cv2.INTER_NEAREST, bilinear=cv2.INTER_LINEAR, bicubic=cv2.INTER_CUBIC, area=cv2.INTER_AREA,
                   lanczos=cv2.INTER_LANCZOS4)


def imresize(img, size, interpolation='bilinear', backend='cv2'):
    """size = (w, h)."""
    if backend == 'pillow':
        from PIL import Image
        assert img.dtype == np.uint8, 'Pillow backend only support uint8 type'
        codes = dict(nearest=Image.NEAREST, bilinear=Image.BILINEAR, bicubic=Image.BICUBIC, box=Image.BOX,
                     lanczos=Image.LANCZOS, hamming=Image.HAMMING)
        return np.array(Image.fromarray(img).resize(size, codes[interpolation]))
    return cv2.resize(img, size, interpolation=_CV2_INTERP[interpolation])


def rescale_size(old_size, scale):
    """mmcv.rescale_size: (w, h) scaled by a factor, or to fit inside (long edge, short edge)."""
    w, h = old_size
    if isinstance(scale, (float, int)):
        if scale <= 0:
            raise ValueError('Invalid scale %s, must be positive.' % scale)
        f = scale
    else:
        f = min(max(scale) / max(h, w), min(scale) / min(h, w))
    return int(w * float(f) + 0.5), int(h * float(f) + 0.5)


def imrescale(img, scale, interpolation='bilinear', backend='cv2'):
    h, w = img.shape[:2]
    return imresize(img, rescale_size((w, h), scale), interpolation, backend)


def imflip(img, direction='horizontal'):
    if direction == 'horizontal':
        return np.flip(img, axis=1)
    if direction == 'vertical':
        return np.flip(img, axis=0)
    assert direction == 'diagonal'
    return np.flip(img, axis=(0, 1))


def imnormalize(img, mean, std, to_rgb=True):
    img = img.copy().astype(np.float32)
    mean = np.float64(np.asarray(mean).reshape(1, -1))
    stdinv = 1 / np.float64(np.asarray(std).reshape(1, -1))
    if to_rgb:
        cv2.cvtColor(img, cv2.COLOR_BGR2RGB, img)
    cv2.subtract(img, mean, img)
    cv2.multiply(img, stdinv, img)
    return img


def impad(img, shape=None, pad_val=0):
    """pad right / bottom to shape = (h, w)."""
    h, w = img.shape[:2]
    pad = (0, 0, max(shape[1] - w, 0), max(shape[0] - h, 0))
    return cv2.copyMakeBorder(img, pad[1], pad[3], pad[0], pad[2], cv2.BORDER_CONSTANT, value=pad_val)


def impad_to_multiple(img, divisor, pad_val=0):
    h = int(np.ceil(img.shape[0] / divisor)) * divisor
    w = int(np.ceil(img.shape[1] / divisor)) * divisor
    return impad(img, (h, w), pad_val)


def bgr2gray(img):
    return cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)


def _blend(img, factor, degenerated, as_float=True):
    if as_float:
        out = cv2.addWeighted(img.astype(np.float32), factor, degenerated.astype(np.float32), 1 - factor, 0)
    else:
        out = cv2.addWeighted(img, factor, degenerated, 1 - factor, 0)
    return np.clip(out, 0, 255).astype(img.dtype)


def adjust_color(img, alpha=1.):
    gray = np.tile(bgr2gray(img)[..., None], [1, 1, 3])
    out = cv2.addWeighted(img, alpha, gray, 1 - alpha, 0)
    if not out.dtype == np.uint8:
        out = np.clip(out, 0, 255)
    return out.astype(img.dtype)


def adjust_brightness(img, factor=1.):
    return _blend(img, factor, np.zeros_like(img))


def adjust_contrast(img, factor=1.):
    gray = bgr2gray(img)
    hist = np.histogram(gray, 256, (0, 255))[0]
    mean = round(np.sum(gray) / np.sum(hist))
    deg = cv2.cvtColor((np.ones_like(img[..., 0]) * mean).astype(img.dtype), cv2.COLOR_GRAY2BGR)
    return _blend(img, factor, deg)


def adjust_sharpness(img, factor=1.):
    kernel = np.array([[1., 1., 1.], [1., 5., 1.], [1., 1., 1.]]) / 13
    return _blend(img, factor, cv2.filter2D(img, -1, kernel), as_float=False)


def auto_contrast(img, cutoff=0):
    def tune(im):
        hist = np.histogram(im, 256, (0, 255))[0]
        cdf = np.cumsum(hist)
        cut_low, cut_high = cdf[-1] * cutoff // 100, cdf[-1] - cdf[-1] * cutoff // 100
        cdf = np.clip(cdf, cut_low, cut_high) - cut_low
        hist = np.concatenate([[cdf[0]], np.diff(cdf)], 0)
        nz = np.nonzero(hist)[0]
        low, high = (nz[0], nz[-1]) if len(nz) else (0, 0)
        if low >= high:
            return im
        scale = 255.0 / (high - low)
        lut = np.clip(np.array(range(256)) * scale + (-low * scale), 0, 255)
        return lut[im]
    return np.stack([tune(img[..., c]) for c in range(3)], -1).astype(img.dtype)


def imequalize(img):
    def scale(im):
        hist = np.histogram(im, 256, (0, 255))[0]
        nz = hist[hist != 0]
        step = (np.sum(nz) - nz[-1]) // 255
        if not step:
            lut = np.array(range(256))
        else:
            lut = (np.cumsum(hist) + (step // 2)) // step
            lut = np.concatenate([[0], lut[:-1]], 0)
            lut[lut > 255] = 255
        return np.where(np.equal(step, 0), im, lut[im])
    return np.stack([scale(img[..., c]) for c in range(3)], -1).astype(img.dtype)


def imrotate(img, angle, center=None, scale=1.0, border_value=0, interpolation='bilinear'):
    h, w = img.shape[:2]
    center = ((w - 1) * 0.5, (h - 1) * 0.5) if center is None else center
    m = cv2.getRotationMatrix2D(center, -angle, scale)
    return cv2.warpAffine(img, m, (w, h), flags=_CV2_INTERP[interpolation], borderValue=border_value)


def _border(v):
    return tuple(v) if isinstance(v, (list, tuple)) else (v, v, v)


def imshear(img, magnitude, direction='horizontal', border_value=0, interpolation='bilinear'):
    h, w = img.shape[:2]
    m = np.float32([[1, magnitude, 0], [0, 1, 0]]) if direction == 'horizontal' else np.float32([[1, 0, 0], [magnitude, 1, 0]])
    return cv2.warpAffine(img, m, (w, h), borderValue=_border(border_value)[:3], flags=_CV2_INTERP[interpolation])


def imtranslate(img, offset, direction='horizontal', border_value=0, interpolation='bilinear'):
    h, w = img.shape[:2]
    m = np.float32([[1, 0, offset], [0, 1, 0]]) if direction == 'horizontal' else np.float32([[1, 0, 0], [0, 1, offset]])
    return cv2.warpAffine(img, m, (w, h), borderValue=_border(border_value)[:3], flags=_CV2_INTERP[interpolation])


# ------------------------------------------------------------------------------------------ loading
@register
class LoadImageFromFile:
    def __init__(self, to_float32=False, color_type='color', **kwargs):
        self.to_float32, self.color_type = to_float32, color_type

    def __call__(self, results):
        info = results['img_info']
        prefix = results.get('img_prefix')
        filename = os.path.join(prefix, info['filename']) if prefix is not None else info['filename']
        img = cv2.imread(filename, cv2.IMREAD_COLOR if self.color_type == 'color' else cv2.IMREAD_UNCHANGED)
        if img is None:
            raise FileNotFoundError(filename)
        if self.to_float32:
            img = img.astype(np.float32)
        results.update(filename=filename, ori_filename=info['filename'], img=img, img_shape=img.shape, ori_shape=img.shape,
                       pad_shape=img.shape, scale_factor=1.0, img_fields=['img'])
        c = 1 if img.ndim < 3 else img.shape[2]
        results['img_norm_cfg'] = dict(mean=np.zeros(c, dtype=np.float32), std=np.ones(c, dtype=np.float32), to_rgb=False)
        return results


@register
class LoadAnnotations:
    """det: with_bbox / with_label (mmdet); seg: reduce_zero_label (mmseg)."""

    def __init__(self, with_bbox=False, with_label=True, with_mask=False, with_seg=False, reduce_zero_label=False,
                 task=None, **kwargs):
        assert not with_mask, 'instance masks are not on any reference pipeline'
        self.with_bbox, self.with_label, self.reduce_zero_label = with_bbox, with_label, reduce_zero_label
        self.seg = task == 'seg' or with_seg

    def __call__(self, results):
        if self.seg:
            from PIL import Image
            prefix = results.get('seg_prefix')
            name = results['ann_info']['seg_map']
            filename = os.path.join(prefix, name) if prefix is not None else name
            seg = np.array(Image.open(filename)).squeeze().astype(np.uint8)
            if results.get('label_map'):
                old = seg.copy()
                for a, b in results['label_map'].items():
                    seg[old == a] = b
            if self.reduce_zero_label:
                seg[seg == 0] = 255
                seg = seg - 1
                seg[seg == 254] = 255
            results['gt_semantic_seg'] = seg
            results.setdefault('seg_fields', []).append('gt_semantic_seg')
            return results
        ann = results['ann_info']
        if self.with_bbox:
            results['gt_bboxes'] = ann['bboxes'].copy()
            results.setdefault('bbox_fields', []).append('gt_bboxes')
            if ann.get('bboxes_ignore') is not None:
                results['gt_bboxes_ignore'] = ann['bboxes_ignore'].copy()
                results['bbox_fields'].append('gt_bboxes_ignore')
        if self.with_label:
            results['gt_labels'] = ann['labels'].copy()
        return results


# ------------------------------------------------------------------------------------------ geometry
@register
class Resize:
    """det (mmdet): img_scale + keep_ratio, boxes scaled and clipped.  seg (mmseg): img_scale + ratio_range
    (one uniform draw), label map rescaled with nearest.  cls (mmcls): size (h, w) or int, pillow/cv2 backend."""

    def __init__(self, img_scale=None, multiscale_mode='range', ratio_range=None, keep_ratio=True, bbox_clip_border=True,
                 backend='cv2', interpolation='bilinear', size=None, adaptive_side='short', override=False, task=None,
                 min_size=None, **kwargs):
        self.task = task
        if task == 'cls' or size is not None:
            self.task = 'cls'
            self.size = (size, size) if isinstance(size, int) else tuple(size)
            self.backend, self.interpolation = backend, interpolation
            return
        if img_scale is None:
            self.img_scale = None
        else:
            self.img_scale = img_scale if isinstance(img_scale, list) else [tuple(img_scale)]
            self.img_scale = [tuple(s) for s in self.img_scale]
        if ratio_range is not None:
            assert self.img_scale is None or len(self.img_scale) == 1
        self.multiscale_mode, self.ratio_range, self.keep_ratio = multiscale_mode, ratio_range, keep_ratio
        self.bbox_clip_border, self.backend, self.interpolation = bbox_clip_border, backend, interpolation

    def _random_scale(self, results):
        if self.ratio_range is not None:
            lo, hi = self.ratio_range
            base = self.img_scale[0] if self.img_scale is not None else results['img'].shape[:2][::-1]
            ratio = np.random.random_sample() * (hi - lo) + lo
            scale, idx = (int(base[0] * ratio), int(base[1] * ratio)), None
        elif len(self.img_scale) == 1:
            scale, idx = self.img_scale[0], 0
        elif self.multiscale_mode == 'range':
            longs = [max(s) for s in self.img_scale]
            shorts = [min(s) for s in self.img_scale]
            scale = (np.random.randint(min(longs), max(longs) + 1), np.random.randint(min(shorts), max(shorts) + 1))
            idx = None
        else:
            idx = np.random.randint(len(self.img_scale))
            scale = self.img_scale[idx]
        results['scale'], results['scale_idx'] = scale, idx

    def __call__(self, results):
        if self.task == 'cls':
            for key in results.get('img_fields', ['img']):
                img = results[key]
                results[key] = imresize(img, (self.size[1], self.size[0]), self.interpolation, self.backend)
                results['img_shape'] = results[key].shape
            return results
        if 'scale' not in results:
            if 'scale_factor' in results and not isinstance(results['scale_factor'], float) and self.img_scale is None:
                h, w = results['img'].shape[:2]
                f = results['scale_factor']
                results['scale'] = (int(w * f), int(h * f))
            else:
                self._random_scale(results)
        elif self.img_scale is None and self.ratio_range is None and isinstance(results['scale'], float):
            h, w = results['img'].shape[:2]
            results['scale'] = (int(w * results['scale']), int(h * results['scale']))
        img = results['img']
        h, w = img.shape[:2]
        if self.keep_ratio:
            img = imrescale(img, results['scale'], self.interpolation, self.backend)
            nh, nw = img.shape[:2]
            w_scale, h_scale = nw / w, nh / h
        else:
            img = imresize(img, tuple(results['scale']), self.interpolation, self.backend)
            nh, nw = img.shape[:2]
            w_scale, h_scale = nw / w, nh / h
        sf = np.array([w_scale, h_scale, w_scale, h_scale], dtype=np.float32)
        results.update(img=img, img_shape=img.shape, pad_shape=img.shape, scale_factor=sf, keep_ratio=self.keep_ratio)
        for key in results.get('bbox_fields', []):
            b = results[key] * sf
            if self.bbox_clip_border:
                b[:, 0::2] = np.clip(b[:, 0::2], 0, img.shape[1])
                b[:, 1::2] = np.clip(b[:, 1::2], 0, img.shape[0])
            results[key] = b
        for key in results.get('seg_fields', []):
            if self.keep_ratio:
                results[key] = imrescale(results[key], results['scale'], 'nearest', self.backend)
            else:
                results[key] = imresize(results[key], tuple(results['scale']), 'nearest', self.backend)
        return results


@register
class RandomFlip:
    def __init__(self, flip_ratio=None, direction='horizontal', prob=None, flip_prob=None, task=None, **kwargs):
        p = flip_ratio if flip_ratio is not None else prob if prob is not None else flip_prob
        self.prob, self.direction, self.task = p, direction, task

    def __call__(self, results):
        if 'flip' not in results:
            if self.task == 'det':
                # mmdet draws with np.random.choice over [direction, None]
                p = self.prob
                if p is None:
                    choice = None
                else:
                    choice = np.random.choice([self.direction, None], p=[p, 1 - p])
                results['flip'] = choice is not None
                results.setdefault('flip_direction', choice)
            else:
                results['flip'] = bool(self.prob is not None and np.random.rand() < self.prob)
                results.setdefault('flip_direction', self.direction)
        else:
            results.setdefault('flip_direction', self.direction)
        if not results['flip']:
            return results
        d = results['flip_direction']
        for key in results.get('img_fields', ['img']):
            results[key] = imflip(results[key], d)
        h, w = results['img'].shape[:2]
        for key in results.get('bbox_fields', []):
            b = results[key]
            f = b.copy()
            if d == 'horizontal':
                f[..., 0::4], f[..., 2::4] = w - b[..., 2::4], w - b[..., 0::4]
            elif d == 'vertical':
                f[..., 1::4], f[..., 3::4] = h - b[..., 3::4], h - b[..., 1::4]
            else:
                f[..., 0::4], f[..., 1::4] = w - b[..., 2::4], h - b[..., 3::4]
                f[..., 2::4], f[..., 3::4] = w - b[..., 0::4], h - b[..., 1::4]
            results[key] = f
        for key in results.get('seg_fields', []):
            results[key] = imflip(results[key], d).copy()
        return results


@register
class RandomCrop:
    """mmseg RandomCrop(crop_size=(h, w), cat_max_ratio, ignore_index=255)."""

    def __init__(self, crop_size, cat_max_ratio=1., ignore_index=255, **kwargs):
        self.crop_size, self.cat_max_ratio, self.ignore_index = tuple(crop_size), cat_max_ratio, ignore_index

    def _bbox(self, img):
        mh = max(img.shape[0] - self.crop_size[0], 0)
        mw = max(img.shape[1] - self.crop_size[1], 0)
        oh, ow = np.random.randint(0, mh + 1), np.random.randint(0, mw + 1)
        return oh, oh + self.crop_size[0], ow, ow + self.crop_size[1]

    def __call__(self, results):
        img = results['img']
        y1, y2, x1, x2 = self._bbox(img)
        if self.cat_max_ratio < 1.:
            for _ in range(10):
                tmp = results['gt_semantic_seg'][y1:y2, x1:x2, ...]
                labels, cnt = np.unique(tmp, return_counts=True)
                cnt = cnt[labels != self.ignore_index]
                if len(cnt) > 1 and np.max(cnt) / np.sum(cnt) < self.cat_max_ratio:
                    break
                y1, y2, x1, x2 = self._bbox(img)
        img = img[y1:y2, x1:x2, ...]
        results.update(img=img, img_shape=img.shape)
        for key in results.get('seg_fields', []):
            results[key] = results[key][y1:y2, x1:x2, ...]
        return results


@register
class Pad:
    def __init__(self, size=None, size_divisor=None, pad_to_square=False, pad_val=0, seg_pad_val=255, **kwargs):
        self.size, self.size_divisor, self.pad_val, self.seg_pad_val = size, size_divisor, pad_val, seg_pad_val
        if isinstance(pad_val, dict):
            self.pad_val, self.seg_pad_val = pad_val.get('img', 0), pad_val.get('seg', 255)
        assert size is not None or size_divisor is not None
        assert size is None or size_divisor is None

    def __call__(self, results):
        for key in results.get('img_fields', ['img']):
            if self.size is not None:
                results[key] = impad(results[key], tuple(self.size), self.pad_val)
            else:
                results[key] = impad_to_multiple(results[key], self.size_divisor, self.pad_val)
        results.update(pad_shape=results['img'].shape, pad_fixed_size=self.size, pad_size_divisor=self.size_divisor)
        for key in results.get('seg_fields', []):
            results[key] = impad(results[key], results['pad_shape'][:2], self.seg_pad_val)
        return results


@register
class RandomResizedCrop:
    """mmcls RandomResizedCrop (torchvision semantics): up to max_attempts draws of (area, log-uniform
    aspect), centre crop fallback, then resize to `size`."""

    def __init__(self, size, scale=(0.08, 1.0), ratio=(3. / 4., 4. / 3.), max_attempts=10, efficientnet_style=False,
                 interpolation='bilinear', backend='cv2', **kwargs):
        assert not efficientnet_style
        self.size = (size, size) if isinstance(size, int) else tuple(size)
        self.scale, self.ratio, self.max_attempts = scale, ratio, max_attempts
        self.interpolation, self.backend = interpolation, backend

    def _params(self, img):
        h, w = img.shape[:2]
        area = h * w
        for _ in range(self.max_attempts):
            target = random.uniform(*self.scale) * area
            aspect = math.exp(random.uniform(math.log(self.ratio[0]), math.log(self.ratio[1])))
            tw, th = int(round(math.sqrt(target * aspect))), int(round(math.sqrt(target / aspect)))
            if 0 < tw <= w and 0 < th <= h:
                y = random.randint(0, h - th)
                x = random.randint(0, w - tw)
                return y, x, th, tw
        in_ratio = float(w) / float(h)
        if in_ratio < min(self.ratio):
            tw = w
            th = int(round(tw / min(self.ratio)))
        elif in_ratio > max(self.ratio):
            th = h
            tw = int(round(th * max(self.ratio)))
        else:
            tw, th = w, h
        return (h - th) // 2, (w - tw) // 2, th, tw

    def __call__(self, results):
        for key in results.get('img_fields', ['img']):
            img = results[key]
            y, x, th, tw = self._params(img)
            img = img[y:y + th, x:x + tw]
            results[key] = imresize(img, (self.size[1], self.size[0]), self.interpolation, self.backend)
        results['img_shape'] = results['img'].shape
        return results


# ------------------------------------------------------------------------------------------ photometric
@register
class PhotoMetricDistortion:
    """mmseg PhotoMetricDistortion: brightness, then contrast first or last (coin), saturation, hue."""

    def __init__(self, brightness_delta=32, contrast_range=(0.5, 1.5), saturation_range=(0.5, 1.5), hue_delta=18):
        self.bd, (self.cl, self.cu), (self.sl, self.su), self.hd = brightness_delta, contrast_range, saturation_range, hue_delta

    @staticmethod
    def _convert(img, alpha=1, beta=0):
        return np.clip(img.astype(np.float32) * alpha + beta, 0, 255).astype(np.uint8)

    def __call__(self, results):
        img = results['img']
        if np.random.randint(2):
            img = self._convert(img, beta=np.random.uniform(-self.bd, self.bd))
        mode = np.random.randint(2)
        if mode == 1 and np.random.randint(2):
            img = self._convert(img, alpha=np.random.uniform(self.cl, self.cu))
        if np.random.randint(2):
            img = cv2.cvtColor(img, cv2.COLOR_BGR2HSV)
            img[:, :, 1] = self._convert(img[:, :, 1], alpha=np.random.uniform(self.sl, self.su))
            img = cv2.cvtColor(img, cv2.COLOR_HSV2BGR)
        if np.random.randint(2):
            img = cv2.cvtColor(img, cv2.COLOR_BGR2HSV)
            img[:, :, 0] = (img[:, :, 0].astype(int) + np.random.randint(-self.hd, self.hd)) % 180
            img = cv2.cvtColor(img, cv2.COLOR_HSV2BGR)
        if mode == 0 and np.random.randint(2):
            img = self._convert(img, alpha=np.random.uniform(self.cl, self.cu))
        results['img'] = img
        return results


@register
class RandomErasing:
    """mmcls RandomErasing (mode 'const' | 'rand')."""

    def __init__(self, erase_prob=0.5, min_area_ratio=0.02, max_area_ratio=0.4, aspect_range=(3 / 10, 10 / 3), mode='const',
                 fill_color=(128, 128, 128), fill_std=None):
        self.p, self.lo, self.hi, self.mode = erase_prob, min_area_ratio, max_area_ratio, mode
        self.aspect = (aspect_range, 1 / aspect_range) if isinstance(aspect_range, float) else tuple(aspect_range)
        self.aspect = (min(self.aspect), max(self.aspect))
        self.fill_color, self.fill_std = fill_color, fill_std

    def __call__(self, results):
        for key in results.get('img_fields', ['img']):
            if np.random.rand() > self.p:
                continue
            img = results[key]
            H, W = img.shape[:2]
            la = np.log(np.array(self.aspect, dtype=np.float32))
            ar = np.exp(np.random.uniform(*la))
            area = H * W * np.random.uniform(self.lo, self.hi)
            h = min(int(round(np.sqrt(area * ar))), H)
            w = min(int(round(np.sqrt(area / ar))), W)
            top = np.random.randint(0, H - h) if H > h else 0
            left = np.random.randint(0, W - w) if W > w else 0
            if self.mode == 'const':
                patch = np.empty((h, w, 3), dtype=np.uint8)
                patch[:, :] = np.array(self.fill_color, dtype=np.uint8)
            elif self.fill_std is None:
                patch = np.random.uniform(0, 256, (h, w, 3)).astype(np.uint8)
            else:
                patch = np.random.normal(self.fill_color, self.fill_std, (h, w, 3))
                patch = np.clip(patch.astype(np.int32), 0, 255).astype(np.uint8)
            img = np.ascontiguousarray(img)
            img[top:top + h, left:left + w] = patch
            results[key] = img
        return results


def _negate(v, prob):
    return -v if np.random.rand() < prob else v


class _Aug:
    """one RandAugment policy: applied with probability `prob`; signed magnitudes flip with random_negative_prob."""
    signed = False

    def __init__(self, prob=0.5, random_negative_prob=0.5, pad_val=128, interpolation='nearest', direction='horizontal', **mag):
        self.prob, self.neg, self.pad_val, self.interpolation, self.direction = prob, random_negative_prob, pad_val, interpolation, direction
        self.pad_val = tuple(pad_val) if isinstance(pad_val, (list, tuple)) else (pad_val,) * 3
        self.mag = next(iter(mag.values())) if mag else None

    def __call__(self, results):
        if np.random.rand() > self.prob:
            return results
        m = _negate(self.mag, self.neg) if self.signed else self.mag
        for key in results.get('img_fields', ['img']):
            results[key] = self.apply(results[key], m).astype(results[key].dtype)
        return results


def _aug(name, fn, signed=False):
    cls = type(name, (_Aug,), dict(apply=lambda self, img, m: fn(self, img, m), signed=signed))
    return register(cls)


_aug('AutoContrast', lambda s, im, m: auto_contrast(im))
_aug('Equalize', lambda s, im, m: imequalize(im))
_aug('Invert', lambda s, im, m: np.full_like(im, 255) - im)
_aug('Rotate', lambda s, im, m: imrotate(im, m, border_value=s.pad_val, interpolation=s.interpolation), signed=True)
_aug('Posterize', lambda s, im, m: np.left_shift(np.right_shift(im, 8 - int(math.ceil(m))), 8 - int(math.ceil(m))))
_aug('Solarize', lambda s, im, m: np.where(im < m, im, 255 - im))
_aug('SolarizeAdd', lambda s, im, m: np.where(im < 128, np.minimum(im.astype(np.int32) + int(m), 255), im))
_aug('ColorTransform', lambda s, im, m: adjust_color(im, 1 + m), signed=True)
_aug('Contrast', lambda s, im, m: adjust_contrast(im, 1 + m), signed=True)
_aug('Brightness', lambda s, im, m: adjust_brightness(im, 1 + m), signed=True)
_aug('Sharpness', lambda s, im, m: adjust_sharpness(im, 1 + m), signed=True)
_aug('Shear', lambda s, im, m: imshear(im, m, s.direction, s.pad_val, s.interpolation), signed=True)
_aug('Translate', lambda s, im, m: imtranslate(im, m * (im.shape[1] if s.direction == 'horizontal' else im.shape[0]),
                                               s.direction, s.pad_val, s.interpolation), signed=True)
_GEOMETRIC = ('Rotate', 'Shear', 'Translate')


@register
class RandAugment:
    """mmcls RandAugment: per sample `num_policies` policies drawn with replacement; each policy's magnitude is
    gauss(magnitude_level, magnitude_std) clipped to [0, total_level], mapped linearly onto its magnitude_range."""

    def __init__(self, policies, num_policies, magnitude_level, magnitude_std=0., total_level=30, hparams=None):
        assert num_policies > 0 and len(policies) > 0
        self.policies = [dict(p) for p in policies]
        self.num_policies, self.level, self.std, self.total = num_policies, magnitude_level, magnitude_std, total_level
        self.hparams = dict(hparams or {})

    def _concrete(self, policy):
        p = dict(policy)
        key = p.pop('magnitude_key', None)
        rng = p.pop('magnitude_range', None)
        if key is not None:
            m = self.level
            if self.std == 'inf':
                m = random.uniform(0, m)
            elif self.std > 0:
                m = random.gauss(m, self.std)
                m = min(self.total, max(0, m))
            p[key] = (m / self.total) * float(rng[1] - rng[0]) + rng[0]
        if p['type'] in _GEOMETRIC:
            for k, v in self.hparams.items():
                p.setdefault(k, v)
        return build_transform(p)

    def __call__(self, results):
        if self.num_policies == 0:
            return results
        for policy in random.choices(self.policies, k=self.num_policies):
            results = self._concrete(policy)(results)
        return results


@register
class Normalize:
    def __init__(self, mean, std, to_rgb=True, defer=False):
        self.mean, self.std = np.array(mean, dtype=np.float32), np.array(std, dtype=np.float32)
        self.to_rgb, self.defer = to_rgb, defer

    def __call__(self, results):
        results['img_norm_cfg'] = dict(mean=self.mean, std=self.std, to_rgb=self.to_rgb)
        if self.defer:
            results['norm_deferred'] = True
            return results
        for key in results.get('img_fields', ['img']):
            results[key] = imnormalize(results[key], self.mean, self.std, self.to_rgb)
        return results


# ------------------------------------------------------------------------------------------ formatting
def _img_tensor(img):
    if img.ndim < 3:
        img = img[..., None]
    return torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1)))


@register
class ImageToTensor:
    def __init__(self, keys):
        self.keys = keys

    def __call__(self, results):
        for k in self.keys:
            results[k] = _img_tensor(results[k])
        return results


@register
class ToTensor:
    def __init__(self, keys):
        self.keys = keys

    def __call__(self, results):
        for k in self.keys:
            v = results[k]
            results[k] = v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v))
        return results


@register
class DefaultFormatBundle:
    """img -> CHW tensor, boxes/labels -> tensors, gt_semantic_seg -> (1, H, W) int64.  (No DataContainer:
    the collate function knows which keys stack and which stay per-image lists.)"""

    def __init__(self, img_to_float=True, pad_val=None, **kwargs):
        self.img_to_float = img_to_float

    def __call__(self, results):
        if 'img' in results:
            img = results['img']
            if self.img_to_float and img.dtype == np.uint8 and not results.get('norm_deferred'):
                img = img.astype(np.float32)
            results.setdefault('pad_shape', img.shape)
            results.setdefault('scale_factor', 1.0)
            results['img'] = _img_tensor(img)
        for k in ('gt_bboxes', 'gt_bboxes_ignore', 'gt_labels'):
            if k in results:
                results[k] = torch.as_tensor(np.ascontiguousarray(results[k]))
        if 'gt_semantic_seg' in results:
            results['gt_semantic_seg'] = torch.from_numpy(np.ascontiguousarray(results['gt_semantic_seg'][None, ...]).astype(np.int64))
        return results


_META_KEYS = ('filename', 'ori_filename', 'ori_shape', 'img_shape', 'pad_shape', 'scale_factor', 'flip', 'flip_direction',
              'img_norm_cfg')


@register
class Collect:
    def __init__(self, keys, meta_keys=_META_KEYS):
        self.keys, self.meta_keys = keys, meta_keys

    def __call__(self, results):
        data = dict(img_metas={k: results[k] for k in self.meta_keys if k in results})
        if results.get('norm_deferred'):
            data['img_metas']['norm_deferred'] = True
        for k in self.keys:
            data[k] = results[k]
        return data


@register
class MultiScaleFlipAug:
    """test-time wrapper; every reference config uses ONE scale and flip=False, so this yields lists of length 1."""

    def __init__(self, transforms, img_scale=None, img_ratios=None, scale_factor=None, flip=False, flip_direction='horizontal',
                 task=None):
        self.transforms = Compose(transforms, task)
        if img_scale is not None and img_ratios is None and scale_factor is None:
            self.scales, self.key = (img_scale if isinstance(img_scale, list) else [img_scale]), 'scale'
        elif img_scale is None:
            r = img_ratios if img_ratios is not None else scale_factor
            self.scales, self.key = [float(x) for x in (r if isinstance(r, (list, tuple)) else [r])], 'scale_factor' if task == 'det' else 'scale'
        else:
            r = img_ratios if isinstance(img_ratios, (list, tuple)) else [img_ratios]
            self.scales, self.key = [(int(img_scale[0] * x), int(img_scale[1] * x)) for x in r], 'scale'
        self.flip = flip
        self.dirs = flip_direction if isinstance(flip_direction, list) else [flip_direction]

    def __call__(self, results):
        outs = []
        flips = [(False, None)] + ([(True, d) for d in self.dirs] if self.flip else [])
        for s in self.scales:
            for f, d in flips:
                r = dict(results)
                r[self.key] = tuple(s) if isinstance(s, (list, tuple)) else s
                r['flip'], r['flip_direction'] = f, d
                outs.append(self.transforms(r))
        return {k: [o[k] for o in outs] for k in outs[0]}
