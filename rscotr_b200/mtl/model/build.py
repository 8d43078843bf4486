"""Thin builders with the reference's signatures (mtl/model/build.py:7-87): a module is
passed through, a cfg dict is resolved in this package's registry whatever `base_mm`
(mmcls / mmdet / mmseg) the reference would have dispatched to."""
import torch.nn as nn

from ...config import MODELS, build_from_cfg


def _build(cfg, base_mm, return_init_requirement):
    if isinstance(cfg, nn.Module):
        return (cfg, False) if return_init_requirement else cfg
    m = build_from_cfg(dict(cfg), MODELS)
    return (m, True) if return_init_requirement else m


def build_backbone(cfg, base_mm='mmdet', return_init_requirement=False):
    return _build(cfg, base_mm, return_init_requirement)


def build_neck(cfg, base_mm='mmdet', return_init_requirement=False):
    return _build(cfg, base_mm, return_init_requirement)


def build_head(cfg, base_mm='mmdet', return_init_requirement=False):
    return _build(cfg, base_mm, return_init_requirement)


def build_transformer(cfg, base_mm='mmdet', return_init_requirement=False):
    return _build(cfg, base_mm, return_init_requirement)


def build_transformer_layer_sequence(cfg, base_mm='mmdet', return_init_requirement=False):
    return _build(cfg, base_mm, return_init_requirement)
