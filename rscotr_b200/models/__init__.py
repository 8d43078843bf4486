"""Importing this package registers every class the reference configs name
(the reference does it with custom_imports=dict(imports='models.multi'))."""
from . import bricks, swin, cls_head, det_head, seg_head, mtl, uper_head, classifier  # noqa: F401
from .mtl import MTL  # noqa: F401
from .classifier import ImageClassifier  # noqa: F401
