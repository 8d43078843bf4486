"""Single-task image classifier of the reference's configs/cls/*.py (mmcls ImageClassifier: backbone -> neck -> head,
batch augments from train_cfg), with the step engine's model interface (BASELINE configs[0])."""
from ..config import MODELS, build_from_cfg
from .bricks import apply_init_cfg
from .cls_head import Augments
from .mtl import SingleTaskModel, normalize_on_device


@MODELS.register_module()
class ImageClassifier(SingleTaskModel):
    """mmcls ImageClassifier(backbone, neck, head, train_cfg.augments) -- the reference's single-task classification
    configs (configs/cls/*.py; BASELINE configs[0]) -- with the step engine's model interface."""
    default_task = 'cls'

    def __init__(self, backbone, neck=None, head=None, pretrained=None, train_cfg=None, init_cfg=None):
        super().__init__()
        self.backbone = build_from_cfg(backbone, MODELS)
        self.neck = build_from_cfg(neck, MODELS) if neck is not None else None
        self.head = build_from_cfg(head, MODELS) if head is not None else None
        self.augments = None
        aug = (train_cfg or {}).get('augments', None)
        if aug is not None:
            self.augments = Augments(aug)
        apply_init_cfg(self, init_cfg)      # (mmcv applies the model-level init_cfg to every matching layer below)

    def init_weights(self):
        bb = self.backbone
        if isinstance(getattr(bb, 'init_cfg', None), dict) and bb.init_cfg.get('type') == 'Pretrained':
            bb.init_weights()

    def extract_feat(self, img):
        x = self.backbone(img)
        x = tuple(x) if isinstance(x, (list, tuple)) else (x,)
        return self.neck(x) if self.neck is not None else x

    def forward_train(self, img, gt_label, **kwargs):
        if self.augments is not None:
            img, gt_label = self.augments(img, gt_label)
        return self.head.forward_train(self.extract_feat(img), gt_label)

    def simple_test(self, img, img_metas=None, **kwargs):
        return self.head.simple_test(self.extract_feat(img), **kwargs)

    def forward(self, img, img_metas=None, return_loss=True, task=None, dataset_name=None, **kwargs):
        if isinstance(img, list):
            img, img_metas = img[0], (img_metas[0] if img_metas else None)
        if img_metas:
            img = normalize_on_device(img, img_metas)
        if return_loss:
            return self.forward_train(img, **kwargs)
        return self.simple_test(img, img_metas, **kwargs)
