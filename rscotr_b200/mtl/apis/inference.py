"""Single-image inference (the useful core of the reference's tools/inference_one_img.py:206-290): push one image file
through a dataset's TEST pipeline, batch it the way the loaders do, run the model for the dataset's task."""
import torch

from ..data.loader import collate
from ..engine.step import _to_device


def inference_one_img(model, dataset, filename, device=None, **kwargs):
    """-> the model's result for this image: class scores (cls), per-class (n, 5) boxes (det), label map (seg).
    `dataset` supplies the pipeline, the task and (for the caller) CLASSES; `filename` is an absolute path or relative to
    the dataset's image prefix."""
    import os
    task = dataset.task
    prefix = None if os.path.isabs(filename) else getattr(dataset, 'img_prefix', None) or getattr(dataset, 'img_dir', None) \
        or getattr(dataset, 'data_prefix', None)
    sample = dict(img_info=dict(filename=filename), img_prefix=prefix, bbox_fields=[], seg_fields=[])
    data = collate([dataset.pipeline(sample)])
    for k in ('gt_label', 'gt_bboxes', 'gt_labels', 'gt_semantic_seg'):
        data.pop(k, None)
    device = device or next(model.parameters()).device
    data = _to_device(data, device)
    img, metas = data.pop('img'), data.pop('img_metas')
    if not isinstance(img, list):
        img, metas = [img], [metas]
    if task == 'det':
        kwargs.setdefault('rescale', True)
    was_training = model.training
    model.eval()
    try:
        with torch.no_grad():
            result = model(return_loss=False, task=task, img=img, img_metas=metas, **kwargs)
    finally:
        model.train(was_training)
    return result[0] if isinstance(result, (list, tuple)) and len(result) == 1 else result
