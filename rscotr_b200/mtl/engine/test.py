"""single_gpu_test / multi_gpu_test with the reference's signatures (mtl/engine/test.py:24-53):
per dataset, run model(return_loss=False, **data) over the loader and collect results."""
import torch

from .step import _to_device


def _run(model, loader, task, show=False, out_dir=None, **kwargs):
    results = []
    device = next(model.parameters()).device
    for data in loader:
        data = _to_device(dict(data), device)
        data.pop('dataset_name', None)
        data.pop('task', None)
        for k in ('gt_label', 'gt_bboxes', 'gt_labels', 'gt_semantic_seg'):
            data.pop(k, None)
        with torch.no_grad():
            result = model(return_loss=False, task=task, img=[data.pop('img')], img_metas=[data.pop('img_metas')],
                           **data)
        results.extend(result if isinstance(result, list) else [result])
    return results


def single_gpu_test(model, data_loaders, show=False, out_dir=None, kwargs_dict=None):
    kwargs_dict = kwargs_dict or {}
    model.eval()
    results = dict()
    for name, loader in data_loaders.items():
        task = loader.dataset.task
        model.CLASSES = getattr(loader.dataset, 'CLASSES', None)
        results[name] = _run(model, loader, task, show, out_dir, **kwargs_dict.get(task, {}))
    return results


def multi_gpu_test(model, data_loaders, tmpdir=None, gpu_collect=False, kwargs_dict=None):
    """The reference raises NotImplementedError for distributed validation
    (mtl/apis/train.py:100-101); each rank evaluates its own shard here."""
    return single_gpu_test(model, data_loaders, kwargs_dict=kwargs_dict)
