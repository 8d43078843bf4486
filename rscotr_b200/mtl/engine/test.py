"""single_gpu_test / multi_gpu_test with the reference's signatures (mtl/engine/test.py:24-53):
per dataset, run model(return_loss=False, **data) over the loader and collect results."""
import torch

from .step import _to_device


def _run(model, loader, task, show=False, out_dir=None, pre_eval=None, **kwargs):
    """mm{cls,det,seg}.apis.single_gpu_test: model(return_loss=False, **data) over the loader.  Seg results are
    reduced to their per-image (intersect, union, pred, label) areas on the fly when the dataset can do it
    (same metrics as keeping every label map, a fraction of the memory)."""
    results = []
    device = next(model.parameters()).device
    dataset = loader.dataset
    reduce_seg = task == 'seg' and hasattr(dataset, 'pre_eval') and pre_eval is not False
    n = 0
    for data in loader:
        data = _to_device(dict(data), device)
        data.pop('dataset_name', None)
        data.pop('task', None)
        for k in ('gt_label', 'gt_bboxes', 'gt_labels', 'gt_semantic_seg'):
            data.pop(k, None)
        img, metas = data.pop('img'), data.pop('img_metas')
        if not isinstance(img, list):                       # (no MultiScaleFlipAug wrapper: cls, synthetic batches)
            img, metas = [img], [metas]
        if task == 'det':
            data.setdefault('rescale', True)            # (mmdet.apis.single_gpu_test)
        with torch.no_grad():
            result = model(return_loss=False, task=task, img=img, img_metas=metas, **data)
        result = list(result) if isinstance(result, (list, tuple)) else list(result.unbind(0)) if torch.is_tensor(result) else [result]
        if reduce_seg:
            result = dataset.pre_eval(result, list(range(n, n + len(result))))
        n += len(result)
        results.extend(result)
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
