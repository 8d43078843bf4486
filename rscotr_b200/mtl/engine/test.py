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
    # mmseg.apis.single_gpu_test: the dataset indices of every batch come from the loader's batch sampler (a sharded
    # validation loader does not visit 0, 1, 2, ...); loaders without one (synthetic) are sequential
    sampler = getattr(loader, 'batch_sampler', None)
    index_iter = iter(sampler) if sampler is not None and not getattr(loader, 'infinite', False) else None
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
        indices = list(next(index_iter)) if index_iter is not None else list(range(n, n + len(result)))
        if reduce_seg:
            result = dataset.pre_eval(result, indices[:len(result)])
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
    """mm{cls,det,seg}.apis.multi_gpu_test: every rank runs its shard of each validation loader (rank-strided, not
    shuffled, padded to equal length), the per-sample results are gathered and re-interleaved into dataset order on rank 0
    (the other ranks get None), as `collect_results_cpu/gpu` do."""
    import torch.distributed as dist
    results = single_gpu_test(model, data_loaders, kwargs_dict=kwargs_dict)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return results
    world, rank = dist.get_world_size(), dist.get_rank()
    out = {}
    for name, part in results.items():
        if not hasattr(data_loaders[name].dataset, '__len__'):      # synthetic loaders are per-rank streams, not shards
            out[name] = part
            continue
        parts = [None] * world
        dist.all_gather_object(parts, part)
        if rank == 0:
            size = len(data_loaders[name].dataset)
            out[name] = [r for group in zip(*parts) for r in group][:size]
        else:
            out[name] = None
    return out
