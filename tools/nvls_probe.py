"""Probe (torchrun, >= 2 GPUs of one NVSwitch box): is a symmetric-memory allocation + NVLS multicast available for the
flat gradient buffer, and how does an in-switch all-reduce (multimem.ld_reduce / multimem.st) compare with ncclAllReduce on
the gradient ranges of the three tasks?   torchrun --nproc-per-node N tools/nvls_probe.py"""
import os
import sys
import time

import torch
import torch.distributed as dist


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    import torch.distributed._symmetric_memory as symm
    total = 63 * (1 << 20)
    try:
        buf = symm.empty(total, dtype=torch.float32, device=dev)
        hdl = symm.rendezvous(buf, dist.group.WORLD)
        mc = hdl.multicast_ptr
        if rank == 0:
            print('symmetric memory ok: backend %s, multicast_ptr %#x, signal pad %d B' % (symm.get_backend(dev), mc, hdl.signal_pad_size),
                  flush=True)
    except Exception as e:
        print('rank %d: symmetric memory unavailable: %r' % (rank, e), flush=True)
        dist.destroy_process_group()
        return
    gname = dist.group.WORLD.group_name
    plain = torch.empty(total, dtype=torch.float32, device=dev)
    # correctness on an offset slice
    lo, hi = 1 << 20, (1 << 20) + 48 * (1 << 20)
    g = torch.Generator(device=dev).manual_seed(rank)
    src = torch.randn(total, device=dev, generator=g)
    buf.copy_(src)
    plain.copy_(src)
    torch.ops.symm_mem.multimem_all_reduce_(buf[lo:hi], 'sum', gname)
    dist.all_reduce(plain[lo:hi])
    torch.cuda.synchronize()
    err = float((buf[lo:hi] - plain[lo:hi]).abs().max() / plain[lo:hi].abs().max())
    untouched = bool(torch.equal(buf[:lo], src[:lo]) and torch.equal(buf[hi:], src[hi:]))
    if rank == 0:
        print('slice all-reduce: max rel err vs NCCL %.2e, outside the slice untouched: %s' % (err, untouched), flush=True)
    for name, n in (('cls 27.6M', 27_600_000), ('det 48M', 48_000_000), ('stage3 14.2M', 14_200_000), ('stage2 10.7M', 10_700_000),
                    ('1.6M', 1_600_000)):
        n = n // 1024 * 1024
        t_nvls = timeit(lambda: torch.ops.symm_mem.multimem_all_reduce_(buf[:n], 'sum', gname))
        t_nccl = timeit(lambda: dist.all_reduce(plain[:n], op=dist.ReduceOp.AVG))
        t_two = timeit(lambda: torch.ops.symm_mem.two_shot_all_reduce_(buf[:n], 'sum', gname))
        if rank == 0:
            print('%-14s %6.1f MB: multimem %7.1f us (%.0f GB/s algbw) | two_shot %7.1f us | nccl %7.1f us (%.0f GB/s)' % (
                name, n * 4 / 1e6, t_nvls, n * 4 / t_nvls / 1e3, t_two, t_nccl, n * 4 / t_nccl / 1e3), flush=True)
    # ---- the repo's own kernel: barrier -> rsc_nvls_allreduce_mean -> barrier
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from rscotr_b200 import _lib

    def own(lo_, hi_, ctas):
        hdl.barrier(channel=0)
        _lib.call('rsc_nvls_allreduce_mean', mc, lo_, hi_, rank, world, 1.0 / world, ctas, torch.cuda.current_stream().cuda_stream)
        hdl.barrier(channel=0)
    buf.copy_(src)
    plain.copy_(src)
    own(lo, hi, 64)
    dist.all_reduce(plain[lo:hi], op=dist.ReduceOp.AVG)
    torch.cuda.synchronize()
    err = float((buf[lo:hi] - plain[lo:hi]).abs().max() / plain[lo:hi].abs().max())
    untouched = bool(torch.equal(buf[:lo], src[:lo]) and torch.equal(buf[hi:], src[hi:]))
    if rank == 0:
        print('own kernel, slice mean: max rel err vs NCCL AVG %.2e, outside the slice untouched: %s' % (err, untouched), flush=True)
    for name, n in (('cls 27.6M', 27_600_000), ('det 48M', 48_000_000), ('stage3 14.2M', 14_200_000), ('1.6M', 1_600_000)):
        n = n // 1024 * 1024
        res = []
        for ctas in (16, 32, 64, 128):
            res.append('%d CTAs %7.1f us' % (ctas, timeit(lambda: own(0, n, ctas))))
        t_nccl = timeit(lambda: dist.all_reduce(plain[:n], op=dist.ReduceOp.AVG))
        if rank == 0:
            print('%-14s %6.1f MB own: %s | nccl %7.1f us' % (name, n * 4 / 1e6, ' | '.join(res), t_nccl), flush=True)
    # CUDA-graph capture of the in-switch all-reduce on a side stream
    try:
        side = torch.cuda.Stream()
        g_ = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(g_):
            cur = torch.cuda.current_stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                torch.ops.symm_mem.multimem_all_reduce_(buf[:14_200_000 // 1024 * 1024], 'sum', gname)
            plain.mul_(1.0)
            cur.wait_stream(side)
        for _ in range(3):
            g_.replay()
        torch.cuda.synchronize()
        if rank == 0:
            print('graph capture + replay of multimem_all_reduce_ on a side stream: ok', flush=True)
    except Exception as e:
        print('rank %d: graph capture failed: %r' % (rank, e), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
