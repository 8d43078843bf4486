"""In-switch gradient exchange (csrc/nvls.cu, row 8e): rsc_nvls_allreduce_mean on a symmetric-memory buffer with an NVLS
multicast mapping against ncclAllReduce(AVG) on the same data, on an offset sub-range.  Needs >= 2 GPUs of one NVSwitch box
(skipped on a single-GPU box); `tools/nvls_probe.py` is the timing companion."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    import torch.distributed._symmetric_memory as symm
    from rscotr_b200 import _lib
    total, lo, hi = 3 << 20, 64 * 1000, 64 * 40000
    buf = symm.empty(total, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(buf, dist.group.WORLD)
    res = dict(multicast=bool(hdl.multicast_ptr))
    if res['multicast']:
        g = torch.Generator(device=dev).manual_seed(rank)
        src = torch.randn(total, device=dev, generator=g)
        buf.copy_(src)
        ref = src.clone()
        hdl.barrier(channel=0)
        _lib.call('rsc_nvls_allreduce_mean', int(hdl.multicast_ptr), lo, hi, rank, world, 1.0 / world, 16,
                  torch.cuda.current_stream().cuda_stream)
        hdl.barrier(channel=0)
        dist.all_reduce(ref[lo:hi], op=dist.ReduceOp.AVG)
        torch.cuda.synchronize()
        res.update(err=float((buf[lo:hi] - ref[lo:hi]).abs().max() / ref[lo:hi].abs().max()),
                   untouched=bool(torch.equal(buf[:lo], src[:lo]) and torch.equal(buf[hi:], src[hi:])))
    torch.save(res, os.path.join(out_dir, 'nvls_rank%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_nvls_allreduce_mean_matches_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs of one NVSwitch box')
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        res = torch.load(tmp_path / ('nvls_rank%d.pt' % r))
        if not res['multicast']:
            pytest.skip('no NVLS multicast mapping on this box')
        assert res['err'] < 1e-6 and res['untouched'], res
