"""GPU parity of rsc_groupnorm_{fwd,bwd} (csrc/norm_ops.cu) through rscotr_b200.ops / the bricks modules: GroupNorm-32
(ChannelMapper, pixel decoder; rows a8 / a17) and training-mode BatchNorm2d (UPerHead / FCNHead; row a20), with and
without the fused ReLU, against torch's own nn.GroupNorm / nn.BatchNorm2d (+ F.relu) in fp32 -- the ops the reference
calls.  fp32: 1e-4 relative (north_star's 1e-3 bar); bf16: outputs rounded to bf16 (2^-9), gradients 1e-2."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _cl(x):
    return x.contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('relu', [False, True])
@pytest.mark.parametrize('B,C,H,W,G', [(2, 256, 100, 100, 32), (1, 256, 25, 25, 32), (3, 256, 13, 7, 32), (2, 128, 9, 11, 8),
                                       (2, 512, 20, 20, 32)])
def test_group_norm_vs_torch(B, C, H, W, G, relu, dtype):
    from rscotr_b200.models import bricks
    torch.manual_seed(0)
    x = torch.randn(B, C, H, W) * 2 + 0.7
    ref = nn.GroupNorm(G, C)
    with torch.no_grad():
        ref.weight.uniform_(0.5, 1.5), ref.bias.uniform_(-0.5, 0.5)
    own = bricks.GroupNorm(G, C).cuda()
    own.load_state_dict(ref.state_dict())
    xr = x.to(dtype).float().clone().requires_grad_()
    want = ref(xr)
    want = F.relu(want) if relu else want
    dy = torch.randn(want.shape)
    want.backward(dy.to(dtype).float())
    xc = _cl(x.detach().to(dtype).cuda()).requires_grad_()
    got = own(xc, relu=relu)
    assert got.shape == want.shape and got.dtype == dtype
    got.backward(_cl(dy.to(dtype).cuda()))
    tol_y, tol_g = (1e-4, 1e-4) if dtype == torch.float32 else (4e-3, 1e-2)
    assert rel(got, want) < tol_y, rel(got, want)
    assert rel(xc.grad, xr.grad) < tol_g, rel(xc.grad, xr.grad)
    assert rel(own.weight.grad, ref.weight.grad) < tol_g and rel(own.bias.grad, ref.bias.grad) < tol_g


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('relu', [False, True])
@pytest.mark.parametrize('B,C,H,W', [(2, 512, 32, 32), (4, 256, 16, 16), (1, 512, 7, 5), (8, 2048, 6, 6)])
def test_batch_norm_training_vs_torch(B, C, H, W, relu, dtype):
    from rscotr_b200.models import bricks
    torch.manual_seed(1)
    x = torch.randn(B, C, H, W) * 1.5 - 0.3
    ref = nn.BatchNorm2d(C)
    with torch.no_grad():
        ref.weight.uniform_(0.5, 1.5), ref.bias.uniform_(-0.5, 0.5)
    own = bricks.BatchNorm2d(C).cuda()
    own.load_state_dict(ref.state_dict())
    ref.train(), own.train()
    xr = x.to(dtype).float().clone().requires_grad_()
    want = ref(xr)
    want = F.relu(want) if relu else want
    dy = torch.randn(want.shape)
    want.backward(dy.to(dtype).float())
    xc = _cl(x.detach().to(dtype).cuda()).requires_grad_()
    got = own(xc, relu=relu)
    got.backward(_cl(dy.to(dtype).cuda()))
    tol_y, tol_g = (1e-4, 1e-4) if dtype == torch.float32 else (4e-3, 1e-2)
    assert rel(got, want) < tol_y, rel(got, want)
    assert rel(xc.grad, xr.grad) < tol_g, rel(xc.grad, xr.grad)
    assert rel(own.weight.grad, ref.weight.grad) < tol_g and rel(own.bias.grad, ref.bias.grad) < tol_g
    # running statistics (momentum 0.1, unbiased variance) and the batch counter
    assert rel(own.running_mean, ref.running_mean) < 1e-3 and rel(own.running_var, ref.running_var) < 1e-3
    assert int(own.num_batches_tracked) == int(ref.num_batches_tracked) == 1
    # evaluation mode goes through the running statistics (torch)
    own.eval(), ref.eval()
    with torch.no_grad():
        assert rel(own(xc.detach()), ref(xr.detach())) < tol_y
