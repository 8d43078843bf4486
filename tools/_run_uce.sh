timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "upsample_ce or normalize_u8" 2>&1 | tail -6
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_uper_head.py -x -q 2>&1 | tail -4
python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
from rscotr_b200 import ops
from tools.kbench import timeit
for (B,C,h,w,H,W) in [(2,100,100,100,800,800),(2,6,100,100,800,800),(8,6,128,128,512,512)]:
    x=torch.randn(B,C,h,w,device='cuda',dtype=torch.bfloat16,requires_grad=True)
    lab=torch.randint(0,C,(B,H,W),device='cuda')
    st=ops.upsample_ce(x,lab,255)
    tf=timeit(lambda: ops.upsample_ce(x,lab,255))
    tb=timeit(lambda: torch.autograd.grad(st[0],x,retain_graph=True))
    print('upsample_ce',(B,C,h,w,H,W),'fwd_us',round(tf*1e3,1),'bwd_us',round(tb*1e3,1))
PY
