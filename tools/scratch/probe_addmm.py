import torch, time
dev='cuda'
torch.manual_seed(0)
for (M,N,K) in [(2048,256,13294),(256,256,792),(288,96,640000)]:
    dy=torch.randn(K,M,device=dev,dtype=torch.bfloat16); x=torch.randn(K,N,device=dev,dtype=torch.bfloat16)
    g=torch.randn(M,N,device=dev)
    ref=g+ (dy.float().t()@x.float())
    g2=g.clone()
    try:
        torch.addmm(g2, dy.t(), x, out_dtype=torch.float32, out=g2)
        print('addmm dtype_out ok', (g2-ref).abs().max().item(), ref.abs().max().item())
    except Exception as e:
        print('addmm dtype_out FAIL', e)
    try:
        r=torch.mm(dy.t(), x, out_dtype=torch.float32)
        print('mm dtype ok', (r+g-ref).abs().max().item())
    except Exception as e:
        print('mm dtype FAIL', e)
    def t(f,n=20):
        for _ in range(3): f()
        torch.cuda.synchronize(); t0=time.perf_counter()
        for _ in range(n): f()
        torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e6
    print(M,N,K,'addmm_acc us', t(lambda: torch.addmm(g2, dy.t(), x, out_dtype=torch.float32, out=g2)),
          'mm bf16 + cast us', t(lambda: torch.mm(dy.t(), x).float()), 'mm only', t(lambda: torch.mm(dy.t(), x)))
