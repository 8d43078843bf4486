import torch, torch.nn.functional as F
dev='cuda'
def t(f,n=30):
    for _ in range(5): f()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n*1000
for (M,K,N) in [(13294,2048,256),(13294,256,2048),(13294,256,256),(792,2048,256),(26588,2048,256)]:
    x3=torch.randn(M,1,K,device=dev,dtype=torch.bfloat16); w=torch.randn(N,K,device=dev,dtype=torch.bfloat16); b=torch.randn(N,device=dev,dtype=torch.bfloat16)
    x2=x3.view(M,K)
    print(M,K,N,'linear3d_nobias %.1f'%t(lambda:F.linear(x3,w)),'linear2d_nobias %.1f'%t(lambda:F.linear(x2,w)),'mm %.1f'%t(lambda:torch.mm(x2,w.t())),
          'linear3d_bias %.1f'%t(lambda:F.linear(x3,w,b)),'addmm %.1f'%t(lambda:torch.addmm(b,x2,w.t())), 'us')
