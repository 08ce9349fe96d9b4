import torch
dev='cuda:0'
n=400*1024*1024//2
a=torch.empty(n,dtype=torch.bfloat16,device=dev); b=torch.empty(n,dtype=torch.bfloat16,device=dev)
def t(f,it=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/it*1e-3
ts=t(lambda: b.copy_(a)); print('copy 400MB->400MB', 2*n*2/ts/1e9,'GB/s')
ts=t(lambda: b.fill_(1.0)); print('fill 400MB', n*2/ts/1e9,'GB/s')
ts=t(lambda: a.sum()); print('sum 400MB (read only)', n*2/ts/1e9,'GB/s')
c=torch.empty(n//4,dtype=torch.bfloat16,device=dev)
ts=t(lambda: torch.cat([c,c,c],out=b[:3*(n//4)])); print('read 100MB write 300MB', (n//4*2+3*(n//4)*2)/ts/1e9,'GB/s')
