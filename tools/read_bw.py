import torch, time
n=100000000
x=torch.randn(n,device='cuda',dtype=torch.float64)
def t(fn,reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
ms=t(lambda: torch.sum(x)); print('torch.sum f64 1e8: %.4f ms %.1f GB/s'%(ms, 0.8/ms*1e3))
ms=t(lambda: torch.max(x)); print('torch.max f64 1e8: %.4f ms %.1f GB/s'%(ms, 0.8/ms*1e3))
y=torch.empty_like(x)
ms=t(lambda: y.copy_(x)); print('copy f64 1e8: %.4f ms %.1f GB/s'%(ms, 1.6/ms*1e3))
xf=x.view(torch.float32)
ms=t(lambda: torch.sum(xf)); print('torch.sum f32 2e8: %.4f ms %.1f GB/s'%(ms, 0.8/ms*1e3))
