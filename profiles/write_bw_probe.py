import torch
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n*1e-3
n = 1216*1048576//4
a = torch.empty(n, dtype=torch.float32, device='cuda'); b = torch.empty(n, dtype=torch.float32, device='cuda')
s = t(lambda: a.fill_(1.0)); print("fill 1.275GB  write GB/s", a.numel()*4/s/1e9)
s = t(lambda: a.zero_()); print("zero (memset)  write GB/s", a.numel()*4/s/1e9)
s = t(lambda: b.copy_(a)); print("copy r+w GB/s", 2*a.numel()*4/s/1e9)
big = torch.empty(2**30, dtype=torch.float32, device='cuda')
s = t(lambda: big.fill_(1.0)); print("fill 4GB write GB/s", big.numel()*4/s/1e9)
s = t(lambda: a.sum()); print("read-only (sum) GB/s", a.numel()*4/s/1e9)
