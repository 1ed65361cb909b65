import torch, time
dev=torch.device('cuda:0')
n=4*(1<<30)
h=torch.empty(n,dtype=torch.uint8).pin_memory()
h2=torch.empty(n,dtype=torch.uint8).pin_memory()
d=torch.empty(n,dtype=torch.uint8,device=dev)
d2=torch.empty(n,dtype=torch.uint8,device=dev)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def t(fn,reps=3):
    fn(); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter()-t0)/reps
a=t(lambda: d.copy_(h,non_blocking=True)); print("H2D GB/s", n/a/1e9)
b=t(lambda: h.copy_(d,non_blocking=True)); print("D2H GB/s", n/b/1e9)
def both():
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
c=t(both); print("bidirectional GB/s each", n/c/1e9)
t0=time.perf_counter(); x=torch.empty(8*(1<<30),dtype=torch.uint8).pin_memory(); print("pin 8GB s", time.perf_counter()-t0)
import ctypes
