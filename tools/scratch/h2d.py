import torch, time
n=444; fb=921600
big=torch.empty(n*fb,dtype=torch.uint8).pin_memory()
big.random_(0,255)
h=[big[i*fb:(i+1)*fb] for i in range(n)]
d=torch.empty(n*fb,dtype=torch.uint8,device='cuda')
for ns in (1,2,4,8):
    ss=[torch.cuda.Stream() for _ in range(ns)]
    for rep in range(3):
        torch.cuda.synchronize(); t=time.perf_counter()
        for i in range(n):
            with torch.cuda.stream(ss[i%ns]): d[i*fb:(i+1)*fb].copy_(h[i],non_blocking=True)
        t1=time.perf_counter()-t
        torch.cuda.synchronize(); dt=time.perf_counter()-t
    print(ns,"streams: distinct 921KB copies GB/s", n*fb/dt/1e9, "ms", dt*1e3, "issue ms", t1*1e3)
# separately allocated pinned buffers (as frames arrive)
hs=[torch.empty(fb,dtype=torch.uint8).pin_memory() for _ in range(n)]
for ns in (1,2,4):
    ss=[torch.cuda.Stream() for _ in range(ns)]
    for rep in range(3):
        torch.cuda.synchronize(); t=time.perf_counter()
        for i in range(n):
            with torch.cuda.stream(ss[i%ns]): d[i*fb:(i+1)*fb].copy_(hs[i],non_blocking=True)
        torch.cuda.synchronize(); dt=time.perf_counter()-t
    print(ns,"streams: separately pinned GB/s", n*fb/dt/1e9, "ms", dt*1e3)
for rep in range(3):
    torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(big,non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
print("h2d one-shot GB/s", n*fb/dt/1e9)
