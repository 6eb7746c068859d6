import torch, time
def rate(fn, reps=20):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t)/reps
n=24576
sizes=[27,26,9,24,24,12,12,20]
hs=[torch.empty(n*s, dtype=torch.float64, pin_memory=True).fill_(1.0) for s in sizes]
ds=[torch.empty(n*s, dtype=torch.float64, device='cuda') for s in sizes]
big_h=torch.empty(n*sum(sizes), dtype=torch.float64, pin_memory=True).fill_(1.0); big_d=torch.empty_like(big_h, device='cuda')
streams=[torch.cuda.Stream() for _ in range(4)]
def sep():
    for h,d in zip(hs,ds): d.copy_(h, non_blocking=True)
def one(): big_d.copy_(big_h, non_blocking=True)
def multi():
    for i,(h,d) in enumerate(zip(hs,ds)):
        with torch.cuda.stream(streams[i%4]): d.copy_(h, non_blocking=True)
for f in (sep,one,multi): rate(f,3)
mb=n*sum(sizes)*8/1e6
for name,f in (("8 separate copies, one stream",sep),("one packed copy",one),("8 copies over 4 streams",multi)):
    dt=rate(f); print(f"{name:32s} {dt*1e3:.3f} ms  {mb/1e3/dt:.1f} GB/s  ({mb:.1f} MB)")
print("asyncEngineCount", torch.cuda.get_device_properties(0).multi_processor_count)
