import torch, time
x = torch.empty(80*1024*1024//8, dtype=torch.float64, pin_memory=True)
y = torch.empty_like(x, device='cuda')
z = torch.empty(40*1024*1024//8, dtype=torch.float64, pin_memory=True)
w = torch.empty_like(z, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for _ in range(3): y.copy_(x, non_blocking=True); z.copy_(w, non_blocking=True)
torch.cuda.synchronize()
t=time.perf_counter(); 
for _ in range(10): y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
print("H2D 80MB: %.3f ms  %.1f GB/s" % (dt*1e3, 80*1.048576/1e3/dt))
t=time.perf_counter()
for _ in range(10): z.copy_(w, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
print("D2H 40MB: %.3f ms  %.1f GB/s" % (dt*1e3, 40*1.048576/1e3/dt))
t=time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): y.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): z.copy_(w, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
print("both concurrently: %.3f ms" % (dt*1e3))
