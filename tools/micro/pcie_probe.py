import torch, time
x = torch.empty(96*1024*1024, dtype=torch.uint8, device='cuda')
t0=time.perf_counter(); h = torch.empty(96*1024*1024, dtype=torch.uint8, pin_memory=True); t1=time.perf_counter()
print('pinned alloc 96MB: %.2f ms'%((t1-t0)*1e3))
for _ in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter(); h.copy_(x, non_blocking=True); torch.cuda.synchronize(); t1=time.perf_counter()
    print('D2H pinned %.1f GB/s'%(x.numel()/(t1-t0)/1e9))
for _ in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter(); x.copy_(h, non_blocking=True); torch.cuda.synchronize(); t1=time.perf_counter()
    print('H2D pinned %.1f GB/s'%(x.numel()/(t1-t0)/1e9))
p = torch.empty(96*1024*1024, dtype=torch.uint8)
for _ in range(2):
    torch.cuda.synchronize(); t0=time.perf_counter(); x.copy_(p); torch.cuda.synchronize(); t1=time.perf_counter()
    print('H2D pageable %.1f GB/s'%(x.numel()/(t1-t0)/1e9))
