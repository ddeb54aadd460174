import torch, time
n = 2_000_000_000
host = [torch.empty(n//4, dtype=torch.uint8).pin_memory() for _ in range(4)]
dev = [torch.empty(n//4, dtype=torch.uint8, device='cuda') for _ in range(4)]
for k in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(k)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(4):
        with torch.cuda.stream(streams[i % k]):
            dev[i].copy_(host[i], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"H2D {k} streams: {n/dt/1e9:.1f} GB/s")
torch.cuda.synchronize(); t0=time.perf_counter()
for i in range(4): host[i].copy_(dev[i], non_blocking=True)
torch.cuda.synchronize(); print(f"D2H: {n/(time.perf_counter()-t0)/1e9:.1f} GB/s")
