"""Pure pinned-host -> device copy bandwidth of the box (the floor of the end-to-end step: 295 MB of inputs per C2 render)."""
import time, torch
for mb, n in ((295, 1), (4.6, 64), (1.92, 128), (0.384, 128)):
    nb = int(mb * 1e6)
    h = [torch.empty(nb, dtype=torch.uint8, pin_memory=True) for _ in range(n)]
    d = [torch.empty(nb, dtype=torch.uint8, device="cuda") for _ in range(n)]
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for a, b in zip(h, d): b.copy_(a, non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{n:4d} copies of {mb:7.3f} MB: {dt*1e3:7.3f} ms  {n*nb/dt/1e9:6.1f} GB/s")
