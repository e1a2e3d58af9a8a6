"""Pure pinned-host -> device copy bandwidth of the box (the floor of the end-to-end step: 295 MB of inputs per C2 render)."""
import time, torch
def run(label, h, d, fn):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for a, b in zip(h, d): fn(a, b)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    nbytes = sum(a.numel() * a.element_size() for a in h)
    print(f"{label:60s} {dt*1e3:7.3f} ms  {nbytes/dt/1e9:6.1f} GB/s")
for mb, n in ((295, 1), (4.6, 64), (1.92, 128), (0.384, 128)):
    nb = int(mb * 1e6)
    h = [torch.empty(nb, dtype=torch.uint8, pin_memory=True) for _ in range(n)]
    d = [torch.empty(nb, dtype=torch.uint8, device="cuda") for _ in range(n)]
    run(f"{n} copies of {mb} MB (1-D)", h, d, lambda a, b: b.copy_(a, non_blocking=True))
# the shape of the e2e step: per voice a [2, 480000] source and a [2, 96000] IR, each one strided (2-D) copy into a padded device buffer
hs = [torch.empty((2, 480000), dtype=torch.float32, pin_memory=True) for _ in range(64)]
hi = [torch.empty((2, 96000), dtype=torch.float32, pin_memory=True) for _ in range(64)]
ds = [torch.empty((2, 480064), dtype=torch.float32, device="cuda") for _ in range(64)]
di = [torch.empty((2, 96064), dtype=torch.float32, device="cuda") for _ in range(64)]
h2 = [x for p in zip(hs, hi) for x in p]
d2 = [x for p in zip(ds, di) for x in p]
run("64 x (source 3.84 MB + IR 0.77 MB), strided 2-D copies", h2, d2, lambda a, b: b[:, :a.shape[1]].copy_(a, non_blocking=True))
run("the same as 256 1-D row copies", h2, d2, lambda a, b: (b[0, :a.shape[1]].copy_(a[0], non_blocking=True), b[1, :a.shape[1]].copy_(a[1], non_blocking=True)))
# do several copy streams hide the per-copy setup gap of the DMA engine?  (the e2e shape as 128 contiguous copies: 64 x 3.84 MB + 64 x 0.77 MB)
hs1 = [x.reshape(-1) for x in hs]; hi1 = [x.reshape(-1) for x in hi]
ds1 = [torch.empty(2 * 480000, dtype=torch.float32, device="cuda") for _ in range(64)]
di1 = [torch.empty(2 * 96000, dtype=torch.float32, device="cuda") for _ in range(64)]
h3 = [x for p in zip(hs1, hi1) for x in p]; d3 = [x for p in zip(ds1, di1) for x in p]
for ns in (1, 2, 3, 4):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i, (a, b) in enumerate(zip(h3, d3)):
            with torch.cuda.stream(streams[i % ns]):
                b.copy_(a, non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    nbytes = sum(a.numel() * 4 for a in h3)
    print(f"{'128 contiguous copies (source, IR alternating) over %d stream(s)' % ns:60s} {dt*1e3:7.3f} ms  {nbytes/dt/1e9:6.1f} GB/s")
