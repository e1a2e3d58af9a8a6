import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import graphaudio_b200 as G
import bench
wl = bench.WORKLOADS["c2"]
voices = bench.make_inputs(wl, 0, pinned=True)
n = int(wl["render_s"] * bench.FS)
out = torch.zeros((2, n), dtype=torch.float32, pin_memory=True).numpy()
for it in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    c = G.OfflineAudioContext(bench.FS, device_id=0, async_upload=True)
    bench.build_into(G, wl, voices, c)
    t2 = time.perf_counter()
    sys.stderr.write(f"[py] build done +{1e3*(t2-t0):.3f} ms\n")
    c.Render(out, n, 0)
    t3 = time.perf_counter()
    sys.stderr.write(f"[py] render returned +{1e3*(t3-t0):.3f} ms (render call {1e3*(t3-t2):.3f}) device {c.last_stats['ms_total']:.3f}\n")
    c.Dispose()
