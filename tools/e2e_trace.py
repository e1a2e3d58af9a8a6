"""Where does an end-to-end step go?  Host timestamps of the phases of bench.py's e2e arm (fresh context, uploads from page-locked
arrays, graph, render) with GAC_TRACE=1 marks of the render itself.

    GAC_TRACE=1 python tools/e2e_trace.py [--workload c3] [--voices 128] [--steps 3] [--sync-upload]
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch  # noqa: E402
import graphaudio_b200 as G  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3")
ap.add_argument("--voices", type=int, default=128)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--sync-upload", dest="sync_upload", action="store_true")
a = ap.parse_args()
wl = dict(bench.WORKLOADS[a.workload])
voices = bench.make_inputs(wl, 0, a.voices, pinned=True)
n = int(wl["render_s"] * wl.get("fs", bench.FS))
out = torch.zeros((2, n), dtype=torch.float32, pin_memory=True).numpy()
h2d = sum(x.nbytes for v in voices for x in (v[0] + v[1]))
for it in range(a.steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    c = G.OfflineAudioContext(wl.get("fs", bench.FS), device_id=0, async_upload=not a.sync_upload)
    t1 = time.perf_counter()
    bench.build_into(G, wl, voices, c)
    c.MarkBus(c.bus)
    t2 = time.perf_counter()
    sys.stderr.write(f"[py] context +{1e3 * (t1 - t0):.3f} ms, graph built (uploads queued) +{1e3 * (t2 - t0):.3f} ms\n")
    c.RenderSharded(out, n, 0)
    t3 = time.perf_counter()
    sys.stderr.write(f"[py] render returned +{1e3 * (t3 - t0):.3f} ms (render call {1e3 * (t3 - t2):.3f}, device {c.last_stats['ms_total']:.3f}); "
                     f"{h2d / 1e6:.0f} MB uploaded = {h2d / (t3 - t0) / 1e9:.1f} GB/s over the step\n")
    c.Dispose()
