"""Host-side phase timestamps of one resident render (GAC_TRACE=1): where does the host spend its time while the stream runs?

    GAC_TRACE=1 python tools/trace_render.py [--voices 1024] [--workload c3]
"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import graphaudio_b200 as G  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--voices", type=int, default=1024)
ap.add_argument("--workload", default="c3")
ap.add_argument("--renders", type=int, default=4)
a = ap.parse_args()
wl = dict(bench.WORKLOADS[a.workload])
n = int(wl["render_s"] * bench.FS)
voices = bench.make_inputs(wl, 0, a.voices, pinned=False, cheap=True)
ctx = bench.build_graph(G, wl, voices, device_id=0)
ctx.MarkBus(ctx.bus)
out = np.zeros((2, n), np.float32)
for r in range(a.renders):
    sys.stderr.write(f"---- render {r}\n")
    t = time.perf_counter()
    ctx.RenderSharded(out, n, 0)
    sys.stderr.write(f"---- render {r}: wall {1e3 * (time.perf_counter() - t):.2f} ms, device {ctx.last_stats['ms_total']:.2f} ms\n")
ctx.Dispose()
