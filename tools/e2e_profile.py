"""cProfile of the host side of bench.py's end-to-end step (fresh context, uploads, graph, flatten, render): where the Python mirror spends
its time.    python tools/e2e_profile.py [--workload c3] [--voices 128] [--steps 6]"""
import argparse, cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch  # noqa: E402
import graphaudio_b200 as G  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3")
ap.add_argument("--voices", type=int, default=128)
ap.add_argument("--steps", type=int, default=6)
a = ap.parse_args()
wl = dict(bench.WORKLOADS[a.workload])
voices = bench.make_inputs(wl, 0, a.voices, pinned=True)
n = int(wl["render_s"] * wl.get("fs", bench.FS))
out = torch.zeros((2, n), dtype=torch.float32, pin_memory=True).numpy()


def step():
    c = G.OfflineAudioContext(wl.get("fs", bench.FS), device_id=0, async_upload=True)
    bench.build_into(G, wl, voices, c)
    c.MarkBus(c.bus)
    c.RenderSharded(out, n, 0)
    c.Dispose()


for _ in range(3):
    step()
import gc
gc.collect(); gc.freeze()
t0 = time.perf_counter()
for _ in range(a.steps):
    step()
print(f"un-profiled: {(time.perf_counter() - t0) / a.steps * 1e3:.2f} ms per step")
pr = cProfile.Profile()
pr.enable()
for _ in range(a.steps):
    step()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(32)
