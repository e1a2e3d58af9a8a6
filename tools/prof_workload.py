"""C2 workload, resident inputs, a few renders: the target of the ncu captures (profiles/)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import graphaudio_b200 as G
from graphaudio_b200 import _native as N
from graphaudio_b200.api import check
from tests import synth
import bench
wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"])
if len(sys.argv) > 2: wl["voices"] = int(sys.argv[2])
renders = int(sys.argv[3]) if len(sys.argv) > 3 else 2
voices = bench.make_inputs(wl, 0, pinned=False)
ctx = bench.build_graph(G, wl, voices, device_id=0)
g = ctx._graph()
n = int(wl["render_s"] * bench.FS)
d_out = torch.empty((2, n), dtype=torch.float32, device="cuda")
st = N.gac_stats()
for i in range(renders):
    check(N.lib().gac_render_device(ctx._h, g, 0, n, C.c_void_p(d_out.data_ptr()), 2, 1))
    check(N.lib().gac_get_stats(ctx._h, C.byref(st)))
    print({k: round(v, 4) if isinstance(v, float) else v for k, v in st.as_dict().items()})
