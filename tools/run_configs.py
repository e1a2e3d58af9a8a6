"""Device-time measurements of the BASELINE configs other than the bench workload (per-GPU shards of C3, C4, C5 and C1),
inputs resident in HBM.  Writes profiles/r02_configs.json.   python tools/run_configs.py [c1 c3 c4 c5]"""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import graphaudio_b200 as G
from graphaudio_b200 import _native as N
from graphaudio_b200.api import check
from tests import synth

which = sys.argv[1:] or ["c1", "c3", "c4", "c5"]
# under torchrun every rank measures its own shard on its own GPU (C4 / C5 shard by render / voice with no exchange step)
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))
RANK = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(LOCAL)
_OAC = G.OfflineAudioContext
G.OfflineAudioContext = lambda fs=48000, **kw: _OAC(fs, **{"device_id": LOCAL, **kw})
res = {}
L = N.lib()


def stats_of(ctx_h):
    st = N.gac_stats()
    check(L.gac_get_stats(ctx_h, C.byref(st)))
    return st.as_dict()


def run_single(name, ctx, n, voices, seconds, reps=5):
    g = ctx._graph()
    d_out = torch.empty((2, n), dtype=torch.float32, device="cuda")
    ms = []
    for i in range(3 + reps):
        check(L.gac_render_device(ctx._h, g, 0, n, C.c_void_p(d_out.data_ptr()), 2, 1))
        s = stats_of(ctx._h)
        if i >= 3:
            ms.append(s)
    t = float(np.mean([s["ms_total"] for s in ms]))
    res[name] = {"voices_per_gpu": voices, "rendered_seconds": seconds, "ms_per_render": t, "voice_s_per_s": voices * seconds / (t * 1e-3),
                 "realtime_factor": seconds / (t * 1e-3),
                 "kernel_ms": {k: float(np.mean([s[k] for s in ms])) for k in ms[0] if k.startswith("ms_")},
                 "conv_units": ms[0]["conv_units"], "algorithmic_bytes": ms[0]["algorithmic_bytes"], "mac_variant_used": ms[0]["mac_variant_used"]}
    print(name, json.dumps(res[name]), flush=True)
    L.gac_graph_destroy(g)


if "c1" in which:
    fs = 48000
    src, ir = synth.make_voice_inputs(0, 10 * fs, fs)
    ctx = synth.build_c1(G, fs, src, ir)
    run_single("C1 (1 voice, 1 s IR, 11 s)", ctx, 11 * fs, 1, 11.0)
    ctx.Dispose()

if "c3" in which:
    fs = 48000
    voices = [synth.make_voice_inputs(v, 10 * fs, 2 * fs) + (synth.voice_gains(v),) for v in range(128)]
    ctx = synth.build_c3(G, fs, voices, 1.0 / 32)
    run_single("C3 shard (128 of 1024 voices: biquad sweep -> gain -> 2 s IR, 12 s)", ctx, 12 * fs, 128, 12.0)
    ctx.Dispose()
    del voices

if "c5" in which or "c5p512" in which:
    fs, src_rate = 96000, 44100
    voices = []
    for v in range(32):
        src = [synth.splitmix_uniform(4 * v + c, 441000) for c in range(2)]
        ir = [synth.decay_ir(4 * v + 2 + c, 960000) for c in range(2)]
        voices.append((src, ir, synth.voice_gains(v)))
    for part in ((512,) if "c5p512" in which else (512, 128)):
        ctx = synth.build_c5(G, fs, src_rate, voices, 1.0 / 16, partition=part)
        run_single(f"C5 shard (32 of 256 voices: 44.1->96 kHz resample -> gain -> 10 s IR, 20 s, partition {part})", ctx, 20 * fs, 32, 20.0, reps=3)
        ctx.Dispose()
    del voices

if "c4" in which or "c4small" in which:
    fs = 48000
    NR = 128 if "c4small" in which else 512
    parent = G.OfflineAudioContext(fs)
    forks = []
    for r in range(NR):
        src, ir = synth.make_voice_inputs(RANK * NR + r, 5 * fs, fs // 2)
        f = parent.Fork()
        s = G.AudioBufferSourceNode(f)
        s.Buffer = G.PlayableAudioBuffer.FromChannelArrays(src, fs)
        lp = G.BiQuadFilterNode(f)
        lp.Type = G.FilterType.Lowpass
        lp.Q.Value = 0.707
        lp.Frequency.SetValueAtTime(2000.0, 0.0)
        lp.Frequency.ExponentialRampToValueAtTime(12000.0, 4.0)
        hp = G.BiQuadFilterNode(f)
        hp.Type = G.FilterType.Highpass
        hp.Frequency.Value = 200.0
        hp.Q.Value = 0.707
        conv = G.ConvolverNode(f)
        conv.Buffer = G.PlayableAudioBuffer.FromChannelArrays(ir, fs)
        s.Connect(lp).Connect(hp).Connect(conv).Connect(f.Destination)
        s.Start()
        forks.append(f)
    n = 5 * fs
    graphs = [f._graph() for f in forks]
    garr = (C.c_void_p * NR)(*graphs)
    out = torch.zeros((NR, 2, n), dtype=torch.float32, pin_memory=True).numpy()
    rows = (N.fp * (2 * NR))(*[out[g, c].ctypes.data_as(N.fp) for g in range(NR) for c in range(2)])
    ms = []
    for i in range(2 + 3):
        t0 = time.perf_counter()
        check(L.gac_render_batch(parent._h, garr, NR, n, rows, 2))
        wall = time.perf_counter() - t0
        s = stats_of(parent._h)
        s["wall_ms"] = wall * 1e3
        if i >= 2:
            ms.append(s)
    t = float(np.mean([s["ms_total"] for s in ms]))
    d2h = float(np.mean([s["ms_d2h"] for s in ms]))
    res[f"C4 shard ({NR} of 4096 independent renders: 2 biquads -> 0.5 s IR, 5 s)"] = {
        "renders_per_gpu": NR, "rendered_seconds": 5.0, "ms_per_batch": t, "ms_d2h_of_983MB_results": d2h, "ms_compute": t - d2h,
        "voice_s_per_s": NR * 5.0 / (t * 1e-3), "voice_s_per_s_compute_only": NR * 5.0 / ((t - d2h) * 1e-3), "wall_ms": float(np.mean([s["wall_ms"] for s in ms])),
        "kernel_ms": {k: float(np.mean([s[k] for s in ms])) for k in ms[0] if k.startswith("ms_")}, "mac_variant_used": ms[0]["mac_variant_used"]}
    print(json.dumps(res[f"C4 shard ({NR} of 4096 independent renders: 2 biquads -> 0.5 s IR, 5 s)"]), flush=True)
    assert np.isfinite(out).all() and np.abs(out).max() > 1e-3
    parent.Dispose()

if "f3" in which:
    # SURVEY.md §8f-3 nodes at the bench workload's size: 64 voices x (stereo 10 s source -> DelayNode -> StereoPannerNode) -> bus, 12 s.
    # Algorithmic bytes per frame and voice: delay 16 (+4 with an a-rate table), panner 16 (+4); the automation tables are written once (4 each).
    fs = 48000
    n = 12 * fs
    for label, automated in (("constant DelayTime / Pan", False), ("a-rate DelayTime and Pan", True)):
        ctx = G.OfflineAudioContext(fs)
        bus = G.GainNode(ctx)
        bus.Gain.Value = 0.125
        bus.Connect(ctx.Destination)
        for v in range(64):
            s = G.AudioBufferSourceNode(ctx)
            s.Buffer = G.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(4 * v + c, 10 * fs) for c in range(2)], fs)
            d = G.DelayNode(ctx, 0.5)
            p = G.StereoPannerNode(ctx)
            if automated:
                d.DelayTime.SetValueAtTime(0.01 + 0.001 * v, 0.0)
                d.DelayTime.LinearRampToValueAtTime(0.3, 8.0)
                p.Pan.SetValueAtTime(-1.0, 0.0)
                p.Pan.LinearRampToValueAtTime(1.0, 6.0 + 0.05 * v)
            else:
                d.DelayTime.Value = 0.01 + 0.001 * v
                p.Pan.Value = -1.0 + v / 32.0
            s.Connect(d).Connect(p).Connect(bus)
            s.Start()
        g = ctx._graph()
        d_out = torch.empty((2, n), dtype=torch.float32, device="cuda")
        spans = {"delay": [], "panner": [], "total": []}
        for i in range(6):
            check(L.gac_render_device(ctx._h, g, 0, n, C.c_void_p(d_out.data_ptr()), 2, 1))
            st = stats_of(ctx._h)
            if i >= 3:
                spans["total"].append(st["ms_total"]); spans["delay"].append(st["ms_delay"]); spans["panner"].append(st["ms_panner"])
        frames = 64 * n
        dms, pms = float(np.mean(spans["delay"])), float(np.mean(spans["panner"]))
        extra = 4 if automated else 0
        res[f"f3: 64 voices x (source -> DelayNode -> StereoPannerNode) -> bus, 12 s, {label}"] = {
            "ms_per_render": float(np.mean(spans["total"])), "ms_delay": dms, "ms_panner": pms,
            "delay_GBps": frames * (16 + extra) / (dms * 1e-3) / 1e9, "panner_GBps": frames * (16 + extra) / (pms * 1e-3) / 1e9,
            "hbm_peak_GBps": 6458.4}
        print(json.dumps(res[f"f3: 64 voices x (source -> DelayNode -> StereoPannerNode) -> bus, 12 s, {label}"]), flush=True)
        L.gac_graph_destroy(g)
        ctx.Dispose()

p = os.path.join(ROOT, "gpurun_out", "r02_configs.json" if "WORLD_SIZE" not in os.environ else f"r02_configs_rank{RANK}of{os.environ['WORLD_SIZE']}.json")
if os.path.exists(p) and len(which) < 5:
    old = json.load(open(p)); old.update(res); res = old
json.dump(res, open(p, "w"), indent=1)
print("wrote", p)
