"""Where does the host spend the build phase of the e2e step (C2, 64 voices)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import graphaudio_b200 as G
from tests import synth
import bench
wl = bench.WORKLOADS["c2"]
voices = bench.make_inputs(wl, 0, pinned=True)
for it in range(4):
    torch.cuda.synchronize()
    T = dict(src=0.0, ir=0.0, nodes=0.0)
    t_all = time.perf_counter()
    ctx = G.OfflineAudioContext(bench.FS, device_id=0, async_upload=True)
    bus = G.GainNode(ctx); bus.Gain.Value = 0.125; bus.Connect(ctx.Destination)
    for src, ir, g in voices:
        t0 = time.perf_counter()
        s = G.AudioBufferSourceNode(ctx)
        s.Buffer = G.PlayableAudioBuffer.FromChannelArrays(src, bench.FS)
        t1 = time.perf_counter()
        gn = G.GainNode(ctx); synth.add_gain_automation(gn.Gain, g, 1.0)
        conv = G.ConvolverNode(ctx)
        t2 = time.perf_counter()
        conv.Buffer = G.PlayableAudioBuffer.FromChannelArrays(ir, bench.FS)
        t3 = time.perf_counter()
        s.Connect(gn).Connect(conv).Connect(bus); s.Start()
        t4 = time.perf_counter()
        T["src"] += t1 - t0; T["ir"] += t3 - t2; T["nodes"] += (t2 - t1) + (t4 - t3)
    total = time.perf_counter() - t_all
    torch.cuda.synchronize()
    print(f"build {total*1e3:.2f} ms: source buffers {T['src']*1e3:.2f}, IR buffers + prepare {T['ir']*1e3:.2f}, nodes/params/connect {T['nodes']*1e3:.2f}")
    ctx.Dispose()
