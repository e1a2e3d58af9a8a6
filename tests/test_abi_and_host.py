"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/graphaudio_cuda.h declares, struct layouts
match the header, the host-side mirror records topology/automation like the reference, and there is NO CPU fallback."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

import graphaudio_b200 as G
from graphaudio_b200 import _native as N
from graphaudio_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "graphaudio_cuda.h")


def _declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gac_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = N.lib()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in graphaudio_cuda.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(N.SIGNATURES) == declared


def test_version_and_struct_layouts():
    assert N.lib().gac_version() == 7
    # gac_event must be bit-compatible with AutomationEvent (AudioParam.cs:360-367): int, float, float, (pad), double, double
    assert C.sizeof(N.gac_event) == 32
    assert N.gac_event.time.offset == 16 and N.gac_event.time_constant.offset == 24
    assert C.sizeof(N.gac_param) == 32
    assert C.sizeof(N.gac_context_desc) == 32
    assert C.sizeof(N.gac_stats) == 184


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    n = C.c_int(-1)
    rc = N.lib().gac_device_count(C.byref(n))
    assert rc == N.GAC_ERR_NO_DEVICE and n.value == 0
    with pytest.raises(G.CudaException) as e:
        G.OfflineAudioContext(48000)
    assert "no CPU fallback" in str(e.value)
    assert N.lib().gac_render(None, None, 0, 128, None, 2, 0) != 0  # null handles are rejected, never rendered on the host


def test_audio_param_event_order_and_clamping():
    """AddEvent keeps time order with a stable upper-bound insert (AudioParam.cs:333-352); values clamp at schedule time."""
    ctx = G.OfflineAudioContext(48000, _record_only=True)
    bq = G.BiQuadFilterNode(ctx)
    f = bq.Frequency
    f.LinearRampToValueAtTime(5000.0, 2.0)
    f.SetValueAtTime(100.0, 1.0)
    f.SetValueAtTime(200.0, 1.0)      # equal times keep call order
    f.SetValueAtTime(1e9, 0.5)        # clamped to fs/2
    f.SetTargetAtTime(0.0, 3.0, 0.1)  # clamped to the minimum (1 Hz)
    ev = f._events
    assert [e[3] for e in ev] == [0.5, 1.0, 1.0, 2.0, 3.0]
    assert [e[1] for e in ev[:4]] == [24000.0, 100.0, 200.0, 5000.0]
    assert ev[4][0] == 3 and ev[4][2] == 1.0
    f.CancelScheduledValues(1.0)      # removes events with time >= 1.0 (:312-331)
    assert [e[3] for e in f._events] == [0.5]
    f.Value = 0.0                     # clamps and clears (:34-49)
    assert f.Value == 1.0 and f._events == []
    with pytest.raises(G.ArgumentException):
        G.GainNode(ctx).Gain.ExponentialRampToValueAtTime(0.0, 1.0)


def test_topology_flattening_matches_connection_order():
    ctx = G.OfflineAudioContext(48000, _record_only=True)
    buf = G.PlayableAudioBuffer.FromMonoArray(np.zeros(256, np.float32), 48000)
    bus = G.GainNode(ctx)
    direct = G.AudioBufferSourceNode(ctx)
    direct.Buffer = buf
    direct.Connect(ctx.Destination)          # destination input 0: a direct voice
    bus.Connect(ctx.Destination)             # destination input 1: the bus
    chains = []
    for v in range(3):
        s = G.AudioBufferSourceNode(ctx)
        s.Buffer = buf
        bq, g, cv = G.BiQuadFilterNode(ctx), G.GainNode(ctx), G.ConvolverNode(ctx)
        s.Connect(bq).Connect(g).Connect(cv).Connect(bus)
        s.Start(0.1 * v)
        chains.append((s, [bq, g, cv]))
    voices, buses, dest = ctx._topology()
    assert dest == [~0, 0]
    assert voices[0][0] is direct and voices[0][2] == -1
    for i, (s, ops) in enumerate(chains):
        assert voices[1 + i][0] is s and voices[1 + i][1] == ops and voices[1 + i][2] == 0
    assert buses == [[bus]]
    with pytest.raises(G.InvalidOperationException):
        chains[0][0].Start()  # "can only be started once" (AudioBufferSourceNode.cs:83-84)
    with pytest.raises(G.InvalidOperationException):
        ctx.Render(128)       # record-only context: no CPU render path exists


def test_fan_out_and_bus_hierarchies_flatten_to_buses_and_bus_fed_chains():
    """ReverbEffect-style dry / wet split (GraphAudio.Kit/Effects/ReverbEffect.cs:63-91) behind a two-source input, into a
    master bus: the fan-in nodes become buses, the node whose output fans out ends a bus, each branch is a chain fed by it."""
    ctx = G.OfflineAudioContext(48000, _record_only=True)
    buf = G.PlayableAudioBuffer.FromMonoArray(np.zeros(256, np.float32), 48000)
    s1, s2 = G.AudioBufferSourceNode(ctx), G.AudioBufferSourceNode(ctx)
    s1.Buffer = s2.Buffer = buf
    inp, dry, conv, wet, out, master = (G.GainNode(ctx), G.GainNode(ctx), G.ConvolverNode(ctx), G.GainNode(ctx), G.GainNode(ctx),
                                        G.GainNode(ctx))
    s1.Connect(inp)
    s2.Connect(inp)
    inp.Connect(dry).Connect(out)
    inp.Connect(conv).Connect(wet).Connect(out)
    out.Connect(master).Connect(ctx.Destination)
    voices, buses, dest, targets, inputs = ctx._topology_full()
    assert [(v[0], v[1], v[2], v[3]) for v in voices] == [(s1, [], 0, -1), (s2, [], 0, -1), (None, [dry], 1, 0), (None, [conv, wet], 1, 0)]
    assert buses == [[inp], [out, master]]
    assert targets == [-1, 0] and inputs == [[~0, ~1], [~2, ~3]] and dest == [1]
    # a source feeding two chains: two voices share the source; a voice chain that fans out is materialised as a bus
    ctx2 = G.OfflineAudioContext(48000, _record_only=True)
    s = G.AudioBufferSourceNode(ctx2)
    s.Buffer = buf
    a, b, pre, bus = G.GainNode(ctx2), G.GainNode(ctx2), G.BiQuadFilterNode(ctx2), G.GainNode(ctx2)
    s.Connect(pre)
    pre.Connect(a).Connect(bus)
    pre.Connect(b).Connect(bus)
    bus.Connect(ctx2.Destination)
    voices, buses, dest, targets, inputs = ctx2._topology_full()
    assert voices[0][0] is s and voices[0][1] == [pre] and voices[0][3] == -1       # source -> pre, materialised in a bus
    mat = voices[0][2]
    assert buses[mat] == [] and inputs[mat] == [~0] and targets[mat] == -1
    assert sorted(v[1][0] is a for v in voices[1:]) == [False, True] and all(v[3] == mat for v in voices[1:])


def test_unsupported_shapes_are_rejected_loudly():
    with pytest.raises(G.InvalidOperationException):  # ConvolverNode.cs:48-49
        G.ConvolverNode(G.OfflineAudioContext(48000, _record_only=True)).Buffer = G.PlayableAudioBuffer.FromMonoArray(np.ones(8, np.float32), 44100)
    with pytest.raises(G.ArgumentException):
        G.PlayableAudioBuffer.FromChannelArrays([np.zeros(4), np.zeros(5)], 48000)  # PlayableAudioBuffer.cs:130-134


def test_shard_ranges_partition_exactly():
    for n in (0, 1, 7, 64, 1024):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = sharding.shard_range(n, r, world)
                got += list(range(lo, hi))
            assert got == list(range(n))
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)
    assert math.isclose(sum(len(sharding.shard_list(list(range(10)), r, 4)) for r in range(4)), 10)


def test_occupancy_critical_kernels_stay_within_their_register_budget():
    """K6 (k_fft2_conv16<M>) is sized for 512 resident threads per SM by shared memory (8192 / M CTAs of M / 16 threads); that only
    holds while ptxas keeps it at <= 128 registers per thread.  A harmless-looking edit once took the 4096-point instantiation to 148
    registers = ONE CTA per SM and K6 of the C5 shard from 1.23 to 1.89 ms — caught only by a measurement.  cuobjdump needs no GPU."""
    import shutil
    import subprocess
    from graphaudio_b200 import _native as N
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "--dump-resource-usage", N.LIB_PATH], capture_output=True, text=True).stdout
    regs = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and name:
            regs[name] = int(m.group(1))
    k6 = {n: r for n, r in regs.items() if "k_fft2_conv16" in n}
    assert len(k6) == 4
    assert all(r <= 128 for r in k6.values()), k6


def test_parameter_epochs_are_recorded_per_render_call():
    """Edits between successive Render calls become epochs (GAC_EVENT_EPOCH markers); host logic only."""
    ctx = G.OfflineAudioContext(48000, _record_only=True)
    g = G.GainNode(ctx)
    g.Gain.Value = 0.5
    g.Gain.LinearRampToValueAtTime(1.0, 0.5)
    keep = []
    p = g.Gain._desc(keep, 0)
    assert (p.value, p.n_events) == (0.5, 1)
    g.Gain.Value = 0.25           # clears the ramp (AudioParam.cs:34-49) — but only from the next unprocessed quantum on
    p = g.Gain._desc(keep, 0)     # nothing rendered in between: the first epoch is simply replaced
    assert (p.value, p.n_events) == (0.25, 0)
    g.Gain.SetValueAtTime(0.75, 1.0)
    p = g.Gain._desc(keep, 8)     # 1000 frames rendered -> the edit acts from quantum 8
    ev = [(p.events[i].type, p.events[i].value, p.events[i].time, p.events[i].time_constant) for i in range(p.n_events)]
    assert p.value == 0.25 and ev == [(N.GAC_EVENT_EPOCH, 0.25, 0.0, 8.0), (0, 0.75, 1.0, 0.0)]
    p = g.Gain._desc(keep, 20)    # unchanged since: no new epoch
    assert p.n_events == 2
    g.Gain.Value = 0.1
    p = g.Gain._desc(keep, 20)
    assert p.n_events == 3 and p.events[2].type == N.GAC_EVENT_EPOCH and p.events[2].time_constant == 20.0 and abs(p.events[2].value - 0.1) < 1e-7
    assert ctx._block_time(3) == (128.0 / 48000 + 128.0 / 48000) + 128.0 / 48000


def test_product_package_never_imports_or_links_the_oracle():
    """oracle/ is test infrastructure: nothing under graphaudio_b200/ (or bench.py's own arm) may import, dlopen or link it."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "graphaudio_b200")
    offenders = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".inl", ".h", ".hpp")):
                continue
            for ln, line in enumerate(open(os.path.join(d, f), errors="replace"), 1):
                code = line.split("//")[0].split("#")[0] if not f.endswith(".py") else line.split("#")[0]
                if re.search(r"\b(import|from)\s+oracle\b|ga_oracle|libga_oracle|#include\s+\".*oracle", code):
                    offenders.append(f"{f}:{ln}: {line.strip()}")
    assert not offenders, offenders
    needed = subprocess.run(["ldd", N.LIB_PATH], capture_output=True, text=True).stdout
    assert "ga_oracle" not in needed


def test_param_modulation_flattens_into_a_mono_bus():
    """oscillator -> gain -> OTHER.Gain : the modulator's branch becomes a voice routed to a GAC_BUS_MONO_INPUT bus that feeds
    nothing but the parameter (AudioParam.cs:60-62: the parameter's input is Explicit, one channel)"""
    ctx = G.OfflineAudioContext(48000, _record_only=True)
    src = G.AudioBufferSourceNode(ctx)
    src.Buffer = G.PlayableAudioBuffer.FromChannelArrays([np.zeros(256, np.float32)], 48000)
    amp = G.GainNode(ctx)
    src.Connect(amp).Connect(ctx.Destination)
    src.Start()
    lfo, depth = G.OscillatorNode(ctx), G.GainNode(ctx)
    lfo.Connect(depth)
    depth.Connect(amp.Gain)
    lfo.Start()
    voices, buses, dest_inputs, targets, inputs = ctx._topology_full()
    assert len(voices) == 2 and len(buses) == 1
    assert ctx._bus_flags == [N.GAC_BUS_MONO_INPUT] and targets == [-1]
    assert amp.Gain._input_node._bus_index == 0
    lfo_voice = [v for v in voices if isinstance(v[0], G.OscillatorNode)][0]
    assert lfo_voice[1] == [depth] and lfo_voice[2] == 0
    # a modulator that is not connected anywhere leaves the graph as it was
    ctx2 = G.OfflineAudioContext(48000, _record_only=True)
    s2 = G.AudioBufferSourceNode(ctx2)
    s2.Buffer = G.PlayableAudioBuffer.FromChannelArrays([np.zeros(256, np.float32)], 48000)
    s2.Connect(ctx2.Destination)
    s2.Start()
    G.OscillatorNode(ctx2)
    assert len(ctx2._topology()[0]) == 1


def test_splitter_and_merger_flatten_into_channel_ops_and_input_slots():
    """source -> splitter ; output 0 -> gain -> merger input 1 ; output 1 -> merger input 0 (a channel swap with one gain)"""
    ctx = G.OfflineAudioContext(48000, _record_only=True)
    s = G.AudioBufferSourceNode(ctx)
    s.Buffer = G.PlayableAudioBuffer.FromChannelArrays([np.zeros(256, np.float32)] * 2, 48000)
    sp, mg, g = G.ChannelSplitterNode(ctx, 2), G.ChannelMergerNode(ctx, 2), G.GainNode(ctx)
    s.Connect(sp)
    sp.Connect(g, 0, 0)
    g.Connect(mg, 0, 1)
    sp.Connect(mg, 1, 0)
    mg.Connect(ctx.Destination)
    s.Start()
    voices, buses, dest_inputs, targets, inputs = ctx._topology_full()
    # bus 0: the splitter's input (no ops), bus 1: the merger (no ops); two chains fed by bus 0 starting with a channel pick
    assert len(buses) == 2 and buses[0] == [] and buses[1] == []
    fed = [v for v in voices if v[0] is None]
    assert len(fed) == 2 and all(v[3] == 0 and v[2] == 1 for v in fed)
    picks = sorted((v[1][0].Index, len(v[1])) for v in fed)
    assert picks == [(0, 2), (1, 1)]
    assert sorted(ctx._bus_slots[1]) == [1, 2] and ctx._bus_slots[0] is None
    with pytest.raises(G.NotSupportedException):
        c3 = G.OfflineAudioContext(48000, _record_only=True)
        m3 = G.ChannelMergerNode(c3, 4)
        x = G.GainNode(c3)
        x.Connect(m3, 0, 2)
        m3.Connect(c3.Destination)
        c3._topology()


def test_k6_segment_planner():
    """plan_segments (host-only): every overlap-save segment spends Lh = ceil16(P - 1) of its M points on history; double-length
    segments in front are used when they save work after pricing in their lower per-point rate."""
    L = N.lib()

    def plan(q, p, uniform=0):
        m, nb, ns = C.c_int(), C.c_int(), C.c_int()
        assert L.gac_plan_segments(q, p, uniform, C.byref(m), C.byref(nb), C.byref(ns)) == 0
        return m.value, nb.value, ns.value
    assert plan(4500, 750) == (2048, 1, 1)        # the bench workload: 4096 + 2048 points instead of 4 x 2048
    assert plan(4500, 750, 1) == (2048, 0, 4)
    assert plan(1300, 100) == (512, 1, 1) and plan(900, 100) == (512, 1, 0) and plan(350, 100) == (512, 0, 1)
    assert plan(3750, 1875) == (4096, 0, 2)       # C5 (512-frame partitions): no radix-16 plan beyond 4096 points
    assert plan(1875, 188) == (512, 0, 6)         # C4: six short segments stay cheaper than 1024-point ones
    assert plan(100, 10)[0] == 0                  # short impulse responses take the direct sum
    for q, p in ((4500, 750), (1300, 100), (777, 300), (20000, 1000)):
        m, nb, ns = plan(q, p)
        lh = ((p - 1 + 15) // 16) * 16
        assert nb * (2 * m - lh) + ns * (m - lh) >= q  # the segments cover every output block
    assert L.gac_plan_segments(0, 10, 0, None, None, None) != 0
