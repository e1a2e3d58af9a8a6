"""GPU: graph shapes beyond source -> chain -> bus against the CPU oracle at the 1e-5 gate:
fan-out of a source, GraphAudio.Kit's ReverbEffect topology (dry / wet split of a mixed input, Effects/ReverbEffect.cs:63-91)
and AudioBus hierarchies with faded bus gains (AudioBus.cs:76-114).  The same builder drives both sides."""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu
TOL = 1e-5
FS = 48000


def _apis():
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    return G, O


def _src(api, ctx, stream, n, channels=2):
    s = api.AudioBufferSourceNode(ctx)
    s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(stream + c, n) for c in range(channels)], FS)
    s.Start()
    return s


def _reverb_effect(api, ctx, ir_stream, ir_frames, dry, wet):
    """input -> dry gain -> output ; input -> ConvolverNode -> wet gain -> output (ReverbEffect.cs:63-91)"""
    inp, out = api.GainNode(ctx), api.GainNode(ctx)
    dg, wg, conv = api.GainNode(ctx), api.GainNode(ctx), api.ConvolverNode(ctx)
    dg.Gain.Value, wg.Gain.Value = dry, wet
    conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.decay_ir(ir_stream + c, ir_frames) for c in range(2)], FS)
    inp.Connect(dg).Connect(out)
    inp.Connect(conv).Connect(wg).Connect(out)
    return inp, out


def _compare(build, n):
    G, O = _apis()
    g = build(G)
    yg = g.Render(n)
    yo = build(O).Render(n)
    assert np.abs(yo).max() > 1e-2
    assert np.abs(yg - yo).max() <= TOL, np.abs(yg - yo).max()
    g.Dispose()
    return yg


def test_source_fan_out_dry_wet():
    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = _src(api, ctx, 300, 20000)
        dry, wet, conv = api.GainNode(ctx), api.GainNode(ctx), api.ConvolverNode(ctx)
        dry.Gain.Value, wet.Gain.Value = 0.4, 0.7
        conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.decay_ir(310 + c, 128 * 70) for c in range(2)], FS)
        s.Connect(dry).Connect(ctx.Destination)
        s.Connect(conv).Connect(wet).Connect(ctx.Destination)
        return ctx
    _compare(build, 20000 + 128 * 72)


@pytest.mark.parametrize("ir_frames", [3000, 128 * 80])
def test_reverb_effect_topology(ir_frames):
    """three voices (one mono) -> effect input (fan-in) -> dry | convolver -> wet -> effect output (fan-in) -> master gain"""
    def build(api):
        ctx = api.OfflineAudioContext(FS)
        inp, out = _reverb_effect(api, ctx, 420, ir_frames, 0.6, 0.5)
        for v in range(3):
            s = _src(api, ctx, 400 + 4 * v, 16000 + 1000 * v, channels=1 if v == 1 else 2)
            g = api.GainNode(ctx)
            g.Gain.SetValueAtTime(0.5, 0.0)
            g.Gain.LinearRampToValueAtTime(0.2 + 0.1 * v, 0.3)
            s.Connect(g).Connect(inp)
        master = api.GainNode(ctx)
        master.Gain.Value = 0.5
        out.Connect(master).Connect(ctx.Destination)
        return ctx
    _compare(build, 19000 + ir_frames + 256)


def test_bus_hierarchy_with_faded_gains_and_a_filtered_send():
    """voices -> two group buses (one faded out exponentially, AudioBus.Fade) -> master bus -> destination, plus a filtered
    reverb send tapped from one group (a bus output read by a second chain)"""
    def build(api):
        ctx = api.OfflineAudioContext(FS)
        master = api.GainNode(ctx)
        master.Gain.Value = 0.5
        master.Connect(ctx.Destination)
        groups = []
        for gidx in range(2):
            grp = api.GainNode(ctx)
            grp.Gain.SetValueAtTime(0.8, 0.0)
            if gidx == 0:
                grp.Gain.ExponentialRampToValueAtTime(0.05, 0.35)
            grp.Connect(master)
            groups.append(grp)
        for v in range(5):
            s = _src(api, ctx, 500 + 4 * v, 14000 + 700 * v)
            bq = api.BiQuadFilterNode(ctx)
            bq.Type = api.FilterType.Lowpass
            bq.Frequency.SetValueAtTime(800.0 + 300 * v, 0.0)
            bq.Frequency.ExponentialRampToValueAtTime(6000.0, 0.3)
            s.Connect(bq).Connect(groups[v % 2])
        # reverb send from group 1: group output -> highpass -> convolver -> send gain -> master
        hp, conv, send = api.BiQuadFilterNode(ctx), api.ConvolverNode(ctx), api.GainNode(ctx)
        hp.Type = api.FilterType.Highpass
        hp.Frequency.Value = 300.0
        conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.decay_ir(560 + c, 128 * 66) for c in range(2)], FS)
        send.Gain.Value = 0.6
        groups[1].Connect(hp).Connect(conv).Connect(send).Connect(master)
        return ctx
    _compare(build, 18000 + 128 * 70)
