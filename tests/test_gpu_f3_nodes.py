"""GPU: the rest of SURVEY.md §8f-3 on the device against the CPU oracle's nodes — AudioParam modulation (node -> param,
AudioParam.cs:97-166), ConstantSourceNode, OscillatorNode, ChannelSplitterNode / ChannelMergerNode.

Tolerances: everything but the oscillator is computed in the reference's operation order (bit-equal or within the convolver's
1e-5 gate); the oscillator's phase is a blocked prefix sum in double instead of the reference's sample-by-sample sum (~1e-10 rad
apart): sine / sawtooth / triangle agree to 1e-6 of full scale, a square wave may switch ONE frame early or late at an edge, which
the test allows explicitly (at most one differing frame per edge, never two in a row)."""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu
FS = 48000


def _apis():
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    return G, O


def _both(build, n):
    G, O = _apis()
    g = build(G)
    yg = g.Render(n)
    g.Dispose()
    return yg, build(O).Render(n)


def _noise(api, ctx, stream, n, channels=2):
    s = api.AudioBufferSourceNode(ctx)
    s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(stream + c, n) for c in range(channels)], FS)
    return s


def test_constant_source_with_sample_accurate_start_and_stop():
    n = 128 * 60

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        c = api.ConstantSourceNode(ctx)
        c.Offset.SetValueAtTime(0.2, 0.0)
        c.Offset.LinearRampToValueAtTime(0.9, 0.1)
        c.Connect(ctx.Destination)
        c.Start(0.0123)      # inside a quantum
        c.Stop(0.1007)
        return ctx

    yg, yo = _both(build, n)
    assert np.abs(yo).max() > 0.5 and np.count_nonzero(yo[0]) < n
    assert np.array_equal(yg, yo)


@pytest.mark.parametrize("otype", ["Sine", "Sawtooth", "Triangle"])
def test_oscillator_with_frequency_automation(otype):
    n = 128 * 300

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        o = api.OscillatorNode(ctx)
        o.Type = getattr(api.OscillatorType, otype)
        o.Frequency.SetValueAtTime(110.0, 0.0)
        o.Frequency.ExponentialRampToValueAtTime(3520.0, 0.6)
        g = api.GainNode(ctx)
        g.Gain.Value = 0.5
        o.Connect(g).Connect(ctx.Destination)
        o.Start(0.0051)
        o.Stop(0.7003)
        return ctx

    yg, yo = _both(build, n)
    assert np.abs(yo).max() > 0.45
    assert np.abs(yg - yo).max() <= 1e-6, np.abs(yg - yo).max()
    assert np.array_equal(yg == 0, yo == 0) or np.count_nonzero((yg == 0) != (yo == 0)) <= 4  # same playing frames


def test_square_oscillator_may_move_an_edge_by_one_frame():
    n = 128 * 200

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        o = api.OscillatorNode(ctx)
        o.Type = api.OscillatorType.Square
        o.Frequency.Value = 441.3
        o.Connect(ctx.Destination)
        o.Start()
        return ctx

    yg, yo = _both(build, n)
    diff = np.flatnonzero(yg[0] != yo[0])
    assert len(diff) <= 4, len(diff)                       # in practice 0; never a run of wrong frames
    assert np.all(np.diff(diff) > 1) if len(diff) > 1 else True


def test_lfo_modulates_a_gain_parameter():
    """tremolo: noise -> GainNode whose Gain (0.5, clamped to its range) is modulated by oscillator -> depth gain"""
    n = 128 * 120

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = _noise(api, ctx, 1300, n + 128)
        amp = api.GainNode(ctx)
        amp.Gain.Value = 0.5
        s.Connect(amp).Connect(ctx.Destination)
        s.Start()
        lfo, depth = api.OscillatorNode(ctx), api.GainNode(ctx)
        lfo.Frequency.Value = 6.0
        depth.Gain.Value = 0.4
        lfo.Connect(depth)
        depth.Connect(amp.Gain)
        lfo.Start(0.02)
        lfo.Stop(0.25)      # before and after: the intrinsic value alone
        return ctx

    yg, yo = _both(build, n)
    assert np.abs(yo).max() > 0.5
    assert np.abs(yg - yo).max() <= 1e-6, np.abs(yg - yo).max()


def test_stereo_noise_modulates_a_filter_frequency_and_the_clamp_bites():
    """an audio-rate STEREO modulator into BiQuadFilterNode.Frequency: mixed to one channel as (L + R) / sqrt 2 (the parameter's
    input is Explicit 1), added to the automation curve, clamped to [1, fs / 2]; a k-rate parameter (Gain of a peaking filter) is
    modulated by a ConstantSourceNode"""
    n = 128 * 100

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = _noise(api, ctx, 1310, n + 128)
        bq = api.BiQuadFilterNode(ctx)
        bq.Type = api.FilterType.Peaking
        bq.Frequency.SetValueAtTime(300.0, 0.0)
        bq.Frequency.LinearRampToValueAtTime(4000.0, 0.2)
        bq.Q.Value = 2.0
        bq.Gain.Value = 3.0
        out = api.GainNode(ctx)
        out.Gain.Value = 0.3
        s.Connect(bq).Connect(out).Connect(ctx.Destination)
        s.Start()
        m = _noise(api, ctx, 1320, 128 * 60)
        mg = api.GainNode(ctx)
        mg.Gain.Value = 900.0     # pushes the frequency below 1 Hz now and then: the clamp matters
        m.Connect(mg)
        mg.Connect(bq.Frequency)
        m.Start()
        c = api.ConstantSourceNode(ctx)
        c.Offset.SetValueAtTime(-6.0, 0.0)
        c.Offset.LinearRampToValueAtTime(9.0, 0.2)
        c.Connect(bq.Gain)
        c.Start(0.03)
        return ctx

    yg, yo = _both(build, n)
    assert np.abs(yo).max() > 0.1
    assert np.abs(yg - yo).max() <= 1e-5, np.abs(yg - yo).max()


def test_two_modulators_into_one_parameter_and_a_modulated_delay():
    n = 128 * 100

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = _noise(api, ctx, 1330, n + 128)
        d = api.DelayNode(ctx, 0.05)
        d.DelayTime.Value = 0.01
        s.Connect(d).Connect(ctx.Destination)
        s.Start()
        for k, f in enumerate((0.9, 2.3)):     # chorus-style: two LFOs summed at the parameter's input
            lfo, depth = api.OscillatorNode(ctx), api.GainNode(ctx)
            lfo.Type = api.OscillatorType.Triangle if k else api.OscillatorType.Sine
            lfo.Frequency.Value = f * 10
            depth.Gain.Value = 0.004
            lfo.Connect(depth)
            depth.Connect(d.DelayTime)
            lfo.Start()
        return ctx

    yg, yo = _both(build, n)
    assert np.abs(yo).max() > 0.5
    # the delay is an integer number of frames: a modulator value within 1e-10 of a frame boundary could move one tap
    bad = np.count_nonzero(np.abs(yg - yo) > 1e-6)
    assert bad <= 4, bad


def test_splitter_merger_channel_swap_with_per_channel_gains():
    n = 128 * 80

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = _noise(api, ctx, 1340, n + 128)
        sp, mg = api.ChannelSplitterNode(ctx, 2), api.ChannelMergerNode(ctx, 2)
        gl, gr = api.GainNode(ctx), api.GainNode(ctx)
        gl.Gain.Value, gr.Gain.Value = 0.8, 0.3
        s.Connect(sp)
        sp.Connect(gl, 0, 0)
        sp.Connect(gr, 1, 0)
        gl.Connect(mg, 0, 1)     # left -> right
        gr.Connect(mg, 0, 0)     # right -> left
        mg.Connect(ctx.Destination)
        s.Start(0.01)
        return ctx

    yg, yo = _both(build, n)
    assert np.abs(yo).max() > 0.5 and not np.array_equal(yo[0], yo[1])
    assert np.array_equal(yg, yo)


def test_mid_side_through_splitter_merger_and_a_convolver():
    """splitter -> (L -> convolver with a mono IR, R untouched) -> merger -> destination; a mono source up-mixes at the splitter"""
    n = 128 * 100

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = _noise(api, ctx, 1350, n, channels=1)
        sp, mg = api.ChannelSplitterNode(ctx, 2), api.ChannelMergerNode(ctx, 2)
        conv = api.ConvolverNode(ctx)
        conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.decay_ir(1360, 128 * 66)], FS)
        s.Connect(sp)
        sp.Connect(conv, 0, 0)
        conv.Connect(mg, 0, 0)
        sp.Connect(mg, 1, 1)
        out = api.GainNode(ctx)
        out.Gain.Value = 0.5
        mg.Connect(out).Connect(ctx.Destination)
        s.Start()
        return ctx

    yg, yo = _both(build, n + 128 * 70)
    assert np.abs(yo).max() > 0.05
    assert np.abs(yg - yo).max() <= 1e-5, np.abs(yg - yo).max()


@pytest.mark.parametrize("loop", [False, True])
def test_modulated_playback_rate(loop):
    """A node connected to AudioBufferSourceNode.PlaybackRate (k-rate: clamp(intrinsic + the modulator's first frame of the quantum) while
    the modulator's block is non-silent, AudioParam.cs:144-158; the value picks the path and the phase increment, Nodes/
    AudioBufferSourceNode.cs:165-186).  The device evaluates the k-rate table behind the modulator's bus, the host replays the positions
    with it.  Modulators whose samples are bit-exact on both sides (a ConstantSourceNode ramp and a buffer source through a GainNode):
    the render is bit-exact."""
    n = 128 * 100

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = _noise(api, ctx, 1400, 30000)
        s.PlaybackRate.Value = 1.0
        if loop:
            s.Loop = True
            s.LoopStart, s.LoopEnd = 500.2 / FS, 9000.2 / FS
        s.Connect(ctx.Destination)
        s.Start(0.003)
        ramp = api.ConstantSourceNode(ctx)
        ramp.Offset.SetValueAtTime(0.0, 0.0)
        ramp.Offset.LinearRampToValueAtTime(0.8, 0.1)
        ramp.Offset.LinearRampToValueAtTime(-0.6, 0.2)
        ramp.Connect(s.PlaybackRate)
        ramp.Start(0.02)
        ramp.Stop(0.22)      # before and after: the intrinsic value alone (exactly 1: the copy path)
        wobble, depth = _noise(api, ctx, 1401, 4000, channels=1), api.GainNode(ctx)
        depth.Gain.Value = 0.05
        wobble.Connect(depth)
        depth.Connect(s.PlaybackRate)
        wobble.Start(0.05)
        return ctx

    yg, yo = _both(build, n)
    assert np.abs(yo).max() > 0.5
    assert np.array_equal(yg, yo), np.abs(yg - yo).max()
