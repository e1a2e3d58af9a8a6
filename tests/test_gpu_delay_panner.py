"""GPU: DelayNode and StereoPannerNode inside rendered graphs, against the CPU oracle.

  DelayNode (Nodes/DelayNode.cs:43-149): out[n] = d >= 1 ? x[n - d] : 0 with d = clamp((int)(delayTime[n] * fs), 0, max); one pooled
  output block that is only ever marked non-silent (:96-97), which matters to a BiQuadFilterNode behind it (it freezes on silent input).
  StereoPannerNode (Nodes/StereoPannerNode.cs:36-153): equal-power gains from MathF.Cos / MathF.Sin of the clamped pan, mono and
  stereo variants.  Both are bit-exact on the device (gather / libm-identical sincos, unfused arithmetic).
"""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu
FS = 48000
TOL = 1e-5


def _apis():
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    return G, O


def _pair(build, n):
    G, O = _apis()
    return build(G).Render(n), build(O).Render(n)


@pytest.mark.parametrize("src_ch", [1, 2])
@pytest.mark.parametrize("delay", [0.0, 1.0 / FS, 0.01, 0.123, 0.25, 0.4])
def test_constant_delay_is_bit_exact(src_ch, delay):
    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(100 + c, 9000) for c in range(src_ch)], FS)
        d = api.DelayNode(ctx, 0.25)  # 0.4 s is clamped to the maximum by AudioParam.Value
        d.DelayTime.Value = delay
        s.Connect(d).Connect(ctx.Destination)
        s.Start(0.05)
        return ctx
    yg, yo = _pair(build, 36000)
    assert (np.abs(yo).max() > 0.1) == (delay > 0)
    assert np.array_equal(yg, yo)


def test_automated_delay_time_sweep():
    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(110 + c, 20000) for c in range(2)], FS)
        d = api.DelayNode(ctx, 0.1)
        d.DelayTime.SetValueAtTime(0.001, 0.0)
        d.DelayTime.LinearRampToValueAtTime(0.08, 0.2)
        d.DelayTime.SetTargetAtTime(0.0, 0.3, 0.05)  # decays through d = 0 (reads nothing)
        s.Connect(d).Connect(ctx.Destination)
        s.Start()
        return ctx
    yg, yo = _pair(build, 30000)
    assert np.abs(yo).max() > 0.1
    assert np.array_equal(yg, yo)


def test_delay_flag_semantics_reach_a_biquad_behind_it():
    # the biquad freezes while its input is flagged silent (BiQuadFilterNode.cs:103-108): before the first delayed sample arrives,
    # and never again afterwards (the delay's block stays flagged once marked, DelayNode.cs:96-97) — so the filter rings out after
    # the source has ended
    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(120 + c, 6000) for c in range(2)], FS)
        d = api.DelayNode(ctx, 0.5)
        d.DelayTime.Value = 0.0301
        f = api.BiQuadFilterNode(ctx)
        f.Frequency.Value = 300.0
        f.Q.Value = 8.0
        g = api.GainNode(ctx)
        g.Gain.Value = 0.25
        s.Connect(d).Connect(f).Connect(g).Connect(ctx.Destination)
        s.Start()
        return ctx
    yg, yo = _pair(build, 16000)
    assert np.abs(yo[:, 9000:]).max() > 1e-6  # the resonance is still ringing after the delayed source has ended
    assert np.abs(yg - yo).max() <= TOL


@pytest.mark.parametrize("start", [0.0, 0.01])
@pytest.mark.parametrize("src_ch", [1, 2])
@pytest.mark.parametrize("pan", [-1.0, -0.3, 0.0, 0.5, 1.0])
def test_constant_pan_is_bit_exact(src_ch, pan, start):
    # covers the odd first quantum (mono up-mixed at block 0, late stereo source mixed down) and the gain pair it leaves cached
    # (tests/test_oracle_kats.py::test_stereo_panner_equal_power_formulas spells both out)
    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(130 + c, 6000) for c in range(src_ch)], FS)
        p = api.StereoPannerNode(ctx)
        p.Pan.Value = pan
        s.Connect(p).Connect(ctx.Destination)
        s.Start(start)
        return ctx
    yg, yo = _pair(build, 7000)
    assert np.abs(yo).max() > 0.1
    assert np.array_equal(yg, yo)


@pytest.mark.parametrize("start", [0.0, 0.02])
@pytest.mark.parametrize("src_ch", [1, 2])
def test_stepped_pan_keeps_the_cached_pair_until_the_first_change(src_ch, start):
    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(135 + c, 12000) for c in range(src_ch)], FS)
        p = api.StereoPannerNode(ctx)
        p.Pan.SetValueAtTime(-0.4, 0.0)
        p.Pan.SetValueAtTime(0.6, 0.1)    # first change at frame 4800
        p.Pan.SetValueAtTime(-0.4, 0.15)  # back to the first value: recomputed by the variant running then
        s.Connect(p).Connect(ctx.Destination)
        s.Start(start)
        return ctx
    yg, yo = _pair(build, 14000)
    assert np.array_equal(yg, yo)


@pytest.mark.parametrize("start", [0.0, 0.02])
@pytest.mark.parametrize("src_ch", [1, 2])
def test_pan_sweep_with_every_sample_a_new_gain_pair(src_ch, start):
    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(140 + c, 30000) for c in range(src_ch)], FS)
        p = api.StereoPannerNode(ctx)
        p.Pan.SetValueAtTime(-1.0, 0.0)
        p.Pan.LinearRampToValueAtTime(1.0, 0.5)  # 24 000 distinct pan values through both branches
        s.Connect(p).Connect(ctx.Destination)
        s.Start(start)
        return ctx
    yg, yo = _pair(build, 32000)
    assert np.array_equal(yg, yo)


def test_voices_with_delay_and_pan_into_a_convolver_bus():
    # a small "positioned voices -> reverb send" graph: source -> gain -> panner -> bus; bus -> delay (pre-delay) -> convolver
    def build(api):
        ctx = api.OfflineAudioContext(FS)
        bus = api.GainNode(ctx)
        bus.Gain.Value = 0.5
        for v in range(4):
            s = api.AudioBufferSourceNode(ctx)
            s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(150 + 2 * v + c, 8000) for c in range(1 + v % 2)], FS)
            g = api.GainNode(ctx)
            g.Gain.SetValueAtTime(0.2, 0.0)
            g.Gain.LinearRampToValueAtTime(0.8, 0.1)
            p = api.StereoPannerNode(ctx)
            p.Pan.Value = -0.75 + 0.5 * v
            s.Connect(g).Connect(p).Connect(bus)
            s.Start(0.01 * v)
        pre = api.DelayNode(ctx, 0.1)
        pre.DelayTime.Value = 0.02
        conv = api.ConvolverNode(ctx)
        conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.decay_ir(160 + c, 6000) for c in range(2)], FS)
        bus.Connect(pre).Connect(conv).Connect(ctx.Destination)
        bus.Connect(ctx.Destination)  # dry path
        return ctx
    yg, yo = _pair(build, 20000)
    assert np.abs(yo).max() > 1e-2
    assert np.abs(yg - yo).max() <= TOL * max(1.0, np.abs(yo).max())


def test_max_delay_time_is_validated():
    G, _ = _apis()
    ctx = G.OfflineAudioContext(FS)
    with pytest.raises(G.ArgumentOutOfRangeException):
        G.DelayNode(ctx, 0.0)
    with pytest.raises(G.ArgumentOutOfRangeException):
        G.DelayNode(ctx, 10.5)


def test_f3_graph_against_the_committed_golden_fixture():
    """The graph of tests/golden/f3_small.npy (looping mono source -> pan sweep; late stereo source -> delay sweep -> pan -> biquad;
    convolver bus), rendered on the device and compared with the committed fixture — no oracle call in this test."""
    import os
    import graphaudio_b200 as G
    from tests import test_oracle_kats as T
    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "f3_small.npy"))
    y = T._golden_f3(G)
    assert y.shape == ref.shape and np.abs(ref).max() > 1e-2
    assert np.abs(y - ref).max() <= TOL
