"""The numpy twin (oracle/ga_twin.py, a second reading of the C# source with another FFT and the same libm) against the C++
oracle (oracle/ga_oracle.cpp) and the committed golden fixtures.  CPU only.

Bit-equality is demanded wherever no FFT is involved (automation, source / resampler, biquad of every type, gain, fan-in);
behind the convolver two different double-precision FFTs may round a float32 spectrum value differently (1 ulp, rarely), so
those comparisons allow 2e-7 of full scale — two orders below the 1e-5 gate the device path is held to."""
import os

import numpy as np
import pytest

from oracle import ga_oracle as O
from oracle import ga_twin as T
from tests import synth

FS = 48000
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _both(build, n):
    return build(O).Render(n), build(T).Render(n)


@pytest.mark.parametrize("ftype", range(8))
def test_biquad_every_type_is_bit_equal(ftype):
    n = 128 * 24

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(60 + c, n + 128) for c in range(2)], FS)
        bq = api.BiQuadFilterNode(ctx)
        bq.Type = ftype
        bq.Frequency.SetValueAtTime(250.0, 0.0)
        bq.Frequency.ExponentialRampToValueAtTime(8000.0, 0.05)
        bq.Q.SetValueAtTime(0.7, 0.0)
        bq.Q.LinearRampToValueAtTime(3.0, 0.04)
        bq.Gain.Value = 5.0
        s.Connect(bq).Connect(ctx.Destination)
        s.Start()
        return ctx

    a, b = _both(build, n)
    assert np.abs(a).max() > 0.05
    assert np.array_equal(a, b), np.abs(a - b).max()


def test_gain_automation_curves_are_bit_equal():
    n = 128 * 40

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([np.ones(n + 128, np.float32)] * 2, FS)
        g = api.GainNode(ctx)
        g.Gain.Value = 0.3
        g.Gain.SetValueAtTime(0.9, 0.01)
        g.Gain.LinearRampToValueAtTime(0.2, 0.03)
        g.Gain.ExponentialRampToValueAtTime(0.8, 0.05)
        g.Gain.SetTargetAtTime(0.1, 0.06, 0.01)
        g.Gain.LinearRampToValueAtTime(0.5, 0.09)   # a ramp that starts from a SetTarget event (prev.Value = 0)
        s.Connect(g).Connect(ctx.Destination)
        s.Start()
        return ctx

    a, b = _both(build, n)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("rate,buf_rate", [(1.0, 44100), (0.7, 48000), (1.9, 32000)])
def test_resampled_source_is_bit_equal(rate, buf_rate):
    n = 128 * 30

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(70 + c, 3000) for c in range(2)], buf_rate)
        s.PlaybackRate.Value = rate
        s.Connect(ctx.Destination)
        s.Start(0.004)
        return ctx

    a, b = _both(build, n)
    assert np.abs(a).max() > 0.1
    assert np.array_equal(a, b), np.abs(a - b).max()


@pytest.mark.parametrize("rate,buf_rate,loop,offset,stop", [
    (1.0, 48000, (100.2, 400.2), 0.0, None),      # copy path, plays into the loop region and cycles it
    (1.0, 48000, (10.5, 50.5), 650.0, 0.05),      # start position behind LoopEnd; Stop()
    (0.5, 48000, (100.2, 400.2), 0.0, None),      # CubicResampler path through the wrap buffer
    (1.37, 44100, (0.0, 0.0), 20.0, None),        # LoopEnd 0 = end of the buffer
    (8.0, 48000, (100.2, 400.2), 0.0, None),      # cleared tails (:334-338)
    (0.9, 48000, (5.2, 8.2), 2.0, None),          # a 3-frame loop
])
def test_looping_source_is_bit_equal(rate, buf_rate, loop, offset, stop):
    """AudioBufferSourceNode.Loop on both paths (Nodes/AudioBufferSourceNode.cs:171-177, :186-235, :236-358): the numpy twin (a second
    reading of the C# source) against the C++ oracle."""
    n = 128 * 24

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(75 + c, 1000) for c in range(2)], buf_rate)
        s.Loop = True
        s.LoopStart, s.LoopEnd = loop[0] / buf_rate, loop[1] / buf_rate
        s.PlaybackRate.Value = rate
        s.Connect(ctx.Destination)
        s.Start(0.004, offset / buf_rate)
        if stop is not None:
            s.Stop(stop)
        return ctx

    a, b = _both(build, n)
    assert np.abs(a).max() > 0.1
    assert np.array_equal(a, b), np.abs(a - b).max()


def test_playback_rate_automation_is_bit_equal():
    """PlaybackRate is a k-rate parameter: its value at the start of a quantum picks the path and the phase increment (:165-186)."""
    n = 128 * 40

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(77 + c, 9000) for c in range(2)], FS)
        s.PlaybackRate.Value = 0.75
        s.PlaybackRate.SetValueAtTime(1.0, 0.02)
        s.PlaybackRate.SetValueAtTime(1.5, 0.04)
        s.PlaybackRate.LinearRampToValueAtTime(0.6, 0.08)
        s.Connect(ctx.Destination)
        s.Start()
        return ctx

    a, b = _both(build, n)
    assert np.abs(a).max() > 0.1
    assert np.array_equal(a, b), np.abs(a - b).max()


def test_partitioned_convolver_spectra_and_output():
    ir = synth.decay_ir(80, 1000)
    x = synth.splitmix_uniform(81, 128 * 30)
    po, pt = O.PartitionedConvolver(ir, 128, True), T.PartitionedConvolver(ir, 128, True)
    assert po.partitions == pt.P == 8
    assert np.float32(O.normalization_scale(ir)) == T.PartitionedConvolver.normalization_scale(ir)
    re, im = po.ir_spectra()
    assert np.abs(re - pt.ir_re).max() <= 1e-6 * np.abs(re).max() and np.abs(im - pt.ir_im).max() <= 1e-6 * np.abs(re).max()
    yo = po.process(x)
    yt = np.concatenate([pt.process(x[b * 128:(b + 1) * 128]) for b in range(30)])
    assert np.abs(yo).max() > 0.01
    assert np.abs(yo - yt).max() <= 2e-7 * max(1.0, float(np.abs(yo).max()))


@pytest.mark.parametrize("name", ["c2_small", "c3_small"])
def test_twin_renders_the_golden_graphs(name):
    """the golden graphs (tests/golden/make_golden.py) rendered by the twin: against the oracle run now and the committed fixture"""
    voices = []
    for v in range(2):
        src, ir = synth.make_voice_inputs(v if name == "c2_small" else 10 + v, 6000, 1500 if name == "c2_small" else 1000)
        voices.append((src, ir, synth.voice_gains(v)))

    def build(api):
        if name == "c2_small":
            return synth.build_c2(api, FS, voices, 0.5, t_scale=0.01)
        return synth.build_c3(api, FS, voices, 0.5, f0=300.0, f1=9000.0, t_scale=0.01, q=2.0)

    a, b = _both(build, 8000)
    ref = np.load(os.path.join(GOLDEN, name + ".npy"))
    assert np.abs(a - b).max() <= 2e-7, np.abs(a - b).max()
    assert np.abs(b - ref).max() <= 1e-6


def test_fan_in_to_a_mono_ir_convolver_mixes_down_per_input():
    n = 128 * 20

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        conv = api.ConvolverNode(ctx)
        conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.decay_ir(90, 600)], FS)
        conv.Connect(ctx.Destination)
        for v, ch in enumerate([1, 2]):
            s = api.AudioBufferSourceNode(ctx)
            s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(91 + 2 * v + c, n) for c in range(ch)], FS)
            s.Connect(conv)
            s.Start()
        return ctx

    a, b = _both(build, n)
    assert np.abs(a).max() > 0.01
    assert np.abs(a - b).max() <= 2e-7
