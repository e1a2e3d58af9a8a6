"""CPU tests: the oracle (oracle/ga_oracle.cpp) against known answers derived from the reference source
and against independent math (numpy / scipy).  The reference ships no tests or golden vectors of its own, so these
known-answer tests — each citing the reference lines it pins — are what anchors the oracle ("parity unpinned" otherwise).
The committed fixtures under tests/golden/ freeze the oracle's output so that later edits cannot drift silently.
"""
import math
import os

import numpy as np
import pytest
from scipy.signal import fftconvolve, lfilter

from oracle import ga_oracle as O
from tests import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ------------------------------------------------------------------ FFT conventions (FftFlat/RealFourierTransform.cs:62-131)
@pytest.mark.parametrize("n", [2, 8, 256, 1024])
def test_rfft_matches_numpy_convention(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    X = O.rfft_forward(x)
    assert np.abs(X - np.fft.rfft(x)).max() < 1e-12 * n  # Forward == numpy.fft.rfft (Math.NET sign convention :80-84)
    assert X[0].imag == 0 and X[-1].imag == 0            # DC and Nyquist purely real (:76-78)
    assert np.abs(O.rfft_inverse(X) - x).max() < 1e-13   # Inverse is the exact inverse (2/N scaling :46,129)


def test_rfft_rejects_non_power_of_two():
    with pytest.raises(ValueError):
        O.rfft_forward(np.zeros(12))  # RealFourierTransform.cs:38-41


# ------------------------------------------------------------------ PartitionedConvolver (PartitionedConvolver.cs)
def test_delta_ir_is_a_pure_delay():
    """Normalize=false, IR = delta[n-d] => output = input delayed by d (ring direction + overlap-add, :104-152)."""
    x = synth.splitmix_uniform(1, 128 * 30)
    d = 200
    ir = np.zeros(300, np.float32)
    ir[d] = 1.0
    y = O.PartitionedConvolver(ir, 128, False).process(x)
    assert np.abs(y[d:] - x[:-d]).max() <= 2.5e-7
    assert np.abs(y[:d]).max() <= 1e-7


@pytest.mark.parametrize("block,ir_len", [(128, 24000), (128, 129), (512, 5000), (128, 1)])
def test_equals_linear_convolution(block, ir_len):
    n = block * 60
    x = synth.splitmix_uniform(2, n)
    ir = synth.decay_ir(3, ir_len) if ir_len > 1 else np.array([0.5], np.float32)
    pc = O.PartitionedConvolver(ir, block, True)
    y = pc.process(x)
    scale = np.float32(O.normalization_scale(ir))
    ref = fftconvolve(x.astype(np.float64), (ir * scale).astype(np.float64))[:n]
    assert np.abs(y - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())
    assert pc.partitions == math.ceil(ir_len / block)  # :44


def test_normalization_scale_known_answers():
    # all-ones IR: rms = 1 => scale = (1/1) * (float)10^(-58 * 0.05f)   (:93-102)
    s = O.normalization_scale(np.ones(1000, np.float32))
    assert s == np.float32(10.0 ** float(np.float32(-58) * np.float32(0.05)))
    assert abs(s - 1.2589e-3) < 1e-6
    # rms below MinPower (1.25e-4) is clamped
    s2 = O.normalization_scale(np.full(100, 1e-6, np.float32))
    assert s2 == np.float32(np.float32(1.0) / np.float32(0.000125)) * np.float32(s)
    # each channel is normalised by its OWN rms (ConvolverNode.cs:51-56): two channels, different scales
    a, b = synth.decay_ir(1, 1000), synth.decay_ir(2, 1000) * np.float32(0.1)
    assert O.normalization_scale(a) != O.normalization_scale(b)


def test_ir_spectra_are_float32_of_double_fft():
    ir = synth.decay_ir(5, 300)
    pc = O.PartitionedConvolver(ir, 128, False)
    re, im = pc.ir_spectra()
    blk = np.zeros(256)
    blk[:300 - 256] = ir[256:]
    ref = np.fft.rfft(blk)
    assert np.array_equal(re[2], ref.real.astype(np.float32)) or np.abs(re[2] - ref.real).max() < 1e-7
    assert re.shape == (3, 129)  # P = ceil(300/128), C = B + 1 (:41,44)


# ------------------------------------------------------------------ CubicResampler (CubicResampler.cs:26-63)
def test_resampler_known_answers():
    x = np.arange(100, dtype=np.float32)
    y, consumed = O.resample(x, 50, 1.0)
    assert y[0] == x[1]                       # primed by 4 samples, first output is in[1] (:31-35,51-52)
    assert np.array_equal(y[:20], x[1:21])    # rate 1, t = 0: passes S1 through
    y2, _ = O.resample(x, 150, 0.5)
    assert np.allclose(y2[:100], 1.0 + 0.5 * np.arange(100), atol=1e-5)  # Catmull-Rom reproduces a linear ramp
    # output count for L inputs at rate r: stops when the next consume would pass the end (:43-44)
    y3, c3 = O.resample(x[:10], 1000, 0.459375)
    assert c3 == 10 and y3.shape[0] == int(np.ceil((10 - 4 + 1) / 0.459375))
    y4, c4 = O.resample(x[:3], 10, 0.5)
    assert y4.shape[0] == 0 and c4 == 3       # fewer than 4 inputs: never primed (:37-38)


# ------------------------------------------------------------------ AudioParam (AudioParam.cs:169-247)
def _gain_param():
    ctx = O.OfflineAudioContext(48000)
    return ctx, O.GainNode(ctx).Gain


def test_automation_schedule():
    ctx, p = _gain_param()
    p.SetValueAtTime(0.5, 0.0)
    p.LinearRampToValueAtTime(1.0, 0.01)
    p.ExponentialRampToValueAtTime(0.25, 0.02)
    p.SetTargetAtTime(0.0, 0.02, 0.005)
    v = p.evaluate(12)
    t = np.zeros(12 * 128)
    bt = 0.0
    for b in range(12):  # blockTime accumulates by repeated addition (AudioContextBase.cs:78-79)
        t[b * 128:(b + 1) * 128] = bt + np.arange(128) * (1.0 / 48000)
        bt = bt + 128.0 / 48000
    ref = np.where(t < 0.01, 0.5 + 0.5 * np.clip(t / 0.01, 0, 1),
                   np.where(t < 0.02, 1.0 * 0.25 ** np.clip((t - 0.01) / 0.01, 0, 1), 0.25 * np.exp(-(t - 0.02) / 0.005)))
    assert np.abs(v - ref).max() < 1e-6


def test_first_event_ramp_is_a_step():
    ctx, p = _gain_param()
    p.Value = 0.25
    p.LinearRampToValueAtTime(1.0, 0.01)  # first event is a ramp: static value until its end time (:181-184)
    v = p.evaluate(6)
    k = int(np.ceil(0.01 * 48000))
    assert np.all(v[:k] == np.float32(0.25)) and np.all(v[k + 1:] == np.float32(1.0))


def test_ramp_after_set_target_starts_from_zero_value_field():
    ctx, p = _gain_param()
    p.SetValueAtTime(0.8, 0.0)
    p.SetTargetAtTime(0.2, 0.001, 0.01)
    p.LinearRampToValueAtTime(1.0, 0.004)  # interpolates from the SetTarget event's Value field, which is 0 (:186-190)
    v = p.evaluate(2)
    i = 100  # t = 100/48000 = 2.08 ms, inside the ramp segment [1 ms, 4 ms]
    u = (i / 48000 - 0.001) / 0.003
    assert abs(v[i] - u * 1.0) < 1e-6


def test_value_setter_cancels_events_and_clamps():
    ctx = O.OfflineAudioContext(48000)
    bq = O.BiQuadFilterNode(ctx)
    bq.Frequency.SetValueAtTime(500.0, 0.0)
    bq.Frequency.Value = 1e9  # clamped to fs/2 and clears the schedule (:34-49)
    assert np.all(bq.Frequency.evaluate(1) == np.float32(24000.0))
    bq.Frequency.ExponentialRampToValueAtTime(-1.0, 1.0)  # clamped to the minimum (1 Hz) BEFORE the > 0 check (:282-284)
    with pytest.raises(O.ArgumentException):
        O.GainNode(ctx).Gain.ExponentialRampToValueAtTime(-1.0, 1.0)  # "Exponential ramp target must be > 0"


# ------------------------------------------------------------------ graph-level semantics
def _source_to_dest(x, fs=48000, ops=None, n=None, stereo=False):
    ctx = O.OfflineAudioContext(fs)
    s = O.AudioBufferSourceNode(ctx)
    s.Buffer = O.PlayableAudioBuffer.FromStereoArrays(x, x[::-1].copy(), fs) if stereo else O.PlayableAudioBuffer.FromMonoArray(x, fs)
    node = s
    made = []
    for mk in ops or []:
        nd = mk(ctx)
        node = node.Connect(nd)
        made.append(nd)
    node.Connect(ctx.Destination)
    s.Start()
    return ctx, ctx.Render(n or len(x)), made


@pytest.mark.parametrize("L", [1280, 1281, 1300, 1407, 1408, 100, 129])
def test_source_end_rule(L):
    """rate == 1, buffer length L => exactly 128*floor((L-1)/128) frames are emitted (AudioBufferSourceNode.cs:224,360-362)."""
    x = synth.splitmix_uniform(4, L) + np.float32(2.0)  # never zero
    _, y, _ = _source_to_dest(x, n=L + 256)
    emitted = 128 * ((L - 1) // 128)
    assert np.array_equal(y[0, :emitted], x[:emitted])
    assert not y[:, emitted:].any()
    assert np.array_equal(y[0], y[1])  # mono -> stereo up-mix copies the channel (AudioNodeInput.cs:201-213)


def test_chunked_render_equals_single_render():
    x = synth.splitmix_uniform(6, 5000)
    ir = synth.decay_ir(7, 700)

    def build():
        ctx = O.OfflineAudioContext(48000)
        s = O.AudioBufferSourceNode(ctx)
        s.Buffer = O.PlayableAudioBuffer.FromMonoArray(x, 48000)
        c = O.ConvolverNode(ctx)
        c.Buffer = O.PlayableAudioBuffer.FromStereoArrays(ir, ir[::-1].copy(), 48000)
        s.Connect(c).Connect(ctx.Destination)
        s.Start()
        return ctx
    whole = build().Render(3000)
    c2 = build()
    parts = np.concatenate([c2.Render(1000), c2.Render(777), c2.Render(1223)], axis=1)  # OfflineAudioContext.cs:55-100
    assert np.array_equal(whole, parts)


def test_biquad_constant_matches_rbj_closed_form():
    """f = 1000, Q = 1 (the defaults): one coefficient set, closed-form RBJ lowpass; checked against scipy.lfilter."""
    fs = 48000
    x = synth.splitmix_uniform(8, 128 * 40)
    _, y, _ = _source_to_dest(x, ops=[lambda c: O.BiQuadFilterNode(c)], n=128 * 40)
    w0 = 2 * np.pi * 1000.0 / fs
    alpha = np.sin(w0) / 2.0
    b = np.array([(1 - np.cos(w0)) / 2, 1 - np.cos(w0), (1 - np.cos(w0)) / 2]) / (1 + alpha)
    a = np.array([1.0, -2 * np.cos(w0) / (1 + alpha), (1 - alpha) / (1 + alpha)])
    emitted = 128 * 39
    ref = lfilter(b, a, x[:emitted].astype(np.float64))
    assert np.abs(y[0, :emitted] - ref).max() < 2e-5  # float32 DF-II vs float64 reference


def test_biquad_nondefault_constant_and_silent_tail():
    fs = 48000
    x = synth.splitmix_uniform(9, 128 * 20)

    def mk(c):
        b = O.BiQuadFilterNode(c)
        b.Type = O.FilterType.Highpass
        b.Frequency.Value = 3000.0
        b.Q.Value = 2.0
        return b
    _, y, _ = _source_to_dest(x, ops=[mk], n=128 * 24)
    w0 = 2 * np.pi * 3000.0 / fs
    alpha = np.sin(w0) / 4.0
    b = np.array([(1 + np.cos(w0)) / 2, -(1 + np.cos(w0)), (1 + np.cos(w0)) / 2]) / (1 + alpha)
    a = np.array([1.0, -2 * np.cos(w0) / (1 + alpha), (1 - alpha) / (1 + alpha)])
    emitted = 128 * 19
    ref = lfilter(b, a, x[:emitted].astype(np.float64))
    assert np.abs(y[0, :emitted] - ref).max() < 2e-5
    # once the source block is flagged silent the filter emits zeros: the tail is CUT, not rung out (BiQuadFilterNode.cs:103-108)
    assert not y[:, emitted:].any()


def test_fan_in_order_is_connection_order():
    fs = 48000
    a = np.full(256, 1e8, np.float32)
    b = np.full(256, -1e8, np.float32)
    c = np.full(256, 1.0, np.float32)

    def run(order):
        ctx = O.OfflineAudioContext(fs)
        for arr in order:
            s = O.AudioBufferSourceNode(ctx)
            s.Buffer = O.PlayableAudioBuffer.FromMonoArray(arr, fs)
            s.Connect(ctx.Destination)
            s.Start()
        return ctx.Render(128)
    assert np.all(run([a, b, c])[0] == 1.0)   # ((0 + 1e8) - 1e8) + 1 = 1     (AudioNodeInput.cs:118-137)
    assert np.all(run([a, c, b])[0] == 0.0)   # ((0 + 1e8) + 1) - 1e8 = 0 in float32


def test_convolver_rate_mismatch_raises():
    ctx = O.OfflineAudioContext(48000)
    c = O.ConvolverNode(ctx)
    with pytest.raises(O.InvalidOperationException):  # ConvolverNode.cs:48-49
        c.Buffer = O.PlayableAudioBuffer.FromMonoArray(np.ones(10, np.float32), 44100)


# ------------------------------------------------------------------ committed golden fixtures
def _golden_c2():
    fs = 48000
    voices = []
    for v in range(2):
        src, ir = synth.make_voice_inputs(v, 6000, 1500)
        voices.append((src, ir, synth.voice_gains(v)))
    return synth.build_c2(O, fs, voices, 0.5, t_scale=0.01).Render(8000)


def _golden_c3():
    fs = 48000
    voices = []
    for v in range(2):
        src, ir = synth.make_voice_inputs(10 + v, 6000, 1000)
        voices.append((src, ir, synth.voice_gains(v)))
    return synth.build_c3(O, fs, voices, 0.5, f0=300.0, f1=9000.0, t_scale=0.01, q=2.0).Render(8000)


def _golden_f3(api=None):
    # looping mono source -> panner sweep (first quantum up-mixed) ; late stereo source -> delay sweep -> panner -> biquad ; both into a convolver bus
    api = api or O
    fs = 48000
    ctx = api.OfflineAudioContext(fs)
    bus = api.GainNode(ctx)
    bus.Gain.Value = 0.5
    a = api.AudioBufferSourceNode(ctx)
    a.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(20, 1000)], fs)
    a.Loop = True
    a.LoopStart, a.LoopEnd = 100.2 / fs, 700.2 / fs
    pa = api.StereoPannerNode(ctx)
    pa.Pan.SetValueAtTime(-0.8, 0.0)
    pa.Pan.LinearRampToValueAtTime(0.8, 0.1)
    a.Connect(pa).Connect(bus)
    a.Start()
    b = api.AudioBufferSourceNode(ctx)
    b.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(21 + c, 5000) for c in range(2)], fs)
    d = api.DelayNode(ctx, 0.05)
    d.DelayTime.SetValueAtTime(0.002, 0.0)
    d.DelayTime.LinearRampToValueAtTime(0.04, 0.12)
    pb = api.StereoPannerNode(ctx)
    pb.Pan.Value = 0.3
    f = api.BiQuadFilterNode(ctx)
    f.Frequency.Value = 1200.0
    b.Connect(d).Connect(pb).Connect(f).Connect(bus)
    b.Start(0.01)
    conv = api.ConvolverNode(ctx)
    conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.decay_ir(23 + c, 1200) for c in range(2)], fs)
    bus.Connect(conv).Connect(ctx.Destination)
    return ctx.Render(8000)


@pytest.mark.parametrize("name,fn", [("c2_small", _golden_c2), ("c3_small", _golden_c3), ("f3_small", _golden_f3)])
def test_golden_fixture(name, fn):
    """tests/golden/*.npy were generated by tests/golden/make_golden.py from this oracle; bit-exact on the same libm."""
    path = os.path.join(GOLDEN, name + ".npy")
    y = fn()
    ref = np.load(path)
    assert y.shape == ref.shape
    # identical on the generating machine; other glibc builds may differ in the last ulp of sinf/cosf/pow
    assert np.abs(y - ref).max() <= 1e-6


# ---------------------------------------------------------------- DelayNode / StereoPannerNode
def _render_chain(make_node, src, n, fs=48000, start=0.0):
    ctx = O.OfflineAudioContext(fs)
    s = O.AudioBufferSourceNode(ctx)
    s.Buffer = O.PlayableAudioBuffer.FromChannelArrays(src, fs)
    node = make_node(ctx)
    s.Connect(node).Connect(ctx.Destination)
    s.Start(start)
    return ctx.Render(n)


def test_delay_node_is_an_integer_shift_and_zero_delay_reads_nothing():
    # DelayNode.cs:68-72: d = clamp((int)(delayTime * fs), 0, max); CircularBuffer.Read returns 0 for d <= 0 (:140-143) and is read
    # before the write, so out[n] = x[n - d] for d >= 1 and silence for d = 0
    x = [synth.splitmix_uniform(700 + c, 128 * 20) for c in range(2)]
    for delay, d in ((0.0, 0), (1.0 / 48000, 1), (0.01, 480), (0.0301, 1444), (0.05, 2400), (0.2, 2400)):  # 0.2 s clamps to the 0.05 s maximum
        def make(ctx):
            node = O.DelayNode(ctx, 0.05)
            node.DelayTime.Value = delay
            return node
        y = _render_chain(make, x, 128 * 40)
        emitted = 128 * 19  # the source drops its final block (AudioBufferSourceNode.cs:360-368)
        want = np.zeros((2, 128 * 40), np.float32)
        if d >= 1:
            for c in range(2):
                want[c, d:d + emitted] = x[c][:emitted]
        assert np.array_equal(y, want), delay


def test_stereo_panner_equal_power_formulas():
    # StereoPannerNode.cs:77-152 with MathF.Cos / MathF.Sin = libm cosf / sinf on x * pi / 2 formed in float32.
    # Two properties of the reference that a restatement has to keep:
    #  (1) the node's input is ClampedMax(2) and AudioNodeInput.Pull computes its channel count from the upstream block of the
    #      PREVIOUS quantum (AudioNodeInput.cs:109 precedes :124): in the very first quantum there is none and the input has 2
    #      channels (a mono source is up-mixed by copy and runs through ProcessStereo); a stereo source that starts later is mixed
    #      down to one channel, (L + R) / sqrt(2), in its first quantum (the idle source block before it had one channel) and runs
    #      through ProcessMono;
    #  (2) the gain pair is recomputed only when the pan value changes (:95-103, :129-137), by whichever variant is running: with a
    #      constant pan the pair of that odd first quantum stays in force for the rest of the render.
    pi = np.float32(np.pi)
    mono = [synth.splitmix_uniform(710, 128 * 10)]
    stereo = [synth.splitmix_uniform(711 + c, 128 * 10) for c in range(2)]
    n = 128 * 9

    def gains(x):
        a = np.float32(np.float32(x) * pi) / np.float32(2.0)
        # float32-rounded cos / sin of a float32 argument (the double evaluation rounds identically for these pan values)
        return np.float32(np.cos(np.float64(a))), np.float32(np.sin(np.float64(a)))

    def g_mono(p):
        return gains((p + np.float32(1.0)) * np.float32(0.5))

    def g_stereo(p):
        return gains(p + np.float32(1.0) if p <= 0 else p)

    def run_mono(x, g):
        return np.stack([x * g[0], x * g[1]])

    def run_stereo(L, R, p, g):
        return np.stack([L + R * g[0], R * g[1]]) if p <= 0 else np.stack([L * g[0], R + L * g[1]])

    for pan in (-1.0, -0.5, 0.0, 0.25, 1.0):
        def make(ctx):
            node = O.StereoPannerNode(ctx)
            node.Pan.Value = pan
            return node
        p = np.float32(pan)
        ym = _render_chain(make, mono, n)
        x = mono[0][:n]
        assert np.array_equal(ym[:, :128], run_stereo(x[:128], x[:128], p, g_stereo(p)))  # first quantum: two equal channels
        assert np.array_equal(ym[:, 128:], run_mono(x[128:], g_stereo(p)))                # ... whose gain pair stays cached
        ys = _render_chain(make, stereo, n)
        assert np.array_equal(ys, run_stereo(stereo[0][:n], stereo[1][:n], p, g_stereo(p)))
        # the same stereo source started in the third quantum: its first quantum arrives mixed down
        yl = _render_chain(make, stereo, n + 256, start=256.0 / 48000)
        assert not yl[:, :256].any()
        down = (stereo[0][:128] + stereo[1][:128]) * (np.float32(1.0) / np.sqrt(np.float32(2.0)))
        assert np.array_equal(yl[:, 256:384], run_mono(down, g_mono(p)))
        assert np.array_equal(yl[:, 384:], run_stereo(stereo[0][128:n], stereo[1][128:n], p, g_mono(p)))
    # full left keeps the left channel and adds the right one at unity; full right mirrors it
    assert gains(np.float32(0.0)) == (np.float32(1.0), np.float32(0.0))

    # a pan that changes with every sample is recomputed with the running variant's formula: mono source, quanta >= 1
    def make_sweep(ctx):
        node = O.StereoPannerNode(ctx)
        node.Pan.SetValueAtTime(-0.9, 0.0)
        node.Pan.LinearRampToValueAtTime(0.9, 0.02)  # 960 samples
        return node
    ym = _render_chain(make_sweep, mono, 128 * 7)
    ctx = O.OfflineAudioContext(48000)
    probe = O.StereoPannerNode(ctx)
    probe.Pan.SetValueAtTime(-0.9, 0.0)
    probe.Pan.LinearRampToValueAtTime(0.9, 0.02)
    pv = probe.Pan.evaluate(7)
    for i in (128, 300, 700, 890):
        g = g_mono(np.float32(pv[i]))
        assert ym[0, i] == mono[0][i] * g[0] and ym[1, i] == mono[0][i] * g[1]


def test_looping_source_copy_path_closed_form():
    # AudioBufferSourceNode.cs:171-177, :186-235: emitted frame j reads buffer frame k = pos0 + j while k < loopEnd, then
    # loopStart + (k - loopEnd) % loopLength; a start position at or behind loopEnd restarts the FIRST quantum at loopStart (:197-200)
    # while _playbackPosition still advances from the original position (:224-234)
    fs = 48000
    x = np.arange(1, 1001, dtype=np.float32)

    def run(loop_start, loop_end, offset, n=128 * 12, stop=None):
        ctx = O.OfflineAudioContext(fs)
        s = O.AudioBufferSourceNode(ctx)
        s.Buffer = O.PlayableAudioBuffer.FromChannelArrays([x], fs)
        s.Loop = True
        s.LoopStart, s.LoopEnd = loop_start / fs, loop_end / fs
        s.Connect(ctx.Destination)
        s.Start(0.0, offset / fs)
        if stop is not None:
            s.Stop(stop)
        return ctx.Render(n)[0]

    k = np.arange(128 * 12)
    assert np.array_equal(run(100.2, 400.2, 0), x[np.where(k < 400, k, 100 + (k - 400) % 300)])
    assert np.array_equal(run(10.5, 50.5, 20), x[np.where(k + 20 < 50, k + 20, 10 + (k + 20 - 50) % 40)])
    late = run(100.2, 400.2, 650)
    assert np.array_equal(late[:128], x[100 + np.arange(128) % 300])
    assert np.array_equal(late[128:], x[100 + (650 + k[128:] - 400) % 300])
    whole = run(0, 0, 0, n=128 * 20)
    assert np.array_equal(whole, x[np.arange(128 * 20) % 1000])  # LoopEnd 0 = end of the buffer; never ends by itself
    stopped = run(0, 0, 0, stop=0.01)  # block-granular stop: quanta with t0 < 0.01 s play (:137-143)
    assert np.nonzero(stopped)[0][-1] == 511


def _looping_resampled(x, fs_buf, fs, rate, loop_start, loop_end, offset, n, stop=None, api=O):
    ctx = api.OfflineAudioContext(fs)
    s = api.AudioBufferSourceNode(ctx)
    s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([x], fs_buf)
    s.Loop = True
    s.LoopStart, s.LoopEnd = loop_start / fs_buf, loop_end / fs_buf
    s.PlaybackRate.Value = rate
    s.Connect(ctx.Destination)
    s.Start(0.0, offset / fs_buf)
    if stop is not None:
        s.Stop(stop)
    return ctx.Render(n)[0]


@pytest.mark.parametrize("rate", [0.5, 0.73, 1.37, 2.0, 3.9])
def test_looping_resampler_equals_the_resampler_over_the_unrolled_loop(rate):
    # Nodes/AudioBufferSourceNode.cs:236-358 with Loop: the wrap buffer (:296-314) presents pos .. loopEnd-1 followed by the loop
    # region, so the CubicResampler sees ONE continuous stream — the same samples, phases and float32 polynomial as a non-looping
    # source playing the unrolled buffer.  (Rates below ~5 never hit the cleared-tail branch :334-338.)
    fs = 48000
    x = synth.splitmix_uniform(901, 1000)
    n = 128 * 10
    k = np.arange(int(n * rate) + 600)
    unrolled = x[np.where(k < 400, k, 100 + (k - 400) % 300)]
    got = _looping_resampled(x, fs, fs, rate, 100.2, 400.2, 0, n)
    ctx = O.OfflineAudioContext(fs)
    s = O.AudioBufferSourceNode(ctx)
    s.Buffer = O.PlayableAudioBuffer.FromChannelArrays([unrolled], fs)
    s.PlaybackRate.Value = rate
    s.Connect(ctx.Destination)
    s.Start(0.0)
    want = ctx.Render(n)[0]
    assert np.count_nonzero(want) > n - 8
    assert np.array_equal(got, want)


@pytest.mark.parametrize("rate,loop,offset,fs_buf", [
    (0.5, (100.2, 400.2), 0, 48000),       # below 1: one Process call per quantum
    (1.37, (100.2, 400.2), 650, 48000),    # start position behind loopEnd: the first call restarts at loopStart (:265-268)
    (1.0, (10.5, 47.5), 20, 44100),        # 44.1 -> 48 kHz, a 37-frame loop: ONE pass over the loop region per call (:308-311)
    (2.5, (0, 0), 0, 48000),               # LoopEnd 0 = end of the buffer
    (8.0, (100.2, 400.2), 0, 48000),       # the tail of every quantum is cleared once fewer than (int)Pos inputs are offered (:334-338)
    (0.9, (5.2, 8.2), 2, 48000),           # a 3-frame loop: priming needs two calls (CubicResampler.cs:31-38)
    (200.0, (0, 0), 0, 48000),             # (int)Pos above the 132 frames a call is offered: one output (phase 0), then the source ends
])
def test_looping_resampler_matches_the_index_model(rate, loop, offset, fs_buf):
    fs = 48000
    x = synth.splitmix_uniform(902, 1000)
    nb = 12
    got = _looping_resampled(x, fs_buf, fs, rate, loop[0], loop[1], offset, 128 * nb)
    le = min(int(loop[1] / fs_buf * fs_buf) if loop[1] > 0 else 1000, 1000)
    ls = min(int(loop[0] / fs_buf * fs_buf), le)
    eff = (fs_buf / float(fs)) * float(np.float32(rate))
    want, n_live = synth.loop_resample_model(x, int(offset / fs_buf * fs_buf), ls, le, eff, nb)
    assert np.array_equal(got, want)
    if rate == 8.0:
        assert np.all(want.reshape(nb, 128)[:, -3:] == 0) and np.count_nonzero(want.reshape(nb, 128)[:, :120]) > 110 * nb
    if rate == 200.0:
        assert n_live == 1 and np.count_nonzero(got) == 1 and got[0] == x[1]


# ---------------------------------------------------------------- OscillatorNode / ConstantSourceNode / AudioParam modulation
# (oracle-side groundwork for SURVEY.md §8f-3: the device path does not accelerate these yet and rejects them)
def test_constant_source_and_sample_accurate_start_stop():
    fs = 48000
    ctx = O.OfflineAudioContext(fs)
    c = O.ConstantSourceNode(ctx)
    c.Offset.SetValueAtTime(0.25, 0.0)
    c.Offset.LinearRampToValueAtTime(0.75, 0.01)
    c.Connect(ctx.Destination)
    c.Start(100.5 / fs)   # inside block 0: first playing frame = ceil((start - t0) * fs) = 101 (ConstantSourceNode.cs:88-93)
    c.Stop(300.5 / fs)    # inside block 2: last playing frame = floor((stop - t0) * fs) - 1 (:95-101)
    y = ctx.Render(128 * 4)
    probe = O.OfflineAudioContext(fs)
    g = O.GainNode(probe)
    g.Gain.SetValueAtTime(0.25, 0.0)
    g.Gain.LinearRampToValueAtTime(0.75, 0.01)
    want = g.Gain.evaluate(4)
    # 300.5 / fs lies in block 2 (frames 256..383): endFrame = floor((stop - t0) * fs) = floor(44.5) = 44 -> frames 256..299 play
    want[:101] = 0
    want[300:] = 0
    assert np.array_equal(y[0], want) and np.array_equal(y[1], want)  # mono output up-mixed by copy at the destination


def test_oscillator_phase_accumulation_and_waveforms():
    fs = 48000
    n = 128 * 6

    def run(kind, freq):
        ctx = O.OfflineAudioContext(fs)
        o = O.OscillatorNode(ctx)
        o.Type = kind
        o.Frequency.Value = freq
        o.Connect(ctx.Destination)
        o.Start()
        return ctx.Render(n)[0]

    def phases(freq):  # OscillatorNode.cs:131-138: double accumulation with a single 2*pi wrap per step
        ph, out = 0.0, np.empty(n)
        inc = (2.0 * math.pi * float(np.float32(freq))) / fs
        for i in range(n):
            out[i] = ph
            ph += inc
            if ph >= 2.0 * math.pi:
                ph -= 2.0 * math.pi
        return out
    ph = phases(997.0)
    assert np.array_equal(run(O.OscillatorType.Sine, 997.0), np.sin(ph).astype(np.float32))
    assert np.array_equal(run(O.OscillatorType.Square, 997.0), np.where(ph < math.pi, 1.0, -1.0).astype(np.float32))
    assert np.array_equal(run(O.OscillatorType.Sawtooth, 997.0), (2.0 * (ph / (2.0 * math.pi)) - 1.0).astype(np.float32))
    t = ph / (2.0 * math.pi)
    assert np.array_equal(run(O.OscillatorType.Triangle, 997.0), (4.0 * np.abs(t - np.floor(t + 0.5)) - 1.0).astype(np.float32))


def test_audio_param_modulation_adds_the_mono_mix_and_clamps():
    # AudioParam.cs:93-166: nodes connected to a param are mixed down to ONE channel (Explicit, channelCount 1) and added to the
    # intrinsic value, clamped to [min, max]; a silent modulation input leaves the intrinsic value alone
    fs = 48000
    x = [synth.splitmix_uniform(720 + c, 128 * 8) for c in range(2)]
    ctx = O.OfflineAudioContext(fs)
    s = O.AudioBufferSourceNode(ctx)
    s.Buffer = O.PlayableAudioBuffer.FromChannelArrays(x, fs)
    d = O.DelayNode(ctx, 0.01)                  # DelayTime in [0, 0.01]: the clamp is easy to hit
    d.DelayTime.Value = 0.004
    lfo = O.ConstantSourceNode(ctx)
    lfo.Offset.SetValueAtTime(-0.006, 0.0)      # 0.004 - 0.006 < 0 -> clamped to 0 (reads nothing)
    lfo.Offset.SetValueAtTime(0.002, 256.0 / fs)  # 0.006 s = 288 frames
    lfo.Offset.SetValueAtTime(0.02, 512.0 / fs)   # clamped to the maximum, 480 frames
    lfo.Connect(d.DelayTime)
    lfo.Start(128.0 / fs)                       # block 0: the modulator is silent -> intrinsic 0.004 s = 192 frames
    s.Connect(d).Connect(ctx.Destination)
    s.Start()
    y = ctx.Render(128 * 7)[0]
    n = np.arange(128 * 7)
    dly = np.select([n < 128, n < 256, n < 512], [192, 0, 288], 480)
    src = np.where((dly >= 1) & (n - dly >= 0), x[0][np.maximum(n - dly, 0)], 0.0).astype(np.float32)
    assert np.array_equal(y, src)


def test_channel_splitter_and_merger_route_single_channels():
    # ChannelSplitterNode.cs:21-55 (output i = input channel i as a mono block), ChannelMergerNode.cs:21-52 (output channel i = channel 0 of
    # input i): swap the channels of a stereo source and scale one of them on the way
    fs = 48000
    x = [synth.splitmix_uniform(730 + c, 128 * 6) for c in range(2)]
    ctx = O.OfflineAudioContext(fs)
    s = O.AudioBufferSourceNode(ctx)
    s.Buffer = O.PlayableAudioBuffer.FromChannelArrays(x, fs)
    split = O.ChannelSplitterNode(ctx, 2)
    merge = O.ChannelMergerNode(ctx, 2)
    g = O.GainNode(ctx)
    g.Gain.Value = 0.5
    s.Connect(split)
    split.Connect(g, 0, 0)       # left -> gain -> merger input 1
    g.Connect(merge, 0, 1)
    split.Connect(merge, 1, 0)   # right -> merger input 0
    merge.Connect(ctx.Destination)
    s.Start()
    y = ctx.Render(128 * 5)
    assert np.array_equal(y[0], x[1][:640]) and np.array_equal(y[1], x[0][:640] * np.float32(0.5))
    # a third splitter output of a stereo input is silence; a merger with one silent input keeps zeros there
    ctx = O.OfflineAudioContext(fs)
    s = O.AudioBufferSourceNode(ctx)
    s.Buffer = O.PlayableAudioBuffer.FromChannelArrays(x, fs)
    split = O.ChannelSplitterNode(ctx, 3)
    merge = O.ChannelMergerNode(ctx, 2)
    s.Connect(split)
    split.Connect(merge, 2, 0)
    split.Connect(merge, 0, 1)
    merge.Connect(ctx.Destination)
    s.Start()
    y = ctx.Render(128 * 5)
    assert not y[0].any() and np.array_equal(y[1], x[0][:640])


def test_edits_between_render_calls_act_from_the_next_unprocessed_block():
    # OfflineAudioContext.cs:55-100 renders whole 128-frame blocks and stashes what a call did not hand out; a parameter set after
    # Render(1000) therefore applies from frame 1024 on (block 8), not from frame 1000 — and a source started "in the past" begins
    # with the next block (AudioBufferSourceNode.cs:137-143 tests t1 > startTime per block)
    fs = 48000
    x = np.ones(128 * 40, np.float32)
    ctx = O.OfflineAudioContext(fs)
    s = O.AudioBufferSourceNode(ctx)
    s.Buffer = O.PlayableAudioBuffer.FromChannelArrays([x], fs)
    g = O.GainNode(ctx)
    g.Gain.Value = 0.5
    s.Connect(g).Connect(ctx.Destination)
    s.Start()
    a = ctx.Render(1000)
    g.Gain.Value = 0.25
    late = O.AudioBufferSourceNode(ctx)
    late.Buffer = O.PlayableAudioBuffer.FromChannelArrays([x * np.float32(4.0)], fs)
    late.Connect(ctx.Destination)
    late.Start(0.0)
    b = ctx.Render(1000)
    y = np.concatenate([a, b], axis=1)[0]
    assert np.all(y[:1024] == 0.5)                    # the block that was already processed keeps the old gain
    assert np.all(y[1024:] == np.float32(0.25) + 4.0)  # new gain and the late source from block 8 on
