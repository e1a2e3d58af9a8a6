"""GPU parity cases added in round 2 (VERDICT r01 "parity holes"), all against the CPU oracle, the same builder
code on both sides:

  * every FilterType of BiQuadFilterNode (Nodes/BiQuadFilterNode.cs:160-245) with swept Frequency and Q and a non-zero
    k-rate Gain (the shelving / peaking `A = pow(10, gain/40)` path), plus a measurement of how often the device's
    coefficient evaluation differs from the oracle's at all;
  * a fan-in whose head node's input is NOT Max/2: a ConvolverNode with a mono impulse response (Explicit 1:
    AudioNodeInput.cs:140-168, mono inputs added as they are, stereo inputs as (L + R)/sqrt(2) each, :214-228) and a
    StereoPannerNode (ClampedMax 2);
  * more distinct resampled source geometries in one batch than the context's phase-table cache holds
    (engine_render.inl: plan_sources).
"""
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


def _src(api, ctx, stream, n, channels=2, rate=FS):
    s = api.AudioBufferSourceNode(ctx)
    s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(stream + c, n) for c in range(channels)], rate)
    return s


def _both(build, n):
    G, O = _apis()
    g = build(G)
    yg = g.Render(n)
    yo = build(O).Render(n)
    g.Dispose()
    return yg, yo


FILTER_NAMES = ["Lowpass", "Highpass", "Bandpass", "Notch", "Allpass", "Peaking", "Lowshelf", "Highshelf"]


@pytest.mark.parametrize("ftype", FILTER_NAMES)
@pytest.mark.parametrize("gain_db", [0.0, 7.5, -11.0])
def test_every_filter_type_with_swept_frequency_q_and_gain(ftype, gain_db):
    """Source -> BiQuadFilterNode(type, f sweep 300 -> 9000 Hz exponential, Q ramp 0.6 -> 4 -> 0.9, Gain k-rate) -> destination.
    The recursion is float32 in the reference's order; the only sanctioned difference is the last-ulp behaviour of the
    transcendental functions behind the coefficients (sinf / cosf restated exactly; powf / sqrtf of the shelving types)."""
    n = 48000

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = _src(api, ctx, 900, n)
        bq = api.BiQuadFilterNode(ctx)
        bq.Type = getattr(api.FilterType, ftype)
        bq.Frequency.SetValueAtTime(300.0, 0.0)
        bq.Frequency.ExponentialRampToValueAtTime(9000.0, 0.8)
        bq.Q.SetValueAtTime(0.6, 0.0)
        bq.Q.LinearRampToValueAtTime(4.0, 0.4)
        bq.Q.LinearRampToValueAtTime(0.9, 0.9)
        bq.Gain.Value = gain_db
        g = api.GainNode(ctx)
        g.Gain.Value = 0.25
        s.Connect(bq).Connect(g).Connect(ctx.Destination)
        s.Start()
        return ctx

    yg, yo = _both(build, n)
    peak = float(np.abs(yo).max())
    err = float(np.abs(yg - yo).max())
    assert peak > 1e-2
    assert err <= TOL, (ftype, gain_db, err, peak)


@pytest.mark.parametrize("ftype", ["Peaking", "Lowshelf", "Highshelf"])
def test_shelving_gain_automation_is_k_rate(ftype):
    """Gain is k-rate (BiQuadFilterNode.cs:77-82, :104): one value per quantum, taken from the first frame's time; a ramp on it
    changes `A` once per quantum."""
    n = 24000

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        s = _src(api, ctx, 910, n)
        bq = api.BiQuadFilterNode(ctx)
        bq.Type = getattr(api.FilterType, ftype)
        bq.Frequency.Value = 1200.0
        bq.Q.Value = 1.3
        bq.Gain.SetValueAtTime(-9.0, 0.0)
        bq.Gain.LinearRampToValueAtTime(12.0, 0.4)
        g = api.GainNode(ctx)
        g.Gain.Value = 0.2
        s.Connect(bq).Connect(g).Connect(ctx.Destination)
        s.Start()
        return ctx

    yg, yo = _both(build, n)
    assert np.abs(yo).max() > 1e-2
    assert np.abs(yg - yo).max() <= TOL


def test_coefficient_mismatch_rate_is_measured():
    """How often does a whole render differ from the oracle AT ALL (not just within the gate)?  Constant-parameter filters of the
    types whose coefficients go through pow / sqrt: the device computes `A` with a double-precision pow rounded to float
    (biquad_math.cuh), the oracle with glibc's powf.  The two agree except on rounding ties; this test records the observed
    rate so that the claim in DESIGN.md is a number: it asserts the gate, and that at least 80 % of the (type, gain) cases are
    bit-identical to the oracle."""
    n = 12800
    cases, exact, worst = 0, 0, 0.0
    for ftype in ["Peaking", "Lowshelf", "Highshelf"]:
        for k in range(12):
            gain_db = -30.0 + 5.37 * k
            freq = 150.0 * (1.6 ** k)

            def build(api, ftype=ftype, gain_db=gain_db, freq=freq):
                ctx = api.OfflineAudioContext(FS)
                s = _src(api, ctx, 920, n)
                bq = api.BiQuadFilterNode(ctx)
                bq.Type = getattr(api.FilterType, ftype)
                bq.Frequency.Value = min(freq, 20000.0)
                bq.Q.Value = 0.9
                bq.Gain.Value = gain_db
                g = api.GainNode(ctx)
                g.Gain.Value = 0.02
                s.Connect(bq).Connect(g).Connect(ctx.Destination)
                s.Start()
                return ctx

            yg, yo = _both(build, n)
            e = float(np.abs(yg - yo).max())
            worst = max(worst, e)
            cases += 1
            exact += int(np.array_equal(yg, yo))
            assert e <= TOL, (ftype, gain_db, freq, e)
    print(f"shelving / peaking coefficient parity: {exact} of {cases} renders bit-identical to the oracle, worst max|err| {worst:.3e}")
    assert exact >= 0.8 * cases, (exact, cases)


# ------------------------------------------------------------------------------------------------ fan-in heads
def test_fan_in_head_is_a_mono_ir_convolver():
    """two mono sources, one stereo source and one processed (stereo) chain -> ConvolverNode(mono IR) -> destination.
    The convolver's input is Explicit(1): mono inputs are added unscaled, stereo ones as (L + R) * (1 / sqrt 2) each."""
    n = 16000

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        conv = api.ConvolverNode(ctx)
        conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.decay_ir(940, 128 * 70)], FS)
        out = api.GainNode(ctx)
        out.Gain.Value = 0.5
        conv.Connect(out).Connect(ctx.Destination)
        for v, ch in enumerate([1, 2, 1]):
            s = _src(api, ctx, 930 + 4 * v, n - 700 * v, channels=ch)
            s.Connect(conv)
            s.Start()
        s = _src(api, ctx, 950, n, channels=1)
        g = api.GainNode(ctx)  # a GainNode's output always has two channels
        g.Gain.Value = 0.7
        s.Connect(g).Connect(conv)
        s.Start()
        return ctx

    yg, yo = _both(build, n + 128 * 72)
    assert np.abs(yo).max() > 1e-2
    assert np.abs(yg - yo).max() <= TOL, np.abs(yg - yo).max()


def test_mono_ir_convolver_outputs_feed_a_mono_ir_convolver():
    """the advisor's example: mono-IR convolver outputs (one channel each) summed into another mono-IR convolver"""
    n = 12000

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        tail = api.ConvolverNode(ctx)
        tail.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.decay_ir(960, 128 * 66)], FS)
        tail.Connect(ctx.Destination)
        for v in range(2):
            s = _src(api, ctx, 962 + 4 * v, n, channels=2)
            c = api.ConvolverNode(ctx)
            c.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.decay_ir(970 + v, 128 * 65)], FS)
            s.Connect(c).Connect(tail)
            s.Start()
        return ctx

    yg, yo = _both(build, n + 128 * 140)
    assert np.abs(yo).max() > 1e-3
    assert np.abs(yg - yo).max() <= TOL * max(1.0, float(np.abs(yo).max())), np.abs(yg - yo).max()


@pytest.mark.parametrize("mixed", [False, True])
def test_fan_in_head_is_a_stereo_panner(mixed):
    """mono inputs only -> the panner's ClampedMax input stays mono (ProcessMono, StereoPannerNode.cs:62-66); with a processed
    stereo chain among them it is stereo throughout (ProcessStereo)."""
    n = 16000

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        pan = api.StereoPannerNode(ctx)
        pan.Pan.SetValueAtTime(-0.6, 0.0)
        pan.Pan.LinearRampToValueAtTime(0.8, 0.3)
        pan.Connect(ctx.Destination)
        for v in range(2):
            s = _src(api, ctx, 980 + 4 * v, n, channels=1)
            s.Connect(pan)
            s.Start()
        if mixed:
            s = _src(api, ctx, 990, n, channels=2)
            g = api.GainNode(ctx)
            g.Gain.Value = 0.5
            s.Connect(g).Connect(pan)
            s.Start()
        return ctx

    yg, yo = _both(build, n)
    assert np.abs(yo).max() > 1e-2
    assert np.abs(yg - yo).max() <= TOL, np.abs(yg - yo).max()


# ------------------------------------------------------------------------------------------------ resample cache
def test_more_resampled_geometries_than_the_phase_table_cache_holds():
    """20 voices with 20 different playback rates and buffer lengths in ONE batch: every voice keeps its own (k, t) phase table
    although the context caches only 16 of them (an eviction while the batch is planned must not free tables the batch still uses)."""
    nv = 20

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        bus = api.GainNode(ctx)
        bus.Gain.Value = 1.0 / 8
        bus.Connect(ctx.Destination)
        for v in range(nv):
            s = _src(api, ctx, 1000 + 4 * v, 9000 + 611 * v, channels=2, rate=44100 if v % 2 else 32000)
            s.PlaybackRate.Value = 0.8 + 0.037 * v
            s.Connect(bus)
            s.Start()
        return ctx

    n = 24000
    yg, yo = _both(build, n)
    assert np.abs(yo).max() > 1e-2
    assert np.array_equal(yg, yo) or np.abs(yg - yo).max() <= 1e-6, np.abs(yg - yo).max()


# ------------------------------------------------------------------------------------------------ §8b kernel hooks
def _param(N, value, events=()):
    import ctypes as C
    p = N.gac_param()
    p.value = value
    p.n_events = len(events)
    keep = None
    if events:
        keep = (N.gac_event * len(events))()
        for i, (t, v, tg, tm, tc) in enumerate(events):
            keep[i].type, keep[i].value, keep[i].target, keep[i].time, keep[i].time_constant = t, v, tg, tm, tc
        p.events = keep
    return p, keep


@pytest.mark.parametrize("ftype", range(8))
def test_gac_biquad_batch_hook_against_the_oracle_node(ftype):
    """kernel-level: gac_biquad_batch (production biquad kernels, no graph around them) vs the oracle's BiQuadFilterNode on the
    same stereo noise with an a-rate frequency sweep.  Types whose coefficients need only sin / cos (restated bit-exactly) must be
    BIT-IDENTICAL; peaking / shelving types (pow, sqrt) stay within 1e-6."""
    import ctypes as C
    import graphaudio_b200 as G
    from graphaudio_b200 import _native as N
    from graphaudio_b200.api import check
    from oracle import ga_oracle as O
    n, ns = 128 * 300, 3
    x = np.stack([np.stack([synth.splitmix_uniform(1100 + 2 * s + c, n) for c in range(2)]) for s in range(ns)]).astype(np.float32)
    f_ev = lambda s: [(0, 200.0 + 150.0 * s, 0.0, 0.0, 0.0), (2, 7000.0 - 900.0 * s, 0.0, 0.6, 0.0)]  # noqa: E731
    q_of = lambda s: 0.5 + 0.8 * s  # noqa: E731
    gdb = 6.0
    ctx = G.OfflineAudioContext(FS)
    keep = []
    fr = (N.gac_param * ns)()
    qq = (N.gac_param * ns)()
    gg = (N.gac_param * ns)()
    for s in range(ns):
        for arr, (p, k) in ((fr, _param(N, 1000.0, f_ev(s))), (qq, _param(N, q_of(s))), (gg, _param(N, gdb))):
            arr[s] = p
            keep.append(k)
    types = (C.c_int * ns)(*[ftype] * ns)
    y = np.zeros_like(x)
    check(N.lib().gac_biquad_batch(ctx._h, x.ctypes.data_as(N.fp), ns, n, types, fr, qq, gg, y.ctypes.data_as(N.fp)))
    ctx.Dispose()
    for s in range(ns):
        o = O.OfflineAudioContext(FS)
        src = O.AudioBufferSourceNode(o)
        src.Buffer = O.PlayableAudioBuffer.FromChannelArrays([x[s, 0], x[s, 1]], FS)
        bq = O.BiQuadFilterNode(o)
        bq.Type = ftype
        for (t, v, tg, tm, tc) in f_ev(s):
            (bq.Frequency.SetValueAtTime if t == 0 else bq.Frequency.ExponentialRampToValueAtTime)(v, tm)
        bq.Q.Value = q_of(s)
        bq.Gain.Value = gdb
        src.Connect(bq).Connect(o.Destination)
        src.Start()
        yo = o.Render(n)
        m = 128 * ((n - 1) // 128)  # the source drops its final block (AudioBufferSourceNode.cs:360-368); the hook filters all of x
        if ftype <= 4:
            assert np.array_equal(y[s][:, :m], yo[:, :m]), (ftype, s, np.abs(y[s][:, :m] - yo[:, :m]).max())
        else:
            assert np.abs(y[s][:, :m] - yo[:, :m]).max() <= 1e-6, (ftype, s)


def test_gac_mix_hook_is_the_sequential_float32_sum():
    import ctypes as C
    import graphaudio_b200 as G
    from graphaudio_b200 import _native as N
    from graphaudio_b200.api import check
    n, ni = 128 * 40, 7
    xs = [np.stack([synth.splitmix_uniform(1200 + 2 * i + c, n) for c in range(2)]).astype(np.float32) for i in range(ni)]
    lo = [0, 128 * 3, 0, 128 * 10, 128 * 39, 0, 128 * 5]
    hi = [n, n, 128 * 20, 128 * 11, n, 0, 128 * 30]
    dm = [0.0, 0.0, float(np.float32(1.0) / np.sqrt(np.float32(2.0))), 0.0, 0.0, 0.0, float(np.float32(1.0) / np.sqrt(np.float32(2.0)))]
    ref = np.zeros((2, n), np.float32)
    for i in range(ni):
        sl = slice(lo[i], hi[i])
        if dm[i]:
            m = ((np.float32(0) + xs[i][0, sl]) + xs[i][1, sl]) * np.float32(dm[i])
            ref[0, sl] += m
            ref[1, sl] += m
        else:
            ref[:, sl] += xs[i][:, sl]
    ctx = G.OfflineAudioContext(FS)
    ptrs = (N.fp * ni)(*[x.ctypes.data_as(N.fp) for x in xs])
    out = np.zeros((2, n), np.float32)
    check(N.lib().gac_mix(ctx._h, ptrs, (C.c_int64 * ni)(*lo), (C.c_int64 * ni)(*hi), (C.c_float * ni)(*dm), ni, n, out.ctypes.data_as(N.fp)))
    ctx.Dispose()
    assert np.array_equal(out, ref)
