"""End-to-end parity (GPU): OfflineAudioContext.Render through the C ABI vs the CPU oracle on identical synthetic
inputs.  Gate: max |y_gpu - y_oracle| <= 1e-5 of full scale (float32) — BASELINE.json north_star.

The graphs are the BASELINE configs at sizes the oracle renders in seconds; the same builder code drives both sides.
"""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _apis():
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    return G, O


def _voices(nv, src_frames, ir_frames, src_channels=2):
    out = []
    for v in range(nv):
        src, ir = synth.make_voice_inputs(v, src_frames, ir_frames, src_channels)
        out.append((src, ir, synth.voice_gains(v)))
    return out


def _compare(gctx, octx, n, tol=TOL):
    yg = gctx.Render(n)
    yo = octx.Render(n)
    peak = np.abs(yo).max()
    err = np.abs(yg - yo).max()
    assert peak > 1e-3, "degenerate test signal"
    assert err <= tol, (err, peak)
    return yg, yo


def test_c1_single_convolver():
    """C1 shape: stereo source -> ConvolverNode(stereo IR) -> destination (48 kHz, 0.25 s IR, 1 s input + tail)."""
    G, O = _apis()
    fs = 48000
    src, ir = synth.make_voice_inputs(0, fs, fs // 4)
    g = synth.build_c1(G, fs, src, ir)
    o = synth.build_c1(O, fs, src, ir)
    yg, yo = _compare(g, o, fs + fs // 2)
    # source end rule (AudioBufferSourceNode.cs:224,360-362): exactly 128*floor((L-1)/128) input frames are emitted;
    # after the tail has rung out the output is exactly silent
    g.Dispose()


def test_c2_gain_automation_and_bus_mix():
    """C2 shape: 6 voices x (gain automation -> per-voice stereo IR) -> bus GainNode(1/8) -> destination."""
    G, O = _apis()
    fs = 48000
    voices = _voices(6, fs, 12000)
    g = synth.build_c2(G, fs, voices, 1.0 / 8, t_scale=0.1)
    o = synth.build_c2(O, fs, voices, 1.0 / 8, t_scale=0.1)
    _compare(g, o, fs + 16000)
    g.Dispose()


@pytest.mark.parametrize("f0,f1,q", [(2000.0, 12000.0, 0.707), (200.0, 2000.0, 0.707), (500.0, 8000.0, 8.0), (40.0, 60.0, 20.0)])
def test_c3_biquad_sweep(f0, f1, q):
    """C3 shape: biquad lowpass with an a-rate exponential cutoff sweep -> gain -> convolver -> bus.
    The recursion runs as concurrent time segments that are verified bit for bit (biquad_lanes.cu); the 40 Hz / Q = 20 case
    never forgets its state within the warm-up, so it exercises the sequential repair path."""
    G, O = _apis()
    fs = 48000
    voices = _voices(4, fs // 2, 6000)
    g = synth.build_c3(G, fs, voices, 1.0 / 4, f0=f0, f1=f1, t_scale=0.05, q=q)
    o = synth.build_c3(O, fs, voices, 1.0 / 4, f0=f0, f1=f1, t_scale=0.05, q=q)
    _compare(g, o, fs // 2 + 8000)
    g.Dispose()


def test_c4_biquad_chain_independent_render():
    G, O = _apis()
    fs = 48000
    src, ir = synth.make_voice_inputs(5, fs // 2, 6000)
    g = synth.build_c4(G, fs, src, ir, t_scale=0.1)
    o = synth.build_c4(O, fs, src, ir, t_scale=0.1)
    _compare(g, o, fs // 2 + 8000)
    g.Dispose()


@pytest.mark.parametrize("partition", [128, 512])
def test_c5_resampled_source(partition):
    """C5 shape: 44.1 kHz buffer in a 96 kHz context (CubicResampler, rate 0.459375) -> gain -> convolver -> bus.
    partition=512 is the offline-only option: same linear convolution, different rounding points, same 1e-5 gate
    against the reference's 128-frame ConvolverNode."""
    G, O = _apis()
    fs, src_rate = 96000, 44100
    voices = []
    for v in range(3):
        src = [synth.splitmix_uniform(4 * v + c, 22050) for c in range(2)]
        ir = [synth.decay_ir(4 * v + 2 + c, 20000) for c in range(2)]
        voices.append((src, ir, synth.voice_gains(v)))
    g = synth.build_c5(G, fs, src_rate, voices, 0.5, t_scale=0.05, partition=partition)
    o = synth.build_c5(O, fs, src_rate, voices, 0.5, t_scale=0.05)
    _compare(g, o, 48000 + 24000)
    g.Dispose()


def test_chunked_render_equals_single_render():
    """KAT: Render(n1) + Render(n2) == Render(n1 + n2) for non-multiples of 128 (OfflineAudioContext.cs:55-100)."""
    G, O = _apis()
    fs = 48000
    src, ir = synth.make_voice_inputs(1, 20000, 3000)
    a = synth.build_c1(G, fs, src, ir)
    b = synth.build_c1(G, fs, src, ir)
    whole = a.Render(10000)
    p1 = b.Render(3333)
    p2 = b.Render(6667)
    assert np.array_equal(np.concatenate([p1, p2], axis=1), whole)
    a.Dispose()
    b.Dispose()


def test_mono_source_upmix_and_source_end_rule():
    """mono source: 1 -> 2 up-mix copies the channel (AudioNodeInput.cs:201-213); emitted length 128*floor((L-1)/128)."""
    G, O = _apis()
    fs = 48000
    L = 128 * 10  # a multiple of 128: the last FULL block is dropped too
    x = synth.splitmix_uniform(9, L)
    outs = []
    for api in (G, O):
        ctx = api.OfflineAudioContext(fs)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromMonoArray(x, fs)
        g = api.GainNode(ctx)
        g.Gain.Value = 0.5
        s.Connect(g).Connect(ctx.Destination)
        s.Start()
        outs.append(ctx.Render(L + 256))
    yg, yo = outs
    assert np.array_equal(yg, yo)  # copy + float32 multiply: bit-exact
    emitted = 128 * ((L - 1) // 128)
    assert np.array_equal(yg[0, :emitted], (x[:emitted] * np.float32(0.5)))
    assert np.array_equal(yg[0], yg[1])
    assert not yg[:, emitted:].any()


def test_start_stop_block_granular():
    """Start(when)/Stop(when) are block-granular: playback begins at frame 0 of the first block whose end is past `when`."""
    G, O = _apis()
    fs = 48000
    x = synth.splitmix_uniform(11, 30000)
    outs = []
    for api in (G, O):
        ctx = api.OfflineAudioContext(fs)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromMonoArray(x, fs)
        s.Connect(ctx.Destination)
        s.Start(0.0301, 0.01, 0.2)
        outs.append(ctx.Render(20000))
    assert np.array_equal(outs[0], outs[1])
    assert outs[0].any()


def test_fan_in_order_and_direct_voices():
    """two voices straight into the destination + one bus: float32 ((0 + a) + b) in connection order (AudioNodeInput.cs:118-137)."""
    G, O = _apis()
    fs = 48000
    outs = []
    for api in (G, O):
        ctx = api.OfflineAudioContext(fs)
        bus = api.GainNode(ctx)
        bus.Gain.Value = 0.7
        for v in range(3):
            s = api.AudioBufferSourceNode(ctx)
            s.Buffer = api.PlayableAudioBuffer.FromStereoArrays(synth.splitmix_uniform(50 + 2 * v, 9000), synth.splitmix_uniform(51 + 2 * v, 9000), fs)
            g = api.GainNode(ctx)
            g.Gain.Value = 0.3 + 0.1 * v
            s.Connect(g).Connect(bus)
            s.Start(0.01 * v)
        d = api.AudioBufferSourceNode(ctx)
        d.Buffer = api.PlayableAudioBuffer.FromMonoArray(synth.splitmix_uniform(60, 5000), fs)
        d.Connect(ctx.Destination)
        bus.Connect(ctx.Destination)
        d.Start()
        outs.append(ctx.Render(10000))
    assert np.array_equal(outs[0], outs[1])  # adds and multiplies only: bit-exact, including the summation order


def test_streaming_mac_variant_matches_oracle_tighter():
    """mac_variant=1 keeps the reference's unfused op order; only the float32 FFTs differ from the oracle."""
    G, O = _apis()
    fs = 48000
    src, ir = synth.make_voice_inputs(2, 24000, 6000)
    g = synth.build_c1(G, fs, src, ir, mac_variant=1)
    o = synth.build_c1(O, fs, src, ir)
    _compare(g, o, 30000, tol=2e-6)
    g.Dispose()


def test_error_mapping():
    G, O = _apis()
    ctx = G.OfflineAudioContext(48000)
    with pytest.raises(G.ArgumentOutOfRangeException):
        ctx.Render(0)
    with pytest.raises(G.ArgumentOutOfRangeException):
        ctx.Render(np.zeros((2, 10), np.float32), 10, -1)
    with pytest.raises(G.ArgumentException):
        ctx.Render(np.zeros((2, 10), np.float32), 11, 0)
    conv = G.ConvolverNode(ctx)
    with pytest.raises(G.InvalidOperationException):  # IR rate mismatch, ConvolverNode.cs:48-49
        conv.Buffer = G.PlayableAudioBuffer.FromMonoArray(np.ones(100, np.float32), 44100)
    ctx.Dispose()
    with pytest.raises(G.ObjectDisposedException):
        ctx.Render(128)


def test_c4_render_batch_of_independent_contexts():
    """BASELINE config 4 shape: a batch of independent renders (biquad chain + convolver each) in one gac_render_batch call
    equals the per-context oracle renders."""
    G, O = _apis()
    fs = 48000
    parent = G.OfflineAudioContext(fs)
    forks, refs = [], []
    for r in range(5):
        src, ir = synth.make_voice_inputs(20 + r, 10000 + 700 * r, 3000)

        class _Api:  # build_c4 against a fork of the shared device context
            pass
        shim = _Api()
        for name in ("AudioBufferSourceNode", "BiQuadFilterNode", "ConvolverNode", "PlayableAudioBuffer", "FilterType"):
            setattr(shim, name, getattr(G, name))
        shim.OfflineAudioContext = lambda _fs, _p=parent: _p.Fork()
        forks.append(synth.build_c4(shim, fs, src, ir, t_scale=0.03))
        refs.append(synth.build_c4(O, fs, src, ir, t_scale=0.03).Render(16000))
    out = G.RenderBatch(forks, 16000)
    assert out.shape == (5, 2, 16000)
    for r in range(5):
        assert np.abs(refs[r]).max() > 1e-3
        assert np.abs(out[r] - refs[r]).max() <= TOL, r
    parent.Dispose()


def test_biquad_slow_drift_takes_the_general_walk_and_the_per_row_layout():
    """A cutoff that drifts slower than the 1e-3 Hz hysteresis recomputes every few frames, at different frames in the two
    channels (channel 1 starts each block from channel 0's last value): exercises the serial walk of k_biquad_select and the
    per-row stream layout of the recursion kernel; a second voice with a constant filter shares the 32-row group."""
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    fs = 48000

    def build(api):
        ctx = api.OfflineAudioContext(fs)
        for v, (f0, f1) in enumerate([(1000.0, 1004.0), (700.0, 700.0), (2500.0, 2500.4)]):
            s = api.AudioBufferSourceNode(ctx)
            s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(700 + 4 * v + c, 30000) for c in range(2)], fs)
            bq = api.BiQuadFilterNode(ctx)
            bq.Type = api.FilterType.Lowpass if v != 1 else api.FilterType.Peaking
            bq.Q.Value = 1.5
            bq.Gain.Value = 3.0
            bq.Frequency.SetValueAtTime(f0, 0.0)
            if f1 != f0:
                bq.Frequency.LinearRampToValueAtTime(f1, 0.6)
            g = api.GainNode(ctx)
            g.Gain.Value = 0.3
            s.Connect(bq).Connect(g).Connect(ctx.Destination)
            s.Start()
        return ctx
    n = 30000
    yg = build(G).Render(n)
    yo = build(O).Render(n)
    assert np.abs(yo).max() > 1e-2
    assert np.abs(yg - yo).max() <= TOL


def test_interleaved_render_equals_planar_render():
    """gac_render_interleaved ≙ ProcessBlockInterleaved block after block (AudioContextBase.cs:88-161)."""
    import graphaudio_b200 as G

    def build():
        ctx = G.OfflineAudioContext(48000)
        s = G.AudioBufferSourceNode(ctx)
        s.Buffer = G.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(900 + c, 5000) for c in range(2)], 48000)
        g = G.GainNode(ctx)
        g.Gain.Value = 0.5
        s.Connect(g).Connect(ctx.Destination)
        s.Start()
        return ctx
    planar = build().Render(4096)
    for ch in (1, 2, 5):
        ctx = build()
        a = ctx.RenderInterleaved(1024, ch).reshape(-1, ch)
        b = ctx.RenderInterleaved(3072, ch).reshape(-1, ch)  # the timeline continues
        inter = np.concatenate([a, b])
        for c in range(ch):
            assert np.array_equal(inter[:, c], planar[c] if c < 2 else np.zeros(4096, np.float32))
        ctx.Dispose()
    with pytest.raises(G.ArgumentOutOfRangeException):
        build().RenderInterleaved(128, 33)


@pytest.mark.parametrize("loop_start,loop_end,offset,start,stop", [
    (0.0, 0.0, 0.0, 0.0, None),            # whole buffer, forever
    (100.2, 400.2, 0.0, 0.0, None),        # plays into the loop region, then cycles it (loop shorter than 3 quanta)
    (100.2, 400.2, 650.0, 0.01, None),     # start position behind the loop end: the first quantum restarts at LoopStart
    (10.5, 50.5, 20.0, 0.0, 0.1),          # loop shorter than a quantum; Stop(0.1)
    (0.0, 0.0, 2990.0, 0.005, 0.2),
])
def test_looping_source_at_rate_one(loop_start, loop_end, offset, start, stop):
    """AudioBufferSourceNode.Loop / LoopStart / LoopEnd on the copy path (Nodes/AudioBufferSourceNode.cs:171-177, :186-235)."""
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    fs = 48000
    src = [synth.splitmix_uniform(950 + c, 3001) for c in range(2)]

    def build(api):
        ctx = api.OfflineAudioContext(fs)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays(src, fs)
        s.Loop = True
        s.LoopStart = loop_start / fs
        s.LoopEnd = loop_end / fs
        g = api.GainNode(ctx)
        g.Gain.Value = 0.5
        s.Connect(g).Connect(ctx.Destination)
        s.Start(start, offset / fs)
        if stop is not None:
            s.Stop(stop)
        return ctx
    yg, yo = build(G).Render(128 * 150), build(O).Render(128 * 150)
    assert np.abs(yo).max() > 0.1
    assert np.array_equal(yg, yo)


def test_looping_source_feeding_a_convolver_and_unsupported_variants():
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    fs = 48000
    src = [synth.splitmix_uniform(960 + c, 2000) for c in range(2)]
    ir = [synth.decay_ir(970 + c, 3000) for c in range(2)]

    def build(api, rate=1.0, ls=0.0, le=0.0):
        ctx = api.OfflineAudioContext(fs)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays(src, fs)
        s.Loop = True
        s.LoopStart, s.LoopEnd = ls, le
        s.PlaybackRate.Value = rate
        conv = api.ConvolverNode(ctx)
        conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays(ir, fs)
        s.Connect(conv).Connect(ctx.Destination)
        s.Start()
        return ctx
    yg, yo = build(G).Render(128 * 100), build(O).Render(128 * 100)
    assert np.abs(yg - yo).max() <= 1e-5
    yg, yo = build(G, rate=0.5).Render(128 * 100), build(O, rate=0.5).Render(128 * 100)   # the looping resampler path in front of K5
    assert np.abs(yo).max() > 0.05 and np.abs(yg - yo).max() <= 1e-5
    with pytest.raises(G.NotSupportedException):
        build(G, ls=0.01, le=0.01).Render(1280)   # an empty loop region: the reference never returns from it on the resampler path


@pytest.mark.parametrize("rate,loop,offset,fs_buf,start,stop,nch", [
    (0.5, (100.2, 400.2), 0.0, 48000, 0.0, None, 2),      # below 1: one Process call per quantum
    (1.37, (100.2, 400.2), 650.0, 48000, 0.01, None, 2),  # start position behind LoopEnd: the first call restarts at LoopStart
    (1.0, (10.5, 47.5), 20.0, 44100, 0.0, 0.1, 2),        # 44.1 -> 48 kHz, a 37-frame loop, Stop(0.1)
    (2.5, (0.0, 0.0), 0.0, 48000, 0.005, None, 1),        # mono buffer, LoopEnd 0 = end of the buffer, late start
    (8.0, (100.2, 400.2), 0.0, 48000, 0.0, None, 2),      # the tail of every quantum is cleared (:334-338)
    (0.9, (5.2, 8.2), 2.0, 48000, 0.0, None, 2),          # a 3-frame loop: priming takes two calls
    (200.0, (0.0, 0.0), 0.0, 48000, 0.0, None, 2),        # one output, then the source ends (:360-368)
    (0.731, (0.0, 2500.0), 1234.0, 32000, 0.0, 0.35, 2),
])
def test_looping_source_on_the_resampler_path(rate, loop, offset, fs_buf, start, stop, nch):
    """AudioBufferSourceNode.Loop on the CubicResampler path (Nodes/AudioBufferSourceNode.cs:236-358: the 512-float wrap buffer, one pass
    over the loop region per Process call, cleared tails, the position bookkeeping of :323-357): the host replays the block loop with
    frame indices, the device evaluates the polynomial — bit-exact against the oracle, which restates the path with samples, and
    against the index-level model the oracle's own KATs use."""
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    fs = 48000
    src = [synth.splitmix_uniform(955 + c, 3001) for c in range(nch)]

    def build(api):
        ctx = api.OfflineAudioContext(fs)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays(src, fs_buf)
        s.Loop = True
        s.LoopStart, s.LoopEnd = loop[0] / fs_buf, loop[1] / fs_buf
        s.PlaybackRate.Value = rate
        g = api.GainNode(ctx)
        g.Gain.Value = 0.5
        s.Connect(g).Connect(ctx.Destination)
        s.Start(start, offset / fs_buf)
        if stop is not None:
            s.Stop(stop)
        return ctx
    nb = 150
    yg, yo = build(G).Render(128 * nb), build(O).Render(128 * nb)
    assert np.count_nonzero(yo) >= 1
    assert np.array_equal(yg, yo)
    if start == 0.0 and stop is None:
        le = min(int(loop[1] / fs_buf * fs_buf) if loop[1] > 0 else 3001, 3001)
        ls = min(int(loop[0] / fs_buf * fs_buf), le)
        eff = (fs_buf / float(fs)) * float(np.float32(rate))
        want, _ = synth.loop_resample_model(src[0], int(offset / fs_buf * fs_buf), ls, le, eff, nb)
        assert np.array_equal(yg[0], np.float32(0.5) * want)


def test_seventeen_looping_resampled_geometries_in_one_batch():
    """More distinct (rate, loop) geometries than the context's table cache holds (kResampleCacheMax = 16), seam windows included."""
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    fs = 48000

    def build(api):
        ctx = api.OfflineAudioContext(fs)
        for v in range(19):
            x = [synth.splitmix_uniform(3000 + 2 * v + c, 900 + 37 * v) for c in range(2)]
            s = api.AudioBufferSourceNode(ctx)
            s.Buffer = api.PlayableAudioBuffer.FromChannelArrays(x, fs)
            s.Loop = v % 5 != 4
            s.LoopStart, s.LoopEnd = (50.2 + 3 * v) / fs, (600.2 + 11 * v) / fs
            s.PlaybackRate.Value = 0.6 + 0.07 * v
            s.Connect(ctx.Destination)
            s.Start(0.0, (10.0 * v) / fs)
        return ctx
    yg, yo = build(G).Render(128 * 40), build(O).Render(128 * 40)
    assert np.abs(yo).max() > 1.0
    assert np.array_equal(yg, yo)


def test_parameter_edits_and_late_starts_between_successive_render_calls():
    """OfflineAudioContext.cs:55-100: successive Render calls continue the timeline, and what the caller does in between (AudioParam.Value,
    scheduling calls, Start / Stop of sources, new branches) acts from the next unprocessed 128-frame quantum on.  The device path
    re-renders the timeline with per-quantum parameter epochs; same call sequence on the oracle."""
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    fs = 48000

    def run(api):
        ctx = api.OfflineAudioContext(fs)
        bus = api.GainNode(ctx)
        bus.Gain.Value = 0.5
        bus.Connect(ctx.Destination)
        a = api.AudioBufferSourceNode(ctx)
        a.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(980 + c, 30000) for c in range(2)], fs)
        f = api.BiQuadFilterNode(ctx)
        f.Frequency.Value = 500.0
        f.Q.Value = 4.0
        g = api.GainNode(ctx)
        g.Gain.SetValueAtTime(0.2, 0.0)
        g.Gain.LinearRampToValueAtTime(1.0, 0.5)
        d = api.DelayNode(ctx, 0.1)
        d.DelayTime.Value = 0.01
        conv = api.ConvolverNode(ctx)
        conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.decay_ir(990 + c, 3000) for c in range(2)], fs)
        a.Connect(f).Connect(g).Connect(d).Connect(conv).Connect(bus)
        a.Start()
        out = [ctx.Render(1000)]                      # ends inside quantum 7: the edits below act from quantum 8 (frame 1024)
        g.Gain.Value = 0.6                            # clears the ramp
        f.Frequency.SetValueAtTime(500.0, 0.0)
        f.Frequency.ExponentialRampToValueAtTime(4000.0, 0.2)
        bus.Gain.Value = 0.25
        b = api.AudioBufferSourceNode(ctx)            # a new branch, started "in the past": plays from the next quantum
        b.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(985, 6000)], fs)
        pb = api.StereoPannerNode(ctx)
        pb.Pan.Value = -0.5
        b.Connect(pb).Connect(bus)
        b.Start(0.001)
        out.append(ctx.Render(3000))                  # frames 1000 .. 4000 (quantum boundary again inside a block)
        d.DelayTime.Value = 0.03
        g.Gain.SetTargetAtTime(0.0, 0.1, 0.05)
        a.Stop(0.0)                                   # "stop now": silent from the next unprocessed quantum
        out.append(ctx.Render(6000))
        return np.concatenate(out, axis=1)
    yg, yo = run(G), run(O)
    assert np.abs(yo[:, :1000]).max() > 1e-3 and np.abs(yo[:, 1100:4000]).max() > 1e-3 and np.abs(yo[:, 6000:]).max() > 1e-5
    assert np.abs(yg - yo).max() <= 1e-5


def _run_both(run):
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    return run(G), run(O)


def test_rewiring_between_render_calls_acts_from_the_next_quantum():
    """Connect / Disconnect after rendering began (Nodes/AudioNode.cs:109-147 post them to the render thread; the offline context
    drains them in front of the next ProcessBlock, AudioContextBase.cs:272-284): a node that already rendered gets a second
    consumer (a reverb send added late), a voice is taken off the bus, a source that existed all along is connected late."""
    fs = 48000

    def run(api):
        ctx = api.OfflineAudioContext(fs)
        bus = api.GainNode(ctx)
        bus.Gain.Value = 0.5
        bus.Connect(ctx.Destination)
        voices = []
        for v in range(3):
            s = api.AudioBufferSourceNode(ctx)
            s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(1400 + 2 * v + c, 30000) for c in range(2)], fs)
            g = api.GainNode(ctx)
            g.Gain.Value = 0.3 + 0.1 * v
            s.Connect(g).Connect(bus)
            s.Start()
            voices.append((s, g))
        late = api.AudioBufferSourceNode(ctx)   # created and started now, connected after the first Render call
        late.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(1410, 9000)], fs)
        late.Start()
        out = [ctx.Render(3000)]
        # (1) a send from a gain that already rendered into a new convolver
        conv, wet = api.ConvolverNode(ctx), api.GainNode(ctx)
        conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.decay_ir(1420 + c, 128 * 70) for c in range(2)], fs)
        wet.Gain.Value = 0.4
        voices[0][1].Connect(conv).Connect(wet).Connect(ctx.Destination)
        # (2) voice 1 leaves the bus; (3) the idle source joins it
        voices[1][1].Disconnect(bus)
        late.Connect(bus)
        out.append(ctx.Render(7000))
        voices[2][0].Disconnect()               # the source itself is unplugged from its gain
        out.append(ctx.Render(128 * 120))
        return np.concatenate(out, axis=1)

    yg, yo = _run_both(run)
    assert np.abs(yo).max() > 0.1
    assert np.abs(yg - yo).max() <= 1e-5, np.abs(yg - yo).max()


def test_impulse_response_swapped_between_render_calls():
    """ConvolverNode.Buffer set again mid-timeline (Nodes/ConvolverNode.cs:25-79): new PartitionedConvolvers with cleared delay
    lines take over from the next unprocessed quantum — the old tail is cut, the new convolvers have not seen the earlier input;
    then Buffer = null (the node clears its output) and a third impulse response of a different length"""
    fs = 48000

    def run(api):
        ctx = api.OfflineAudioContext(fs)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(1430 + c, 40000) for c in range(2)], fs)
        g = api.GainNode(ctx)
        g.Gain.SetValueAtTime(0.8, 0.0)
        g.Gain.LinearRampToValueAtTime(0.3, 0.5)
        conv = api.ConvolverNode(ctx)
        conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.decay_ir(1440 + c, 128 * 90) for c in range(2)], fs)
        s.Connect(g).Connect(conv).Connect(ctx.Destination)
        s.Start()
        out = [ctx.Render(5000)]
        conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.decay_ir(1450 + c, 128 * 70) for c in range(2)], fs)
        out.append(ctx.Render(9000))
        conv.Buffer = None
        out.append(ctx.Render(2000))
        conv.Normalize = False
        conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.decay_ir(1460 + c, 2000) * np.float32(0.02) for c in range(2)], fs)
        out.append(ctx.Render(12000))
        return np.concatenate(out, axis=1)

    yg, yo = _run_both(run)
    assert np.abs(yo[:, :5000]).max() > 0.05 and np.abs(yo[:, 5200:14000]).max() > 0.05 and np.abs(yo[:, 16200:]).max() > 0.01
    assert np.abs(yo[:, 14100:15900]).max() == 0.0
    assert np.abs(yg - yo).max() <= 1e-5, np.abs(yg - yo).max()


def test_reconnecting_a_removed_connection_is_refused_not_approximated():
    """the one re-wiring that is NOT reproduced: nodes that nothing pulled for a while would resume where they stopped"""
    import graphaudio_b200 as G
    fs = 48000
    ctx = G.OfflineAudioContext(fs)
    s = G.AudioBufferSourceNode(ctx)
    s.Buffer = G.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(995, 4000)], fs)
    g = G.GainNode(ctx)
    s.Connect(g).Connect(ctx.Destination)
    s.Start()
    ctx.Render(512)
    g.Disconnect(ctx.Destination)
    ctx.Render(512)
    g.Connect(ctx.Destination)
    with pytest.raises(G.NotSupportedException):
        ctx.Render(512)
    ctx.Dispose()


def test_distinct_contexts_render_concurrently_from_distinct_threads():
    """A gac_context is not thread-safe, but distinct contexts may be used from distinct threads (AudioContextBase.cs:266-305 is the
    same model): context creation, the process-wide table caches, staging blocks and the stream-ordered pool are shared state."""
    import threading
    import graphaudio_b200 as G
    fs = 48000
    n = 40000

    def build(seed, async_upload):
        voices = []
        for v in range(3):
            src, ir = synth.make_voice_inputs(seed * 10 + v, 30000, 9000)
            voices.append((src, ir, synth.voice_gains(seed + v)))
        return synth.build_c2(G, fs, voices, 0.25, t_scale=0.05, async_upload=async_upload)
    want = [build(s, False).Render(n) for s in range(4)]
    got = [None] * 4
    errors = []
    gate = threading.Barrier(4)

    def worker(s):
        try:
            gate.wait()
            for _ in range(3):  # fresh context every time: creation and destruction race with the other threads' renders
                ctx = build(s, async_upload=(s % 2 == 1))
                got[s] = ctx.Render(n)
                ctx.Dispose()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))
    threads = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for s in range(4):
        assert np.array_equal(got[s], want[s])


@pytest.mark.parametrize("case", ["ramp", "steps_through_one", "looping_ramp", "set_target", "resampled_buffer"])
def test_playback_rate_automation_is_evaluated_per_quantum(case):
    """AudioBufferSourceNode.PlaybackRate is a k-rate AudioParam (Nodes/AudioBufferSourceNode.cs:76,165-169): its value at the start of
    every quantum picks the path (copy at an effective rate of exactly 1, else CubicResampler — whose window and phase survive the
    quanta in between) and the phase increment.  The host evaluates the schedule and replays the positions; bit-exact against the oracle."""
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    fs = 48000
    fs_buf = 44100 if case == "resampled_buffer" else fs
    src = [synth.splitmix_uniform(975 + c, 20000) for c in range(2)]

    def build(api):
        ctx = api.OfflineAudioContext(fs)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays(src, fs_buf)
        r = s.PlaybackRate
        if case == "ramp":
            r.SetValueAtTime(0.5, 0.0)
            r.LinearRampToValueAtTime(2.0, 0.05)
            r.ExponentialRampToValueAtTime(0.8, 0.12)
        elif case == "steps_through_one":   # resampler -> copy path -> resampler (stale window) -> copy
            r.Value = 0.75
            r.SetValueAtTime(1.0, 0.02)
            r.SetValueAtTime(1.5, 0.04)
            r.SetValueAtTime(1.0, 0.06)
            r.SetValueAtTime(0.3, 0.09)
        elif case == "looping_ramp":
            s.Loop = True
            s.LoopStart, s.LoopEnd = 300.2 / fs, 2500.2 / fs
            r.SetValueAtTime(1.0, 0.0)
            r.SetValueAtTime(0.9, 0.03)
            r.LinearRampToValueAtTime(3.0, 0.2)
        elif case == "set_target":
            r.Value = 2.0
            r.SetTargetAtTime(0.5, 0.01, 0.03)
        else:
            r.SetValueAtTime(1.0, 0.0)
            r.LinearRampToValueAtTime(1.2, 0.1)
        g = api.GainNode(ctx)
        g.Gain.Value = 0.5
        s.Connect(g).Connect(ctx.Destination)
        s.Start(0.004, 100.0 / fs_buf)
        return ctx
    n = 128 * 120
    yg, yo = build(G).Render(n), build(O).Render(n)
    assert np.count_nonzero(yo) > 4000
    assert np.array_equal(yg, yo)


def test_playback_rate_value_edited_between_render_calls():
    """PlaybackRate.Value set between successive Render calls acts from the next unprocessed quantum on (an epoch of the parameter)."""
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    fs = 48000
    src = [synth.splitmix_uniform(985 + c, 30000) for c in range(2)]

    def run(api):
        ctx = api.OfflineAudioContext(fs)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays(src, fs)
        s.Connect(ctx.Destination)
        s.Start()
        a = ctx.Render(128 * 20)
        s.PlaybackRate.Value = 1.7
        b = ctx.Render(128 * 20)
        s.PlaybackRate.Value = 1.0
        c = ctx.Render(128 * 20)
        return np.concatenate([np.asarray(a), np.asarray(b), np.asarray(c)], axis=1)
    yg, yo = run(G), run(O)
    assert np.array_equal(yg, yo)
