"""GPU: BASELINE.json's full sizes.

The oracle renders ~30 voice-seconds per second, so full parity is run on a full-LENGTH subset (8 voices x 12 s x 2 s IR =
BASELINE configs[1] geometry: P = 750, Q = 4500) and the full 64-voice workload is checked through size-independent
properties of the path:
  * linearity in the sources: scaling every source by 2 scales the output by exactly 2 (powers of two commute with every
    rounding on the path: copy, float32 multiply by the gain, FFT butterflies, MAC, overlap-add, fan-in sum);
  * the bus is the float32 sum of its voices: rendering the voices in two halves and adding the results reproduces the
    full render up to float32 association;
  * a delta source turns the convolver into a table lookup: the output is the normalised impulse response times the gain
    automation at frame 0 ... i.e. y[n] = float32(h[n] * scale) * g[0], checked against the oracle's normalisation scale.
"""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu

FS = 48000
SRC, IR, N = 10 * FS, 2 * FS, 12 * FS  # BASELINE configs[1]


def _voices(nv, scale=1.0):
    out = []
    for v in range(nv):
        src, ir = synth.make_voice_inputs(v, SRC, IR)
        if scale != 1.0:
            src = [s * np.float32(scale) for s in src]
        out.append((src, ir, synth.voice_gains(v)))
    return out


def test_c2_full_length_subset_matches_oracle():
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    voices = _voices(8)
    g = synth.build_c2(G, FS, voices, 1.0 / 8)
    yg = g.Render(N)
    yo = synth.build_c2(O, FS, voices, 1.0 / 8).Render(N)
    assert np.abs(yo).max() > 0.05
    assert np.abs(yg - yo).max() <= 1e-5
    st = g.last_stats
    assert st["conv_units"] == 8 * 2 * 4500 and st["kernel_launches"] > 0
    # the production K6 plan of the bench workload: second-level FFT, one 4096-point overlap-save segment in front of one 2048-point one
    assert st["mac_variant_used"] == 3 and st["mac_big_segments"] == 1
    g.Dispose()


def test_c2_full_size_linearity_and_bus_additivity():
    import graphaudio_b200 as G
    nv = 64
    voices = _voices(nv)
    ctx = synth.build_c2(G, FS, voices, 1.0 / 8)
    y = ctx.Render(N)
    peak = np.abs(y).max()
    assert 0.25 <= peak <= 1.0, peak  # synthetic inputs scaled so that the bus peak lies in [0.25, 1]
    ctx.Dispose()
    # (1) exact linearity under a power-of-two scaling of every source
    ctx2 = synth.build_c2(G, FS, _voices(nv, 0.5), 1.0 / 8)
    y2 = ctx2.Render(N)
    ctx2.Dispose()
    assert np.array_equal(y2 * np.float32(2.0), y)
    # (2) bus additivity: two half-renders add up to the full bus (float32 association only)
    a = synth.build_c2(G, FS, voices[:nv // 2], 1.0 / 8)
    b = synth.build_c2(G, FS, voices[nv // 2:], 1.0 / 8)
    ya, yb = a.Render(N), b.Render(N)
    a.Dispose()
    b.Dispose()
    assert np.abs((ya + yb) - y).max() <= 2e-6


def test_delta_source_reproduces_normalised_ir_at_full_ir_length():
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    ir = [synth.decay_ir(2 + c, IR) for c in range(2)]
    d = np.zeros(4096, np.float32)
    d[0] = 1.0
    ctx = G.OfflineAudioContext(FS)
    s = G.AudioBufferSourceNode(ctx)
    s.Buffer = G.PlayableAudioBuffer.FromStereoArrays(d, d, FS)
    conv = G.ConvolverNode(ctx)
    conv.Buffer = G.PlayableAudioBuffer.FromChannelArrays(ir, FS)
    s.Connect(conv).Connect(ctx.Destination)
    s.Start()
    y = ctx.Render(IR + 256)
    ctx.Dispose()
    for c in range(2):
        ref = ir[c] * np.float32(O.normalization_scale(ir[c]))
        assert np.abs(y[c, :IR] - ref).max() <= 2e-7
        assert np.abs(y[c, IR:]).max() <= 1e-7


def test_c3_full_length_subset_matches_oracle():
    """BASELINE configs[2] geometry (biquad lowpass sweep -> gain -> 2 s IR, 12 s) on a 4-voice subset against the oracle:
    the biquad recursion runs as verified concurrent time segments over the full 576 000 frames."""
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    voices = _voices(4)
    g = synth.build_c3(G, FS, voices, 1.0 / 2)
    yg = g.Render(N)
    yo = synth.build_c3(O, FS, voices, 1.0 / 2).Render(N)
    assert np.abs(yo).max() > 0.05
    assert np.abs(yg - yo).max() <= 1e-5
    g.Dispose()


def test_c4_full_length_renders_match_oracle():
    """BASELINE configs[3] at its real size: independent renders of 5 s (240 000 frames), each Source -> BiQuad(lowpass sweep) ->
    BiQuad(highpass 200 Hz, constant) -> ConvolverNode(0.5 s stereo IR) -> destination, batched by gac_render_batch.  The constant
    200 Hz highpass is the filter whose speculative time segments never re-join bitwise (poles at radius 0.98): the whole
    signal goes through the sequential repair path, which is what this test pins at full length (3 renders: the batch crosses
    the 16-row group of the biquad lanes only with more, so a fourth case runs 17 renders on a shorter signal elsewhere)."""
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    fs, n_src, n_ir, n = 48000, 5 * 48000, 24000, 240000
    parent = G.OfflineAudioContext(fs)
    ctxs, refs = [], []
    for r in range(3):
        src, ir = synth.make_voice_inputs(40 + r, n_src, n_ir)

        class _Api:  # the builder creates its own context: hand it a fork of the shared device context instead
            pass
        api = _Api()
        for name in dir(G):
            setattr(api, name, getattr(G, name))
        api.OfflineAudioContext = lambda fs_, **kw: parent.Fork()
        ctxs.append(synth.build_c4(api, fs, src, ir))
        refs.append(synth.build_c4(O, fs, src, ir).Render(n))
    out = G.RenderBatch(ctxs, n)
    for r in range(3):
        assert np.abs(refs[r]).max() > 0.02
        err = np.abs(out[r] - refs[r]).max()
        assert err <= 1e-5, (r, err)
    assert parent.last_stats["conv_units"] == 3 * 2 * 1875
    parent.Dispose()


def test_c5_real_geometry_one_voice_matches_partition_512_oracle():
    """BASELINE configs[4] at its real geometry on one voice: 44.1 kHz stereo source of 10 s in a 96 kHz context (CubicResampler,
    rate 0.459375, 960 000 output frames) -> GainNode automation -> convolver with a 10 s stereo IR (960 000 frames) at 512-frame
    partitions (P = 1875, second-level transform M = 4096) -> bus gain, 20 s = 1 920 000 frames.
    Oracle (note on C5): the graph up to the convolver's input rendered by the oracle's nodes, then
    PartitionedConvolver(blockSize = 512) per channel driven with 512-frame blocks, then the bus gain — ConvolverNode itself
    always constructs with 128 (ConvolverNode.cs:55), the 512 partition is the constructor argument of PartitionedConvolver.cs:37."""
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    fs, src_rate = 96000, 44100
    n_src, n_ir, n = 441000, 960000, 1920000
    src = [synth.splitmix_uniform(4 * 7 + c, n_src) for c in range(2)]
    ir = [synth.decay_ir(4 * 7 + 2 + c, n_ir) for c in range(2)]
    gains = synth.voice_gains(7)
    bus_gain = 0.125  # bus peak ~0.5 (inputs scaled so that the bus peak lies in [0.25, 1])
    g = synth.build_c5(G, fs, src_rate, [(src, ir, gains)], bus_gain, partition=512)
    yg = g.Render(n)
    st = g.last_stats
    assert st["conv_units"] == 2 * 3750 and st["mac_variant_used"] == 3
    g.Dispose()
    # oracle: source -> gain (same automation) -> destination, then the 512-frame convolvers
    o = O.OfflineAudioContext(fs)
    s = O.AudioBufferSourceNode(o)
    s.Buffer = O.PlayableAudioBuffer.FromChannelArrays(src, src_rate)
    gn = O.GainNode(o)
    synth.add_gain_automation(gn.Gain, gains)
    s.Connect(gn).Connect(o.Destination)
    s.Start()
    x = o.Render(n)
    yo = np.stack([O.PartitionedConvolver(ir[c], 512, True).process(x[c]) for c in range(2)]) * np.float32(bus_gain)
    assert np.abs(yo).max() > 0.01
    err = np.abs(yg - yo).max()
    assert err <= 1e-5, err
