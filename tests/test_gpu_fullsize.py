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
