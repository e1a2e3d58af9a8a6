"""GPU: the spectral MAC as a fast convolution along block time (csrc/fft2.cu, mac_variant 3 / the default for IRs of
>= 64 partitions) against the oracle.

Kernel level: gac_spectral_mac(variant=3) vs ProcessSpectralConvolution restated in float32 numpy in the reference's op
order (PartitionedConvolver.cs:154-223), for every second-level transform length (512 ... 8192), single- and multi-segment
overlap-save, ragged ends.  Render level: the BASELINE graph shapes with the second-level FFT forced (small IRs would
otherwise take the direct MAC) and with the direct MAC forced, both against the CPU oracle at the 1e-5 gate.
"""
import ctypes as C

import numpy as np
import pytest

from tests import synth
from tests.test_gpu_kernels import _fp, _mac_reference_unfused, _native

pytestmark = pytest.mark.gpu
TOL = 1e-5
FS = 48000


def _apis():
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    return G, O


# (Q blocks, P partitions) -> second-level length M: P-1 rounded up to 16 = Lh, M = smallest 2^n >= 512 with M - Lh >= M/2
@pytest.mark.parametrize("Q,P", [
    (1, 1), (70, 21), (33, 100),          # M = 512, one segment
    (600, 100),                            # M = 512, two segments (V = 400)
    (900, 300),                            # M = 1024, V = 720: two segments, ragged end
    (3000, 750),                           # M = 2048 (BASELINE geometry), V = 1296: three segments
    (500, 1200),                           # M = 4096, fewer blocks than partitions
    (300, 2500),                           # M = 8192
])
def test_spectral_mac_second_level_fft(Q, P):
    import graphaudio_b200 as G
    N, check = _native()
    ctx = G.OfflineAudioContext(FS)
    S, B = 2, 128
    rng = np.random.default_rng(Q * 1000 + P)
    X = (rng.uniform(-1, 1, (S, Q, B)) + 1j * rng.uniform(-1, 1, (S, Q, B))).astype(np.complex64)
    # decaying partitions, like an impulse response: keeps the accumulated magnitude O(1)
    env = np.exp(-3.0 * np.arange(P) / max(P, 1))[None, :, None]
    H = ((rng.uniform(-1, 1, (S, P, B)) + 1j * rng.uniform(-1, 1, (S, P, B))) * env * 0.05).astype(np.complex64)
    Y = np.zeros((S, Q, B), np.complex64)
    check(N.lib().gac_spectral_mac(ctx._h, _fp(X.view(np.float32)), _fp(H.view(np.float32)), S, Q, P, 3, _fp(Y.view(np.float32))))
    for s in range(S):
        ref = _mac_reference_unfused(X[s], H[s])
        err = np.abs(Y[s] - ref).max()
        # float32 FFT convolution vs float32 direct sum: both carry ~1e-7 relative to the accumulated magnitude
        assert err <= 2e-6 * max(1.0, np.abs(ref).max()), (err, np.abs(ref).max())
    ctx.Dispose()


@pytest.mark.parametrize("B", [256, 512])
def test_spectral_mac_second_level_fft_other_partitions(B):
    import graphaudio_b200 as G
    N, check = _native()
    ctx = G.OfflineAudioContext(FS, partition=B)
    S, Q, P = 1, 130, 70
    rng = np.random.default_rng(B)
    X = (rng.uniform(-1, 1, (S, Q, B)) + 1j * rng.uniform(-1, 1, (S, Q, B))).astype(np.complex64)
    H = ((rng.uniform(-1, 1, (S, P, B)) + 1j * rng.uniform(-1, 1, (S, P, B))) * 0.05).astype(np.complex64)
    Y = np.zeros((S, Q, B), np.complex64)
    check(N.lib().gac_spectral_mac(ctx._h, _fp(X.view(np.float32)), _fp(H.view(np.float32)), S, Q, P, 3, _fp(Y.view(np.float32))))
    ref = _mac_reference_unfused(X[0], H[0])
    assert np.abs(Y[0] - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())
    ctx.Dispose()


def _voices(nv, src_frames, ir_frames):
    return [synth.make_voice_inputs(v, src_frames, ir_frames) + (synth.voice_gains(v),) for v in range(nv)]


@pytest.mark.parametrize("variant", [3, 4])
def test_c2_render_forced_variants(variant):
    """C2 shape with the second-level FFT (3) and the register-tiled direct MAC (4) forced; odd block count."""
    G, O = _apis()
    voices = _voices(5, FS // 2, 12000)
    n = FS // 2 + 14000 + 77
    yg = synth.build_c2(G, FS, voices, 1.0 / 4, t_scale=0.05, mac_variant=variant).Render(n)
    yo = synth.build_c2(O, FS, voices, 1.0 / 4, t_scale=0.05).Render(n)
    assert np.abs(yo).max() > 1e-2
    assert np.abs(yg - yo).max() <= TOL


def test_default_picks_second_level_fft_for_long_irs_only():
    G, O = _apis()
    src, ir_long = synth.make_voice_inputs(0, 20000, 128 * 64)  # 64 partitions
    _, ir_short = synth.make_voice_inputs(0, 20000, 128 * 63)
    a = synth.build_c1(G, FS, src, ir_long)
    ya = a.Render(30000)
    assert a.last_stats["mac_variant_used"] == 3
    b = synth.build_c1(G, FS, src, ir_short)
    yb = b.Render(30000)
    assert b.last_stats["mac_variant_used"] == 4
    assert np.abs(ya - synth.build_c1(O, FS, src, ir_long).Render(30000)).max() <= TOL
    assert np.abs(yb - synth.build_c1(O, FS, src, ir_short).Render(30000)).max() <= TOL
    a.Dispose()
    b.Dispose()


def test_mixed_ir_lengths_in_one_render():
    """voices with different IR lengths land in different second-level transform lengths (and the direct MAC) in one render"""
    G, O = _apis()

    def build(api, **kw):
        ctx = api.OfflineAudioContext(FS, **kw)
        for v, L in enumerate([3000, 128 * 70, 128 * 300, 128 * 600]):
            src = [synth.splitmix_uniform(4 * v + c, 30000) for c in range(2)]
            ir = [synth.decay_ir(4 * v + 2 + c, L) for c in range(2)]
            s = api.AudioBufferSourceNode(ctx)
            s.Buffer = api.PlayableAudioBuffer.FromChannelArrays(src, FS)
            conv = api.ConvolverNode(ctx)
            conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays(ir, FS)
            g = api.GainNode(ctx)
            g.Gain.Value = 0.25
            s.Connect(conv).Connect(g).Connect(ctx.Destination)
            s.Start()
        return ctx
    n = 30000 + 128 * 600
    yg = build(G).Render(n)
    yo = build(O).Render(n)
    assert np.abs(yo).max() > 1e-2
    assert np.abs(yg - yo).max() <= TOL


@pytest.mark.parametrize("partition", [128, 256, 512])
def test_c5_resampled_source_second_level_fft(partition):
    G, O = _apis()
    fs, src_rate = 96000, 44100
    voices = []
    for v in range(2):
        src = [synth.splitmix_uniform(4 * v + c, 22050) for c in range(2)]
        ir = [synth.decay_ir(4 * v + 2 + c, 40000) for c in range(2)]
        voices.append((src, ir, synth.voice_gains(v)))
    g = synth.build_c5(G, fs, src_rate, voices, 0.5, t_scale=0.05, partition=partition, mac_variant=3)
    o = synth.build_c5(O, fs, src_rate, voices, 0.5, t_scale=0.05)
    n = 48000 + 44000
    yg, yo = g.Render(n), o.Render(n)
    assert np.abs(yo).max() > 1e-3
    assert np.abs(yg - yo).max() <= TOL
    g.Dispose()


def _render_modes(api, src_channels, ir_channels, n, **kw):
    ctx = api.OfflineAudioContext(FS, **kw)
    s = api.AudioBufferSourceNode(ctx)
    s.Buffer = api.PlayableAudioBuffer.FromChannelArrays(src_channels, FS)
    g = api.GainNode(ctx)
    g.Gain.SetValueAtTime(0.8, 0.0)
    g.Gain.LinearRampToValueAtTime(0.4, 0.2)
    conv = api.ConvolverNode(ctx)
    conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays(ir_channels, FS)
    s.Connect(g).Connect(conv).Connect(ctx.Destination)
    s.Start()
    return ctx.Render(n)


@pytest.mark.parametrize("n_ir", [1, 4])
def test_mono_and_true_stereo_irs_second_level_fft(n_ir):
    G, O = _apis()
    src = [synth.splitmix_uniform(50 + c, 20000) for c in range(2)]
    ir = [synth.decay_ir(60 + c, 9000) * np.float32(0.5 + 0.1 * c) for c in range(n_ir)]
    yg = _render_modes(G, src, ir, 31000, mac_variant=3)
    yo = _render_modes(O, src, ir, 31000)
    assert np.abs(yo).max() > 1e-3
    assert np.abs(yg - yo).max() <= TOL
    if n_ir == 1:
        assert np.array_equal(yg[0], yg[1])


def test_chunked_render_equals_single_render_second_level_fft():
    import graphaudio_b200 as G
    src, ir = synth.make_voice_inputs(1, 40000, 128 * 80)
    a = synth.build_c1(G, FS, src, ir)
    b = synth.build_c1(G, FS, src, ir)
    whole = a.Render(30000)
    p1 = b.Render(13333)
    p2 = b.Render(16667)
    # the second render recomputes frames [0, 30000) with a different number of blocks: segment boundaries move, so the
    # spectra differ in the last bits (float32 FFT), unlike the direct MAC which is bit-reproducible
    assert np.abs(np.concatenate([p1, p2], axis=1) - whole).max() <= 2e-6
    a.Dispose()
    b.Dispose()


@pytest.mark.parametrize("partition", [128, 512])
def test_long_ir_at_96k_through_the_render_path(partition):
    """3 s impulse response at 96 kHz: 2250 partitions of 128 frames -> the M = 8192 radix-8 second-level plan;
    563 partitions of 512 frames -> M = 2048 with the B = 512 first-level kernels.  Against the oracle (128-frame partitions)."""
    G, O = _apis()
    fs = 96000
    src = [synth.splitmix_uniform(900 + c, fs // 2) for c in range(2)]
    ir = [synth.decay_ir(910 + c, 3 * fs) for c in range(2)]
    g = synth.build_c1(G, fs, src, ir, partition=partition)
    o = synth.build_c1(O, fs, src, ir)
    n = fs // 2 + 3 * fs
    yg, yo = g.Render(n), o.Render(n)
    assert g.last_stats["mac_variant_used"] == 3
    assert np.abs(yo).max() > 1e-2
    assert np.abs(yg - yo).max() <= TOL
    g.Dispose()


@pytest.mark.parametrize("when,offset,duration", [
    (0.0, 0.0, None),          # read in place from frame 0
    (0.011, 0.0, None),        # starts in block 4: the alias pointer sits before the buffer start, only [lo, hi) is touched
    (0.0, 0.0101, None),       # even buffer offset (484 frames): in place
    (0.0, 0.01002083, None),   # odd buffer offset (481 frames): falls back to the copy
    (0.004, 0.0025, 0.2),      # duration-limited playback, the tail of the buffer is never read
])
def test_sources_read_in_place_by_the_convolver(when, offset, duration):
    """rate-1 sources that feed a convolver (directly or through a fused GainNode) are not copied: K5 reads the source
    buffer in place inside the non-silent range.  Start / offset / duration variants against the oracle."""
    G, O = _apis()

    def build(api):
        ctx = api.OfflineAudioContext(FS)
        for v, with_gain in enumerate([True, False]):
            s = api.AudioBufferSourceNode(ctx)
            s.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.splitmix_uniform(800 + 4 * v + c, 20000) for c in range(1 + v)], FS)
            conv = api.ConvolverNode(ctx)
            conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays([synth.decay_ir(810 + 4 * v + c, 128 * 65 + 1) for c in range(2)], FS)
            node = s
            if with_gain:
                g = api.GainNode(ctx)
                g.Gain.SetValueAtTime(0.9, 0.0)
                g.Gain.LinearRampToValueAtTime(0.3, 0.3)
                node = s.Connect(g)
            node.Connect(conv).Connect(ctx.Destination)
            if duration is None:
                s.Start(when, offset)
            else:
                s.Start(when, offset, duration)
        return ctx
    n = 20000 + 128 * 70
    yg = build(G).Render(n)
    yo = build(O).Render(n)
    assert np.abs(yo).max() > 1e-2
    assert np.abs(yg - yo).max() <= TOL


@pytest.mark.parametrize("n_ir,n_frames,mixed_expected", [
    (12800, 128 * 1300, True),   # P = 100 (M = 512, V = 400): one 1024-point segment (912 outputs) + one 512-point segment instead of four
    (12800, 128 * 900, True),    # ... one 1024-point segment alone instead of three 512-point ones
    (12800, 128 * 350, False),   # a single short segment stays
    (38400, 128 * 1700, True),   # P = 300 (M = 1024, V = 720): one 2048-point segment (V = 1744) instead of three 1024-point ones
])
def test_mixed_segment_lengths_match_uniform_segments_and_the_oracle(n_ir, n_frames, mixed_expected):
    """K6 with double-length overlap-save segments in front (the default; plan_segments) against the single-length plan and
    the CPU oracle."""
    G, O = _apis()
    src = [synth.splitmix_uniform(800 + c, n_frames - 256) for c in range(2)]
    ir = [synth.decay_ir(810 + c, n_ir) for c in range(2)]

    def build(api, **kw):
        ctx = api.OfflineAudioContext(FS, **kw)
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays(src, FS)
        g = api.GainNode(ctx)
        g.Gain.SetValueAtTime(0.9, 0.0)
        g.Gain.LinearRampToValueAtTime(0.3, 1.0)
        conv = api.ConvolverNode(ctx)
        conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays(ir, FS)
        s.Connect(g).Connect(conv).Connect(ctx.Destination)
        s.Start()
        return ctx
    cm, cu = build(G), build(G, uniform_segments=True)
    ym, yu, yo = cm.Render(n_frames), cu.Render(n_frames), build(O).Render(n_frames)
    assert cm.last_stats["mac_variant_used"] == 3 and cu.last_stats["mac_variant_used"] == 3
    assert (cm.last_stats["mac_flops"] < 0.99 * cu.last_stats["mac_flops"]) == mixed_expected
    assert np.abs(yo).max() > 1e-2
    assert np.abs(ym - yo).max() <= TOL and np.abs(yu - yo).max() <= TOL
    assert np.abs(ym - yu).max() <= 2e-6
    ym2 = cm.Render(128 * 10)  # the cached double-length spectra serve the next render of the context too
    assert np.isfinite(ym2).all()
