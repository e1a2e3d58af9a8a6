"""Fan-in fusion (GAC_FLAG_NO_FANIN_FUSION off, the default): ConvolverNodes that end their chains and meet in one fan-in
(AudioNodeInput.MixBuffer, AudioNodeInput.cs:118-137) are summed as second-level spectra — one inverse transform pair and one
fan-in input for the whole group (csrc/fft2.cu k_fft2_sum16, engine.cu conv_batch_fft2_sum).  Same linear combination, another
float32 rounding order: every case is compared with the CPU oracle AND with the unfused CUDA path."""
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


def _voices(n, src_frames, ir_frames, seed=0):
    out = []
    for v in range(n):
        src, ir = synth.make_voice_inputs(seed + v, src_frames, ir_frames if np.isscalar(ir_frames) else ir_frames[v])
        out.append((src, ir, synth.voice_gains(seed + v)))
    return out


@pytest.mark.parametrize("n_voices,ir_frames,n_frames", [
    (6, 12800, 128 * 420),      # P = 100: one group, one chunk per channel
    (37, 9600, 128 * 300),      # P = 75 (M = 512): several chunks, ragged last chunk
    (5, 38400, 128 * 1800),     # P = 300 (M = 1024): mixed segment lengths inside the group on the second render
])
def test_voices_into_a_bus_are_summed_as_spectra(n_voices, ir_frames, n_frames):
    G, O = _apis()
    voices = _voices(n_voices, n_frames - 2000, ir_frames)
    gain = 1.0 / np.sqrt(n_voices)
    cf = synth.build_c2(G, FS, voices, gain, t_scale=0.05)
    cn = synth.build_c2(G, FS, voices, gain, t_scale=0.05, fanin_fusion=False)
    yo = synth.build_c2(O, FS, voices, gain, t_scale=0.05).Render(n_frames)
    for rep in range(2):  # (the second render of a context runs the mixed-segment plan where it pays)
        yf = _render_fresh(cf, n_frames)
        yn = _render_fresh(cn, n_frames)
        assert cf.last_stats["fanin_groups"] == 1 and cf.last_stats["fanin_members"] == n_voices
        assert cn.last_stats["fanin_groups"] == 0
        assert np.abs(yo).max() > 1e-2
        assert np.abs(yf - yo).max() <= TOL
        assert np.abs(yn - yo).max() <= TOL
        assert np.abs(yf - yn).max() <= 2e-6
    cf.Dispose()
    cn.Dispose()


def _render_fresh(ctx, n):
    """Renders frames [0, n) of the context's graph again (a new graph handle from frame 0), as bench.py's resident arm does."""
    import ctypes as C
    from graphaudio_b200 import _native as N
    from graphaudio_b200.api import check
    g = ctx._graph()
    out = np.zeros((2, n), np.float32)
    ptrs = (N.fp * 2)(*[out[c].ctypes.data_as(N.fp) for c in range(2)])
    check(N.lib().gac_render(ctx._h, g, 0, n, ptrs, 2, 0))
    st = N.gac_stats()
    check(N.lib().gac_get_stats(ctx._h, C.byref(st)))
    ctx.last_stats = st.as_dict()
    N.lib().gac_graph_destroy(g)
    return out


def test_groups_split_by_impulse_response_length_and_singletons_stay_unfused():
    """Voices whose impulse responses need different segment plans form groups of their own; a lone voice keeps the plain path."""
    G, O = _apis()
    n = 128 * 500
    irs = [12800, 12800, 12800, 38400, 38400, 3000]   # P = 100 x3 (a group), P = 300 x2 (a group), P = 24 (direct sum, never fused)
    voices = _voices(len(irs), n - 1000, irs, seed=40)
    cf = synth.build_c2(G, FS, voices, 0.4, t_scale=0.05)
    yf = cf.Render(n)
    yo = synth.build_c2(O, FS, voices, 0.4, t_scale=0.05).Render(n)
    assert cf.last_stats["fanin_groups"] == 2 and cf.last_stats["fanin_members"] == 5
    assert np.abs(yo).max() > 1e-2
    assert np.abs(yf - yo).max() <= TOL
    cf.Dispose()


def test_only_chain_ends_with_a_single_plain_consumer_are_fused():
    """(1) a convolver followed by another node, (2) a convolver whose output fans out to two consumers, (3) convolvers straight
    into the destination (a fan-in too), (4) mono impulse responses: the first two and the last stay unfused, all match the oracle."""
    G, O = _apis()
    n = 128 * 400
    src = [[synth.splitmix_uniform(900 + 2 * v + c, n - 500) for c in range(2)] for v in range(6)]
    ir = [[synth.decay_ir(950 + 2 * v + c, 12800) for c in range(2)] for v in range(6)]

    def build(api, **kw):
        ctx = api.OfflineAudioContext(FS, **kw)
        bus = api.GainNode(ctx)
        bus.Gain.Value = 0.5
        bus.Connect(ctx.Destination)

        def voice(v, ir_ch=None):
            s = api.AudioBufferSourceNode(ctx)
            s.Buffer = api.PlayableAudioBuffer.FromChannelArrays(src[v], FS)
            c = api.ConvolverNode(ctx)
            c.Buffer = api.PlayableAudioBuffer.FromChannelArrays(ir_ch if ir_ch else ir[v], FS)
            s.Connect(c)
            s.Start()
            return c
        voice(0).Connect(bus)                      # fused with voice 1
        voice(1).Connect(bus)
        tail = api.GainNode(ctx)                   # (1) the convolver does not end its chain
        tail.Gain.Value = 0.7
        voice(2).Connect(tail).Connect(bus)
        fan = voice(3)                             # (2) two consumers
        fan.Connect(bus)
        fan.Connect(ctx.Destination)
        voice(4).Connect(ctx.Destination)          # (3) direct voices meet in the destination's fan-in (with the bus and `fan`)
        voice(5, [ir[5][0]]).Connect(bus)          # (4) mono impulse response
        return ctx
    cf, cn = build(G), build(G, fanin_fusion=False)
    yf, yn, yo = cf.Render(n), cn.Render(n), build(O).Render(n)
    assert cf.last_stats["fanin_groups"] == 1 and cf.last_stats["fanin_members"] == 2
    assert np.abs(yo).max() > 1e-2
    assert np.abs(yf - yo).max() <= TOL and np.abs(yn - yo).max() <= TOL
    assert np.abs(yf - yn).max() <= 2e-6
    cf.Dispose()
    cn.Dispose()


def test_direct_voices_into_the_destination_form_a_group():
    G, O = _apis()
    n = 128 * 300
    src = [[synth.splitmix_uniform(700 + 2 * v + c, n - 300) for c in range(2)] for v in range(4)]
    ir = [[synth.decay_ir(750 + 2 * v + c, 12800) for c in range(2)] for v in range(4)]

    def build(api, **kw):
        ctx = api.OfflineAudioContext(FS, **kw)
        for v in range(4):
            s = api.AudioBufferSourceNode(ctx)
            s.Buffer = api.PlayableAudioBuffer.FromChannelArrays(src[v], FS)
            g = api.GainNode(ctx)
            g.Gain.Value = 0.3
            c = api.ConvolverNode(ctx)
            c.Buffer = api.PlayableAudioBuffer.FromChannelArrays(ir[v], FS)
            s.Connect(g).Connect(c).Connect(ctx.Destination)
            s.Start(0.01 * v)
        return ctx
    cf = build(G)
    yf, yo = cf.Render(n), build(O).Render(n)
    assert cf.last_stats["fanin_groups"] == 1 and cf.last_stats["fanin_members"] == 4
    assert np.abs(yo).max() > 1e-2
    assert np.abs(yf - yo).max() <= TOL
    cf.Dispose()


def test_chunked_render_calls_with_fusion_equal_one_call():
    """Successive Render calls (OfflineAudioContext.cs:55-100) of a fused graph against one call and the oracle."""
    G, O = _apis()
    n = 128 * 360
    voices = _voices(5, n - 1000, 12800, seed=70)
    c1 = synth.build_c2(G, FS, voices, 0.4, t_scale=0.05)
    c2 = synth.build_c2(G, FS, voices, 0.4, t_scale=0.05)
    y1 = c1.Render(n)
    y2 = np.concatenate([c2.Render(128 * 100), c2.Render(128 * 7 + 40), c2.Render(n - 128 * 107 - 40)], axis=1)
    yo = synth.build_c2(O, FS, voices, 0.4, t_scale=0.05).Render(n)
    assert c2.last_stats["fanin_groups"] == 1
    assert np.abs(y1 - yo).max() <= TOL and np.abs(y2 - yo).max() <= TOL
    assert np.abs(y1 - y2).max() <= 2e-6
    c1.Dispose()
    c2.Dispose()


def test_impulse_responses_beyond_the_radix16_plans_stay_unfused():
    """P = 2100 partitions need 8192-point second-level segments (radix-8 plan): no k_fft2_sum16 instantiation, the voices keep
    one inverse transform each and still meet the oracle."""
    G, O = _apis()
    n = 128 * 2300
    voices = _voices(2, 128 * 300, 128 * 2100 - 17, seed=90)
    cf = synth.build_c2(G, FS, voices, 0.5, t_scale=0.05)
    yf = cf.Render(n)
    yo = synth.build_c2(O, FS, voices, 0.5, t_scale=0.05).Render(n)
    assert cf.last_stats["mac_variant_used"] == 3 and cf.last_stats["fanin_groups"] == 0
    assert np.abs(yo).max() > 1e-2
    assert np.abs(yf - yo).max() <= TOL
    cf.Dispose()
