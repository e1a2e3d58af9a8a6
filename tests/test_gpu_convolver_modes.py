"""GPU: the remaining ConvolverNode modes against the CPU oracle.

  * mono IR: the node's input is forced to 1 channel (ConvolverNode.cs:72-76) — a stereo upstream is down-mixed
    (L + R) * (1/sqrt(2)), a mono upstream passes as is (AudioNodeInput.cs:188-228) — and the mono result is up-mixed by copy
    at the next input (:201-213);
  * 4-channel IR with EnableTrueStereo: L = c0(inL) + c2(inR), R = c1(inL) + c3(inR) (ConvolverNode.cs:127-144);
  * Normalize = false; each IR channel normalised by its own RMS.
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


def _render(api, src_channels, ir_channels, n, with_gain, normalize=True, true_stereo=True):
    ctx = api.OfflineAudioContext(FS)
    s = api.AudioBufferSourceNode(ctx)
    s.Buffer = api.PlayableAudioBuffer.FromChannelArrays(src_channels, FS)
    node = s
    if with_gain:
        g = api.GainNode(ctx)
        g.Gain.SetValueAtTime(0.8, 0.0)
        g.Gain.LinearRampToValueAtTime(0.4, 0.2)
        node = node.Connect(g)
    conv = api.ConvolverNode(ctx)
    conv.Normalize = normalize
    conv.EnableTrueStereo = true_stereo
    conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays(ir_channels, FS)
    node.Connect(conv).Connect(ctx.Destination)
    s.Start()
    return ctx.Render(n)


@pytest.mark.parametrize("src_ch,with_gain", [(2, False), (2, True), (1, False), (1, True)])
def test_mono_ir(src_ch, with_gain):
    G, O = _apis()
    src = [synth.splitmix_uniform(30 + c, 20000) for c in range(src_ch)]
    ir = [synth.decay_ir(40, 4000)]
    yg = _render(G, src, ir, 26000, with_gain)
    yo = _render(O, src, ir, 26000, with_gain)
    assert np.abs(yo).max() > 1e-3
    assert np.abs(yg - yo).max() <= TOL
    assert np.array_equal(yg[0], yg[1])  # mono output up-mixed by copy


def test_true_stereo_ir():
    G, O = _apis()
    src = [synth.splitmix_uniform(50 + c, 20000) for c in range(2)]
    ir = [synth.decay_ir(60 + c, 5000) * np.float32(0.5 + 0.1 * c) for c in range(4)]
    yg = _render(G, src, ir, 27000, True)
    yo = _render(O, src, ir, 27000, True)
    assert np.abs(yo).max() > 1e-3
    assert np.abs(yg - yo).max() <= TOL
    assert not np.array_equal(yg[0], yg[1])


def test_normalize_false_and_per_channel_rms():
    G, O = _apis()
    src = [synth.splitmix_uniform(70 + c, 12000) for c in range(2)]
    ir = [synth.decay_ir(80, 3000) * np.float32(0.02), synth.decay_ir(81, 3000) * np.float32(0.2)]
    for normalize in (False, True):
        yg = _render(G, src, ir, 16000, False, normalize=normalize)
        yo = _render(O, src, ir, 16000, False, normalize=normalize)
        assert np.abs(yg - yo).max() <= TOL


def test_discrete_four_channel_ir_is_rejected():
    G, _ = _apis()
    ctx = G.OfflineAudioContext(FS)
    conv = G.ConvolverNode(ctx)
    conv.EnableTrueStereo = False
    with pytest.raises(G.NotSupportedException):
        conv.Buffer = G.PlayableAudioBuffer.FromChannelArrays([np.ones(256, np.float32)] * 4, FS)
    ctx.Dispose()
