"""GPU: the steps either side of the render: decoder output uploaded as interleaved file samples and converted on
the device (gac_buffer_create_interleaved ≙ AudioDecoder.LoadFromStream + DecodePlanar, GraphAudio.IO/LibsndfileDecoder.cs:195-220),
and the interleaved writer (gac_render_interleaved ≙ ProcessBlockInterleaved, AudioContextBase.cs:88-161)."""
import io
import struct

import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu
FS = 48000


def _wav(samples_bytes, code, channels, bits, extensible=False):
    if extensible:
        fmt = struct.pack("<HHIIHHHHIH", 0xFFFE, channels, FS, FS * channels * bits // 8, channels * bits // 8, bits, 22, bits, 3, code) + b"\0" * 14
    else:
        fmt = struct.pack("<HHIIHH", code, channels, FS, FS * channels * bits // 8, channels * bits // 8, bits)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"LIST" + struct.pack("<I", 3) + b"abc\0"  # an odd-sized chunk to skip
    body += b"data" + struct.pack("<I", len(samples_bytes)) + samples_bytes
    return b"RIFF" + struct.pack("<I", len(body)) + body


def _cases():
    rng = np.random.default_rng(5)
    n, c = 5000, 2
    i16 = rng.integers(-32768, 32768, (n, c), dtype=np.int64).astype("<i2")
    i16[0], i16[1] = (-32768, 32767), (0, -1)
    i24 = rng.integers(-(1 << 23), 1 << 23, (n, c), dtype=np.int64)
    i24[0] = (-(1 << 23), (1 << 23) - 1)
    b24 = np.stack([(i24 >> s) & 255 for s in (0, 8, 16)], axis=-1).astype(np.uint8)
    i32 = rng.integers(-(1 << 31), 1 << 31, (n, c), dtype=np.int64).astype("<i4")
    i32[0] = (-(1 << 31), (1 << 31) - 1)
    f32 = rng.standard_normal((n, c)).astype("<f4")
    return [
        ("s16", i16.tobytes(), 1, 16, i16.astype(np.float32) * np.float32(2.0 ** -15)),
        ("s24", b24.tobytes(), 1, 24, i24.astype(np.float32) * np.float32(2.0 ** -23)),
        ("s32", i32.tobytes(), 1, 32, i32.astype(np.float32) * np.float32(2.0 ** -31)),
        ("f32", f32.tobytes(), 3, 32, f32),
    ]


@pytest.mark.parametrize("async_upload", [False, True])
def test_decoded_wav_is_converted_on_the_device_exactly_like_sf_readf_float(async_upload):
    import graphaudio_b200 as G
    from graphaudio_b200.io import AudioDecoder
    for name, raw, code, bits, want in _cases():
        buf = AudioDecoder.LoadFromStream(io.BytesIO(_wav(raw, code, 2, bits, extensible=(name == "s24"))))
        assert (buf.NumberOfChannels, buf.Length, buf.SampleRate) == (2, 5000, FS)
        assert np.array_equal(np.stack(buf.channels), want.T)  # the host view of the same samples (GetChannelData)
        ctx = G.OfflineAudioContext(FS, async_upload=async_upload)
        s = G.AudioBufferSourceNode(ctx)
        s.Buffer = buf
        s.Connect(ctx.Destination)
        s.Start()
        y = ctx.Render(4992)  # 39 full quanta (the source drops its final, partial block)
        assert np.array_equal(y, want.T[:, :4992]), name
        ctx.Dispose()


def test_mono_and_validation():
    import graphaudio_b200 as G
    from graphaudio_b200.io import AudioDecoder
    x = (synth.splitmix_uniform(3, 1280) * 32767).astype("<i2")
    buf = AudioDecoder.LoadFromStream(io.BytesIO(_wav(x.tobytes(), 1, 1, 16)))
    ctx = G.OfflineAudioContext(FS)
    s = G.AudioBufferSourceNode(ctx)
    s.Buffer = buf
    s.Connect(ctx.Destination)
    s.Start()
    y = ctx.Render(1152)
    want = x.astype(np.float32)[:1152] * np.float32(2.0 ** -15)
    assert np.array_equal(y[0], want) and np.array_equal(y[1], want)  # mono -> stereo up-mix by copy at the destination
    with pytest.raises(G.InvalidOperationException):
        AudioDecoder.LoadFromStream(io.BytesIO(b"RIFF\0\0\0\0WAVX"))
    with pytest.raises(G.InvalidOperationException):
        AudioDecoder.LoadFromStream(io.BytesIO(_wav(b"\0" * 64, 1, 2, 8)))  # 8-bit PCM is not on this path
    ctx.Dispose()


def test_write_wav_round_trip():
    import graphaudio_b200 as G
    from graphaudio_b200.io import AudioDecoder, WriteWav
    src = [synth.splitmix_uniform(40 + c, 3000) for c in range(2)]

    def build():
        ctx = G.OfflineAudioContext(FS)
        s = G.AudioBufferSourceNode(ctx)
        s.Buffer = G.PlayableAudioBuffer.FromChannelArrays(src, FS)
        g = G.GainNode(ctx)
        g.Gain.Value = 0.25
        s.Connect(g).Connect(ctx.Destination)
        s.Start()
        return ctx
    out = io.BytesIO()
    WriteWav(out, build(), 2944)
    back = AudioDecoder.LoadFromStream(io.BytesIO(out.getvalue()))
    assert np.array_equal(np.stack(back.channels), build().Render(2944))
