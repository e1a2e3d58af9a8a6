"""Host side of the steps either side of the render path, mirroring GraphAudio.IO:

  AudioDecoder.LoadFromFile / LoadFromStream  ≙ GraphAudio.IO/LibsndfileDecoder.cs:185-220 — here a RIFF/WAVE container parser
      (libsndfile is not in this image; WAV PCM 16/24/32 and IEEE float 32 cover the formats its sf_readf_float path normalises
      the same way).  Only the container is parsed on the host: the interleaved samples go to the device as they are and are
      converted and de-interleaved there (gac_buffer_create_interleaved), so 16-bit material costs half the PCIe bytes.
  WriteWav                                    ≙ a consumer of ProcessBlockInterleaved (AudioContextBase.cs:88-161): the render is
      interleaved on the device (gac_render_interleaved) and written as IEEE-float WAV.
"""
from __future__ import annotations

import io
import struct

import numpy as np

from . import _native as N
from .api import InvalidOperationException, PlayableAudioBuffer

_KSDATAFORMAT_PCM, _KSDATAFORMAT_FLOAT = 1, 3


class AudioDecoder:
    @staticmethod
    def LoadFromFile(filePath) -> PlayableAudioBuffer:
        with open(filePath, "rb") as f:
            return AudioDecoder.LoadFromStream(f)

    @staticmethod
    def LoadFromStream(stream) -> PlayableAudioBuffer:
        data = stream.read()
        if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
            raise InvalidOperationException("not a RIFF/WAVE stream")
        pos, fmt, payload = 12, None, None
        while pos + 8 <= len(data):
            tag, size = data[pos:pos + 4], struct.unpack_from("<I", data, pos + 4)[0]
            body = data[pos + 8:pos + 8 + size]
            if tag == b"fmt ":
                code, channels, rate, _, _, bits = struct.unpack_from("<HHIIHH", body, 0)
                if code == 0xFFFE and len(body) >= 26:  # WAVE_FORMAT_EXTENSIBLE: the sub-format GUID starts with the real code
                    code = struct.unpack_from("<H", body, 24)[0]
                fmt = (code, channels, rate, bits)
            elif tag == b"data":
                payload = body
            pos += 8 + size + (size & 1)
        if fmt is None or payload is None:
            raise InvalidOperationException("WAVE stream without fmt / data chunk")
        code, channels, rate, bits = fmt
        kind = {(_KSDATAFORMAT_PCM, 16): N.GAC_SAMPLE_S16, (_KSDATAFORMAT_PCM, 24): N.GAC_SAMPLE_S24,
                (_KSDATAFORMAT_PCM, 32): N.GAC_SAMPLE_S32, (_KSDATAFORMAT_FLOAT, 32): N.GAC_SAMPLE_F32}.get((code, bits))
        if kind is None:
            raise InvalidOperationException(f"unsupported WAVE sample format (code {code}, {bits} bits)")
        frames = len(payload) // (channels * bits // 8)
        if frames <= 0:  # LibsndfileDecoder.cs:200-201
            raise InvalidOperationException(f"Invalid audio duration or frame count: {frames}")
        return PlayableAudioBuffer.FromInterleaved(np.frombuffer(payload, np.uint8), kind, channels, rate)


def WriteWav(path_or_stream, context, frameCount, channels=2):
    """Renders `frameCount` frames of `context` interleaved on the device and writes them as a 32-bit float WAV."""
    inter = context.RenderInterleaved(frameCount, channels)
    payload = inter.astype("<f4").tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(payload)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, _KSDATAFORMAT_FLOAT, channels, context.SampleRate, context.SampleRate * channels * 4, channels * 4, 32)
    hdr += b"data" + struct.pack("<I", len(payload))
    if isinstance(path_or_stream, (str, bytes)):
        with open(path_or_stream, "wb") as f:
            f.write(hdr + payload)
    else:
        path_or_stream.write(hdr + payload)
    return inter
