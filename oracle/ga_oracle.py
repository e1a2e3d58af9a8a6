"""ctypes front-end of the CPU oracle (oracle/ga_oracle.cpp).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (graphaudio_b200/) never imports it.
PARITY UNPINNED: see the header of ga_oracle.cpp.

The classes mirror the reference's public API names (GraphAudio.Core/OfflineAudioContext.cs:30,108,
AudioParam.cs:252-312, Nodes/*.cs) so that a graph-building function written once can be run
against this oracle and against graphaudio_b200 (the CUDA path).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libga_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ga_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libga_oracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    fpp = C.POINTER(C.POINTER(C.c_float))
    fp = C.POINTER(C.c_float)
    dp = C.POINTER(C.c_double)
    L.ora_rfft_forward.argtypes = [C.c_int, dp, dp, dp]
    L.ora_rfft_inverse.argtypes = [C.c_int, dp, dp, dp]
    L.ora_normalization_scale.argtypes = [fp, C.c_int64]
    L.ora_normalization_scale.restype = C.c_float
    L.ora_pc_create.argtypes = [fp, C.c_int64, C.c_int, C.c_int]
    L.ora_pc_create.restype = C.c_void_p
    L.ora_pc_destroy.argtypes = [C.c_void_p]
    L.ora_pc_partitions.argtypes = [C.c_void_p]
    L.ora_pc_ir_spectra.argtypes = [C.c_void_p, fp, fp]
    L.ora_pc_process.argtypes = [C.c_void_p, fp, fp, C.c_int, C.c_int64]
    L.ora_resample.argtypes = [fp, C.c_int64, fp, C.c_int64, C.c_double, C.POINTER(C.c_int64)]
    L.ora_resample.restype = C.c_int64
    L.ora_context_create.argtypes = [C.c_int]
    L.ora_context_create.restype = C.c_void_p
    L.ora_context_destroy.argtypes = [C.c_void_p]
    L.ora_buffer_create.argtypes = [C.c_void_p, fpp, C.c_int, C.c_int64, C.c_int]
    L.ora_node_create.argtypes = [C.c_void_p, C.c_int]
    L.ora_delay_create.argtypes = [C.c_void_p, C.c_double]
    L.ora_connect_param.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.ora_connect_io.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.ora_splitter_create.argtypes = [C.c_void_p, C.c_int]
    L.ora_merger_create.argtypes = [C.c_void_p, C.c_int]
    L.ora_scheduled_start.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
    L.ora_scheduled_stop.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.ora_oscillator_set_type.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ora_connect.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ora_disconnect.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ora_param_set_value.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float]
    L.ora_param_event.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_double, C.c_double]
    L.ora_param_cancel.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
    L.ora_param_eval.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, fp]
    L.ora_source_set_buffer.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ora_source_start.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
    L.ora_source_stop.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.ora_source_set_loop.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
    L.ora_unsupported.argtypes = [C.c_void_p]
    L.ora_biquad_set_type.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ora_convolver_set_buffer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.ora_render.argtypes = [C.c_void_p, fpp, C.c_int, C.c_int, C.c_int]
    _lib = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


# ------------------------------------------------------------------ primitives
def rfft_forward(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[0]
    re = np.empty(n // 2 + 1)
    im = np.empty(n // 2 + 1)
    if lib().ora_rfft_forward(n, _dptr(x), _dptr(re), _dptr(im)) != 0:
        raise ValueError("The FFT length must be a power of two.")
    return re + 1j * im


def rfft_inverse(spec):
    spec = np.asarray(spec, dtype=np.complex128)
    n = (spec.shape[0] - 1) * 2
    re = np.ascontiguousarray(spec.real)
    im = np.ascontiguousarray(spec.imag)
    x = np.empty(n)
    if lib().ora_rfft_inverse(n, _dptr(re), _dptr(im), _dptr(x)) != 0:
        raise ValueError("The FFT length must be a power of two.")
    return x


def normalization_scale(ir):
    ir = _f32(ir)
    return float(lib().ora_normalization_scale(_fptr(ir), ir.shape[0]))


class PartitionedConvolver:
    """GraphAudio.Core/PartitionedConvolver.cs:37,104 (internal class; exposed for kernel-level parity)."""

    def __init__(self, impulse_response, block_size=128, normalize=True):
        ir = _f32(impulse_response)
        self.block_size = block_size
        self._h = lib().ora_pc_create(_fptr(ir), ir.shape[0], block_size, int(normalize))
        self.partitions = lib().ora_pc_partitions(self._h)

    def ir_spectra(self):
        c = self.block_size + 1
        re = np.empty((self.partitions, c), np.float32)
        im = np.empty((self.partitions, c), np.float32)
        lib().ora_pc_ir_spectra(self._h, _fptr(re), _fptr(im))
        return re, im

    def process(self, x):
        """Feeds len(x)//block_size consecutive blocks through Process()."""
        x = _f32(x)
        nb = x.shape[0] // self.block_size
        out = np.empty(nb * self.block_size, np.float32)
        lib().ora_pc_process(self._h, _fptr(x), _fptr(out), self.block_size, nb)
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ora_pc_destroy(self._h)
            self._h = None


def resample(x, n_out, rate):
    """One CubicResampler (CubicResampler.cs:26-63) over the whole input; returns (out[:produced], consumed)."""
    x = _f32(x)
    out = np.zeros(n_out, np.float32)
    consumed = C.c_int64(0)
    n = lib().ora_resample(_fptr(x), x.shape[0], _fptr(out), n_out, float(rate), C.byref(consumed))
    return out[:n], consumed.value


# ------------------------------------------------------------------ graph API (reference names)
class ArgumentException(ValueError):
    pass


class ArgumentOutOfRangeException(ArgumentException):
    pass


class InvalidOperationException(RuntimeError):
    pass


class FilterType:
    Lowpass, Highpass, Bandpass, Notch, Allpass, Peaking, Lowshelf, Highshelf = range(8)


class PlayableAudioBuffer:
    def __init__(self, channels, sample_rate):
        self.channels = [_f32(c) for c in channels]
        self.SampleRate = int(sample_rate)
        self.NumberOfChannels = len(self.channels)
        self.Length = int(self.channels[0].shape[0])
        self._ids = {}

    @staticmethod
    def FromChannelArrays(channel_data, sample_rate):
        if len(channel_data) == 0:
            raise ArgumentException("Channel data cannot be empty")
        n = len(channel_data[0])
        if any(len(c) != n for c in channel_data):
            raise ArgumentException("All channels must have the same length")
        return PlayableAudioBuffer(channel_data, sample_rate)

    @staticmethod
    def FromMonoArray(data, sample_rate):
        return PlayableAudioBuffer([data], sample_rate)

    @staticmethod
    def FromStereoArrays(left, right, sample_rate):
        if len(left) != len(right):
            raise ArgumentException("Left and right channels must have the same length")
        return PlayableAudioBuffer([left, right], sample_rate)

    def _id(self, ctx):
        key = id(ctx)
        if key not in self._ids:
            ptrs = (C.POINTER(C.c_float) * self.NumberOfChannels)(*[_fptr(c) for c in self.channels])
            self._ids[key] = lib().ora_buffer_create(ctx._h, ptrs, self.NumberOfChannels, self.Length, self.SampleRate)
        return self._ids[key]


class AudioParam:
    def __init__(self, node, index, default, mn, mx):
        self._node, self._idx = node, index
        self._value = default
        self.MinValue, self.MaxValue, self.DefaultValue = mn, mx, default

    @property
    def Value(self):
        return self._value

    @Value.setter
    def Value(self, v):
        self._value = min(max(float(v), self.MinValue), self.MaxValue)
        lib().ora_param_set_value(self._node._ctx._h, self._node._id, self._idx, float(v))

    def _ev(self, typ, v, t, tc=0.0):
        r = lib().ora_param_event(self._node._ctx._h, self._node._id, self._idx, typ, float(v), float(t), float(tc))
        if r == -2:
            raise ArgumentException("Exponential ramp target must be > 0")

    def SetValueAtTime(self, value, startTime):
        self._ev(0, value, startTime)

    def LinearRampToValueAtTime(self, value, endTime):
        self._ev(1, value, endTime)

    def ExponentialRampToValueAtTime(self, value, endTime):
        self._ev(2, value, endTime)

    def SetTargetAtTime(self, target, startTime, timeConstant):
        self._ev(3, target, startTime, timeConstant)

    def CancelScheduledValues(self, cancelTime):
        lib().ora_param_cancel(self._node._ctx._h, self._node._id, self._idx, float(cancelTime))

    def evaluate(self, n_blocks):
        """Test helper: values for blocks 0..n_blocks-1 as ComputeValues would produce them."""
        out = np.empty(n_blocks * 128, np.float32)
        lib().ora_param_eval(self._node._ctx._h, self._node._id, self._idx, n_blocks, _fptr(out))
        return out


class AudioNode:
    _KIND = -1

    def __init__(self, context):
        self._ctx = context
        self.Context = context
        self._id = lib().ora_node_create(context._h, self._KIND) if self._KIND >= 0 else 0

    def Connect(self, destination, outputIndex=0, inputIndex=0):
        if outputIndex or inputIndex:  # Connect(destination, outputIndex, inputIndex) — Nodes/AudioNode.cs:68-84
            if lib().ora_connect_io(self._ctx._h, self._id, int(outputIndex), destination._id, int(inputIndex)) != 0:
                raise ArgumentOutOfRangeException("cannot connect")
            return destination
        if isinstance(destination, AudioParam):  # AudioNode.Connect(AudioParam param) — Nodes/AudioNode.cs:86-92
            if lib().ora_connect_param(self._ctx._h, self._id, destination._node._id, destination._idx) != 0:
                raise ArgumentOutOfRangeException("cannot connect to the parameter")
            return None
        if lib().ora_connect(self._ctx._h, self._id, destination._id) != 0:
            raise ArgumentOutOfRangeException("cannot connect")
        return destination


    def Disconnect(self, destination=None):  # Nodes/AudioNode.cs:78-84
        lib().ora_disconnect(self._ctx._h, self._id, destination._id if destination is not None else -1)


class AudioDestinationNode(AudioNode):
    pass


class AudioBufferSourceNode(AudioNode):
    _KIND = 0

    def __init__(self, context):
        super().__init__(context)
        self.PlaybackRate = AudioParam(self, 0, 1.0, 0.001, 1000.0)
        self._buffer = None
        self._loop, self._loop_start, self._loop_end = False, 0.0, 0.0

    @property
    def Buffer(self):
        return self._buffer

    @Buffer.setter
    def Buffer(self, b):
        self._buffer = b
        lib().ora_source_set_buffer(self._ctx._h, self._id, b._id(self._ctx))

    def _set_loop(self):
        lib().ora_source_set_loop(self._ctx._h, self._id, int(self._loop), self._loop_start, self._loop_end)

    @property
    def Loop(self):
        return self._loop

    @Loop.setter
    def Loop(self, v):
        self._loop = bool(v)
        self._set_loop()

    @property
    def LoopStart(self):
        return self._loop_start

    @LoopStart.setter
    def LoopStart(self, v):
        self._loop_start = max(0.0, float(v))
        self._set_loop()

    @property
    def LoopEnd(self):
        return self._loop_end

    @LoopEnd.setter
    def LoopEnd(self, v):
        self._loop_end = max(0.0, float(v))
        self._set_loop()

    def Start(self, when=0.0, offset=0.0, duration=math.inf):
        if lib().ora_source_start(self._ctx._h, self._id, float(when), float(offset), float(duration)) != 0:
            raise InvalidOperationException("AudioBufferSourceNode can only be started once, with a buffer set.")

    def Stop(self, when=0.0):
        lib().ora_source_stop(self._ctx._h, self._id, float(when))


class BiQuadFilterNode(AudioNode):
    _KIND = 1

    def __init__(self, context):
        super().__init__(context)
        self.Frequency = AudioParam(self, 0, 1000.0, 1.0, context.SampleRate / 2.0)
        self.Q = AudioParam(self, 1, 1.0, 0.001, 1000.0)
        self.Gain = AudioParam(self, 2, 0.0, -60.0, 60.0)
        self._type = FilterType.Lowpass

    @property
    def Type(self):
        return self._type

    @Type.setter
    def Type(self, t):
        self._type = t
        lib().ora_biquad_set_type(self._ctx._h, self._id, int(t))


class GainNode(AudioNode):
    _KIND = 2

    def __init__(self, context):
        super().__init__(context)
        self.Gain = AudioParam(self, 0, 1.0, -3.4028234663852886e38, 3.4028234663852886e38)


class ConvolverNode(AudioNode):
    _KIND = 3

    def __init__(self, context):
        super().__init__(context)
        self.Normalize = True
        self.EnableTrueStereo = True
        self._buffer = None

    @property
    def Buffer(self):
        return self._buffer

    @Buffer.setter
    def Buffer(self, b):
        if b is self._buffer:
            return
        if b is None:
            lib().ora_convolver_set_buffer(self._ctx._h, self._id, -1, 0, 0)
            self._buffer = None
            return
        r = lib().ora_convolver_set_buffer(self._ctx._h, self._id, b._id(self._ctx), int(self.Normalize), int(self.EnableTrueStereo))
        if r != 0:
            raise InvalidOperationException("Impulse response buffer sample rate must match the audio context sample rate.")
        self._buffer = b


class StereoPannerNode(AudioNode):
    """Nodes/StereoPannerNode.cs"""
    _KIND = 4

    def __init__(self, context):
        super().__init__(context)
        self.Pan = AudioParam(self, 0, 0.0, -1.0, 1.0)


class OscillatorType:  # Nodes/OscillatorNode.cs:207-213
    Sine, Square, Sawtooth, Triangle = range(4)


class _ScheduledSource(AudioNode):
    def Start(self, when=0.0, offset=0.0, duration=math.nan):
        if lib().ora_scheduled_start(self._ctx._h, self._id, float(when), float(duration)) != 0:
            raise InvalidOperationException("The node can only be started once.")

    def Stop(self, when=0.0):
        lib().ora_scheduled_stop(self._ctx._h, self._id, float(when))


class OscillatorNode(_ScheduledSource):
    """Nodes/OscillatorNode.cs (oracle only so far: the device path is SURVEY.md §8f-3 "next")"""
    _KIND = 5

    def __init__(self, context):
        super().__init__(context)
        self.Frequency = AudioParam(self, 0, 440.0, 0.0, context.SampleRate / 2.0)
        self._type = OscillatorType.Sine

    @property
    def Type(self):
        return self._type

    @Type.setter
    def Type(self, t):
        if lib().ora_oscillator_set_type(self._ctx._h, self._id, int(t)) != 0:
            raise ArgumentOutOfRangeException("Type")
        self._type = t


class ConstantSourceNode(_ScheduledSource):
    """Nodes/ConstantSourceNode.cs (oracle only so far)"""
    _KIND = 6

    def __init__(self, context):
        super().__init__(context)
        self.Offset = AudioParam(self, 0, 1.0, -3.4028234663852886e38, 3.4028234663852886e38)


class ChannelSplitterNode(AudioNode):
    """Nodes/ChannelSplitterNode.cs (oracle only so far)"""

    def __init__(self, context, numberOfOutputs=2):
        if numberOfOutputs < 1 or numberOfOutputs > 32:
            raise ArgumentOutOfRangeException("numberOfOutputs")
        self._ctx = context
        self.Context = context
        self._id = lib().ora_splitter_create(context._h, int(numberOfOutputs))


class ChannelMergerNode(AudioNode):
    """Nodes/ChannelMergerNode.cs (oracle only so far)"""

    def __init__(self, context, numberOfInputs=2):
        if numberOfInputs < 1 or numberOfInputs > 32:
            raise ArgumentOutOfRangeException("numberOfInputs")
        self._ctx = context
        self.Context = context
        self._id = lib().ora_merger_create(context._h, int(numberOfInputs))


class DelayNode(AudioNode):
    """Nodes/DelayNode.cs"""

    def __init__(self, context, maxDelayTime=1.0):
        if maxDelayTime <= 0 or maxDelayTime > 10:
            raise ArgumentOutOfRangeException("maxDelayTime")  # :25-26
        self._ctx = context
        self.Context = context
        self._id = lib().ora_delay_create(context._h, float(maxDelayTime))
        self.DelayTime = AudioParam(self, 0, 0.0, 0.0, float(np.float32(maxDelayTime)))


class OfflineAudioContext:
    def __init__(self, sampleRate=48000):
        if sampleRate <= 0:
            raise ArgumentOutOfRangeException("sampleRate")
        self.SampleRate = int(sampleRate)
        self._h = lib().ora_context_create(self.SampleRate)
        self.Destination = AudioDestinationNode(self)

    def Render(self, output_or_count, frameCount=None, startIndex=0):
        """Render(float[][] output, int frameCount, int startIndex=0) or Render(int frameCount)."""
        if frameCount is None:
            n = int(output_or_count)
            if n <= 0:
                raise ArgumentOutOfRangeException("Frame count must be positive.")
            out = np.zeros((2, n), np.float32)
            self.Render(out, n, 0)
            return out
        out = output_or_count
        if len(out) == 0:
            raise ArgumentException("Output buffer must have at least one channel.")
        if frameCount <= 0:
            raise ArgumentOutOfRangeException("Frame count must be positive.")
        if startIndex < 0:
            raise ArgumentOutOfRangeException("Start index must be non-negative.")
        rows = [out[c] for c in range(len(out))]
        for r in rows:
            if r.dtype != np.float32 or not r.flags.c_contiguous:
                raise ArgumentException("channels must be contiguous float32")
            if r.shape[0] < startIndex + frameCount:
                raise ArgumentException("Channel buffer is too small.")
        ptrs = (C.POINTER(C.c_float) * len(rows))(*[_fptr(r) for r in rows])
        rc = lib().ora_render(self._h, ptrs, len(rows), int(frameCount), int(startIndex))
        if rc == -2:
            raise InvalidOperationException("Audio graph cycle detected")
        if lib().ora_unsupported(self._h):
            raise NotImplementedError("the oracle does not restate this path (resampled looping source with an empty loop region: the reference spins forever)")
        if rc != 0:
            raise ArgumentOutOfRangeException("render failed (%d)" % rc)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ora_context_destroy(self._h)
            self._h = None
