"""Host-side mirror of the reference's public API for the offline render path, over the C ABI.

Same type and member names as GraphAudio.Core (OfflineAudioContext.cs:18,30,108; AudioParam.cs:252-312;
Nodes/AudioBufferSourceNode.cs:32,67,79,116; Nodes/BiQuadFilterNode.cs:21,42,47,52; Nodes/GainNode.cs:14;
Nodes/ConvolverNode.cs:25,87,95; Nodes/AudioNode.cs:68), so that graph-building code written for the reference
reads the same here.  The nodes only RECORD topology and automation; `OfflineAudioContext.Render` flattens the
recorded graph into gac_voice_desc / gac_bus_desc arrays and makes ONE call into libgraphaudio_cuda.so.

This module is the Python twin of the C++ mirror in graphaudio_b200/host/graphaudio_cuda.hpp and of the C#
package sketched in INTEGRATION.md; pytest drives the library through it.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C
import itertools
import math
import struct
from typing import List, Optional

import numpy as np

from . import _native as N


# ---- exception types the reference throws, keyed by gac_status ------------------------------------------
class ArgumentException(ValueError):
    pass


class ArgumentOutOfRangeException(ArgumentException):
    pass


class InvalidOperationException(RuntimeError):
    pass


class ObjectDisposedException(InvalidOperationException):
    pass


class NotSupportedException(RuntimeError):
    """Graph shape outside the accelerated hot path."""


class CudaException(RuntimeError):
    pass


_ERR = {
    N.GAC_ERR_INVALID_ARGUMENT: ArgumentException,
    N.GAC_ERR_OUT_OF_RANGE: ArgumentOutOfRangeException,
    N.GAC_ERR_INVALID_OPERATION: InvalidOperationException,
    N.GAC_ERR_DISPOSED: ObjectDisposedException,
    N.GAC_ERR_UNSUPPORTED: NotSupportedException,
}


def check(rc: int):
    if rc != N.GAC_OK:
        raise _ERR.get(rc, CudaException)(f"[gac {rc}] {N.last_error()}")


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fptr(a):
    return a.ctypes.data_as(N.fp)


_context_serial = itertools.count(1)


class FilterType:  # Nodes/BiQuadFilterNode.cs:288-298
    Lowpass, Highpass, Bandpass, Notch, Allpass, Peaking, Lowshelf, Highshelf = range(8)


class PlayableAudioBuffer:
    """PlayableAudioBuffer.cs — planar float32 sample container; uploaded to HBM on first use by a context."""

    def __init__(self, channels, sample_rate):
        if len(channels) < 1 or len(channels) > 32:
            raise ArgumentOutOfRangeException("Channel count must be between 1 and 32")
        if sample_rate <= 0:
            raise ArgumentOutOfRangeException("Sample rate must be positive")
        self._channels = [_f32(c) for c in channels]
        self._raw = None  # (interleaved samples as uint8, gac_sample_format) when the buffer came from a decoder (FromInterleaved)
        self.SampleRate = int(sample_rate)
        self.NumberOfChannels = len(self._channels)
        self.Length = int(self._channels[0].shape[0])
        self._handles = {}

    _SAMPLE_BYTES = {N.GAC_SAMPLE_S16: 2, N.GAC_SAMPLE_S24: 3, N.GAC_SAMPLE_S32: 4, N.GAC_SAMPLE_F32: 4}

    @staticmethod
    def FromInterleaved(samples, sampleFormat, numberOfChannels, sampleRate):
        """The buffer a decoder produces (GraphAudio.IO AudioDecoder.LoadFromStream, LibsndfileDecoder.cs:195-220), kept as the
        file's interleaved little-endian samples: they are uploaded as they are and converted + de-interleaved on the device
        (gac_buffer_create_interleaved).  `samples`: bytes-like / uint8 array."""
        if numberOfChannels < 1 or numberOfChannels > 32:
            raise ArgumentOutOfRangeException("Channel count must be between 1 and 32")
        if sampleRate <= 0:
            raise ArgumentOutOfRangeException("Sample rate must be positive")
        if sampleFormat not in PlayableAudioBuffer._SAMPLE_BYTES:
            raise ArgumentException("unknown sample format")
        raw = samples if isinstance(samples, np.ndarray) else np.frombuffer(samples, np.uint8)
        raw = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
        frame_bytes = PlayableAudioBuffer._SAMPLE_BYTES[sampleFormat] * numberOfChannels
        b = PlayableAudioBuffer.__new__(PlayableAudioBuffer)
        b._channels = None
        b.SampleRate, b.NumberOfChannels, b.Length = int(sampleRate), int(numberOfChannels), int(raw.shape[0] // frame_bytes)
        b._raw = (raw[:b.Length * frame_bytes], int(sampleFormat))
        b._handles = {}
        return b

    @property
    def channels(self):
        """Planar float32 channel data (GetChannelData, PlayableAudioBuffer.cs:72); converted on the host on first access for
        buffers that hold interleaved file samples."""
        if self._channels is None:
            raw, fmt = self._raw
            n, c = self.Length, self.NumberOfChannels
            if fmt == N.GAC_SAMPLE_S16:
                x = raw.view("<i2").astype(np.float32) * np.float32(2.0 ** -15)
            elif fmt == N.GAC_SAMPLE_S24:
                t = raw.reshape(-1, 3).astype(np.int32)
                x = (((t[:, 0] << 8) | (t[:, 1] << 16) | (t[:, 2] << 24)) >> 8).astype(np.float32) * np.float32(2.0 ** -23)
            elif fmt == N.GAC_SAMPLE_S32:
                x = raw.view("<i4").astype(np.float32) * np.float32(2.0 ** -31)
            else:
                x = raw.view("<f4").astype(np.float32)
            self._channels = [np.ascontiguousarray(x.reshape(n, c)[:, k]) for k in range(c)]
        return self._channels

    @staticmethod
    def FromChannelArrays(channelData, sampleRate):  # :122-143
        if len(channelData) == 0:
            raise ArgumentException("Channel data cannot be or empty")
        n = len(channelData[0])
        if any(len(c) != n for c in channelData):
            raise ArgumentException("All channels must have the same length")
        return PlayableAudioBuffer(channelData, sampleRate)

    @staticmethod
    def FromMonoArray(audioData, sampleRate):  # :148-157
        return PlayableAudioBuffer([audioData], sampleRate)

    @staticmethod
    def FromStereoArrays(leftChannel, rightChannel, sampleRate):  # :162-173
        if len(leftChannel) != len(rightChannel):
            raise ArgumentException("Left and right channels must have the same length")
        return PlayableAudioBuffer([leftChannel, rightChannel], sampleRate)

    def _handle(self, ctx: "OfflineAudioContext", member: int = 0):
        """Device copy of the buffer in `ctx` (member `member` of a multi-GPU context)."""
        ctx = ctx._root()  # forks share the parent's device handle
        # keyed by the context's serial number, never by id(): CPython reuses ids, and a disposed context's handles are freed
        key = (ctx._serial, member)
        h = self._handles.get(key)
        if h is None:
            out = C.c_void_p()
            ch = ctx._member_handle(member)
            if self._raw is not None:
                raw, fmt = self._raw
                check(N.lib().gac_buffer_create_interleaved(ch, raw.ctypes.data_as(C.c_void_p), fmt, self.NumberOfChannels, self.Length,
                                                            self.SampleRate, C.byref(out)))
            else:
                ptrs = (N.fp * self.NumberOfChannels)(*[_fptr(c) for c in self.channels])
                check(N.lib().gac_buffer_create(ch, ptrs, self.NumberOfChannels, self.Length, self.SampleRate, C.byref(out)))
            h = out.value
            self._handles[key] = h
            ctx._owned_buffers.append(h)
            ctx._buffer_objects.append(self)
        return h


class AudioParam:
    """AudioParam.cs — value + time-sorted automation events (clamped at schedule time, :254,268,282,299)."""

    def __init__(self, default, mn, mx):
        self.DefaultValue, self.MinValue, self.MaxValue = float(default), float(mn), float(mx)
        r = AudioParam._RANGES.get((default, mn, mx))
        if r is None:  # (default, min, max) rounded to float32, once per distinct triple (Math.Clamp works on floats)
            r = AudioParam._RANGES[(default, mn, mx)] = (float(np.float32(default)), float(np.float32(mn)), float(np.float32(mx)))
        self._value, self._mn32, self._mx32 = r  # Python floats that hold float32 values
        self._events: List[tuple] = []  # (type, value, target, time, time_constant)
        # (first quantum, value, events) as each Render call saw the parameter: edits made between successive Render calls act from
        # the next unprocessed quantum on (OfflineAudioContext.cs:55-100), earlier quanta keep what they were rendered with
        self._epochs: List[tuple] = []
        self._input_node = None  # the parameter's own fan-in (AudioParam.cs:60-62), created by the first AudioNode.Connect(param)

    def _clamp(self, v):
        v = float(np.float32(v))
        return self._mn32 if v < self._mn32 else (self._mx32 if v > self._mx32 else v)

    @property
    def Value(self):
        return float(self._value)

    @Value.setter
    def Value(self, v):  # :34-49: clamps and clears every scheduled event
        self._value = self._clamp(v)
        self._events = []

    def _add(self, ev):  # AddEvent :333-352, stable upper-bound insert
        if not self._events or not ev[3] < self._events[-1][3]:  # (the usual case: scheduled in time order)
            self._events.append(ev)
            return
        lo, hi = 0, len(self._events)
        while lo < hi:
            mid = (lo + hi) >> 1
            if ev[3] < self._events[mid][3]:
                hi = mid
            else:
                lo = mid + 1
        self._events.insert(lo, ev)

    def SetValueAtTime(self, value, startTime):
        self._add((0, float(self._clamp(value)), 0.0, float(startTime), 0.0))

    def LinearRampToValueAtTime(self, value, endTime):
        self._add((1, float(self._clamp(value)), 0.0, float(endTime), 0.0))

    def ExponentialRampToValueAtTime(self, value, endTime):
        v = self._clamp(value)
        if v <= 0:
            raise ArgumentException("Exponential ramp target must be > 0")  # :283-284
        self._add((2, float(v), 0.0, float(endTime), 0.0))

    def SetTargetAtTime(self, target, startTime, timeConstant):
        self._add((3, 0.0, float(self._clamp(target)), float(startTime), float(timeConstant)))

    def CancelScheduledValues(self, cancelTime):  # :312-331
        self._events = [e for e in self._events if e[3] < cancelTime]

    def _commit(self, q_now: int):
        """Called when a graph is flattened for a render that starts at quantum q_now."""
        state = (self._value, tuple(self._events))
        if not self._epochs:
            self._epochs = [(0,) + state]
        elif state != self._epochs[-1][1:]:
            if q_now > self._epochs[-1][0]:
                self._epochs.append((q_now,) + state)
            else:  # edited again before anything further was rendered
                self._epochs[-1] = (self._epochs[-1][0],) + state

    _RANGES = {}
    _EVENT_ARRAYS = {}  # n -> (gac_event * n, struct.Struct of n events): array types and packers are built once per length

    def _desc(self, keep: list, q_now: int = 0) -> N.gac_param:
        mod_bus = 0
        if self._input_node is not None and self._input_node._in:
            mod_bus = self._input_node._bus_index + 1  # (set by the flattening of the graph this parameter belongs to)
        eps = self._epochs
        if not self._events and (not eps or (len(eps) == 1 and eps[0][1] == self._value and not eps[0][2])):
            # a plain value that was never scheduled or edited (most parameters of a graph): nothing to flatten
            if not eps:
                self._epochs = [(0, self._value, ())]
            return N.gac_param(self._value, 0, None, mod_bus, self._mn32, self._mx32)
        self._commit(q_now)
        eps = self._epochs
        if len(eps) == 1:
            flat = eps[0][2]
        else:
            flat = []
            for k, (q0, value, events) in enumerate(eps):
                if k > 0:
                    flat.append((N.GAC_EVENT_EPOCH, value, 0.0, 0.0, float(q0)))
                flat.extend(events)
        arr = None
        n = len(flat)
        if n:
            ta = AudioParam._EVENT_ARRAYS.get(n)
            if ta is None:
                ta = AudioParam._EVENT_ARRAYS[n] = (N.gac_event * n, struct.Struct("<" + "iffxxxxdd" * n))
            # (type, value, target, time, time_constant): the struct's field order, packed in one call (a ctypes array built from
            # tuples converts field by field: 5 us instead of 1.3 for four events, ~400 parameters per 128-voice graph)
            arr = ta[0].from_buffer_copy(ta[1].pack(*[x for e in flat for x in e]))
            keep.append(arr)
        # positional initialiser: value, n_events, events, mod_bus, min_value, max_value
        return N.gac_param(eps[0][1], n, arr, mod_bus, self._mn32, self._mx32)


class AudioNode:
    """Nodes/AudioNode.cs — records connections; Connect returns the destination to allow chaining (:68-73)."""
    _is_source = False   # AudioBufferSourceNode / scheduled sources: the head of a voice
    _force_bus = False   # nodes that always start a bus, whatever their number of inputs (set per instance)
    _accelerated = False  # node types the device path renders (anything else must never render as silence: the graph is refused)

    def __init__(self, context: "OfflineAudioContext", n_inputs=1, n_outputs=1):
        self.Context = context
        self._in: List["AudioNode"] = []   # upstream nodes in connection order (AudioNodeInput._connectedOutputs)
        self._out: List["AudioNode"] = []
        self._n_inputs, self._n_outputs = n_inputs, n_outputs
        self._born_frames = getattr(context, "_frames_rendered", 0)  # frames the context had rendered when the node was created
        self._edge_on = {}    # id(destination) -> first quantum of the connection (0: before anything was rendered)
        self._edge_off = {}   # id(destination) -> quantum from which the connection is gone (Disconnect after rendering began)
        self._first_live_q = None  # first quantum in which the node was pulled (set when a graph that contains it is flattened)
        context._nodes.append(self)

    def _existed_in_a_render(self):
        return self._born_frames < getattr(self.Context, "_frames_rendered", 0)

    def _upstream_existed_in_a_render(self):
        seen, stack = set(), [self]
        while stack:
            n = stack.pop()
            if id(n) in seen:
                continue
            seen.add(id(n))
            if n._existed_in_a_render():
                return True
            stack.extend(n._in)
        return False

    def Connect(self, destination: "AudioNode", outputIndex=0, inputIndex=0):
        if outputIndex < 0 or outputIndex >= self._n_outputs:
            raise ArgumentOutOfRangeException("outputIndex")
        src = self._output_node(outputIndex)
        if isinstance(destination, AudioParam):  # AudioNode.Connect(AudioParam) — Nodes/AudioNode.cs:86-92
            if destination._input_node is None:
                destination._input_node = _ParamInputNode(self.Context)
            if src is not self:
                return src.Connect(destination)
            if destination._input_node not in self._out:
                self._out.append(destination._input_node)
                destination._input_node._in.append(self)
            return None
        if inputIndex < 0 or inputIndex >= destination._n_inputs:
            raise ArgumentOutOfRangeException("inputIndex")
        if destination is self:
            raise InvalidOperationException("Cannot connect a node to itself")  # AudioNodeOutput.cs:43-44
        if src is not self:  # an output of a ChannelSplitterNode: the connection starts at that output's own node
            return src.Connect(destination, 0, inputIndex)
        if isinstance(destination, ChannelMergerNode):
            destination._slots[id(self)] = inputIndex + 1
        if destination not in self._out:
            # Between successive Render calls the device path re-renders the timeline from frame 0 with the CURRENT graph.  A
            # connection made after rendering began acts from the next unprocessed quantum on (Nodes/AudioNode.cs:109-123 posts it to
            # the render thread): the edge carries that quantum and is flattened into a GAC_OP_GATE.
            self._out.append(destination)
            destination._in.append(self)
            self._edge_on[id(destination)] = self.Context._q_now()
        elif id(destination) in self._edge_off:
            # the reference would resume nodes that were not pulled in between where they stopped: not reproduced
            self.Context._unsupported_edit = "re-connecting a connection that was removed after rendering began"
        return destination

    def _output_node(self, outputIndex):
        return self

    def _params(self):
        # a node's parameters are created by its constructor, under the same attribute names for every instance of the class: the
        # names are found once per class (scanning __dict__ per node cost 1 ms per 128-voice graph)
        cls = type(self)
        names = cls.__dict__.get("_param_names")
        if names is None:
            names = tuple(k for k, v in self.__dict__.items() if isinstance(v, AudioParam))
            cls._param_names = names
        if not names:
            return ()
        d = self.__dict__
        return [d[k] for k in names]

    def Disconnect(self, destination: Optional["AudioNode"] = None):
        q = self.Context._q_now()
        for d in ([destination] if destination is not None else list(self._out)):
            if d in self._out and id(d) not in self._edge_off:
                if q > self._edge_on.get(id(d), 0):
                    # rendered quanta keep the connection; it ends with the next unprocessed quantum (Nodes/AudioNode.cs:125-147).  What
                    # hangs on it alone is no longer pulled from then on, which a cut reproduces as long as it is not connected again.
                    self._edge_off[id(d)] = q
                else:  # nothing was rendered with this connection
                    self._out.remove(d)
                    d._in.remove(self)
                    self._edge_on.pop(id(d), None)

    def _gates(self, d):
        """GAC_OP_GATE pseudo-nodes for the connection self -> d (none for a connection that always existed)."""
        g = []
        if self._edge_on.get(id(d), 0) > 0:
            g.append(_Gate(0, self._edge_on[id(d)]))
        if id(d) in self._edge_off:
            g.append(_Gate(1, self._edge_off[id(d)]))
        return g


class AudioDestinationNode(AudioNode):
    _accelerated = True
    def __init__(self, context):
        super().__init__(context, 1, 0)


class AudioBufferSourceNode(AudioNode):
    _accelerated = True
    _is_source = True
    def __init__(self, context):
        super().__init__(context, 0, 1)
        self.PlaybackRate = AudioParam(1.0, 0.001, 1000.0)  # k-rate, Nodes/AudioBufferSourceNode.cs:76
        self._buffer: Optional[PlayableAudioBuffer] = None
        self.Loop = False            # :40-44 (rate 1: in-place wrap; any other effective rate: the wrap-buffer resampler path :236-358)
        self._loop_start = 0.0
        self._loop_end = 0.0
        self._started = False
        self._when = math.nan
        self._start_frames = 0
        self._offset = 0.0
        self._duration = math.inf
        self._stop = math.nan

    @property
    def Buffer(self):  # :67-71
        return self._buffer

    @Buffer.setter
    def Buffer(self, value):
        self._buffer = value
        # upload now rather than at Render: with async_upload the copy engine works while the rest of the graph is built
        # (a multi-GPU context uploads at Render, when it knows which device renders the voice)
        if value is not None and not self.Context._record_only and self.Context._h is not None and self.Context._root()._group is None:
            value._handle(self.Context)

    @property
    def LoopStart(self):  # :49-53
        return self._loop_start

    @LoopStart.setter
    def LoopStart(self, v):
        self._loop_start = max(0.0, float(v))

    @property
    def LoopEnd(self):  # :58-62 (0 = end of the buffer)
        return self._loop_end

    @LoopEnd.setter
    def LoopEnd(self, v):
        self._loop_end = max(0.0, float(v))

    def Start(self, when=0.0, offset=0.0, duration=math.inf):  # :79-114
        if self._started:
            raise InvalidOperationException("AudioBufferSourceNode can only be started once.")
        if self.Buffer is None:
            raise InvalidOperationException("Cannot start without a buffer set")
        self._started = True
        self._when, self._offset, self._duration = float(when), float(offset), float(duration)
        self._start_frames = getattr(self.Context, "_frames_rendered", 0)  # Start() between Render calls acts from the next quantum on

    def Stop(self, when=0.0):  # :116-129
        # called between Render calls it cannot silence quanta that were already rendered: the stop time is at least the start
        # time of the next unprocessed quantum (the reference tests t0 < stopTime per block, :139)
        at = max(0.0, float(when), self.Context._block_time(self.Context._q_now()))
        if math.isnan(self._stop):
            self._stop = at
        else:
            self._stop = min(self._stop, at)


class BiQuadFilterNode(AudioNode):
    _accelerated = True
    def __init__(self, context):
        super().__init__(context)
        self._type = FilterType.Lowpass
        self.Frequency = AudioParam(1000.0, 1.0, context.SampleRate / 2.0)  # :63-68
        self.Q = AudioParam(1.0, 0.001, 1000.0)                             # :70-75
        self.Gain = AudioParam(0.0, -60.0, 60.0)                            # :77-82 (k-rate)


    @property
    def Type(self):  # :42-52
        return self._type

    @Type.setter
    def Type(self, value):
        if value != self._type and self._existed_in_a_render():
            self.Context._unsupported_edit = "BiQuadFilterNode.Type changed after the node took part in a Render call"
        self._type = value


class GainNode(AudioNode):
    _accelerated = True
    def __init__(self, context):
        super().__init__(context)
        fmax = float(np.finfo(np.float32).max)
        self.Gain = AudioParam(1.0, -fmax, fmax)  # Nodes/GainNode.cs:19-24


class DelayNode(AudioNode):
    _accelerated = True
    """Nodes/DelayNode.cs — integer-sample delay, a-rate DelayTime in seconds."""

    def __init__(self, context, maxDelayTime=1.0):
        if maxDelayTime <= 0 or maxDelayTime > 10:  # :25-26
            raise ArgumentOutOfRangeException("maxDelayTime")
        super().__init__(context)
        self.MaxDelayTime = float(maxDelayTime)
        self.DelayTime = AudioParam(0.0, 0.0, float(np.float32(maxDelayTime)))  # :35-40


class StereoPannerNode(AudioNode):
    _accelerated = True
    """Nodes/StereoPannerNode.cs — equal-power pan, a-rate Pan in [-1, 1]."""

    def __init__(self, context):
        super().__init__(context)
        self.Pan = AudioParam(0.0, -1.0, 1.0)  # :28-33


class OscillatorType:  # Nodes/OscillatorNode.cs:207-213
    Sine, Square, Sawtooth, Triangle = range(4)


class _Gate:
    """A connection window in a flattened chain (GAC_OP_GATE): kind 0 = from quantum q on, 1 = until quantum q."""

    def __init__(self, kind, q):
        self.kind, self.q = int(kind), int(q)

    def __eq__(self, other):
        return isinstance(other, _Gate) and (self.kind, self.q) == (other.kind, other.q)

    def __repr__(self):
        return f"_Gate({'until' if self.kind else 'from'} {self.q})"


class _ParamInputNode(AudioNode):
    _accelerated = True
    """The AudioNodeInput an AudioParam owns (Explicit, one channel: AudioParam.cs:60-62): what AudioNode.Connect(param) connects
    to.  Flattened into a GAC_BUS_MONO_INPUT bus that feeds nothing but the parameter."""

    def __init__(self, context):
        super().__init__(context, 1, 0)
        self._force_bus = True
        self._bus_index = -1


class _ScheduledSourceNode(AudioNode):
    _accelerated = True
    _is_source = True
    """IAudioScheduledSourceNode: Start(when, offset, duration = NaN) / Stop(when), sample-accurate inside a quantum
    (Nodes/OscillatorNode.cs:54-89, Nodes/ConstantSourceNode.cs:38-73)."""

    def __init__(self, context):
        super().__init__(context, 0, 1)
        self._started = False
        self._when, self._duration, self._stop = math.nan, math.nan, math.nan
        self._start_frames = 0

    def Start(self, when=0.0, offset=0.0, duration=math.nan):
        if self._started:
            raise InvalidOperationException(f"{type(self).__name__} can only be started once.")
        self._started = True
        self._when, self._duration = float(when), float(duration)
        self._start_frames = getattr(self.Context, "_frames_rendered", 0)

    def Stop(self, when=0.0):
        at = max(0.0, float(when), self.Context._block_time(self.Context._q_now()))
        self._stop = at if math.isnan(self._stop) else min(self._stop, at)


class OscillatorNode(_ScheduledSourceNode):  # Nodes/OscillatorNode.cs
    def __init__(self, context):
        super().__init__(context)
        self.Type = OscillatorType.Sine
        self.Frequency = AudioParam(440.0, 0.0, context.SampleRate / 2.0)


class ConstantSourceNode(_ScheduledSourceNode):  # Nodes/ConstantSourceNode.cs
    def __init__(self, context):
        super().__init__(context)
        fmax = float(np.finfo(np.float32).max)
        self.Offset = AudioParam(1.0, -fmax, fmax)


class _SplitterOutput(AudioNode):
    _accelerated = True
    """Output `index` of a ChannelSplitterNode: channel `index` of the splitter's input as a one-channel signal (GAC_OP_CHANNEL)."""

    def __init__(self, context, index):
        super().__init__(context)
        self.Index = int(index)


class ChannelSplitterNode(AudioNode):
    _accelerated = True
    """Nodes/ChannelSplitterNode.cs.  Flattened as: the splitter's input becomes a bus (no ops), every used output a chain fed
    by that bus that starts with GAC_OP_CHANNEL."""

    def __init__(self, context, numberOfOutputs=2):
        if numberOfOutputs < 1 or numberOfOutputs > 32:
            raise ArgumentOutOfRangeException("numberOfOutputs")
        super().__init__(context, 1, numberOfOutputs)
        self._force_bus = True
        self._outputs = {}

    def _output_node(self, outputIndex):
        node = self._outputs.get(outputIndex)
        if node is None:
            node = self._outputs[outputIndex] = _SplitterOutput(self.Context, outputIndex)
            self._out.append(node)
            node._in.append(self)
        return node


class ChannelMergerNode(AudioNode):
    _accelerated = True
    """Nodes/ChannelMergerNode.cs: output channel i = channel 0 of what input i mixes.  Flattened into a bus whose inputs carry
    their merger input (gac_bus_desc.input_slots); two inputs at most on the accelerated path."""

    def __init__(self, context, numberOfInputs=2):
        if numberOfInputs < 1 or numberOfInputs > 32:
            raise ArgumentOutOfRangeException("numberOfInputs")
        super().__init__(context, numberOfInputs, 1)
        self._force_bus = True
        self._slots = {}  # id(upstream node) -> merger input + 1


class ConvolverNode(AudioNode):
    _accelerated = True
    def __init__(self, context):
        super().__init__(context)
        self.Normalize = True          # Nodes/ConvolverNode.cs:87
        self.EnableTrueStereo = True   # :95
        self._buffer = None
        # the impulse responses the node ran with over time: ConvolverNode.Buffer set again between two Render calls builds new
        # convolvers (cleared delay lines) from the next unprocessed quantum on (Nodes/ConvolverNode.cs:51-77)
        self._epochs = []              # dicts: q0, buffer (or None), args (normalize, true stereo), irs {member: handle}

    @property
    def Buffer(self):
        return self._buffer

    @Buffer.setter
    def Buffer(self, value):  # :25-79: the convolvers are built here, with the Normalize value of this moment
        if value is self._buffer:
            return
        ctx = self.Context
        if value is not None and value.SampleRate != ctx.SampleRate:  # Nodes/ConvolverNode.cs:48-49 (the library checks again)
            raise InvalidOperationException(
                "Impulse response buffer sample rate must match the audio context sample rate. "
                f"Impulse response buffer sample rate: {value.SampleRate}, Audio context sample rate: {ctx.SampleRate}.")
        # the convolvers are built with the Normalize / EnableTrueStereo values of THIS moment (:87-95)
        ep = dict(q0=ctx._q_now() if self._first_live_q is not None else 0, buffer=value,
                  args=(int(self.Normalize), int(self.EnableTrueStereo)), irs={})
        if self._epochs and self._epochs[-1]["q0"] >= ep["q0"]:
            self._epochs[-1] = ep  # set again before anything further was rendered
        else:
            self._epochs.append(ep)
        self._buffer = value
        # prepare now (the reference does the work in the setter); a multi-GPU context prepares per member at Render, on the device
        # that renders the voice
        if value is not None and not ctx._record_only and ctx._root()._group is None:
            self._epoch_ir(ep, 0)

    @property
    def _ir(self):
        """prepared impulse response of the current epoch in member 0 (None without a Buffer)"""
        if not self._epochs or self._epochs[-1]["buffer"] is None or self.Context._record_only:
            return None
        return self._epoch_ir(self._epochs[-1], 0)


def _convolver_epoch_ir(self, ep, member):
    """The prepared impulse response of epoch `ep` of this ConvolverNode in member `member` of its context (prepared on first use)."""
    h = ep["irs"].get(member)
    if h is None:
        ctx = self.Context
        out = C.c_void_p()
        check(N.lib().gac_ir_prepare(ctx._root()._member_handle(member), ep["buffer"]._handle(ctx, member), ep["args"][0], ep["args"][1], C.byref(out)))
        ctx._root()._owned_irs.append(out.value)
        h = ep["irs"][member] = out.value
    return h


ConvolverNode._epoch_ir = _convolver_epoch_ir


class CudaConvolverNode:
    """The literal plugin seam: a ConvolverNode that runs INSIDE a reference-style block loop, one native call
    per render quantum, the way GraphAudio.SteamAudio's nodes call their native effect from `Process()`
    (GraphAudio.SteamAudio/Nodes/SteamAudioNodeBase.cs:50-135).  Same members as ConvolverNode (Buffer / Normalize /
    EnableTrueStereo, Nodes/ConvolverNode.cs:25-95); `Process(input)` is the body of the C# node's `Process()` override:
    `input` is what `Inputs[0].Buffer` holds (a [channels][128] block already mixed to `InputChannelCount`), the return value
    is the block handed to `SetOutputBuffer`.  `ProcessFrames` feeds several quanta in one call.  State (delay line, overlap,
    IR spectra) lives on the device.  Latency-bound by construction — the batched renderer is `OfflineAudioContext.Render`."""

    def __init__(self, context: "OfflineAudioContext"):
        self.Context = context
        self.Normalize = True
        self.EnableTrueStereo = True
        self._buffer = None
        self._ir = None
        self._h = None
        self.InputChannelCount = 0   # what the C# node sets on Inputs[0] (SetChannelCount + Explicit, ConvolverNode.cs:62-76)
        self.OutputChannelCount = 0  # _effectiveOutputChannels (:60)

    @property
    def Buffer(self):
        return self._buffer

    @Buffer.setter
    def Buffer(self, value):
        if value is self._buffer:
            return
        L = N.lib()
        if self._h is not None:
            self.Context._root()._owned_convolvers.remove(self._h)
            L.gac_convolver_destroy(self._h)
            self._h = None
        self._buffer, self._ir = None, None
        self.InputChannelCount = self.OutputChannelCount = 0
        if value is None:
            return
        ctx = self.Context
        if value.SampleRate != ctx.SampleRate:  # Nodes/ConvolverNode.cs:48-49
            raise InvalidOperationException(
                "Impulse response buffer sample rate must match the audio context sample rate. "
                f"Impulse response buffer sample rate: {value.SampleRate}, Audio context sample rate: {ctx.SampleRate}.")
        ir = C.c_void_p()
        check(L.gac_ir_prepare(ctx._h, value._handle(ctx), int(self.Normalize), int(self.EnableTrueStereo), C.byref(ir)))
        ctx._root()._owned_irs.append(ir.value)
        h = C.c_void_p()
        check(L.gac_convolver_create(ctx._h, ir.value, C.byref(h)))
        ctx._root()._owned_convolvers.append(h.value)
        ni, no = C.c_int(), C.c_int()
        check(L.gac_convolver_channels(h.value, C.byref(ni), C.byref(no)))
        self._buffer, self._ir, self._h = value, ir.value, h.value
        self.InputChannelCount, self.OutputChannelCount = ni.value, no.value

    def Reset(self):
        if self._h is not None:
            check(N.lib().gac_convolver_reset(self._h))

    def ProcessFrames(self, block):
        """`block`: float32 [InputChannelCount][n], n a multiple of the partition; returns [OutputChannelCount][n]."""
        x = _f32(block)
        if x.ndim != 2:
            raise ArgumentException("input must be [channels][frames]")
        if self._h is None:  # no Buffer: the node outputs silence with the input's channel count (ConvolverNode.cs:107-119)
            return np.zeros_like(x)
        y = np.empty((self.OutputChannelCount, x.shape[1]), np.float32)
        ip = (N.fp * x.shape[0])(*[_fptr(x[c]) for c in range(x.shape[0])])
        op = (N.fp * y.shape[0])(*[_fptr(y[c]) for c in range(y.shape[0])])
        if x.shape[1] == 128:
            check(N.lib().gac_convolver_process_block(self._h, ip, x.shape[0], op, y.shape[0]))
        else:
            check(N.lib().gac_convolver_process(self._h, ip, x.shape[0], op, y.shape[0], x.shape[1]))
        return y

    def Process(self, input_block):
        return self.ProcessFrames(input_block)


class OfflineAudioContext:
    """OfflineAudioContext.cs — `Render` is the one call that crosses into libgraphaudio_cuda.so."""

    def __init__(self, sampleRate=48000, partition=128, device_id=-1, mac_variant=0, tile_blocks=32, async_upload=False,
                 uniform_segments=False, fanin_fusion=True, device_ids=None, _record_only=False):
        """partition / device_id / mac_variant / tile_blocks map onto gac_context_desc.  `_record_only=True` builds a context
        without a device handle: nodes, automation and topology can be recorded and inspected (`_topology()`), Render raises.
        It exists for the CPU unit tests of the host-side logic; it is not a fallback."""
        self._h = None
        self._serial = next(_context_serial)
        self._buffer_objects: List[PlayableAudioBuffer] = []  # buffers holding a device handle of this context (purged by Dispose)
        if sampleRate <= 0:
            raise ArgumentOutOfRangeException("sampleRate")  # AudioContextBase.cs:37-38
        self.SampleRate = int(sampleRate)
        self._nodes: List[AudioNode] = []
        self._owned_buffers: List[int] = []
        self._owned_irs: List[int] = []
        self._owned_convolvers: List[int] = []
        self._record_only = bool(_record_only)
        self._group = None       # gac_group handle of a context that spans several GPUs (device_ids=[...])
        self._members = None     # its member gac_context handles
        if not self._record_only:
            desc = N.gac_context_desc()
            desc.sample_rate, desc.quantum, desc.partition, desc.device_id, desc.mac_variant = self.SampleRate, 128, partition, device_id, mac_variant
            desc.tile_blocks = tile_blocks
            # async_upload: page-locked source arrays are uploaded asynchronously (GAC_FLAG_ASYNC_UPLOAD); the arrays handed
            # to PlayableAudioBuffer must then stay untouched until Render returns
            # fanin_fusion=False keeps one inverse transform and one fan-in input per ConvolverNode (GAC_FLAG_NO_FANIN_FUSION)
            desc.flags = ((N.GAC_FLAG_ASYNC_UPLOAD if async_upload else 0) | (N.GAC_FLAG_UNIFORM_SEGMENTS if uniform_segments else 0) |
                          (0 if fanin_fusion else N.GAC_FLAG_NO_FANIN_FUSION))
            out = C.c_void_p()
            if device_ids is not None:
                # ONE OfflineAudioContext over several GPUs of this process: voices -> bus -> destination graphs
                # are sharded by voice over the members at Render, anything else renders on member 0
                ids = (C.c_int * len(device_ids))(*[int(d) for d in device_ids])
                check(N.lib().gac_group_create(C.byref(desc), ids, len(device_ids), C.byref(out)))
                self._group = out.value
                self._members = []
                for i in range(len(device_ids)):
                    m = C.c_void_p()
                    check(N.lib().gac_group_context(self._group, i, C.byref(m)))
                    self._members.append(m.value)
                self._h = self._members[0]
            else:
                check(N.lib().gac_context_create(C.byref(desc), C.byref(out)))
                self._h = out.value
        self.Destination = AudioDestinationNode(self)
        self._frames_rendered = 0
        self.last_stats = None

    def _member_handle(self, member=0):
        return self._members[member] if self._members else self._h

    # ---- graph flattening: destination <- (bus chain <- fan-in)? <- voice chain <- source
    def _q_now(self):
        """First quantum the next Render call will process (the reference renders whole 128-frame blocks, OfflineAudioContext.cs:55-100)."""
        return -(-self._frames_rendered // 128)

    def _block_time(self, q):
        """Start time of quantum q, accumulated like AudioContextBase.cs:78-79 (_currentTime += 128.0 / SampleRate)."""
        t, inc = 0.0, 128.0 / self.SampleRate
        for _ in range(int(q)):
            t = t + inc
        return t

    def _op_descs(self, nodes, keep, member=0):
        """gac_op_desc entries of a flattened chain (a ConvolverNode that was given several impulse responses over time yields one
        op per epoch)"""
        out = []
        for n in nodes:
            if isinstance(n, ConvolverNode):
                eps = n._epochs or [dict(q0=0, buffer=None, args=(1, 1), irs={})]
                for k, ep in enumerate(eps):
                    op = N.gac_op_desc()
                    op.kind, op.filter_type, op.aux = N.GAC_OP_CONVOLVER, (1 if k else 0), float(ep["q0"] if k else 0)
                    op.ir = n._epoch_ir(ep, member) if ep["buffer"] is not None else None
                    out.append(op)
            else:
                out.append(self._op_desc(n, keep, member))
        return out

    def _op_desc(self, node, keep, member=0):
        op = N.gac_op_desc()
        q = self._q_now()
        if isinstance(node, BiQuadFilterNode):
            op.kind, op.filter_type = N.GAC_OP_BIQUAD, int(node.Type)
            op.p0, op.p1, op.p2 = node.Frequency._desc(keep, q), node.Q._desc(keep, q), node.Gain._desc(keep, q)
        elif isinstance(node, GainNode):
            op.kind = N.GAC_OP_GAIN
            op.p0 = node.Gain._desc(keep, q)
        elif isinstance(node, _Gate):
            op.kind, op.filter_type, op.aux = N.GAC_OP_GATE, node.kind, float(node.q)
        elif isinstance(node, DelayNode):
            op.kind, op.aux = N.GAC_OP_DELAY, node.MaxDelayTime
            op.p0 = node.DelayTime._desc(keep, q)
        elif isinstance(node, StereoPannerNode):
            op.kind, op.aux = N.GAC_OP_PANNER, float(node._first_live_q or 0)
            op.p0 = node.Pan._desc(keep, q)
        elif isinstance(node, _SplitterOutput):
            op.kind, op.aux = N.GAC_OP_CHANNEL, float(node.Index)
        else:
            raise NotSupportedException(f"{type(node).__name__} is outside the accelerated path")
        return op

    def _topology_full(self):
        """Pure host logic (no device needed).  Cuts the node graph into the three things the C ABI knows:
          * voices  — chains of single-input / single-output nodes fed by an AudioBufferSourceNode: (source, [ops], bus, -1)
          * buses   — a node with several inputs (AudioNodeInput fan-in, summed in connection order) plus the chain behind it
          * chains fed by a bus output — (None, [ops], bus, input_bus): a node whose output feeds several nodes ends a bus
            (one is inserted if the signal was not a bus yet), and every branch reads that bus
        which together express any acyclic graph of the supported nodes, e.g. GraphAudio.Kit's ReverbEffect
        (input -> dry gain -> output ; input -> convolver -> wet gain -> output, Effects/ReverbEffect.cs:63-91) and AudioBus
        hierarchies (AudioBus.cs:76-114).  Returns (voices, bus_ops, dest_inputs, bus_targets, bus_inputs)."""
        dest = self.Destination
        # nodes that reach the destination
        live, stack = set(), [dest]
        while stack:
            n = stack.pop()
            if id(n) in live:
                continue
            live.add(id(n))
            stack.extend(n._in)
            for prm in n._params():  # a connected modulator keeps its branch alive (AudioParam.ComputeValues pulls it, AudioParam.cs:97-101)
                if prm._input_node is not None and prm._input_node._in:
                    stack.append(prm._input_node)
            # a node type the device path does not accelerate must never render as silence: refuse the whole graph
            if not n._accelerated:
                raise NotSupportedException(f"{type(n).__name__} is outside the accelerated path (SURVEY.md §8f-3; the CPU oracle has it)")
            if isinstance(n, ChannelMergerNode) and any(sl > 2 for sl in n._slots.values()):
                raise NotSupportedException("ChannelMergerNode inputs beyond the second are outside the accelerated path (signals carry two channels)")

        q_now = self._q_now()
        for n in self._nodes:  # the first quantum in which a node is pulled: it is pulled while it reaches the destination
            if id(n) in live and n._first_live_q is None:
                n._first_live_q = q_now

        def outs(n):
            return [d for d in n._out if id(d) in live]

        def is_src(n):
            return n._is_source

        def ops_of(chain):  # nodes that only shape the flattening (fan-in points without a node behind them) carry no op
            return [n for n in chain if not isinstance(n, (_ParamInputNode, ChannelSplitterNode, ChannelMergerNode))]

        def fan_in(n):  # starts a bus
            return n is not dest and not n._is_source and (len(n._in) != 1 or n._force_bus)

        voices, bus_ops, bus_targets, bus_inputs, dest_inputs = [], [], [], [], []
        bus_of_head = {}   # id(fan-in node or materialised tail) -> bus index
        edge_code = {}     # (id(upstream), id(downstream)) -> input code at the downstream fan-in (>= 0 bus, < 0 ~voice)

        def chain_from(first):
            """[first, ...] downstream while the link is single-output -> single-input and does not enter a bus / the destination"""
            chain, n = [], first
            while True:
                chain.append(n)
                o = outs(n)
                if len(o) != 1 or o[0] is dest or fan_in(o[0]):
                    return chain
                chain.extend(n._gates(o[0]))  # a connection made / removed after rendering began (GAC_OP_GATE)
                n = o[0]

        def last_node(chain):
            return [x for x in chain if not isinstance(x, _Gate)][-1]

        def new_bus(ops):
            bus_ops.append(ops)
            bus_targets.append(-1)  # -1: no direct target (yet); 0: destination; k > 0: bus k - 1
            bus_inputs.append([])
            return len(bus_ops) - 1

        def emit(tail, code_is_bus, code):
            """records where the signal at `tail` (a bus, or the voice `code`) goes; fan-out materialises a bus"""
            o = outs(tail)
            if len(o) == 1:
                edge_code[(id(tail), id(o[0]))] = code if code_is_bus else ~code
                if code_is_bus:
                    bus_ops[code] = bus_ops[code] + tail._gates(o[0])
                    pending_bus_target.append((code, o[0]))
                else:
                    voices[code][1] = voices[code][1] + tail._gates(o[0])
                    pending_voice_target.append((code, o[0]))
                return
            # fan-out (or a dead end): every branch reads a bus
            if not code_is_bus:
                b = new_bus([])
                bus_inputs[b].append(~code)
                voices[code][2] = b
                code = b
            for d in o:
                branch(code, tail, d)

        def branch(bus, tail, d):
            """the edge tail -> d of a signal that lives in bus `bus`"""
            if d is dest or fan_in(d):
                voices.append([None, tail._gates(d), -2, bus])  # pass-through chain: the bus output as an input of d
                v = len(voices) - 1
                edge_code[(id(tail), id(d))] = ~v
                pending_voice_target.append((v, d))
            else:
                ch = chain_from(d)
                first_real = ch[0]
                voices.append([None, (tail._gates(d) if not isinstance(first_real, _SplitterOutput) else []) + ch, -2, bus])
                emit(last_node(ch), False, len(voices) - 1)

        pending_voice_target, pending_bus_target = [], []
        # 1. source-fed voices: one per live out-edge of every source
        for n in self._nodes:
            if not is_src(n) or id(n) not in live:
                continue
            for d in outs(n):
                if d is dest or fan_in(d):
                    voices.append([n, n._gates(d), -2, -1])
                    v = len(voices) - 1
                    edge_code[(id(n), id(d))] = ~v
                    pending_voice_target.append((v, d))
                else:
                    ch = chain_from(d)
                    voices.append([n, n._gates(d) + ch, -2, -1])
                    emit(last_node(ch), False, len(voices) - 1)
        # 2. buses: every live fan-in node, with the chain behind it
        for n in self._nodes:
            if id(n) not in live or not fan_in(n):
                continue
            if len(n._in) == 0 and not n._force_bus:
                continue  # nothing connected: contributes silence (its consumers see a missing input)
            ch = chain_from(n)
            b = new_bus(ops_of(ch))
            bus_of_head[id(n)] = b
            if isinstance(n, _ParamInputNode):
                n._bus_index = b
            emit(last_node(ch), True, b)
        # 3. resolve targets now that every fan-in has its bus index
        for v, d in pending_voice_target:
            voices[v][2] = -1 if d is dest else bus_of_head[id(d)]
        for b, d in pending_bus_target:
            bus_targets[b] = 0 if d is dest else bus_of_head[id(d)] + 1
        # 4. connection order at every fan-in and at the destination
        for n in self._nodes:
            if id(n) in bus_of_head:
                b = bus_of_head[id(n)]
                bus_inputs[b] = [edge_code[(id(u), id(n))] for u in n._in if (id(u), id(n)) in edge_code] + bus_inputs[b]
        dest_inputs = [edge_code[(id(u), id(dest))] for u in dest._in if (id(u), id(dest)) in edge_code]
        # bus flags (the input of an AudioParam has one channel) and ChannelMergerNode input slots, for _flatten
        self._bus_flags = [0] * len(bus_ops)
        self._bus_slots = [None] * len(bus_ops)
        for n in self._nodes:
            if id(n) in bus_of_head:
                b = bus_of_head[id(n)]
                if isinstance(n, _ParamInputNode):
                    self._bus_flags[b] = N.GAC_BUS_MONO_INPUT
                if isinstance(n, ChannelMergerNode):
                    self._bus_slots[b] = [n._slots.get(id(u), 0) for u in n._in if (id(u), id(n)) in edge_code]
        # materialised fan-out buses keep the single input recorded in emit()
        emit = branch = None  # (the two closures refer to each other: cut the cycle, or every flattening leaves its lists to the collector)
        return [tuple(v) for v in voices], bus_ops, dest_inputs, bus_targets, bus_inputs

    def _topology(self):
        """([(source, [ops], bus, input_bus)], [[bus ops]], [dest input codes]) — see _topology_full."""
        voices, buses, dest_inputs, _, _ = self._topology_full()
        return voices, buses, dest_inputs

    def _flatten(self, member=0, voice_range=None):
        """gac_graph_desc of the recorded graph for member `member`; voice_range = (lo, hi) keeps only those source-fed voices (a
        shard of a flat graph: the buses are described in full on every member)."""
        if getattr(self, "_unsupported_edit", None):
            raise NotSupportedException(
                f"{self._unsupported_edit}: successive Render calls re-render the timeline on the device, which is exact for parameter "
                "edits, for sources started or stopped in between and for new branches, not for re-wiring what was already rendered")
        keep = []
        voices, buses, dest_inputs, bus_targets, bus_inputs = self._topology_full()
        if voice_range is not None:
            lo, hi = voice_range
            voices = voices[lo:hi]
            bus_inputs = [[] for _ in buses]  # (default order: the shard's voices in index order)
        vdesc = (N.gac_voice_desc * max(1, len(voices)))()
        for i, (src, ops, bus, input_bus) in enumerate(voices):
            v = vdesc[i]
            if isinstance(src, _ScheduledSourceNode):
                when = src._when if src._started else math.nan
                q_first = max(-(-src._start_frames // 128), src._first_live_q or 0)
                if src._started and q_first > 0:
                    when = max(when, self._block_time(q_first))
                osc = isinstance(src, OscillatorNode)
                v.source = None
                v.source_kind = N.GAC_SOURCE_OSCILLATOR if osc else N.GAC_SOURCE_CONSTANT
                v.source_param = (src.Frequency if osc else src.Offset)._desc(keep, self._q_now())
                v.oscillator_type = int(src.Type) if osc else 0
                v.start_when, v.start_offset, v.start_duration, v.stop_when = when, 0.0, src._duration, src._stop
                v.playback_rate = 1.0
            elif src is not None:
                if not src._started or src.Buffer is None:
                    when = math.nan
                else:
                    when = src._when
                    # started, or first connected, between Render calls: the node is first processed in a later quantum and begins there
                    q_first = max(-(-src._start_frames // 128), src._first_live_q or 0)
                    if q_first > 0:
                        when = max(when, self._block_time(q_first))
                v.loop, v.loop_start, v.loop_end = int(bool(src.Loop)), src.LoopStart, src.LoopEnd
                v.source = src.Buffer._handle(self, member) if src.Buffer is not None else None
                v.start_when, v.start_offset, v.start_duration, v.stop_when = when, src._offset, src._duration, src._stop
                v.playback_rate = src.PlaybackRate.Value
                # PlaybackRate as a full parameter: read by the library when it carries events (automation, or epochs of a Value that
                # was edited between Render calls); k-rate, evaluated per quantum on the host
                v.source_param = src.PlaybackRate._desc(keep, self._q_now())
            else:
                v.source = None
                v.playback_rate = 1.0
            descs = self._op_descs(ops, keep, member)
            arr = (N.gac_op_desc * max(1, len(descs)))(*descs)
            keep.append(arr)
            v.n_ops, v.ops, v.bus, v.input = len(descs), arr, bus, input_bus + 1
        bdesc = (N.gac_bus_desc * max(1, len(buses)))()
        for i, ops in enumerate(buses):
            descs = self._op_descs(ops, keep, member)
            arr = (N.gac_op_desc * max(1, len(descs)))(*descs)
            inp = (C.c_int32 * max(1, len(bus_inputs[i])))(*bus_inputs[i])
            keep += [arr, inp]
            bdesc[i].n_ops, bdesc[i].ops = len(descs), arr
            bdesc[i].target, bdesc[i].n_inputs, bdesc[i].inputs = bus_targets[i], len(bus_inputs[i]), inp
            bdesc[i].flags = self._bus_flags[i]
            if self._bus_slots[i] is not None and voice_range is None:
                sl = (C.c_int32 * max(1, len(self._bus_slots[i])))(*self._bus_slots[i])
                keep.append(sl)
                bdesc[i].input_slots = sl
        darr = (C.c_int32 * max(1, len(dest_inputs)))(*dest_inputs)
        g = N.gac_graph_desc()
        g.n_voices, g.voices, g.n_buses, g.buses = len(voices), vdesc, len(buses), bdesc
        g.n_dest_inputs, g.dest_inputs = len(dest_inputs), darr
        keep += [vdesc, bdesc, darr]
        return g, keep

    def _graph(self, member=0, voice_range=None):
        g, keep = self._flatten(member, voice_range)
        out = C.c_void_p()
        check(N.lib().gac_graph_create(self._member_handle(member), C.byref(g), C.byref(out)))
        return out.value

    def _shardable(self):
        """True when the recorded graph is voices -> ONE bus -> destination with every voice fed by a source: the shape whose
        voices a multi-GPU context spreads over its members (the bus is what the single ncclReduce sums)."""
        voices, buses, dest_inputs, bus_targets, _ = self._topology_full()
        return (len(buses) == 1 and bus_targets[0] == 0 and list(dest_inputs) == [0] and len(voices) > 0
                and all(src is not None and bus == 0 and input_bus < 0 for src, _, bus, input_bus in voices))

    def _render_group(self, rows, frameCount, startIndex):
        """Render of a multi-GPU context: contiguous voice ranges per member, one gac_group_render call."""
        from . import sharding
        L = N.lib()
        M = len(self._members)
        nv = len(self._topology_full()[0])
        graphs = []
        try:
            for m in range(M):
                graphs.append(self._graph(m, sharding.shard_range(nv, m, M)))
            garr = (C.c_void_p * M)(*graphs)
            ptrs = (N.fp * len(rows))(*[_fptr(r) for r in rows])
            check(L.gac_group_render(self._group, garr, self._frames_rendered, int(frameCount), ptrs, len(rows), int(startIndex)))
            self._frames_rendered += int(frameCount)
            self.last_stats_members = []
            for m in range(M):
                st = N.gac_stats()
                check(L.gac_get_stats(self._members[m], C.byref(st)))
                self.last_stats_members.append(st.as_dict())
            self.last_stats = self.last_stats_members[0]
        finally:
            for g in graphs:
                L.gac_graph_destroy(g)

    def Render(self, output_or_count, frameCount=None, startIndex=0):
        """Render(float[][] output, int frameCount, int startIndex = 0)  (OfflineAudioContext.cs:30)
        or  float[][] Render(int frameCount)  (:108).  Successive calls continue the timeline (:55-100)."""
        if self._record_only:
            raise InvalidOperationException("record-only context: there is no CPU render path (build the library and use a B200)")
        if self._h is None:
            raise ObjectDisposedException("OfflineAudioContext")
        if frameCount is None:
            n = int(output_or_count)
            if n <= 0:
                raise ArgumentOutOfRangeException("Frame count must be positive.")
            out = np.zeros((2, n), np.float32)
            self.Render(out, n, 0)
            return out
        output = output_or_count
        if len(output) == 0:
            raise ArgumentException("Output buffer must have at least one channel.")
        if frameCount <= 0:
            raise ArgumentOutOfRangeException("Frame count must be positive.")
        if startIndex < 0:
            raise ArgumentOutOfRangeException("Start index must be non-negative.")
        rows = [output[c] for c in range(len(output))]
        for c, r in enumerate(rows):
            if r is None:
                raise ArgumentException(f"Channel {c} buffer is null.")
            if r.dtype != np.float32 or not r.flags.c_contiguous:
                raise ArgumentException("channel buffers must be contiguous float32")
            if r.shape[0] < startIndex + frameCount:
                raise ArgumentException(f"Channel {c} buffer is too small. Required: {startIndex + frameCount}, Available: {r.shape[0]}")
        if self._group is not None and self._shardable():
            return self._render_group(rows, frameCount, startIndex)
        graph = self._graph()
        try:
            ptrs = (N.fp * len(rows))(*[_fptr(r) for r in rows])
            check(N.lib().gac_render(self._h, graph, self._frames_rendered, int(frameCount), ptrs, len(rows), int(startIndex)))
            self._frames_rendered += int(frameCount)
            st = N.gac_stats()
            check(N.lib().gac_get_stats(self._h, C.byref(st)))
            self.last_stats = st.as_dict()
        finally:
            N.lib().gac_graph_destroy(graph)

    def RenderInterleaved(self, frameCount, channels=2, output=None, startIndex=0):
        """The render as interleaved frames [frameCount * channels] — the array a caller fills by driving
        `ProcessBlockInterleaved(float[] interleavedBuffer, int channels)` block after block (AudioContextBase.cs:88-161):
        destination channels first, zeros in the others.  Continues the timeline like Render."""
        if self._record_only:
            raise InvalidOperationException("record-only context: there is no CPU render path (build the library and use a B200)")
        if self._h is None:
            raise ObjectDisposedException("OfflineAudioContext")
        if channels < 1 or channels > 32:
            raise ArgumentOutOfRangeException("channels")  # :93
        if frameCount <= 0:
            raise ArgumentOutOfRangeException("Frame count must be positive.")
        if output is None:
            output = np.zeros((startIndex + frameCount) * channels, np.float32)
        if output.dtype != np.float32 or not output.flags.c_contiguous or output.size < (startIndex + frameCount) * channels:
            raise ArgumentException("Buffer too small for interleaved output.")  # :94-95
        graph = self._graph()
        try:
            check(N.lib().gac_render_interleaved(self._h, graph, self._frames_rendered, int(frameCount), _fptr(output), int(channels),
                                                 int(startIndex)))
            self._frames_rendered += int(frameCount)
        finally:
            N.lib().gac_graph_destroy(graph)
        return output

    # ---- multi-GPU: voices sharded over processes, one NCCL reduce of the bus (gac_render_sharded)
    def MarkBus(self, node):
        """Declares `node`'s input as the fan-in that is reduced across ranks.  Needed because a shard may hold one voice
        (or none): the flattening would otherwise fold `voice -> node -> destination` into a single direct voice and
        apply the bus ops before the reduce instead of after it."""
        node._force_bus = True
        return node

    def CommInit(self, unique_id: bytes, rank: int, world: int):
        check(N.lib().gac_comm_init(self._h, unique_id, int(rank), int(world)))

    @staticmethod
    def CommUniqueId() -> bytes:
        buf = (C.c_char * 128)()
        check(N.lib().gac_comm_unique_id(buf))
        return bytes(buf.raw)

    def RenderSharded(self, output, frameCount, root=0):
        """Collective: every rank renders its shard up to the bus, the buses are summed onto `root` by one ncclReduce, the
        root applies the bus ops and fills `output` (float32 [2][>= frameCount]); other ranks leave it untouched."""
        if frameCount <= 0:
            raise ArgumentOutOfRangeException("Frame count must be positive.")
        graph = self._graph()
        try:
            rows = [output[c] for c in range(len(output))]
            ptrs = (N.fp * len(rows))(*[_fptr(r) for r in rows])
            check(N.lib().gac_render_sharded(self._h, graph, int(frameCount), int(root), ptrs, len(rows)))
            st = N.gac_stats()
            check(N.lib().gac_get_stats(self._h, C.byref(st)))
            self.last_stats = st.as_dict()
        finally:
            N.lib().gac_graph_destroy(graph)

    # ---- batches of independent renders share one device context
    def Fork(self):
        """A new, empty OfflineAudioContext (own Destination, own timeline) that records against this context's device
        handle.  Used with RenderBatch; disposing the parent disposes the shared handle."""
        child = OfflineAudioContext.__new__(OfflineAudioContext)
        child._h = self._h
        child._group, child._members = None, None  # (forks record against member 0)
        child._serial = self._serial
        child._buffer_objects = self._buffer_objects
        child.SampleRate = self.SampleRate
        child._nodes = []
        child._owned_buffers = self._owned_buffers
        child._owned_irs = self._owned_irs
        child._record_only = self._record_only
        child._parent = self
        child.Destination = AudioDestinationNode(child)
        child._frames_rendered = 0
        child.last_stats = None
        return child

    def _root(self):
        c = self
        while getattr(c, "_parent", None) is not None:
            c = c._parent
        return c

    def Dispose(self):
        if getattr(self, "_parent", None) is not None:
            self._h = None  # forks do not own the handle
            return
        if self._h is not None:
            L = N.lib()
            for h in self._owned_convolvers:
                L.gac_convolver_destroy(h)
            for h in self._owned_irs:
                L.gac_ir_destroy(h)
            for h in self._owned_buffers:
                L.gac_buffer_destroy(h)
            for b in self._buffer_objects:
                for key in [k for k in b._handles if k[0] == self._serial]:
                    b._handles.pop(key, None)
            self._buffer_objects = []
            if self._group is not None:
                L.gac_group_destroy(self._group)  # destroys the member contexts
                self._group, self._members = None, None
            else:
                L.gac_context_destroy(self._h)
            self._h = None
        # The recorded graph is a web of mutual references (node <-> node, node -> context -> node): left alone it is cyclic garbage
        # that only the generational collector frees — one OfflineAudioContext per render (the reference's usual pattern) then pays a
        # full collection of ~40 000 objects every tenth 128-voice render (+10 ms).  Cut the web here; reference counting does the rest.
        nodes, self._nodes = self._nodes, []
        for n in nodes:
            n._in, n._out = [], []
            for prm in n._params():
                prm._input_node = None

    def __del__(self):
        try:
            self.Dispose()
        except Exception:
            pass


def RenderBatch(contexts, frameCount):
    """BASELINE config 4: many independent OfflineAudioContexts rendered in ONE native call (gac_render_batch).

    `contexts` are graphs recorded against a shared device context: build each with `OfflineAudioContext.Fork()` of a
    common parent (they share the parent's gac_context, hence its stream and scratch pool).  Returns float32
    [len(contexts), 2, frameCount].  Equivalent to calling Render(frameCount) on each context in turn, as the reference
    would (one context per render, OfflineAudioContext.cs:30), but batched across renders on the device."""
    if frameCount <= 0:
        raise ArgumentOutOfRangeException("Frame count must be positive.")
    if not contexts:
        raise ArgumentException("no contexts")
    root = contexts[0]._root()
    if any(c._root() is not root for c in contexts):
        raise ArgumentException("all contexts of a batch must be forks of the same parent context")
    L = N.lib()
    graphs = [c._graph() for c in contexts]
    try:
        out = np.zeros((len(contexts), 2, frameCount), np.float32)
        rows = (N.fp * (2 * len(contexts)))(*[_fptr(out[g, c]) for g in range(len(contexts)) for c in range(2)])
        garr = (C.c_void_p * len(graphs))(*graphs)
        check(L.gac_render_batch(root._h, garr, len(graphs), int(frameCount), rows, 2))
        st = N.gac_stats()
        check(L.gac_get_stats(root._h, C.byref(st)))
        root.last_stats = st.as_dict()
        return out
    finally:
        for g in graphs:
            L.gac_graph_destroy(g)
