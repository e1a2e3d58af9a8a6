"""ga_twin.py — a SECOND, independent CPU restatement of the reference's offline render path, in numpy.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): imported by tests/ alone.

Why it exists: the reference is C#/.NET 9 and cannot run in this image, so the C++ oracle (ga_oracle.cpp) is "parity
unpinned" — and oracle and kernels were written from one reading of the source.  This file is a second reading: it was
written from the C# files again (not from ga_oracle.cpp), in another language, with another FFT (numpy's pocketfft instead of
the oracle's radix-2) and the SAME libm entry points the .NET runtime forwards MathF.* / Math.* to on Linux
(sinf, cosf, powf, sqrtf, pow, exp — called through ctypes / the `math` module).  tests/test_twin_vs_oracle.py renders the golden
graphs with both and demands bit-equality wherever no FFT is involved, and <= 1 ulp-level agreement (2e-7) behind the
convolver, where two different double-precision FFTs may round a float32 spectrum value differently.

It follows the reference's own structure (pull graph, one 128-frame block at a time), slow and literal:
  AudioContextBase.ProcessBlock        AudioContextBase.cs:52-81
  OfflineAudioContext.Render           OfflineAudioContext.cs:30-124
  AudioNode.ProcessInternal            Nodes/AudioNode.cs:152-183
  AudioNodeInput.Pull / MixBuffer      AudioNodeInput.cs:100-244
  AudioParam                           AudioParam.cs:93-352
  AudioBufferSourceNode                Nodes/AudioBufferSourceNode.cs:79-402 (incl. Loop / LoopStart / LoopEnd on both paths)
  CubicResampler                       CubicResampler.cs:19-97
  BiQuadFilterNode                     Nodes/BiQuadFilterNode.cs:87-258
  GainNode                             Nodes/GainNode.cs:29-61
  ConvolverNode                        Nodes/ConvolverNode.cs:25-164
  PartitionedConvolver                 PartitionedConvolver.cs:37-223
All paths relative to /root/reference/GraphAudio.Core/.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math

import numpy as np

F = np.float32
FRAMES = 128  # AudioBuffer.FramesPerBlock (AudioBuffer.cs:10)

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("sinf", "cosf", "sqrtf"):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float]
_libm.powf.restype = ctypes.c_float
_libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]


def sinf(x):
    return F(_libm.sinf(float(x)))


def cosf(x):
    return F(_libm.cosf(float(x)))


def sqrtf(x):
    return F(_libm.sqrtf(float(x)))


def powf(x, y):
    return F(_libm.powf(float(x), float(y)))


class ArgumentException(ValueError):
    pass


class ArgumentOutOfRangeException(ArgumentException):
    pass


class InvalidOperationException(RuntimeError):
    pass


class FilterType:  # Nodes/BiQuadFilterNode.cs:288-298
    Lowpass, Highpass, Bandpass, Notch, Allpass, Peaking, Lowshelf, Highshelf = range(8)


# ------------------------------------------------------------------------------------------------ AudioBuffer.cs
class Block:
    def __init__(self, channels):
        self.data = np.zeros((channels, FRAMES), F)
        self.silent = True  # BufferPool.Rent hands out cleared blocks (BufferPool.cs:66-85)

    @property
    def channels(self):
        return self.data.shape[0]

    def clear(self):  # :60-67
        self.data[:] = 0
        self.silent = True


class PlayableAudioBuffer:  # PlayableAudioBuffer.cs
    def __init__(self, channels, sample_rate):
        self.channels = [np.ascontiguousarray(c, dtype=F) for c in channels]
        self.SampleRate = int(sample_rate)
        self.NumberOfChannels = len(self.channels)
        self.Length = int(self.channels[0].shape[0])

    @staticmethod
    def FromChannelArrays(channelData, sampleRate):
        return PlayableAudioBuffer(channelData, sampleRate)

    @staticmethod
    def FromMonoArray(audioData, sampleRate):
        return PlayableAudioBuffer([audioData], sampleRate)

    @staticmethod
    def FromStereoArrays(l, r, sampleRate):
        return PlayableAudioBuffer([l, r], sampleRate)


# ------------------------------------------------------------------------------------------------ AudioNodeInput.cs
class Input:
    def __init__(self):
        self.connected = []          # outputs, in connection order (:60-67)
        self.channel_count = 2       # :19
        self.mode = "max"            # :21
        self.buffer = None

    def pull(self, block, time):     # :100-138
        if not self.connected:
            if self.buffer is None or self.buffer.channels != self.channel_count:
                self.buffer = Block(self.channel_count)
            self.buffer.clear()
            return
        n_out = self._output_channels()
        if self.buffer is None or self.buffer.channels != n_out:
            self.buffer = Block(n_out)
        self.buffer.clear()
        mixed = False
        for out in self.connected:
            out.owner.process_internal(block, time)
            src = out.buffer
            if src is not None and not src.silent:
                _mix(src, self.buffer)
                mixed = True
        if mixed:
            self.buffer.silent = False

    def _output_channels(self):      # :140-168
        if self.mode == "explicit":
            return self.channel_count
        if self.mode == "clamped-max":
            m = 0
            for o in self.connected:
                if o.buffer is not None:
                    m = max(m, o.buffer.channels)
            return min(self.channel_count if m == 0 else m, self.channel_count)
        m = self.channel_count
        for o in self.connected:
            if o.buffer is not None:
                m = max(m, o.buffer.channels)
        return m


def _mix(src, dst):                   # MixBuffer :182-244
    s, d = src.channels, dst.channels
    if s == d:
        dst.data += src.data
    elif s == 1 and d > 1:
        dst.data += src.data[0]
    elif s > 1 and d == 1:
        scale = F(1.0) / sqrtf(F(s))  # 1.0f / MathF.Sqrt(srcChannels)
        total = np.zeros(FRAMES, F)
        for ch in range(s):
            total = total + src.data[ch]
        dst.data[0] += total * scale
    else:
        m = min(s, d)
        dst.data[:m] += src.data[:m]


class Output:
    def __init__(self, owner):
        self.owner = owner
        self.buffer = None


# ------------------------------------------------------------------------------------------------ AudioParam.cs
class AudioParam:
    def __init__(self, owner, default, mn, mx, a_rate=True):
        self._owner = owner
        self._value = F(default)
        self.MinValue, self.MaxValue = F(mn), F(mx)
        self._events = []            # (type, value, target, time, time_constant); types as :368-374
        self._a_rate = a_rate
        self._input = Input()        # modulation input (:97-101)
        self.values = np.zeros(FRAMES, F)

    def _clamp(self, v):
        return F(min(max(F(v), self.MinValue), self.MaxValue))

    @property
    def Value(self):
        return float(self._value)

    @Value.setter
    def Value(self, v):              # :34-49
        self._value = self._clamp(v)
        self._events = []

    def _add(self, ev):              # AddEvent :333-352
        lo, hi = 0, len(self._events)
        while lo < hi:
            mid = (lo + hi) >> 1
            if ev[3] < self._events[mid][3]:
                hi = mid
            else:
                lo = mid + 1
        self._events.insert(lo, ev)

    def SetValueAtTime(self, value, startTime):
        self._add((0, self._clamp(value), F(0), float(startTime), 0.0))

    def LinearRampToValueAtTime(self, value, endTime):
        self._add((1, self._clamp(value), F(0), float(endTime), 0.0))

    def ExponentialRampToValueAtTime(self, value, endTime):
        v = self._clamp(value)
        if v <= 0:
            raise ArgumentException("Exponential ramp target must be > 0")
        self._add((2, v, F(0), float(endTime), 0.0))

    def SetTargetAtTime(self, target, startTime, timeConstant):
        self._add((3, F(0), self._clamp(target), float(startTime), float(timeConstant)))

    def CancelScheduledValues(self, cancelTime):  # :312-331
        keep = 0
        for e in self._events:
            if e[3] < cancelTime:
                keep += 1
            else:
                break
        self._events = self._events[:keep]

    # ---- evaluation
    def compute(self, block, time):  # ComputeValues :93-112
        has_mod = len(self._input.connected) > 0
        if has_mod:
            self._input.pull(block, time)
        mod = self._input.buffer if has_mod and self._input.buffer is not None and not self._input.buffer.silent else None
        if self._a_rate:             # ComputeARate :114-141
            dt = 1.0 / self._owner.Context.SampleRate
            for i in range(FRAMES):
                v = self._value_at(time + i * dt)
                if mod is not None:
                    v = F(min(max(F(v + mod.data[0][i]), self.MinValue), self.MaxValue))
                self.values[i] = v
        else:                        # ComputeKRate :143-166
            v = self._value_at(time)
            if mod is not None:
                v = F(min(max(F(v + mod.data[0][0]), self.MinValue), self.MaxValue))
            self.values[:] = v

    def _value_at(self, t):          # ComputeValueAtTime :169-217
        ev = self._events
        if not ev:
            return self._value
        boundary = self._value
        for i, e in enumerate(ev):
            if t < e[3]:
                if i == 0:
                    return boundary
                p = ev[i - 1]
                if e[0] == 1:
                    return _lin(p[1], p[3], e[1], e[3], t)
                if e[0] == 2:
                    return _exp(p[1], p[3], e[1], e[3], t)
                if p[0] == 3:
                    return _target(p, boundary, t)
                return p[1]
            if e[0] != 3:
                boundary = e[1]
        last = ev[-1]
        if last[0] == 3:
            return _target(last, boundary, t)
        return last[1]


def _lin(v0, t0, v1, t1, t):         # InterpolateLinear :220-225
    u = (t - t0) / (t1 - t0)
    u = min(max(u, 0.0), 1.0)
    return F(float(v0) + float(F(v1 - v0)) * u)


def _exp(v0, t0, v1, t1, t):         # InterpolateExponential :228-237
    if v0 <= 0 or v1 <= 0:
        return _lin(v0, t0, v1, t1, t)
    u = (t - t0) / (t1 - t0)
    u = min(max(u, 0.0), 1.0)
    return F(float(v0) * math.pow(float(F(v1 / v0)), u))


def _target(e, base, t):             # ComputeSetTargetFromBaseline :240-247
    elapsed = t - e[3]
    if elapsed <= 0:
        return F(base)
    tc = max(e[4], 0.001)
    return F(float(e[2]) + float(F(base - e[2])) * math.exp(-elapsed / tc))


# ------------------------------------------------------------------------------------------------ Nodes/AudioNode.cs
class AudioNode:
    def __init__(self, context, n_in=1, n_out=1):
        self.Context = context
        self.inputs = [Input() for _ in range(n_in)]
        self.outputs = [Output(self) for _ in range(n_out)]
        self.params = []
        self._last_block = -1
        self._processing = False

    def _param(self, default, mn, mx, a_rate=True):
        p = AudioParam(self, default, mn, mx, a_rate)
        self.params.append(p)
        return p

    def Connect(self, destination, outputIndex=0, inputIndex=0):  # :68-92
        if isinstance(destination, AudioParam):
            inp = destination._input
        else:
            inp = destination.inputs[inputIndex]
        out = self.outputs[outputIndex]
        if out not in inp.connected:
            inp.connected.append(out)
        return destination

    def process_internal(self, block, time):  # :152-183
        if self._last_block == block:
            return
        if self._processing:
            raise InvalidOperationException("Audio graph cycle detected")
        self._processing = True
        self._last_block = block
        try:
            for p in self.params:
                p.compute(block, time)
            for i in self.inputs:
                i.pull(block, time)
            self.process()
        finally:
            self._processing = False

    def process(self):
        raise NotImplementedError


class AudioDestinationNode(AudioNode):  # Nodes/AudioDestinationNode.cs
    def __init__(self, context):
        super().__init__(context, 1, 0)
        self.inputs[0].channel_count = 2  # :17
        self.out = None

    def process(self):                    # :42-64
        self.out = self.inputs[0].buffer


class GainNode(AudioNode):                # Nodes/GainNode.cs
    def __init__(self, context):
        super().__init__(context)
        fmax = float(np.finfo(F).max)
        self.Gain = self._param(1.0, -fmax, fmax)
        self._out = None

    def process(self):                    # :29-61
        inp = self.inputs[0].buffer
        if self._out is None or self._out.channels != inp.channels:
            self._out = Block(inp.channels)
        if inp.silent:
            self._out.clear()
            self.outputs[0].buffer = self._out
            return
        self._out.data[:] = inp.data * self.Gain.values  # float32 * float32, per channel
        self._out.silent = False
        self.outputs[0].buffer = self._out


class BiQuadFilterNode(AudioNode):        # Nodes/BiQuadFilterNode.cs
    def __init__(self, context):
        super().__init__(context)
        self._type = FilterType.Lowpass
        self.Frequency = self._param(1000.0, 1.0, context.SampleRate / 2.0)
        self.Q = self._param(1.0, 0.001, 1000.0)
        self.Gain = self._param(0.0, -60.0, 60.0, a_rate=False)
        self._last_f, self._last_q = F(1000.0), F(1.0)  # :13-14 (never written again)
        self._dirty = True                               # :17
        self._coef = (F(1), F(0), F(0), F(0), F(0))      # b0 b1 b2 a1 a2
        self._state = []
        self._out = None

    @property
    def Type(self):
        return self._type

    @Type.setter
    def Type(self, v):                    # :21-36
        if v != self._type:
            self._type = v
            self._dirty = True

    def process(self):                    # :87-147
        fv, qv = self.Frequency.values, self.Q.values
        gain_db = F(self.Gain.values[0])
        inp = self.inputs[0].buffer
        ch_n = inp.channels
        while len(self._state) < ch_n:
            self._state.append([F(0), F(0)])
        if self._out is None or self._out.channels != ch_n:
            self._out = Block(ch_n)
        if inp.silent:
            self._out.clear()
            self.outputs[0].buffer = self._out
            return
        b0, b1, b2, a1, a2 = self._coef
        used_f, used_q, used_g = self._last_f, self._last_q, gain_db
        half = F(F(self.Context.SampleRate) / F(2.0))
        for ch in range(ch_n):
            x_row, y_row = inp.data[ch], self._out.data[ch]
            w1, w2 = self._state[ch]
            for i in range(FRAMES):
                f = F(min(max(fv[i], F(1.0)), half))
                q = F(max(F(0.001), qv[i]))
                if self._dirty or abs(F(f - used_f)) > F(0.001) or abs(F(q - used_q)) > F(0.0001) or abs(F(gain_db - used_g)) > F(0.001):
                    self._update(f, q, gain_db)
                    used_f, used_q, used_g = f, q, gain_db
                    self._dirty = False
                    b0, b1, b2, a1, a2 = self._coef
                x = x_row[i]
                w = F(F(x - F(a1 * w1)) - F(a2 * w2))
                y = F(F(F(b0 * w) + F(b1 * w1)) + F(b2 * w2))
                w2 = w1
                w1 = w
                y_row[i] = y
            self._state[ch] = [w1, w2]
        self._out.silent = False
        self.outputs[0].buffer = self._out

    def _update(self, frequency, q, gain):  # UpdateCoefficients :149-258, every operation in float32, left to right
        fs = F(self.Context.SampleRate)
        w0 = F(F(F(F(2.0) * F(math.pi)) * frequency) / fs)
        c, s = cosf(w0), sinf(w0)
        alpha = F(s / F(F(2.0) * q))
        one, two = F(1.0), F(2.0)
        t = self._type
        if t == FilterType.Lowpass:
            b0 = F(F(one - c) / two); b1 = F(one - c); b2 = F(F(one - c) / two)
            a0 = F(one + alpha); a1 = F(F(-2.0) * c); a2 = F(one - alpha)
        elif t == FilterType.Highpass:
            b0 = F(F(one + c) / two); b1 = F(-F(one + c)); b2 = F(F(one + c) / two)
            a0 = F(one + alpha); a1 = F(F(-2.0) * c); a2 = F(one - alpha)
        elif t == FilterType.Bandpass:
            b0 = alpha; b1 = F(0); b2 = F(-alpha)
            a0 = F(one + alpha); a1 = F(F(-2.0) * c); a2 = F(one - alpha)
        elif t == FilterType.Notch:
            b0 = one; b1 = F(F(-2.0) * c); b2 = one
            a0 = F(one + alpha); a1 = F(F(-2.0) * c); a2 = F(one - alpha)
        elif t == FilterType.Allpass:
            b0 = F(one - alpha); b1 = F(F(-2.0) * c); b2 = F(one + alpha)
            a0 = F(one + alpha); a1 = F(F(-2.0) * c); a2 = F(one - alpha)
        elif t == FilterType.Peaking:
            A = powf(F(10.0), F(gain / F(40.0)))
            b0 = F(one + F(alpha * A)); b1 = F(F(-2.0) * c); b2 = F(one - F(alpha * A))
            a0 = F(one + F(alpha / A)); a1 = F(F(-2.0) * c); a2 = F(one - F(alpha / A))
        elif t in (FilterType.Lowshelf, FilterType.Highshelf):
            A = powf(F(10.0), F(gain / F(40.0)))
            beta = F(sqrtf(A) / q)
            ap1, am1 = F(A + one), F(A - one)
            bs = F(beta * s)
            if t == FilterType.Lowshelf:
                b0 = F(A * F(F(ap1 - F(am1 * c)) + bs))
                b1 = F(F(two * A) * F(am1 - F(ap1 * c)))
                b2 = F(A * F(F(ap1 - F(am1 * c)) - bs))
                a0 = F(F(ap1 + F(am1 * c)) + bs)
                a1 = F(F(-2.0) * F(am1 + F(ap1 * c)))
                a2 = F(F(ap1 + F(am1 * c)) - bs)
            else:
                b0 = F(A * F(F(ap1 + F(am1 * c)) + bs))
                b1 = F(F(F(-2.0) * A) * F(am1 + F(ap1 * c)))
                b2 = F(A * F(F(ap1 + F(am1 * c)) - bs))
                a0 = F(F(ap1 - F(am1 * c)) + bs)
                a1 = F(two * F(am1 - F(ap1 * c)))
                a2 = F(F(ap1 - F(am1 * c)) - bs)
        else:
            b0, b1, b2, a0, a1, a2 = one, F(0), F(0), one, F(0), F(0)
        self._coef = (F(b0 / a0), F(b1 / a0), F(b2 / a0), F(a1 / a0), F(a2 / a0))


# ------------------------------------------------------------------------------------------------ PartitionedConvolver.cs
class PartitionedConvolver:
    def __init__(self, ir, block_size=128, normalize=True):  # :37-63
        ir = np.ascontiguousarray(ir, dtype=F)
        self.B = block_size
        self.N = 2 * block_size
        self.C = self.N // 2 + 1
        self.P = int(math.ceil(ir.shape[0] / block_size))
        self.ir_re = np.zeros((self.P, self.C), F)
        self.ir_im = np.zeros((self.P, self.C), F)
        self.dl_re = np.zeros((self.P, self.C), F)
        self.dl_im = np.zeros((self.P, self.C), F)
        self.overlap = np.zeros(block_size, F)
        self.w = 0
        scale = F(1.0)
        if normalize:
            scale = self.normalization_scale(ir)
        for p in range(self.P):                               # PrepareImpulseResponse :65-91
            seg = ir[p * block_size:(p + 1) * block_size]
            t = np.zeros(self.N, np.float64)
            t[:seg.shape[0]] = (seg * scale).astype(np.float64)   # float32 product, then widened (:80)
            X = np.fft.rfft(t)
            self.ir_re[p] = X.real.astype(F)
            self.ir_im[p] = X.imag.astype(F)

    @staticmethod
    def normalization_scale(ir):                              # :93-102
        s = 0.0
        for v in ir:                                          # double sum of float32 squares, in order
            s += float(F(v * v))
        power = F(math.sqrt(s / ir.shape[0]))
        if math.isnan(power) or math.isinf(power) or power < F(0.000125):
            power = F(0.000125)
        return F(F(F(1.0) / power) * F(math.pow(10.0, float(F(F(-58.0) * F(0.05))))))

    def process(self, x):                                     # Process :104-152
        t = np.zeros(self.N, np.float64)
        t[:self.B] = x
        X = np.fft.rfft(t)
        self.dl_re[self.w] = X.real.astype(F)
        self.dl_im[self.w] = X.imag.astype(F)
        acc_r = np.zeros(self.C, F)                           # ProcessSpectralConvolution :154-223
        acc_i = np.zeros(self.C, F)
        for p in range(self.P):
            d = self.w + p
            if d >= self.P:
                d -= self.P
            dr, di, ir_, ii = self.dl_re[d], self.dl_im[d], self.ir_re[p], self.ir_im[p]
            acc_r = acc_r + ((dr * ir_) - (di * ii))          # separate float32 multiply / subtract / add
            acc_i = acc_i + ((dr * ii) + (di * ir_))
        self.w -= 1
        if self.w < 0:
            self.w = self.P - 1
        r = np.fft.irfft(acc_r.astype(np.float64) + 1j * acc_i.astype(np.float64), self.N)
        out = r[:self.B].astype(F) + self.overlap
        self.overlap = r[self.B:].astype(F)
        return out


class ConvolverNode(AudioNode):           # Nodes/ConvolverNode.cs
    def __init__(self, context):
        super().__init__(context)
        self.Normalize = True
        self.EnableTrueStereo = True
        self._buffer = None
        self._conv = None
        self._true_stereo = False
        self._n_out = 0
        self._out = None

    @property
    def Buffer(self):
        return self._buffer

    @Buffer.setter
    def Buffer(self, value):              # :25-79
        if value is self._buffer:
            return
        if value is None:
            self._buffer, self._conv, self._n_out, self._true_stereo = None, None, 0, False
            self.inputs[0].mode = "max"
            return
        if value.SampleRate != self.Context.SampleRate:
            raise InvalidOperationException("Impulse response buffer sample rate must match the audio context sample rate.")
        self._conv = [PartitionedConvolver(c, FRAMES, self.Normalize) for c in value.channels]
        self._buffer = value
        ch = value.NumberOfChannels
        self._true_stereo = ch == 4 and self.EnableTrueStereo
        self._n_out = 2 if self._true_stereo else ch
        self.inputs[0].channel_count = 2 if self._true_stereo else ch
        self.inputs[0].mode = "explicit"

    def process(self):                    # :102-155
        inp = self.inputs[0].buffer
        if self._conv is None:
            if self._out is None or self._out.channels != inp.channels:
                self._out = Block(inp.channels)
            self._out.clear()
            self.outputs[0].buffer = self._out
            return
        if self._out is None or self._out.channels != self._n_out:
            self._out = Block(self._n_out)
        if self._true_stereo:
            self._out.data[0] = self._conv[0].process(inp.data[0]) + self._conv[2].process(inp.data[1])
            self._out.data[1] = self._conv[1].process(inp.data[0]) + self._conv[3].process(inp.data[1])
        else:
            for ch in range(self._n_out):
                self._out.data[ch] = self._conv[ch].process(inp.data[ch])
        self._out.silent = False
        self.outputs[0].buffer = self._out


# ------------------------------------------------------------------------------------------------ CubicResampler.cs
class CubicResampler:
    def __init__(self):
        self.s = [F(0), F(0), F(0), F(0)]
        self.pos = 0.0
        self.ready = 0

    def _shift(self, v):
        self.s = [self.s[1], self.s[2], self.s[3], F(v)]

    def process(self, inp, out, rate):    # :26-63 -> (consumed, produced)
        ip, op = 0, 0
        while self.ready < 4 and ip < len(inp):
            self._shift(inp[ip])
            ip += 1
            self.ready += 1
        if self.ready < 4:
            return ip, op
        while op < len(out):
            consume = int(self.pos)
            if ip + consume > len(inp):
                break
            for _ in range(consume):
                self._shift(inp[ip])
                ip += 1
            self.pos -= consume
            t = F(self.pos)
            s0, s1, s2, s3 = self.s
            inner3 = F(F(F(0.5) * F(s3 - s0)) + F(F(1.5) * F(s1 - s2)))
            inner2 = F(F(F(F(s0 - F(F(2.5) * s1)) + F(F(2.0) * s2)) - F(F(0.5) * s3)) + F(t * inner3))
            inner1 = F(F(F(0.5) * F(s2 - s0)) + F(t * inner2))
            out[op] = F(s1 + F(t * inner1))
            op += 1
            self.pos += rate
        return ip, op


class AudioBufferSourceNode(AudioNode):   # Nodes/AudioBufferSourceNode.cs
    def __init__(self, context):
        super().__init__(context, 0, 1)
        self.PlaybackRate = self._param(1.0, 0.001, 1000.0, a_rate=False)
        self.Buffer = None
        self._started = False
        self._start, self._stop = 0.0, math.nan
        self._offset, self._duration = 0.0, math.inf
        self._position = 0
        self._resamplers = None
        self._out = None
        self.Loop = False                 # :40-44
        self.LoopStart = 0.0              # :49-53 (seconds; the setter clamps at 0)
        self.LoopEnd = 0.0                # :58-62 (0 = end of the buffer)
        self._wrap = None                 # _loopWrapBuffer, 512 floats (:247-250)

    def Start(self, when=0.0, offset=0.0, duration=math.inf):  # :79-114
        if self._started:
            raise InvalidOperationException("AudioBufferSourceNode can only be started once.")
        if self.Buffer is None:
            raise InvalidOperationException("Cannot start without a buffer set")
        self._started = True
        self._start = max(0.0, when)
        self._offset = max(0.0, offset)
        self._duration = duration
        self._position = int(self._offset * self.Buffer.SampleRate)
        if not math.isinf(duration) and duration >= 0:
            self._stop = self._start + duration
            self._has_stopped = True

    def Stop(self, when=0.0):             # :116-129
        if getattr(self, "_has_stopped", False):
            return
        at = max(0.0, when)
        self._stop = at if math.isnan(self._stop) else min(self._stop, at)
        self._has_stopped = True

    def _silence(self):                   # ProduceSilence :391-402
        if self._out is None or self._out.channels != 1:
            self._out = Block(1)
        self._out.clear()
        self.outputs[0].buffer = self._out

    def process(self):                    # :131-376
        ctx = self.Context
        t0 = ctx.current_time
        t1 = t0 + FRAMES / ctx.SampleRate
        play = self._started and t1 > self._start and (math.isnan(self._stop) or t0 < self._stop)
        if not play or self.Buffer is None:
            self._silence()
            return
        buf = self.Buffer
        nch = buf.NumberOfChannels
        if self._out is None or self._out.channels != nch:
            self._out = Block(nch)
        rate = float(F(self.PlaybackRate.values[0]))
        eff = (buf.SampleRate / float(ctx.SampleRate)) * rate
        if self._duration < math.inf:
            dur_end = int(self._offset * buf.SampleRate) + int(self._duration * buf.SampleRate)
        else:
            dur_end = buf.Length
        dur_end = min(dur_end, buf.Length)
        loop = bool(self.Loop)
        loop_start = int(max(0.0, self.LoopStart) * buf.SampleRate)                   # :171
        loop_end = int(max(0.0, self.LoopEnd) * buf.SampleRate) if self.LoopEnd > 0 else buf.Length  # :172-174
        loop_end = min(loop_end, buf.Length)                                          # :176
        loop_start = min(loop_start, loop_end)                                        # :177
        more = False
        if eff == 1.0:                    # :186-235
            for ch in range(nch):
                data, row = buf.channels[ch], self._out.data[ch]
                pos, oi = self._position, 0
                while oi < FRAMES:
                    if loop and pos >= loop_end:                                      # :197-200
                        pos = loop_start
                    if pos >= dur_end and not loop:                                   # :202-206
                        row[oi:] = 0
                        break
                    end = loop_end if loop else min(dur_end, buf.Length)              # :208
                    avail = int(min(end - pos, FRAMES - oi))
                    if avail <= 0:
                        row[oi:] = 0
                        break
                    row[oi:oi + avail] = data[pos:pos + avail]
                    pos += avail
                    oi += avail
                    more = True
            self._position += FRAMES                                                  # :224
            if loop and self._position >= loop_end and loop_end - loop_start > 0:     # :226-234
                self._position = loop_start + (self._position - loop_end) % (loop_end - loop_start)
        else:                             # :236-358
            if self._resamplers is None or len(self._resamplers) != nch:
                self._resamplers = [CubicResampler() for _ in range(nch)]
            total = 0
            for ch in range(nch):
                data, row = buf.channels[ch], self._out.data[ch]
                pos, consumed_ch, oi = self._position, 0, 0
                rs = self._resamplers[ch]
                if self._wrap is None:
                    self._wrap = np.zeros(512, F)
                while oi < FRAMES:
                    if loop and pos >= loop_end:                                      # :265-268
                        pos = loop_start
                    if pos >= dur_end and not loop:                                   # :270-274
                        row[oi:] = 0
                        break
                    end = loop_end if loop else min(dur_end, buf.Length)              # :276
                    avail = int(min(end - pos, buf.Length - pos))                     # :277
                    if avail <= 0:                                                    # :279-292
                        if loop:
                            raise NotImplementedError("empty loop region: the reference never leaves this branch")
                        row[oi:] = 0
                        break
                    out_slice = row[oi:]
                    if loop and pos + avail >= loop_end - 4:                          # :296 (always, since avail = loop_end - pos)
                        loop_len = loop_end - loop_start
                        from_end = int(loop_end - pos)
                        needed = min(FRAMES - oi + 4, 512)                            # :301
                        copied = 0
                        i = 0
                        while i < from_end and copied < needed:                       # :303-306
                            self._wrap[copied] = data[pos + i]
                            copied += 1
                            i += 1
                        i = 0
                        while copied < needed and i < loop_len:                       # :308-311
                            self._wrap[copied] = data[loop_start + i]
                            copied += 1
                            i += 1
                        c, p = rs.process(self._wrap[:copied], out_slice, eff)        # :313
                    else:
                        c, p = rs.process(data[pos:pos + avail], out_slice, eff)      # :317
                    if p > 0:
                        more = True
                    new_pos = pos + c
                    if loop and new_pos >= loop_end:                                  # :323-328
                        new_pos = loop_start + (new_pos - loop_end)
                    consumed_ch += (new_pos - pos) if new_pos >= pos else (loop_end - pos + new_pos - loop_start)  # :330
                    pos = new_pos
                    oi += p
                    if c == 0 and p == 0:
                        row[oi:] = 0
                        break
                if ch == 0:
                    total = consumed_ch
            self._position += total                                                   # :347
            if loop and self._position >= loop_end and loop_end - loop_start > 0:     # :349-357
                self._position = loop_start + (self._position - loop_end) % (loop_end - loop_start)
        if not more or (not loop and self._position >= dur_end):  # :360-372
            self._out.clear()
            if math.isnan(self._stop):
                self._stop = t1
                self._has_stopped = True
        else:
            self._out.silent = False
        self.outputs[0].buffer = self._out


# ------------------------------------------------------------------------------------------------ contexts
class OfflineAudioContext:                # AudioContextBase.cs + OfflineAudioContext.cs
    def __init__(self, sampleRate=48000):
        if sampleRate <= 0:
            raise ArgumentOutOfRangeException("sampleRate")
        self.SampleRate = int(sampleRate)
        self.current_time = 0.0
        self._block = 0
        self.Destination = AudioDestinationNode(self)
        self._cache = None                # unread tail of the last block (:55-100)

    def _process_block(self):             # AudioContextBase.ProcessBlock :52-81
        self._block += 1
        t = self.current_time
        self.Destination.process_internal(self._block, t)
        self.current_time = t + FRAMES / self.SampleRate
        return self.Destination.out

    def Render(self, frameCount):         # OfflineAudioContext.Render :30-124 (two output channels)
        out = np.zeros((2, frameCount), F)
        written = 0
        if self._cache is not None and self._cache.shape[1] > 0:
            n = min(self._cache.shape[1], frameCount)
            out[:, :n] = self._cache[:, :n]
            self._cache = self._cache[:, n:]
            written = n
        while written < frameCount:
            blk = self._process_block()
            n = min(FRAMES, frameCount - written)
            rows = blk.data if blk.channels >= 2 else np.vstack([blk.data[0], np.zeros(FRAMES, F)])
            out[:, written:written + n] = rows[:2, :n]
            written += n
            if n < FRAMES:
                self._cache = rows[:2, n:].copy()
        return out
