"""Synthetic inputs and graph builders shared by the parity tests, bench.py and smoke().

Inputs follow SURVEY.md §8(d): SplitMix64 streams, u = ((x >> 40) * 2^-23) - 1 in [-1, 1) (exact in float32, no
transcendental functions in generation); source = white noise; IR = noise * exponential envelope reaching -60 dB.
Graph builders take an API module (`graphaudio_b200` = CUDA path, or `oracle.ga_oracle` = CPU oracle) so the very
same construction code drives both sides of a parity check — written the way a user of the reference would write it
(OfflineAudioContext / AudioBufferSourceNode / BiQuadFilterNode / GainNode / ConvolverNode / Connect / Render).
"""
from __future__ import annotations

import numpy as np

MASK = (1 << 64) - 1


def splitmix_uniform(stream: int, n: int) -> np.ndarray:
    """n samples of stream `stream`, float32 in [-1, 1)."""
    seed = (0x9E3779B97F4A7C15 * (1 + stream)) & MASK
    idx = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(40)).astype(np.float64) * 2.0 ** -23 - 1.0).astype(np.float32)


def decay_ir(stream: int, n: int) -> np.ndarray:
    """noise * env, env[0] = 1, env[k+1] = env[k] * d in float32, d = float32(10^(-3/n)) (-60 dB at the end)."""
    d = np.float32(10.0 ** (-3.0 / n))
    # float32 running product, reproduced exactly with a cumulative product in float32
    env = np.cumprod(np.concatenate([[np.float32(1.0)], np.full(n - 1, d, np.float32)]), dtype=np.float32)
    return (splitmix_uniform(stream, n) * env).astype(np.float32)


def voice_gains(v: int):
    """g0, g1, g2 in [0.25, 1] from the PRNG."""
    u = splitmix_uniform(100000 + v, 3).astype(np.float64)
    return [float(np.float32(0.625 + 0.375 * x)) for x in u]


def make_voice_inputs(v: int, src_frames: int, ir_frames: int, src_channels: int = 2):
    src = [splitmix_uniform(4 * v + c, src_frames) for c in range(src_channels)]
    ir = [decay_ir(4 * v + 2 + c, ir_frames) for c in range(2)]
    return src, ir


# ----------------------------------------------------------------------------------------------- graph builders
def build_c1(api, fs, src, ir, **ctx_kw):
    """C1: AudioBufferSourceNode(stereo) -> ConvolverNode(stereo IR) -> destination."""
    ctx = api.OfflineAudioContext(fs, **ctx_kw)
    s = api.AudioBufferSourceNode(ctx)
    s.Buffer = api.PlayableAudioBuffer.FromChannelArrays(src, fs)
    conv = api.ConvolverNode(ctx)
    conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays(ir, fs)
    s.Connect(conv).Connect(ctx.Destination)
    s.Start()
    return ctx


def add_gain_automation(gain_param, g, t_scale=1.0):
    """C2 automation: SetValueAtTime(g0,0), LinearRamp(g1,5), ExponentialRamp(g2,10), SetTarget(0,10,0.5)."""
    gain_param.SetValueAtTime(g[0], 0.0)
    gain_param.LinearRampToValueAtTime(g[1], 5.0 * t_scale)
    gain_param.ExponentialRampToValueAtTime(g[2], 10.0 * t_scale)
    gain_param.SetTargetAtTime(0.0, 10.0 * t_scale, 0.5 * t_scale)


def build_c2(api, fs, voices, bus_gain, t_scale=1.0, **ctx_kw):
    """C2: per voice Source -> GainNode(automation) -> ConvolverNode(per-voice stereo IR) -> bus GainNode -> destination.
    voices: list of (src_channels, ir_channels, gains)."""
    ctx = api.OfflineAudioContext(fs, **ctx_kw)
    bus = api.GainNode(ctx)
    bus.Gain.Value = bus_gain
    bus.Connect(ctx.Destination)
    ctx.bus = bus  # handle for callers that shard the voices over ranks (OfflineAudioContext.MarkBus)
    for src, ir, g in voices:
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays(src, fs)
        gn = api.GainNode(ctx)
        add_gain_automation(gn.Gain, g, t_scale)
        conv = api.ConvolverNode(ctx)
        conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays(ir, fs)
        s.Connect(gn).Connect(conv).Connect(bus)
        s.Start()
    return ctx


def build_c3(api, fs, voices, bus_gain, f0=2000.0, f1=12000.0, t_scale=1.0, q=0.707, **ctx_kw):
    """C3: Source -> BiQuadFilterNode(Lowpass, Q, a-rate cutoff sweep) -> GainNode -> ConvolverNode -> bus GainNode -> destination."""
    ctx = api.OfflineAudioContext(fs, **ctx_kw)
    bus = api.GainNode(ctx)
    bus.Gain.Value = bus_gain
    bus.Connect(ctx.Destination)
    ctx.bus = bus  # handle for callers that shard the voices over ranks (OfflineAudioContext.MarkBus)
    for src, ir, g in voices:
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays(src, fs)
        bq = api.BiQuadFilterNode(ctx)
        bq.Type = api.FilterType.Lowpass
        bq.Q.Value = q
        bq.Frequency.SetValueAtTime(f0, 0.0)
        bq.Frequency.ExponentialRampToValueAtTime(f1, 10.0 * t_scale)
        gn = api.GainNode(ctx)
        add_gain_automation(gn.Gain, g, t_scale)
        conv = api.ConvolverNode(ctx)
        conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays(ir, fs)
        s.Connect(bq).Connect(gn).Connect(conv).Connect(bus)
        s.Start()
    return ctx


def build_c4(api, fs, src, ir, t_scale=1.0, **ctx_kw):
    """C4 (one of the independent renders): Source -> BiQuad(Lowpass sweep) -> BiQuad(Highpass 200 Hz) -> Convolver -> destination."""
    ctx = api.OfflineAudioContext(fs, **ctx_kw)
    s = api.AudioBufferSourceNode(ctx)
    s.Buffer = api.PlayableAudioBuffer.FromChannelArrays(src, fs)
    lp = api.BiQuadFilterNode(ctx)
    lp.Type = api.FilterType.Lowpass
    lp.Q.Value = 0.707
    lp.Frequency.SetValueAtTime(2000.0, 0.0)
    lp.Frequency.ExponentialRampToValueAtTime(12000.0, 4.0 * t_scale)
    hp = api.BiQuadFilterNode(ctx)
    hp.Type = api.FilterType.Highpass
    hp.Frequency.Value = 200.0
    hp.Q.Value = 0.707
    conv = api.ConvolverNode(ctx)
    conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays(ir, fs)
    s.Connect(lp).Connect(hp).Connect(conv).Connect(ctx.Destination)
    s.Start()
    return ctx


def build_c5(api, fs, src_rate, voices, bus_gain, t_scale=1.0, loop=None, rate_ramp=None, **ctx_kw):
    """C5: Source(44.1 kHz buffer in a 96 kHz context -> CubicResampler) -> GainNode -> Convolver -> bus -> destination.
    loop = (LoopStart, LoopEnd) in seconds: the sources loop; rate_ramp = (r0, r1, t1): PlaybackRate.SetValueAtTime(r0, 0) and
    LinearRampToValueAtTime(r1, t1) (the "loop" cross-check case: the wrap-buffer resampler path under a k-rate rate sweep)."""
    ctx = api.OfflineAudioContext(fs, **ctx_kw)
    bus = api.GainNode(ctx)
    bus.Gain.Value = bus_gain
    bus.Connect(ctx.Destination)
    ctx.bus = bus  # handle for callers that shard the voices over ranks (OfflineAudioContext.MarkBus)
    for src, ir, g in voices:
        s = api.AudioBufferSourceNode(ctx)
        s.Buffer = api.PlayableAudioBuffer.FromChannelArrays(src, src_rate)
        if loop is not None:
            s.Loop = True
            s.LoopStart, s.LoopEnd = loop
        if rate_ramp is not None:
            s.PlaybackRate.SetValueAtTime(rate_ramp[0], 0.0)
            s.PlaybackRate.LinearRampToValueAtTime(rate_ramp[1], rate_ramp[2])
        gn = api.GainNode(ctx)
        add_gain_automation(gn.Gain, g, t_scale)
        conv = api.ConvolverNode(ctx)
        conv.Buffer = api.PlayableAudioBuffer.FromChannelArrays(ir, fs)
        s.Connect(gn).Connect(conv).Connect(bus)
        s.Start()
    return ctx


def loop_resample_model(x, pos0, loop_start, loop_end, eff, n_blocks):
    """Index-level model of a LOOPING AudioBufferSourceNode on the CubicResampler path (Nodes/AudioBufferSourceNode.cs:236-358 with
    _loop set), one channel.  Which buffer frames are shifted into the resampler and the phase of every output depend on positions only,
    never on sample values, so the model replays the block loop with indices: per Process call the wrap buffer holds the frames
    pos .. loopEnd-1 followed by ONE pass over the loop region, cut at min(128 - outIdx + 4, 512) entries (:296-314); a call that
    neither consumes nor produces clears the rest of the quantum (:334-338); a quantum without any output ends the source (:360-368).
    Returns (y, n_live_blocks): y[n_blocks * 128] float32 and the number of leading quanta that are flagged non-silent."""
    x = np.asarray(x, dtype=np.float32)
    idx = np.zeros((n_blocks * 128, 4), dtype=np.int64)
    tt = np.zeros(n_blocks * 128, dtype=np.float32)
    live = np.zeros(n_blocks * 128, dtype=bool)
    win, ready, Pos = [0, 0, 0, 0], 0, 0.0
    position = pos0
    n_live = 0
    for b in range(n_blocks):
        pos, consumed_ch, oi, more = position, 0, 0, False
        while oi < 128:
            if pos >= loop_end:
                pos = loop_start
            from_end = loop_end - pos
            needed = min(128 - oi + 4, 512)
            wrap = list(range(pos, pos + min(from_end, needed)))
            wrap += list(range(loop_start, loop_start + min(loop_end - loop_start, needed - len(wrap))))
            ip, op = 0, 0
            while ready < 4 and ip < len(wrap):   # CubicResampler.cs:31-35
                win = win[1:] + [wrap[ip]]
                ip += 1
                ready += 1
            if ready == 4:
                while oi + op < 128:               # :40-60
                    consume = int(Pos)
                    if ip + consume > len(wrap):
                        break
                    for _ in range(consume):
                        win = win[1:] + [wrap[ip]]
                        ip += 1
                    Pos -= consume
                    idx[b * 128 + oi + op] = win
                    tt[b * 128 + oi + op] = np.float32(Pos)
                    live[b * 128 + oi + op] = True
                    op += 1
                    Pos += eff
            more = more or op > 0
            new_pos = pos + ip
            if new_pos >= loop_end:
                new_pos = loop_start + (new_pos - loop_end)
            consumed_ch += (new_pos - pos) if new_pos >= pos else (loop_end - pos + new_pos - loop_start)
            pos, oi = new_pos, oi + op
            if ip == 0 and op == 0:
                break
        position += consumed_ch
        if position >= loop_end:
            position = loop_start + (position - loop_end) % (loop_end - loop_start)
        if not more:
            live[b * 128:(b + 1) * 128] = False
            break
        n_live = b + 1
    S0, S1, S2, S3 = (x[idx[:, i]] for i in range(4))
    h, q, t = np.float32(0.5), np.float32(1.5), tt
    y = S1 + t * (h * (S2 - S0) + t * ((S0 - np.float32(2.5) * S1 + np.float32(2.0) * S2 - h * S3) + t * (h * (S3 - S0) + q * (S1 - S2))))
    return np.where(live, y, np.float32(0)).astype(np.float32), n_live
