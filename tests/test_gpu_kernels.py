"""Kernel-level parity (GPU): every sm_100a kernel against the CPU oracle, through the C ABI.

Tolerances (written per test):
  * integer / index / unfused-float paths (stream MAC, automation, resampler): bit-exact
  * float32-FFT paths: max |err| <= 1e-5 of full scale (BASELINE.json north_star); observed ~2e-7
"""
import ctypes as C

import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu

TOL = 1e-5  # BASELINE.json: max |err| <= 1e-5 of full scale (float32)


@pytest.fixture(scope="module")
def ctx128():
    import graphaudio_b200 as G
    c = G.OfflineAudioContext(48000)
    yield c
    c.Dispose()


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _native():
    from graphaudio_b200 import _native as N
    from graphaudio_b200.api import check
    return N, check


def pack(spec):
    """numpy rfft output [.., B+1] complex -> packed [.., B] complex64 (bin 0 = DC.re + i*Nyquist.re)."""
    out = spec[..., :-1].astype(np.complex64).copy()
    out[..., 0] = spec[..., 0].real + 1j * spec[..., -1].real
    return out


def unpack(p):
    B = p.shape[-1]
    out = np.zeros(p.shape[:-1] + (B + 1,), np.complex128)
    out[..., :B] = p
    out[..., 0] = p[..., 0].real
    out[..., B] = p[..., 0].imag
    return out


@pytest.mark.parametrize("B", [128, 256, 512])
def test_rfft_forward_matches_oracle(B):
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    N, check = _native()
    ctx = G.OfflineAudioContext(48000, partition=B)
    S, Q = 3, 37
    x = np.stack([synth.splitmix_uniform(10 + s, Q * B) for s in range(S)])
    spec = np.zeros((S, Q, B), np.complex64)
    check(N.lib().gac_rfft_fwd_batch(ctx._h, _fp(x), S, Q, _fp(spec.view(np.float32))))
    # oracle: zero-padded 2B-point double rFFT per block (PartitionedConvolver.cs:106-109), cast to float32
    ref = np.zeros((S, Q, B + 1), np.complex128)
    for s in range(S):
        for b in range(Q):
            blk = np.zeros(2 * B)
            blk[:B] = x[s, b * B:(b + 1) * B]
            ref[s, b] = O.rfft_forward(blk)
    err = np.abs(unpack(spec) - ref).max()
    scale = np.abs(ref).max()
    assert err <= 2e-6 * scale, (err, scale)  # float32 FFT vs double FFT rounded to float32
    ctx.Dispose()


def _mac_reference_unfused(X, H):
    """ProcessSpectralConvolution restated with float32 numpy ops in the reference's order (p ascending, mul/sub/add)."""
    Q, Bn = X.shape
    P = H.shape[0]
    xr, xi = X.real.astype(np.float32), X.imag.astype(np.float32)
    hr, hi = H.real.astype(np.float32), H.imag.astype(np.float32)
    yr = np.zeros((Q, Bn), np.float32)
    yi = np.zeros((Q, Bn), np.float32)
    for p in range(min(P, Q)):
        a = xr[:Q - p] if p else xr
        b = xi[:Q - p] if p else xi
        re = a * hr[p] - b * hi[p]
        im = a * hi[p] + b * hr[p]
        # bin 0 packs two real bins: (DC, Nyquist)
        re[:, 0] = a[:, 0] * hr[p, 0]
        im[:, 0] = b[:, 0] * hi[p, 0]
        yr[p:] = yr[p:] + re
        yi[p:] = yi[p:] + im
    return yr + 1j * yi


@pytest.mark.parametrize("variant,Q,P", [(1, 70, 21), (0, 70, 21), (0, 200, 47), (0, 33, 100), (0, 64, 16), (0, 1, 1)])
def test_spectral_mac(ctx128, variant, Q, P):
    N, check = _native()
    S, B = 2, 128
    rng = np.random.default_rng(Q * 1000 + P)
    X = (rng.uniform(-1, 1, (S, Q, B)) + 1j * rng.uniform(-1, 1, (S, Q, B))).astype(np.complex64)
    H = (rng.uniform(-1, 1, (S, P, B)) + 1j * rng.uniform(-1, 1, (S, P, B))).astype(np.complex64) * np.float32(0.05)
    Y = np.zeros((S, Q, B), np.complex64)
    check(N.lib().gac_spectral_mac(ctx128._h, _fp(X.view(np.float32)), _fp(H.view(np.float32)), S, Q, P, variant, _fp(Y.view(np.float32))))
    for s in range(S):
        ref = _mac_reference_unfused(X[s], H[s])
        if variant == 1:
            # the streaming kernel keeps the reference's op order and rounding: bit-exact
            assert np.array_equal(Y[s].view(np.float32), ref.astype(np.complex64).view(np.float32))
        else:
            # the tiled kernel uses FMA (one rounding fewer per term): within a few ulp of the accumulated magnitude
            err = np.abs(Y[s] - ref).max()
            assert err <= 1e-5 * max(1.0, np.abs(ref).max()), err


def test_irfft_ola_matches_numpy(ctx128):
    N, check = _native()
    S, Q, B = 2, 50, 128
    rng = np.random.default_rng(7)
    t = rng.uniform(-1, 1, (S, Q, 2 * B))
    spec = np.fft.rfft(t, axis=-1)
    Yp = pack(spec)
    y = np.zeros((S, Q * B), np.float32)
    check(N.lib().gac_irfft_ola_batch(ctx128._h, _fp(Yp.view(np.float32)), S, Q, _fp(y)))
    # reference: r = irfft(Y) ; out[b] = float32(r[:B]) + float32(r_prev[B:])   (PartitionedConvolver.cs:146-150)
    r = np.fft.irfft(unpack(Yp), axis=-1).astype(np.float32)
    ref = r[:, :, :B].copy()
    ref[:, 1:, :] += r[:, :-1, B:]
    assert np.abs(y.reshape(S, Q, B) - ref).max() <= 2e-6


@pytest.mark.parametrize("B,ir_len,n,normalize", [(128, 4800, 24000, True), (128, 300, 5000, False), (512, 9000, 30000, True), (256, 1000, 10000, True)])
def test_convolve_batch_matches_oracle(B, ir_len, n, normalize):
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    N, check = _native()
    ctx = G.OfflineAudioContext(48000, partition=B)
    S = 3
    x = np.stack([synth.splitmix_uniform(20 + s, n) for s in range(S)])
    ir = np.stack([synth.decay_ir(40 + s, ir_len) for s in range(S)])
    y = np.zeros((S, n), np.float32)
    check(N.lib().gac_convolve_batch(ctx._h, _fp(x), S, n, _fp(ir), ir_len, int(normalize), _fp(y)))
    nb = n // B
    for s in range(S):
        pc = O.PartitionedConvolver(ir[s], B, normalize)
        ref = pc.process(x[s, :nb * B])
        err = np.abs(y[s, :nb * B] - ref).max()
        assert err <= TOL, (s, err)
    ctx.Dispose()


def test_convolver_delta_ir_is_a_delay(ctx128):
    """KAT: Normalize=false, IR = delta[n-d] => output = input delayed by d."""
    N, check = _native()
    n, d = 128 * 20, 200
    x = synth.splitmix_uniform(3, n)[None, :].copy()
    ir = np.zeros((1, 300), np.float32)
    ir[0, d] = 1.0
    y = np.zeros((1, n), np.float32)
    check(N.lib().gac_convolve_batch(ctx128._h, _fp(x), 1, n, _fp(ir), 300, 0, _fp(y)))
    assert np.abs(y[0, d:] - x[0, :n - d]).max() <= 5e-7
    assert np.abs(y[0, :d]).max() <= 5e-7


def _param_struct(N, value, events):
    p = N.gac_param()
    p.value = value
    p.n_events = len(events)
    arr = (N.gac_event * max(1, len(events)))()
    for i, (t, v, tg, tm, tc) in enumerate(events):
        arr[i].type, arr[i].value, arr[i].target, arr[i].time, arr[i].time_constant = t, v, tg, tm, tc
    p.events = arr
    return p, arr


@pytest.mark.parametrize("a_rate", [1, 0])
def test_automation_matches_oracle(ctx128, a_rate):
    """SetValue -> LinearRamp -> ExponentialRamp -> SetTarget schedule, sampled at t = T_b + i/fs (AudioParam.cs:114-247)."""
    from oracle import ga_oracle as O
    N, check = _native()
    octx = O.OfflineAudioContext(48000)
    if a_rate:
        node = O.GainNode(octx)
        op = node.Gain
    else:
        node = O.BiQuadFilterNode(octx)
        op = node.Gain
    sched = [(0, 0.3, 0.0), (1, 0.9, 0.11), (2, 0.2, 0.23), (3, 0.05, 0.23, 0.04), (1, 0.7, 0.5), (0, 0.4, 0.61)]
    events = []
    for e in sched:
        if e[0] == 0:
            op.SetValueAtTime(e[1], e[2]); events.append((0, e[1], 0.0, e[2], 0.0))
        elif e[0] == 1:
            op.LinearRampToValueAtTime(e[1], e[2]); events.append((1, e[1], 0.0, e[2], 0.0))
        elif e[0] == 2:
            op.ExponentialRampToValueAtTime(e[1], e[2]); events.append((2, e[1], 0.0, e[2], 0.0))
        else:
            op.SetTargetAtTime(e[1], e[2], e[3]); events.append((3, 0.0, e[1], e[2], e[3]))
    nq = 300
    ref = op.evaluate(nq)
    p, keep = _param_struct(N, op.DefaultValue, events)
    out = np.zeros(nq * 128, np.float32)
    check(N.lib().gac_automation_eval(ctx128._h, C.byref(p), a_rate, nq * 128, _fp(out)))
    # double-precision pow/exp on the device differ from glibc by <= 2 ulp(double): identical after the float32 cast
    # except on rounding ties; allow 1 ulp(float32) on a vanishing fraction
    diff = np.abs(out - ref)
    assert diff.max() <= 1.2e-7 * max(1.0, np.abs(ref).max())
    assert (diff > 0).mean() < 1e-3


def test_first_event_ramp_is_a_step(ctx128):
    """KAT: a ramp that is the first event does not ramp (AudioParam.cs:181-184)."""
    N, check = _native()
    p, keep = _param_struct(N, 0.25, [(1, 1.0, 0.0, 0.01, 0.0)])
    out = np.zeros(1024, np.float32)
    check(N.lib().gac_automation_eval(ctx128._h, C.byref(p), 1, 1024, _fp(out)))
    k = int(np.ceil(0.01 * 48000))
    assert np.all(out[:k] == np.float32(0.25))
    assert np.all(out[k + 1:] == np.float32(1.0))


@pytest.mark.parametrize("rate,n_in,n_out", [(0.459375, 5000, 12000), (1.5, 5000, 2000), (2.75, 3000, 4000), (0.459375, 3, 10), (1.0001, 4000, 3000)])
def test_cubic_resampler_matches_oracle(ctx128, rate, n_in, n_out):
    from oracle import ga_oracle as O
    N, check = _native()
    x = synth.splitmix_uniform(77, n_in)
    ref, ref_consumed = O.resample(x, n_out, rate)
    out = np.zeros(n_out, np.float32)
    produced, consumed = C.c_int64(0), C.c_int64(0)
    check(N.lib().gac_resample_cubic(ctx128._h, _fp(x), n_in, rate, n_out, _fp(out), C.byref(produced), C.byref(consumed)))
    assert produced.value == ref.shape[0]
    assert consumed.value == ref_consumed
    # phase recurrence replayed exactly + unfused float32 polynomial: bit-exact
    assert np.array_equal(out[:produced.value], ref)
    if produced.value:
        assert out[0] == x[1]  # KAT: primed by 4 samples, first output is in[1] (CubicResampler.cs:31-35,51-52)
