"""GPU: GAC_FLAG_ASYNC_UPLOAD — page-locked source arrays are uploaded on a copy stream while IR preparation and the first
voice batches run; the result must be bit-identical to the synchronous path (same kernels, same order per voice)."""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu


def _pinned(a):
    import torch
    t = torch.empty(a.shape[0], dtype=torch.float32, pin_memory=True)
    v = t.numpy()
    v[:] = a
    return v, t


def test_async_upload_matches_sync_and_oracle():
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    fs = 48000
    keep = []
    voices = []
    for v in range(20):  # >= 16 voices: exercises the batched schedule behind pending uploads
        src, ir = synth.make_voice_inputs(v, 60000, 5000)
        ps, pi = [], []
        for a in src:
            b, t = _pinned(a); ps.append(b); keep.append(t)
        for a in ir:
            b, t = _pinned(a); pi.append(b); keep.append(t)
        voices.append((ps, pi, synth.voice_gains(v)))
    n = 70000
    ya = synth.build_c2(G, fs, voices, 0.1, t_scale=0.1, async_upload=True).Render(n)
    ys = synth.build_c2(G, fs, voices, 0.1, t_scale=0.1, async_upload=False).Render(n)
    assert np.array_equal(ya, ys)
    yo = synth.build_c2(O, fs, voices[:4], 0.1, t_scale=0.1).Render(20000)
    y4 = synth.build_c2(G, fs, voices[:4], 0.1, t_scale=0.1, async_upload=True).Render(20000)
    assert np.abs(y4 - yo).max() <= 1e-5


def test_async_flag_with_pageable_memory_falls_back_to_copy_during_call():
    import graphaudio_b200 as G
    fs = 48000
    src, ir = synth.make_voice_inputs(3, 20000, 3000)
    ya = synth.build_c1(G, fs, src, ir, async_upload=True).Render(24000)   # pageable numpy arrays: copied during the call
    yb = synth.build_c1(G, fs, src, ir, async_upload=False).Render(24000)
    assert np.array_equal(ya, yb) and ya.any()


def test_deferred_ir_preparation_survives_early_buffer_destroy_and_unused_irs():
    """Async mode defers IR preparation to the first render that uses the IR; until then the IR keeps its source buffer alive even
    if the caller destroys the buffer handle, and an IR that is never used releases it when it is destroyed."""
    import ctypes as C
    import graphaudio_b200 as G
    from graphaudio_b200 import _native as N
    from graphaudio_b200.api import check
    fs = 48000
    L = N.lib()
    src, ir = synth.make_voice_inputs(7, 30000, 9000)
    keep = []
    pin = lambda xs: [(_pinned(a)) for a in xs]  # noqa: E731
    ps, pi, pu = pin(src), pin(ir), pin(ir)
    keep += [t for _, t in ps + pi + pu]

    def render(async_upload, destroy_ir_buffer_early):
        ctx = G.OfflineAudioContext(fs, async_upload=async_upload)
        s = G.AudioBufferSourceNode(ctx)
        s.Buffer = G.PlayableAudioBuffer.FromChannelArrays([a for a, _ in ps], fs)
        conv = G.ConvolverNode(ctx)
        irbuf = G.PlayableAudioBuffer.FromChannelArrays([a for a, _ in pi], fs)
        conv.Buffer = irbuf
        unused = G.ConvolverNode(ctx)  # prepared (deferred) but never connected: its IR is never used by a render
        unused.Buffer = G.PlayableAudioBuffer.FromChannelArrays([a for a, _ in pu], fs)
        if destroy_ir_buffer_early:
            h = irbuf._handles.pop((ctx._serial, 0))
            ctx._owned_buffers.remove(h)
            check(L.gac_buffer_destroy(h))  # the handle is gone for the caller; the deferred preparation still reads the data
        s.Connect(conv).Connect(ctx.Destination)
        s.Start()
        y = ctx.Render(40000)
        y2 = ctx.Render(2000)  # a second render of the same context: the IR is prepared now, nothing is deferred any more
        ctx.Dispose()
        return y, y2
    ya, ya2 = render(True, True)
    yb, yb2 = render(True, False)
    ys, ys2 = render(False, False)
    assert ya.any() and np.array_equal(ya, yb) and np.array_equal(ya, ys)
    assert np.array_equal(ya2, ys2) and np.array_equal(yb2, ys2)


def test_impulse_responses_prepared_ahead_of_the_first_render():
    """Async mode: every 32 registrations the impulse responses whose samples have landed are prepared in one batch, ahead of the render
    that first uses them (kick_deferred_irs).  70 voices with IRs long enough for the second-level-FFT path: the result equals the
    synchronous context's, the oracle's on a subset, and a SECOND render of the same context (which may switch to mixed segment lengths:
    the double-length spectra are prepared then) agrees with the first."""
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    fs = 48000
    voices = [synth.make_voice_inputs(300 + v, 30000, 9000) + (synth.voice_gains(v),) for v in range(70)]
    n = 128 * 330
    ca = synth.build_c3(G, fs, voices, 0.1, t_scale=0.05, async_upload=True)
    ya = np.array(ca.Render(n))
    ys = synth.build_c3(G, fs, voices, 0.1, t_scale=0.05, async_upload=False).Render(n)
    assert np.abs(ya).max() > 0.05
    assert np.abs(ya - ys).max() <= 2e-6   # (the synchronous context prepares at creation and may pick another segment plan)
    k = 6
    yo = synth.build_c3(O, fs, voices[:k], 0.1, t_scale=0.05).Render(n)
    yk = synth.build_c3(G, fs, voices[:k], 0.1, t_scale=0.05, async_upload=True).Render(n)
    assert np.abs(yk - yo).max() <= 1e-5
    # the same context again, from frame 0 of a fresh timeline: a second context built from the same buffers shares nothing, so
    # render twice through the C ABI (gac_render with first_frame = 0 re-renders the timeline)
    import ctypes as C
    from graphaudio_b200 import _native as N
    from graphaudio_b200.api import check
    g = ca._graph()
    out = np.zeros((2, n), np.float32)
    ptrs = (N.fp * 2)(*[out[c].ctypes.data_as(N.fp) for c in range(2)])
    check(N.lib().gac_render(ca._h, g, 0, n, ptrs, 2, 0))
    N.lib().gac_graph_destroy(g)
    assert np.abs(out - ya).max() <= 2e-6
