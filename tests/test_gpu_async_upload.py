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
