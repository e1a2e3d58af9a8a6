"""The C++ host-side mirror (graphaudio_b200/host/graphaudio_cuda.hpp) over the C ABI: compiles everywhere; on a GPU
box it renders the C2-shaped graph and is compared with the CPU oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "mirror_smoke.cpp")
LIBDIR = os.path.join(ROOT, "graphaudio_b200", "lib")


def _build(tmp):
    exe = os.path.join(tmp, "mirror_smoke")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", SRC, "-o", exe, "-L" + LIBDIR, "-lgraphaudio_cuda",
                           "-Wl,-rpath," + LIBDIR])
    return exe


def test_cpp_mirror_compiles_and_fails_loudly_without_a_device(tmp_path):
    from graphaudio_b200 import build
    build.build()
    exe = _build(str(tmp_path))
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; see the gpu test")
    r = subprocess.run([exe, str(tmp_path), "1", "256", "128", "512"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr  # the context constructor throws; nothing renders on the host


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle(tmp_path):
    from oracle import ga_oracle as O
    exe = _build(str(tmp_path))
    nv, ns, ni, n = 3, 20000, 5000, 26000
    voices = []
    for v in range(nv):
        src, ir = synth.make_voice_inputs(v, ns, ni)
        for c in range(2):
            src[c].tofile(os.path.join(tmp_path, f"src_{v}_{c}.f32"))
            ir[c].tofile(os.path.join(tmp_path, f"ir_{v}_{c}.f32"))
        voices.append((src, ir))
    r = subprocess.run([exe, str(tmp_path), str(nv), str(ns), str(ni), str(n)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = np.stack([np.fromfile(os.path.join(tmp_path, f"out_{c}.f32"), np.float32) for c in range(2)])
    ctx = O.OfflineAudioContext(48000)
    bus = O.GainNode(ctx)
    bus.Gain.Value = 0.25
    bus.Connect(ctx.Destination)
    for v, (src, ir) in enumerate(voices):
        s = O.AudioBufferSourceNode(ctx)
        s.Buffer = O.PlayableAudioBuffer.FromChannelArrays(src, 48000)
        g = O.GainNode(ctx)
        g.Gain.SetValueAtTime(0.9, 0.0)
        g.Gain.LinearRampToValueAtTime(0.3, 0.05 * (v + 1))
        g.Gain.ExponentialRampToValueAtTime(0.8, 0.2)
        g.Gain.SetTargetAtTime(0.0, 0.25, 0.05)
        cv = O.ConvolverNode(ctx)
        cv.Buffer = O.PlayableAudioBuffer.FromChannelArrays(ir, 48000)
        s.Connect(g).Connect(cv).Connect(bus)
        s.Start()
    ref = ctx.Render(n)
    assert np.abs(ref).max() > 1e-3
    assert np.abs(out - ref).max() <= 1e-5
