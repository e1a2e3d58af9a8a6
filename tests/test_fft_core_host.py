"""CPU: the per-thread phases of the radix-8 / radix-16 FFT cores (graphaudio_b200/csrc/fft2_core.cuh) replayed on the host.

The device kernels (K5 / K6 / K7) call these phases with barriers in between; tools/fft2_host_test.cu calls the SAME functions
thread by thread, which validates the in-place index algebra, the twiddle tables, the forward/inverse mirror property and the
frequency index held by every register slot (what the split step of the real transforms relies on) against double-precision
DFTs and circular convolutions — without a GPU.  nvcc is only used as the host compiler here."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fft_cores_replayed_on_the_host(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "fft2_host_test")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "graphaudio_b200", "csrc"),
                           os.path.join(ROOT, "tools", "fft2_host_test.cu"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.splitlines()
    assert sum(l.startswith("map") for l in lines) == 3          # slot -> frequency maps of the 128 / 256 / 512 plans
    assert sum(l.startswith("r16 M=") for l in lines) == 6       # radix-16 plans 128 ... 4096
    assert sum(l.startswith("M=") for l in lines) == 5           # radix-8 plans 512 ... 8192
