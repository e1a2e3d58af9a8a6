"""GPU: ONE process, several devices.

  * gac_group / OfflineAudioContext(device_ids=[...]): one context whose voices are sharded over the GPUs of the process, one
    ncclReduce of the bus (ncclCommInitAll), against the single-GPU render and the CPU oracle;
  * two independent contexts on two devices of one process rendering concurrently from two threads (every opt-in to more than
    48 KB of dynamic shared memory is per device);
  * the degenerate group of one device runs everywhere (no NCCL involved).
The two-device cases skip on a box with one GPU (run them with `gpurun --gpus 2`)."""
import threading

import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu
FS = 48000


def _n_devices():
    import ctypes as C
    from graphaudio_b200 import _native as N
    n = C.c_int(0)
    N.lib().gac_device_count(C.byref(n))
    return n.value


def _voices(nv, src_frames=30000, ir_frames=9000):
    out = []
    for v in range(nv):
        src, ir = synth.make_voice_inputs(v, src_frames, ir_frames)
        out.append((src, ir, synth.voice_gains(v)))
    return out


def test_group_of_one_device_matches_the_plain_context():
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    voices = _voices(5)
    n = 40000
    g = synth.build_c3(G, FS, voices, 0.25, t_scale=0.05, device_ids=[0])
    yg = g.Render(n)
    g.Dispose()
    p = synth.build_c3(G, FS, voices, 0.25, t_scale=0.05)
    yp = p.Render(n)
    p.Dispose()
    yo = synth.build_c3(O, FS, voices, 0.25, t_scale=0.05).Render(n)
    assert np.abs(yo).max() > 0.02
    assert np.abs(yg - yo).max() <= 1e-5
    assert np.abs(yg - yp).max() <= 1e-6


@pytest.mark.parametrize("nv", [11, 2, 1])
def test_group_render_over_all_devices_of_the_process(nv):
    """uneven shards (11 voices), one voice per member, and fewer voices than members"""
    nd = _n_devices()
    if nd < 2:
        pytest.skip("needs at least two GPUs in one process")
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    voices = _voices(nv)
    n = 40000
    g = synth.build_c2(G, FS, voices, 0.25, t_scale=0.05, device_ids=list(range(nd)))
    y = np.zeros((2, n + 100), np.float32)
    g.Render(y, 3000, 100)            # successive Render calls continue the timeline on every member
    g.Render(y, n - 3000, 3100)
    if nv > 1:
        assert len(g.last_stats_members) == nd
        assert sum(int(s["voices"]) for s in g.last_stats_members) == nv
    g.Dispose()
    yo = synth.build_c2(O, FS, voices, 0.25, t_scale=0.05).Render(n)
    assert np.abs(yo).max() > 0.02
    assert np.abs(y[:, 100:] - yo).max() <= 1e-5


def test_two_contexts_on_two_devices_render_concurrently():
    nd = _n_devices()
    if nd < 2:
        pytest.skip("needs at least two GPUs in one process")
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    n = 128 * 400
    out, errs = {}, []

    def work(dev):
        try:
            voices = [(s, i, g) for s, i, g in _voices(3 + dev, 40000, 128 * 80)]
            ctx = synth.build_c3(G, FS, voices, 0.3, t_scale=0.05, device_id=dev)
            out[dev] = (ctx.Render(n), voices)
            ctx.Dispose()
        except Exception as e:  # noqa: BLE001
            errs.append((dev, repr(e)))

    ts = [threading.Thread(target=work, args=(d,)) for d in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for dev in range(2):
        y, voices = out[dev]
        yo = synth.build_c3(O, FS, voices, 0.3, t_scale=0.05).Render(n)
        assert np.abs(y - yo).max() <= 1e-5
