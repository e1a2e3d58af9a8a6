"""GPU: the per-quantum plugin seam (gac_convolver_process_block, SURVEY.md §8b "literal plugin-seam shim") against the CPU oracle.

The C# `CudaConvolverNode.Process()` calls the library once per 128-frame block; here the block loop is driven from Python the
same way and compared with the oracle's PartitionedConvolver / ConvolverNode fed the same blocks
(PartitionedConvolver.cs:104-152, Nodes/ConvolverNode.cs:102-155).
"""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu
TOL = 1e-5
FS = 48000


def _apis():
    import graphaudio_b200 as G
    from oracle import ga_oracle as O
    return G, O


def _node(G, ir_channels, normalize=True, true_stereo=True, **kw):
    ctx = G.OfflineAudioContext(FS, **kw)
    node = G.CudaConvolverNode(ctx)
    node.Normalize = normalize
    node.EnableTrueStereo = true_stereo
    node.Buffer = G.PlayableAudioBuffer.FromChannelArrays(ir_channels, FS)
    return ctx, node


@pytest.mark.parametrize("ir_len", [1, 100, 128, 129, 4000])
def test_mono_block_by_block_matches_partitioned_convolver(ir_len):
    G, O = _apis()
    ir = synth.decay_ir(5, ir_len)
    x = synth.splitmix_uniform(6, 128 * 60)
    ctx, node = _node(G, [ir])
    assert (node.InputChannelCount, node.OutputChannelCount) == (1, 1)
    y = np.concatenate([node.Process(x[None, b * 128:(b + 1) * 128])[0] for b in range(60)])
    yo = O.PartitionedConvolver(ir, 128, True).process(x)
    assert np.abs(yo).max() > 1e-3
    assert np.abs(y - yo).max() <= TOL
    ctx.Dispose()


def test_stereo_tail_continues_through_silent_input_and_reset():
    G, O = _apis()
    irs = [synth.decay_ir(10 + c, 3000) for c in range(2)]
    x = np.stack([synth.splitmix_uniform(20 + c, 128 * 40) for c in range(2)])
    x[:, 128 * 10:] = 0.0  # the convolver keeps running on silent input: the tail rings out (ConvolverNode.cs:146-153)
    ctx, node = _node(G, irs)
    assert (node.InputChannelCount, node.OutputChannelCount) == (2, 2)
    y = np.concatenate([node.Process(x[:, b * 128:(b + 1) * 128]) for b in range(40)], axis=1)
    yo = np.stack([O.PartitionedConvolver(irs[c], 128, True).process(x[c]) for c in range(2)])
    assert np.abs(yo[:, 128 * 12:128 * 20]).max() > 1e-4
    assert np.abs(y - yo).max() <= TOL
    node.Reset()  # a fresh delay line gives the same result again
    y2 = np.concatenate([node.Process(x[:, b * 128:(b + 1) * 128]) for b in range(40)], axis=1)
    assert np.array_equal(y, y2)
    ctx.Dispose()


def test_multi_quantum_call_equals_block_calls_and_history_wraps():
    G, O = _apis()
    ir = synth.decay_ir(31, 2000)  # P = 16: the linear history (16 + 256 + 32 rows) wraps several times in 1000 blocks
    n = 128 * 1000
    x = synth.splitmix_uniform(32, n)
    ctx, node = _node(G, [ir])
    y = node.ProcessFrames(x[None, :])[0]
    node.Reset()
    yb = np.concatenate([node.ProcessFrames(x[None, s:s + 128 * k])[0]
                         for s, k in zip(np.arange(0, n, 128 * 50), [50] * 20)])
    assert np.array_equal(y, yb)
    yo = O.PartitionedConvolver(ir, 128, True).process(x)
    assert np.abs(y - yo).max() <= TOL
    ctx.Dispose()


def test_true_stereo_and_normalize_false_match_the_oracle_node():
    G, O = _apis()
    src = [synth.splitmix_uniform(50 + c, 128 * 80) for c in range(2)]
    ir = [synth.decay_ir(60 + c, 5000) * np.float32(0.5 + 0.1 * c) for c in range(4)]
    for normalize in (True, False):
        ctx, node = _node(G, ir, normalize=normalize)
        assert (node.InputChannelCount, node.OutputChannelCount) == (2, 2)
        y = node.ProcessFrames(np.stack(src))
        # the oracle's ConvolverNode inside its block loop: source -> convolver -> destination
        octx = O.OfflineAudioContext(FS)
        s = O.AudioBufferSourceNode(octx)
        s.Buffer = O.PlayableAudioBuffer.FromChannelArrays(src, FS)
        conv = O.ConvolverNode(octx)
        conv.Normalize = normalize
        conv.Buffer = O.PlayableAudioBuffer.FromChannelArrays(ir, FS)
        s.Connect(conv).Connect(octx.Destination)
        s.Start()
        yo = octx.Render(128 * 79)  # (the source drops its final block, AudioBufferSourceNode.cs:360-368)
        assert np.abs(yo).max() > (1e-3 if normalize else 1e-2)
        assert np.abs(y[:, :128 * 79] - yo).max() <= TOL * max(1.0, np.abs(yo).max())
        ctx.Dispose()


def test_works_in_an_async_upload_context_and_validates_arguments():
    G, O = _apis()
    import torch
    ir_t = torch.from_numpy(np.stack([synth.decay_ir(70, 1000), synth.decay_ir(71, 1000)])).pin_memory()
    ctx, node = _node(G, [ir_t[0].numpy(), ir_t[1].numpy()], async_upload=True)  # the deferred IR preparation happens at create
    x = np.stack([synth.splitmix_uniform(72 + c, 128 * 12) for c in range(2)])
    y = node.ProcessFrames(x)
    yo = np.stack([O.PartitionedConvolver(ir_t[c].numpy(), 128, True).process(x[c]) for c in range(2)])
    assert np.abs(y - yo).max() <= TOL
    with pytest.raises(G.ArgumentException):
        node.ProcessFrames(x[:1])  # wrong channel count
    with pytest.raises(G.ArgumentOutOfRangeException):
        node.ProcessFrames(x[:, :100])  # not a multiple of the quantum
    none = G.CudaConvolverNode(ctx)  # no Buffer: silence with the input's channel count
    assert not none.Process(x[:, :128]).any()
    ctx.Dispose()
