"""graphaudio_b200 — B200-native (sm_100a) batched offline render path of GraphAudio.

The product is the C-ABI shared library graphaudio_b200/lib/libgraphaudio_cuda.so (include/graphaudio_cuda.h);
this package is its host-side mirror of the reference API (api.py) plus the build recipe (build.py).
There is no CPU or pure-Python fallback.
"""
from .api import (  # noqa: F401
    ArgumentException,
    ArgumentOutOfRangeException,
    AudioBufferSourceNode,
    AudioDestinationNode,
    AudioNode,
    AudioParam,
    BiQuadFilterNode,
    ConvolverNode,
    CudaConvolverNode,
    ChannelMergerNode,
    ChannelSplitterNode,
    ConstantSourceNode,
    CudaException,
    DelayNode,
    FilterType,
    GainNode,
    InvalidOperationException,
    NotSupportedException,
    ObjectDisposedException,
    OfflineAudioContext,
    OscillatorNode,
    OscillatorType,
    PlayableAudioBuffer,
    RenderBatch,
    StereoPannerNode,
)

__all__ = [
    "OfflineAudioContext", "PlayableAudioBuffer", "AudioBufferSourceNode", "BiQuadFilterNode", "GainNode", "ConvolverNode",
    "AudioDestinationNode", "AudioNode", "AudioParam", "FilterType", "CudaConvolverNode", "DelayNode", "StereoPannerNode",
]
