// OfflineAudioContext.cs (GraphAudio.Cuda) — the render entry point of the mirror API.  The node / param / buffer
// mirror types live in Nodes.cs, the graph cutting in GraphFlattener.cs; they record topology and automation exactly like
// graphaudio_b200/api.py and graphaudio_b200/host/graphaudio_cuda.hpp (the Python and C++ twins, which ARE built and tested in
// this repository).
// Source only: no dotnet toolchain in this image.
using System;
using System.Collections.Generic;

namespace GraphAudio.Cuda;

public sealed unsafe class OfflineAudioContext : IDisposable
{
    private IntPtr _ctx;
    private long _framesRendered;
    public int SampleRate { get; }
    public AudioDestinationNode Destination { get; }

    public OfflineAudioContext(int sampleRate = 48000)   // OfflineAudioContext.cs:18
    {
        if (sampleRate <= 0) throw new ArgumentOutOfRangeException(nameof(sampleRate));   // AudioContextBase.cs:37-38
        SampleRate = sampleRate;
        var desc = new GacContextDesc { SampleRate = sampleRate, Quantum = 128, Partition = 128, DeviceId = -1 };
        Native.Check(Native.ContextCreate(&desc, out _ctx));
        Destination = new AudioDestinationNode(this);
    }

    internal IntPtr Handle => _ctx != IntPtr.Zero ? _ctx : throw new ObjectDisposedException(nameof(OfflineAudioContext));
    private readonly List<AudioNode> _nodes = new();
    internal void Register(AudioNode node) => _nodes.Add(node);
    internal long FramesRendered => _framesRendered;
    /// <summary>First quantum the next Render call processes (the reference renders whole 128-frame blocks, :55-100).</summary>
    internal long QuantumNow => (_framesRendered + 127) / 128;
    /// <summary>Start time of quantum q, accumulated like AudioContextBase.cs:78-79.</summary>
    internal double BlockTime(long q) { double t = 0, inc = 128.0 / SampleRate; for (long i = 0; i < q; i++) t += inc; return t; }
    /// <summary>Set by an edit that the re-rendering model cannot reproduce exactly (Nodes.cs); Render then refuses.</summary>
    internal string? UnsupportedEdit;

    /// <summary>Render(float[][] output, int frameCount, int startIndex = 0) — OfflineAudioContext.cs:30.</summary>
    public void Render(float[][] output, int frameCount, int startIndex = 0)
    {
        if (output.Length == 0) throw new ArgumentException("Output buffer must have at least one channel.", nameof(output));
        if (frameCount <= 0) throw new ArgumentOutOfRangeException(nameof(frameCount), "Frame count must be positive.");
        if (startIndex < 0) throw new ArgumentOutOfRangeException(nameof(startIndex), "Start index must be non-negative.");
        for (int ch = 0; ch < output.Length; ch++)
        {
            if (output[ch] is null) throw new ArgumentException($"Channel {ch} buffer is null.", nameof(output));
            if (output[ch].Length < startIndex + frameCount)
                throw new ArgumentException($"Channel {ch} buffer is too small. Required: {startIndex + frameCount}, Available: {output[ch].Length}", nameof(output));
        }
        // GraphFlattener cuts the recorded node graph into what the ABI (v3 and later) knows — voices (chains fed by a source), buses (a
        // node with several inputs plus the chain behind it; `Target` = destination / parent bus) and chains fed by a bus
        // output (a node whose output fans out ends a bus) — with the connection order of every fan-in preserved, and pins
        // the event arrays.  The algorithm is the one of graphaudio_b200/api.py::_topology_full (tested against the oracle on
        // ReverbEffect / AudioBus shaped graphs, tests/test_gpu_graphs.py); Flatten() in graphaudio_cuda.hpp is its C++ twin for
        // the flat source -> chain -> bus -> destination shape.
        using var flat = GraphFlattener.Flatten(this);
        Native.Check(Native.GraphCreate(Handle, flat.Desc, out IntPtr graph));
        try
        {
            // planar float[][] pinned and passed as float**, the way SteamAudioNodeBase.PinBuffersAndProcess does
            // (GraphAudio.SteamAudio/Nodes/SteamAudioNodeBase.cs:74-135)
            var handles = new System.Runtime.InteropServices.GCHandle[output.Length];
            float** rows = stackalloc float*[output.Length];
            try
            {
                for (int ch = 0; ch < output.Length; ch++)
                {
                    handles[ch] = System.Runtime.InteropServices.GCHandle.Alloc(output[ch], System.Runtime.InteropServices.GCHandleType.Pinned);
                    rows[ch] = (float*)handles[ch].AddrOfPinnedObject();
                }
                Native.Check(Native.Render(Handle, graph, _framesRendered, frameCount, rows, output.Length, startIndex));
                _framesRendered += frameCount;   // successive Render calls continue the timeline (:55-100)
            }
            finally
            {
                foreach (var h in handles) if (h.IsAllocated) h.Free();
            }
        }
        finally
        {
            Native.GraphDestroy(graph);
        }
    }

    /// <summary>float[][] Render(int frameCount) — OfflineAudioContext.cs:108.</summary>
    public float[][] Render(int frameCount)
    {
        if (frameCount <= 0) throw new ArgumentOutOfRangeException(nameof(frameCount), "Frame count must be positive.");
        var output = new[] { new float[frameCount], new float[frameCount] };
        Render(output, frameCount);
        return output;
    }

    public void Dispose()
    {
        if (_ctx != IntPtr.Zero) { Native.ContextDestroy(_ctx); _ctx = IntPtr.Zero; }
    }
}
