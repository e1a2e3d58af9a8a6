// GraphFlattener.cs (GraphAudio.Cuda) — cuts the recorded node graph into what the C ABI knows and pins everything the
// descriptors point at for the duration of one native call.  This file spells out the FLAT shape
// (source -> chain -> [fan-in node + chain] -> destination), the twin of Flatten() in graphaudio_b200/host/graphaudio_cuda.hpp;
// bus hierarchies and chains fed by bus outputs follow graphaudio_b200/api.py::_topology_full (ABI v3 fields Target / Inputs / Input).
// Source only: no dotnet toolchain in this image.
using System;
using System.Collections.Generic;
using System.Runtime.InteropServices;

namespace GraphAudio.Cuda;

internal sealed unsafe class GraphFlattener : IDisposable
{
    private readonly List<IntPtr> _blocks = new();   // unmanaged blocks the descriptors point into
    public GacGraphDesc* Desc { get; private set; }

    private T* Alloc<T>(int count) where T : unmanaged
    {
        var p = (T*)NativeMemory.AllocZeroed((nuint)Math.Max(1, count), (nuint)sizeof(T));
        _blocks.Add((IntPtr)p);
        return p;
    }

    private GacParam Param(AudioParam p, long quantumNow)
    {
        var (value, flat) = p.Commit(quantumNow);
        var d = new GacParam { Value = value, EventCount = flat.Length, Events = null };
        if (flat.Length > 0)
        {
            d.Events = Alloc<GacEvent>(flat.Length);
            for (int i = 0; i < flat.Length; i++) d.Events[i] = flat[i];
        }
        return d;
    }

    private GacOpDesc Op(AudioNode n, long q) => n switch
    {
        BiQuadFilterNode b => new GacOpDesc { Kind = 1, FilterType = (int)b.Type, P0 = Param(b.Frequency, q), P1 = Param(b.Q, q), P2 = Param(b.Gain, q) },
        GainNode g => new GacOpDesc { Kind = 2, P0 = Param(g.Gain, q) },
        ConvolverNode c => new GacOpDesc { Kind = 3, Ir = c.Ir },
        DelayNode d => new GacOpDesc { Kind = 4, P0 = Param(d.DelayTime, q), Aux = d.MaxDelayTime },
        StereoPannerNode s => new GacOpDesc { Kind = 5, P0 = Param(s.Pan, q), Aux = (s.BornFrames + 127) / 128 },   // first quantum the node processes
        _ => throw new NotSupportedException($"{n.GetType().Name} is outside the accelerated path")
    };

    private GacOpDesc* Ops(List<AudioNode> chain, long q)
    {
        var ops = Alloc<GacOpDesc>(chain.Count);
        for (int i = 0; i < chain.Count; i++) ops[i] = Op(chain[i], q);
        return ops;
    }

    public static GraphFlattener Flatten(OfflineAudioContext ctx)
    {
        if (ctx.UnsupportedEdit is not null)
            throw new NotSupportedException(ctx.UnsupportedEdit + ": successive Render calls re-render the timeline on the device, which is exact for " +
                                            "parameter edits, sources started or stopped in between and new branches, not for re-wiring what was already rendered");
        var f = new GraphFlattener();
        long q = ctx.QuantumNow;
        var voices = new List<GacVoiceDesc>();
        var buses = new List<GacBusDesc>();
        var dest = new List<int>();

        void AddVoice(AudioBufferSourceNode s, List<AudioNode> chain, int bus)
        {
            double when = (s.Started && s.Buffer is not null) ? s.When : double.NaN;
            if (!double.IsNaN(when) && s.StartFrames > 0) when = Math.Max(when, ctx.BlockTime((s.StartFrames + 127) / 128));
            voices.Add(new GacVoiceDesc
            {
                Source = s.Buffer?.Handle(ctx) ?? IntPtr.Zero,
                StartWhen = when, StartOffset = s.Offset, StartDuration = s.Duration, StopWhen = s.StopWhen,
                PlaybackRate = s.PlaybackRate.Value, OpCount = chain.Count, Ops = f.Ops(chain, q), Bus = bus, Input = 0,
                Loop = s.Loop ? 1 : 0, LoopStart = s.LoopStart, LoopEnd = s.LoopEnd,
                SourceParam = f.Param(s.PlaybackRate, q)   // read by the library when it carries events / epochs (k-rate, per quantum on the host)
            });
        }
        // node .. upstream through single-input nodes; returns the chain in processing order and the node it starts from
        static AudioNode Walk(AudioNode head, List<AudioNode> chain)
        {
            var node = head;
            while (node is not AudioBufferSourceNode && node.In.Count == 1)
            {
                if (node.Out.Count > 1) throw new NotSupportedException("fan-out inside a chain needs the bus-fed-chain form (api.py::_topology_full)");
                chain.Add(node);
                node = node.In[0];
            }
            return node;
        }

        foreach (var head in ctx.Destination.In)   // connection order of the destination's fan-in
        {
            var chain = new List<AudioNode>();
            var node = Walk(head, chain);
            if (node is AudioBufferSourceNode src)
            {
                chain.Reverse();
                AddVoice(src, chain, -1);
                dest.Add(~(voices.Count - 1));
                continue;
            }
            if (node.In.Count == 0) continue;   // nothing connected: contributes silence
            chain.Add(node);                     // the fan-in node (typically the bus GainNode) heads the bus chain
            chain.Reverse();
            int bus = buses.Count;
            var inputs = new List<int>();
            foreach (var up in node.In)          // connection order at the fan-in (AudioNodeInput.cs:118-137)
            {
                var vchain = new List<AudioNode>();
                var start = Walk(up, vchain);
                if (start is not AudioBufferSourceNode vs) throw new NotSupportedException("nested fan-in needs the bus-hierarchy form (api.py::_topology_full)");
                vchain.Reverse();
                AddVoice(vs, vchain, bus);
                inputs.Add(~(voices.Count - 1));
            }
            var inp = f.Alloc<int>(inputs.Count);
            for (int i = 0; i < inputs.Count; i++) inp[i] = inputs[i];
            buses.Add(new GacBusDesc { OpCount = chain.Count, Ops = f.Ops(chain, q), Target = 0, InputCount = inputs.Count, Inputs = inp });
            dest.Add(bus);
        }

        var v = f.Alloc<GacVoiceDesc>(voices.Count);
        for (int i = 0; i < voices.Count; i++) v[i] = voices[i];
        var b = f.Alloc<GacBusDesc>(buses.Count);
        for (int i = 0; i < buses.Count; i++) b[i] = buses[i];
        var d = f.Alloc<int>(dest.Count);
        for (int i = 0; i < dest.Count; i++) d[i] = dest[i];
        var g = f.Alloc<GacGraphDesc>(1);
        *g = new GacGraphDesc { VoiceCount = voices.Count, Voices = v, BusCount = buses.Count, Buses = b, DestInputCount = dest.Count, DestInputs = d };
        f.Desc = g;
        return f;
    }

    public void Dispose()
    {
        foreach (var p in _blocks) NativeMemory.Free((void*)p);
        _blocks.Clear();
    }
}
