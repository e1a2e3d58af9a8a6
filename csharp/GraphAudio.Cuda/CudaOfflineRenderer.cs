// CudaOfflineRenderer.cs (GraphAudio.Cuda) — the IMPORTER (SURVEY.md §8b (2), INTEGRATION.md §3): renders an EXISTING reference graph,
// built with the unmodified GraphAudio.Core types, on the GPU.  FromContext walks the public topology of the reference graph
//     AudioContextBase.Destination                      (AudioContextBase.cs:33)
//     AudioNode.Inputs / AudioNode.Outputs              (Nodes/AudioNode.cs:49-50)
//     AudioNodeInput.ConnectedOutputs, connection order (AudioNodeInput.cs:26; the order of the fan-in sum, :118-137)
//     AudioNodeOutput.Owner                             (AudioNodeOutput.cs:18)
// creates the mirror node (Nodes.cs) of every reachable node, copies the public state (parameter values, FilterType, Normalize,
// EnableTrueStereo, Loop*, buffers through GetChannelData) and reads by reflection the three things the reference keeps private:
//     AudioParam._events            (AudioParam.cs:22; struct AutomationEvent :360-367, enum order SetValue / LinearRamp / ExponentialRamp / SetTarget)
//     AudioBufferSourceNode._hasStarted / _startTime / _offset / _duration / _stopTime   (Nodes/AudioBufferSourceNode.cs:15-22)
//     DelayNode._maxDelaySamples    (Nodes/DelayNode.cs:15,28)
// Needed because PartitionedConvolver and AudioNodeOutput.SetBuffer are internal (PartitionedConvolver.cs:12, AudioNodeOutput.cs:37):
// Core cannot be patched from outside, so the graph is re-recorded and rendered by the one native call of OfflineAudioContext.Render.
// Node types outside the accelerated set throw NotSupportedException (never rendered with a node left out).
// Source only: no dotnet toolchain in this image; the Python twin of the recording side is graphaudio_b200/api.py.
//
// Build note: this file needs a reference to GraphAudio.Core (the csproj carries it as an optional ProjectReference, condition
// '$(GraphAudioCoreProject)' != '').  The reference types are addressed through the aliases below, the mirror types by their plain names.
using System;
using System.Collections.Generic;
using System.Reflection;
using RefCore = GraphAudio.Core;
using RefNodes = GraphAudio.Nodes;

namespace GraphAudio.Cuda;

public sealed class CudaOfflineRenderer : IDisposable
{
    private readonly OfflineAudioContext _ctx;
    private readonly Dictionary<RefNodes.AudioNode, AudioNode> _map = new(ReferenceEqualityComparer.Instance);
    private readonly Dictionary<RefCore.PlayableAudioBuffer, PlayableAudioBuffer> _buffers = new(ReferenceEqualityComparer.Instance);

    private CudaOfflineRenderer(int sampleRate) { _ctx = new OfflineAudioContext(sampleRate); }

    /// <summary>Imports the graph that reaches <c>context.Destination</c>.  The reference context is left untouched and can still render on the CPU.</summary>
    public static CudaOfflineRenderer FromContext(RefCore.AudioContextBase context)
    {
        if (context is null) throw new ArgumentNullException(nameof(context));
        var r = new CudaOfflineRenderer(context.SampleRate);
        r.Import(context.Destination);
        return r;
    }

    /// <summary>Render(float[][] output, int frameCount, int startIndex = 0) — same contract as OfflineAudioContext.cs:30.</summary>
    public void Render(float[][] output, int frameCount, int startIndex = 0) => _ctx.Render(output, frameCount, startIndex);
    /// <summary>float[][] Render(int frameCount) — OfflineAudioContext.cs:108.</summary>
    public float[][] Render(int frameCount) => _ctx.Render(frameCount);
    public void Dispose()
    {
        foreach (var b in _buffers.Values) b.Dispose();
        _ctx.Dispose();
    }

    // ---- graph walk: depth first from the destination; every node is mirrored once, connections are replayed in the order of
    // AudioNodeInput.ConnectedOutputs so that every fan-in adds its inputs in the reference's order
    private AudioNode Import(RefNodes.AudioNode node)
    {
        if (_map.TryGetValue(node, out var done)) return done;
        AudioNode mirror = node switch
        {
            RefNodes.AudioDestinationNode => _ctx.Destination,
            RefNodes.AudioBufferSourceNode s => ImportSource(s),
            RefNodes.BiQuadFilterNode b => ImportBiquad(b),
            RefNodes.GainNode g => With(new GainNode(_ctx), m => CopyParam(g.Gain, m.Gain)),
            RefNodes.DelayNode d => ImportDelay(d),
            RefNodes.StereoPannerNode p => With(new StereoPannerNode(_ctx), m => CopyParam(p.Pan, m.Pan)),
            RefNodes.ConvolverNode c => ImportConvolver(c),
            _ => throw new NotSupportedException($"{node.GetType().Name} is outside the accelerated path"),
        };
        _map[node] = mirror;
        foreach (var input in node.Inputs)
        {
            if (input.Index > 0 && input.ConnectedOutputs.Count > 0)
                throw new NotSupportedException($"{node.GetType().Name}: inputs beyond the first are outside the importer (ChannelMergerNode: use the GraphAudio.Cuda mirror types)");
            foreach (var output in input.ConnectedOutputs)   // connection order == mixing order (AudioNodeInput.cs:118-137)
            {
                if (output.Index > 0)
                    throw new NotSupportedException($"{output.Owner.GetType().Name}: outputs beyond the first are outside the importer (ChannelSplitterNode: use the GraphAudio.Cuda mirror types)");
                Import(output.Owner).Connect(mirror);
            }
        }
        return mirror;
    }

    private static T With<T>(T node, Action<T> init) { init(node); return node; }

    private AudioNode ImportSource(RefNodes.AudioBufferSourceNode s)
    {
        var m = new AudioBufferSourceNode(_ctx) { Loop = s.Loop, LoopStart = s.LoopStart, LoopEnd = s.LoopEnd };
        if (s.Buffer is not null) m.Buffer = ImportBuffer(s.Buffer);
        CopyParam(s.PlaybackRate, m.PlaybackRate);
        if (Private<bool>(s, "_hasStarted"))
        {
            // Start(when, offset, duration) as recorded by the reference (:79-114); Stop(when) (:116-129)
            m.Start(Private<double>(s, "_startTime"), Private<double>(s, "_offset"), Private<double>(s, "_duration"));
            double stop = Private<double>(s, "_stopTime");
            if (!double.IsNaN(stop)) m.Stop(stop);
        }
        return m;
    }

    private AudioNode ImportBiquad(RefNodes.BiQuadFilterNode b)
    {
        var m = new BiQuadFilterNode(_ctx) { Type = (FilterType)(int)b.Type };   // same member order (BiQuadFilterNode.cs:288-298)
        CopyParam(b.Frequency, m.Frequency);
        CopyParam(b.Q, m.Q);
        CopyParam(b.Gain, m.Gain);
        return m;
    }

    private AudioNode ImportDelay(RefNodes.DelayNode d)
    {
        // the constructor argument is not kept as such: _maxDelaySamples = (int)(maxDelayTime * sampleRate) (:28); the parameter's
        // MaxValue is the maxDelayTime the node was created with (:30-38)
        var m = new DelayNode(_ctx, d.DelayTime.MaxValue);
        CopyParam(d.DelayTime, m.DelayTime);
        return m;
    }

    private AudioNode ImportConvolver(RefNodes.ConvolverNode c)
    {
        // Normalize / EnableTrueStereo are read when Buffer is set (ConvolverNode.cs:25-79): set them first
        var m = new ConvolverNode(_ctx) { Normalize = c.Normalize, EnableTrueStereo = c.EnableTrueStereo };
        if (c.Buffer is not null) m.Buffer = ImportBuffer(c.Buffer);
        return m;
    }

    private PlayableAudioBuffer ImportBuffer(RefCore.PlayableAudioBuffer b)
    {
        if (_buffers.TryGetValue(b, out var done)) return done;   // one upload per buffer object, however many nodes share it
        var channels = new float[b.NumberOfChannels][];
        for (int ch = 0; ch < channels.Length; ch++) channels[ch] = b.GetChannelData(ch).ToArray();   // PlayableAudioBuffer.cs:72
        return _buffers[b] = PlayableAudioBuffer.FromChannelArrays(channels, b.SampleRate);
    }

    // ---- AudioParam: value + the private event list, replayed through the public scheduling calls of the mirror (which clamp and
    // sort exactly like AudioParam.cs:252-352, so replaying is idempotent)
    private static void CopyParam(RefCore.AudioParam src, AudioParam dst)
    {
        dst.Value = src.Value;
        var events = (Array?)typeof(RefCore.AudioParam).GetField("_events", BindingFlags.NonPublic | BindingFlags.Instance)?.GetValue(src);
        if (events is null) return;
        foreach (var e in events)
        {
            var t = e!.GetType();
            int type = Convert.ToInt32(t.GetField("Type")!.GetValue(e));   // AutomationEventType: SetValue, LinearRamp, ExponentialRamp, SetTarget (:369-375)
            float value = (float)t.GetField("Value")!.GetValue(e)!;
            float target = (float)t.GetField("Target")!.GetValue(e)!;
            double time = (double)t.GetField("Time")!.GetValue(e)!;
            double tc = (double)t.GetField("TimeConstant")!.GetValue(e)!;
            switch (type)
            {
                case 0: dst.SetValueAtTime(value, time); break;
                case 1: dst.LinearRampToValueAtTime(value, time); break;
                case 2: dst.ExponentialRampToValueAtTime(value, time); break;
                case 3: dst.SetTargetAtTime(target, time, tc); break;
                default: throw new NotSupportedException($"AudioParam event type {type}");
            }
        }
        foreach (var input in ParamInputs(src))
            if (input.ConnectedOutputs.Count > 0)
                throw new NotSupportedException("node -> AudioParam connections are outside the importer (use the GraphAudio.Cuda mirror types: AudioNode.Connect(AudioParam))");
    }

    // the modulation input of a parameter (AudioParam.cs:60-62) is private too
    private static IEnumerable<RefCore.AudioNodeInput> ParamInputs(RefCore.AudioParam p)
    {
        foreach (var f in typeof(RefCore.AudioParam).GetFields(BindingFlags.NonPublic | BindingFlags.Instance))
            if (f.FieldType == typeof(RefCore.AudioNodeInput) && f.GetValue(p) is RefCore.AudioNodeInput i) yield return i;
    }

    private static T Private<T>(object o, string field) =>
        (T)(o.GetType().GetField(field, BindingFlags.NonPublic | BindingFlags.Instance)?.GetValue(o)
            ?? throw new MissingFieldException(o.GetType().Name, field));
}
