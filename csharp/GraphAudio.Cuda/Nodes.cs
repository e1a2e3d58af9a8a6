// Nodes.cs (GraphAudio.Cuda) — the mirror types of the source-level drop-in: same names and members as GraphAudio.Core /
// GraphAudio.Nodes, but they only RECORD topology, automation and buffers; OfflineAudioContext.Render flattens them
// (GraphFlattener.cs) and makes one native call.  Twin of graphaudio_b200/api.py (Python, tested) and
// graphaudio_b200/host/graphaudio_cuda.hpp (C++, tested).  Source only: no dotnet toolchain in this image.
using System;
using System.Collections.Generic;
using System.Runtime.InteropServices;

namespace GraphAudio.Cuda;

public enum FilterType { Lowpass, Highpass, Bandpass, Notch, Allpass, Peaking, Lowshelf, Highshelf }   // BiQuadFilterNode.cs:288-298

/// <summary>PlayableAudioBuffer.cs — planar float32 samples; uploaded to HBM on first use by a context.</summary>
public sealed unsafe class PlayableAudioBuffer : IDisposable
{
    private readonly float[][] _channels;
    private IntPtr _handle;
    public int SampleRate { get; }
    public int NumberOfChannels => _channels.Length;
    public int Length => _channels[0].Length;

    private PlayableAudioBuffer(float[][] channels, int sampleRate) { _channels = channels; SampleRate = sampleRate; }

    public static PlayableAudioBuffer FromChannelArrays(float[][] channelData, int sampleRate)   // PlayableAudioBuffer.cs:122-143
    {
        if (channelData is null || channelData.Length == 0) throw new ArgumentException("Channel data cannot be or empty", nameof(channelData));
        if (channelData.Length > 32) throw new ArgumentOutOfRangeException(nameof(channelData), "Channel count must be between 1 and 32");
        if (sampleRate <= 0) throw new ArgumentOutOfRangeException(nameof(sampleRate), "Sample rate must be positive");
        foreach (var c in channelData)
            if (c.Length != channelData[0].Length) throw new ArgumentException("All channels must have the same length", nameof(channelData));
        var copy = new float[channelData.Length][];
        for (int c = 0; c < copy.Length; c++) copy[c] = (float[])channelData[c].Clone();   // CopyToChannel copies (:84-93)
        return new PlayableAudioBuffer(copy, sampleRate);
    }
    public static PlayableAudioBuffer FromMonoArray(float[] audioData, int sampleRate) => FromChannelArrays(new[] { audioData }, sampleRate);   // :148-157
    public static PlayableAudioBuffer FromStereoArrays(float[] left, float[] right, int sampleRate)   // :162-173
    {
        if (left.Length != right.Length) throw new ArgumentException("Left and right channels must have the same length");
        return FromChannelArrays(new[] { left, right }, sampleRate);
    }

    internal IntPtr Handle(OfflineAudioContext ctx)
    {
        if (_handle != IntPtr.Zero) return _handle;
        var pins = new GCHandle[_channels.Length];
        float** rows = stackalloc float*[_channels.Length];
        try
        {
            for (int c = 0; c < _channels.Length; c++)
            {
                pins[c] = GCHandle.Alloc(_channels[c], GCHandleType.Pinned);
                rows[c] = (float*)pins[c].AddrOfPinnedObject();
            }
            Native.Check(Native.BufferCreate(ctx.Handle, rows, _channels.Length, Length, SampleRate, out _handle));
        }
        finally { foreach (var p in pins) if (p.IsAllocated) p.Free(); }
        return _handle;
    }
    public void Dispose() { if (_handle != IntPtr.Zero) { Native.BufferDestroy(_handle); _handle = IntPtr.Zero; } }
}

/// <summary>AudioParam.cs — value + time-sorted automation events, clamped at schedule time (:254,268,282,299).</summary>
public sealed class AudioParam
{
    private float _value;
    internal readonly List<GacEvent> Events = new();
    // (first quantum, value, events) as each Render call saw the parameter: edits between Render calls act from the next
    // unprocessed quantum on (GAC_EVENT_EPOCH, include/graphaudio_cuda.h)
    internal readonly List<(long Quantum, float Value, GacEvent[] Events)> Epochs = new();
    public float DefaultValue { get; }
    public float MinValue { get; }
    public float MaxValue { get; }
    internal AudioParam(float defaultValue, float min, float max) { DefaultValue = _value = defaultValue; MinValue = min; MaxValue = max; }

    public float Value { get => _value; set { _value = Math.Clamp(value, MinValue, MaxValue); Events.Clear(); } }   // :34-49

    private void Add(GacEvent e)   // AddEvent: stable upper-bound insert (:333-352)
    {
        int lo = 0, hi = Events.Count;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (e.Time < Events[mid].Time) hi = mid; else lo = mid + 1; }
        Events.Insert(lo, e);
    }
    public void SetValueAtTime(float value, double startTime) => Add(new GacEvent { Type = 0, Value = Math.Clamp(value, MinValue, MaxValue), Time = startTime });
    public void LinearRampToValueAtTime(float value, double endTime) => Add(new GacEvent { Type = 1, Value = Math.Clamp(value, MinValue, MaxValue), Time = endTime });
    public void ExponentialRampToValueAtTime(float value, double endTime)
    {
        float v = Math.Clamp(value, MinValue, MaxValue);
        if (v <= 0) throw new ArgumentException("Exponential ramp target must be > 0", nameof(value));   // :283-284
        Add(new GacEvent { Type = 2, Value = v, Time = endTime });
    }
    public void SetTargetAtTime(float target, double startTime, double timeConstant) =>
        Add(new GacEvent { Type = 3, Target = Math.Clamp(target, MinValue, MaxValue), Time = startTime, TimeConstant = timeConstant });
    public void CancelScheduledValues(double cancelTime) => Events.RemoveAll(e => e.Time >= cancelTime);   // :312-331

    /// <summary>The event list that crosses the ABI: epoch 0's events, then for every later epoch a marker followed by its events.</summary>
    internal (float Value, GacEvent[] Flat) Commit(long quantumNow)
    {
        var now = Events.ToArray();
        bool changed = Epochs.Count == 0 || Epochs[^1].Value != _value || !SameEvents(Epochs[^1].Events, now);
        if (Epochs.Count == 0) Epochs.Add((0, _value, now));
        else if (changed && quantumNow > Epochs[^1].Quantum) Epochs.Add((quantumNow, _value, now));
        else if (changed) Epochs[^1] = (Epochs[^1].Quantum, _value, now);
        var flat = new List<GacEvent>();
        for (int k = 0; k < Epochs.Count; k++)
        {
            if (k > 0) flat.Add(new GacEvent { Type = 4, Value = Epochs[k].Value, TimeConstant = Epochs[k].Quantum });
            flat.AddRange(Epochs[k].Events);
        }
        return (Epochs[0].Value, flat.ToArray());
    }
    private static bool SameEvents(GacEvent[] a, GacEvent[] b)
    {
        if (a.Length != b.Length) return false;
        for (int i = 0; i < a.Length; i++)
            if (a[i].Type != b[i].Type || a[i].Value != b[i].Value || a[i].Target != b[i].Target || a[i].Time != b[i].Time || a[i].TimeConstant != b[i].TimeConstant) return false;
        return true;
    }
}

/// <summary>Nodes/AudioNode.cs — records connections; Connect returns the destination to allow chaining (:68-73).</summary>
public abstract class AudioNode
{
    public OfflineAudioContext Context { get; }
    internal readonly List<AudioNode> In = new();    // upstream nodes in connection order (AudioNodeInput._connectedOutputs)
    internal readonly List<AudioNode> Out = new();
    internal readonly long BornFrames;               // frames the context had rendered when the node was created
    protected AudioNode(OfflineAudioContext context) { Context = context; BornFrames = context.FramesRendered; context.Register(this); }

    public AudioNode Connect(AudioNode destination)
    {
        if (ReferenceEquals(destination, this)) throw new InvalidOperationException("Cannot connect a node to itself");   // AudioNodeOutput.cs:43-44
        if (!Out.Contains(destination))
        {
            if (UpstreamExistedInARender()) Context.UnsupportedEdit = "Connect() of a node that already took part in a Render call";
            Out.Add(destination);
            destination.In.Add(this);
        }
        return destination;
    }
    public void Disconnect(AudioNode? destination = null)
    {
        foreach (var d in destination is null ? Out.ToArray() : new[] { destination })
            if (Out.Remove(d))
            {
                if (UpstreamExistedInARender()) Context.UnsupportedEdit = "Disconnect() of a node that already took part in a Render call";
                d.In.Remove(this);
            }
    }
    internal bool ExistedInARender => BornFrames < Context.FramesRendered;
    private bool UpstreamExistedInARender()
    {
        var seen = new HashSet<AudioNode>();
        var stack = new Stack<AudioNode>();
        stack.Push(this);
        while (stack.Count > 0)
        {
            var n = stack.Pop();
            if (!seen.Add(n)) continue;
            if (n.ExistedInARender) return true;
            foreach (var u in n.In) stack.Push(u);
        }
        return false;
    }
}

public sealed class AudioDestinationNode : AudioNode { internal AudioDestinationNode(OfflineAudioContext c) : base(c) { } }

public sealed class AudioBufferSourceNode : AudioNode   // Nodes/AudioBufferSourceNode.cs
{
    public AudioBufferSourceNode(OfflineAudioContext c) : base(c) { }
    public AudioParam PlaybackRate { get; } = new(1.0f, 0.001f, 1000.0f);   // :76 (k-rate)
    public PlayableAudioBuffer? Buffer { get; set; }                          // :67-71
    public bool Loop { get; set; }                                            // :40-44 (any effective rate)
    private double _loopStart, _loopEnd;
    public double LoopStart { get => _loopStart; set => _loopStart = Math.Max(0, value); }   // :49-53
    public double LoopEnd { get => _loopEnd; set => _loopEnd = Math.Max(0, value); }         // :58-62
    internal bool Started;
    internal double When = double.NaN, Offset, Duration = double.PositiveInfinity, StopWhen = double.NaN;
    internal long StartFrames;

    public void Start(double when = 0, double offset = 0, double duration = double.PositiveInfinity)   // :79-114
    {
        if (Started) throw new InvalidOperationException("AudioBufferSourceNode can only be started once.");
        if (Buffer is null) throw new InvalidOperationException("Cannot start without a buffer set");
        Started = true; When = when; Offset = offset; Duration = duration;
        StartFrames = Context.FramesRendered;   // started between Render calls: plays from the next unprocessed quantum
    }
    public void Stop(double when = 0)   // :116-129; cannot silence quanta that were already rendered
    {
        double at = Math.Max(Math.Max(0, when), Context.BlockTime(Context.QuantumNow));
        StopWhen = double.IsNaN(StopWhen) ? at : Math.Min(StopWhen, at);
    }
}

public sealed class BiQuadFilterNode : AudioNode   // Nodes/BiQuadFilterNode.cs:54-85
{
    private FilterType _type = FilterType.Lowpass;
    public BiQuadFilterNode(OfflineAudioContext c) : base(c) { Frequency = new AudioParam(1000.0f, 1.0f, c.SampleRate / 2.0f); }
    public FilterType Type
    {
        get => _type;
        set { if (value != _type && ExistedInARender) Context.UnsupportedEdit = "BiQuadFilterNode.Type changed after the node took part in a Render call"; _type = value; }
    }
    public AudioParam Frequency { get; }                                    // :63-68 (a-rate)
    public AudioParam Q { get; } = new(1.0f, 0.001f, 1000.0f);             // :70-75 (a-rate)
    public AudioParam Gain { get; } = new(0.0f, -60.0f, 60.0f);            // :77-82 (k-rate, dB)
}

public sealed class GainNode : AudioNode   // Nodes/GainNode.cs:16-25
{
    public GainNode(OfflineAudioContext c) : base(c) { }
    public AudioParam Gain { get; } = new(1.0f, float.MinValue, float.MaxValue);
}

public sealed class DelayNode : AudioNode   // Nodes/DelayNode.cs:22-41
{
    public double MaxDelayTime { get; }
    public AudioParam DelayTime { get; }
    public DelayNode(OfflineAudioContext c, double maxDelayTime = 1.0) : base(c)
    {
        if (maxDelayTime <= 0 || maxDelayTime > 10) throw new ArgumentOutOfRangeException(nameof(maxDelayTime));   // :25-26
        MaxDelayTime = maxDelayTime;
        DelayTime = new AudioParam(0.0f, 0.0f, (float)maxDelayTime);
    }
}

public sealed class StereoPannerNode : AudioNode   // Nodes/StereoPannerNode.cs:21-34
{
    public StereoPannerNode(OfflineAudioContext c) : base(c) { }
    public AudioParam Pan { get; } = new(0.0f, -1.0f, 1.0f);
}

public sealed class ConvolverNode : AudioNode   // Nodes/ConvolverNode.cs
{
    private PlayableAudioBuffer? _buffer;
    internal IntPtr Ir;
    public ConvolverNode(OfflineAudioContext c) : base(c) { }
    public bool Normalize { get; set; } = true;          // :87
    public bool EnableTrueStereo { get; set; } = true;   // :95
    public PlayableAudioBuffer? Buffer                    // :25-79: the convolvers are built here, with the Normalize value of this moment
    {
        get => _buffer;
        set
        {
            if (ReferenceEquals(value, _buffer)) return;
            if (ExistedInARender) Context.UnsupportedEdit = "ConvolverNode.Buffer changed after the node took part in a Render call";
            if (Ir != IntPtr.Zero) { Native.IrDestroy(Ir); Ir = IntPtr.Zero; }
            _buffer = value;
            if (value is null) return;
            // (the library checks the sample rate and answers with the reference's message, :48-49)
            Native.Check(Native.IrPrepare(Context.Handle, value.Handle(Context), Normalize ? 1 : 0, EnableTrueStereo ? 1 : 0, out Ir));
        }
    }
}
