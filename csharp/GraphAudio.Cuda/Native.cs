// Native.cs — P/Invoke surface of libgraphaudio_cuda.so (include/graphaudio_cuda.h), in the idiom of
// GraphAudio.IO/Libsndfile.cs:36-68 and GraphAudio.Realtime/Miniaudio.cs:305-349 ([LibraryImport] + cdecl).
// Source only: this repository's image has no dotnet toolchain, so the file is reviewed by reading.
using System;
using System.Runtime.CompilerServices;
using System.Runtime.InteropServices;

namespace GraphAudio.Cuda;

internal enum GacStatus
{
    Ok = 0, InvalidArgument = -1, OutOfRange = -2, InvalidOperation = -3, Disposed = -4,
    NoDevice = -5, Cuda = -6, OutOfMemory = -7, Nccl = -8, Unsupported = -9
}

[StructLayout(LayoutKind.Sequential)]
internal unsafe struct GacContextDesc
{
    public int SampleRate, Quantum, Partition, DeviceId, MacVariant, TileBlocks, Flags, Reserved;
}

/// <summary>Bit-compatible with AudioParam.AutomationEvent (AudioParam.cs:360-367).</summary>
[StructLayout(LayoutKind.Sequential)]
internal struct GacEvent
{
    public int Type;
    public float Value;
    public float Target;
    public double Time;
    public double TimeConstant;
}

[StructLayout(LayoutKind.Sequential)]
internal unsafe struct GacParam   // gac_param (ABI v6+): 32 bytes
{
    public float Value;
    public int EventCount;
    public GacEvent* Events;
    public int ModBus;            // 0 = no modulation input; k > 0: the fan-in is bus k-1 (GAC_BUS_MONO_INPUT) — AudioNode.Connect(AudioParam)
    public float MinValue, MaxValue;
    public int Reserved;
}

[StructLayout(LayoutKind.Sequential)]
internal unsafe struct GacOpDesc
{
    public int Kind;        // 1 biquad, 2 gain, 3 convolver, 4 delay, 5 stereo panner, 6 splitter channel, 7 connection gate
    public int FilterType;  // (int)FilterType; gate: 0 from / 1 until quantum Aux; convolver: 1 = later epoch of the preceding convolver op
    public GacParam P0, P1, P2;
    public IntPtr Ir;
    public double Aux;      // delay: maxDelayTime (seconds)
}

[StructLayout(LayoutKind.Sequential)]
internal unsafe struct GacVoiceDesc
{
    public IntPtr Source;
    public double StartWhen, StartOffset, StartDuration, StopWhen;
    public float PlaybackRate;
    public int OpCount;
    public GacOpDesc* Ops;
    public int Bus;
    public int Input;   // 0 = fed by Source; k > 0 = fed by the output of bus k-1
    public int Loop;    // AudioBufferSourceNode.Loop (any effective rate)
    public int SourceKind;              // 0 buffer, 1 ConstantSourceNode, 2 OscillatorNode
    public double LoopStart, LoopEnd;   // seconds; LoopEnd 0 = end of the buffer
    public GacParam SourceParam;        // constant: Offset; oscillator: Frequency; buffer source: PlaybackRate when it carries events
    public int OscillatorType;          // 0 sine, 1 square, 2 sawtooth, 3 triangle
    public int Reserved;
}

[StructLayout(LayoutKind.Sequential)]
internal unsafe struct GacBusDesc
{
    public int OpCount;
    public GacOpDesc* Ops;
    public int Target;      // 0 = destination, k > 0 = input of bus k-1, -1 = only read by bus-fed chains
    public int InputCount;  // connection order at the fan-in (entries >= 0: bus indices, < 0: ~voice index); 0 / null = default
    public int* Inputs;
    public int Flags;       // GAC_BUS_MONO_INPUT = 1: the fan-in of an AudioParam (one channel)
    public int Reserved;
    public int* InputSlots; // ChannelMergerNode: per input 0 ordinary, 1 / 2 = merger input 0 / 1; null = all ordinary
}

[StructLayout(LayoutKind.Sequential)]
internal unsafe struct GacGraphDesc
{
    public int VoiceCount;
    public GacVoiceDesc* Voices;
    public int BusCount;
    public GacBusDesc* Buses;
    public int DestInputCount;
    public int* DestInputs;
}

internal static unsafe partial class Native
{
    private const string Lib = "graphaudio_cuda";
    internal const int AbiVersion = 7;                 // GAC_ABI_VERSION this binding was written against (checked once at start-up)
    internal const int FlagAsyncUpload = 1;            // GAC_FLAG_ASYNC_UPLOAD
    internal const int FlagUniformSegments = 2;        // GAC_FLAG_UNIFORM_SEGMENTS
    internal const int FlagNoFaninFusion = 4;          // GAC_FLAG_NO_FANIN_FUSION: one inverse transform and one fan-in input per ConvolverNode

    [LibraryImport(Lib, EntryPoint = "gac_version")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int Version();

    [LibraryImport(Lib, EntryPoint = "gac_last_error")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial IntPtr LastErrorPtr();   // thread-local, owned by the library (cf. sf_strerror)

    [LibraryImport(Lib, EntryPoint = "gac_context_create")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int ContextCreate(GacContextDesc* desc, out IntPtr ctx);

    [LibraryImport(Lib, EntryPoint = "gac_context_destroy")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int ContextDestroy(IntPtr ctx);

    [LibraryImport(Lib, EntryPoint = "gac_buffer_create")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int BufferCreate(IntPtr ctx, float** channels, int channelCount, long frames, int sampleRate, out IntPtr buffer);

    /// <summary>Interleaved file samples (0 s16, 1 s24, 2 s32, 3 f32), converted and de-interleaved on the device — the
    /// upload behind AudioDecoder.LoadFromStream (GraphAudio.IO/LibsndfileDecoder.cs:195-220).</summary>
    [LibraryImport(Lib, EntryPoint = "gac_buffer_create_interleaved")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int BufferCreateInterleaved(IntPtr ctx, void* samples, int sampleFormat, int channelCount, long frames, int sampleRate, out IntPtr buffer);

    [LibraryImport(Lib, EntryPoint = "gac_buffer_destroy")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int BufferDestroy(IntPtr buffer);

    [LibraryImport(Lib, EntryPoint = "gac_ir_prepare")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int IrPrepare(IntPtr ctx, IntPtr buffer, int normalize, int trueStereo, out IntPtr ir);

    [LibraryImport(Lib, EntryPoint = "gac_ir_destroy")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int IrDestroy(IntPtr ir);

    /// <summary>Host-only: how K6 covers a render (transform length, double-length segments in front, segments behind them).</summary>
    [LibraryImport(Lib, EntryPoint = "gac_plan_segments")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int PlanSegments(long blocks, int partitions, int uniform, out int m, out int big, out int small);

    // ---- the per-quantum plugin seam (CudaConvolverNode.cs)
    [LibraryImport(Lib, EntryPoint = "gac_convolver_create")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int ConvolverCreate(IntPtr ctx, IntPtr ir, out IntPtr convolver);

    [LibraryImport(Lib, EntryPoint = "gac_convolver_destroy")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int ConvolverDestroy(IntPtr convolver);

    [LibraryImport(Lib, EntryPoint = "gac_convolver_channels")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int ConvolverChannels(IntPtr convolver, out int inputChannels, out int outputChannels);

    [LibraryImport(Lib, EntryPoint = "gac_convolver_reset")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int ConvolverReset(IntPtr convolver);

    [LibraryImport(Lib, EntryPoint = "gac_convolver_process_block")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int ConvolverProcessBlock(IntPtr convolver, float** input, int inputChannels, float** output, int outputChannels);

    [LibraryImport(Lib, EntryPoint = "gac_convolver_process")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int ConvolverProcess(IntPtr convolver, float** input, int inputChannels, float** output, int outputChannels, long frames);

    [LibraryImport(Lib, EntryPoint = "gac_graph_create")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int GraphCreate(IntPtr ctx, GacGraphDesc* desc, out IntPtr graph);

    [LibraryImport(Lib, EntryPoint = "gac_graph_destroy")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int GraphDestroy(IntPtr graph);

    [LibraryImport(Lib, EntryPoint = "gac_render")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int Render(IntPtr ctx, IntPtr graph, long firstFrame, long frames, float** outChannels, int channelCount, long startIndex);

    /// <summary>The render written interleaved (≙ ProcessBlockInterleaved block after block, AudioContextBase.cs:88-161).</summary>
    [LibraryImport(Lib, EntryPoint = "gac_render_interleaved")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int RenderInterleaved(IntPtr ctx, IntPtr graph, long firstFrame, long frames, float* interleaved, int channels, long startIndex);

    [LibraryImport(Lib, EntryPoint = "gac_render_batch")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int RenderBatch(IntPtr ctx, IntPtr* graphs, int graphCount, long frames, float** outChannels, int channelCount);

    // ---- one process, several GPUs (gac_group: a member context per device, ncclCommInitAll, one ncclReduce of the bus)
    [LibraryImport(Lib, EntryPoint = "gac_group_create")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int GroupCreate(GacContextDesc* desc, int* deviceIds, int deviceCount, out IntPtr group);

    [LibraryImport(Lib, EntryPoint = "gac_group_destroy")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int GroupDestroy(IntPtr group);

    [LibraryImport(Lib, EntryPoint = "gac_group_size")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int GroupSize(IntPtr group, out int members);

    [LibraryImport(Lib, EntryPoint = "gac_group_context")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int GroupContext(IntPtr group, int index, out IntPtr member);

    /// <summary>shards[i] belongs to member i (voices sharded contiguously, the bus described on every member).</summary>
    [LibraryImport(Lib, EntryPoint = "gac_group_render")]
    [UnmanagedCallConv(CallConvs = new[] { typeof(CallConvCdecl) })]
    internal static partial int GroupRender(IntPtr group, IntPtr* shards, long firstFrame, long frames, float** outChannels, int channelCount, long startIndex);

    internal static string LastError() => Marshal.PtrToStringUTF8(LastErrorPtr()) ?? string.Empty;

    /// <summary>Maps gac_status onto the exception types the reference throws at the same places
    /// (OfflineAudioContext.cs:32-51, ConvolverNode.cs:45-49, AudioContextBase.cs:54-55).</summary>
    internal static void Check(int status)
    {
        if (status == 0) return;
        string msg = LastError();
        throw (GacStatus)status switch
        {
            GacStatus.InvalidArgument => new ArgumentException(msg),
            GacStatus.OutOfRange => new ArgumentOutOfRangeException(null, msg),
            GacStatus.InvalidOperation => new InvalidOperationException(msg),
            GacStatus.Disposed => new ObjectDisposedException(nameof(OfflineAudioContext), msg),
            GacStatus.Unsupported => new NotSupportedException(msg),
            GacStatus.OutOfMemory => new OutOfMemoryException(msg),
            _ => new InvalidOperationException($"graphaudio_cuda error {status}: {msg}")
        };
    }
}
