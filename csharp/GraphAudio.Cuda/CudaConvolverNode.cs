// CudaConvolverNode.cs — the literal plugin seam (SURVEY.md §8b): a convolver that sits in an ORDINARY GraphAudio.Core graph
// (realtime or offline) and hands each render quantum to libgraphaudio_cuda.so.  It subclasses the reference's AudioNode
// exactly where GraphAudio.SteamAudio's nodes do (GraphAudio.SteamAudio/Nodes/SteamAudioNodeBase.cs:8-72): Process() reads
// Inputs[0].Buffer, rents its output block from Context.BufferPool, pins both and makes ONE native call, then publishes the
// block with the protected SetOutputBuffer (Nodes/AudioNode.cs:194-200).  Members mirror ConvolverNode
// (Nodes/ConvolverNode.cs:25-95): Buffer, Normalize, EnableTrueStereo.
//
// One call = one 128-frame block = a handful of kernel launches and two small copies: this node is latency-bound and exists
// for mixed graphs; the throughput path is GraphAudio.Cuda.OfflineAudioContext.Render (gac_render).
// Source only: this repository's image has no dotnet toolchain.  The same call sequence is exercised from Python in
// tests/test_gpu_stream_convolver.py (graphaudio_b200.CudaConvolverNode).
using System;
using GraphAudio.Core;
using GraphAudio.Cuda;

namespace GraphAudio.Nodes;

public sealed unsafe class CudaConvolverNode : AudioNode
{
    private IntPtr _ctx;        // a gac_context of its own (device stream, twiddle tables)
    private IntPtr _irBuffer, _ir, _conv;
    private int _inChannels, _outChannels;
    private PlayableAudioBuffer? _buffer;
    private AudioBuffer? _outputBuffer;

    public bool Normalize { get; set; } = true;          // ConvolverNode.cs:87
    public bool EnableTrueStereo { get; set; } = true;   // ConvolverNode.cs:95

    public CudaConvolverNode(AudioContextBase context, int deviceId = -1)
        : base(context, inputCount: 1, outputCount: 1, "CudaConvolver")
    {
        var desc = new GacContextDesc { SampleRate = context.SampleRate, Quantum = 128, Partition = 128, DeviceId = deviceId };
        Native.Check(Native.ContextCreate(&desc, out _ctx));
    }

    public PlayableAudioBuffer? Buffer
    {
        get => _buffer;
        set
        {
            if (_buffer == value) return;
            if (value is null)
            {
                Context.Post(_ => { ReleaseConvolver(); _buffer = null; Inputs[0].SetChannelCountMode(ChannelCountMode.Max); });
                return;
            }
            if (!value.IsInitialized)
                throw new InvalidOperationException("Impulse response buffer must be initialized before being assigned to the ConvolverNode.");
            // (the library checks the sample rate again and reports the reference's message, ConvolverNode.cs:48-49)

            // upload + prepare on the calling thread, swap on the render thread — the split ConvolverNode.Buffer makes (:51-77)
            int channels = value.NumberOfChannels;
            var pins = new System.Runtime.InteropServices.GCHandle[channels];
            float** rows = stackalloc float*[channels];
            IntPtr buf = IntPtr.Zero, ir = IntPtr.Zero, conv = IntPtr.Zero;
            try
            {
                for (int c = 0; c < channels; c++)
                {
                    // GetChannelData hands out a ReadOnlySpan (PlayableAudioBuffer.cs:72): one managed copy, pinned for the upload
                    pins[c] = System.Runtime.InteropServices.GCHandle.Alloc(value.GetChannelData(c).ToArray(), System.Runtime.InteropServices.GCHandleType.Pinned);
                    rows[c] = (float*)pins[c].AddrOfPinnedObject();
                }
                Native.Check(Native.BufferCreate(_ctx, rows, channels, value.Length, value.SampleRate, out buf));
                Native.Check(Native.IrPrepare(_ctx, buf, Normalize ? 1 : 0, EnableTrueStereo ? 1 : 0, out ir));
                Native.Check(Native.ConvolverCreate(_ctx, ir, out conv));
                Native.Check(Native.ConvolverChannels(conv, out int nIn, out int nOut));
                IntPtr b = buf, i = ir, k = conv;
                buf = ir = conv = IntPtr.Zero;   // ownership moves to the posted swap
                Context.Post(_ =>
                {
                    ReleaseConvolver();
                    _irBuffer = b; _ir = i; _conv = k; _inChannels = nIn; _outChannels = nOut; _buffer = value;
                    Inputs[0].SetChannelCount(nIn);                           // ConvolverNode.cs:62-76
                    Inputs[0].SetChannelCountMode(ChannelCountMode.Explicit);
                });
            }
            finally
            {
                foreach (var h in pins) if (h.IsAllocated) h.Free();
                if (conv != IntPtr.Zero) Native.ConvolverDestroy(conv);
                if (ir != IntPtr.Zero) Native.IrDestroy(ir);
                if (buf != IntPtr.Zero) Native.BufferDestroy(buf);
            }
        }
    }

    protected override void Process()
    {
        var input = Inputs[0].Buffer;
        int outCh = _conv == IntPtr.Zero ? input.ChannelCount : _outChannels;
        if (_outputBuffer is null || _outputBuffer.ChannelCount != outCh)
        {
            if (_outputBuffer is not null) Context.BufferPool.Return(_outputBuffer);
            _outputBuffer = Context.BufferPool.Rent(outCh);
        }
        if (_conv == IntPtr.Zero)   // no impulse response: silence with the input's channel count (ConvolverNode.cs:107-119)
        {
            _outputBuffer.Clear();
            SetOutputBuffer(0, _outputBuffer);
            return;
        }
        // a silent input is still convolved: the tail of earlier blocks keeps sounding (ConvolverNode.cs:121-153 has no IsSilent test)
        float** inRows = stackalloc float*[2];
        float** outRows = stackalloc float*[2];
        fixed (float* i0 = input.GetChannelData(0))
        fixed (float* i1 = input.GetChannelData(_inChannels > 1 ? 1 : 0))
        fixed (float* o0 = _outputBuffer.GetChannelData(0))
        fixed (float* o1 = _outputBuffer.GetChannelData(_outChannels > 1 ? 1 : 0))
        {
            inRows[0] = i0; inRows[1] = i1;
            outRows[0] = o0; outRows[1] = o1;
            Native.Check(Native.ConvolverProcessBlock(_conv, inRows, _inChannels, outRows, _outChannels));
        }
        _outputBuffer.MarkAsNonSilent();   // ConvolverNode.cs:153
        SetOutputBuffer(0, _outputBuffer);
    }

    private void ReleaseConvolver()
    {
        if (_conv != IntPtr.Zero) { Native.ConvolverDestroy(_conv); _conv = IntPtr.Zero; }
        if (_ir != IntPtr.Zero) { Native.IrDestroy(_ir); _ir = IntPtr.Zero; }
        if (_irBuffer != IntPtr.Zero) { Native.BufferDestroy(_irBuffer); _irBuffer = IntPtr.Zero; }
    }

    protected override void OnDispose()
    {
        ReleaseConvolver();
        if (_outputBuffer is not null) { Context.BufferPool.Return(_outputBuffer); _outputBuffer = null; }
        if (_ctx != IntPtr.Zero) { Native.ContextDestroy(_ctx); _ctx = IntPtr.Zero; }
        base.OnDispose();
    }
}
