// Renders the cross-check cases with the REAL reference (GraphAudio.Core, unmodified) and writes the results next to the
// inputs, so that tests/test_reference_crosscheck.py can pin the CPU oracle (and through it the CUDA path) to the reference
// binary instead of to a reading of its source.
//
//   dotnet run -c Release --project tools/dotnet_crosscheck -- <cases dir> [<out dir>]
//
// A case is a directory written by tools/dotnet_crosscheck/dump_cases.py:
//   case.txt        key = value lines (kind, sample_rate, frames, voices, bus_gain, automation times, ...)
//   v<k>_src<c>.f32 / v<k>_ir<c>.f32   little-endian float32 channel data of voice k
// and the output is <out dir>/ref_<case>.f32: the two rendered channels, planar, little-endian float32, preceded by nothing.
// Every graph is built exactly like tests/synth.py builds it (build_c1 .. build_c5, the latter also with looping, rate-swept sources): the same node order, the same automation calls.
// Timing of each Render call is printed as well: this is the reference's own CPU baseline (one context = one thread).
using System.Diagnostics;
using System.Globalization;
using GraphAudio.Core;
using GraphAudio.Nodes;

static float[] ReadF32(string path)
{
    var bytes = File.ReadAllBytes(path);
    var data = new float[bytes.Length / 4];
    Buffer.BlockCopy(bytes, 0, data, 0, data.Length * 4);
    return data;
}

static Dictionary<string, string> ReadCase(string path)
{
    var d = new Dictionary<string, string>();
    foreach (var line in File.ReadAllLines(path))
    {
        var t = line.Trim();
        if (t.Length == 0 || t.StartsWith('#')) continue;
        int eq = t.IndexOf('=');
        d[t[..eq].Trim()] = t[(eq + 1)..].Trim();
    }
    return d;
}

static double D(Dictionary<string, string> c, string k, double def = 0) =>
    c.TryGetValue(k, out var v) ? double.Parse(v, CultureInfo.InvariantCulture) : def;
static int I(Dictionary<string, string> c, string k, int def = 0) =>
    c.TryGetValue(k, out var v) ? int.Parse(v, CultureInfo.InvariantCulture) : def;

// tests/synth.py: add_gain_automation
static void GainAutomation(AudioParam g, float g0, float g1, float g2, double ts)
{
    g.SetValueAtTime(g0, 0.0);
    g.LinearRampToValueAtTime(g1, 5.0 * ts);
    g.ExponentialRampToValueAtTime(g2, 10.0 * ts);
    g.SetTargetAtTime(0.0f, 10.0 * ts, 0.5 * ts);
}

static float[][] RenderCase(string dir, Dictionary<string, string> c, out double seconds)
{
    string kind = c["kind"];
    int fs = I(c, "sample_rate", 48000);
    int frames = I(c, "frames");
    int voices = I(c, "voices", 1);
    int srcRate = I(c, "source_rate", fs);
    int srcCh = I(c, "source_channels", 2);
    int irCh = I(c, "ir_channels", 2);
    double ts = D(c, "t_scale", 1.0);
    var ctx = new OfflineAudioContext(fs);
    AudioNode sink = ctx.Destination;
    if (kind is "c2" or "c3" or "c5" or "loop")
    {
        var bus = new GainNode(ctx);
        bus.Gain.Value = (float)D(c, "bus_gain", 1.0);
        bus.Connect(ctx.Destination);
        sink = bus;
    }
    for (int v = 0; v < voices; v++)
    {
        var src = new float[srcCh][];
        for (int ch = 0; ch < srcCh; ch++) src[ch] = ReadF32(Path.Combine(dir, $"v{v}_src{ch}.f32"));
        var ir = new float[irCh][];
        for (int ch = 0; ch < irCh; ch++) ir[ch] = ReadF32(Path.Combine(dir, $"v{v}_ir{ch}.f32"));
        var s = new AudioBufferSourceNode(ctx);
        s.Buffer = PlayableAudioBuffer.FromChannelArrays(src, srcRate);
        if (kind == "loop")   // tests/synth.py: build_c5(loop = ..., rate_ramp = ...)
        {
            s.Loop = true;
            s.LoopStart = D(c, "loop_start");
            s.LoopEnd = D(c, "loop_end");
            s.PlaybackRate.SetValueAtTime((float)D(c, "rate0", 1.0), 0.0);
            s.PlaybackRate.LinearRampToValueAtTime((float)D(c, "rate1", 1.0), D(c, "rate_t1", 1.0));
        }
        AudioNode tail = s;
        if (kind is "c3" or "c4")
        {
            var lp = new BiQuadFilterNode(ctx);
            lp.Type = FilterType.Lowpass;
            lp.Q.Value = (float)D(c, "q", 0.707);
            lp.Frequency.SetValueAtTime((float)D(c, "f0", 2000.0), 0.0);
            lp.Frequency.ExponentialRampToValueAtTime((float)D(c, "f1", 12000.0), D(c, "sweep_end", 10.0) * ts);
            tail = tail.Connect(lp);
        }
        if (kind == "c4")
        {
            var hp = new BiQuadFilterNode(ctx);
            hp.Type = FilterType.Highpass;
            hp.Frequency.Value = 200.0f;
            hp.Q.Value = 0.707f;
            tail = tail.Connect(hp);
        }
        if (kind is "c2" or "c3" or "c5" or "loop")
        {
            var gn = new GainNode(ctx);
            GainAutomation(gn.Gain, (float)D(c, $"v{v}_g0"), (float)D(c, $"v{v}_g1"), (float)D(c, $"v{v}_g2"), ts);
            tail = tail.Connect(gn);
        }
        var conv = new ConvolverNode(ctx);
        conv.Normalize = I(c, "normalize", 1) != 0;
        conv.Buffer = PlayableAudioBuffer.FromChannelArrays(ir, fs);
        tail.Connect(conv).Connect(sink);
        s.Start();
    }
    var sw = Stopwatch.StartNew();
    var outp = ctx.Render(frames);
    seconds = sw.Elapsed.TotalSeconds;
    ctx.Dispose();
    return outp;
}

if (args.Length < 1)
{
    Console.Error.WriteLine("usage: Crosscheck <cases dir> [<out dir>]");
    return 2;
}
string casesDir = args[0];
string outDir = args.Length > 1 ? args[1] : casesDir;
Directory.CreateDirectory(outDir);
foreach (var dir in Directory.GetDirectories(casesDir).OrderBy(x => x))
{
    string caseFile = Path.Combine(dir, "case.txt");
    if (!File.Exists(caseFile)) continue;
    var c = ReadCase(caseFile);
    string name = Path.GetFileName(dir);
    var y = RenderCase(dir, c, out double sec);
    using (var f = File.Create(Path.Combine(outDir, $"ref_{name}.f32")))
    {
        for (int ch = 0; ch < 2; ch++)
        {
            var row = ch < y.Length ? y[ch] : new float[y[0].Length];
            var bytes = new byte[row.Length * 4];
            Buffer.BlockCopy(row, 0, bytes, 0, bytes.Length);
            f.Write(bytes);
        }
    }
    int voices = I(c, "voices", 1);
    double renderS = I(c, "frames") / (double)I(c, "sample_rate", 48000);
    Console.WriteLine($"{name}: {voices} voice(s) x {renderS:F3} s rendered in {sec:F3} s = {voices * renderS / sec:F1} voice-s/s (1 thread)");
}
return 0;
