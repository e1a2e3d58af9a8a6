// ga_oracle.cpp — CPU restatement of GraphAudio's offline render hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under graphaudio_b200/ may include, link or call this
// file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs use it, and only as the checker / CPU baseline.
//
// PARITY UNPINNED: the reference (C#/.NET 9) ships no tests, golden vectors or fixtures
//  and cannot be executed in this image (no dotnet/mono).  This file is a
// from-source restatement of the algorithms, every function citing the reference lines it
// follows; it is cross-checked against independent math (numpy/scipy) in tests/.
//
// Build: g++ -O2 -std=c++17 -mavx2 -ffp-contract=off -fno-fast-math -shared -fPIC
// (-ffp-contract=off keeps every float op separately rounded, as the reference's
// unfused AVX Multiply/Subtract/Add do: PartitionedConvolver.cs:195-204.)
//
// All paths below are relative to /root/reference/GraphAudio.Core/.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <functional>
#include <limits>
#include <memory>
#include <vector>

namespace ora {

static constexpr int kQuantum = 128;  // AudioBuffer.cs:10 FramesPerBlock

// ---------------------------------------------------------------------------------------------
// Real FFT, double precision.  FftFlat/RealFourierTransform.cs:62-131 fixes only the
// CONVENTION (Forward == numpy.fft.rfft, Inverse == numpy.fft.irfft, i.e. exact inverse with
// the 2/N scaling of :46,129); Ooura's fftsg internals are not transliterated — any
// double-precision transform with ~1e-16 relative error gives the same float32 spectra
// after the casts in PartitionedConvolver.cs:87-88,117-118 (up to rare 1-ulp ties).
// Implementation here: N/2-point complex radix-2 FFT of the even/odd packing + split step.
// ---------------------------------------------------------------------------------------------
class RealFFT {
 public:
  explicit RealFFT(int n) : n_(n), h_(n / 2) {
    // RealFourierTransform.cs:28-41 argument checks (power of two, even, >= 2)
    valid_ = n >= 2 && (n & (n - 1)) == 0;
    if (!valid_) return;
    int lg = 0;
    while ((1 << lg) < h_) lg++;
    rev_.resize(h_);
    for (int i = 0; i < h_; i++) {
      int r = 0;
      for (int b = 0; b < lg; b++)
        if (i & (1 << b)) r |= 1 << (lg - 1 - b);
      rev_[i] = r;
    }
    const double pi = 3.14159265358979323846;
    twc_.resize(h_ / 2 + 1);
    tws_.resize(h_ / 2 + 1);
    for (int k = 0; k <= h_ / 2; k++) {  // e^{-2 pi i k / h}
      twc_[k] = std::cos(2.0 * pi * k / h_);
      tws_[k] = -std::sin(2.0 * pi * k / h_);
    }
    splc_.resize(h_ + 1);
    spls_.resize(h_ + 1);
    for (int k = 0; k <= h_; k++) {  // e^{-2 pi i k / n}
      splc_[k] = std::cos(2.0 * pi * k / n_);
      spls_[k] = -std::sin(2.0 * pi * k / n_);
    }
    zr_.resize(h_);
    zi_.resize(h_);
  }
  bool valid() const { return valid_; }
  int size() const { return n_; }

  // x[n] real -> re/im[n/2+1]
  void forward(const double* x, double* re, double* im) {
    for (int i = 0; i < h_; i++) {
      zr_[rev_[i]] = x[2 * i];
      zi_[rev_[i]] = x[2 * i + 1];
    }
    cfft(false);
    // split: X[k] = E[k] + e^{-2 pi i k/n} O[k]
    for (int k = 0; k <= h_; k++) {
      int a = k % h_, b = (h_ - k) % h_;
      double er = 0.5 * (zr_[a] + zr_[b]), ei = 0.5 * (zi_[a] - zi_[b]);
      double orr = 0.5 * (zi_[a] + zi_[b]), oi = -0.5 * (zr_[a] - zr_[b]);
      re[k] = er + (splc_[k] * orr - spls_[k] * oi);
      im[k] = ei + (splc_[k] * oi + spls_[k] * orr);
    }
    im[0] = 0.0;   // RealFourierTransform.cs:76-78: DC and Nyquist are purely real
    im[h_] = 0.0;
  }

  // re/im[n/2+1] -> x[n] real, exact inverse (includes the 1/n)
  void inverse(const double* re, const double* im, double* x) {
    for (int k = 0; k < h_; k++) {
      // E[k] = (X[k] + conj X[h-k]) / 2 ; O[k] = (X[k] - conj X[h-k]) / 2 * e^{+2 pi i k/n}
      double ar = re[k], ai = (k == 0 ? 0.0 : im[k]);
      double br = re[h_ - k], bi = (k == 0 ? 0.0 : -im[h_ - k]);  // conj X[h-k]; im[h] ignored
      double er = 0.5 * (ar + br), ei = 0.5 * (ai + bi);
      double dr = 0.5 * (ar - br), di = 0.5 * (ai - bi);
      double c = splc_[k], s = -spls_[k];  // e^{+i theta}
      double orr = dr * c - di * s, oi = dr * s + di * c;
      // z[k] = E[k] + i O[k]
      zr_[rev_[k]] = er - oi;
      zi_[rev_[k]] = ei + orr;
    }
    cfft(true);
    const double sc = 1.0 / h_;
    for (int i = 0; i < h_; i++) {
      x[2 * i] = zr_[i] * sc;
      x[2 * i + 1] = zi_[i] * sc;
    }
  }

 private:
  // in-place radix-2 DIT on bit-reversed input, length h_
  void cfft(bool inv) {
    for (int len = 2; len <= h_; len <<= 1) {
      int half = len >> 1, step = h_ / len;
      for (int s = 0; s < h_; s += len) {
        for (int j = 0; j < half; j++) {
          int ti = j * step;  // 0 .. h/2-1
          double wr = twc_[ti], wi = inv ? -tws_[ti] : tws_[ti];
          int a = s + j, b = a + half;
          double tr = zr_[b] * wr - zi_[b] * wi, tim = zr_[b] * wi + zi_[b] * wr;
          zr_[b] = zr_[a] - tr;
          zi_[b] = zi_[a] - tim;
          zr_[a] += tr;
          zi_[a] += tim;
        }
      }
    }
  }
  int n_, h_;
  bool valid_ = false;
  std::vector<int> rev_;
  std::vector<double> twc_, tws_, splc_, spls_, zr_, zi_;
};

// ---------------------------------------------------------------------------------------------
// PartitionedConvolver.cs (whole file)
// ---------------------------------------------------------------------------------------------
class PartitionedConvolver {
 public:
  // ctor :37-63
  PartitionedConvolver(const float* ir, int64_t irLen, int blockSize, bool normalize)
      : B_(blockSize), N_(2 * blockSize), C_(blockSize + 1), fft_(2 * blockSize) {
    P_ = (int)std::ceil((double)irLen / blockSize);  // :44
    size_t total = (size_t)P_ * C_;
    irRe_.assign(total, 0.f);
    irIm_.assign(total, 0.f);
    dlRe_.assign(total, 0.f);
    dlIm_.assign(total, 0.f);
    overlap_.assign(B_, 0.f);
    accRe_.assign(C_, 0.f);
    accIm_.assign(C_, 0.f);
    tin_.assign(N_, 0.0);
    tre_.assign(C_, 0.0);
    tim_.assign(C_, 0.0);
    prepare(ir, irLen, normalize);
  }

  // CalculateNormalizationScale :93-102
  static float normalizationScale(const float* r, int64_t n) {
    const float GainCalibration = -58;
    const float MinPower = 0.000125f;
    double sumSquared = 0;
    for (int64_t i = 0; i < n; i++) sumSquared += (double)(float)(r[i] * r[i]);  // float*float, then widened (:98)
    float power = (float)std::sqrt(sumSquared / (double)n);
    if (std::isnan(power) || std::isinf(power) || power < MinPower) power = MinPower;
    return (1.0f / power) * (float)std::pow(10.0, (double)(GainCalibration * 0.05f));
  }

  // Process :104-152
  void process(const float* in, float* out) {
    for (int i = 0; i < B_; i++) tin_[i] = in[i];          // :106
    for (int i = B_; i < N_; i++) tin_[i] = 0.0;           // :107
    fft_.forward(tin_.data(), tre_.data(), tim_.data());  // :109
    size_t off = (size_t)w_ * C_;
    for (int i = 0; i < C_; i++) {                         // :115-124
      dlRe_[off + i] = (float)tre_[i];
      dlIm_[off + i] = (float)tim_[i];
    }
    spectralConvolution();                                 // :125
    w_--;                                                  // :127-128
    if (w_ < 0) w_ = P_ - 1;
    for (int i = 0; i < C_; i++) {                         // :134-137
      tre_[i] = accRe_[i];
      tim_[i] = accIm_[i];
    }
    fft_.inverse(tre_.data(), tim_.data(), tin_.data());  // :140
    for (int i = 0; i < B_; i++) {                         // :146-150
      out[i] = (float)tin_[i] + overlap_[i];
      overlap_[i] = (float)tin_[i + B_];
    }
  }

  int partitions() const { return P_; }
  int bins() const { return C_; }
  const float* irRe() const { return irRe_.data(); }
  const float* irIm() const { return irIm_.data(); }

 private:
  // PrepareImpulseResponse :65-91
  void prepare(const float* ir, int64_t irLen, bool normalize) {
    float scale = 1.0f;
    if (normalize) scale = normalizationScale(ir, irLen);
    std::vector<double> t(N_), re(C_), im(C_);
    for (int p = 0; p < P_; p++) {
      std::fill(t.begin(), t.end(), 0.0);
      int64_t offset = (int64_t)p * B_;
      int len = (int)std::min<int64_t>(B_, irLen - offset);
      for (int i = 0; i < len; i++) t[i] = (double)(float)(ir[offset + i] * scale);  // :80
      fft_.forward(t.data(), re.data(), im.data());
      size_t po = (size_t)p * C_;
      for (int i = 0; i < C_; i++) {
        irRe_[po + i] = (float)re[i];
        irIm_[po + i] = (float)im[i];
      }
    }
  }

  // ProcessSpectralConvolution :154-223 — p ascending, separate mul/sub/add per term
  void spectralConvolution() {
    std::fill(accRe_.begin(), accRe_.end(), 0.f);
    std::fill(accIm_.begin(), accIm_.end(), 0.f);
    const int count = C_;
    float* __restrict ar = accRe_.data();
    float* __restrict ai = accIm_.data();
    for (int p = 0; p < P_; p++) {
      int dp = w_ + p;  // :173-174
      if (dp >= P_) dp -= P_;
      const float* __restrict dr = dlRe_.data() + (size_t)dp * count;
      const float* __restrict di = dlIm_.data() + (size_t)dp * count;
      const float* __restrict hr = irRe_.data() + (size_t)p * count;
      const float* __restrict hi = irIm_.data() + (size_t)p * count;
      for (int i = 0; i < count; i++) {  // :195-204 and :213-219 round identically
        float ac = dr[i] * hr[i];
        float bd = di[i] * hi[i];
        float re = ac - bd;
        float ad = dr[i] * hi[i];
        float bc = di[i] * hr[i];
        float im = ad + bc;
        ar[i] = ar[i] + re;
        ai[i] = ai[i] + im;
      }
    }
  }

  int B_, N_, C_, P_ = 0, w_ = 0;
  RealFFT fft_;
  std::vector<float> irRe_, irIm_, dlRe_, dlIm_, overlap_, accRe_, accIm_;
  std::vector<double> tin_, tre_, tim_;
};

// ---------------------------------------------------------------------------------------------
// CubicResampler.cs:19-97
// ---------------------------------------------------------------------------------------------
struct CubicResampler {
  float S0 = 0, S1 = 0, S2 = 0, S3 = 0;
  double Pos = 0;
  int Ready = 0;
  void clear() { S0 = S1 = S2 = S3 = 0; Pos = 0; Ready = 0; }  // :66-71
  void shift(float s) { S0 = S1; S1 = S2; S2 = S3; S3 = s; }    // :91-97
  // Process :26-63.  Returns (consumed, produced).
  void process(const float* in, int inLen, float* out, int outLen, double rate, int* consumed, int* produced) {
    int inPos = 0, outPos = 0;
    while (Ready < 4 && inPos < inLen) { shift(in[inPos++]); Ready++; }
    if (Ready < 4) { *consumed = inPos; *produced = outPos; return; }
    while (outPos < outLen) {
      int consume = (int)Pos;
      if (inPos + consume > inLen) break;
      for (int i = 0; i < consume; i++) shift(in[inPos++]);
      Pos -= consume;
      float t = (float)Pos;
      out[outPos++] = S1 + t * (0.5f * (S2 - S0) + t * ((S0 - 2.5f * S1 + 2.f * S2 - 0.5f * S3) + t * (0.5f * (S3 - S0) + 1.5f * (S1 - S2))));
      Pos += rate;
    }
    *consumed = inPos;
    *produced = outPos;
  }
};

// ---------------------------------------------------------------------------------------------
// AudioBuffer.cs:8-181 (block with the IsSilent flag)
// ---------------------------------------------------------------------------------------------
struct Block {
  int channels;
  bool silent = true;
  std::vector<float> data;  // [channels][128]
  explicit Block(int ch) : channels(ch), data((size_t)ch * kQuantum, 0.f) {}
  float* ch(int c) { return data.data() + (size_t)c * kQuantum; }
  void clear() { std::fill(data.begin(), data.end(), 0.f); silent = true; }  // :60-67
  void markNonSilent() { silent = false; }                                   // :73-76
  void copyFrom(Block& s) {                                                  // :81-104
    if (s.silent) { clear(); return; }
    int m = std::min(channels, s.channels);
    for (int c = 0; c < m; c++) std::memcpy(ch(c), s.ch(c), sizeof(float) * kQuantum);
    for (int c = m; c < channels; c++) std::fill(ch(c), ch(c) + kQuantum, 0.f);
    silent = false;
  }
};
using BlockPtr = std::shared_ptr<Block>;
static BlockPtr rent(int ch) { return std::make_shared<Block>(ch); }  // BufferPool.Rent returns a cleared block (BufferPool.cs:66-85)

struct PlayBuffer {  // PlayableAudioBuffer.cs
  int channels = 0, rate = 0;
  int64_t length = 0;
  std::vector<std::vector<float>> data;
};

class Context;
class Node;

// AudioParam.cs:360-375
enum EvType { SetValue = 0, LinearRamp = 1, ExponentialRamp = 2, SetTarget = 3 };
struct AutomationEvent {
  int type;
  float value;
  float target;
  double time;
  double timeConstant;
};

enum ChannelCountMode { ModeMax = 0, ModeClampedMax = 1, ModeExplicit = 2 };

struct Output {
  Node* owner;
  BlockPtr buffer;                // AudioNodeOutput.cs:14
  std::vector<struct Input*> to;  // connected inputs
};

// AudioNodeInput.cs
struct Input {
  Node* owner;
  std::vector<Output*> from;  // connection order (:16)
  BlockPtr buffer;
  bool dirty = true;
  int channelCount = 2;
  int mode = ModeMax;
  void pull(int blockNumber, double blockTime);
  int computeOutputChannelCount();
  void ensureBuffer() {  // :170-180
    if (!buffer || dirty) { buffer = rent(channelCount); dirty = false; }
  }
};

// AudioParam.cs
class Param {
 public:
  Param(Node* owner, float def, float mn, float mx, bool arate) : owner_(owner), min_(mn), max_(mx), arate_(arate), value_(def) {}
  static float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
  void setValue(float v) { value_ = clampf(v, min_, max_); events_.clear(); }  // :34-49
  void addEvent(AutomationEvent e) {                                           // :333-352 (upper bound: equal times keep call order)
    size_t lo = 0, hi = events_.size();
    while (lo < hi) {
      size_t mid = (lo + hi) >> 1;
      if (e.time < events_[mid].time) hi = mid; else lo = mid + 1;
    }
    events_.insert(events_.begin() + lo, e);
  }
  void setValueAtTime(float v, double t) { addEvent({SetValue, clampf(v, min_, max_), 0.f, t, 0.0}); }          // :252-261
  void linearRamp(float v, double t) { addEvent({LinearRamp, clampf(v, min_, max_), 0.f, t, 0.0}); }             // :266-275
  bool exponentialRamp(float v, double t) {                                                                     // :280-292
    v = clampf(v, min_, max_);
    if (v <= 0.f) return false;  // ArgumentException
    addEvent({ExponentialRamp, v, 0.f, t, 0.0});
    return true;
  }
  void setTarget(float tgt, double t, double tc) { addEvent({SetTarget, 0.f, clampf(tgt, min_, max_), t, tc}); }  // :297-307
  void cancel(double t) {                                                                                       // :312-331
    size_t s = 0;
    while (s < events_.size() && events_[s].time < t) s++;
    events_.resize(s);
  }
  // ComputeValues :93-111.  `mod` is the param's own AudioNodeInput (Explicit, 1 channel, :68-70): nodes connected to the param are
  // pulled and mixed down to mono, and the sum is added to the intrinsic value, clamped to the param's range (:125-131, :150-156).
  void computeValues(int blockNumber, double blockTime, int sampleRate) {
    const bool hasModulation = mod && !mod->from.empty();
    if (hasModulation) mod->pull(blockNumber, blockTime);
    const Block* mb = hasModulation ? mod->buffer.get() : nullptr;
    const bool useMod = mb && !mb->silent;
    if (arate_) {  // ComputeARate :114-141
      double deltaTime = 1.0 / sampleRate;
      for (int i = 0; i < kQuantum; i++) {
        float v = valueAtTime(blockTime + i * deltaTime);
        values[i] = useMod ? clampf(v + mb->data[i], min_, max_) : v;
      }
    } else {  // ComputeKRate :144-166
      float v = valueAtTime(blockTime);
      if (useMod) v = clampf(v + mb->data[0], min_, max_);
      for (int i = 0; i < kQuantum; i++) values[i] = v;
    }
  }
  std::unique_ptr<Input> mod;  // created by Node::addParam (needs the owner)
  // ComputeValueAtTime :169-217
  float valueAtTime(double time) const {
    size_t count = events_.size();
    if (count == 0) return value_;
    float boundary = value_;
    for (size_t i = 0; i < count; i++) {
      const AutomationEvent& e = events_[i];
      if (time < e.time) {
        if (i == 0) return boundary;
        const AutomationEvent& prev = events_[i - 1];
        if (e.type == LinearRamp) return lerp(prev.value, prev.time, e.value, e.time, time);
        if (e.type == ExponentialRamp) return eerp(prev.value, prev.time, e.value, e.time, time);
        if (prev.type == SetTarget) return target(prev, boundary, time);
        return prev.value;
      }
      if (e.type != SetTarget) boundary = e.value;
    }
    const AutomationEvent& last = events_[count - 1];
    if (last.type == SetTarget) return target(last, boundary, time);
    return last.value;
  }
  static float lerp(float v0, double t0, float v1, double t1, double t) {  // :220-225
    double u = (t - t0) / (t1 - t0);
    u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
    float d = v1 - v0;
    return (float)((double)v0 + (double)d * u);
  }
  static float eerp(float v0, double t0, float v1, double t1, double t) {  // :228-237
    if (v0 <= 0 || v1 <= 0) return lerp(v0, t0, v1, t1, t);
    double u = (t - t0) / (t1 - t0);
    u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
    float ratio = v1 / v0;
    return (float)((double)v0 * std::pow((double)ratio, u));
  }
  static float target(const AutomationEvent& e, float base, double time) {  // :240-247
    double elapsed = time - e.time;
    if (elapsed <= 0) return base;
    double tc = std::max(e.timeConstant, 0.001);
    float d = base - e.target;
    return (float)((double)e.target + (double)d * std::exp(-elapsed / tc));
  }
  float values[kQuantum];
  const std::vector<AutomationEvent>& events() const { return events_; }
  float staticValue() const { return value_; }

 private:
  Node* owner_;
  float min_, max_;
  bool arate_;
  float value_;
  std::vector<AutomationEvent> events_;
};

// Nodes/AudioNode.cs
class Node {
 public:
  Node(Context* c, int nin, int nout) : ctx(c) {
    for (int i = 0; i < nin; i++) { inputs.emplace_back(new Input()); inputs.back()->owner = this; }
    for (int i = 0; i < nout; i++) { outputs.emplace_back(new Output()); outputs.back()->owner = this; }
  }
  virtual ~Node() {}
  Param* addParam(float def, float mn, float mx, bool arate) {
    params.emplace_back(new Param(this, def, mn, mx, arate));
    Param* p = params.back().get();
    p->mod.reset(new Input());  // AudioParam.cs:68-70: SetChannelCount(1), Explicit
    p->mod->owner = this;
    p->mod->channelCount = 1;
    p->mod->mode = ModeExplicit;
    return p;
  }
  // ProcessInternal :152-183
  bool processInternal(int blockNumber, double blockTime);
  virtual void process() = 0;
  void disposeNow();  // Dispose :203-238 body
  Context* ctx;
  std::vector<std::unique_ptr<Input>> inputs;
  std::vector<std::unique_ptr<Output>> outputs;
  std::vector<std::unique_ptr<Param>> params;
  int lastBlock = 0;
  bool processing = false;
  bool disposed = false;
};

class Destination;

// AudioContextBase.cs + OfflineAudioContext.cs
class Context {
 public:
  explicit Context(int fs);
  ~Context();
  int sampleRate;
  int currentBlock = 0;       // AudioContextBase.cs:16
  double currentTime = 0.0;   // :17
  bool cycle = false;
  bool unsupported = false;  // a path this oracle does not restate was reached (a looping source with an empty loop region: the reference never returns)
  std::deque<std::function<void()>> commands;  // :15
  std::vector<std::unique_ptr<Node>> nodes;
  std::vector<std::unique_ptr<PlayBuffer>> buffers;
  Destination* destination = nullptr;
  // OfflineAudioContext cache of the unread tail of the last quantum (:10-13)
  std::vector<std::vector<float>> cache;
  int cached = 0;
  Block* processBlock();
  int render(float* const* out, int channels, int frameCount, int startIndex);
};

// Nodes/AudioDestinationNode.cs:42-64
class Destination : public Node {
 public:
  explicit Destination(Context* c) : Node(c, 1, 0) { inputs[0]->channelCount = 2; inputs[0]->dirty = true; }
  void process() override {
    if (inputs[0]->buffer) out = inputs[0]->buffer;
    else { out = rent(inputs[0]->channelCount); out->clear(); }
  }
  BlockPtr out;
};

int Input::computeOutputChannelCount() {  // AudioNodeInput.cs:140-168
  switch (mode) {
    case ModeExplicit: return channelCount;
    case ModeClampedMax: {
      int m = 0;
      for (auto* o : from) if (o->buffer) m = std::max(m, o->buffer->channels);
      return std::min(m == 0 ? channelCount : m, channelCount);
    }
    default: {
      int m = channelCount;
      for (auto* o : from) if (o->buffer) m = std::max(m, o->buffer->channels);
      return m;
    }
  }
}

// MixBuffer AudioNodeInput.cs:182-244
static void mixBuffer(Block& s, Block& d) {
  int sc = s.channels, dc = d.channels;
  if (sc == dc) {
    for (int c = 0; c < sc; c++) { float* a = s.ch(c); float* b = d.ch(c); for (int i = 0; i < kQuantum; i++) b[i] += a[i]; }
  } else if (sc == 1 && dc > 1) {
    float* a = s.ch(0);
    for (int c = 0; c < dc; c++) { float* b = d.ch(c); for (int i = 0; i < kQuantum; i++) b[i] += a[i]; }
  } else if (sc > 1 && dc == 1) {
    float* b = d.ch(0);
    float scale = 1.0f / sqrtf((float)sc);
    for (int i = 0; i < kQuantum; i++) {
      float sum = 0;
      for (int c = 0; c < sc; c++) sum += s.ch(c)[i];
      b[i] += sum * scale;
    }
  } else {
    int m = std::min(sc, dc);
    for (int c = 0; c < m; c++) { float* a = s.ch(c); float* b = d.ch(c); for (int i = 0; i < kQuantum; i++) b[i] += a[i]; }
  }
}

void Input::pull(int blockNumber, double blockTime) {  // AudioNodeInput.cs:100-138
  if (from.empty()) { ensureBuffer(); buffer->clear(); return; }
  int oc = computeOutputChannelCount();  // uses upstream buffers of the PREVIOUS block (:109 precedes :124)
  ensureBuffer();
  if (buffer->channels != oc) buffer = rent(oc);
  buffer->clear();
  bool mixed = false;
  for (size_t i = 0; i < from.size(); i++) {
    Output* o = from[i];
    o->owner->processInternal(blockNumber, blockTime);  // AudioNodeOutput.cs:75
    if (o->buffer && !o->buffer->silent) { mixBuffer(*o->buffer, *buffer); mixed = true; }
  }
  if (mixed) buffer->markNonSilent();
}

bool Node::processInternal(int blockNumber, double blockTime) {
  if (lastBlock == blockNumber) return true;
  if (processing) { ctx->cycle = true; return false; }  // InvalidOperationException :157-160
  processing = true;
  lastBlock = blockNumber;
  for (auto& p : params) p->computeValues(blockNumber, blockTime, ctx->sampleRate);
  for (auto& in : inputs) in->pull(blockNumber, blockTime);
  process();
  processing = false;
  return true;
}

static void disconnect(Output* o, Input* in) {
  auto it = std::find(o->to.begin(), o->to.end(), in);
  if (it == o->to.end()) return;
  o->to.erase(it);
  auto jt = std::find(in->from.begin(), in->from.end(), o);
  if (jt != in->from.end()) in->from.erase(jt);
  in->dirty = true;  // RemoveConnection AudioNodeInput.cs:69-73
}
static void connect(Output* o, Input* in) {  // AudioNodeOutput.ConnectTo :41-51, AddConnection AudioNodeInput.cs:60-67
  if (std::find(o->to.begin(), o->to.end(), in) != o->to.end()) return;
  o->to.push_back(in);
  if (std::find(in->from.begin(), in->from.end(), o) == in->from.end()) { in->from.push_back(o); in->dirty = true; }
}

void Node::disposeNow() {
  if (disposed) return;
  disposed = true;
  for (auto& o : outputs) { auto ins = o->to; for (auto* in : ins) disconnect(o.get(), in); }
  for (auto& in : inputs) { auto outs = in->from; for (auto* o : outs) disconnect(o, in.get()); in->buffer.reset(); }
  for (auto& p : params)  // AudioParam.Dispose: its modulation input is disconnected too
    if (p->mod) { auto outs = p->mod->from; for (auto* o : outs) disconnect(o, p->mod.get()); p->mod->buffer.reset(); }
}

// Nodes/GainNode.cs:29-61
class Gain : public Node {
 public:
  explicit Gain(Context* c) : Node(c, 1, 1) {
    gain = addParam(1.0f, -std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), true);
  }
  void process() override {
    Block& in = *inputs[0]->buffer;
    if (!ob || ob->channels != in.channels) ob = rent(in.channels);
    if (in.silent) { ob->clear(); outputs[0]->buffer = ob; return; }  // :41-46
    ob->copyFrom(in);
    for (int c = 0; c < in.channels; c++) {
      float* s = ob->ch(c);
      for (int i = 0; i < kQuantum; i++) s[i] *= gain->values[i];  // :49-58
    }
    outputs[0]->buffer = ob;
  }
  Param* gain;
  BlockPtr ob;
};

// Nodes/DelayNode.cs
class Delay : public Node {
 public:
  Delay(Context* c, double maxDelayTime);
  void process() override;  // :43-100
  struct Ring {              // CircularBuffer :124-149
    std::vector<float> buf;
    int writePos = 0;
    explicit Ring(int size) : buf((size_t)size, 0.f) {}
    void write(float v) {
      buf[writePos] = v;
      writePos = (writePos + 1) % (int)buf.size();
    }
    float read(int d) const {
      if (d <= 0 || d > (int)buf.size()) return 0.f;
      int pos = (writePos - d + (int)buf.size()) % (int)buf.size();
      return buf[pos];
    }
  };
  Param* delayTime;
  int maxDelaySamples;
  std::vector<Ring> rings;
  BlockPtr ob;
};

// Nodes/StereoPannerNode.cs
class StereoPanner : public Node {
 public:
  explicit StereoPanner(Context* c) : Node(c, 1, 1) {
    inputs[0]->channelCount = 2;  // :24-26 (Speakers interpretation is the default mixing rule of MixBuffer)
    inputs[0]->mode = ModeClampedMax;
    pan = addParam(0.0f, -1.0f, 1.0f, true);
  }
  void process() override {
    Block& in = *inputs[0]->buffer;
    if (!ob || ob->channels != 2) ob = rent(2);
    if (in.silent) { ob->clear(); outputs[0]->buffer = ob; return; }  // :49-54
    float* oL = ob->ch(0);
    float* oR = ob->ch(1);
    const float kPi = 3.14159265358979323846f;  // MathF.PI
    float gL = lastGainL, gR = lastGainR, lp = lastPan;
    if (in.channels == 1) {  // ProcessMono :77-108
      const float* x = in.ch(0);
      for (int i = 0; i < kQuantum; i++) {
        float p = Param::clampf(pan->values[i], -1.0f, 1.0f);
        if (p != lp) {
          float u = (p + 1.0f) * 0.5f;
          gL = cosf(u * kPi / 2.0f);
          gR = sinf(u * kPi / 2.0f);
          lp = p;
        }
        float v = x[i];
        oL[i] = v * gL;
        oR[i] = v * gR;
      }
    } else {  // ProcessStereo :110-152
      const float* xL = in.ch(0);
      const float* xR = in.ch(1);
      for (int i = 0; i < kQuantum; i++) {
        float p = Param::clampf(pan->values[i], -1.0f, 1.0f);
        if (p != lp) {
          float u = p <= 0.0f ? p + 1.0f : p;
          gL = cosf(u * kPi / 2.0f);
          gR = sinf(u * kPi / 2.0f);
          lp = p;
        }
        float a = xL[i], b = xR[i];
        if (p <= 0.0f) {
          oL[i] = a + b * gL;
          oR[i] = b * gR;
        } else {
          oL[i] = a * gL;
          oR[i] = b + a * gR;
        }
      }
    }
    lastPan = lp;
    lastGainL = gL;
    lastGainR = gR;
    ob->markNonSilent();
    outputs[0]->buffer = ob;
  }
  Param* pan;
  float lastPan = std::numeric_limits<float>::quiet_NaN(), lastGainL = 0.5f, lastGainR = 0.5f;
  BlockPtr ob;
};

// Nodes/BiQuadFilterNode.cs
class Biquad : public Node {
 public:
  explicit Biquad(Context* c) : Node(c, 1, 1) {
    freq = addParam(1000.f, 1.f, c->sampleRate / 2.f, true);  // :63-68
    q = addParam(1.0f, 0.001f, 1000.f, true);                 // :70-75
    gain = addParam(0.f, -60.f, 60.f, false);                 // :77-82 (k-rate)
    st.resize(2);
    update(lastFreq, lastQ, lastGain);  // :84
  }
  void setType(int t) { if (type != t) { type = t; dirty = true; } }  // :21-37
  void process() override {  // :87-147
    float gainDb = gain->values[0];
    Block& in = *inputs[0]->buffer;
    int channels = in.channels;
    if ((int)st.size() < channels) st.resize(channels);
    if (!ob || ob->channels != channels) ob = rent(channels);
    if (in.silent) { ob->clear(); outputs[0]->buffer = ob; return; }  // :103-108
    float lb0 = b0, lb1 = b1, lb2 = b2, la1 = a1, la2 = a2;
    float usedFreq = lastFreq;  // never updated: hysteresis reference resets every block (:13-14,111-112)
    float usedQ = lastQ;
    float usedGain = gainDb;
    const float nyq = ctx->sampleRate / 2.f;
    for (int c = 0; c < channels; c++) {
      float* x = in.ch(c);
      float* y = ob->ch(c);
      State& s = st[c];
      for (int i = 0; i < kQuantum; i++) {
        float f = Param::clampf(freq->values[i], 1.f, nyq);  // :123
        float qq = std::max(0.001f, q->values[i]);           // :124
        if (dirty || std::fabs(f - usedFreq) > 0.001f || std::fabs(qq - usedQ) > 0.0001f || std::fabs(gainDb - usedGain) > 0.001f) {
          update(f, qq, gainDb);
          usedFreq = f; usedQ = qq; usedGain = gainDb; dirty = false;
          lb0 = b0; lb1 = b1; lb2 = b2; la1 = a1; la2 = a2;
        }
        float xin = x[i];
        float w = xin - la1 * s.w1 - la2 * s.w2;            // :137
        float yo = lb0 * w + lb1 * s.w1 + lb2 * s.w2;       // :138
        s.w2 = s.w1;
        s.w1 = w;
        y[i] = yo;
      }
    }
    ob->markNonSilent();
    outputs[0]->buffer = ob;
  }
  // UpdateCoefficients :149-258
  void update(float frequency, float qv, float g) {
    float w0 = 2.f * 3.14159274f /* MathF.PI */ * frequency / (float)ctx->sampleRate;
    float cosW0 = cosf(w0);
    float sinW0 = sinf(w0);
    float alpha = sinW0 / (2.f * qv);
    float A0, A1, A2, B0, B1, B2;
    switch (type) {
      case 0:  // Lowpass :160-167
        B0 = (1.f - cosW0) / 2.f; B1 = 1.f - cosW0; B2 = (1.f - cosW0) / 2.f;
        A0 = 1.f + alpha; A1 = -2.f * cosW0; A2 = 1.f - alpha; break;
      case 1:  // Highpass :169-176
        B0 = (1.f + cosW0) / 2.f; B1 = -(1.f + cosW0); B2 = (1.f + cosW0) / 2.f;
        A0 = 1.f + alpha; A1 = -2.f * cosW0; A2 = 1.f - alpha; break;
      case 2:  // Bandpass :178-185
        B0 = alpha; B1 = 0.f; B2 = -alpha;
        A0 = 1.f + alpha; A1 = -2.f * cosW0; A2 = 1.f - alpha; break;
      case 3:  // Notch :187-194
        B0 = 1.f; B1 = -2.f * cosW0; B2 = 1.f;
        A0 = 1.f + alpha; A1 = -2.f * cosW0; A2 = 1.f - alpha; break;
      case 4:  // Allpass :196-203
        B0 = 1.f - alpha; B1 = -2.f * cosW0; B2 = 1.f + alpha;
        A0 = 1.f + alpha; A1 = -2.f * cosW0; A2 = 1.f - alpha; break;
      case 5: {  // Peaking :205-215
        float A = powf(10.f, g / 40.f);
        B0 = 1.f + alpha * A; B1 = -2.f * cosW0; B2 = 1.f - alpha * A;
        A0 = 1.f + alpha / A; A1 = -2.f * cosW0; A2 = 1.f - alpha / A; break;
      }
      case 6: {  // Lowshelf :217-230
        float A = powf(10.f, g / 40.f);
        float sqrtA = sqrtf(A);
        float beta = sqrtA / qv;
        B0 = A * ((A + 1.f) - (A - 1.f) * cosW0 + beta * sinW0);
        B1 = 2.f * A * ((A - 1.f) - (A + 1.f) * cosW0);
        B2 = A * ((A + 1.f) - (A - 1.f) * cosW0 - beta * sinW0);
        A0 = (A + 1.f) + (A - 1.f) * cosW0 + beta * sinW0;
        A1 = -2.f * ((A - 1.f) + (A + 1.f) * cosW0);
        A2 = (A + 1.f) + (A - 1.f) * cosW0 - beta * sinW0; break;
      }
      case 7: {  // Highshelf :232-245
        float A = powf(10.f, g / 40.f);
        float sqrtA = sqrtf(A);
        float beta = sqrtA / qv;
        B0 = A * ((A + 1.f) + (A - 1.f) * cosW0 + beta * sinW0);
        B1 = -2.f * A * ((A - 1.f) + (A + 1.f) * cosW0);
        B2 = A * ((A + 1.f) + (A - 1.f) * cosW0 - beta * sinW0);
        A0 = (A + 1.f) - (A - 1.f) * cosW0 + beta * sinW0;
        A1 = 2.f * ((A - 1.f) - (A + 1.f) * cosW0);
        A2 = (A + 1.f) - (A - 1.f) * cosW0 - beta * sinW0; break;
      }
      default:
        B0 = 1.f; B1 = 0.f; B2 = 0.f; A0 = 1.f; A1 = 0.f; A2 = 0.f; break;
    }
    b0 = B0 / A0; b1 = B1 / A0; b2 = B2 / A0; a1 = A1 / A0; a2 = A2 / A0;  // :253-257
  }
  struct State { float w1 = 0, w2 = 0; };
  Param *freq, *q, *gain;
  int type = 0;
  float lastFreq = 1000.f, lastQ = 1.0f, lastGain = 0.f;
  float b0 = 0, b1 = 0, b2 = 0, a1 = 0, a2 = 0;
  bool dirty = true;
  std::vector<State> st;
  BlockPtr ob;
};

// Nodes/ConvolverNode.cs
class Convolver : public Node {
 public:
  explicit Convolver(Context* c) : Node(c, 1, 1) {}
  // Buffer.set :25-79.  Returns 0 ok, -1 rate mismatch (InvalidOperationException :48-49)
  int setBuffer(PlayBuffer* b, bool normalize, bool enableTrueStereo, int blockSize = kQuantum) {
    if (!b) { convs.clear(); effCh = 0; trueStereo = false; inputs[0]->mode = ModeMax; return 0; }
    if (b->rate != ctx->sampleRate) return -1;
    convs.clear();
    for (int i = 0; i < b->channels; i++)
      convs.emplace_back(new PartitionedConvolver(b->data[i].data(), b->length, blockSize, normalize));  // :51-56
    int ch = b->channels;
    trueStereo = (ch == 4 && enableTrueStereo);  // :64
    effCh = trueStereo ? 2 : ch;
    inputs[0]->channelCount = trueStereo ? 2 : ch;  // :67-76
    inputs[0]->dirty = true;
    inputs[0]->mode = ModeExplicit;
    return 0;
  }
  void process() override {  // :102-155
    Block& in = *inputs[0]->buffer;
    if (convs.empty()) {
      if (!ob || ob->channels != in.channels) ob = rent(in.channels);
      ob->clear();
      outputs[0]->buffer = ob;
      return;
    }
    if (!ob || ob->channels != effCh) ob = rent(effCh);
    if (trueStereo) {  // :127-144
      float t1[kQuantum], t2[kQuantum];
      convs[0]->process(in.ch(0), t1);
      convs[2]->process(in.ch(1), t2);
      for (int i = 0; i < kQuantum; i++) ob->ch(0)[i] = t1[i] + t2[i];
      convs[1]->process(in.ch(0), t1);
      convs[3]->process(in.ch(1), t2);
      for (int i = 0; i < kQuantum; i++) ob->ch(1)[i] = t1[i] + t2[i];
    } else {
      for (int c = 0; c < effCh; c++) convs[c]->process(in.ch(c), ob->ch(c));  // runs regardless of the silent flag
    }
    ob->markNonSilent();  // :153
    outputs[0]->buffer = ob;
  }
  std::vector<std::unique_ptr<PartitionedConvolver>> convs;
  int effCh = 0;
  bool trueStereo = false;
  BlockPtr ob;
};

// Nodes/AudioBufferSourceNode.cs:79-402 (all four paths: rate 1 / resampled, with and without Loop)
class BufferSource : public Node {
 public:
  explicit BufferSource(Context* c) : Node(c, 0, 1) { rate = addParam(1.f, 0.001f, 1000.f, false); }  // :76
  // Start :79-114
  int start(double when, double off, double dur) {
    if (started || !buf) return -1;
    started = true;
    startTime = std::max(0.0, when);
    offset = std::max(0.0, off);
    duration = dur;
    pos = (int64_t)(offset * buf->rate);  // :96
    for (auto& r : rs) r.clear();
    if (!std::isinf(dur) && dur >= 0) { stopTime = startTime + dur; stopped = true; }
    return 0;
  }
  void stop(double when) {  // :116-129
    if (stopped) return;
    double at = std::max(0.0, when);
    stopTime = std::isnan(stopTime) ? at : std::min(stopTime, at);
    stopped = true;
  }
  void silence() {  // ProduceSilence :391-402 (1-channel silent block)
    if (!ob || ob->channels != 1) ob = rent(1);
    ob->clear();
    outputs[0]->buffer = ob;
  }
  void process() override {  // :131-376
    double t0 = ctx->currentTime;
    double t1 = t0 + (double)kQuantum / ctx->sampleRate;
    bool shouldPlay = started && (t1 > startTime && (std::isnan(stopTime) || t0 < stopTime));  // :137-143
    if (!shouldPlay || !buf) { silence(); return; }
    int oc = buf->channels;
    if (!ob || ob->channels != oc) ob = rent(oc);
    float playbackRate = rate->values[0];
    const int frames = kQuantum;
    double ratio = buf->rate / (double)ctx->sampleRate;  // :168
    double eff = ratio * playbackRate;                   // :169
    int64_t durEnd = duration < std::numeric_limits<double>::infinity()
                         ? (int64_t)(offset * buf->rate) + (int64_t)(duration * buf->rate)
                         : buf->length;  // :179-182
    durEnd = std::min(durEnd, buf->length);
    int64_t loopStartFrame = (int64_t)(loopStart * buf->rate);  // :171-177
    int64_t loopEndFrame = loopEnd > 0 ? (int64_t)(loopEnd * buf->rate) : buf->length;
    loopEndFrame = std::min(loopEndFrame, buf->length);
    loopStartFrame = std::min(loopStartFrame, loopEndFrame);
    bool hasMore = false;
    if (eff == 1.0) {  // :186-235
      for (int c = 0; c < oc; c++) {
        const float* d = buf->data[c].data();
        float* o = ob->ch(c);
        int64_t p = pos;
        int oi = 0;
        while (oi < frames) {
          if (loop && p >= loopEndFrame) p = loopStartFrame;                                     // :197-200
          if (p >= durEnd && !loop) { std::fill(o + oi, o + frames, 0.f); break; }               // :202-206
          int64_t endFrame = loop ? loopEndFrame : std::min(durEnd, buf->length);                // :208
          int avail = (int)std::min<int64_t>(endFrame - p, frames - oi);
          if (avail <= 0) { std::fill(o + oi, o + frames, 0.f); break; }
          std::memcpy(o + oi, d + p, sizeof(float) * avail);
          p += avail; oi += avail; hasMore = true;
        }
      }
      pos += frames;  // :224
      if (loop && pos >= loopEndFrame) {  // :226-234
        int64_t loopLength = loopEndFrame - loopStartFrame;
        if (loopLength > 0) pos = loopStartFrame + ((pos - loopEndFrame) % loopLength);
      }
    } else if (loop) {  // :236-358 with _loop set: every Process call is fed from the 512-float wrap buffer (:296-314)
      // (available = loopEndFrame - pos because loopEndFrame <= Length, so `pos + available >= loopEndFrame - 4` (:296) always holds)
      if ((int)rs.size() != oc) { rs.assign(oc, CubicResampler()); }
      float wrap[512];
      int64_t totalConsumed = 0;
      for (int c = 0; c < oc; c++) {
        const float* d = buf->data[c].data();
        float* o = ob->ch(c);
        int64_t p = pos, consumedCh = 0;
        int oi = 0;
        while (oi < frames) {
          if (p >= loopEndFrame) p = loopStartFrame;                                          // :265-268
          int avail = (int)std::min<int64_t>(loopEndFrame - p, buf->length - p);             // :276-277
          if (avail <= 0) {                                                                   // :279-285 (an empty loop region never leaves this branch)
            ctx->unsupported = true;
            std::fill(o + oi, o + frames, 0.f);
            break;
          }
          int64_t loopLength = loopEndFrame - loopStartFrame;
          int fromEnd = (int)(loopEndFrame - p);
          int needed = std::min(frames - oi + 4, 512);                                        // :301
          int copied = 0;
          for (int i = 0; i < fromEnd && copied < needed; i++) wrap[copied++] = d[p + i];     // :303-306
          for (int64_t i = 0; copied < needed && i < loopLength; i++) wrap[copied++] = d[loopStartFrame + i];  // :308-311
          int ic, op;
          rs[c].process(wrap, copied, o + oi, frames - oi, eff, &ic, &op);                    // :313
          if (op > 0) hasMore = true;
          int64_t np = p + ic;
          if (np >= loopEndFrame) np = loopStartFrame + (np - loopEndFrame);                  // :323-328 (no modulo here)
          consumedCh += (np >= p) ? (np - p) : (loopEndFrame - p + np - loopStartFrame);      // :330
          p = np; oi += op;
          if (ic == 0 && op == 0) { std::fill(o + oi, o + frames, 0.f); break; }              // :334-338
        }
        if (c == 0) totalConsumed = consumedCh;
      }
      pos += totalConsumed;                                                                   // :347
      if (pos >= loopEndFrame) {                                                              // :349-357
        int64_t loopLength = loopEndFrame - loopStartFrame;
        if (loopLength > 0) pos = loopStartFrame + ((pos - loopEndFrame) % loopLength);
      }
    } else {  // :236-358
      if ((int)rs.size() != oc) { rs.assign(oc, CubicResampler()); }
      int64_t totalConsumed = 0;
      for (int c = 0; c < oc; c++) {
        const float* d = buf->data[c].data();
        float* o = ob->ch(c);
        int64_t p = pos, consumedCh = 0;
        int oi = 0;
        while (oi < frames) {
          if (p >= durEnd) { std::fill(o + oi, o + frames, 0.f); break; }
          int64_t endFrame = std::min(durEnd, buf->length);
          int avail = (int)std::min<int64_t>(endFrame - p, buf->length - p);
          if (avail <= 0) { std::fill(o + oi, o + frames, 0.f); break; }
          int ic, op;
          rs[c].process(d + p, avail, o + oi, frames - oi, eff, &ic, &op);  // :330
          if (op > 0) hasMore = true;
          int64_t np = p + ic;
          consumedCh += np - p;
          p = np; oi += op;
          if (ic == 0 && op == 0) { std::fill(o + oi, o + frames, 0.f); break; }  // :334-338
        }
        if (c == 0) totalConsumed = consumedCh;  // :341-344
      }
      pos += totalConsumed;
    }
    double tEnd = t1;
    if (!hasMore || (!loop && pos >= durEnd)) {  // :360-368
      ob->clear();
      if (std::isnan(stopTime)) { stopTime = t1; stopped = true; }
    } else {
      ob->markNonSilent();
    }
    outputs[0]->buffer = ob;
    // TryRaiseEndedEvent :378-389 -> Dispose() is posted (we are mid-render) and runs at the next block
    if (started && !std::isnan(stopTime) && tEnd >= stopTime && !ended) {
      ended = true;
      ctx->commands.push_back([this]() { disposeNow(); });
    }
  }
  PlayBuffer* buf = nullptr;
  Param* rate;
  bool loop = false;                    // :40-44
  double loopStart = 0, loopEnd = 0;    // :49-62 (seconds; loopEnd 0 = end of buffer)
  bool started = false, stopped = false, ended = false;
  double startTime = std::numeric_limits<double>::quiet_NaN(), stopTime = std::numeric_limits<double>::quiet_NaN();
  double offset = 0, duration = std::numeric_limits<double>::infinity();
  int64_t pos = 0;
  std::vector<CubicResampler> rs;
  BlockPtr ob;
};

// Nodes/ChannelSplitterNode.cs: output i = channel i of the input as a 1-channel block (cleared when the input has fewer channels or is silent)
class ChannelSplitter : public Node {
 public:
  ChannelSplitter(Context* c, int n) : Node(c, 1, n), obs((size_t)n) {}
  void process() override {  // :21-55
    Block* in = inputs[0]->buffer.get();
    const bool silent = !in || in->silent;
    for (size_t i = 0; i < obs.size(); i++) {
      if (!obs[i]) obs[i] = rent(1);
      if (!silent && (int)i < in->channels) {
        std::memcpy(obs[i]->ch(0), in->ch((int)i), sizeof(float) * kQuantum);  // CopyChannelFrom clears the silent flag (AudioBuffer.cs:110-121)
        obs[i]->markNonSilent();
      } else {
        obs[i]->clear();
      }
      outputs[i]->buffer = obs[i];
    }
  }
  std::vector<BlockPtr> obs;
};

// Nodes/ChannelMergerNode.cs: output channel i = channel 0 of input i (silent inputs leave zeros)
class ChannelMerger : public Node {
 public:
  ChannelMerger(Context* c, int n) : Node(c, n, 1), n_(n) {}
  void process() override {  // :21-52
    if (!ob || ob->channels != n_) ob = rent(n_);
    ob->clear();
    bool hasAudio = false;
    for (int i = 0; i < n_; i++) {
      Block* in = inputs[i]->buffer.get();
      if (in && !in->silent && in->channels > 0) {
        std::memcpy(ob->ch(i), in->ch(0), sizeof(float) * kQuantum);
        hasAudio = true;
      }
    }
    if (hasAudio) ob->markNonSilent();
    outputs[0]->buffer = ob;
  }
  int n_;
  BlockPtr ob;
};

// Start / Stop bookkeeping shared by OscillatorNode and ConstantSourceNode (IAudioScheduledSourceNode): sample-accurate start and
// stop inside a block (Nodes/OscillatorNode.cs:97-118, Nodes/ConstantSourceNode.cs:82-110)
class ScheduledSource : public Node {
 public:
  explicit ScheduledSource(Context* c) : Node(c, 0, 1) {}
  int start(double when, double duration) {
    if (started) return -1;
    started = true;
    startTime = std::max(0.0, when);
    if (!std::isnan(duration) && duration >= 0) { stopTime = startTime + duration; stopped = true; }
    return 0;
  }
  void stop(double when) {
    if (stopped) return;
    double at = std::max(0.0, when);
    stopTime = std::isnan(stopTime) ? at : std::min(stopTime, at);
    stopped = true;
  }
  // returns false when the block is silent; else [startFrame, endFrame) is the playing part of the block
  bool window(int* startFrame, int* endFrame) const {
    double t0 = ctx->currentTime;
    double t1 = t0 + (double)kQuantum / ctx->sampleRate;
    *startFrame = 0;
    *endFrame = kQuantum;
    if (!started || !(t1 > startTime && (std::isnan(stopTime) || t0 < stopTime))) return false;
    auto clampi = [](double v) { return (int)(v < 0 ? 0 : (v > kQuantum ? kQuantum : v)); };
    if (t0 < startTime && startTime < t1) *startFrame = clampi(std::ceil((startTime - t0) * ctx->sampleRate));
    if (!std::isnan(stopTime) && t0 < stopTime && stopTime < t1) *endFrame = clampi(std::floor((stopTime - t0) * ctx->sampleRate));
    return true;
  }
  void endCheck() {  // TryRaiseEndedAndDispose: Dispose() is posted mid-render and runs at the next block
    double t1 = ctx->currentTime + (double)kQuantum / ctx->sampleRate;
    if (started && stopped && !ended && !std::isnan(stopTime) && t1 >= stopTime) {
      ended = true;
      ctx->commands.push_back([this]() { disposeNow(); });
    }
  }
  bool started = false, stopped = false, ended = false;
  double startTime = std::numeric_limits<double>::quiet_NaN(), stopTime = std::numeric_limits<double>::quiet_NaN();
  BlockPtr ob;
};

// Nodes/OscillatorNode.cs
class Oscillator : public ScheduledSource {
 public:
  explicit Oscillator(Context* c) : ScheduledSource(c) { frequency = addParam(440.f, 0.f, c->sampleRate / 2.f, true); }  // :42-47
  static float sample(double phase, int type) {  // GenerateSample :164-188
    const double kPi = 3.14159265358979323846;
    switch (type) {
      case 0: return (float)std::sin(phase);
      case 1: return phase < kPi ? 1.0f : -1.0f;
      case 2: return (float)(2.0 * (phase / (2.0 * kPi)) - 1.0);
      case 3: { double t = phase / (2.0 * kPi); return (float)(4.0 * std::fabs(t - std::floor(t + 0.5)) - 1.0); }
      default: return 0.f;
    }
  }
  void process() override {  // :91-150
    if (!ob) ob = rent(1);
    int sf, ef;
    if (!window(&sf, &ef)) { ob->clear(); outputs[0]->buffer = ob; endCheck(); return; }
    const double kPi = 3.14159265358979323846;
    float* o = ob->ch(0);
    for (int i = 0; i < sf; i++) o[i] = 0.f;
    for (int i = sf; i < ef; i++) {
      o[i] = sample(phase, type);
      double inc = (2.0 * kPi * frequency->values[i]) / ctx->sampleRate;  // :133
      phase += inc;
      if (phase >= 2.0 * kPi) phase -= 2.0 * kPi;
    }
    for (int i = ef; i < kQuantum; i++) o[i] = 0.f;
    ob->markNonSilent();
    outputs[0]->buffer = ob;
    endCheck();
  }
  Param* frequency;
  int type = 0;  // OscillatorType: Sine, Square, Sawtooth, Triangle (:207-213)
  double phase = 0.0;
};

// Nodes/ConstantSourceNode.cs
class ConstantSource : public ScheduledSource {
 public:
  explicit ConstantSource(Context* c) : ScheduledSource(c) {
    offset = addParam(1.f, -std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), true);  // :22-28
  }
  void process() override {  // :68-137
    if (!ob) ob = rent(1);
    int sf, ef;
    if (!window(&sf, &ef)) { ob->clear(); outputs[0]->buffer = ob; endCheck(); return; }
    float* o = ob->ch(0);
    for (int i = 0; i < sf; i++) o[i] = 0.f;
    for (int i = sf; i < ef; i++) o[i] = offset->values[i];
    for (int i = ef; i < kQuantum; i++) o[i] = 0.f;
    ob->markNonSilent();
    outputs[0]->buffer = ob;
    endCheck();
  }
  Param* offset;
};

Delay::Delay(Context* c, double maxDelayTime) : Node(c, 1, 1) {  // :22-41
  maxDelaySamples = (int)(maxDelayTime * c->sampleRate);
  for (int i = 0; i < 2; i++) rings.emplace_back(maxDelaySamples);
  delayTime = addParam(0.0f, 0.0f, (float)maxDelayTime, true);
}
void Delay::process() {
  Block* in = inputs[0]->buffer.get();
  int channels = in ? in->channels : 2;
  while ((int)rings.size() < channels) rings.emplace_back(maxDelaySamples);  // EnsureChannelCount :102-113
  if (!ob || ob->channels != channels) ob = rent(channels);
  bool hasAudio = false;
  const bool silentIn = !in || in->silent;
  for (int ch = 0; ch < channels; ch++) {
    float* o = ob->ch(ch);
    const float* x = silentIn ? nullptr : in->ch(ch);
    for (int i = 0; i < kQuantum; i++) {
      int d = (int)(delayTime->values[i] * ctx->sampleRate);  // float * int -> float, truncated (:68,85)
      d = d < 0 ? 0 : (d > maxDelaySamples ? maxDelaySamples : d);
      o[i] = rings[ch].read(d);
      rings[ch].write(x ? x[i] : 0.f);
      if (o[i] != 0.f) hasAudio = true;
    }
  }
  // the node keeps ONE pooled block and only ever sets the flag (:96-97): silent until the first block that carries audio,
  // non-silent from then on (until a channel-count change rents a fresh block)
  if (hasAudio) ob->markNonSilent();
  outputs[0]->buffer = ob;
}

Context::Context(int fs) : sampleRate(fs) {
  destination = new Destination(this);
  nodes.emplace_back(destination);
}
Context::~Context() {}

// AudioContextBase.ProcessBlock :52-81
Block* Context::processBlock() {
  while (!commands.empty()) {  // DrainCommands :272-284
    auto cmd = std::move(commands.front());
    commands.pop_front();
    cmd();
  }
  int next = currentBlock + 1;
  currentBlock = next;
  double blockTime = currentTime;
  destination->processInternal(next, blockTime);
  double inc = (double)kQuantum / sampleRate;
  currentTime = blockTime + inc;  // accumulated, not block*dt (:78-79)
  return destination->out.get();
}

// OfflineAudioContext.Render :30-102
int Context::render(float* const* out, int channels, int frameCount, int startIndex) {
  if (channels <= 0 || frameCount <= 0 || startIndex < 0) return -1;
  int written = 0;
  if (cached > 0) {  // :55-75
    int n = std::min(cached, frameCount);
    for (int c = 0; c < channels; c++)
      if (c < (int)cache.size()) std::memcpy(out[c] + startIndex, cache[c].data(), sizeof(float) * n);
    if (n < cached)
      for (auto& cc : cache) std::memmove(cc.data(), cc.data() + n, sizeof(float) * (cached - n));
    written = n;
    cached -= n;
  }
  while (written < frameCount) {  // :77-101
    Block* b = processBlock();
    if (cycle) return -2;
    int n = std::min(kQuantum, frameCount - written);
    for (int c = 0; c < channels; c++) {
      if (c < b->channels) std::memcpy(out[c] + startIndex + written, b->ch(c), sizeof(float) * n);
      else return -3;  // GetChannelSpan would throw ArgumentOutOfRange (AudioBuffer.cs:39-40)
    }
    written += n;
    int excess = kQuantum - n;
    if (excess > 0) {
      if ((int)cache.size() < channels) cache.resize(channels);
      for (int c = 0; c < channels; c++) {
        cache[c].resize(cached + excess);
        std::memcpy(cache[c].data() + cached, b->ch(c) + n, sizeof(float) * excess);
      }
      cached += excess;
    }
  }
  return 0;
}

}  // namespace ora

// ---------------------------------------------------------------------------------------------
// C entry points (ctypes-friendly)
// ---------------------------------------------------------------------------------------------
using namespace ora;

extern "C" {

// ---- primitives (kernel-level parity) ----
int ora_rfft_forward(int n, const double* x, double* re, double* im) {
  RealFFT f(n);
  if (!f.valid()) return -1;
  f.forward(x, re, im);
  return 0;
}
int ora_rfft_inverse(int n, const double* re, const double* im, double* x) {
  RealFFT f(n);
  if (!f.valid()) return -1;
  f.inverse(re, im, x);
  return 0;
}
float ora_normalization_scale(const float* ir, int64_t n) { return PartitionedConvolver::normalizationScale(ir, n); }

void* ora_pc_create(const float* ir, int64_t n, int blockSize, int normalize) {
  return new PartitionedConvolver(ir, n, blockSize, normalize != 0);
}
void ora_pc_destroy(void* h) { delete (PartitionedConvolver*)h; }
int ora_pc_partitions(void* h) { return ((PartitionedConvolver*)h)->partitions(); }
// copies the planar float32 IR spectra [P][C]
void ora_pc_ir_spectra(void* h, float* re, float* im) {
  auto* pc = (PartitionedConvolver*)h;
  size_t n = (size_t)pc->partitions() * pc->bins();
  std::memcpy(re, pc->irRe(), n * sizeof(float));
  std::memcpy(im, pc->irIm(), n * sizeof(float));
}
// runs nBlocks consecutive Process() calls
void ora_pc_process(void* h, const float* in, float* out, int blockSize, int64_t nBlocks) {
  auto* pc = (PartitionedConvolver*)h;
  for (int64_t b = 0; b < nBlocks; b++) pc->process(in + b * blockSize, out + b * blockSize);
}

// CubicResampler over a whole buffer in one Process call; returns produced count
int64_t ora_resample(const float* in, int64_t inLen, float* out, int64_t outLen, double rate, int64_t* consumed) {
  CubicResampler r;
  int64_t ip = 0, op = 0;
  // chunked so that int counters never overflow
  while (op < outLen) {
    int ic, oc;
    int il = (int)std::min<int64_t>(inLen - ip, 1 << 30);
    int ol = (int)std::min<int64_t>(outLen - op, 1 << 30);
    r.process(in + ip, il, out + op, ol, rate, &ic, &oc);
    ip += ic; op += oc;
    if (ic == 0 && oc == 0) break;
  }
  if (consumed) *consumed = ip;
  return op;
}

// ---- graph ----
void* ora_context_create(int sampleRate) { return sampleRate > 0 ? new Context(sampleRate) : nullptr; }
void ora_context_destroy(void* c) { delete (Context*)c; }

int ora_buffer_create(void* c, const float* const* ch, int nch, int64_t len, int rate) {
  auto* ctx = (Context*)c;
  if (nch < 1 || nch > 32 || len < 0 || rate <= 0) return -1;
  auto b = std::make_unique<PlayBuffer>();
  b->channels = nch; b->rate = rate; b->length = len;
  for (int i = 0; i < nch; i++) b->data.emplace_back(ch[i], ch[i] + len);
  ctx->buffers.push_back(std::move(b));
  return (int)ctx->buffers.size() - 1;
}

// kind: 0 AudioBufferSourceNode, 1 BiQuadFilterNode, 2 GainNode, 3 ConvolverNode.  Node 0 is the destination.
int ora_node_create(void* c, int kind) {
  auto* ctx = (Context*)c;
  Node* n = nullptr;
  switch (kind) {
    case 0: n = new BufferSource(ctx); break;
    case 1: n = new Biquad(ctx); break;
    case 2: n = new Gain(ctx); break;
    case 3: n = new Convolver(ctx); break;
    case 4: n = new StereoPanner(ctx); break;
    case 5: n = new Oscillator(ctx); break;
    case 6: n = new ConstantSource(ctx); break;
    default: return -1;
  }
  ctx->nodes.emplace_back(n);
  return (int)ctx->nodes.size() - 1;
}

int ora_delay_create(void* c, double maxDelayTime) {  // DelayNode(context, maxDelayTime) :22-26
  auto* ctx = (Context*)c;
  if (maxDelayTime <= 0 || maxDelayTime > 10) return -1;  // ArgumentOutOfRangeException
  ctx->nodes.emplace_back(new Delay(ctx, maxDelayTime));
  return (int)ctx->nodes.size() - 1;
}

static Node* nodeAt(void* c, int id) {
  auto* ctx = (Context*)c;
  if (id < 0 || id >= (int)ctx->nodes.size()) return nullptr;
  return ctx->nodes[id].get();
}

int ora_connect(void* c, int src, int dst) {  // AudioNode.Connect :68-73 (applied in call order)
  Node *a = nodeAt(c, src), *b = nodeAt(c, dst);
  if (!a || !b || a->outputs.empty() || b->inputs.empty() || a == b) return -1;
  connect(a->outputs[0].get(), b->inputs[0].get());
  return 0;
}

int ora_disconnect(void* c, int src, int dst) {  // AudioNode.Disconnect(destination) :78-84,129-147; dst < 0: every connection of output 0
  Node* a = nodeAt(c, src);
  if (!a || a->outputs.empty()) return -1;
  Output* o = a->outputs[0].get();
  if (dst < 0) {
    while (!o->to.empty()) disconnect(o, o->to.back());
    return 0;
  }
  Node* b = nodeAt(c, dst);
  if (!b || b->inputs.empty()) return -1;
  disconnect(o, b->inputs[0].get());
  return 0;
}

int ora_splitter_create(void* c, int n) {  // ChannelSplitterNode(context, numberOfOutputs) :12-20
  auto* ctx = (Context*)c;
  if (n < 1 || n > 32) return -1;
  ctx->nodes.emplace_back(new ChannelSplitter(ctx, n));
  return (int)ctx->nodes.size() - 1;
}
int ora_merger_create(void* c, int n) {  // ChannelMergerNode(context, numberOfInputs) :12-19
  auto* ctx = (Context*)c;
  if (n < 1 || n > 32) return -1;
  ctx->nodes.emplace_back(new ChannelMerger(ctx, n));
  return (int)ctx->nodes.size() - 1;
}
int ora_connect_io(void* c, int src, int outIdx, int dst, int inIdx) {  // AudioNode.Connect(destination, outputIndex, inputIndex) :68-84
  Node *a = nodeAt(c, src), *b = nodeAt(c, dst);
  if (!a || !b || a == b || outIdx < 0 || outIdx >= (int)a->outputs.size() || inIdx < 0 || inIdx >= (int)b->inputs.size()) return -1;
  connect(a->outputs[outIdx].get(), b->inputs[inIdx].get());
  return 0;
}
static Param* paramAt(void* c, int node, int pidx);
int ora_connect_param(void* c, int src, int dstNode, int pidx) {  // AudioNode.Connect(AudioParam) :86-92
  Node* a = nodeAt(c, src);
  Param* p = paramAt(c, dstNode, pidx);
  if (!a || !p || a->outputs.empty()) return -1;
  connect(a->outputs[0].get(), p->mod.get());
  return 0;
}
int ora_scheduled_start(void* c, int node, double when, double duration) {
  auto* s = dynamic_cast<ScheduledSource*>(nodeAt(c, node));
  return s ? s->start(when, duration) : -1;
}
int ora_scheduled_stop(void* c, int node, double when) {
  auto* s = dynamic_cast<ScheduledSource*>(nodeAt(c, node));
  if (!s) return -1;
  s->stop(when);
  return 0;
}
int ora_oscillator_set_type(void* c, int node, int type) {
  auto* o = dynamic_cast<Oscillator*>(nodeAt(c, node));
  if (!o || type < 0 || type > 3) return -1;
  o->type = type;
  return 0;
}

static Param* paramAt(void* c, int node, int pidx) {
  Node* n = nodeAt(c, node);
  if (!n || pidx < 0 || pidx >= (int)n->params.size()) return nullptr;
  return n->params[pidx].get();
}
int ora_param_set_value(void* c, int node, int pidx, float v) {
  Param* p = paramAt(c, node, pidx);
  if (!p) return -1;
  p->setValue(v);
  return 0;
}
// type: 0 SetValueAtTime, 1 LinearRampToValueAtTime, 2 ExponentialRampToValueAtTime, 3 SetTargetAtTime
int ora_param_event(void* c, int node, int pidx, int type, float value, double time, double tc) {
  Param* p = paramAt(c, node, pidx);
  if (!p) return -1;
  switch (type) {
    case 0: p->setValueAtTime(value, time); return 0;
    case 1: p->linearRamp(value, time); return 0;
    case 2: return p->exponentialRamp(value, time) ? 0 : -2;
    case 3: p->setTarget(value, time, tc); return 0;
  }
  return -1;
}
int ora_param_cancel(void* c, int node, int pidx, double t) {
  Param* p = paramAt(c, node, pidx);
  if (!p) return -1;
  p->cancel(t);
  return 0;
}
// evaluates n consecutive blocks of an a-rate/k-rate param starting at block 0 (time accumulates as in ProcessBlock)
int ora_param_eval(void* c, int node, int pidx, int64_t nBlocks, float* out) {
  Param* p = paramAt(c, node, pidx);
  auto* ctx = (Context*)c;
  if (!p) return -1;
  double t = 0.0;
  for (int64_t b = 0; b < nBlocks; b++) {
    p->computeValues((int)b + 1, t, ctx->sampleRate);
    std::memcpy(out + b * kQuantum, p->values, sizeof(float) * kQuantum);
    t = t + (double)kQuantum / ctx->sampleRate;
  }
  return 0;
}

int ora_source_set_buffer(void* c, int node, int buf) {
  auto* n = dynamic_cast<BufferSource*>(nodeAt(c, node));
  auto* ctx = (Context*)c;
  if (!n || buf < 0 || buf >= (int)ctx->buffers.size()) return -1;
  n->buf = ctx->buffers[buf].get();
  return 0;
}
int ora_source_start(void* c, int node, double when, double offset, double duration) {
  auto* n = dynamic_cast<BufferSource*>(nodeAt(c, node));
  return n ? n->start(when, offset, duration) : -1;
}
int ora_source_set_loop(void* c, int node, int loop, double loopStart, double loopEnd) {
  auto* s = dynamic_cast<BufferSource*>(nodeAt(c, node));
  if (!s) return -1;
  s->loop = loop != 0;
  s->loopStart = std::max(0.0, loopStart);  // :52
  s->loopEnd = std::max(0.0, loopEnd);      // :61
  return 0;
}
int ora_unsupported(void* c) { return ((Context*)c)->unsupported ? 1 : 0; }
int ora_source_stop(void* c, int node, double when) {
  auto* n = dynamic_cast<BufferSource*>(nodeAt(c, node));
  if (!n) return -1;
  n->stop(when);
  return 0;
}
int ora_biquad_set_type(void* c, int node, int type) {
  auto* n = dynamic_cast<Biquad*>(nodeAt(c, node));
  if (!n) return -1;
  n->setType(type);
  return 0;
}
// blockSize is always 128 through ConvolverNode (ConvolverNode.cs:55)
int ora_convolver_set_buffer(void* c, int node, int buf, int normalize, int trueStereo) {
  auto* n = dynamic_cast<Convolver*>(nodeAt(c, node));
  auto* ctx = (Context*)c;
  if (!n) return -1;
  if (buf < 0) return n->setBuffer(nullptr, normalize != 0, trueStereo != 0);
  if (buf >= (int)ctx->buffers.size()) return -1;
  return n->setBuffer(ctx->buffers[buf].get(), normalize != 0, trueStereo != 0);
}
int ora_render(void* c, float* const* out, int channels, int frameCount, int startIndex) {
  return ((Context*)c)->render(out, channels, frameCount, startIndex);
}

}  // extern "C"
