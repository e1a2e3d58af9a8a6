// graphaudio_cuda.hpp — header-only C++17 mirror of the reference's public API for the offline render path,
// layered on the C ABI of include/graphaudio_cuda.h.
//
// The reference is C# (net9.0); its toolchain is absent from this image, so the host side above the ABI is written in
// C++ here (and in Python in graphaudio_b200/api.py) with the reference's type and member names, argument meaning and
// error behaviour, so that code written against GraphAudio.Core reads the same:
//
//     GraphAudio::Cuda::OfflineAudioContext ctx(48000);
//     auto src  = ctx.CreateBufferSource();   src->Buffer = PlayableAudioBuffer::FromStereoArrays(l, r, 48000);
//     auto gain = ctx.CreateGain();           gain->Gain.LinearRampToValueAtTime(0.5f, 2.0);
//     auto conv = ctx.CreateConvolver();      conv->SetBuffer(ir);            // ConvolverNode.Buffer = ir
//     src->Connect(gain)->Connect(conv)->Connect(ctx.Destination());
//     src->Start();
//     ctx.Render(output, frameCount);          // OfflineAudioContext.Render(float[][] output, int frameCount, int startIndex = 0)
//
// Nodes only RECORD topology and automation (AudioNode.Connect is queued in the reference as well,
// Nodes/AudioNode.cs:109-123); Render flattens the graph into gac_voice_desc / gac_bus_desc and makes ONE call into
// libgraphaudio_cuda.so.  There is no CPU fallback.  Citations are relative to /root/reference/GraphAudio.Core/.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/graphaudio_cuda.h"

namespace GraphAudio {
namespace Cuda {

// ---- the exception types the reference throws, keyed by gac_status -----------------------------------------
struct ArgumentException : std::invalid_argument { using std::invalid_argument::invalid_argument; };
struct ArgumentOutOfRangeException : ArgumentException { using ArgumentException::ArgumentException; };
struct InvalidOperationException : std::logic_error { using std::logic_error::logic_error; };
struct ObjectDisposedException : InvalidOperationException { using InvalidOperationException::InvalidOperationException; };
struct NotSupportedException : std::runtime_error { using std::runtime_error::runtime_error; };
struct CudaException : std::runtime_error { using std::runtime_error::runtime_error; };

inline void Check(int rc) {
  if (rc == GAC_OK) return;
  std::string msg = gac_last_error();
  switch (rc) {
    case GAC_ERR_INVALID_ARGUMENT: throw ArgumentException(msg);
    case GAC_ERR_OUT_OF_RANGE: throw ArgumentOutOfRangeException(msg);
    case GAC_ERR_INVALID_OPERATION: throw InvalidOperationException(msg);
    case GAC_ERR_DISPOSED: throw ObjectDisposedException(msg);
    case GAC_ERR_UNSUPPORTED: throw NotSupportedException(msg);
    default: throw CudaException(msg);
  }
}

enum class FilterType { Lowpass, Highpass, Bandpass, Notch, Allpass, Peaking, Lowshelf, Highshelf };  // BiQuadFilterNode.cs:288-298

class OfflineAudioContext;

// ---- PlayableAudioBuffer.cs -----------------------------------------------------------------------------------
class PlayableAudioBuffer {
 public:
  static std::shared_ptr<PlayableAudioBuffer> FromChannelArrays(const std::vector<std::vector<float>>& channelData, int sampleRate) {  // :122-143
    if (channelData.empty()) throw ArgumentException("Channel data cannot be or empty");
    for (auto& c : channelData)
      if (c.size() != channelData[0].size()) throw ArgumentException("All channels must have the same length");
    if (sampleRate <= 0) throw ArgumentOutOfRangeException("Sample rate must be positive");
    auto b = std::shared_ptr<PlayableAudioBuffer>(new PlayableAudioBuffer());
    b->channels_ = channelData;
    b->sampleRate_ = sampleRate;
    return b;
  }
  static std::shared_ptr<PlayableAudioBuffer> FromMonoArray(const std::vector<float>& audioData, int sampleRate) {  // :148-157
    return FromChannelArrays({audioData}, sampleRate);
  }
  static std::shared_ptr<PlayableAudioBuffer> FromStereoArrays(const std::vector<float>& l, const std::vector<float>& r, int sampleRate) {  // :162-173
    if (l.size() != r.size()) throw ArgumentException("Left and right channels must have the same length");
    return FromChannelArrays({l, r}, sampleRate);
  }
  int NumberOfChannels() const { return (int)channels_.size(); }
  int64_t Length() const { return (int64_t)channels_[0].size(); }
  int SampleRate() const { return sampleRate_; }
  ~PlayableAudioBuffer() { if (handle_) gac_buffer_destroy(handle_); }

 private:
  friend class OfflineAudioContext;
  friend class ConvolverNode;
  PlayableAudioBuffer() = default;
  gac_buffer* Handle(gac_context* ctx) {  // uploaded to HBM on first use (the copy of CopyToChannel, :84-93)
    if (!handle_) {
      std::vector<const float*> p;
      for (auto& c : channels_) p.push_back(c.data());
      Check(gac_buffer_create(ctx, p.data(), NumberOfChannels(), Length(), sampleRate_, &handle_));
    }
    return handle_;
  }
  std::vector<std::vector<float>> channels_;
  int sampleRate_ = 0;
  gac_buffer* handle_ = nullptr;
};

// ---- AudioParam.cs --------------------------------------------------------------------------------------------
class AudioParam {
 public:
  AudioParam(float def, float mn, float mx) : value_(def), min_(mn), max_(mx) {}
  float Value() const { return value_; }
  void SetValue(float v) { value_ = Clamp(v); events_.clear(); }  // Value setter cancels scheduled events (:34-49)
  void SetValueAtTime(float v, double t) { Add({GAC_EVENT_SET_VALUE, Clamp(v), 0.f, t, 0.0}); }                // :252
  void LinearRampToValueAtTime(float v, double t) { Add({GAC_EVENT_LINEAR_RAMP, Clamp(v), 0.f, t, 0.0}); }     // :266
  void ExponentialRampToValueAtTime(float v, double t) {                                                       // :280
    v = Clamp(v);
    if (v <= 0.f) throw ArgumentException("Exponential ramp target must be > 0");
    Add({GAC_EVENT_EXPONENTIAL_RAMP, v, 0.f, t, 0.0});
  }
  void SetTargetAtTime(float target, double t, double tc) { Add({GAC_EVENT_SET_TARGET, 0.f, Clamp(target), t, tc}); }  // :297
  void CancelScheduledValues(double t) {                                                                       // :312
    events_.erase(std::remove_if(events_.begin(), events_.end(), [t](const gac_event& e) { return e.time >= t; }), events_.end());
  }
  gac_param Desc() const { return gac_param{value_, (int32_t)events_.size(), events_.empty() ? nullptr : events_.data()}; }

 private:
  float Clamp(float v) const { return std::min(std::max(v, min_), max_); }
  void Add(gac_event e) {  // AddEvent: stable upper-bound insert (:333-352)
    auto it = std::upper_bound(events_.begin(), events_.end(), e, [](const gac_event& a, const gac_event& b) { return a.time < b.time; });
    events_.insert(it, e);
  }
  float value_, min_, max_;
  std::vector<gac_event> events_;
};

// ---- Nodes/AudioNode.cs ---------------------------------------------------------------------------------------
class AudioNode : public std::enable_shared_from_this<AudioNode> {
 public:
  virtual ~AudioNode() = default;
  // Connect returns the destination to allow chaining (:68-73)
  std::shared_ptr<AudioNode> Connect(std::shared_ptr<AudioNode> destination) {
    if (destination.get() == this) throw InvalidOperationException("Cannot connect a node to itself");  // AudioNodeOutput.cs:43-44
    if (std::find(out_.begin(), out_.end(), destination.get()) == out_.end()) {
      out_.push_back(destination.get());
      destination->in_.push_back(this);
    }
    return destination;
  }
  enum class Kind { Destination, Source, Biquad, Gain, Convolver, Delay, Panner };
  virtual Kind NodeKind() const = 0;

 protected:
  friend class OfflineAudioContext;
  std::vector<AudioNode*> in_, out_;  // connection order (AudioNodeInput._connectedOutputs)
};

class AudioDestinationNode : public AudioNode {
 public:
  Kind NodeKind() const override { return Kind::Destination; }
};

class AudioBufferSourceNode : public AudioNode {  // Nodes/AudioBufferSourceNode.cs
 public:
  std::shared_ptr<PlayableAudioBuffer> Buffer;
  AudioParam PlaybackRate{1.f, 0.001f, 1000.f};  // :76
  bool Loop = false;
  double LoopStart = 0, LoopEnd = 0;  // seconds; LoopEnd 0 = end of the buffer (:49-62)
  void Start(double when = 0, double offset = 0, double duration = std::numeric_limits<double>::infinity()) {  // :79-114
    if (started_) throw InvalidOperationException("AudioBufferSourceNode can only be started once.");
    if (!Buffer) throw InvalidOperationException("Cannot start without a buffer set");
    started_ = true;
    when_ = when; offset_ = offset; duration_ = duration;
  }
  void Stop(double when = 0) { stop_ = std::isnan(stop_) ? std::max(0.0, when) : std::min(stop_, std::max(0.0, when)); }  // :116-129
  Kind NodeKind() const override { return Kind::Source; }

 private:
  friend class OfflineAudioContext;
  bool started_ = false;
  double when_ = std::numeric_limits<double>::quiet_NaN(), offset_ = 0, duration_ = std::numeric_limits<double>::infinity();
  double stop_ = std::numeric_limits<double>::quiet_NaN();
};

class BiQuadFilterNode : public AudioNode {  // Nodes/BiQuadFilterNode.cs:54-85
 public:
  explicit BiQuadFilterNode(int sampleRate) : Frequency(1000.f, 1.f, sampleRate / 2.f) {}
  FilterType Type = FilterType::Lowpass;
  AudioParam Frequency;
  AudioParam Q{1.0f, 0.001f, 1000.f};
  AudioParam Gain{0.f, -60.f, 60.f};
  Kind NodeKind() const override { return Kind::Biquad; }
};

class GainNode : public AudioNode {  // Nodes/GainNode.cs:16-25
 public:
  AudioParam Gain{1.0f, std::numeric_limits<float>::lowest(), std::numeric_limits<float>::max()};
  Kind NodeKind() const override { return Kind::Gain; }
};

class DelayNode : public AudioNode {  // Nodes/DelayNode.cs:22-41
 public:
  explicit DelayNode(double maxDelayTime = 1.0) : MaxDelayTime(maxDelayTime), DelayTime(0.f, 0.f, (float)maxDelayTime) {
    if (maxDelayTime <= 0 || maxDelayTime > 10) throw ArgumentOutOfRangeException("maxDelayTime");  // :25-26
  }
  const double MaxDelayTime;
  AudioParam DelayTime;
  Kind NodeKind() const override { return Kind::Delay; }
};

class StereoPannerNode : public AudioNode {  // Nodes/StereoPannerNode.cs:21-34
 public:
  AudioParam Pan{0.f, -1.f, 1.f};
  Kind NodeKind() const override { return Kind::Panner; }
};

class ConvolverNode : public AudioNode {  // Nodes/ConvolverNode.cs
 public:
  explicit ConvolverNode(gac_context* ctx, int sampleRate) : ctx_(ctx), sampleRate_(sampleRate) {}
  bool Normalize = true;         // :87
  bool EnableTrueStereo = true;  // :95
  // ConvolverNode.Buffer = value (:25-79): the convolvers are built here, with the Normalize value of this moment
  void SetBuffer(std::shared_ptr<PlayableAudioBuffer> value) {
    if (value == buffer_) return;
    if (ir_) { gac_ir_destroy(ir_); ir_ = nullptr; }
    buffer_ = value;
    if (!value) return;
    if (value->SampleRate() != sampleRate_)
      throw InvalidOperationException("Impulse response buffer sample rate must match the audio context sample rate.");  // :48-49
    Check(gac_ir_prepare(ctx_, value->Handle(ctx_), Normalize ? 1 : 0, EnableTrueStereo ? 1 : 0, &ir_));
  }
  std::shared_ptr<PlayableAudioBuffer> Buffer() const { return buffer_; }
  ~ConvolverNode() override { if (ir_) gac_ir_destroy(ir_); }
  Kind NodeKind() const override { return Kind::Convolver; }

 private:
  friend class OfflineAudioContext;
  gac_context* ctx_;
  int sampleRate_;
  std::shared_ptr<PlayableAudioBuffer> buffer_;
  gac_ir* ir_ = nullptr;
};

// ---- OfflineAudioContext.cs -----------------------------------------------------------------------------------
class OfflineAudioContext {
 public:
  explicit OfflineAudioContext(int sampleRate = 48000, int partition = 128, int deviceId = -1) : sampleRate_(sampleRate) {
    if (sampleRate <= 0) throw ArgumentOutOfRangeException("sampleRate");  // AudioContextBase.cs:37-38
    gac_context_desc d{};
    d.sample_rate = sampleRate;
    d.quantum = 128;
    d.partition = partition;
    d.device_id = deviceId;
    Check(gac_context_create(&d, &ctx_));
    destination_ = std::make_shared<AudioDestinationNode>();
  }
  ~OfflineAudioContext() {
    nodes_.clear();  // ConvolverNodes release their IR handles before the context goes away
    if (ctx_) gac_context_destroy(ctx_);
  }
  OfflineAudioContext(const OfflineAudioContext&) = delete;
  OfflineAudioContext& operator=(const OfflineAudioContext&) = delete;

  int SampleRate() const { return sampleRate_; }
  std::shared_ptr<AudioNode> Destination() const { return destination_; }
  std::shared_ptr<AudioBufferSourceNode> CreateBufferSource() { return Keep(std::make_shared<AudioBufferSourceNode>()); }
  std::shared_ptr<BiQuadFilterNode> CreateBiQuadFilter() { return Keep(std::make_shared<BiQuadFilterNode>(sampleRate_)); }
  std::shared_ptr<GainNode> CreateGain() { return Keep(std::make_shared<GainNode>()); }
  std::shared_ptr<DelayNode> CreateDelay(double maxDelayTime = 1.0) { return Keep(std::make_shared<DelayNode>(maxDelayTime)); }
  std::shared_ptr<StereoPannerNode> CreateStereoPanner() { return Keep(std::make_shared<StereoPannerNode>()); }
  std::shared_ptr<ConvolverNode> CreateConvolver() { return Keep(std::make_shared<ConvolverNode>(ctx_, sampleRate_)); }

  // Render(float[][] output, int frameCount, int startIndex = 0)  (OfflineAudioContext.cs:30-102)
  void Render(float* const* output, int channels, int frameCount, int startIndex = 0) {
    if (!output || channels <= 0) throw ArgumentException("Output buffer must have at least one channel.");
    if (frameCount <= 0) throw ArgumentOutOfRangeException("Frame count must be positive.");
    if (startIndex < 0) throw ArgumentOutOfRangeException("Start index must be non-negative.");
    Flat f;
    Flatten(f);
    gac_graph* g = nullptr;
    Check(gac_graph_create(ctx_, &f.desc, &g));
    int rc = gac_render(ctx_, g, framesRendered_, frameCount, output, channels, startIndex);
    gac_graph_destroy(g);
    Check(rc);
    framesRendered_ += frameCount;
  }
  // float[][] Render(int frameCount)  (:108-124)
  std::vector<std::vector<float>> Render(int frameCount) {
    if (frameCount <= 0) throw ArgumentOutOfRangeException("Frame count must be positive.");
    std::vector<std::vector<float>> out(2, std::vector<float>((size_t)frameCount));
    float* rows[2] = {out[0].data(), out[1].data()};
    Render(rows, 2, frameCount, 0);
    return out;
  }
  gac_stats LastStats() const {
    gac_stats s{};
    Check(gac_get_stats(ctx_, &s));
    return s;
  }

 private:
  template <typename T>
  std::shared_ptr<T> Keep(std::shared_ptr<T> n) { nodes_.push_back(n); return n; }

  struct Flat {
    std::vector<std::vector<gac_op_desc>> ops;  // backing store for every op list
    std::vector<gac_voice_desc> voices;
    std::vector<gac_bus_desc> buses;
    std::vector<int32_t> dest;
    gac_graph_desc desc{};
  };
  static gac_op_desc OpDesc(AudioNode* n) {
    gac_op_desc o{};
    switch (n->NodeKind()) {
      case AudioNode::Kind::Biquad: {
        auto* b = static_cast<BiQuadFilterNode*>(n);
        o.kind = GAC_OP_BIQUAD;
        o.filter_type = (int32_t)b->Type;
        o.p0 = b->Frequency.Desc(); o.p1 = b->Q.Desc(); o.p2 = b->Gain.Desc();
        break;
      }
      case AudioNode::Kind::Gain: o.kind = GAC_OP_GAIN; o.p0 = static_cast<GainNode*>(n)->Gain.Desc(); break;
      case AudioNode::Kind::Convolver: o.kind = GAC_OP_CONVOLVER; o.ir = static_cast<ConvolverNode*>(n)->ir_; break;
      case AudioNode::Kind::Delay: {
        auto* d = static_cast<DelayNode*>(n);
        o.kind = GAC_OP_DELAY; o.p0 = d->DelayTime.Desc(); o.aux = d->MaxDelayTime;
        break;
      }
      case AudioNode::Kind::Panner: o.kind = GAC_OP_PANNER; o.p0 = static_cast<StereoPannerNode*>(n)->Pan.Desc(); break;
      default: throw NotSupportedException("node type outside the accelerated path");
    }
    return o;
  }
  // node .. upstream through single-input nodes to a source; false if nothing is connected
  static bool WalkVoice(AudioNode* node, AudioBufferSourceNode** src, std::vector<AudioNode*>* ops) {
    ops->clear();
    while (node->NodeKind() != AudioNode::Kind::Source) {
      if (node->out_.size() > 1) throw NotSupportedException("fan-out inside a voice chain is outside the accelerated path");
      ops->push_back(node);
      if (node->in_.empty()) return false;
      if (node->in_.size() > 1) throw NotSupportedException("nested fan-in is outside the accelerated path");
      node = node->in_[0];
    }
    if (node->out_.size() > 1) throw NotSupportedException("a source feeding several nodes is outside the accelerated path");
    std::reverse(ops->begin(), ops->end());
    *src = static_cast<AudioBufferSourceNode*>(node);
    return true;
  }
  void AddVoice(Flat& f, AudioBufferSourceNode* s, const std::vector<AudioNode*>& ops, int bus) {
    f.ops.emplace_back();
    for (auto* n : ops) f.ops.back().push_back(OpDesc(n));
    gac_voice_desc v{};
    v.source = s->Buffer ? s->Buffer->Handle(ctx_) : nullptr;
    v.start_when = (s->started_ && s->Buffer) ? s->when_ : std::numeric_limits<double>::quiet_NaN();
    v.start_offset = s->offset_;
    v.start_duration = s->duration_;
    v.stop_when = s->stop_;
    v.playback_rate = s->PlaybackRate.Value();
    v.source_param = s->PlaybackRate.Desc();  // read when it carries automation events (k-rate, evaluated per quantum on the host)
    v.loop = s->Loop ? 1 : 0;  // any effective rate; an empty loop region answers GAC_ERR_UNSUPPORTED
    v.loop_start = std::max(0.0, s->LoopStart);
    v.loop_end = std::max(0.0, s->LoopEnd);
    v.n_ops = (int32_t)f.ops.back().size();
    v.ops = nullptr;  // patched once every list is in place (vectors may reallocate)
    v.bus = bus;
    f.voices.push_back(v);
    voiceOps_.push_back(f.ops.size() - 1);
  }
  void Flatten(Flat& f) {
    voiceOps_.clear();
    std::vector<size_t> busOps;
    for (AudioNode* head : destination_->in_) {
      std::vector<AudioNode*> chain;
      AudioNode* node = head;
      while (node->NodeKind() != AudioNode::Kind::Source && node->in_.size() == 1) { chain.push_back(node); node = node->in_[0]; }
      if (node->NodeKind() == AudioNode::Kind::Source) {
        std::reverse(chain.begin(), chain.end());
        AddVoice(f, static_cast<AudioBufferSourceNode*>(node), chain, -1);
        f.dest.push_back(~(int32_t)(f.voices.size() - 1));
        continue;
      }
      if (node->in_.empty()) continue;  // nothing connected: contributes silence
      chain.push_back(node);
      std::reverse(chain.begin(), chain.end());
      const int bus = (int)f.buses.size();
      f.ops.emplace_back();
      for (auto* n : chain) f.ops.back().push_back(OpDesc(n));
      busOps.push_back(f.ops.size() - 1);
      f.buses.push_back(gac_bus_desc{(int32_t)chain.size(), nullptr, 0, 0, nullptr});
      f.dest.push_back(bus);
      for (AudioNode* up : node->in_) {
        AudioBufferSourceNode* s = nullptr;
        std::vector<AudioNode*> ops;
        if (WalkVoice(up, &s, &ops)) AddVoice(f, s, ops, bus);
      }
    }
    for (size_t i = 0; i < f.voices.size(); i++) f.voices[i].ops = f.ops[voiceOps_[i]].data();
    for (size_t i = 0; i < f.buses.size(); i++) f.buses[i].ops = f.ops[busOps[i]].data();
    f.desc.n_voices = (int32_t)f.voices.size();
    f.desc.voices = f.voices.data();
    f.desc.n_buses = (int32_t)f.buses.size();
    f.desc.buses = f.buses.data();
    f.desc.n_dest_inputs = (int32_t)f.dest.size();
    f.desc.dest_inputs = f.dest.data();
  }

  int sampleRate_;
  gac_context* ctx_ = nullptr;
  std::shared_ptr<AudioDestinationNode> destination_;
  std::vector<std::shared_ptr<AudioNode>> nodes_;
  std::vector<size_t> voiceOps_;
  int64_t framesRendered_ = 0;
};

}  // namespace Cuda
}  // namespace GraphAudio
