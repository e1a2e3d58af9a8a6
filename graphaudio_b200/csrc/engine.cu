// engine.cu — the C ABI of include/graphaudio_cuda.h: handles, error plumbing, graph flattening and the
// batched render schedule.  The reference renders one 128-frame quantum at a time by pulling through the node
// graph (AudioContextBase.ProcessBlock, AudioContextBase.cs:52-81; AudioNode.ProcessInternal, Nodes/AudioNode.cs:152-183).
// Offline, every input is known up front, so a render here is a short sequence of large kernels over
// [voice, channel, time]:  source -> (automation, biquad, gain)* -> rFFT -> spectral MAC -> irFFT+OLA -> mix.
//
// No CPU fallback exists: if there is no CUDA device the entry points fail with GAC_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <memory>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "../../include/graphaudio_cuda.h"
#include "gac_kernels.h"

using namespace gac;

// ------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CU(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      int code_ = (e_ == cudaErrorMemoryAllocation) ? GAC_ERR_OUT_OF_MEMORY                              \
                  : (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? GAC_ERR_NO_DEVICE   \
                                                                                   : GAC_ERR_CUDA;       \
      return fail(code_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);   \
    }                                                                                                    \
  } while (0)

// ------------------------------------------------------------------------------------------ handles
struct ParamH {
  float value = 0.f;
  std::vector<gac_event> ev;
  // further epochs of a parameter that was edited between successive Render calls (GAC_EVENT_EPOCH): from quantum q0 on the
  // parameter is (value, ev) of that epoch
  struct Epoch {
    int64_t q0;
    float value;
    std::vector<gac_event> ev;
  };
  std::vector<Epoch> later;
  int mod_bus = -1;  // >= 0: the bus whose (mono) output is the parameter's modulation input
  float minv = 0.f, maxv = 0.f;
};
// AudioParam.ComputeValueAtTime (AudioParam.cs:169-217) with InterpolateLinear / InterpolateExponential / ComputeSetTargetFromBaseline
// (:220-247) on the HOST, for the one parameter whose values steer host-side planning: AudioBufferSourceNode.PlaybackRate (k-rate: the
// value at the quantum's start time, :146-165).  `q` selects the epoch (parameters edited between successive Render calls).
static float host_param_value(const ParamH& p, int64_t q, double time) {
  float value = p.value;
  const std::vector<gac_event>* ev = &p.ev;
  for (const auto& e : p.later) {
    if (e.q0 > q) break;
    value = e.value;
    ev = &e.ev;
  }
  const int count = (int)ev->size();
  if (count == 0) return value;
  auto lin = [](float v0, double t0, float v1, double t1, double t) {
    double u = (t - t0) / (t1 - t0);
    u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
    return (float)((double)v0 + (double)(v1 - v0) * u);
  };
  auto set_target = [](const gac_event& e, float base, double t) {
    const double elapsed = t - e.time;
    if (elapsed <= 0) return base;
    const double tc = std::max(e.time_constant, 0.001);
    return (float)((double)e.target + (double)(base - e.target) * std::exp(-elapsed / tc));
  };
  float boundary = value;
  for (int i = 0; i < count; i++) {
    const gac_event& e = (*ev)[i];
    if (time < e.time) {
      if (i == 0) return boundary;
      const gac_event& pv = (*ev)[i - 1];
      if (e.type == GAC_EVENT_LINEAR_RAMP) return lin(pv.value, pv.time, e.value, e.time, time);
      if (e.type == GAC_EVENT_EXPONENTIAL_RAMP) {
        if (pv.value <= 0 || e.value <= 0) return lin(pv.value, pv.time, e.value, e.time, time);
        double u = (time - pv.time) / (e.time - pv.time);
        u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
        return (float)((double)pv.value * std::pow((double)(e.value / pv.value), u));
      }
      if (pv.type == GAC_EVENT_SET_TARGET) return set_target(pv, boundary, time);
      return pv.value;
    }
    if (e.type != GAC_EVENT_SET_TARGET) boundary = e.value;
  }
  const gac_event& last = (*ev)[count - 1];
  return last.type == GAC_EVENT_SET_TARGET ? set_target(last, boundary, time) : last.value;
}
struct OpH {
  int kind = 0;
  int ftype = 0;
  double aux = 0;
  ParamH p0, p1, p2;
  const gac_ir* ir = nullptr;
};
struct VoiceH {
  const gac_buffer* src = nullptr;
  double when = 0, offset = 0, duration = 0, stop_when = 0;
  float rate = 1.f;
  bool loop = false;
  double loop_start = 0, loop_end = 0;
  std::vector<OpH> ops;
  int bus = -1;
  int input_bus = -1;  // >= 0: the chain is fed by that bus's output instead of a source buffer
  int kind = GAC_SOURCE_BUFFER;  // gac_source_kind
  ParamH src_param;              // CONSTANT: Offset, OSCILLATOR: Frequency, BUFFER: PlaybackRate when it carries events (rate_events)
  bool rate_events = false;
  int osc_type = 0;
};
struct BusH {
  std::vector<OpH> ops;
  int target = -1;          // -1 destination, >= 0 parent bus, -2 none (only read by bus-fed chains)
  std::vector<int> inputs;  // connection order at the fan-in: >= 0 bus index, < 0 ~voice index
  std::vector<int> slots;   // per input: 0 ordinary, 1 / 2 = ChannelMergerNode input 0 / 1 (empty: all ordinary)
  bool mono = false;        // GAC_BUS_MONO_INPUT
};

struct NcclApi;

// CubicResampler phase table of one source geometry: (k, t) per output frame (engine_render.inl, plan_sources)
struct ResampleTable {
  std::vector<int32_t> k;
  std::vector<float> t;
  std::vector<int32_t> x;  // looping sources: windows that straddle the loop seam, four source frame indices each (ResampleJob::x)
  int64_t n_active_blocks = 0;
  int64_t n_zero_from = 0;
  int32_t* d_k = nullptr;
  float* d_t = nullptr;
  int32_t* d_x = nullptr;
};

constexpr size_t kResampleCacheMax = 16;  // distinct source geometries whose phase tables a context keeps
struct gac_context {
  uint32_t magic = 0x47414331;  // "GAC1"
  int device = 0;
  int fs = 48000;
  int B = 128;
  int mac_variant = 0;
  int tile_blocks = 32;
  bool mixed_segments = true;  // K6: double-length overlap-save segments in front when that saves work (GAC_FLAG_UNIFORM_SEGMENTS clears it)
  bool fuse_fanin = true;      // convolvers that end their chains and meet in one fan-in are summed as spectra (GAC_FLAG_NO_FANIN_FUSION clears it)
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  float2* d_tw = nullptr;  // e^{-2 pi i k/(2B)}, k < B
  float2* d_tw2 = nullptr; // e^{-2 pi i e/8192}, e < 8192: twiddles of the second-level (block-time) FFT, fft2.cu
  float2* d_tab16 = nullptr; // (points into the d_tw2 allocation) per-M twiddle tables of the radix-16 plan
  std::shared_ptr<struct BlockTimes> bt;  // block start times (shared by the contexts of one device and sample rate)
  double* d_bt = nullptr;                 // = bt->d
  int64_t bt_cap = 0;
  gac_stats stats{};
  // NCCL (optional)
  void* comm = nullptr;
  int rank = 0, n_ranks = 1;
  size_t scratch_budget = (size_t)24 << 30;
  // opt-in asynchronous uploads (GAC_FLAG_ASYNC_UPLOAD): H2D copies run on their own stream and overlap IR
  // preparation and the first voice batches of the next render
  bool async_upload = false;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t d2h_stream = nullptr;  // results of batch renders leave here while the next sub-batch computes (gac_render_batch)
  // ... and IR preparation is DEFERRED: gac_ir_prepare only sizes and allocates, the render that first uses an impulse
  // response prepares it together with all the others its voice batch needs (three launches per batch instead of three per
  // IR; a render never queues behind the preparation of an impulse response whose upload is still in flight)
  std::vector<cudaEvent_t> event_pool;  // recycled `ready` events (creating one costs a driver call per buffer)
  std::map<std::tuple<double, int64_t, int64_t, int64_t, int64_t>, std::shared_ptr<ResampleTable>> resample_cache;  // last: loop start, -1 = no loop
  // page-locked staging for the small job tables of a render: they reach the device through a copy KERNEL (SM loads over
  // PCIe), not through the DMA engine, whose queue may hold hundreds of megabytes of asynchronous buffer uploads — a
  // cudaMemcpyAsync of a 2 KB table would wait behind all of them, and with it every kernel of the render
  // Every context stages this way (not only async-upload ones): a cudaMemcpyAsync from PAGEABLE memory larger than the driver's
  // 64 KB staging slot blocks the host until the stream has drained, which serialised planning and execution of large renders
  // (1024 voices: 57 ms per render of which 25 ms were gaps).  Blocks are taken from the process-wide pool on demand.
  std::vector<char*> stage_blocks;
  size_t stage_block = 0, stage_used = 0;  // current block / bytes used in it
  CopySegments pending;                    // staged tables whose copy to the device has not been launched yet (kstream() flushes)
  std::vector<gac_ir*> deferred_irs;        // async mode: impulse responses waiting for their preparation (kick_deferred_irs / first render)
  int deferred_since_kick = 0;
  cudaEvent_t kick_done = nullptr;         // behind the launches of the last kick: its staged job tables are dead once it has fired
  bool kick_pending = false;
  int64_t copy_launches = 0;               // k_copy_from_host_multi launches so far (counted into gac_stats.kernel_launches per render)
  bool defer_copies = false;               // inside render_core: table copies wait for the next kernel launch and travel together
  // render scratch arena (struct Scratch): device chunks kept between renders, bump-allocated
  std::vector<std::pair<char*, size_t>> arena;
  size_t arena_chunk = 0, arena_used = 0;
  size_t arena_keep_limit = (size_t)64 << 30;
};
// Staging blocks are recycled process-wide: page-locking 8 MB costs milliseconds (and cudaFreeHost synchronises the device),
// far more than the render of a context that lives for one graph.  Blocks are portable (any device's context may take one).
static constexpr size_t kStageBytes = (size_t)8 << 20;
static std::mutex g_stage_mu;
static std::vector<char*> g_stage_free;
static char* take_stage_block() {
  {
    std::lock_guard<std::mutex> lk(g_stage_mu);
    if (!g_stage_free.empty()) {
      char* p = g_stage_free.back();
      g_stage_free.pop_back();
      return p;
    }
  }
  char* p = nullptr;
  if (cudaHostAlloc((void**)&p, kStageBytes, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
static void give_stage_block(char* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_stage_mu);
  g_stage_free.push_back(p);
}
// ---- constant device tables, shared process-wide.
// They are uploaded ONCE per device with a synchronous copy and never through a context's stream: a stream that has used the
// copy engine orders its next operation behind whatever other streams have queued on that engine since, so one small table
// copy at context creation made the whole render wait for the last asynchronous buffer upload (tools/overlap_probe.cu:
// variants 2, 4, 6 against 1, 8, 9).  They are never freed (a few hundred KB per device).
struct BlockTimes {
  std::vector<double> h;  // _currentTime += 128.0 / SampleRate, accumulated block after block (AudioContextBase.cs:78-79)
  double* d = nullptr;
  ~BlockTimes() {
    if (d) cudaFree(d);
  }
};
struct DeviceTables {
  float2* d_tw[3] = {nullptr, nullptr, nullptr};  // partition 128, 256, 512
  float2* d_tw2 = nullptr;
  std::map<int, std::shared_ptr<BlockTimes>> bt;  // by sample rate: the largest table built so far
};
static std::mutex g_tables_mu;
static DeviceTables* device_tables(int dev) {
  static DeviceTables* tabs = new DeviceTables[64];  // (leaked on purpose: no CUDA calls from static destructors)
  return &tabs[dev];
}
static cudaEvent_t take_event(gac_context* ctx) {
  if (!ctx->event_pool.empty()) {
    cudaEvent_t e = ctx->event_pool.back();
    ctx->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  return e;
}
struct gac_buffer {
  gac_context* ctx;
  int nch;
  int64_t n;
  int rate;
  float* d = nullptr;  // [nch][stride]
  int64_t stride;
  cudaEvent_t ready = nullptr;  // async upload: recorded on the copy stream behind the last H2D copy
  int ir_refs = 0;              // impulse responses whose (deferred) preparation still has to read this buffer
  bool zombie = false;          // gac_buffer_destroy was called while ir_refs > 0: freed when the last reference goes
};
// orders the context stream behind a buffer's (possibly still running) upload
static inline void wait_ready(gac_context* ctx, const gac_buffer* b) {
  if (b && b->ready) cudaStreamWaitEvent(ctx->stream, b->ready, 0);
}
struct gac_ir {
  gac_context* ctx;
  int nch;
  bool true_stereo;
  int P, P16;
  int64_t frames;
  float2* d_H = nullptr;  // [nch][P16][B]
  float* d_scale = nullptr;
  // second-level spectra (fft2.cu): [nch][B+1][M2], or null when the context's mac_variant never uses them
  float2* d_H2 = nullptr;  // (inside the d_H allocation)
  int M2 = 0, Lh = 0;
  // second-level spectra for transforms of TWICE the length (own allocation, prepared by the first render that mixes segment
  // lengths: conv_batch_fft2 / plan_segments), [nch][B+1][fft2_h2_row_elems(2 * M2)]
  float2* d_H2b = nullptr;
  // deferred preparation (async mode): the source buffer is read by the first render that uses the impulse response
  bool prepared = true;
  gac_buffer* src = nullptr;
  bool normalize = true;
  // prepared AHEAD of its first render by kick_deferred_irs: that render still counts as the one that prepares it (single-length
  // segment plan; double-length spectra pay for themselves from the second render on)
  bool fresh = false;
};
struct gac_graph {
  gac_context* ctx;
  std::vector<VoiceH> voices;
  std::vector<BusH> buses;
  std::vector<int> dest_inputs;
};

static bool ctx_ok(gac_context* c) { return c && c->magic == 0x47414331; }

// mac_variant: 0 = auto (second-level FFT when the IR has >= kFft2MinP partitions, else register-tiled FFMA2),
// 1 = streaming (reference op order), 2 = register-tiled scalar FFMA, 3 = second-level FFT (forced), 4 = register-tiled FFMA2 (forced)
constexpr int kFft2MinP = 64;
static bool wants_fft2(const gac_context* c) { return c->mac_variant == 0 || c->mac_variant == 3; }
static bool use_fft2(const gac_context* c, int P, int M2) {
  if (M2 <= 0) return false;
  return c->mac_variant == 3 || (c->mac_variant == 0 && P >= kFft2MinP);
}

// ------------------------------------------------------------------------------------------ scratch + timing
// Stream-ordered scratch allocations, all released at the end of a render.
// Host table -> device, ordered on the context stream.  With asynchronous uploads the table goes through the context's
// page-locked staging area and a copy kernel (see gac_context::h_stage); the staging area is recycled by every render
// (renders synchronise before they return).  The destination must be allocated in multiples of 16 bytes.
static void flush_copies(gac_context* ctx) {
  if (ctx->pending.n == 0) return;
  ctx->copy_launches++;
  launch_copy_from_host_multi(ctx->pending, ctx->stream);
  ctx->pending.n = 0;
}
// the stream kernels are launched on, with every table they may read on its way
static inline cudaStream_t kstream(gac_context* ctx) {
  flush_copies(ctx);
  return ctx->stream;
}
static int table_h2d(gac_context* ctx, void* d_dst, const void* h_src, size_t bytes) {
  if (bytes == 0) return GAC_OK;
  const size_t need = (bytes + 15) & ~(size_t)15;
  if (need <= kStageBytes) {
    if (ctx->stage_block < ctx->stage_blocks.size() && ctx->stage_used + need > kStageBytes) {
      ctx->stage_block++;
      ctx->stage_used = 0;
    }
    if (ctx->stage_block >= ctx->stage_blocks.size()) {
      if (char* p = take_stage_block()) {
        ctx->stage_blocks.push_back(p);
        ctx->stage_block = ctx->stage_blocks.size() - 1;
        ctx->stage_used = 0;
      }
    }
    if (ctx->stage_block < ctx->stage_blocks.size()) {
      char* slot = ctx->stage_blocks[ctx->stage_block] + ctx->stage_used;
      ctx->stage_used += need;
      memcpy(slot, h_src, bytes);
      // the copy is deferred: every kernel launch on the context stream goes through kstream(), which sends the pending tables
      // with one launch first
      if (ctx->pending.n == kCopySegs) flush_copies(ctx);
      CopySegments& pc = ctx->pending;
      pc.dst[pc.n] = d_dst;
      pc.src[pc.n] = slot;
      pc.n16[pc.n] = need / 16;
      pc.n++;
      if (!ctx->defer_copies) flush_copies(ctx);
      return GAC_OK;
    }
  }
  flush_copies(ctx);  // (keeps the order of writes to a destination that is uploaded twice)
  cudaError_t e = cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) return fail(GAC_ERR_CUDA, "cudaMemcpyAsync failed: %s", cudaGetErrorString(e));
  return GAC_OK;
}

// Render scratch comes from a context-owned ARENA: chunks of device memory that stay with the context between renders, handed out
// by bumping a pointer.  The first render of a context grows the arena chunk by chunk (stream-ordered allocations from the device
// pool); the second render replaces several chunks by ONE of their total size, so that from then on a render makes no allocation
// call at all.  (Allocating and freeing ~20 blocks of up to 24 GB per render through cudaMallocAsync was correct but not
// steady: whenever a long-lived allocation — an impulse response's double-length spectra — was carved out of the pool's cached
// blocks in between, the next render paid hundreds of milliseconds of re-mapping: 1024 voices rendered in 33 ms, 164 ms, 477 ms,
// 33 ms ... on consecutive calls.)  Arenas beyond `arena_keep_limit` are returned to the pool when the render ends.
struct Scratch {
  gac_context* ctx;
  size_t bytes = 0;
  explicit Scratch(gac_context* c) : ctx(c) {
    ctx->arena_chunk = 0;
    ctx->arena_used = 0;
    // several chunks left by the previous render: one chunk of their total size replaces them NOW (not when that render ended: a
    // context that renders once and is disposed — one OfflineAudioContext per render is the reference's usual pattern — must not
    // pay for an arena it never uses again)
    auto& chunks = ctx->arena;
    if (chunks.size() > 1) {
      size_t total = 0;
      for (auto& ch : chunks) {
        total += ch.second;
        cudaFreeAsync(ch.first, ctx->stream);
      }
      chunks.clear();
      void* p = nullptr;
      if (cudaMallocAsync(&p, total, ctx->stream) == cudaSuccess) chunks.emplace_back((char*)p, total);
      else cudaGetLastError();  // (the render grows the arena chunk by chunk again)
    }
  }
  template <typename T>
  int alloc(T** out, size_t count) {
    const size_t b = (std::max<size_t>(count * sizeof(T), 16) + 255) & ~(size_t)255;
    auto& chunks = ctx->arena;
    while (ctx->arena_chunk < chunks.size() && ctx->arena_used + b > chunks[ctx->arena_chunk].second) {
      ctx->arena_chunk++;
      ctx->arena_used = 0;
    }
    if (ctx->arena_chunk >= chunks.size()) {
      void* p = nullptr;
      const size_t want = std::max(b, (size_t)64 << 20);
      cudaError_t e = cudaMallocAsync(&p, want, ctx->stream);
      if (e != cudaSuccess) {
        return fail(e == cudaErrorMemoryAllocation ? GAC_ERR_OUT_OF_MEMORY : GAC_ERR_CUDA, "cudaMallocAsync(%zu bytes) failed: %s", want,
                    cudaGetErrorString(e));
      }
      chunks.emplace_back((char*)p, want);
      ctx->arena_chunk = chunks.size() - 1;
      ctx->arena_used = 0;
    }
    *out = (T*)(chunks[ctx->arena_chunk].first + ctx->arena_used);
    ctx->arena_used += b;
    bytes += b;
    return GAC_OK;
  }
  // upload a host vector (stream-ordered; the vector must stay alive until the stream is synchronised)
  template <typename T>
  int upload(T** out, const std::vector<T>& v) {
    int rc = alloc(out, v.size());
    if (rc) return rc;
    if (v.empty()) return GAC_OK;
    return table_h2d(ctx, *out, v.data(), v.size() * sizeof(T));  // (the allocation is 16-byte granular)
  }
  // end of a render: oversized arenas go back to the pool
  void release() {
    auto& chunks = ctx->arena;
    size_t total = 0;
    for (auto& c : chunks) total += c.second;
    if (total <= ctx->arena_keep_limit) return;
    for (auto& c : chunks) cudaFreeAsync(c.first, ctx->stream);
    chunks.clear();
  }
  ~Scratch() { release(); }
};

enum Cat { C_SOURCE, C_AUTO, C_BIQUAD, C_GAIN, C_FFT_FWD, C_MAC, C_FFT_INV, C_MIX, C_D2H, C_DELAY, C_PANNER, C_COUNT };
struct Timer {
  gac_context* ctx;
  struct Span {
    cudaEvent_t a, b;
    int cat;
  };
  std::vector<Span> spans;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  explicit Timer(gac_context* c) : ctx(c) {
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    cudaEventRecord(t0, ctx->stream);
  }
  int begin(int cat) {
    Span s;
    cudaEventCreate(&s.a);
    cudaEventCreate(&s.b);
    s.cat = cat;
    cudaEventRecord(s.a, ctx->stream);
    spans.push_back(s);
    return (int)spans.size() - 1;
  }
  void end(int id) { cudaEventRecord(spans[id].b, ctx->stream); }
  void finish(gac_stats* st) {
    cudaEventRecord(t1, ctx->stream);
    cudaEventSynchronize(t1);
    double acc[C_COUNT] = {0};
    for (auto& s : spans) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, s.a, s.b);
      acc[s.cat] += ms;
    }
    float tot = 0.f;
    cudaEventElapsedTime(&tot, t0, t1);
    st->ms_total = tot;
    st->ms_source = acc[C_SOURCE];
    st->ms_automation = acc[C_AUTO];
    st->ms_biquad = acc[C_BIQUAD];
    st->ms_gain = acc[C_GAIN];
    st->ms_fft_fwd = acc[C_FFT_FWD];
    st->ms_mac = acc[C_MAC];
    st->ms_fft_inv = acc[C_FFT_INV];
    st->ms_mix = acc[C_MIX];
    st->ms_d2h = acc[C_D2H];
    st->ms_delay = acc[C_DELAY];
    st->ms_panner = acc[C_PANNER];
  }
  ~Timer() {
    for (auto& s : spans) {
      cudaEventDestroy(s.a);
      cudaEventDestroy(s.b);
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
  }
};

// keeps host-side job arrays alive until the stream has consumed them
struct HostKeep {
  std::vector<std::shared_ptr<void>> items;
  template <typename T>
  std::vector<T>& make() {
    auto p = std::make_shared<std::vector<T>>();
    items.push_back(p);
    return *p;
  }
};

// GAC_TRACE=1: host-side timestamps of a render's phases on stderr (diagnostics only)
struct HostTrace {
  bool on;
  std::chrono::steady_clock::time_point t0;
  HostTrace() : on(getenv("GAC_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
  void mark(const char* what) {
    if (!on) return;
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[gac_trace] %-28s +%.3f ms\n", what, ms);
  }
};
static thread_local HostTrace* g_trace = nullptr;
#define TRACE_MARK(what)              \
  do {                                \
    if (g_trace) g_trace->mark(what); \
  } while (0)

// ------------------------------------------------------------------------------------------ library
extern "C" int gac_version(void) { return GAC_ABI_VERSION; }
extern "C" const char* gac_last_error(void) { return g_err; }
extern "C" int gac_device_count(int* count) {
  if (!count) return fail(GAC_ERR_INVALID_ARGUMENT, "count is null");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(GAC_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  }
  *count = n;
  return GAC_OK;
}

static int ensure_block_times(gac_context* ctx, int64_t nq) {
  if (ctx->bt && (int64_t)ctx->bt->h.size() >= nq + 1) return GAC_OK;
  std::lock_guard<std::mutex> lk(g_tables_mu);
  std::shared_ptr<BlockTimes>& cached = device_tables(ctx->device)->bt[ctx->fs];
  if (!cached || (int64_t)cached->h.size() < nq + 1) {
    auto t = std::make_shared<BlockTimes>();
    if (cached) t->h = cached->h;  // (the accumulation continues where the smaller table stopped)
    const int64_t want = std::max<int64_t>(nq + 1, 8192);
    const double inc = (double)128 / (double)ctx->fs;
    double acc = t->h.empty() ? 0.0 : t->h.back();
    if (t->h.empty()) t->h.push_back(0.0);
    while ((int64_t)t->h.size() < want) {
      acc = acc + inc;
      t->h.push_back(acc);
    }
    CU(cudaMalloc(&t->d, t->h.size() * sizeof(double)));
    CU(cudaMemcpy(t->d, t->h.data(), t->h.size() * sizeof(double), cudaMemcpyHostToDevice));
    cached = t;  // contexts that still hold the smaller table keep it alive
  }
  ctx->bt = cached;
  ctx->d_bt = cached->d;
  ctx->bt_cap = (int64_t)cached->h.size();
  return GAC_OK;
}

extern "C" int gac_context_create(const gac_context_desc* desc, gac_context** out) {
  if (!desc || !out) return fail(GAC_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  if (desc->sample_rate <= 0) return fail(GAC_ERR_OUT_OF_RANGE, "sampleRate must be positive");  // AudioContextBase.cs:37-38
  if (desc->quantum != 0 && desc->quantum != 128) return fail(GAC_ERR_INVALID_ARGUMENT, "quantum must be 128 (AudioBuffer.FramesPerBlock)");
  int B = desc->partition == 0 ? 128 : desc->partition;
  if (B != 128 && B != 256 && B != 512) return fail(GAC_ERR_INVALID_ARGUMENT, "partition must be 128, 256 or 512");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(GAC_ERR_NO_DEVICE, "no CUDA device (%s); libgraphaudio_cuda has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
  int dev = desc->device_id;
  if (dev < 0) CU(cudaGetDevice(&dev));
  if (dev >= ndev) return fail(GAC_ERR_OUT_OF_RANGE, "device_id %d out of range (%d devices)", dev, ndev);
  // per-device facts are queried once per process (cudaGetDeviceProperties / cudaMemGetInfo cost milliseconds)
  struct DevInfo {
    bool known = false;
    int major = 0, minor = 0, sms = 148;
    size_t budget = (size_t)24 << 30;
  };
  static DevInfo dev_info[64];
  if (dev >= 64) return fail(GAC_ERR_OUT_OF_RANGE, "device_id %d not supported", dev);
  DevInfo& di = dev_info[dev];
  CU(cudaSetDevice(dev));
  std::unique_lock<std::mutex> dev_lock(g_tables_mu);  // contexts may be created from several threads at once
  if (!di.known) {
    CU(cudaDeviceGetAttribute(&di.major, cudaDevAttrComputeCapabilityMajor, dev));
    CU(cudaDeviceGetAttribute(&di.minor, cudaDevAttrComputeCapabilityMinor, dev));
    CU(cudaDeviceGetAttribute(&di.sms, cudaDevAttrMultiProcessorCount, dev));
    // keep freed scratch in the pool between renders
    cudaMemPool_t pool;
    CU(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t thr = UINT64_MAX;
    CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) di.budget = std::max<size_t>((size_t)1 << 30, free_b / 3);
    di.known = true;
  }
  dev_lock.unlock();
  if (di.major != 10) return fail(GAC_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, di.major, di.minor);
  auto ctx = std::make_unique<gac_context>();
  ctx->device = dev;
  ctx->fs = desc->sample_rate;
  ctx->B = B;
  ctx->mac_variant = desc->mac_variant;
  ctx->tile_blocks = desc->tile_blocks == 64 ? 64 : 32;
  CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ctx->async_upload = (desc->flags & GAC_FLAG_ASYNC_UPLOAD) != 0;
  ctx->mixed_segments = (desc->flags & GAC_FLAG_UNIFORM_SEGMENTS) == 0;
  ctx->fuse_fanin = (desc->flags & GAC_FLAG_NO_FANIN_FUSION) == 0 && !getenv("GAC_NO_FANIN_FUSION");
  if (ctx->async_upload) CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  ctx->scratch_budget = di.budget;
  ctx->sm_count = di.sms > 0 ? di.sms : 148;
  ctx->arena_keep_limit = di.budget + di.budget / 2;  // half of the memory that was free when the device was first used
  {
    std::lock_guard<std::mutex> lk(g_tables_mu);
    DeviceTables* T = device_tables(dev);
    const double pi = 3.14159265358979323846;
    const int bi = B == 128 ? 0 : B == 256 ? 1 : 2;
    if (!T->d_tw[bi]) {
      // twiddles e^{-2 pi i k / N}, N = 2B, in double then rounded once
      std::vector<float2> tw(B);
      for (int k = 0; k < B; k++) {
        double a = -2.0 * pi * (double)k / (double)(2 * B);
        tw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
      }
      CU(cudaMalloc(&T->d_tw[bi], sizeof(float2) * B));
      CU(cudaMemcpy(T->d_tw[bi], tw.data(), sizeof(float2) * B, cudaMemcpyHostToDevice));
    }
    if (!T->d_tw2) {
      std::vector<float2> tw2(kFft2TwLen + fft2_table_total());
      for (int e = 0; e < kFft2TwLen; e++) {
        double a = -2.0 * pi * (double)e / (double)kFft2TwLen;
        tw2[e] = make_float2((float)std::cos(a), (float)std::sin(a));
      }
      fft2_fill_tables(tw2.data() + kFft2TwLen);
      CU(cudaMalloc(&T->d_tw2, sizeof(float2) * tw2.size()));
      CU(cudaMemcpy(T->d_tw2, tw2.data(), sizeof(float2) * tw2.size(), cudaMemcpyHostToDevice));
    }
    ctx->d_tw = T->d_tw[bi];
    ctx->d_tw2 = T->d_tw2;
    ctx->d_tab16 = ctx->d_tw2 + kFft2TwLen;
  }
  int rc = ensure_block_times(ctx.get(), 8192);
  if (rc) return rc;
  *out = ctx.release();
  return GAC_OK;
}

extern "C" int gac_synchronize(gac_context* ctx) {
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or destroyed");
  CU(cudaSetDevice(ctx->device));
  if (ctx->copy_stream) CU(cudaStreamSynchronize(ctx->copy_stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return GAC_OK;
}

extern "C" int gac_comm_destroy(gac_context* ctx);
extern "C" int gac_context_destroy(gac_context* ctx) {
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or already destroyed");
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->comm) gac_comm_destroy(ctx);
  for (auto& c : ctx->arena) cudaFreeAsync(c.first, ctx->stream);
  ctx->arena.clear();
  ctx->deferred_irs.clear();
  if (ctx->kick_done) cudaEventDestroy(ctx->kick_done);
  ctx->kick_done = nullptr;
  for (auto& kv : ctx->resample_cache) {
    if (kv.second->d_k) cudaFreeAsync(kv.second->d_k, ctx->stream);
    if (kv.second->d_t) cudaFreeAsync(kv.second->d_t, ctx->stream);
    if (kv.second->d_x) cudaFreeAsync(kv.second->d_x, ctx->stream);
  }
  cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
  }
  if (ctx->d2h_stream) {
    cudaStreamSynchronize(ctx->d2h_stream);
    cudaStreamDestroy(ctx->d2h_stream);
  }
  for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
  for (char* p : ctx->stage_blocks) give_stage_block(p);
  cudaStreamDestroy(ctx->stream);
  ctx->magic = 0;
  delete ctx;
  return GAC_OK;
}

// ------------------------------------------------------------------------------------------ buffers
extern "C" int gac_buffer_create(gac_context* ctx, const float* const* channels, int n_channels, int64_t n_frames, int sample_rate,
                                 gac_buffer** out) {
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or destroyed");
  if (!channels || !out) return fail(GAC_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  if (n_channels < 1 || n_channels > 32) return fail(GAC_ERR_OUT_OF_RANGE, "Channel count must be between 1 and 32");  // PlayableAudioBuffer.cs:47-48
  if (n_frames < 0) return fail(GAC_ERR_OUT_OF_RANGE, "Length must be non-negative");
  if (sample_rate <= 0) return fail(GAC_ERR_OUT_OF_RANGE, "Sample rate must be positive");
  for (int c = 0; c < n_channels; c++)
    if (!channels[c]) return fail(GAC_ERR_INVALID_ARGUMENT, "channel %d is null", c);
  CU(cudaSetDevice(ctx->device));
  auto b = std::make_unique<gac_buffer>();
  b->ctx = ctx;
  b->nch = n_channels;
  b->n = n_frames;
  b->rate = sample_rate;
  // Channel arrays that are rows of one host block (a pinned staging buffer) are uploaded with ONE 1-D copy into rows of the
  // same pitch: a strided 2-D copy reaches 41 GB/s on this box, one copy per row 46, one contiguous copy 52-55
  // (tools/h2d_bandwidth.py).  The device rows keep 16-byte alignment (float2 / float4 loads) only if the pitch is a multiple
  // of 4 frames; otherwise the rows get their own padded pitch and one copy each.
  bool contiguous = n_channels > 1 && n_frames > 0 && n_frames % 4 == 0;
  for (int c = 1; c < n_channels && contiguous; c++) contiguous = channels[c] == channels[c - 1] + n_frames;
  b->stride = contiguous ? n_frames : ((n_frames + 8 + 63) / 64) * 64;
  // Asynchronous path (opt-in): the copy is queued on the copy stream and the call returns; the arrays must stay valid and
  // unmodified until the next gac_render* / gac_synchronize returns.  (Pageable arrays are still correct: the runtime stages
  // them before cudaMemcpyAsync returns.)
  const bool async = ctx->async_upload;
  cudaStream_t st = async ? ctx->copy_stream : ctx->stream;
  // a little slack at the end so that 4-tap reads never leave the allocation; the frames behind a channel's last one are never
  // USED: k_source_copy stays inside [0, n) and the resampler's taps are the last four CONSUMED frames (k + 3 <= n - 1)
  CU(cudaMallocAsync(&b->d, sizeof(float) * ((size_t)b->stride * n_channels + 64), st));
  if (contiguous) {
    CU(cudaMemcpyAsync(b->d, channels[0], sizeof(float) * (size_t)n_frames * n_channels, cudaMemcpyHostToDevice, st));
  } else {
    for (int c = 0; c < n_channels; c++)
      CU(cudaMemcpyAsync(b->d + c * b->stride, channels[c], sizeof(float) * n_frames, cudaMemcpyHostToDevice, st));
  }
  if (async) {
    b->ready = take_event(ctx);
    CU(cudaEventRecord(b->ready, st));
  } else {
    CU(cudaStreamSynchronize(st));  // the caller may reuse its arrays as soon as we return
  }
  *out = b.release();
  return GAC_OK;
}
// ≙ AudioDecoder.LoadFromStream (GraphAudio.IO/LibsndfileDecoder.cs:195-220) behind the container parser: the interleaved samples
// of the file go to the device as they are (half the PCIe bytes for 16-bit material) and are converted and de-interleaved there
extern "C" int gac_buffer_create_interleaved(gac_context* ctx, const void* samples, int sample_format, int n_channels, int64_t n_frames,
                                             int sample_rate, gac_buffer** out) {
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or destroyed");
  if (!samples || !out) return fail(GAC_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  if (n_channels < 1 || n_channels > 32) return fail(GAC_ERR_OUT_OF_RANGE, "Channel count must be between 1 and 32");
  if (n_frames < 0) return fail(GAC_ERR_OUT_OF_RANGE, "Length must be non-negative");
  if (sample_rate <= 0) return fail(GAC_ERR_OUT_OF_RANGE, "Sample rate must be positive");
  if (sample_format < GAC_SAMPLE_S16 || sample_format > GAC_SAMPLE_F32) return fail(GAC_ERR_INVALID_ARGUMENT, "unknown sample format %d", sample_format);
  static const int bytes_of[4] = {2, 3, 4, 4};
  CU(cudaSetDevice(ctx->device));
  auto b = std::make_unique<gac_buffer>();
  b->ctx = ctx;
  b->nch = n_channels;
  b->n = n_frames;
  b->rate = sample_rate;
  b->stride = ((n_frames + 8 + 63) / 64) * 64;
  const bool async = ctx->async_upload;
  cudaStream_t st = async ? ctx->copy_stream : ctx->stream;
  const size_t raw_bytes = (size_t)n_frames * n_channels * bytes_of[sample_format];
  void* d_raw = nullptr;
  CU(cudaMallocAsync(&b->d, sizeof(float) * ((size_t)b->stride * n_channels + 64), st));
  CU(cudaMallocAsync(&d_raw, std::max<size_t>(raw_bytes, 16), st));
  if (raw_bytes) CU(cudaMemcpyAsync(d_raw, samples, raw_bytes, cudaMemcpyHostToDevice, st));
  launch_deinterleave(d_raw, sample_format, n_channels, n_frames, b->d, b->stride, st);
  CU(cudaGetLastError());
  CU(cudaFreeAsync(d_raw, st));
  if (async) {
    b->ready = take_event(ctx);
    CU(cudaEventRecord(b->ready, st));
  } else {
    CU(cudaStreamSynchronize(st));
  }
  *out = b.release();
  return GAC_OK;
}
extern "C" int gac_buffer_destroy(gac_buffer* buf) {
  if (!buf) return fail(GAC_ERR_INVALID_ARGUMENT, "buffer is null");
  if (buf->ir_refs > 0) {  // an impulse response prepared from it has not been used yet: it goes when that one is prepared / destroyed
    buf->zombie = true;
    return GAC_OK;
  }
  cudaSetDevice(buf->ctx->device);
  // stream-ordered free: safe behind any render still queued on the context stream
  if (buf->ready) {
    cudaStreamWaitEvent(buf->ctx->stream, buf->ready, 0);
    buf->ctx->event_pool.push_back(buf->ready);
  }
  cudaFreeAsync(buf->d, buf->ctx->stream);
  delete buf;
  return GAC_OK;
}

// ------------------------------------------------------------------------------------------ IR prepare (K0)
// sizes and allocates the spectra of an impulse response (no kernel is launched)
static int ir_allocate(gac_context* ctx, int nch, int64_t frames, gac_ir* ir) {
  const int B = ctx->B;
  ir->P = (int)((frames + B - 1) / B);  // ceil(L / blockSize)  PartitionedConvolver.cs:44
  ir->P16 = std::max(16, ((ir->P + 15) / 16) * 16);
  ir->M2 = wants_fft2(ctx) ? fft2_pick_m(ir->P, &ir->Lh) : 0;
  if (ir->M2 > 0 && !use_fft2(ctx, ir->P, ir->M2)) ir->M2 = 0;
  // one stream-ordered allocation: spectra [nch][P16][B] float2, the second-level spectra, the per-channel scales
  const size_t hbytes = sizeof(float2) * (size_t)nch * ir->P16 * B;
  const size_t h2bytes = ir->M2 > 0 ? sizeof(float2) * (size_t)nch * (B + 1) * fft2_h2_row_elems(ir->M2) : 0;
  CU(cudaMallocAsync(&ir->d_H, hbytes + h2bytes + sizeof(float) * nch, ctx->stream));
  ir->d_H2 = h2bytes ? reinterpret_cast<float2*>(reinterpret_cast<char*>(ir->d_H) + hbytes) : nullptr;
  ir->d_scale = reinterpret_cast<float*>(reinterpret_cast<char*>(ir->d_H) + hbytes + h2bytes);
  return GAC_OK;
}

// prepares one ALLOCATED impulse response on the context stream (PartitionedConvolver.cs:65-102 per channel)
static int ir_prepare_run(gac_context* ctx, const float* d_ir, int64_t stride, int nch, int64_t frames, bool normalize, gac_ir* ir) {
  const int B = ctx->B;
  cudaStream_t st = ctx->stream;
  // rows P .. P16 stay zero (the register-tiled MAC reads whole 16-row chunks); rows < P are written by the transform
  if (ir->P16 > ir->P)
    CU(cudaMemset2DAsync(ir->d_H + (size_t)ir->P * B, sizeof(float2) * (size_t)ir->P16 * B, 0, sizeof(float2) * (size_t)(ir->P16 - ir->P) * B, nch, st));
  // (float)Math.Pow(10, GainCalibration * 0.05f) with GainCalibration = -58  (PartitionedConvolver.cs:95,101)
  const float cal = (float)std::pow(10.0, (double)(-58.f * 0.05f));
  launch_ir_scale(d_ir, stride, nch, frames, normalize && frames > 0 ? 1 : 0, cal, ir->d_scale, st);
  FftFwdUniform u;
  u.in_base = d_ir;
  u.in_stride = stride;
  u.out_base = ir->d_H;
  u.out_stride = (int64_t)ir->P16 * B;
  u.scale_base = ir->d_scale;
  u.n_valid = frames;
  u.n_blocks = ir->P;
  launch_rfft_fwd_uniform(u, nch, B, ctx->d_tw, st);
  CU(cudaGetLastError());
  // second-level spectra: FFT of every bin's partition sequence along p (fft2.cu)
  if (ir->M2 > 0) {
    launch_fft2_prep(ir->d_H, (int64_t)ir->P16 * B, nch, B, ir->P, ir->M2, ir->d_H2, ctx->d_tw2, ctx->d_tab16, st);
    CU(cudaGetLastError());
  }
  // no host synchronisation: every later use of the spectra is ordered behind the same stream
  ir->prepared = true;
  return GAC_OK;
}
// allocates and prepares at once (the reference's semantics: ConvolverNode.Buffer = ir does the work)
static int ir_prepare_device(gac_context* ctx, const float* d_ir, int64_t stride, int nch, int64_t frames, bool normalize, gac_ir* ir) {
  int rc = ir_allocate(ctx, nch, frames, ir);
  if (rc) return rc;
  return ir_prepare_run(ctx, d_ir, stride, nch, frames, normalize, ir);
}

static void buffer_unref(gac_buffer* b) {
  if (b && --b->ir_refs == 0 && b->zombie) {
    b->zombie = false;
    gac_buffer_destroy(b);
  }
}

static void kick_deferred_irs(gac_context* ctx, size_t min_ready);
extern "C" int gac_ir_prepare(gac_context* ctx, const gac_buffer* buf, int normalize, int true_stereo, gac_ir** out) {
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or destroyed");
  if (!buf || !out) return fail(GAC_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  if (buf->ctx != ctx) return fail(GAC_ERR_INVALID_ARGUMENT, "buffer belongs to another context");
  if (buf->rate != ctx->fs)  // Nodes/ConvolverNode.cs:48-49
    return fail(GAC_ERR_INVALID_OPERATION,
                "Impulse response buffer sample rate must match the audio context sample rate. Impulse response buffer sample rate: %d, "
                "Audio context sample rate: %d.",
                buf->rate, ctx->fs);
  bool ts = (buf->nch == 4 && true_stereo);
  if (!(buf->nch == 1 || buf->nch == 2 || ts))
    return fail(GAC_ERR_UNSUPPORTED, "impulse responses with %d discrete channels are outside the accelerated path (1, 2, or 4 with true stereo)", buf->nch);
  if (buf->n <= 0) return fail(GAC_ERR_INVALID_ARGUMENT, "impulse response is empty");
  CU(cudaSetDevice(ctx->device));
  auto ir = std::make_unique<gac_ir>();
  ir->ctx = ctx;
  ir->nch = buf->nch;
  ir->true_stereo = ts;
  ir->frames = buf->n;
  int rc;
  if (ctx->async_upload && buf->ready) {
    // deferred: the render that first uses the impulse response prepares it, batched with the others (prepare_irs)
    rc = ir_allocate(ctx, buf->nch, buf->n, ir.get());
    ir->prepared = false;
    ir->src = const_cast<gac_buffer*>(buf);
    ir->normalize = normalize != 0;
    if (!rc) {
      ir->src->ir_refs++;
      ctx->deferred_irs.push_back(ir.get());
    }
  } else {
    wait_ready(ctx, buf);
    rc = ir_prepare_device(ctx, buf->d, buf->stride, buf->nch, buf->n, normalize != 0, ir.get());
  }
  if (rc) {
    if (ir->d_H) cudaFreeAsync(ir->d_H, ctx->stream);
    return rc;
  }
  *out = ir.release();
  // every 32 deferred impulse responses: those whose samples have landed are prepared now, in one batch, while the copy engine
  // delivers the rest and the host builds the graph — the first render then finds them ready
  if (!(*out)->prepared && ++ctx->deferred_since_kick >= 32) {
    ctx->deferred_since_kick = 0;
    kick_deferred_irs(ctx, 16);
  }
  return GAC_OK;
}
extern "C" int gac_ir_destroy(gac_ir* ir) {
  if (!ir) return fail(GAC_ERR_INVALID_ARGUMENT, "ir is null");
  cudaSetDevice(ir->ctx->device);
  {
    auto& L = ir->ctx->deferred_irs;
    L.erase(std::remove(L.begin(), L.end(), ir), L.end());
  }
  // stream-ordered free, behind any render still queued on the context stream
  if (!ir->prepared) buffer_unref(ir->src);  // never used: its source buffer is released too
  if (ir->d_H2b) cudaFreeAsync(ir->d_H2b, ir->ctx->stream);
  cudaFreeAsync(ir->d_H, ir->ctx->stream);
  delete ir;
  return GAC_OK;
}

// ------------------------------------------------------------------------------------------ graph
static int copy_param(const gac_param& p, ParamH* out, const char* what) {
  out->value = p.value;
  out->mod_bus = p.mod_bus > 0 ? p.mod_bus - 1 : -1;
  out->minv = p.min_value;
  out->maxv = p.max_value;
  if (p.mod_bus < 0) return fail(GAC_ERR_OUT_OF_RANGE, "%s: negative modulation bus", what);
  if (p.mod_bus > 0 && !(p.min_value <= p.max_value)) return fail(GAC_ERR_INVALID_ARGUMENT, "%s: a modulated parameter needs its [min, max] range", what);
  if (p.n_events < 0) return fail(GAC_ERR_INVALID_ARGUMENT, "%s: negative event count", what);
  if (p.n_events > 0 && !p.events) return fail(GAC_ERR_INVALID_ARGUMENT, "%s: events is null", what);
  out->ev.clear();
  out->later.clear();
  std::vector<gac_event>* cur = &out->ev;
  for (int i = 0; i < p.n_events; i++) {
    const gac_event& e = p.events[i];
    if (e.type == GAC_EVENT_EPOCH) {
      const int64_t q0 = (int64_t)e.time_constant;
      const int64_t prev = out->later.empty() ? 0 : out->later.back().q0;
      if (!(e.time_constant >= 1.0) || q0 <= prev) return fail(GAC_ERR_INVALID_ARGUMENT, "%s: epoch markers must carry increasing quantum indices >= 1", what);
      out->later.push_back({q0, e.value, {}});
      cur = &out->later.back().ev;
      continue;
    }
    if (e.type < 0 || e.type > 3) return fail(GAC_ERR_INVALID_ARGUMENT, "%s: bad event type", what);
    if (!cur->empty() && e.time < cur->back().time) return fail(GAC_ERR_INVALID_ARGUMENT, "%s: events must be sorted by time (AudioParam.AddEvent order)", what);
    cur->push_back(e);
  }
  return GAC_OK;
}
static int copy_ops(gac_context* ctx, int n, const gac_op_desc* ops, std::vector<OpH>* out) {
  if (n < 0 || (n > 0 && !ops)) return fail(GAC_ERR_INVALID_ARGUMENT, "bad op list");
  out->resize(n);
  for (int i = 0; i < n; i++) {
    OpH& o = (*out)[i];
    o.kind = ops[i].kind;
    o.ftype = ops[i].filter_type;
    int rc;
    switch (o.kind) {
      case GAC_OP_BIQUAD:
        if (o.ftype < 0 || o.ftype > 7) return fail(GAC_ERR_INVALID_ARGUMENT, "bad filter type %d", o.ftype);
        if ((rc = copy_param(ops[i].p0, &o.p0, "biquad.frequency"))) return rc;
        if ((rc = copy_param(ops[i].p1, &o.p1, "biquad.Q"))) return rc;
        if ((rc = copy_param(ops[i].p2, &o.p2, "biquad.gain"))) return rc;
        break;
      case GAC_OP_GAIN:
        if ((rc = copy_param(ops[i].p0, &o.p0, "gain.gain"))) return rc;
        break;
      case GAC_OP_GATE:
        o.aux = ops[i].aux;
        if (o.ftype < 0 || o.ftype > 1 || !(o.aux >= 0.0)) return fail(GAC_ERR_INVALID_ARGUMENT, "bad connection window");
        break;
      case GAC_OP_CONVOLVER:
        o.aux = ops[i].aux;
        if (o.ftype == 1) {
          if (i == 0 || (*out)[i - 1].kind != GAC_OP_CONVOLVER) return fail(GAC_ERR_INVALID_ARGUMENT, "a convolver epoch must follow the convolver op it continues");
          if (!(o.aux >= 1.0) || o.aux <= (*out)[i - 1].aux) return fail(GAC_ERR_INVALID_ARGUMENT, "convolver epochs must carry increasing quantum indices >= 1");
        } else {
          o.ftype = 0;
          o.aux = 0;
        }
        o.ir = ops[i].ir;
        if (o.ir && o.ir->ctx != ctx) return fail(GAC_ERR_INVALID_ARGUMENT, "impulse response belongs to another context");
        if (o.ir && !(o.ir->nch == 1 || o.ir->nch == 2 || (o.ir->nch == 4 && o.ir->true_stereo))) return fail(GAC_ERR_UNSUPPORTED, "discrete impulse responses with %d channels are outside the accelerated path", o.ir->nch);
        break;
      case GAC_OP_DELAY:
        o.aux = ops[i].aux;
        if (!(o.aux > 0.0 && o.aux <= 10.0)) return fail(GAC_ERR_OUT_OF_RANGE, "maxDelayTime must be in (0, 10] seconds");  // DelayNode.cs:25-26
        if ((rc = copy_param(ops[i].p0, &o.p0, "delay.delayTime"))) return rc;
        break;
      case GAC_OP_PANNER:
        o.aux = ops[i].aux;
        if (!(o.aux >= 0.0)) return fail(GAC_ERR_OUT_OF_RANGE, "panner: first processed quantum must be >= 0");
        if ((rc = copy_param(ops[i].p0, &o.p0, "panner.pan"))) return rc;
        break;
      case GAC_OP_CHANNEL:
        o.aux = ops[i].aux;
        if (!(o.aux >= 0.0 && o.aux < 32.0)) return fail(GAC_ERR_OUT_OF_RANGE, "channel index must be in 0 .. 31");  // ChannelSplitterNode.cs:17-18
        if (i != 0) return fail(GAC_ERR_INVALID_ARGUMENT, "GAC_OP_CHANNEL must be the first op of a chain fed by a bus");
        break;
      default:
        return fail(GAC_ERR_INVALID_ARGUMENT, "unknown op kind %d", o.kind);
    }
  }
  return GAC_OK;
}

extern "C" int gac_graph_create(gac_context* ctx, const gac_graph_desc* desc, gac_graph** out) {
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or destroyed");
  if (!desc || !out) return fail(GAC_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  if (desc->n_voices < 0 || desc->n_buses < 0) return fail(GAC_ERR_INVALID_ARGUMENT, "negative counts");
  if (desc->n_voices > 0 && !desc->voices) return fail(GAC_ERR_INVALID_ARGUMENT, "voices is null");
  if (desc->n_buses > 0 && !desc->buses) return fail(GAC_ERR_INVALID_ARGUMENT, "buses is null");
  auto g = std::make_unique<gac_graph>();
  g->ctx = ctx;
  g->voices.resize(desc->n_voices);
  for (int v = 0; v < desc->n_voices; v++) {
    const gac_voice_desc& d = desc->voices[v];
    VoiceH& h = g->voices[v];
    if (d.input < 0 || d.input > desc->n_buses) return fail(GAC_ERR_OUT_OF_RANGE, "voice %d: input bus %d out of range", v, d.input - 1);
    h.input_bus = d.input - 1;
    h.kind = h.input_bus < 0 ? d.source_kind : GAC_SOURCE_BUFFER;
    if (h.kind < GAC_SOURCE_BUFFER || h.kind > GAC_SOURCE_OSCILLATOR) return fail(GAC_ERR_INVALID_ARGUMENT, "voice %d: unknown source kind %d", v, h.kind);
    if (h.input_bus < 0 && h.kind != GAC_SOURCE_BUFFER) {
      int rc = copy_param(d.source_param, &h.src_param, h.kind == GAC_SOURCE_CONSTANT ? "constantSource.offset" : "oscillator.frequency");
      if (rc) return rc;
      h.osc_type = d.oscillator_type;
      if (h.kind == GAC_SOURCE_OSCILLATOR && (h.osc_type < 0 || h.osc_type > 3)) return fail(GAC_ERR_INVALID_ARGUMENT, "voice %d: unknown oscillator type %d", v, h.osc_type);
    } else if (h.input_bus < 0) {
      if (!d.source) return fail(GAC_ERR_INVALID_OPERATION, "voice %d: Cannot start without a buffer set", v);  // AudioBufferSourceNode.cs:86-87
      if (d.source->ctx != ctx) return fail(GAC_ERR_INVALID_ARGUMENT, "voice %d: buffer belongs to another context", v);
      if (d.source->nch > 2) return fail(GAC_ERR_UNSUPPORTED, "voice %d: sources with more than 2 channels are outside the accelerated path", v);
      if (!(d.playback_rate >= 0.001f && d.playback_rate <= 1000.f)) return fail(GAC_ERR_OUT_OF_RANGE, "voice %d: playbackRate outside [0.001, 1000]", v);
      if (d.source_param.n_events > 0 || d.source_param.mod_bus > 0) {
        // PlaybackRate with automation events / epochs (evaluated per quantum on the host) or with a modulation input (its k-rate
        // table is evaluated on the device behind the modulator's bus and read back before the positions are replayed)
        int rc = copy_param(d.source_param, &h.src_param, "bufferSource.playbackRate");
        if (rc) return rc;
        h.rate_events = true;
      }
    }
    if (d.bus < -1 || d.bus >= desc->n_buses) return fail(GAC_ERR_OUT_OF_RANGE, "voice %d: bus index %d out of range", v, d.bus);
    h.src = (h.input_bus < 0 && h.kind == GAC_SOURCE_BUFFER) ? d.source : nullptr;
    h.when = d.start_when;
    h.offset = d.start_offset;
    h.duration = d.start_duration;
    h.stop_when = d.stop_when;
    h.rate = d.playback_rate;
    h.loop = d.loop != 0;
    h.loop_start = std::max(0.0, d.loop_start);  // AudioBufferSourceNode.cs:52
    h.loop_end = std::max(0.0, d.loop_end);      // :61
    h.bus = d.bus;
    int rc = copy_ops(ctx, d.n_ops, d.ops, &h.ops);
    if (rc) return rc;
    if (!h.ops.empty() && h.ops[0].kind == GAC_OP_CHANNEL && h.input_bus < 0) return fail(GAC_ERR_INVALID_ARGUMENT, "voice %d: GAC_OP_CHANNEL needs a chain fed by a bus", v);
  }
  g->buses.resize(desc->n_buses);
  for (int b = 0; b < desc->n_buses; b++) {
    const gac_bus_desc& d = desc->buses[b];
    int rc = copy_ops(ctx, d.n_ops, d.ops, &g->buses[b].ops);
    if (rc) return rc;
    if (d.target < -1 || d.target > desc->n_buses || d.target == b + 1) return fail(GAC_ERR_OUT_OF_RANGE, "bus %d: target %d out of range", b, d.target);
    g->buses[b].target = d.target == 0 ? -1 : (d.target < 0 ? -2 : d.target - 1);
  }
  for (int b = 0; b < desc->n_buses; b++) {
    const gac_bus_desc& d = desc->buses[b];
    BusH& h = g->buses[b];
    h.mono = (d.flags & GAC_BUS_MONO_INPUT) != 0;
    if (d.input_slots && d.inputs && d.n_inputs > 0) {
      h.slots.assign(d.input_slots, d.input_slots + d.n_inputs);
      for (int sl : h.slots)
        if (sl < 0 || sl > 2) return fail(GAC_ERR_UNSUPPORTED, "bus %d: channel mergers with more than two inputs are outside the accelerated path", b);
      if (h.mono) return fail(GAC_ERR_INVALID_ARGUMENT, "bus %d: a merger bus cannot be a mono fan-in", b);
    }
    if (d.inputs && d.n_inputs > 0) {
      h.inputs.assign(d.inputs, d.inputs + d.n_inputs);
      for (int x : h.inputs) {
        if (x >= 0 && (x >= desc->n_buses || g->buses[x].target != b)) return fail(GAC_ERR_INVALID_ARGUMENT, "bus %d: input bus %d does not target it", b, x);
        if (x < 0 && (~x >= desc->n_voices || g->voices[~x].bus != b)) return fail(GAC_ERR_INVALID_ARGUMENT, "bus %d: input voice %d is not routed to it", b, ~x);
      }
    } else {
      for (int v = 0; v < desc->n_voices; v++)
        if (g->voices[v].bus == b) h.inputs.push_back(~v);
      for (int c = 0; c < desc->n_buses; c++)
        if (g->buses[c].target == b) h.inputs.push_back(c);
    }
  }
  if (desc->dest_inputs && desc->n_dest_inputs > 0) {
    g->dest_inputs.assign(desc->dest_inputs, desc->dest_inputs + desc->n_dest_inputs);
    for (int x : g->dest_inputs) {
      if (x >= 0 && (x >= desc->n_buses || g->buses[x].target != -1)) return fail(GAC_ERR_OUT_OF_RANGE, "dest_inputs: bus %d out of range or not connected to the destination", x);
      if (x < 0 && (~x >= desc->n_voices || g->voices[~x].bus != -1)) return fail(GAC_ERR_INVALID_ARGUMENT, "dest_inputs: voice %d is not a direct voice", ~x);
    }
  } else {
    for (int b = 0; b < desc->n_buses; b++)
      if (g->buses[b].target == -1) g->dest_inputs.push_back(b);
    for (int v = 0; v < desc->n_voices; v++)
      if (g->voices[v].bus == -1) g->dest_inputs.push_back(~v);
  }
  // modulation inputs name buses of this graph
  auto check_mod = [&](const ParamH& p) { return p.mod_bus < desc->n_buses && (p.mod_bus < 0 || g->buses[p.mod_bus].mono); };
  auto check_ops = [&](const std::vector<OpH>& ops) {
    for (const OpH& o : ops)
      if (!check_mod(o.p0) || !check_mod(o.p1) || !check_mod(o.p2)) return false;
    return true;
  };
  for (auto& v : g->voices)
    if (!check_ops(v.ops) || !check_mod(v.src_param)) return fail(GAC_ERR_OUT_OF_RANGE, "a parameter's modulation bus is out of range or not a GAC_BUS_MONO_INPUT bus");
  for (auto& b : g->buses)
    if (!check_ops(b.ops)) return fail(GAC_ERR_OUT_OF_RANGE, "a parameter's modulation bus is out of range or not a GAC_BUS_MONO_INPUT bus");
  *out = g.release();
  return GAC_OK;
}
extern "C" int gac_graph_destroy(gac_graph* g) {
  if (!g) return fail(GAC_ERR_INVALID_ARGUMENT, "graph is null");
  delete g;
  return GAC_OK;
}

// ------------------------------------------------------------------------------------------ render
struct Sig {
  float* p[2];
  // not yet materialised: frames [lo, hi) of the signal are read in place from the source buffer (lazy[c][n] for n in [lo, hi));
  // set for rate-1 sources that feed a convolver directly (or through a fused GainNode), so that no copy of the source is made
  const float* lazy[2] = {nullptr, nullptr};
  int64_t lo = 0, hi = 0;  // frames flagged non-silent (multiples of 128)
  int ch = 2;              // logical channel count of the block the reference would carry here (rows are always 2; mono = duplicated)
  bool from_source = false;  // the chain is fed by an AudioBufferSourceNode (whose idle blocks have ONE channel, AudioBufferSourceNode.cs:391-402)
  const std::vector<OpH>* ops = nullptr;
  size_t bus_base = 0;       // index of the graph's first bus in RenderEnv::buses (modulation inputs name graph-local buses)
  // >= 0: the signal's only consumer is a plain two-channel fan-in (this is its id): a ConvolverNode that ends the chain may be
  // summed with the others that carry the same id (run_chains; the sum lands in the first one's rows, the rest turn silent)
  int64_t sum_key = -1;
};

struct RenderEnv {
  gac_context* ctx;
  Scratch* scratch;
  HostKeep* keep;
  Timer* timer;
  int64_t Npad, NQ, QB;
  int64_t launches = 0;
  int64_t conv_units = 0;
  double alg_bytes = 0, macs = 0;
  double mac_h2_single = 0;             // bytes of ONE set of IR spectra per channel-convolver (the compulsory part of the H reads)
  double mac_flops = 0, mac_bytes = 0;  // flops issued / bytes the K6 variant in use has to move (X, H, Y once)
  int mac_big = 0;                      // double-length segments per channel-convolver of the last second-level-FFT batch
  int mac_used = 0;                     // K6 variant of the last convolver batch (1 stream, 2/4 tiled, 3 second-level FFT)
  int sum_groups = 0, sum_members = 0;  // fan-in groups summed as spectra / convolvers in them
  std::map<Sig*, std::pair<const float*, float>> fused;  // GainNode folded into the next convolver's forward FFT
  std::map<std::string, float*> param_tables;  // automation tables already evaluated in this render, by (rate, value, events)
  // automation tables are carved out of chunks (one stream-ordered allocation per 16 tables) and the events of a batch of
  // parameter jobs travel in one upload: a driver call per table and per event list costs more than evaluating them
  float* tab_chunk = nullptr;
  size_t tab_left = 0;
  std::vector<DevEvent>* ev_host = nullptr;
  std::vector<Sig>* buses = nullptr;     // every bus of the render (modulation inputs are read from here)
  std::vector<ModJob> mod_jobs;          // modulation sums queued behind the parameter jobs of the current batch
};

static int param_table(RenderEnv& env, const ParamH& p, bool a_rate, std::vector<ParamJob>& jobs, float** out_table, size_t bus_base = 0) {
  *out_table = nullptr;
  const Sig* mod = nullptr;
  if (p.mod_bus >= 0 && env.buses) {
    const Sig& m = (*env.buses)[bus_base + (size_t)p.mod_bus];
    if (m.hi > m.lo) mod = &m;  // a modulator that is silent throughout leaves the intrinsic value alone (AudioParam.cs:118-126)
  }
  if (p.ev.empty() && p.later.empty() && !mod) return GAC_OK;
  // voices that schedule the same automation (same value, same events) share one table: the curve depends on nothing else
  std::string key(1, a_rate ? 'a' : 'k');
  if (mod) {  // ... unless something is added to it: a modulated parameter owns its table
    key[0] = a_rate ? 'A' : 'K';
    const void* who = mod->p[0];
    key.append(reinterpret_cast<const char*>(&who), sizeof(who));
    key.append(reinterpret_cast<const char*>(&p.minv), sizeof(float));
    key.append(reinterpret_cast<const char*>(&p.maxv), sizeof(float));
  }
  key.append(reinterpret_cast<const char*>(&p.value), sizeof(float));
  auto key_events = [&](const std::vector<gac_event>& ev) {
    for (const gac_event& e : ev) {  // field by field: the struct has 4 bytes of padding
      key.append(reinterpret_cast<const char*>(&e.type), sizeof(e.type));
      key.append(reinterpret_cast<const char*>(&e.value), sizeof(e.value));
      key.append(reinterpret_cast<const char*>(&e.target), sizeof(e.target));
      key.append(reinterpret_cast<const char*>(&e.time), sizeof(e.time));
      key.append(reinterpret_cast<const char*>(&e.time_constant), sizeof(e.time_constant));
    }
  };
  key_events(p.ev);
  for (const ParamH::Epoch& ep : p.later) {
    key.append("|", 1);
    key.append(reinterpret_cast<const char*>(&ep.q0), sizeof(ep.q0));
    key.append(reinterpret_cast<const char*>(&ep.value), sizeof(ep.value));
    key_events(ep.ev);
  }
  auto hit = env.param_tables.find(key);
  if (hit != env.param_tables.end()) {
    *out_table = hit->second;
    return GAC_OK;
  }
  const size_t need = (((a_rate ? (size_t)env.Npad : (size_t)env.NQ) + 63) / 64) * 64;  // 256-byte aligned slices
  if (env.tab_left < need) {
    const size_t chunk = std::max(need, (size_t)16 * (size_t)env.Npad);
    int rc = env.scratch->alloc(&env.tab_chunk, chunk);
    if (rc) return rc;
    env.tab_left = chunk;
  }
  float* tab = env.tab_chunk;
  env.tab_chunk += need;
  env.tab_left -= need;
  env.param_tables[key] = tab;
  if (!env.ev_host) env.ev_host = &env.keep->make<DevEvent>();
  static_assert(sizeof(DevEvent) == sizeof(gac_event), "event layout");
  // one job per epoch (a parameter that was never edited between Render calls has one), each writing its own quanta
  auto add_job = [&](float value, const std::vector<gac_event>& ev, int64_t q_lo, int64_t q_hi) {
    const size_t off = env.ev_host->size();
    env.ev_host->resize(off + ev.size());
    if (!ev.empty()) memcpy(env.ev_host->data() + off, ev.data(), sizeof(gac_event) * ev.size());
    ParamJob j;
    j.value = value;
    j.n_events = (int)ev.size();
    j.events = reinterpret_cast<const DevEvent*>(off);  // offset into the batch's event block; made a pointer by run_param_jobs
    j.out = tab;
    j.a_rate = a_rate ? 1 : 0;
    j.q_lo = q_lo;
    j.q_hi = q_hi;
    jobs.push_back(j);
  };
  add_job(p.value, p.ev, 0, p.later.empty() ? std::numeric_limits<int64_t>::max() : p.later[0].q0);
  for (size_t e = 0; e < p.later.size(); e++)
    add_job(p.later[e].value, p.later[e].ev, p.later[e].q0, e + 1 < p.later.size() ? p.later[e + 1].q0 : std::numeric_limits<int64_t>::max());
  if (mod) {
    ModJob m;
    m.table = tab;
    m.mod = mod->p[0];  // the parameter's input has one channel (a GAC_BUS_MONO_INPUT bus keeps it in both rows)
    m.lo = mod->lo;
    m.hi = mod->hi;
    m.minv = p.minv;
    m.maxv = p.maxv;
    m.a_rate = a_rate ? 1 : 0;
    env.mod_jobs.push_back(m);
  }
  *out_table = tab;
  return GAC_OK;
}

static int run_param_jobs(RenderEnv& env, std::vector<ParamJob>& jobs) {
  if (jobs.empty()) return GAC_OK;
  DevEvent* dev = nullptr;
  int rc = env.scratch->upload(&dev, *env.ev_host);
  if (rc) return rc;
  env.ev_host = nullptr;  // the next batch starts its own block (this one stays alive in `keep`)
  auto& hj = env.keep->make<ParamJob>();
  hj = jobs;
  for (ParamJob& j : hj) j.events = dev + reinterpret_cast<size_t>(j.events);
  ParamJob* dj = nullptr;
  rc = env.scratch->upload(&dj, hj);
  if (rc) return rc;
  int t = env.timer->begin(C_AUTO);
  // a-rate and k-rate jobs share a launch; the kernel branches per job
  launch_param_eval(dj, (int)hj.size(), env.ctx->d_bt, env.NQ, env.ctx->fs, kstream(env.ctx));
  if (!env.mod_jobs.empty()) {  // intrinsic + modulation, clamped (AudioParam.cs:125-131, :150-155)
    auto& hm = env.keep->make<ModJob>();
    hm = env.mod_jobs;
    env.mod_jobs.clear();
    ModJob* dm = nullptr;
    if ((rc = env.scratch->upload(&dm, hm))) return rc;
    launch_param_modulate(dm, (int)hm.size(), env.Npad, kstream(env.ctx));
    env.launches++;
  }
  env.timer->end(t);
  env.launches += ((int64_t)hj.size() + 65534) / 65535;
  CU(cudaGetLastError());
  return GAC_OK;
}

// ---- the convolver: K5 -> K6 -> K7.  One ConvItem = one ConvolverNode instance of one signal
// (Nodes/ConvolverNode.cs:102-155): 1-2 forward transforms, 1-4 channel-convolvers (PartitionedConvolver each),
// 1-2 inverse transforms.
//   stereo IR     : fwd L, R        mac (L,H0) (R,H1)                    inv Y0 -> L ; Y1 -> R              (:145-151)
//   mono IR       : fwd (L+R)/sqrt2 mac (M,H0)                           inv Y0 -> L and R (1 -> 2 up-mix)   (AudioNodeInput.cs:214-228, 201-213)
//   true stereo   : fwd L, R        mac (L,H0) (R,H2) (L,H1) (R,H3)      inv Y0+Y1 -> L ; Y2+Y3 -> R         (:127-144)
struct ConvItem {
  int n_fwd = 0;
  struct {
    const float* in;
    const float* in2;  // down-mix partner or nullptr
    float mix_scale;
  } fwd[2];
  int64_t lo = 0, hi = 0;            // non-silent input frames
  const float* gain_tab = nullptr;   // fused preceding GainNode (may be null)
  float gain_const = 1.0f;
  int n_mac = 0;
  struct {
    int x;            // which forward spectrum feeds this channel-convolver
    const float2* H;  // packed IR spectra [P16][B]
    const float2* H2; // second-level IR spectra [B+1][M2] (null: direct MAC)
    const float2* H2b = nullptr;  // the same for transforms of length 2 * M2 (n_big > 0)
  } mac[4];
  int P = 0;
  int M2 = 0, Lh = 0;  // second-level transform length / history blocks (0: direct MAC)
  int n_big = 0, n_small = 0;  // segments of length 2 * M2 in front, of length M2 behind them (plan_segments)
  int n_inv = 0;
  struct {
    int y, y2;        // spectrogram(s) of which channel-convolver(s); y2 = -1: none
    float* out;
    float* out2;      // duplicate destination or nullptr
  } inv[2];
  int64_t sum_key = -1;  // fan-in fusion: items with the same key (and segment plan) are summed as spectra (conv_batch_fft2_sum)
};

static int conv_batch_direct(RenderEnv& env, std::vector<ConvItem>& items) {
  gac_context* ctx = env.ctx;
  const int B = ctx->B;
  const int TB = ctx->tile_blocks;
  const int64_t QB = env.QB;
  const int64_t QBpad = ((QB + TB - 1) / TB) * TB;
  const int64_t rowsX = TB + QBpad;  // TB zero rows in front (blocks -TB..-1)
  const int groups = B / 128;
  // sub-batches bounded by the scratch budget (an item needs up to 2 X and 4 Y spectrograms)
  const size_t xbytes = (size_t)rowsX * B * sizeof(float2), ybytes = (size_t)QBpad * B * sizeof(float2);
  for (size_t i0 = 0; i0 < items.size();) {
    size_t ni = 0, nx = 0, ny = 0;
    while (i0 + ni < items.size()) {
      const ConvItem& it = items[i0 + ni];
      if (ni > 0 && (nx + it.n_fwd) * xbytes + (ny + it.n_mac) * ybytes > ctx->scratch_budget) break;
      nx += it.n_fwd;
      ny += it.n_mac;
      ni++;
    }
    float2 *dX = nullptr, *dY = nullptr;
    int rc = env.scratch->alloc(&dX, nx * rowsX * B);
    if (rc) return rc;
    rc = env.scratch->alloc(&dY, ny * QBpad * B);
    if (rc) return rc;
    // zero the front pad and the tail rows of X (rows the forward FFT does not write)
    CU(cudaMemset2DAsync(dX, xbytes, 0, (size_t)TB * B * sizeof(float2), nx, ctx->stream));
    if (QBpad > QB)
      CU(cudaMemset2DAsync(dX + (size_t)(TB + QB) * B, xbytes, 0, (size_t)(QBpad - QB) * B * sizeof(float2), nx, ctx->stream));
    auto& fj = env.keep->make<FftFwdJob>();
    auto& mj = env.keep->make<MacJob>();
    auto& ij = env.keep->make<FftInvJob>();
    auto& tiles = env.keep->make<MacTile>();
    size_t xi = 0, yi = 0;
    for (size_t i = 0; i < ni; i++) {
      ConvItem& it = items[i0 + i];
      float2* Xc[2] = {nullptr, nullptr};
      float2* Yc[4] = {nullptr, nullptr, nullptr, nullptr};
      for (int k = 0; k < it.n_fwd; k++) {
        Xc[k] = dX + (xi++) * rowsX * B + (size_t)TB * B;  // row 0
        FftFwdJob f;
        f.in = it.fwd[k].in;
        f.out = Xc[k];
        f.scale = nullptr;
        f.gain = it.gain_tab;
        f.gain_const = it.gain_const;
        f.n_valid = env.Npad;
        f.n_blocks = QB;
        f.gate_lo = it.lo;
        f.gate_hi = it.hi;
        f.in2 = it.fwd[k].in2;
        f.mix_scale = it.fwd[k].mix_scale;
        fj.push_back(f);
      }
      for (int k = 0; k < it.n_mac; k++) {
        Yc[k] = dY + (yi++) * QBpad * B;
        for (int g = 0; g < groups; g++) {
          MacJob m;
          m.X = Xc[it.mac[k].x] + g * 128;
          m.H = it.mac[k].H + g * 128;
          m.Y = Yc[k] + g * 128;
          m.P = it.P;
          m.has_dc = (g == 0);
          mj.push_back(m);
        }
        // accounting: one unit = one channel-convolver block of B frames through P partitions
        const double P = it.P, C = B + 1;
        env.conv_units += QB;
        env.alg_bytes += (double)QB * (16.0 * P * C + 8.0 * C + 8.0 * B);
        env.mac_bytes += 8.0 * B * ((double)QB * 2 + P);  // X and Y once, H once
        env.mac_h2_single += 8.0 * B * P;
      }
      for (int k = 0; k < it.n_inv; k++) {
        FftInvJob v;
        v.in = Yc[it.inv[k].y];
        v.out = it.inv[k].out;
        v.n_blocks = QB;
        v.in2 = it.inv[k].y2 >= 0 ? Yc[it.inv[k].y2] : nullptr;
        v.out2 = it.inv[k].out2;
        ij.push_back(v);
      }
    }
    // tiles, heaviest first (stable: tiles of one job stay adjacent for L2 reuse of its H and X rows)
    for (int j = 0; j < (int)mj.size(); j++)
      for (int64_t b0 = 0; b0 < QB; b0 += TB) tiles.push_back(MacTile{j, (int)b0});
    auto stages = [&](const MacTile& t) {
      int p16 = (mj[t.job].P + 15) / 16;
      int causal = (t.b0 + TB) / 16;
      return std::min(p16, causal);
    };
    std::stable_sort(tiles.begin(), tiles.end(), [&](const MacTile& a, const MacTile& b) { return stages(a) > stages(b); });
    for (auto& t : tiles) {
      env.macs += (double)stages(t) * 16.0 * TB * 128.0;
      env.mac_flops += 8.0 * (double)stages(t) * 16.0 * TB * 128.0;
    }
    env.mac_used = ctx->mac_variant == 1 ? 1 : (ctx->mac_variant == 2 ? 2 : 4);

    FftFwdJob* dfj = nullptr;
    MacJob* dmj = nullptr;
    FftInvJob* dij = nullptr;
    MacTile* dt = nullptr;
    if ((rc = env.scratch->upload(&dfj, fj))) return rc;
    if ((rc = env.scratch->upload(&dmj, mj))) return rc;
    if ((rc = env.scratch->upload(&dij, ij))) return rc;
    if ((rc = env.scratch->upload(&dt, tiles))) return rc;

    int t = env.timer->begin(C_FFT_FWD);
    launch_rfft_fwd(dfj, (int)fj.size(), QB, B, ctx->d_tw, kstream(ctx));
    env.timer->end(t);
    CU(cudaGetLastError());
    t = env.timer->begin(C_MAC);
    if (ctx->mac_variant == 1)
      launch_mac_stream(dmj, (int)mj.size(), QB, B, kstream(ctx));
    else
    {
      int pmax = 1;
      for (auto& m : mj) pmax = std::max(pmax, m.P);
      launch_mac_tiled(dmj, (int)mj.size(), dt, (int)tiles.size(), QB, pmax, B, TB, ctx->mac_variant == 2 ? 2 : 0, kstream(ctx));
    }
    env.timer->end(t);
    CU(cudaGetLastError());
    t = env.timer->begin(C_FFT_INV);
    launch_irfft_ola(dij, (int)ij.size(), QB, B, ctx->d_tw, kstream(ctx));
    env.timer->end(t);
    CU(cudaGetLastError());
    env.launches += (ctx->mac_variant == 1) ? 3 : 4;  // K5, K6 (+ k_mac_dc), K7
    i0 += ni;
  }
  return GAC_OK;
}

// Overlap-save segments for QB output blocks: every segment spends Lh of its M points on history, so a render of QB = 4500 blocks with
// Lh = 752 needs four 2048-point segments (V = 1296 valid outputs each) — or one 4096-point segment (V = 3344) followed by one
// 2048-point segment, 25 % fewer transform points.  Picks the number of double-length segments in front (radix-16 plan: 2 M <= 4096).
static void plan_segments(int64_t QB, int M, int Lh, int* n_big, int* n_small) {
  int log2m = 0;
  while ((1 << log2m) < M) log2m++;
  const int64_t V = M - Lh;
  *n_big = 0;
  *n_small = (int)((QB + V - 1) / V);
  if (2 * M > fft2_r16_max()) return;
  const int64_t Vb = 2 * (int64_t)M - Lh;
  double best = (double)*n_small * M * log2m;
  for (int nb = 1; (int64_t)(nb - 1) * Vb < QB; nb++) {
    const int64_t rem = std::max<int64_t>(0, QB - (int64_t)nb * Vb);
    const int ns = (int)((rem + V - 1) / V);
    // (the double-length kernel keeps half as many CTAs resident: measured at ~0.8 of the per-point rate, 4096 against 2048 points)
    const double cost = (double)nb * 2 * M * (log2m + 1) / 0.8 + (double)ns * M * log2m;
    if (cost < best * 0.97) {  // (a tie keeps the single-length plan: one launch less)
      best = cost;
      *n_big = nb;
      *n_small = ns;
    }
  }
}

extern "C" int gac_plan_segments(int64_t n_blocks, int n_partitions, int uniform, int* m, int* n_big, int* n_small) {
  if (!m || !n_big || !n_small) return fail(GAC_ERR_INVALID_ARGUMENT, "null argument");
  if (n_blocks <= 0 || n_partitions <= 0) return fail(GAC_ERR_OUT_OF_RANGE, "block and partition counts must be positive");
  int Lh = 0;
  const int M = n_partitions >= kFft2MinP ? fft2_pick_m(n_partitions, &Lh) : 0;
  *m = M;
  *n_big = *n_small = 0;
  if (M == 0) return GAC_OK;
  if (uniform) *n_small = (int)((n_blocks + (M - Lh) - 1) / (M - Lh));
  else plan_segments(n_blocks, M, Lh, n_big, n_small);
  return GAC_OK;
}

// Second-level spectra of length 2 * M2 for the impulse responses a voice batch wants to run with mixed segment lengths: allocated
// on first use, prepared by ONE launch per transform length for the whole batch (the first-level spectra exist: prepare_irs ran).
static int ensure_h2b_batch(RenderEnv& env, const std::vector<const gac_ir*>& irs) {
  gac_context* ctx = env.ctx;
  const int B = ctx->B;
  std::map<int, std::vector<IrChanJob>> by_m;
  for (const gac_ir* cir : irs) {
    gac_ir* ir = const_cast<gac_ir*>(cir);
    if (ir->d_H2b || ir->M2 <= 0) continue;
    int n_big = 0, n_small = 0;
    plan_segments(env.QB, ir->M2, ir->Lh, &n_big, &n_small);
    if (n_big == 0) continue;
    const int Mb = 2 * ir->M2;
    const size_t row = (size_t)fft2_h2_row_elems(Mb);
    CU(cudaMallocAsync(&ir->d_H2b, sizeof(float2) * (size_t)ir->nch * (B + 1) * row, ctx->stream));
    for (int c = 0; c < ir->nch; c++) {
      IrChanJob j{};
      j.H = ir->d_H + (size_t)c * ir->P16 * B;
      j.P = ir->P;
      j.P16 = ir->P16;
      j.H2 = ir->d_H2b + (size_t)c * (B + 1) * row;
      by_m[Mb].push_back(j);
    }
  }
  for (auto& kv : by_m) {
    auto& hj = env.keep->make<IrChanJob>();
    hj = kv.second;
    IrChanJob* dj = nullptr;
    int rc = env.scratch->upload(&dj, hj);
    if (rc) return rc;
    launch_fft2_prep_batch(dj, (int)hj.size(), B, kv.first, ctx->d_tab16, kstream(ctx));
    env.launches++;
    CU(cudaGetLastError());
  }
  return GAC_OK;
}

// The convolver with the spectral MAC done as a fast convolution along block time (fft2.cu):
// K5 (transposing) -> k_fft2_conv -> K7 (transposing), over transposed spectrograms XT/YT[chan][B+1][Qs].
static int conv_batch_fft2(RenderEnv& env, std::vector<ConvItem>& items, int M) {
  gac_context* ctx = env.ctx;
  const int B = ctx->B, C = B + 1;
  const int64_t QB = env.QB;
  const int64_t Qs = ((QB + 15) / 16) * 16;  // row stride: rows start 128-byte aligned
  const size_t sbytes = (size_t)C * Qs * sizeof(float2);
  int log2m = 0;
  while ((1 << log2m) < M) log2m++;
  for (size_t i0 = 0; i0 < items.size();) {
    size_t ni = 0, nx = 0, ny = 0;
    while (i0 + ni < items.size()) {
      const ConvItem& it = items[i0 + ni];
      if (ni > 0 && (nx + it.n_fwd + ny + it.n_mac) * sbytes > ctx->scratch_budget) break;
      nx += it.n_fwd;
      ny += it.n_mac;
      ni++;
    }
    float2 *dX = nullptr, *dY = nullptr;
    int rc = env.scratch->alloc(&dX, nx * (size_t)C * Qs);
    if (rc) return rc;
    rc = env.scratch->alloc(&dY, ny * (size_t)C * Qs);
    if (rc) return rc;
    auto& fj = env.keep->make<FftFwdJob>();
    auto& cj = env.keep->make<Fft2Job>();
    auto& cjb = env.keep->make<Fft2Job>();  // double-length segments in front (plan_segments)
    auto& ij = env.keep->make<FftInvJob>();
    size_t xi = 0, yi = 0;
    int max_seg = 0, max_seg_big = 0;
    for (size_t i = 0; i < ni; i++) {
      ConvItem& it = items[i0 + i];
      float2* Xc[2] = {nullptr, nullptr};
      float2* Yc[4] = {nullptr, nullptr, nullptr, nullptr};
      for (int k = 0; k < it.n_fwd; k++) {
        Xc[k] = dX + (xi++) * (size_t)C * Qs;
        FftFwdJob f;
        f.in = it.fwd[k].in;
        f.out = Xc[k];
        f.scale = nullptr;
        f.gain = it.gain_tab;
        f.gain_const = it.gain_const;
        f.n_valid = env.Npad;
        f.n_blocks = QB;
        f.gate_lo = it.lo;
        f.gate_hi = it.hi;
        f.in2 = it.fwd[k].in2;
        f.mix_scale = it.fwd[k].mix_scale;
        fj.push_back(f);
      }
      const int V = M - it.Lh;
      const int n_big = it.mac[0].H2b ? it.n_big : 0;
      const int nseg = n_big > 0 ? it.n_small : (int)((QB + V - 1) / V);
      const int64_t b0 = (int64_t)n_big * (2 * (int64_t)M - it.Lh);  // (multiple of 16: M and Lh are)
      max_seg = std::max(max_seg, nseg);
      max_seg_big = std::max(max_seg_big, n_big);
      env.mac_big = std::max(env.mac_big, n_big);
      for (int k = 0; k < it.n_mac; k++) {
        Yc[k] = dY + (yi++) * (size_t)C * Qs;
        Fft2Job j;
        j.X = Xc[it.mac[k].x];
        j.H2 = it.mac[k].H2;
        j.Y = Yc[k];
        j.Lh = it.Lh;
        j.nseg = nseg;
        j.b0 = b0;
        if (nseg > 0) cj.push_back(j);
        if (n_big > 0) {
          j.H2 = it.mac[k].H2b;
          j.nseg = n_big;
          j.b0 = 0;
          cjb.push_back(j);
        }
        const double P = it.P;
        env.conv_units += QB;
        env.alg_bytes += (double)QB * (16.0 * P * C + 8.0 * C + 8.0 * B);
        env.macs += (double)QB * P * C;  // complex MACs the direct sum would need (not issued: see mac_flops)
        env.mac_flops += (double)nseg * C * (2.0 * 5.0 * M * log2m + 6.0 * M) + (double)n_big * C * (2.0 * 5.0 * 2 * M * (log2m + 1) + 6.0 * 2 * M);
        // XT once (window overlaps hit L2), the H2 row(s), YT
        env.mac_h2_single += 8.0 * C * (double)fft2_h2_row_elems(M);
        env.mac_bytes += 8.0 * C * ((double)QB + (nseg > 0 ? (double)fft2_h2_row_elems(M) : 0.0) + (n_big > 0 ? (double)fft2_h2_row_elems(2 * M) : 0.0) + (double)QB);
      }
      for (int k = 0; k < it.n_inv; k++) {
        FftInvJob v;
        v.in = Yc[it.inv[k].y];
        v.out = it.inv[k].out;
        v.n_blocks = QB;
        v.in2 = it.inv[k].y2 >= 0 ? Yc[it.inv[k].y2] : nullptr;
        v.out2 = it.inv[k].out2;
        ij.push_back(v);
      }
    }
    env.mac_used = 3;
    FftFwdJob* dfj = nullptr;
    Fft2Job* dcj = nullptr;
    Fft2Job* dcjb = nullptr;
    FftInvJob* dij = nullptr;
    if ((rc = env.scratch->upload(&dfj, fj))) return rc;
    if ((rc = env.scratch->upload(&dcj, cj))) return rc;
    if ((rc = env.scratch->upload(&dcjb, cjb))) return rc;
    if ((rc = env.scratch->upload(&dij, ij))) return rc;
    TRACE_MARK("conv fft2: tables uploaded");
    int t = env.timer->begin(C_FFT_FWD);
    launch_rfft_fwd_t8(dfj, (int)fj.size(), QB, B, Qs, ctx->d_tab16, ctx->d_tw, kstream(ctx));
    env.timer->end(t);
    CU(cudaGetLastError());
    t = env.timer->begin(C_MAC);
    launch_fft2_conv(dcjb, (int)cjb.size(), max_seg_big, C, 2 * M, ctx->d_tw2, ctx->d_tab16, QB, Qs, Qs, kstream(ctx));
    launch_fft2_conv(dcj, (int)cj.size(), max_seg, C, M, ctx->d_tw2, ctx->d_tab16, QB, Qs, Qs, kstream(ctx));
    env.timer->end(t);
    if (!cjb.empty()) env.launches += 1;
    CU(cudaGetLastError());
    t = env.timer->begin(C_FFT_INV);
    launch_irfft_ola_t8(dij, (int)ij.size(), QB, B, Qs, ctx->d_tab16, ctx->d_tw, kstream(ctx));
    env.timer->end(t);
    CU(cudaGetLastError());
    env.launches += 3;
    i0 += ni;
  }
  return GAC_OK;
}

// Fan-in fusion: how many voices one CTA of k_fft2_sum16 walks.  A chunk costs (members + ~0.8) transforms (one forward transform
// and a multiply-accumulate per member, one inverse per chunk); the launches (double-length segments, then single-length ones) run
// in whole waves of resident CTAs, so the chunk count is picked to fill the last wave.
static int pick_sum_chunks(int n, int n_big, int n_small, int M, int C, int sms) {
  int log2m = 0;
  while ((1 << log2m) < M) log2m++;
  const char* force = getenv("GAC_SUM_CHUNK");
  if (force && atoi(force) > 0) return (n + atoi(force) - 1) / atoi(force);
  int best_chunks = 1;
  double best = 1e300;
  for (int nch = 1; nch <= n; nch++) {
    const int chunk = (n + nch - 1) / nch;
    if (chunk < 4 && nch > 1) break;
    double cost = 0.0;
    if (n_big > 0) {
      const double ctas = (double)n_big * C * 2 * nch, slots = (double)sms * fft2_sum_ctas_per_sm(2 * M);
      cost += std::ceil(ctas / slots) * (chunk + 0.8) * 2.0 * M * (log2m + 1) / 0.9;
    }
    if (n_small > 0) {
      const double ctas = (double)n_small * C * 2 * nch, slots = (double)sms * fft2_sum_ctas_per_sm(M);
      cost += std::ceil(ctas / slots) * (chunk + 0.8) * (double)M * log2m;
    }
    if (cost < best) {
      best = cost;
      best_chunks = nch;
    }
  }
  return best_chunks;
}

// One fan-in group (ConvItem::sum_key; stereo impulse responses, one segment plan): K5 per voice, k_fft2_sum16 per chunk of voices
// and output channel, ONE K7 pair for the group.  The sum lands in the rows of the first item.
static int conv_batch_fft2_sum(RenderEnv& env, std::vector<ConvItem>& items, int M) {
  gac_context* ctx = env.ctx;
  const int B = ctx->B, C = B + 1;
  const int64_t QB = env.QB;
  const int64_t Qs = ((QB + 15) / 16) * 16;
  const size_t plane = (size_t)C * Qs;
  int log2m = 0;
  while ((1 << log2m) < M) log2m++;
  const int n = (int)items.size();
  const ConvItem& first = items[0];
  const int V = M - first.Lh;
  const int n_big = first.n_big;
  const int nseg = n_big > 0 ? first.n_small : (int)((QB + V - 1) / V);
  const int64_t b0 = (int64_t)n_big * (2 * (int64_t)M - first.Lh);
  const int n_chunks = pick_sum_chunks(n, n_big, nseg, M, C, ctx->sm_count);
  float2 *dX = nullptr, *dYp = nullptr;
  int rc;
  if ((rc = env.scratch->alloc(&dX, (size_t)n * 2 * plane))) return rc;
  if ((rc = env.scratch->alloc(&dYp, (size_t)n_chunks * 2 * plane))) return rc;  // [channel][chunk][C][Qs]
  auto& fj = env.keep->make<FftFwdJob>();
  auto& mem = env.keep->make<Fft2SumMember>();    // [channel][voice], single-length spectra
  auto& memb = env.keep->make<Fft2SumMember>();   // the same with the double-length spectra
  for (int c = 0; c < 2; c++)
    for (int i = 0; i < n; i++) {
      const ConvItem& it = items[i];
      float2* X = dX + ((size_t)i * 2 + c) * plane;
      {  // (stereo layout: channel-convolver c reads forward spectrum c)
        FftFwdJob f;
        f.in = it.fwd[c].in;
        f.out = X;
        f.scale = nullptr;
        f.gain = it.gain_tab;
        f.gain_const = it.gain_const;
        f.n_valid = env.Npad;
        f.n_blocks = QB;
        f.gate_lo = it.lo;
        f.gate_hi = it.hi;
        f.in2 = it.fwd[c].in2;
        f.mix_scale = it.fwd[c].mix_scale;
        fj.push_back(f);
      }
      mem.push_back(Fft2SumMember{X, it.mac[c].H2});
      if (n_big > 0) memb.push_back(Fft2SumMember{X, it.mac[c].H2b});
      const double P = it.P;
      env.conv_units += QB;
      env.alg_bytes += (double)QB * (16.0 * P * C + 8.0 * C + 8.0 * B);
      env.macs += (double)QB * P * C;
      // one forward transform and a multiply-accumulate per member and segment (the inverse is counted per chunk below)
      env.mac_flops += (double)nseg * C * (5.0 * M * log2m + 8.0 * M) + (double)n_big * C * (5.0 * 2 * M * (log2m + 1) + 8.0 * 2 * M);
      env.mac_h2_single += 8.0 * C * (double)fft2_h2_row_elems(M);
      env.mac_bytes += 8.0 * C * ((double)QB + (nseg > 0 ? (double)fft2_h2_row_elems(M) : 0.0) + (n_big > 0 ? (double)fft2_h2_row_elems(2 * M) : 0.0));
    }
  Fft2SumMember *dmem = nullptr, *dmemb = nullptr;
  if ((rc = env.scratch->upload(&dmem, mem))) return rc;
  if ((rc = env.scratch->upload(&dmemb, memb))) return rc;
  auto& sj = env.keep->make<Fft2SumJob>();
  auto& sjb = env.keep->make<Fft2SumJob>();
  auto& ij = env.keep->make<FftInvJob>();
  for (int c = 0; c < 2; c++) {
    int at = 0;
    for (int k = 0; k < n_chunks; k++) {
      const int cnt = n / n_chunks + (k < n % n_chunks ? 1 : 0);
      Fft2SumJob j;
      j.members = dmem + (size_t)c * n + at;
      j.n_members = cnt;
      j.Y = dYp + ((size_t)c * n_chunks + k) * plane;
      j.Lh = first.Lh;
      j.nseg = nseg;
      j.b0 = b0;
      if (nseg > 0) sj.push_back(j);
      if (n_big > 0) {
        j.members = dmemb + (size_t)c * n + at;
        j.nseg = n_big;
        j.b0 = 0;
        sjb.push_back(j);
      }
      at += cnt;
      env.mac_flops += (double)nseg * C * (5.0 * M * log2m) + (double)n_big * C * (5.0 * 2 * M * (log2m + 1));
      env.mac_bytes += 8.0 * C * (double)QB;  // the chunk's partial spectrogram
    }
    FftInvJob v;
    v.in = dYp + (size_t)c * n_chunks * plane;
    v.out = first.inv[c].out;
    v.n_blocks = QB;
    v.n_parts = n_chunks;
    v.part_stride = (int64_t)plane;
    ij.push_back(v);
  }
  env.mac_big = std::max(env.mac_big, n_big);
  env.mac_used = 3;
  env.sum_groups++;
  env.sum_members += n;
  FftFwdJob* dfj = nullptr;
  Fft2SumJob *dsj = nullptr, *dsjb = nullptr;
  FftInvJob* dij = nullptr;
  if ((rc = env.scratch->upload(&dfj, fj))) return rc;
  if ((rc = env.scratch->upload(&dsj, sj))) return rc;
  if ((rc = env.scratch->upload(&dsjb, sjb))) return rc;
  if ((rc = env.scratch->upload(&dij, ij))) return rc;
  TRACE_MARK("conv fft2 (fan-in group): tables uploaded");
  int t = env.timer->begin(C_FFT_FWD);
  launch_rfft_fwd_t8(dfj, (int)fj.size(), QB, B, Qs, ctx->d_tab16, ctx->d_tw, kstream(ctx));
  env.timer->end(t);
  CU(cudaGetLastError());
  t = env.timer->begin(C_MAC);
  launch_fft2_sum(dsjb, (int)sjb.size(), n_big, C, 2 * M, ctx->d_tab16, QB, Qs, Qs, kstream(ctx));
  launch_fft2_sum(dsj, (int)sj.size(), nseg, C, M, ctx->d_tab16, QB, Qs, Qs, kstream(ctx));
  env.timer->end(t);
  CU(cudaGetLastError());
  t = env.timer->begin(C_FFT_INV);
  launch_irfft_ola_t8(dij, (int)ij.size(), QB, B, Qs, ctx->d_tab16, ctx->d_tw, kstream(ctx));
  env.timer->end(t);
  CU(cudaGetLastError());
  env.launches += 3 + (n_big > 0 && nseg > 0 ? 1 : 0);
  return GAC_OK;
}

// splits the items by K6 variant (direct MAC / second-level FFT of each length) and runs each class batched
static int conv_batch(RenderEnv& env, std::vector<ConvItem>& items) {
  std::map<int, std::vector<ConvItem>> classes;
  std::map<std::tuple<int64_t, int, int, int>, std::vector<ConvItem>> groups;  // fan-in groups: (key, M2, Lh, double-length segments)
  for (auto& it : items) {
    bool f2 = it.M2 > 0;
    for (int k = 0; k < it.n_mac; k++) f2 = f2 && it.mac[k].H2 != nullptr;
    if (f2 && it.sum_key >= 0) groups[std::make_tuple(it.sum_key, it.M2, it.Lh, it.n_big)].push_back(it);
    else classes[f2 ? it.M2 : 0].push_back(it);
  }
  for (auto& kv : classes) {
    int rc = kv.first == 0 ? conv_batch_direct(env, kv.second) : conv_batch_fft2(env, kv.second, kv.first);
    if (rc) return rc;
  }
  for (auto& kv : groups) {
    int rc = conv_batch_fft2_sum(env, kv.second, std::get<1>(kv.first));
    if (rc) return rc;
  }
  return GAC_OK;
}

// Deferred preparation of the impulse responses a batch of convolvers is about to use (async mode): one scale launch, one
// first-level transform launch and one second-level launch per transform length for ALL of them
// (PartitionedConvolver.cs:65-102 per channel; three launches per impulse response when done eagerly).
static int prepare_irs(RenderEnv& env, const std::vector<const gac_ir*>& irs) {
  gac_context* ctx = env.ctx;
  const int B = ctx->B;
  auto& cj = env.keep->make<IrChanJob>();
  auto& fj = env.keep->make<FftFwdJob>();
  std::vector<gac_ir*> todo;
  int pmax = 0;
  for (const gac_ir* cir : irs) {
    gac_ir* ir = const_cast<gac_ir*>(cir);
    if (ir->prepared) continue;
    ir->prepared = true;  // (also de-duplicates the list)
    todo.push_back(ir);
    wait_ready(ctx, ir->src);
    const int64_t h2row = ir->M2 > 0 ? fft2_h2_row_elems(ir->M2) : 0;
    for (int c = 0; c < ir->nch; c++) {
      IrChanJob j;
      j.ir = ir->src->d + (size_t)c * ir->src->stride;
      j.n_frames = ir->frames;
      j.normalize = ir->normalize && ir->frames > 0 ? 1 : 0;
      j.scale = ir->d_scale + c;
      j.H = ir->d_H + (size_t)c * ir->P16 * B;
      j.P = ir->P;
      j.P16 = ir->P16;
      j.H2 = ir->d_H2 ? ir->d_H2 + (size_t)c * (B + 1) * h2row : nullptr;
      cj.push_back(j);
      FftFwdJob f;
      f.in = j.ir;
      f.out = j.H;
      f.scale = j.scale;
      f.gain = nullptr;
      f.gain_const = 1.0f;
      f.n_valid = ir->frames;
      f.n_blocks = ir->P;
      f.gate_lo = 0;
      f.gate_hi = std::numeric_limits<int64_t>::max();
      fj.push_back(f);
      pmax = std::max(pmax, ir->P);
    }
  }
  if (todo.empty()) return GAC_OK;
  // channels that share a second-level transform length are contiguous in the job array
  std::vector<size_t> order(cj.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = i;
  std::vector<int> m_of(cj.size());
  {
    size_t k = 0;
    for (gac_ir* ir : todo)
      for (int c = 0; c < ir->nch; c++) m_of[k++] = ir->d_H2 ? ir->M2 : 0;
  }
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return m_of[a] < m_of[b]; });
  auto& cjs = env.keep->make<IrChanJob>();
  std::vector<int> ms;
  for (size_t i : order) {
    cjs.push_back(cj[i]);
    ms.push_back(m_of[i]);
  }
  IrChanJob* dcj = nullptr;
  FftFwdJob* dfj = nullptr;
  int rc;
  if ((rc = env.scratch->upload(&dcj, cjs))) return rc;
  if ((rc = env.scratch->upload(&dfj, fj))) return rc;
  const float cal = (float)std::pow(10.0, (double)(-58.f * 0.05f));  // PartitionedConvolver.cs:95,101
  launch_ir_scale_batch(dcj, (int)cjs.size(), cal, B, kstream(ctx));
  launch_rfft_fwd(dfj, (int)fj.size(), pmax, B, ctx->d_tw, kstream(ctx));
  env.launches += 2;
  for (size_t i0 = 0; i0 < ms.size();) {
    size_t i1 = i0;
    while (i1 < ms.size() && ms[i1] == ms[i0]) i1++;
    const int M = ms[i0];
    if (M > 0 && M <= fft2_r16_max()) {
      launch_fft2_prep_batch(dcj + i0, (int)(i1 - i0), B, M, ctx->d_tab16, kstream(ctx));
      env.launches++;
    } else if (M > 0) {  // radix-8 plan: per channel
      for (size_t i = i0; i < i1; i++) {
        launch_fft2_prep(cjs[i].H, 0, 1, B, cjs[i].P, M, cjs[i].H2, ctx->d_tw2, ctx->d_tab16, kstream(ctx));
        env.launches++;
      }
    }
    i0 = i1;
  }
  CU(cudaGetLastError());
  for (gac_ir* ir : todo) {
    buffer_unref(ir->src);
    ir->src = nullptr;
  }
  if (!ctx->deferred_irs.empty()) {
    auto& L = ctx->deferred_irs;
    L.erase(std::remove_if(L.begin(), L.end(), [](gac_ir* ir) { return ir->prepared; }), L.end());
  }
  return GAC_OK;
}

// Asynchronous uploads: the deferred impulse responses whose samples have LANDED are prepared now (one batch, no waiting), ahead of
// the render that first uses them.  Called from gac_ir_prepare every 32 registrations: while the copy engine delivers the rest of a
// graph's buffers and the host keeps building it, the device — idle until the first render — works off the preparation (C3 shard:
// 0.9 of the 3.6 ms the first render otherwise spends behind the last upload).
static void kick_deferred_irs(gac_context* ctx, size_t min_ready) {
  std::vector<const gac_ir*> ready;
  for (gac_ir* ir : ctx->deferred_irs)
    if (!ir->prepared && ir->src && (!ir->src->ready || cudaEventQuery(ir->src->ready) == cudaSuccess)) ready.push_back(ir);
  cudaGetLastError();
  if (ready.size() < min_ready) return;
  RenderEnv env;
  Scratch scratch(ctx);
  HostKeep keep;
  env.ctx = ctx;
  env.scratch = &scratch;
  env.keep = &keep;
  env.timer = nullptr;
  env.Npad = env.NQ = env.QB = 0;
  const bool was_deferring = ctx->defer_copies;
  ctx->defer_copies = true;  // the job tables travel with the first launch
  const int rc = prepare_irs(env, ready);
  flush_copies(ctx);
  ctx->defer_copies = was_deferring;
  if (rc) {
    cudaGetLastError();  // (the render that uses them reports what is wrong)
    return;
  }
  for (const gac_ir* ir : ready) const_cast<gac_ir*>(ir)->fresh = true;
  auto& L = ctx->deferred_irs;
  L.erase(std::remove_if(L.begin(), L.end(), [](gac_ir* ir) { return ir->prepared; }), L.end());
  // the staged tables live in page-locked blocks that the next render re-uses from the start: it waits for this event first
  if (!ctx->kick_done) cudaEventCreateWithFlags(&ctx->kick_done, cudaEventDisableTiming);
  cudaEventRecord(ctx->kick_done, ctx->stream);
  ctx->kick_pending = true;
}

// Runs every signal's op chain, position by position, batching equal kinds across signals.
static int run_chains(RenderEnv& env, std::vector<Sig>& sigs) {
  gac_context* ctx = env.ctx;
  size_t maxlen = 0;
  for (auto& s : sigs) maxlen = std::max(maxlen, s.ops ? s.ops->size() : 0);
  for (size_t pos = 0; pos < maxlen; pos++) {
    std::vector<size_t> gains, biquads, convs, delays, panners;
    for (size_t i = 0; i < sigs.size(); i++) {
      if (!sigs[i].ops || pos >= sigs[i].ops->size()) continue;
      const OpH& opx = (*sigs[i].ops)[pos];
      if (opx.kind == GAC_OP_GATE) {
        // a connection that exists only from / until a quantum: outside the window the next input sees nothing connected
        Sig& s = sigs[i];
        const int64_t f = std::min<int64_t>(env.Npad, (int64_t)opx.aux * 128);
        if (opx.ftype == 0) s.lo = std::max(s.lo, f); else s.hi = std::min(s.hi, f);
        if (s.hi <= s.lo) s.lo = s.hi = 0;
        continue;
      }
      if (opx.kind == GAC_OP_CONVOLVER && opx.ftype == 1) continue;  // a later epoch: handled with the convolver op it continues
      switch (opx.kind) {
        case GAC_OP_GAIN: gains.push_back(i); break;
        case GAC_OP_BIQUAD: biquads.push_back(i); break;
        case GAC_OP_CONVOLVER: convs.push_back(i); break;
        case GAC_OP_DELAY: delays.push_back(i); break;
        case GAC_OP_PANNER: panners.push_back(i); break;
        default: break;  // GAC_OP_CHANNEL: applied where the chain is fed from its bus (render_core)
      }
    }
    // ---------------- DelayNode: a gather with a per-sample delay (K2 + k_delay), into fresh rows
    if (!delays.empty()) {
      std::vector<ParamJob> pj;
      std::vector<float*> tabs(delays.size(), nullptr);
      for (size_t k = 0; k < delays.size(); k++) {
        int rc = param_table(env, (*sigs[delays[k]].ops)[pos].p0, true, pj, &tabs[k], sigs[delays[k]].bus_base);
        if (rc) return rc;
      }
      int rc = run_param_jobs(env, pj);
      if (rc) return rc;
      float* rows = nullptr;
      if ((rc = env.scratch->alloc(&rows, delays.size() * 2 * (size_t)env.Npad))) return rc;
      auto& dj = env.keep->make<DelayJob>();
      for (size_t k = 0; k < delays.size(); k++) {
        Sig& s = sigs[delays[k]];
        const OpH& op = (*s.ops)[pos];
        DelayJob j;
        j.in[0] = s.p[0];
        j.in[1] = s.p[1];
        j.out[0] = rows + (k * 2 + 0) * (size_t)env.Npad;
        j.out[1] = rows + (k * 2 + 1) * (size_t)env.Npad;
        j.dt = tabs[k];
        j.dt_const = op.p0.value;
        j.max_delay = (int)(op.aux * ctx->fs);  // (int)(maxDelayTime * context.SampleRate)  DelayNode.cs:28
        j.in_lo = s.lo;
        j.in_hi = s.hi;
        dj.push_back(j);
        s.p[0] = j.out[0];
        s.p[1] = j.out[1];
        s.ch = 2;  // Max-mode input with channelCount 2: a mono upstream is up-mixed by copy (AudioNodeInput.cs:157-166,201-213)
        // Flags: the node keeps one pooled block and only ever MARKS it non-silent (DelayNode.cs:96-97): silent until the first
        // block that carries audio, flagged from then on.  With a constant DelayTime that block is known here (assuming the
        // flagged input blocks carry non-zero samples); with automation the range starts conservatively at the input's.
        if (s.hi > s.lo) {
          int64_t first = s.lo;
          if (!j.dt) {
            int d = (int)(j.dt_const * (float)ctx->fs);
            d = d < 0 ? 0 : (d > j.max_delay ? j.max_delay : d);
            first = d >= 1 ? s.lo + d : env.Npad;  // d = 0 reads nothing (CircularBuffer.Read :140-143)
          }
          s.lo = std::min<int64_t>(env.Npad, (first / 128) * 128);
          s.hi = s.lo < env.Npad ? env.Npad : s.lo;
        }
      }
      DelayJob* dd = nullptr;
      if ((rc = env.scratch->upload(&dd, dj))) return rc;
      int t = env.timer->begin(C_DELAY);
      launch_delay(dd, (int)dj.size(), env.Npad, ctx->fs, kstream(ctx));
      env.timer->end(t);
      env.launches += 1;
      CU(cudaGetLastError());
    }
    // ---------------- StereoPannerNode (K2 + k_panner), in place
    if (!panners.empty()) {
      std::vector<ParamJob> pj;
      std::vector<float*> tabs(panners.size(), nullptr);
      for (size_t k = 0; k < panners.size(); k++) {
        int rc = param_table(env, (*sigs[panners[k]].ops)[pos].p0, true, pj, &tabs[k], sigs[panners[k]].bus_base);
        if (rc) return rc;
      }
      int rc = run_param_jobs(env, pj);
      if (rc) return rc;
      auto& qj = env.keep->make<PannerJob>();
      unsigned long long* d_first = nullptr;
      if ((rc = env.scratch->alloc(&d_first, panners.size()))) return rc;
      CU(cudaMemsetAsync(d_first, 0x7f, sizeof(unsigned long long) * panners.size(), ctx->stream));
      bool scan = false;
      for (size_t k = 0; k < panners.size(); k++) {
        Sig& s = sigs[panners[k]];
        PannerJob j;
        j.sig[0] = s.p[0];
        j.sig[1] = s.p[1];
        j.pan = tabs[k];
        j.pan_const = (*s.ops)[pos].p0.value;
        j.mono = s.ch == 1 ? 1 : 0;  // the input is ClampedMax(2): a mono upstream stays mono (StereoPannerNode.cs:24-26, :62-66)
        j.lo = s.lo;
        j.hi = s.hi;
        j.sp_block = -(int64_t)1 << 40;
        j.sp_mode = 0;
        // `first` = the first quantum this node processes (0, or later for a node created between two Render calls): no upstream
        // block exists yet then, so the input takes its own channelCount (2)
        const int64_t first = (int64_t)(*s.ops)[pos].aux * 128;
        if (s.ch == 1 && s.lo <= first && first < s.hi) {
          j.sp_block = first;
          j.sp_mode = 1;
        } else if (s.from_source && pos == 0 && s.ch == 2 && s.lo > first) {  // the source's idle block before its start had one channel
          j.sp_block = s.lo;
          j.sp_mode = 2;
        }
        j.first_change = j.pan ? d_first + k : nullptr;
        scan = scan || (j.pan && j.sp_mode != 0);
        qj.push_back(j);
        s.ch = 2;  // the output block always has two channels (:41-47)
      }
      PannerJob* dq = nullptr;
      if ((rc = env.scratch->upload(&dq, qj))) return rc;
      int t = env.timer->begin(C_PANNER);
      launch_panner(dq, (int)qj.size(), env.Npad, scan, kstream(ctx));
      env.timer->end(t);
      env.launches += scan ? 2 : 1;
      CU(cudaGetLastError());
    }
    // ---------------- GainNode (K2 + K4).  A gain directly followed by a convolver is fused into K5.
    if (!gains.empty()) {
      std::vector<ParamJob> pj;
      std::vector<float*> tabs(gains.size(), nullptr);
      for (size_t k = 0; k < gains.size(); k++) {
        int rc = param_table(env, (*sigs[gains[k]].ops)[pos].p0, true, pj, &tabs[k], sigs[gains[k]].bus_base);
        if (rc) return rc;
      }
      int rc = run_param_jobs(env, pj);
      if (rc) return rc;
      auto& gj = env.keep->make<GainJob>();
      for (size_t k = 0; k < gains.size(); k++) {
        Sig& s = sigs[gains[k]];
        s.ch = 2;  // the node's input is Max-mode with channelCount 2: a mono upstream is up-mixed by copy (AudioNodeInput.cs:157-166,201-213)
        const auto& ops = *s.ops;
        const bool next_is_conv = pos + 1 < ops.size() && ops[pos + 1].kind == GAC_OP_CONVOLVER && ops[pos + 1].ir;
        if (next_is_conv) {
          // fused into the convolver's forward FFT load (K4 inside K5): same float32 multiply, same gating
          env.fused[&s] = std::make_pair((const float*)tabs[k], ops[pos].p0.value);
          continue;
        }
        GainJob g;
        g.sig[0] = s.p[0];
        g.sig[1] = s.p[1];
        g.gain = tabs[k];
        g.gain_const = ops[pos].p0.value;
        g.lo = s.lo;
        g.hi = s.hi;
        gj.push_back(g);
      }
      if (!gj.empty()) {
        GainJob* dg = nullptr;
        if ((rc = env.scratch->upload(&dg, gj))) return rc;
        int t = env.timer->begin(C_GAIN);
        launch_gain(dg, (int)gj.size(), env.Npad, kstream(ctx));
        env.timer->end(t);
        env.launches += 1;
        CU(cudaGetLastError());
      }
    }
    // ---------------- BiQuadFilterNode (K2 + K3)
    TRACE_MARK("chain position: before biquads");
    if (!biquads.empty()) {
      // parameter tables and jobs for every filter of this position
      std::vector<ParamJob> pj;
      std::vector<BiquadJob> all(biquads.size());
      for (size_t k = 0; k < biquads.size(); k++) {
        Sig& s = sigs[biquads[k]];
        s.ch = 2;
        const OpH& op = (*s.ops)[pos];
        BiquadJob j{};
        if (s.lazy[0]) {  // a rate-1 source feeds this filter directly: its frames are read where they lie, no source copy was made
          j.in[0] = s.lazy[0];
          j.in[1] = s.lazy[1];
          s.lazy[0] = s.lazy[1] = nullptr;  // the filter's output is written to the signal's own rows
        }
        float *tf = nullptr, *tq = nullptr, *tg = nullptr;
        int rc;
        if ((rc = param_table(env, op.p0, true, pj, &tf, s.bus_base))) return rc;
        if ((rc = param_table(env, op.p1, true, pj, &tq, s.bus_base))) return rc;
        if ((rc = param_table(env, op.p2, false, pj, &tg, s.bus_base))) return rc;
        j.sig[0] = s.p[0];
        j.sig[1] = s.p[1];
        j.freq = tf;
        j.q = tq;
        j.gain = tg;
        j.freq_const = op.p0.value;
        j.q_const = op.p1.value;
        j.gain_const = op.p2.value;
        j.type = op.ftype;
        j.lo = s.lo;
        j.hi = s.hi;
        all[k] = j;
      }
      TRACE_MARK("biquad: jobs built");
      int rc = run_param_jobs(env, pj);
      if (rc) return rc;
      TRACE_MARK("biquad: params queued");
      // Classes: filters with the same type, the same parameter tables / constants and the same non-silent range compute the same
      // coefficients at every (channel, frame).  A class of at least kSharedMin voices whose input rows are 16-byte aligned takes the
      // shared-coefficient path (biquad.cu); everything else the general one.
      constexpr size_t kSharedMin = 4;
      typedef std::tuple<int, const float*, const float*, const float*, float, float, float, int64_t, int64_t> ClassKey;
      std::map<ClassKey, std::vector<size_t>> classes;
      std::vector<size_t> general;
      for (size_t k = 0; k < all.size(); k++) {
        const BiquadJob& j = all[k];
        const float* x0 = j.in[0] ? j.in[0] : j.sig[0];
        const float* x1 = j.in[1] ? j.in[1] : j.sig[1];
        const bool aligned = reinterpret_cast<uintptr_t>(x0) % 16 == 0 && reinterpret_cast<uintptr_t>(x1) % 16 == 0;
        if (!aligned || j.hi <= j.lo || getenv("GAC_BIQUAD_GENERAL")) {
          general.push_back(k);
          continue;
        }
        classes[ClassKey(j.type, j.freq, j.q, j.gain, j.freq_const, j.q_const, j.gain_const, j.lo, j.hi)].push_back(k);
      }
      auto& sj = env.keep->make<BiquadJob>();   // jobs of the shared path, class after class
      auto& reps = env.keep->make<BiquadJob>();  // one representative per class
      auto& groups = env.keep->make<BqGroup>();  // speculative groups first, then the sequential ones
      std::vector<BqGroup> slow_groups;
      // A filter with constant parameters whose poles lie close to the unit circle forgets its state too slowly for speculative
      // segments to re-join bit for bit (measured: a constant 200 Hz highpass never does).  Pole radius^2 = a2 = (1 - alpha) /
      // (1 + alpha) for every RBJ type at unit shelf gain; the state decays by 2^-24 in 16.6 / -ln(r) frames.
      auto forgets_slowly = [&](const BiquadJob& j) {
        if (j.freq || j.q) return false;  // automated: unknown here, the verified speculation decides
        const double f = std::min(std::max((double)j.freq_const, 1.0), ctx->fs / 2.0), q = std::max((double)j.q_const, 0.001);
        const double w0 = 2.0 * 3.14159265358979323846 * f / ctx->fs, alpha = std::sin(w0) / (2.0 * q);
        const double a2 = (1.0 - alpha) / (1.0 + alpha);
        if (!(a2 > 0.0)) return false;
        const double frames = 16.6 / (-0.5 * std::log(a2));
        return frames > 400.0;
      };
      // Warm-up of the speculative segments (32-frame slabs): a filter whose poles stay well inside the unit circle over the whole
      // automation forgets its state within a few hundred frames, so 8192 frames of warm-up per segment would double the recursion's
      // work for nothing.  Pole radius^2 = a2 = (1 - alpha) / (1 + alpha), alpha = sin(w0) / (2 Q) (divided by the shelf / peaking
      // amplitude in the worst case); the bound is taken over the values the parameters can reach (static value, event values and
      // targets: ramps stay between their end points).  Modulated parameters keep the default.  Only speed depends on this: every
      // segment is verified bit for bit and repaired where it did not re-join.
      auto warm_slabs_of = [&](const OpH& op) -> int {
        auto range = [](const ParamH& p, double& lo, double& hi) {
          if (p.mod_bus >= 0) return false;
          lo = hi = p.value;
          auto take = [&](const std::vector<gac_event>& ev) {
            for (const gac_event& e : ev) {
              const double v = e.type == 3 ? e.target : e.value;
              lo = std::min(lo, v);
              hi = std::max(hi, v);
            }
          };
          take(p.ev);
          for (const auto& ep : p.later) {
            lo = std::min(lo, (double)ep.value);
            hi = std::max(hi, (double)ep.value);
            take(ep.ev);
          }
          return true;
        };
        double f0, f1, q0, q1, g0, g1;
        if (!range(op.p0, f0, f1) || !range(op.p1, q0, q1) || !range(op.p2, g0, g1)) return 0;
        const double nyq = ctx->fs / 2.0, pi = 3.14159265358979323846;
        f0 = std::min(std::max(f0, 1.0), nyq);
        f1 = std::min(std::max(f1, 1.0), nyq);
        const double smin = std::min(std::sin(2.0 * pi * f0 / ctx->fs), std::sin(2.0 * pi * f1 / ctx->fs));  // (sin is concave on [0, pi])
        const double amp = std::pow(10.0, std::max(std::fabs(g0), std::fabs(g1)) / 40.0);
        const double alpha = smin / (2.0 * std::max(q1, 0.001)) / amp;
        const double a2 = (1.0 - alpha) / (1.0 + alpha);
        if (!(alpha > 0.0) || !(a2 > 0.0)) return alpha >= 1.0 ? 32 : 0;
        const double frames = 16.6 / (-0.5 * std::log(a2));  // the state decays by 2^-24
        return frames <= 128.0 ? 32 : frames <= 400.0 ? 128 : 0;
      };
      int warm_hint = -1;  // the longest warm-up any speculative class asks for (0 = the default)
      for (auto& kv : classes) {
        if (kv.second.size() < kSharedMin) {
          general.insert(general.end(), kv.second.begin(), kv.second.end());
          continue;
        }
        const int cls = (int)reps.size();
        reps.push_back(all[kv.second[0]]);
        const bool slow = forgets_slowly(all[kv.second[0]]) && !getenv("GAC_BIQUAD_SPECULATE");
        if (!slow) {
          const int w = warm_slabs_of((*sigs[biquads[kv.second[0]]].ops)[pos]);
          warm_hint = warm_hint < 0 ? w : (w == 0 || warm_hint == 0 ? 0 : std::max(warm_hint, w));
        }
        for (size_t m0 = 0; m0 < kv.second.size(); m0 += 16) {
          const size_t cnt = std::min<size_t>(16, kv.second.size() - m0);
          (slow ? slow_groups : groups).push_back(BqGroup{(int)sj.size(), (int)cnt, cls});
          for (size_t m = 0; m < cnt; m++) sj.push_back(all[kv.second[m0 + m]]);
        }
      }
      const int n_spec_groups = (int)groups.size();
      groups.insert(groups.end(), slow_groups.begin(), slow_groups.end());
      if (!sj.empty()) {
        const int n_cls = (int)reps.size(), n_groups = (int)groups.size();
        const size_t cs_stride = (size_t)(env.Npad / kBqChunkFrames) * kBqChunkBytes;
        int32_t *idx = nullptr, *dlast = nullptr, *dent = nullptr;
        int* dwide = nullptr;
        unsigned char* dcs = nullptr;
        if ((rc = env.scratch->alloc(&idx, (size_t)n_cls * 2 * (size_t)env.Npad))) return rc;
        if ((rc = env.scratch->alloc(&dlast, (size_t)n_cls * 2 * (size_t)env.NQ))) return rc;
        if ((rc = env.scratch->alloc(&dent, (size_t)n_cls * 2 * (size_t)env.NQ))) return rc;
        if ((rc = env.scratch->alloc(&dwide, (size_t)n_cls))) return rc;
        if ((rc = env.scratch->alloc(&dcs, (size_t)n_cls * cs_stride))) return rc;
        for (int c = 0; c < n_cls; c++) reps[c].idx = idx + (size_t)c * 2 * (size_t)env.Npad;
        BiquadJob *dreps = nullptr, *dsj = nullptr;
        BqGroup* dgroups = nullptr;
        if ((rc = env.scratch->upload(&dreps, reps))) return rc;
        if ((rc = env.scratch->upload(&dsj, sj))) return rc;
        if ((rc = env.scratch->upload(&dgroups, groups))) return rc;
        const int n_slow_groups = n_groups - n_spec_groups;
        size_t n_f2 = 0, n_i = 0, n_f2s = 0, n_is = 0;
        if (warm_hint < 0) warm_hint = 0;
        biquad_shared_scratch_sizes(std::max(1, n_spec_groups), env.Npad, &n_f2, &n_i, warm_hint);
        biquad_shared_scratch_sizes(std::max(1, n_slow_groups), env.Npad, &n_f2s, &n_is);  // (an upper bound for the one-segment launch)
        float2 *dstates = nullptr, *dstates_slow = nullptr;
        int* dflags = nullptr;
        if ((rc = env.scratch->alloc(&dstates, n_f2))) return rc;
        if ((rc = env.scratch->alloc(&dstates_slow, n_f2s))) return rc;
        if ((rc = env.scratch->alloc(&dflags, n_i))) return rc;
        int t = env.timer->begin(C_BIQUAD);
        launch_biquad_classes(dreps, n_cls, env.Npad, env.NQ, ctx->fs, dlast, dent, dwide, dcs, cs_stride, kstream(ctx));
        bool partial = false;
        for (const BiquadJob& j : sj) partial = partial || j.lo > 0 || j.hi < env.Npad;
        if (partial) launch_biquad_zero_outside(dsj, (int)sj.size(), env.Npad, kstream(ctx));
        launch_biquad_lanes_shared(dsj, dgroups, n_spec_groups, dcs, cs_stride, env.Npad, dstates, dflags, false, kstream(ctx), warm_hint);
        launch_biquad_lanes_shared(dsj, dgroups + n_spec_groups, n_slow_groups, dcs, cs_stride, env.Npad, dstates_slow, nullptr, true, kstream(ctx));
        env.timer->end(t);
        env.launches += 3 + (partial ? 1 : 0) + (n_slow_groups > 0 ? 1 : 0) +
                        (n_spec_groups > 0 ? 1 + (biquad_shared_segments(n_spec_groups, env.Npad, nullptr, warm_hint) > 1 ? 2 : 0) : 0);
        CU(cudaGetLastError());
        TRACE_MARK("biquad (shared coefficients): queued");
      }
      const size_t per_job = (size_t)env.Npad * (2 * (4 + 16 + 16) + 4 + 4) + (size_t)env.NQ * (16 + 4);
      const size_t max_jobs = std::min<size_t>(65535, std::max<size_t>(1, ctx->scratch_budget / per_job));
      for (size_t k0 = 0; k0 < general.size(); k0 += max_jobs) {
        const size_t nk = std::min(max_jobs, general.size() - k0);
        auto& bj = env.keep->make<BiquadJob>();
        int32_t* idx_all = nullptr;
        float4 *s1_all = nullptr, *s2_all = nullptr;
        {
          const size_t rows = ((nk + 15) / 16) * 32;  // slab-transposed streams cover whole 32-row groups
          if ((rc = env.scratch->alloc(&idx_all, nk * 2 * (size_t)env.Npad))) return rc;
          if ((rc = env.scratch->alloc(&s1_all, rows * (size_t)env.Npad))) return rc;
          if ((rc = env.scratch->alloc(&s2_all, rows * (size_t)env.Npad))) return rc;
        }
        for (size_t k = 0; k < nk; k++) {
          BiquadJob j = all[general[k0 + k]];
          j.idx = idx_all + k * 2 * (size_t)env.Npad;
          bj.push_back(j);
        }
        BiquadJob* dbj = nullptr;
        int32_t *dlast = nullptr, *dent = nullptr;
        if ((rc = env.scratch->upload(&dbj, bj))) return rc;
        if ((rc = env.scratch->alloc(&dlast, nk * 2 * (size_t)env.NQ))) return rc;
        if ((rc = env.scratch->alloc(&dent, nk * 2 * (size_t)env.NQ))) return rc;
        float2* dstates = nullptr;
        int* dbad = nullptr;
        const int n_seg = biquad_lane_segments((int)nk, env.Npad, nullptr);
        size_t n_f2 = 0, n_i = 0;
        biquad_scratch_sizes((int)nk, env.Npad, &n_f2, &n_i);
        if ((rc = env.scratch->alloc(&dstates, n_f2))) return rc;
        if ((rc = env.scratch->alloc(&dbad, n_i))) return rc;
        int t = env.timer->begin(C_BIQUAD);
        launch_biquad(dbj, (int)nk, env.Npad, env.NQ, ctx->fs, dlast, dent, s1_all, s2_all, dstates, dbad, kstream(ctx));
        env.timer->end(t);
        env.launches += n_seg > 1 ? 9 : 6;
        CU(cudaGetLastError());
        TRACE_MARK("biquad: queued");
      }
    }
    // ---------------- ConvolverNode (K5, K6, K7)
    if (!convs.empty()) {
      // the epochs of a ConvolverNode whose Buffer was set again between Render calls follow its op (filter_type 1)
      auto epochs_of = [&](const Sig& s) {
        std::vector<const OpH*> e;
        e.push_back(&(*s.ops)[pos]);
        for (size_t q = pos + 1; q < s.ops->size() && (*s.ops)[q].kind == GAC_OP_CONVOLVER && (*s.ops)[q].ftype == 1; q++) e.push_back(&(*s.ops)[q]);
        return e;
      };
      {
        std::vector<const gac_ir*> need, reused;
        for (size_t k : convs)
          for (const OpH* op : epochs_of(sigs[k]))
            if (const gac_ir* ir = op->ir) {
              need.push_back(ir);
              // double-length spectra pay for themselves only when the impulse response serves more than one render: an IR whose
              // (deferred) preparation happens in this very render keeps the single-length plan
              if (ir->prepared && !ir->fresh) reused.push_back(ir);
              const_cast<gac_ir*>(ir)->fresh = false;
            }
        int rc = prepare_irs(env, need);
        if (rc) return rc;
        if (ctx->mixed_segments && (rc = ensure_h2b_batch(env, reused))) return rc;
      }
      std::vector<ConvItem> items;
      std::vector<size_t> item_sig;  // which signal an item belongs to
      auto& zj = env.keep->make<GainJob>();
      struct Window {  // what a later epoch contributes: frames [w0, w1) of `from` (null: silence) replace the signal's rows
        float* dst[2];
        const float* from[2];
        int64_t w0, w1;
      };
      std::vector<Window> windows;
      for (size_t k = 0; k < convs.size(); k++) {
        Sig& s = sigs[convs[k]];
        const std::vector<const OpH*> eps = epochs_of(s);
        const float* gtab = nullptr;
        float gconst = 1.0f;
        auto f = env.fused.find(&s);
        if (f != env.fused.end()) {
          gtab = f->second.first;
          gconst = f->second.second;
          env.fused.erase(f);
        }
        const float* in0 = s.lazy[0] ? s.lazy[0] : s.p[0];
        const float* in1 = s.lazy[0] ? s.lazy[1] : s.p[1];
        s.lazy[0] = s.lazy[1] = nullptr;  // the convolver's output is written to the signal's own rows
        const int in_ch = s.ch;
        const int64_t in_lo = s.lo, in_hi = s.hi;
        int64_t out_lo = env.Npad, out_hi = 0;
        int layout = -1;
        for (size_t e = 0; e < eps.size(); e++) {
          const OpH& op = *eps[e];
          const int64_t w0 = std::min<int64_t>(env.Npad, (int64_t)op.aux * 128);
          const int64_t w1 = e + 1 < eps.size() ? std::min<int64_t>(env.Npad, (int64_t)eps[e + 1]->aux * 128) : env.Npad;
          float* out0 = s.p[0];
          float* out1 = s.p[1];
          // with several epochs EVERY epoch renders into rows of its own and its window is copied back afterwards: the signal's rows
          // are the input of all of them, and epochs of different impulse-response lengths run in different kernel batches
          if (eps.size() > 1) {
            if (w1 <= w0) continue;
            if (op.ir) {
              float* tmp = nullptr;
              int rc = env.scratch->alloc(&tmp, 2 * (size_t)env.Npad);
              if (rc) return rc;
              out0 = tmp;
              out1 = tmp + env.Npad;
            }
            windows.push_back(Window{{s.p[0], s.p[1]}, {op.ir ? out0 : nullptr, op.ir ? out1 : nullptr}, w0, w1});
          }
          if (!op.ir) {
            if (eps.size() == 1) {
              // ConvolverNode without a Buffer clears its output (ConvolverNode.cs:107-119): silence (until a later epoch)
              GainJob g;
              g.sig[0] = s.p[0];
              g.sig[1] = s.p[1];
              g.gain = nullptr;
              g.gain_const = 0.f;
              g.lo = 0;
              g.hi = 0;
              zj.push_back(g);
            }
            continue;
          }
          const gac_ir* ir = op.ir;
          const int lay = ir->true_stereo ? 4 : ir->nch;
          if (layout >= 0 && lay != layout)
            return fail(GAC_ERR_UNSUPPORTED, "the impulse responses a ConvolverNode is given over time must share the channel layout on the accelerated path");
          layout = lay;
          out_lo = std::min(out_lo, w0);
          out_hi = std::max(out_hi, w1);
          auto Hch = [&](int c) { return (const float2*)(ir->d_H + (size_t)c * ir->P16 * ctx->B); };
          auto H2ch = [&](int c) { return ir->d_H2 ? (const float2*)(ir->d_H2 + (size_t)c * (ctx->B + 1) * fft2_h2_row_elems(ir->M2)) : (const float2*)nullptr; };
          ConvItem it;
          it.P = ir->P;
          it.M2 = ir->d_H2 ? ir->M2 : 0;
          it.Lh = ir->Lh;
          const float2* h2b = nullptr;
          if (it.M2 > 0 && ctx->mixed_segments) {
            plan_segments(env.QB, it.M2, it.Lh, &it.n_big, &it.n_small);
            if (it.n_big > 0) h2b = ir->d_H2b;  // (prepared above for the whole batch)
            if (!h2b) it.n_big = 0;
          }
          auto H2bch = [&](int c) { return h2b ? h2b + (size_t)c * (ctx->B + 1) * fft2_h2_row_elems(2 * ir->M2) : (const float2*)nullptr; };
          // an epoch's convolvers are created when the Buffer is set: they see the input from their first quantum on
          it.lo = std::max(in_lo, w0);
          it.hi = in_hi;
          if (it.hi <= it.lo) it.lo = it.hi = 0;
          it.gain_tab = gtab;
          it.gain_const = gconst;
          if (ir->nch == 1) {
            // input forced to 1 channel (ConvolverNode.cs:72-76): a stereo upstream is down-mixed (L + R) * (1/sqrt(2)),
            // a mono upstream is taken as is (AudioNodeInput.cs:188-228); the mono result is copied to both rows
            it.n_fwd = 1;
            it.fwd[0] = {in0, in_ch == 2 ? in1 : nullptr, 1.0f / sqrtf(2.0f)};
            it.n_mac = 1;
            it.mac[0] = {0, Hch(0), H2ch(0), H2bch(0)};
            it.n_inv = 1;
            it.inv[0] = {0, -1, out0, out1};
            s.ch = 1;
          } else if (ir->nch == 2) {
            it.n_fwd = 2;
            it.fwd[0] = {in0, nullptr, 1.0f};
            it.fwd[1] = {in1, nullptr, 1.0f};
            it.n_mac = 2;
            it.mac[0] = {0, Hch(0), H2ch(0), H2bch(0)};
            it.mac[1] = {1, Hch(1), H2ch(1), H2bch(1)};
            it.n_inv = 2;
            it.inv[0] = {0, -1, out0, nullptr};
            it.inv[1] = {1, -1, out1, nullptr};
            s.ch = 2;
            // fan-in fusion: this convolver ends its chain and its only consumer is a plain fan-in
            // (k_fft2_sum16 exists for the radix-16 plans: impulse responses that need 8192-point segments keep the plain path)
            if (ctx->fuse_fanin && s.sum_key >= 0 && eps.size() == 1 && pos + 1 == s.ops->size() && it.M2 > 0 && it.M2 <= fft2_r16_max() && it.mac[0].H2 &&
                it.mac[1].H2)
              it.sum_key = s.sum_key;
          } else {
            // true stereo (ConvolverNode.cs:127-144): L = c0(inL) + c2(inR), R = c1(inL) + c3(inR); X spectra shared
            it.n_fwd = 2;
            it.fwd[0] = {in0, nullptr, 1.0f};
            it.fwd[1] = {in1, nullptr, 1.0f};
            it.n_mac = 4;
            it.mac[0] = {0, Hch(0), H2ch(0), H2bch(0)};
            it.mac[1] = {1, Hch(2), H2ch(2), H2bch(2)};
            it.mac[2] = {0, Hch(1), H2ch(1), H2bch(1)};
            it.mac[3] = {1, Hch(3), H2ch(3), H2bch(3)};
            it.n_inv = 2;
            it.inv[0] = {0, 1, out0, nullptr};
            it.inv[1] = {2, 3, out1, nullptr};
            s.ch = 2;
          }
          items.push_back(it);
          item_sig.push_back(convs[k]);
        }
        // ConvolverNode always marks its output non-silent while it has convolvers (ConvolverNode.cs:153), and clears it while it
        // has none (:107-119); over several epochs the flagged range is the hull of the epochs that had a Buffer
        s.lo = out_hi > out_lo ? out_lo : 0;
        s.hi = out_hi > out_lo ? out_hi : 0;
      }
      if (!zj.empty()) {
        GainJob* dz = nullptr;
        int rc = env.scratch->upload(&dz, zj);
        if (rc) return rc;
        launch_gain(dz, (int)zj.size(), env.Npad, kstream(ctx));
        env.launches += 1;
      }
      {
        // fan-in groups: at least two members with the same consumer and segment plan; the first member's rows receive the sum,
        // the others carry nothing from here on (a silent input is never mixed, AudioNodeInput.cs:127)
        std::map<std::tuple<int64_t, int, int, int>, std::vector<size_t>> groups;
        for (size_t i = 0; i < items.size(); i++)
          if (items[i].sum_key >= 0) groups[std::make_tuple(items[i].sum_key, items[i].M2, items[i].Lh, items[i].n_big)].push_back(i);
        for (auto& kv : groups) {
          if (kv.second.size() < 2) {
            items[kv.second[0]].sum_key = -1;
            continue;
          }
          Sig& lead = sigs[item_sig[kv.second[0]]];
          for (size_t m = 1; m < kv.second.size(); m++) {
            Sig& s = sigs[item_sig[kv.second[m]]];
            if (s.hi > s.lo) {
              lead.lo = lead.hi > lead.lo ? std::min(lead.lo, s.lo) : s.lo;
              lead.hi = std::max(lead.hi, s.hi);
            }
            s.lo = s.hi = 0;
          }
        }
      }
      TRACE_MARK("conv: items built");
      int rc = conv_batch(env, items);
      if (rc) return rc;
      for (const Window& w : windows)  // later epochs take over from their first quantum on
        for (int c = 0; c < 2; c++) {
          if (w.from[c]) CU(cudaMemcpyAsync(w.dst[c] + w.w0, w.from[c] + w.w0, sizeof(float) * (size_t)(w.w1 - w.w0), cudaMemcpyDeviceToDevice, ctx->stream));
          else CU(cudaMemsetAsync(w.dst[c] + w.w0, 0, sizeof(float) * (size_t)(w.w1 - w.w0), ctx->stream));
        }
      TRACE_MARK("conv: queued");
    }
  }
  return GAC_OK;
}

// source stage, render drivers, NCCL bus reduce and the kernel-level entry points
#include "engine_render.inl"
#include "engine_stream.inl"
