// nodes.cu — the per-quantum node work around the convolver, batched over voices (K1-K4, mix).
// Compiled with --fmad=false: every float/double op below rounds separately, exactly like the reference's
// scalar C# (no FMA contraction), so these kernels are bit-faithful to the CPU oracle wherever the same
// libm results are available.
//
//   K1  k_source_copy / k_resample    AudioBufferSourceNode.Process (Nodes/AudioBufferSourceNode.cs:131-376),
//                                     CubicResampler.Process (CubicResampler.cs:26-63)
//   K2  k_param_eval                  AudioParam.ComputeARate/ComputeKRate/ComputeValueAtTime (AudioParam.cs:114-247)
//   K3  k_biquad_select/_coef/_lanes  BiQuadFilterNode.Process/UpdateCoefficients (Nodes/BiQuadFilterNode.cs:87-258)
//   K4  k_gain                        GainNode.Process (Nodes/GainNode.cs:29-61)
//   mix k_mix                         AudioNodeInput.Pull/MixBuffer (AudioNodeInput.cs:100-138,182-244)
//   K0  k_ir_scale                    PartitionedConvolver.CalculateNormalizationScale (PartitionedConvolver.cs:93-102)
#include <math_constants.h>
#include <algorithm>
#include <cstdint>

#include "gac_kernels.h"
#include "biquad_math.cuh"

namespace gac {

// ============================================================================================ K2
__device__ __forceinline__ float interp_linear(float v0, double t0, float v1, double t1, double t) {  // AudioParam.cs:220-225
  double u = (t - t0) / (t1 - t0);
  u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
  float d = v1 - v0;
  return (float)((double)v0 + (double)d * u);
}
// v0 * Math.Pow(v1 / v0, u) (:236) evaluated as v0 * exp(u * log(ratio)) with log(ratio) computed once per event interval
// (`lr`, cached by the caller): within 2 ulp(double) of pow, i.e. the same float32 except on rounding ties, like the device's
// own pow against glibc's; half the FP64 work of a pow per frame.
__device__ __forceinline__ float interp_exponential(float v0, double t0, float v1, double t1, double t, double lr) {  // :228-237
  double u = (t - t0) / (t1 - t0);
  u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
  return (float)((double)v0 * exp(u * lr));
}
__device__ __forceinline__ float set_target(const DevEvent& e, float base, double t) {  // :240-247
  double elapsed = t - e.time;
  if (elapsed <= 0.0) return base;
  double tc = e.time_constant > 0.001 ? e.time_constant : 0.001;
  float d = base - e.target;
  return (float)((double)e.target + (double)d * exp(-elapsed / tc));
}
// ComputeValueAtTime AudioParam.cs:169-217, split into "which event interval is `time` in" and "evaluate inside it" so that a
// thread walking consecutive samples scans the event list once: `i` = first event with time < e.time (count = past the last),
// `boundary` = value of the last non-SetTarget event before i (the param's static value if there is none).
__device__ __forceinline__ void advance_interval(const DevEvent* __restrict__ ev, int count, double time, int& i, float& boundary) {
  while (i < count && !(time < ev[i].time)) {
    if (ev[i].type != 3) boundary = ev[i].value;
    i++;
  }
}
// A thread walks 16 consecutive frames of one quantum.  Inside an exponential ramp or a SetTarget curve the exponential of frame
// n + 1 is the exponential of frame n times a constant (the frames are dt apart), so only the first frame of the thread in an
// interval pays for exp(): `Rec` carries e^(..) and its per-frame factor.  15 multiplications add ~2e-15 of relative error to a
// number that is rounded to float32 (6e-8): the same float as the direct evaluation except on one sample in ~1e7.
struct Rec {
  int i = -2;       // interval the recurrence belongs to (-2: none)
  int kind = 0;     // 2 exponential ramp, 3 SetTarget
  double val = 0.0, mul = 1.0;
};
__device__ __forceinline__ float eval_interval(float value, const DevEvent* __restrict__ ev, int count, int i, float boundary, double time,
                                               double dt, Rec& rec) {
  if (count == 0) return value;
  const DevEvent* tgt = nullptr;  // the SetTarget event in force, if any
  if (i < count) {
    if (i == 0) return boundary;
    const DevEvent e = ev[i];
    const DevEvent prev = ev[i - 1];
    if (e.type == 1) return interp_linear(prev.value, prev.time, e.value, e.time, time);
    if (e.type == 2) {
      if (prev.value <= 0.f || e.value <= 0.f) return interp_linear(prev.value, prev.time, e.value, e.time, time);  // :230-231
      if (rec.i == i && rec.kind == 2) {
        rec.val *= rec.mul;
      } else {
        const float ratio = e.value / prev.value;  // formed in float32 (:235)
        const double lr = log((double)ratio);
        double u = (time - prev.time) / (e.time - prev.time);
        u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
        rec.val = exp(u * lr);  // Math.Pow(v1 / v0, u) (:236) as exp(u log r): within 2 ulp(double) of pow
        rec.mul = exp(lr * (dt / (e.time - prev.time)));
        rec.i = i;
        rec.kind = 2;
      }
      return (float)((double)prev.value * rec.val);
    }
    if (prev.type != 3) return prev.value;
    tgt = &ev[i - 1];
  } else {
    if (ev[count - 1].type != 3) return ev[count - 1].value;
    tgt = &ev[count - 1];
  }
  // ComputeSetTargetFromBaseline :240-247
  const DevEvent t3 = *tgt;
  const double elapsed = time - t3.time;
  if (elapsed <= 0.0) {
    rec.i = -2;
    return boundary;
  }
  const double tc = t3.time_constant > 0.001 ? t3.time_constant : 0.001;
  if (rec.i == i && rec.kind == 3) {
    rec.val *= rec.mul;
  } else {
    rec.val = exp(-elapsed / tc);
    rec.mul = exp(-dt / tc);
    rec.i = i;
    rec.kind = 3;
  }
  const float d = boundary - t3.target;
  return (float)((double)t3.target + (double)d * rec.val);
}

// a-rate: a thread evaluates 16 consecutive frames (four float4 stores); k-rate: one quantum per thread
constexpr int kParamFrames = 16;
constexpr int kParamSmemEvents = 32;  // event lists up to this length are staged in shared memory (32 bytes per event)
static_assert(sizeof(DevEvent) == 32, "DevEvent is staged as two int4 words");
__global__ void __launch_bounds__(256) k_param_eval(const ParamJob* __restrict__ jobs, const double* __restrict__ block_time,
                                                    int64_t n_quanta, int sample_rate) {
  const ParamJob job = jobs[blockIdx.y];
  // The CTA's threads walk the same event list: short lists are staged in shared memory first (resolving the interval is a chain of
  // dependent loads — long-scoreboard stalls were the kernel's top stall reason with the list in global memory).
  __shared__ int4 s_ev[2 * kParamSmemEvents];
  const DevEvent* __restrict__ ev = job.events;
  if (job.n_events > 0 && job.n_events <= kParamSmemEvents) {  // (uniform over the CTA)
    const int4* __restrict__ src = reinterpret_cast<const int4*>(job.events);
    for (int w = threadIdx.x; w < 2 * job.n_events; w += 256) s_ev[w] = src[w];
    __syncthreads();
    ev = reinterpret_cast<const DevEvent*>(s_ev);
  }
  const double dt = 1.0 / (double)sample_rate;  // AudioParam.cs:116
  int i = 0;
  Rec rec;
  float boundary = job.value;
  if (job.a_rate) {
    const int64_t n = ((int64_t)blockIdx.x * 256 + threadIdx.x) * kParamFrames;
    if (n >= n_quanta * 128 || (n >> 7) < job.q_lo || (n >> 7) >= job.q_hi) return;
    const double t0 = block_time[n >> 7];  // the frames share a quantum (16 divides 128)
    const int f0 = (int)(n & 127);
    // Fast path: no event time falls between this thread's first and last frame (all but a handful of threads per event).  The
    // interval is resolved once, its constants are hoisted, and a frame costs a few FP64 operations; same arithmetic per frame as
    // eval_interval, the quotient (t - t0) / (t1 - t0) of a linear ramp through one IEEE reciprocal per thread and a Markstein
    // correction step (q = a r; q += fma(-q, b, a) r: the correctly rounded quotient).
    {
      const double t_first = t0 + (double)f0 * dt, t_last = t0 + (double)(f0 + kParamFrames - 1) * dt;
      advance_interval(ev, job.n_events, t_first, i, boundary);
      if (i >= job.n_events || t_last < ev[i].time) {
        const int count = job.n_events;
        int mode = 0;  // 0 constant, 1 linear, 2 exponential ramp, 3 SetTarget
        float cval = job.value, v0 = 0.f, dv = 0.f, tgt = 0.f;
        double te = 0.0, span = 1.0, rspan = 1.0, lr = 0.0, tc = 1.0;
        if (count > 0) {
          const DevEvent* st = nullptr;
          if (i < count) {
            if (i == 0) {
              cval = boundary;
            } else {
              const DevEvent e = ev[i], prev = ev[i - 1];
              if (e.type == 1 || (e.type == 2 && (prev.value <= 0.f || e.value <= 0.f))) {
                mode = 1;
                v0 = prev.value;
                dv = e.value - prev.value;
                te = prev.time;
                span = e.time - prev.time;
                rspan = 1.0 / span;
              } else if (e.type == 2) {
                mode = 2;
                v0 = prev.value;
                te = prev.time;
                span = e.time - prev.time;
                lr = log((double)(e.value / prev.value));  // ratio formed in float32 (:235)
              } else if (prev.type != 3) {
                cval = prev.value;
              } else {
                st = &ev[i - 1];
              }
            }
          } else if (ev[count - 1].type != 3) {
            cval = ev[count - 1].value;
          } else {
            st = &ev[count - 1];
          }
          if (st) {
            mode = 3;
            te = st->time;
            tgt = st->target;
            dv = boundary - st->target;
            tc = st->time_constant > 0.001 ? st->time_constant : 0.001;
          }
        }
        // mode-specialised loops (small code: the transcendental calls sit outside of them)
        float v[kParamFrames];
        if (mode == 0) {
#pragma unroll
          for (int k = 0; k < kParamFrames; k++) v[k] = cval;
        } else if (mode == 1) {
          const double dv0 = (double)v0, ddv = (double)dv;
#pragma unroll
          for (int k = 0; k < kParamFrames; k++) {
            const double a = (t0 + (double)(f0 + k) * dt) - te;  // :120
            double u = a * rspan;
            u = fma(fma(-u, span, a), rspan, u);  // (in [0, 1): prev.time <= t < e.time, the clamp of :222 is a no-op)
            v[k] = (float)(dv0 + ddv * u);
          }
        } else if (mode == 2) {
          double u = (t_first - te) / span;
          u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
          double val = exp(u * lr);
          const double mul = exp(lr * (dt / span)), dv0 = (double)v0;
#pragma unroll
          for (int k = 0; k < kParamFrames; k++) {
            v[k] = (float)(dv0 * val);
            val *= mul;
          }
        } else {
          // elapsed = t - te >= 0 here; it is 0 only on a first frame that coincides with the event, which keeps the baseline
          int k0 = 0;
          if (t_first - te <= 0.0) {
            v[0] = boundary;
            k0 = 1;
          }
          double val = exp(-((t0 + (double)(f0 + k0) * dt) - te) / tc);
          const double mul = exp(-dt / tc), dt0 = (double)tgt, ddv = (double)dv;
#pragma unroll
          for (int k = 0; k < kParamFrames; k++) {
            if (k >= k0) {
              v[k] = (float)(dt0 + ddv * val);
              val *= mul;
            }
          }
        }
#pragma unroll
        for (int e4 = 0; e4 < kParamFrames / 4; e4++)
          *reinterpret_cast<float4*>(job.out + n + 4 * e4) = make_float4(v[4 * e4], v[4 * e4 + 1], v[4 * e4 + 2], v[4 * e4 + 3]);
        return;
      }
      i = 0;  // an event inside the thread's frames: the general walk below, from the start
      boundary = job.value;
    }
#pragma unroll 1
    for (int e4 = 0; e4 < kParamFrames / 4; e4++) {
      float v[4];
#pragma unroll 1
      for (int e = 0; e < 4; e++) {
        const double t = t0 + (double)((int)(n & 127) + 4 * e4 + e) * dt;  // :120
        const int before = i;
        advance_interval(ev, job.n_events, t, i, boundary);
        if (i != before) rec.i = -2;  // a new interval starts its own recurrence
        v[e] = eval_interval(job.value, ev, job.n_events, i, boundary, t, dt, rec);
      }
      *reinterpret_cast<float4*>(job.out + n + 4 * e4) = make_float4(v[0], v[1], v[2], v[3]);
    }
  } else {
    const int64_t b = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (b >= n_quanta || b < job.q_lo || b >= job.q_hi) return;
    const double t = block_time[b];
    advance_interval(ev, job.n_events, t, i, boundary);
    job.out[b] = eval_interval(job.value, ev, job.n_events, i, boundary, t, dt, rec);  // ComputeKRate :144-146
  }
}

void launch_param_eval(const ParamJob* d_jobs, int n_jobs, const double* d_block_time, int64_t n_quanta, int sample_rate, cudaStream_t s) {
  if (n_jobs <= 0 || n_quanta <= 0) return;
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
    dim3 grid((unsigned)((n_quanta * (128 / kParamFrames) + 255) / 256), (unsigned)nj);  // a-rate: 16 frames per thread (k-rate jobs use the first CTAs)
    k_param_eval<<<grid, 256, 0, s>>>(d_jobs + j0, d_block_time, n_quanta, sample_rate);
  }
}

// ============================================================================================ K1
__global__ void __launch_bounds__(256) k_source_copy(const SourceJob* __restrict__ jobs, int64_t n_frames) {
  const SourceJob job = jobs[blockIdx.y];
  int64_t n4 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (n4 >= n_frames) return;
#pragma unroll
  for (int c = 0; c < 2; c++) {
    float4 v;
    float* pv = &v.x;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      int64_t m = n4 + u - job.out0;
      int64_t k = job.pos0 + m;
      if (job.loop_len > 0 && k >= job.loop_end)
        k = (job.pos0 >= job.loop_end && m < 128) ? job.loop_start + m % job.loop_len : job.loop_start + (k - job.loop_end) % job.loop_len;
      pv[u] = (m >= 0 && m < job.n_emit) ? job.src[c][k] : 0.f;
    }
    *reinterpret_cast<float4*>(job.dst[c] + n4) = v;
  }
}
void launch_source_copy(const SourceJob* d_jobs, int n_jobs, int64_t n_frames, cudaStream_t s) {
  if (n_jobs <= 0 || n_frames <= 0) return;
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
    dim3 grid((unsigned)((n_frames / 4 + 255) / 256), (unsigned)nj);
    k_source_copy<<<grid, 256, 0, s>>>(d_jobs + j0, n_frames);
  }
}

__global__ void __launch_bounds__(256) k_resample(const ResampleJob* __restrict__ jobs, int64_t n_frames) {
  const ResampleJob job = jobs[blockIdx.y];
  int64_t n = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (n >= n_frames) return;
  int64_t m = n - job.out0;
  const bool live = m >= 0 && m < job.n_emit && m < job.n_zero_from;
  int k = 0;
  float t = 0.f;
  if (live) {
    k = job.k[m];
    t = job.t[m];
  }
  int4 w = make_int4(k, k + 1, k + 2, k + 3);
  bool live_w = live;
  if (k < 0) {  // looping sources: a cleared frame (AudioBufferSourceNode.cs:334-338) or a window that straddles the loop seam
    live_w = live && k != kResampleCleared;
    if (live_w) w = *reinterpret_cast<const int4*>(job.x + 4 * (int64_t)(-(k + 1)));
  }
#pragma unroll
  for (int c = 0; c < 2; c++) {
    float y = 0.f;
    if (live_w) {
      const float* p = job.src[c];
      float S0 = p[w.x], S1 = p[w.y], S2 = p[w.z], S3 = p[w.w];
      // CubicResampler.cs:52-57 (float32, left to right, unfused)
      y = S1 + t * (0.5f * (S2 - S0) + t * ((S0 - 2.5f * S1 + 2.f * S2 - 0.5f * S3) + t * (0.5f * (S3 - S0) + 1.5f * (S1 - S2))));
    }
    job.dst[c][n] = y;
  }
}
void launch_resample(const ResampleJob* d_jobs, int n_jobs, int64_t n_frames, cudaStream_t s) {
  if (n_jobs <= 0 || n_frames <= 0) return;
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
    dim3 grid((unsigned)((n_frames + 255) / 256), (unsigned)nj);
    k_resample<<<grid, 256, 0, s>>>(d_jobs + j0, n_frames);
  }
}

// ============================================================================================ K4
__global__ void __launch_bounds__(256) k_gain(const GainJob* __restrict__ jobs, int64_t n_frames) {
  const GainJob job = jobs[blockIdx.y];
  int64_t n4 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (n4 >= n_frames) return;
  float4 g = make_float4(job.gain_const, job.gain_const, job.gain_const, job.gain_const);
  if (job.gain) g = *reinterpret_cast<const float4*>(job.gain + n4);
  const float* pg = &g.x;
#pragma unroll
  for (int c = 0; c < 2; c++) {
    float4 v = *reinterpret_cast<float4*>(job.sig[c] + n4);
    float* pv = &v.x;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      int64_t n = n4 + u;
      pv[u] = (n >= job.lo && n < job.hi) ? pv[u] * pg[u] : 0.f;  // silent-flagged input -> cleared output (GainNode.cs:41-46)
    }
    *reinterpret_cast<float4*>(job.sig[c] + n4) = v;
  }
}
void launch_gain(const GainJob* d_jobs, int n_jobs, int64_t n_frames, cudaStream_t s) {
  if (n_jobs <= 0 || n_frames <= 0) return;
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
    dim3 grid((unsigned)((n_frames / 4 + 255) / 256), (unsigned)nj);
    k_gain<<<grid, 256, 0, s>>>(d_jobs + j0, n_frames);
  }
}

// ============================================================================================ delay / stereo panner
// DelayNode.Process (Nodes/DelayNode.cs:43-100): per sample d = clamp((int)(delayTime[i] * sampleRate), 0, max); the ring is read
// BEFORE the input sample is written, and Read returns 0 for d <= 0 (:140-147) — so out[n] = d >= 1 ? x[n - d] : 0, with x = 0
// before the render starts and on silent-flagged input blocks (which write zeros, :63-75).  A gather: nothing recursive.
__global__ void __launch_bounds__(256) k_delay(const DelayJob* __restrict__ jobs, int64_t n_frames, int sample_rate) {
  const DelayJob job = jobs[blockIdx.y];
  const int64_t n4 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;  // 4 frames per thread: 16-byte stores, neighbouring gathers
  if (n4 >= n_frames) return;
  float4 t4 = make_float4(job.dt_const, job.dt_const, job.dt_const, job.dt_const);
  if (job.dt) t4 = *reinterpret_cast<const float4*>(job.dt + n4);
  const float* pt = &t4.x;
  float4 o0, o1;
  float* p0 = &o0.x;
  float* p1 = &o1.x;
#pragma unroll
  for (int u = 0; u < 4; u++) {
    int d = (int)(pt[u] * (float)sample_rate);  // float * int -> float product, truncated (:68,85)
    d = d < 0 ? 0 : (d > job.max_delay ? job.max_delay : d);
    const int64_t m = n4 + u - d;
    const bool live = d >= 1 && m >= job.in_lo && m < job.in_hi;
    p0[u] = live ? job.in[0][m] : 0.f;
    p1[u] = live ? job.in[1][m] : 0.f;
  }
  *reinterpret_cast<float4*>(job.out[0] + n4) = o0;
  *reinterpret_cast<float4*>(job.out[1] + n4) = o1;
}
void launch_delay(const DelayJob* d_jobs, int n_jobs, int64_t n_frames, int sample_rate, cudaStream_t s) {
  if (n_jobs <= 0 || n_frames <= 0) return;
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
    k_delay<<<dim3((unsigned)((n_frames / 4 + 255) / 256), (unsigned)nj), 256, 0, s>>>(d_jobs + j0, n_frames, sample_rate);
  }
}

// StereoPannerNode.Process (Nodes/StereoPannerNode.cs:36-153).  The node's cached gains are a pure function of the clamped pan
// value (recomputed whenever it changes, :95-103 / :129-137), so every sample is independent; MathF.Cos / MathF.Sin are the
// platform libm's cosf / sinf (sincosf_libm), the products and sums stay unfused (--fmad=false).
__global__ void __launch_bounds__(256) k_panner(const PannerJob* __restrict__ jobs, int64_t n_frames) {
  const PannerJob job = jobs[blockIdx.y];
  const int64_t n4 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;  // 4 frames per thread (a quantum never straddles a thread)
  if (n4 >= n_frames) return;
  float4 L4 = make_float4(0.f, 0.f, 0.f, 0.f), R4 = L4;
  if (n4 >= job.lo && n4 < job.hi) {  // silent-flagged input -> cleared output (:49-54); lo / hi are multiples of 128
    float4 pan4 = make_float4(job.pan_const, job.pan_const, job.pan_const, job.pan_const);
    if (job.pan) pan4 = *reinterpret_cast<const float4*>(job.pan + n4);
    const float4 a4 = *reinterpret_cast<const float4*>(job.sig[0] + n4), b4 = *reinterpret_cast<const float4*>(job.sig[1] + n4);
    const float kPi = 3.14159265358979323846f;
    // The input's channel count is computed from the upstream block of the PREVIOUS quantum (AudioNodeInput.cs:109 precedes :124),
    // so one block can differ from the rest: a mono signal is up-mixed to two equal channels in the very first quantum (mode 1,
    // rows are already duplicates), a stereo source that starts later is mixed DOWN to one channel in its first quantum (mode 2:
    // (0 + L + R) * (1 / sqrt(2)), AudioNodeInput.cs:214-228).
    bool mono = job.mono != 0;
    const bool odd = n4 >= job.sp_block && n4 < job.sp_block + 128;
    if (odd) mono = job.sp_mode == 2;
    const int64_t first_change = job.first_change ? (int64_t)*job.first_change : INT64_MAX;
    float gl = 0.f, gr = 0.f, last = CUDART_NAN_F;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      float pan = (&pan4.x)[u];
      pan = fminf(fmaxf(pan, -1.0f), 1.0f);
      float a = (&a4.x)[u];
      const float b = (&b4.x)[u];
      if (odd && mono) a = (a + b) * (1.0f / sqrtf(2.0f));
      // The node recomputes its gain pair only when the pan value CHANGES (:95-103, :129-137), with the formula of the variant
      // that is running at that moment — so the pair computed in the odd first quantum stays in force, across the switch of
      // variants, until the first sample whose pan differs from its predecessor's (first_change; never, for a constant pan).
      bool mono_formula = mono;
      if (job.sp_mode != 0 && n4 >= job.sp_block && n4 + u < first_change) mono_formula = job.sp_mode == 2;
      if (pan != last || u == 0 || n4 + u == first_change) {  // (the pair is a pure function of pan and the formula: reuse it within the thread)
        const float x = mono_formula ? (pan + 1.0f) * 0.5f : (pan <= 0.0f ? pan + 1.0f : pan);
        sincosf_libm(x * kPi / 2.0f, &gr, &gl);
        last = pan;
      }
      float l, r;
      if (mono) {  // ProcessMono :77-108
        l = a * gl;
        r = a * gr;
      } else if (pan <= 0.0f) {  // ProcessStereo :110-152
        l = a + b * gl;
        r = b * gr;
      } else {
        l = a * gl;
        r = b + a * gr;
      }
      (&L4.x)[u] = l;
      (&R4.x)[u] = r;
    }
  }
  *reinterpret_cast<float4*>(job.sig[0] + n4) = L4;
  *reinterpret_cast<float4*>(job.sig[1] + n4) = R4;
}
// first frame behind the odd quantum whose (clamped) pan differs from its predecessor's; first_change must be preset to a huge value
__global__ void __launch_bounds__(256) k_pan_first_change(const PannerJob* __restrict__ jobs, int64_t n_frames) {
  const PannerJob job = jobs[blockIdx.y];
  if (!job.pan || !job.first_change || job.sp_mode == 0) return;
  const int64_t n0 = job.sp_block + 128 + ((int64_t)blockIdx.x * 256 + threadIdx.x) * 8;
  for (int64_t n = n0; n < n0 + 8 && n < n_frames && n < job.hi; n++) {
    const float p = fminf(fmaxf(job.pan[n], -1.0f), 1.0f), q = fminf(fmaxf(job.pan[n - 1], -1.0f), 1.0f);
    if (p != q) {
      atomicMin(job.first_change, (unsigned long long)n);
      break;
    }
  }
}
void launch_panner(const PannerJob* d_jobs, int n_jobs, int64_t n_frames, bool scan_changes, cudaStream_t s) {
  if (n_jobs <= 0 || n_frames <= 0) return;
  if (scan_changes)
    for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
      int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
      k_pan_first_change<<<dim3((unsigned)((n_frames + 2047) / 2048), (unsigned)nj), 256, 0, s>>>(d_jobs + j0, n_frames);
    }
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
    k_panner<<<dim3((unsigned)((n_frames / 4 + 255) / 256), (unsigned)nj), 256, 0, s>>>(d_jobs + j0, n_frames);
  }
}

// ============================================================================================ AudioParam modulation
// AudioParam.ComputeARate / ComputeKRate with a connected modulator (AudioParam.cs:114-166): value = Math.Clamp(intrinsic +
// modulation, min, max) in float32 while the modulation block is non-silent, the intrinsic value otherwise.
__device__ __forceinline__ float clamp_param(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }  // Math.Clamp(float)
__global__ void __launch_bounds__(256) k_param_modulate(const ModJob* __restrict__ jobs, int64_t n_frames) {
  const ModJob job = jobs[blockIdx.y];
  if (job.a_rate) {
    const int64_t n4 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (n4 < job.lo || n4 >= job.hi || n4 >= n_frames) return;  // (ranges are multiples of 128 frames)
    float4 t = *reinterpret_cast<float4*>(job.table + n4);
    const float4 m = *reinterpret_cast<const float4*>(job.mod + n4);
    t.x = clamp_param(t.x + m.x, job.minv, job.maxv);
    t.y = clamp_param(t.y + m.y, job.minv, job.maxv);
    t.z = clamp_param(t.z + m.z, job.minv, job.maxv);
    t.w = clamp_param(t.w + m.w, job.minv, job.maxv);
    *reinterpret_cast<float4*>(job.table + n4) = t;
  } else {
    const int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x;  // one value per quantum, modulated by the block's first frame (:152)
    if (q * 128 < job.lo || q * 128 >= job.hi || q * 128 >= n_frames) return;
    job.table[q] = clamp_param(job.table[q] + job.mod[q * 128], job.minv, job.maxv);
  }
}
void launch_param_modulate(const ModJob* d_jobs, int n_jobs, int64_t n_frames, cudaStream_t s) {
  if (n_jobs <= 0 || n_frames <= 0) return;
  dim3 grid((unsigned)((n_frames / 4 + 255) / 256), (unsigned)n_jobs);
  k_param_modulate<<<grid, 256, 0, s>>>(d_jobs, n_frames);
}

// ============================================================================================ scheduled sources
// ConstantSourceNode (Nodes/ConstantSourceNode.cs:75-141): the Offset values on frames [s0, s1), zeros on the rest of the quanta
// the node plays in.  OscillatorNode (Nodes/OscillatorNode.cs:91-196): sample n >= s0 is GenerateSample(phase[n]) with
// phase[s0] = 0, phase[n + 1] = phase[n] + 2*pi*f[n]/fs, minus 2*pi whenever it reaches 2*pi (:150-156).  The reference adds frame by
// frame; here the increments are summed by a three-pass blocked prefix sum in double and reduced modulo 2*pi — the same number up
// to ~1e-10 rad (gac_source_kind in the header says what that means for the samples).
constexpr int kSchedChunk = 1024;
__device__ __forceinline__ double osc_increment(const SchedJob& job, int64_t n, int sample_rate) {
  const float f = job.table ? job.table[n] : job.value;
  return (2.0 * 3.14159265358979323846 * (double)f) / (double)sample_rate;  // (2.0 * Math.PI * freqValues[i]) / Context.SampleRate
}
// pass 1: chunk sums of the phase increments
__global__ void __launch_bounds__(256) k_osc_chunk_sums(const SchedJob* __restrict__ jobs, int sample_rate) {
  const SchedJob job = jobs[blockIdx.y];
  if (job.osc_type < 0) return;
  const int64_t c0 = job.s0 + (int64_t)blockIdx.x * kSchedChunk;
  if (c0 >= job.s1) return;
  double acc = 0.0;
  for (int e = threadIdx.x; e < kSchedChunk; e += 256) {
    const int64_t n = c0 + e;
    if (n < job.s1) acc += osc_increment(job, n, sample_rate);
  }
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) job.chunk_sum[blockIdx.x] = red[0];
}
// pass 2: exclusive scan over the chunk sums (one thread per oscillator: a few hundred chunks)
__global__ void k_osc_scan_chunks(const SchedJob* __restrict__ jobs, int n_jobs) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_jobs) return;
  const SchedJob job = jobs[j];
  if (job.osc_type < 0 || job.s1 <= job.s0) return;
  const int64_t nc = (job.s1 - job.s0 + kSchedChunk - 1) / kSchedChunk;
  double run = 0.0;
  for (int64_t c = 0; c < nc; c++) {
    const double v = job.chunk_sum[c];
    job.chunk_sum[c] = run;
    run += v;
  }
}
__device__ __forceinline__ float osc_sample(double phase, int type) {  // GenerateSample :176-199
  const double two_pi = 2.0 * 3.14159265358979323846;
  switch (type) {
    case 0: return (float)sin(phase);
    case 1: return phase < 3.14159265358979323846 ? 1.0f : -1.0f;
    case 2: return (float)(2.0 * (phase / two_pi) - 1.0);
    default: {
      const double t = phase / two_pi;
      return (float)(4.0 * fabs(t - floor(t + 0.5)) - 1.0);
    }
  }
}
// pass 3: the samples.  CTA = one chunk of one job; constant sources need no phase
__global__ void __launch_bounds__(256) k_sched_fill(const SchedJob* __restrict__ jobs, int sample_rate) {
  const SchedJob job = jobs[blockIdx.y];
  const int64_t c0 = job.lo + (int64_t)blockIdx.x * kSchedChunk;
  if (c0 >= job.hi) return;
  if (job.osc_type < 0) {
    for (int e = threadIdx.x; e < kSchedChunk; e += 256) {
      const int64_t n = c0 + e;
      if (n >= job.hi) break;
      const float v = (n >= job.s0 && n < job.s1) ? (job.table ? job.table[n] : job.value) : 0.f;
      job.dst[0][n] = v;
      job.dst[1][n] = v;
    }
    return;
  }
  // oscillator: this CTA covers frames [c0, c0 + 1024) of [lo, hi); the phase chunks are aligned to s0
  __shared__ double sc[kSchedChunk];
  __shared__ double part[256];
  const double two_pi = 2.0 * 3.14159265358979323846;
  for (int half = 0; half < 2; half++) {
    // frames of this CTA belong to at most two phase chunks (lo and s0 differ by less than a quantum): handle them one after the other
    const int64_t k = (c0 - job.s0 >= 0 ? (c0 - job.s0) / kSchedChunk : -1) + half;  // phase chunk index
    if (k < 0) continue;
    const int64_t p0 = job.s0 + k * kSchedChunk;  // first frame of the phase chunk
    if (p0 >= job.s1 || p0 >= c0 + kSchedChunk) continue;
    // exclusive scan of the increments of phase chunk k (4 per thread, then a scan of the 256 partial sums)
    double v[4], run = 0.0;
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int64_t n = p0 + threadIdx.x * 4 + e;
      v[e] = n < job.s1 ? osc_increment(job, n, sample_rate) : 0.0;
      run += v[e];
    }
    part[threadIdx.x] = run;
    __syncthreads();
    if (threadIdx.x == 0) {
      double acc = 0.0;
      for (int t = 0; t < 256; t++) {
        const double x = part[t];
        part[t] = acc;
        acc += x;
      }
    }
    __syncthreads();
    double ph = job.chunk_sum[k] + part[threadIdx.x];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      sc[threadIdx.x * 4 + e] = ph;
      ph += v[e];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kSchedChunk; e += 256) {
      const int64_t n = p0 + e;
      if (n < c0 || n >= c0 + kSchedChunk || n >= job.hi || n >= job.s1) continue;
      const double phase = fmod(sc[e], two_pi);
      const float s = osc_sample(phase, job.osc_type);
      job.dst[0][n] = s;
      job.dst[1][n] = s;
    }
    __syncthreads();
  }
  // frames of the playing quanta outside [s0, s1) are zero
  for (int e = threadIdx.x; e < kSchedChunk; e += 256) {
    const int64_t n = c0 + e;
    if (n < job.hi && (n < job.s0 || n >= job.s1)) {
      job.dst[0][n] = 0.f;
      job.dst[1][n] = 0.f;
    }
  }
}
void launch_scheduled_sources(const SchedJob* d_jobs, int n_jobs, int64_t n_frames, int sample_rate, bool any_oscillator, cudaStream_t s) {
  if (n_jobs <= 0 || n_frames <= 0) return;
  const unsigned chunks = (unsigned)((n_frames + kSchedChunk - 1) / kSchedChunk) + 1;
  if (any_oscillator) {
    k_osc_chunk_sums<<<dim3(chunks, (unsigned)n_jobs), 256, 0, s>>>(d_jobs, sample_rate);
    k_osc_scan_chunks<<<(unsigned)((n_jobs + 63) / 64), 64, 0, s>>>(d_jobs, n_jobs);
  }
  k_sched_fill<<<dim3(chunks, (unsigned)n_jobs), 256, 0, s>>>(d_jobs, sample_rate);
}

// ============================================================================================ mix
// N -> 1 channel mix of one input block (AudioNodeInput.cs:214-228): sum = 0; sum += L; sum += R; dst += sum * (1 / sqrt(N)).
// The mono result is kept in both rows.
__device__ __forceinline__ void mix_down(float4& l, float4& r, float scale) {
  l.x = (l.x + r.x) * scale; l.y = (l.y + r.y) * scale; l.z = (l.z + r.z) * scale; l.w = (l.w + r.w) * scale;
  r = l;
}
__global__ void __launch_bounds__(256) k_mix(const MixJob* __restrict__ jobs, const MixInput* __restrict__ inputs, int64_t n_frames) {
  const MixJob job = jobs[blockIdx.y];
  int64_t n4 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (n4 >= n_frames) return;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;  // AudioNodeInput.cs:118 buffer cleared
  // the active ranges are multiples of 128 frames, so one test covers the float4.  Four inputs at a time: their eight 16-byte loads
  // are requested before the first sum, the sums then run in connection order (:121-132) — the loop used to be 93 % long-scoreboard
  // with two loads in flight per thread (52 % of the DRAM peak).
  const MixInput* __restrict__ ins = inputs + job.first_input;
  int i = 0;
  for (; i + 4 <= job.n_inputs; i += 4) {
    float4 x0[4], x1[4];
    bool on[4];
    float dm[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const MixInput in = ins[i + k];
      on[k] = n4 >= in.lo && n4 < in.hi;
      dm[k] = in.downmix;
      if (on[k]) {
        x0[k] = *reinterpret_cast<const float4*>(in.src[0] + n4);
        x1[k] = *reinterpret_cast<const float4*>(in.src[1] + n4);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (on[k]) {  // silent inputs are skipped, not added as zeros (:127)
        if (dm[k] != 0.f) mix_down(x0[k], x1[k], dm[k]);
        a0.x += x0[k].x; a0.y += x0[k].y; a0.z += x0[k].z; a0.w += x0[k].w;
        a1.x += x1[k].x; a1.y += x1[k].y; a1.z += x1[k].z; a1.w += x1[k].w;
      }
  }
  for (; i < job.n_inputs; i++) {
    const MixInput in = ins[i];
    if (n4 >= in.lo && n4 < in.hi) {
      float4 x0 = *reinterpret_cast<const float4*>(in.src[0] + n4);
      float4 x1 = *reinterpret_cast<const float4*>(in.src[1] + n4);
      if (in.downmix != 0.f) mix_down(x0, x1, in.downmix);
      a0.x += x0.x; a0.y += x0.y; a0.z += x0.z; a0.w += x0.w;
      a1.x += x1.x; a1.y += x1.y; a1.z += x1.z; a1.w += x1.w;
    }
  }
  *reinterpret_cast<float4*>(job.dst[0] + n4) = a0;
  *reinterpret_cast<float4*>(job.dst[1] + n4) = a1;
}
void launch_mix(const MixJob* d_jobs, int n_jobs, const MixInput* d_inputs, int64_t n_frames, cudaStream_t s) {
  if (n_jobs <= 0 || n_frames <= 0) return;
  dim3 grid((unsigned)((n_frames / 4 + 255) / 256), (unsigned)n_jobs);
  k_mix<<<grid, 256, 0, s>>>(d_jobs, d_inputs, n_frames);
}

// ============================================================================================ K0
// scale = (1 / max(rms, 1.25e-4)) * calibration ; rms via a double sum of float32 squares (PartitionedConvolver.cs:93-102).
// The reference sums sequentially; a tree sum in double differs by ~1e-16 relative, far below float32 resolution.
__global__ void __launch_bounds__(1024) k_ir_scale(const float* __restrict__ base, int64_t stride, int64_t n_frames, int normalize,
                                                   float calibration, float* __restrict__ scale) {
  if (!normalize) {  // Normalize = false: scale stays 1 (PartitionedConvolver.cs:67-68)
    if (threadIdx.x == 0) scale[blockIdx.x] = 1.0f;
    return;
  }
  const float* r = base + (int64_t)blockIdx.x * stride;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n_frames; i += 1024) {
    float sq = r[i] * r[i];  // float * float (:98)
    acc += (double)sq;
  }
  __shared__ double sm[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = sm[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) {
      float power = (float)sqrt(acc / (double)n_frames);
      if (isnan(power) || isinf(power) || power < 0.000125f) power = 0.000125f;
      scale[blockIdx.x] = (1.0f / power) * calibration;
    }
  }
}
void launch_ir_scale(const float* d_base, int64_t stride, int n_channels, int64_t n_frames, int normalize, float calibration, float* d_scale,
                     cudaStream_t s) {
  if (n_channels <= 0) return;
  k_ir_scale<<<n_channels, 1024, 0, s>>>(d_base, stride, n_frames, normalize, calibration, d_scale);
}

// the same for a batch of impulse-response channels described by a job array (deferred preparation: one launch for all the
// impulse responses a voice batch needs); also clears the rows P .. P16 of the channel's spectra
__global__ void __launch_bounds__(1024) k_ir_scale_batch(const IrChanJob* __restrict__ jobs, float calibration, int B) {
  const IrChanJob job = jobs[blockIdx.x];
  float2* tail = job.H + (size_t)job.P * B;
  for (int i = threadIdx.x; i < (job.P16 - job.P) * B; i += 1024) tail[i] = make_float2(0.f, 0.f);
  if (!job.normalize) {
    if (threadIdx.x == 0) *job.scale = 1.0f;
    return;
  }
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < job.n_frames; i += 1024) {
    float sq = job.ir[i] * job.ir[i];  // float * float (:98)
    acc += (double)sq;
  }
  __shared__ double sm[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = sm[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) {
      float power = (float)sqrt(acc / (double)job.n_frames);
      if (isnan(power) || isinf(power) || power < 0.000125f) power = 0.000125f;
      *job.scale = (1.0f / power) * calibration;
    }
  }
}
void launch_ir_scale_batch(const IrChanJob* d_jobs, int n_jobs, float calibration, int B, cudaStream_t s) {
  if (n_jobs <= 0) return;
  k_ir_scale_batch<<<n_jobs, 1024, 0, s>>>(d_jobs, calibration, B);
}

// device <- page-locked host memory by SM loads (16 bytes per thread): for small job tables that must not queue behind the
// buffer uploads in the DMA engine (engine.cu, Scratch::upload)
__global__ void __launch_bounds__(256) k_copy_from_host(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n16) dst[i] = src[i];
}
void launch_copy_from_host(void* d_dst, const void* h_src, size_t bytes, cudaStream_t s) {
  const size_t n16 = bytes / 16;
  if (n16 == 0) return;
  k_copy_from_host<<<(unsigned)((n16 + 255) / 256), 256, 0, s>>>(reinterpret_cast<uint4*>(d_dst), reinterpret_cast<const uint4*>(h_src), n16);
}
// several such tables with ONE launch (the segments travel as a kernel argument): a render uploads a few dozen job tables, and a
// launch per table cost more than the copies
__global__ void __launch_bounds__(256) k_copy_from_host_multi(const __grid_constant__ CopySegments segs) {
  const int g = blockIdx.y;
  const size_t n16 = segs.n16[g];
  const uint4* __restrict__ src = reinterpret_cast<const uint4*>(segs.src[g]);
  uint4* __restrict__ dst = reinterpret_cast<uint4*>(segs.dst[g]);
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n16; i += (size_t)gridDim.x * 256) dst[i] = src[i];
}
void launch_copy_from_host_multi(const CopySegments& segs, cudaStream_t s) {
  if (segs.n <= 0) return;
  size_t mx = 0;
  for (int g = 0; g < segs.n; g++) mx = segs.n16[g] > mx ? segs.n16[g] : mx;
  if (mx == 0) return;
  const unsigned bx = (unsigned)std::min<size_t>((mx + 255) / 256, 64);
  k_copy_from_host_multi<<<dim3(bx, (unsigned)segs.n), 256, 0, s>>>(segs);
}

// interleaved file samples -> planar float32 rows, the conversion libsndfile's sf_readf_float applies (normalised floats:
// 16-bit * 2^-15, 24-bit * 2^-23, 32-bit (float)x * 2^-31, float32 as is) followed by the de-interleave of
// AudioDecoder.DecodePlanar (GraphAudio.IO/LibsndfileDecoder.cs:195-220).  thread = frame; fmt: 0 s16, 1 s24 (3 bytes, LE), 2 s32, 3 f32
__global__ void __launch_bounds__(256) k_deinterleave(const unsigned char* __restrict__ raw, int fmt, int n_channels, int64_t n_frames,
                                                      float* __restrict__ out, int64_t stride) {
  const int64_t f = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (f >= n_frames) return;
  for (int c = 0; c < n_channels; c++) {
    const int64_t i = f * n_channels + c;
    float v;
    if (fmt == 0) v = (float)reinterpret_cast<const short*>(raw)[i] * (1.0f / 32768.0f);
    else if (fmt == 1) {
      const unsigned char* p = raw + 3 * i;
      const int x = (int)((unsigned)p[0] << 8 | (unsigned)p[1] << 16 | (unsigned)p[2] << 24) >> 8;  // sign-extended 24 bits
      v = (float)x * (1.0f / 8388608.0f);
    } else if (fmt == 2) v = (float)reinterpret_cast<const int*>(raw)[i] * (1.0f / 2147483648.0f);
    else v = reinterpret_cast<const float*>(raw)[i];
    out[c * stride + f] = v;
  }
}
void launch_deinterleave(const void* d_raw, int fmt, int n_channels, int64_t n_frames, float* d_out, int64_t stride, cudaStream_t s) {
  if (n_frames <= 0) return;
  k_deinterleave<<<(unsigned)((n_frames + 255) / 256), 256, 0, s>>>(reinterpret_cast<const unsigned char*>(d_raw), fmt, n_channels, n_frames, d_out, stride);
}

// planar destination rows -> interleaved frames with `channels` channels (AudioContextBase.cs:127-160: the destination's
// channels first, the remaining ones zero)
__global__ void __launch_bounds__(256) k_interleave(const float* __restrict__ c0, const float* __restrict__ c1, float* __restrict__ out,
                                                    int64_t n_frames, int channels) {
  const int64_t total = n_frames * channels;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t f = i / channels;
    const int c = (int)(i - f * channels);
    out[i] = c == 0 ? c0[f] : (c == 1 && c1 ? c1[f] : 0.0f);
  }
}
void launch_interleave(const float* c0, const float* c1, float* out, int64_t n_frames, int channels, cudaStream_t s) {
  const int64_t total = n_frames * channels;
  if (total <= 0) return;
  const int64_t want = (total + 255) / 256;
  k_interleave<<<(unsigned)(want < 148 * 16 ? want : 148 * 16), 256, 0, s>>>(c0, c1, out, n_frames, channels);
}

void launch_fill_zero(void* p, size_t bytes, cudaStream_t s) { cudaMemsetAsync(p, 0, bytes, s); }

}  // namespace gac
