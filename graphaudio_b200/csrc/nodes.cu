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

#include "gac_kernels.h"

namespace gac {

// ============================================================================================ K2
__device__ __forceinline__ float interp_linear(float v0, double t0, float v1, double t1, double t) {  // AudioParam.cs:220-225
  double u = (t - t0) / (t1 - t0);
  u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
  float d = v1 - v0;
  return (float)((double)v0 + (double)d * u);
}
__device__ __forceinline__ float interp_exponential(float v0, double t0, float v1, double t1, double t) {  // :228-237
  if (v0 <= 0.f || v1 <= 0.f) return interp_linear(v0, t0, v1, t1, t);
  double u = (t - t0) / (t1 - t0);
  u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
  float ratio = v1 / v0;
  return (float)((double)v0 * pow((double)ratio, u));
}
__device__ __forceinline__ float set_target(const DevEvent& e, float base, double t) {  // :240-247
  double elapsed = t - e.time;
  if (elapsed <= 0.0) return base;
  double tc = e.time_constant > 0.001 ? e.time_constant : 0.001;
  float d = base - e.target;
  return (float)((double)e.target + (double)d * exp(-elapsed / tc));
}
// ComputeValueAtTime AudioParam.cs:169-217
__device__ float value_at_time(float value, const DevEvent* __restrict__ ev, int count, double time) {
  if (count == 0) return value;
  float boundary = value;
  for (int i = 0; i < count; i++) {
    const DevEvent e = ev[i];
    if (time < e.time) {
      if (i == 0) return boundary;
      const DevEvent prev = ev[i - 1];
      if (e.type == 1) return interp_linear(prev.value, prev.time, e.value, e.time, time);
      if (e.type == 2) return interp_exponential(prev.value, prev.time, e.value, e.time, time);
      if (prev.type == 3) return set_target(prev, boundary, time);
      return prev.value;
    }
    if (e.type != 3) boundary = e.value;
  }
  const DevEvent last = ev[count - 1];
  if (last.type == 3) return set_target(last, boundary, time);
  return last.value;
}

__global__ void __launch_bounds__(256) k_param_eval(const ParamJob* __restrict__ jobs, const double* __restrict__ block_time,
                                                    int64_t n_quanta, int sample_rate) {
  const ParamJob job = jobs[blockIdx.y];
  const double dt = 1.0 / (double)sample_rate;  // AudioParam.cs:116
  if (job.a_rate) {
    int64_t n = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (n >= n_quanta * 128) return;
    int64_t b = n >> 7;
    int i = (int)(n & 127);
    double t = block_time[b] + (double)i * dt;  // :120
    job.out[n] = value_at_time(job.value, job.events, job.n_events, t);
  } else {
    int64_t b = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (b >= n_quanta) return;
    job.out[b] = value_at_time(job.value, job.events, job.n_events, block_time[b]);  // ComputeKRate :144-146
  }
}

void launch_param_eval(const ParamJob* d_jobs, int n_jobs, const double* d_block_time, int64_t n_quanta, int sample_rate, cudaStream_t s) {
  if (n_jobs <= 0 || n_quanta <= 0) return;
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
    dim3 grid((unsigned)((n_quanta * 128 + 255) / 256), (unsigned)nj);
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
      pv[u] = (m >= 0 && m < job.n_emit) ? job.src[c][job.pos0 + m] : 0.f;
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
#pragma unroll
  for (int c = 0; c < 2; c++) {
    float y = 0.f;
    if (live) {
      const float* p = job.src[c] + k;
      float S0 = p[0], S1 = p[1], S2 = p[2], S3 = p[3];
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

// ============================================================================================ mix
__global__ void __launch_bounds__(256) k_mix(const MixJob* __restrict__ jobs, const MixInput* __restrict__ inputs, int64_t n_frames) {
  const MixJob job = jobs[blockIdx.y];
  int64_t n4 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (n4 >= n_frames) return;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;  // AudioNodeInput.cs:118 buffer cleared
  // the active ranges are multiples of 128 frames, so one test covers the float4
  for (int i = 0; i < job.n_inputs; i++) {  // connection order (:121-132)
    const MixInput in = inputs[job.first_input + i];
    if (n4 >= in.lo && n4 < in.hi) {
      float4 x0 = *reinterpret_cast<const float4*>(in.src[0] + n4);
      float4 x1 = *reinterpret_cast<const float4*>(in.src[1] + n4);
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
__global__ void __launch_bounds__(1024) k_ir_scale(const float* const* __restrict__ channels, int64_t n_frames, float calibration,
                                                   float* __restrict__ scale) {
  const float* r = channels[blockIdx.x];
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
void launch_ir_scale(const float* const* d_channels, int n_channels, int64_t n_frames, float calibration, float* d_scale, cudaStream_t s) {
  if (n_channels <= 0) return;
  k_ir_scale<<<n_channels, 1024, 0, s>>>(d_channels, n_frames, calibration, d_scale);
}

void launch_fill_zero(void* p, size_t bytes, cudaStream_t s) { cudaMemsetAsync(p, 0, bytes, s); }

// ============================================================================================ K3
// glibc's sinf/cosf (sysdeps/ieee754/flt-32/s_sincosf.h — the ARM optimized-routines algorithm that .NET's
// MathF.Sin/Cos reach through the platform libm on Linux): double-precision pi/2 reduction + double polynomial,
// rounded once to float.  Restated here so that the RBJ coefficients (BiQuadFilterNode.cs:151-153) come out
// bit-identical to the CPU oracle; verified exhaustively on the host for every float in [1e-5, 3.2].
// Valid for 0 <= x < 120 (w0 = 2*pi*f/fs lies in (0, pi]).
__device__ __forceinline__ void sincosf_libm(float y, float* sn, float* cs) {
  const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
  const double c0 = 1.0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10, c4 = 0x1.99343027bf8c3p-16;
  const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
  double x = (double)y;
  int n = 0;
  double sgn = 1.0;
  bool neg = false;
  const unsigned top = (__float_as_uint(y) >> 20) & 0x7ff;
  if (top >= 0x3f4) {  // |y| >= pi/4  (abstop12(pio4) = 0x3f4)
    double r = x * hpi_inv;
    n = ((int)r + 0x800000) >> 24;
    x = x - (double)n * hpi;
    sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    neg = (n & 2) != 0;
  } else if (top < 0x398) {  // |y| < 2^-12
    *sn = y;
    *cs = 1.0f;
    return;
  }
  const double x2 = x * x;
  const double xs = x * sgn;
  // sine polynomial on (xs, x2), cosine polynomial on (x2) with the sign of table 1 when n & 2
  double sinp, cosp;
  {
    double x3 = xs * x2;
    double S1 = s2 + x2 * s3;
    double x7 = x3 * x2;
    double sv = xs + x3 * s1;
    sinp = sv + x7 * S1;
  }
  {
    double C0 = neg ? -c0 : c0, C1 = neg ? -c1 : c1, C2 = neg ? -c2 : c2, C3 = neg ? -c3 : c3, C4 = neg ? -c4 : c4;
    double x4 = x2 * x2;
    double cc2 = C3 + x2 * C4;
    double cc1 = C0 + x2 * C1;
    double x6 = x4 * x2;
    double cv = cc1 + x4 * C2;
    cosp = cv + x6 * cc2;
  }
  // sinf uses poly(n), cosf uses poly(n ^ 1): even -> sine polynomial, odd -> cosine polynomial
  if ((n & 1) == 0) {
    *sn = (float)sinp;
    *cs = (float)cosp;
  } else {
    *sn = (float)cosp;
    // cosf with odd n evaluates the sine polynomial; the table-1 switch only negates cosine coefficients,
    // the sign of the sine polynomial is carried by xs
    *cs = (float)sinp;
  }
}

struct Coef {
  float b0, b1, b2, a1, a2;
};

// UpdateCoefficients, BiQuadFilterNode.cs:149-258
__device__ Coef rbj(int type, float frequency, float q, float gain, int sample_rate) {
  float w0 = 2.f * 3.14159274f * frequency / (float)sample_rate;  // left to right in float32 (:151)
  float sinW0, cosW0;
  sincosf_libm(w0, &sinW0, &cosW0);
  float alpha = sinW0 / (2.f * q);
  float a0, a1, a2, b0, b1, b2;
  switch (type) {
    case 0: b0 = (1.f - cosW0) / 2.f; b1 = 1.f - cosW0; b2 = (1.f - cosW0) / 2.f; a0 = 1.f + alpha; a1 = -2.f * cosW0; a2 = 1.f - alpha; break;
    case 1: b0 = (1.f + cosW0) / 2.f; b1 = -(1.f + cosW0); b2 = (1.f + cosW0) / 2.f; a0 = 1.f + alpha; a1 = -2.f * cosW0; a2 = 1.f - alpha; break;
    case 2: b0 = alpha; b1 = 0.f; b2 = -alpha; a0 = 1.f + alpha; a1 = -2.f * cosW0; a2 = 1.f - alpha; break;
    case 3: b0 = 1.f; b1 = -2.f * cosW0; b2 = 1.f; a0 = 1.f + alpha; a1 = -2.f * cosW0; a2 = 1.f - alpha; break;
    case 4: b0 = 1.f - alpha; b1 = -2.f * cosW0; b2 = 1.f + alpha; a0 = 1.f + alpha; a1 = -2.f * cosW0; a2 = 1.f - alpha; break;
    case 5: {
      float A = (float)pow(10.0, (double)(gain / 40.f));  // MathF.Pow -> powf; double pow rounded to float agrees except on rare ties
      b0 = 1.f + alpha * A; b1 = -2.f * cosW0; b2 = 1.f - alpha * A; a0 = 1.f + alpha / A; a1 = -2.f * cosW0; a2 = 1.f - alpha / A; break;
    }
    case 6: {
      float A = (float)pow(10.0, (double)(gain / 40.f));
      float sqrtA = sqrtf(A);
      float beta = sqrtA / q;
      b0 = A * ((A + 1.f) - (A - 1.f) * cosW0 + beta * sinW0);
      b1 = 2.f * A * ((A - 1.f) - (A + 1.f) * cosW0);
      b2 = A * ((A + 1.f) - (A - 1.f) * cosW0 - beta * sinW0);
      a0 = (A + 1.f) + (A - 1.f) * cosW0 + beta * sinW0;
      a1 = -2.f * ((A - 1.f) + (A + 1.f) * cosW0);
      a2 = (A + 1.f) + (A - 1.f) * cosW0 - beta * sinW0;
      break;
    }
    case 7: {
      float A = (float)pow(10.0, (double)(gain / 40.f));
      float sqrtA = sqrtf(A);
      float beta = sqrtA / q;
      b0 = A * ((A + 1.f) + (A - 1.f) * cosW0 + beta * sinW0);
      b1 = -2.f * A * ((A - 1.f) + (A + 1.f) * cosW0);
      b2 = A * ((A + 1.f) + (A - 1.f) * cosW0 - beta * sinW0);
      a0 = (A + 1.f) - (A - 1.f) * cosW0 + beta * sinW0;
      a1 = 2.f * ((A - 1.f) - (A + 1.f) * cosW0);
      a2 = (A + 1.f) - (A - 1.f) * cosW0 - beta * sinW0;
      break;
    }
    default: b0 = 1.f; b1 = 0.f; b2 = 0.f; a0 = 1.f; a1 = 0.f; a2 = 0.f; break;
  }
  Coef c;
  c.b0 = b0 / a0; c.b1 = b1 / a0; c.b2 = b2 / a0; c.a1 = a1 / a0; c.a2 = a2 / a0;  // :253-257
  return c;
}

__device__ __forceinline__ float clamped_freq(const BiquadJob& job, int64_t n, float nyq) {
  float f = job.freq ? job.freq[n] : job.freq_const;
  return f < 1.f ? 1.f : (f > nyq ? nyq : f);  // Math.Clamp(freq, 1, fs/2)  :123
}
__device__ __forceinline__ float clamped_q(const BiquadJob& job, int64_t n) {
  float q = job.q ? job.q[n] : job.q_const;
  return q > 0.001f ? q : 0.001f;  // Math.Max(0.001f, q)  :124
}

// K3a: recompute decisions.  One thread per (job, quantum): walks channel 0 then channel 1 through the
// hysteresis test of :126.  usedFreq/usedQ restart from (1000, 1) at the top of EVERY block because
// _lastFrequency/_lastQ are never written (:13-14,111-112); channel 1 inherits channel 0's end state.
// sel[c][n] = 1 when the coefficients are recomputed at frame n for channel c.
// last[c][b] = frame of the last recompute of channel c in quantum b, or -1.
__global__ void __launch_bounds__(128) k_biquad_select(const BiquadJob* __restrict__ jobs, int sample_rate, int64_t n_quanta,
                                                       uint8_t* __restrict__ sel_base, int32_t* __restrict__ last_base, int64_t n_frames) {
  const int jid = blockIdx.y;
  const BiquadJob job = jobs[jid];
  int64_t b = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (b >= n_quanta) return;
  uint8_t* sel = sel_base + (size_t)jid * 2 * n_frames;
  int32_t* last = last_base + (size_t)jid * 2 * n_quanta;
  const int64_t n0 = b * 128;
  if (n0 < job.lo || n0 >= job.hi) {
    last[b] = -1;
    last[n_quanta + b] = -1;
    return;  // silent-flagged input block: state and coefficients untouched (:103-108)
  }
  const float nyq = (float)sample_rate / 2.f;
  float usedF = 1000.f, usedQ = 1.0f;
  bool dirty = (n0 == job.lo);  // _coefficientsDirty is still set when the first non-silent block arrives
  for (int c = 0; c < 2; c++) {
    int32_t lastc = -1;
    for (int i = 0; i < 128; i++) {
      const int64_t n = n0 + i;
      float f = clamped_freq(job, n, nyq);
      float q = clamped_q(job, n);
      bool re = dirty || fabsf(f - usedF) > 0.001f || fabsf(q - usedQ) > 0.0001f;
      if (re) {
        usedF = f;
        usedQ = q;
        dirty = false;
        lastc = (int32_t)n;
      }
      sel[(size_t)c * n_frames + n] = re ? 1 : 0;
    }
    last[(size_t)c * n_quanta + b] = lastc;
  }
}

// K3b: coefficient table for every frame of the active range (parallel; sin/cos/5 divisions hoisted out of the lanes)
__global__ void __launch_bounds__(256) k_biquad_coef(const BiquadJob* __restrict__ jobs, int sample_rate, int64_t n_frames) {
  const BiquadJob job = jobs[blockIdx.y];
  int64_t n = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (n >= n_frames || n < job.lo || n >= job.hi) return;
  const float nyq = (float)sample_rate / 2.f;
  float f = clamped_freq(job, n, nyq);
  float q = clamped_q(job, n);
  float g = job.gain ? job.gain[n >> 7] : job.gain_const;
  Coef c = rbj(job.type, f, q, g, sample_rate);
  float* o = job.coef + (size_t)n * 5;
  o[0] = c.b0; o[1] = c.b1; o[2] = c.b2; o[3] = c.a1; o[4] = c.a2;
}

// K3c: the recursion.  One lane per (job, channel): float32 Direct-Form-II exactly as :136-141,
//   w = x - a1*w1 - a2*w2 ; y = b0*w + b1*w1 + b2*w2   (left to right, unfused).
// A warp owns 32 lanes (16 voices x 2 channels) and stages 32-frame slabs of x / sel through shared memory
// so that global traffic is coalesced although each lane walks its own stream.
__global__ void __launch_bounds__(32) k_biquad_lanes(const BiquadJob* __restrict__ jobs, int n_jobs, int64_t n_quanta,
                                                     const uint8_t* __restrict__ sel_base, const int32_t* __restrict__ last_base,
                                                     int64_t n_frames) {
  __shared__ float xs[32][33];
  __shared__ uint8_t ss[32][36];
  const int lane = threadIdx.x;
  const int jid = blockIdx.x * 16 + (lane >> 1);
  const int c = lane & 1;
  const bool valid = jid < n_jobs;
  BiquadJob job;
  if (valid) job = jobs[jid];
  // warp-wide frame range
  int64_t lo = valid ? job.lo : INT64_MAX, hi = valid ? job.hi : 0;
  for (int o = 16; o > 0; o >>= 1) {
    int64_t lo2 = __shfl_xor_sync(0xffffffffu, lo, o), hi2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = lo2 < lo ? lo2 : lo;
    hi = hi2 > hi ? hi2 : hi;
  }
  const int64_t my_lo = valid ? job.lo : 0, my_hi = valid ? job.hi : 0;
  const int32_t* last0 = last_base + (size_t)jid * 2 * n_quanta;
  const int32_t* last1 = last0 + n_quanta;
  const float* coef = valid ? job.coef : nullptr;
  float b0 = 0.f, b1 = 0.f, b2 = 0.f, a1 = 0.f, a2 = 0.f, w1 = 0.f, w2 = 0.f;

  // rows of the staging tile: row r belongs to lane r
  const int row_j = blockIdx.x * 16;
  for (int64_t base = lo; base < hi; base += 32) {
    // cooperative load: for each row r, 32 consecutive frames
#pragma unroll 4
    for (int r = 0; r < 32; r++) {
      int rj = row_j + (r >> 1);
      float v = 0.f;
      uint8_t sv = 0;
      if (rj < n_jobs) {
        const BiquadJob& jr = jobs[rj];
        int64_t n = base + lane;
        if (n >= jr.lo && n < jr.hi) {
          v = jr.sig[r & 1][n];
          sv = sel_base[((size_t)rj * 2 + (r & 1)) * n_frames + n];
        }
      }
      xs[r][lane] = v;
      ss[r][lane] = sv;
    }
    __syncwarp();
    const bool act = valid && base >= my_lo && base < my_hi;  // ranges are multiples of 128, slabs of 32 never straddle
    if (act) {
      if ((base & 127) == 0) {
        // block entry: pick up the coefficients the reference would hold in its fields here
        const int64_t b = base >> 7;
        if (c == 0) {
          if (base > my_lo) {
            int32_t l1 = last1[b - 1];
            if (l1 >= 0) { const float* p = coef + (size_t)l1 * 5; b0 = p[0]; b1 = p[1]; b2 = p[2]; a1 = p[3]; a2 = p[4]; }
          }
        } else {
          int32_t l0 = last0[b];
          if (l0 >= 0) { const float* p = coef + (size_t)l0 * 5; b0 = p[0]; b1 = p[1]; b2 = p[2]; a1 = p[3]; a2 = p[4]; }
          else if (base > my_lo) {
            // channel 0 kept the fields it found: those are channel 1's own values from the previous block
          }
        }
      }
#pragma unroll 4
      for (int i = 0; i < 32; i++) {
        if (ss[lane][i]) {
          const float* p = coef + (size_t)(base + i) * 5;
          b0 = p[0]; b1 = p[1]; b2 = p[2]; a1 = p[3]; a2 = p[4];
        }
        float x = xs[lane][i];
        float w = x - a1 * w1 - a2 * w2;
        float y = b0 * w + b1 * w1 + b2 * w2;
        w2 = w1;
        w1 = w;
        xs[lane][i] = y;
      }
    } else {
#pragma unroll 4
      for (int i = 0; i < 32; i++) xs[lane][i] = 0.f;  // silent-flagged block -> cleared output
    }
    __syncwarp();
#pragma unroll 4
    for (int r = 0; r < 32; r++) {
      int rj = row_j + (r >> 1);
      if (rj < n_jobs) {
        const BiquadJob& jr = jobs[rj];
        int64_t n = base + lane;
        if (n >= jr.lo && n < jr.hi) jr.sig[r & 1][n] = xs[r][lane];
      }
    }
    __syncwarp();
  }
}

// zero the frames outside the active range (biquad output is cleared on silent-flagged input)
__global__ void __launch_bounds__(256) k_gate(const BiquadJob* __restrict__ jobs, int64_t n_frames) {
  const BiquadJob job = jobs[blockIdx.y];
  int64_t n4 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (n4 >= n_frames) return;
  if (n4 >= job.lo && n4 < job.hi) return;
  *reinterpret_cast<float4*>(job.sig[0] + n4) = make_float4(0.f, 0.f, 0.f, 0.f);
  *reinterpret_cast<float4*>(job.sig[1] + n4) = make_float4(0.f, 0.f, 0.f, 0.f);
}

void launch_biquad(const BiquadJob* d_jobs, int n_jobs, int64_t n_frames, int64_t n_quanta, int sample_rate, uint8_t* d_sel,
                   int32_t* d_last, cudaStream_t s) {
  if (n_jobs <= 0 || n_frames <= 0) return;
  for (int j0 = 0; j0 < n_jobs; j0 += 32768) {
    int nj = n_jobs - j0 < 32768 ? n_jobs - j0 : 32768;
    uint8_t* sel = d_sel + (size_t)j0 * 2 * n_frames;
    int32_t* last = d_last + (size_t)j0 * 2 * n_quanta;
    k_biquad_select<<<dim3((unsigned)((n_quanta + 127) / 128), (unsigned)nj), 128, 0, s>>>(d_jobs + j0, sample_rate, n_quanta, sel, last, n_frames);
    k_biquad_coef<<<dim3((unsigned)((n_frames + 255) / 256), (unsigned)nj), 256, 0, s>>>(d_jobs + j0, sample_rate, n_frames);
    k_biquad_lanes<<<(unsigned)((nj + 15) / 16), 32, 0, s>>>(d_jobs + j0, nj, n_quanta, sel, last, n_frames);
    k_gate<<<dim3((unsigned)((n_frames / 4 + 255) / 256), (unsigned)nj), 256, 0, s>>>(d_jobs + j0, n_frames);
  }
}

}  // namespace gac
