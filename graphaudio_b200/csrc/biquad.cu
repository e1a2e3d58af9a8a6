// biquad.cu — K3: BiQuadFilterNode (Nodes/BiQuadFilterNode.cs:87-258), batched over voices.
// Compiled with --fmad=false: the RBJ formulas and the Direct-Form-II recursion round like the reference's scalar float32 code.
//
// The reference recomputes coefficients inside the per-sample loop (sin, cos, five divisions) whenever the a-rate
// frequency/Q moved beyond a hysteresis (:126).  Here that is split so that only the unavoidable recursion is sequential:
//   k_biquad_select   parallel over quanta   the hysteresis walk (usedFreq restarts at 1000 Hz in EVERY block, :13-14,111-112;
//                                            channel 1 inherits channel 0's end state): for every (channel, frame) the frame
//                                            whose (f, Q) produced the coefficients in force, or -1 = "what the block started with"
//   k_biquad_entry    one thread per voice   which coefficient set the reference's fields hold when a quantum starts
//   k_biquad_resolve  parallel over frames   RBJ (glibc-exact sinf/cosf, :149-258) of the frame in force -> (x, a1, a2, b0), (b1, b2)
//   k_biquad_lanes    one lane per (voice, channel): ONLY w = x - a1*w1 - a2*w2 (:137) is sequential; 3-stage cp.async
//                     pipeline of 32-frame slabs so that the lane never waits on HBM
//   k_biquad_output   parallel over frames   y = b0*w + b1*w1 + b2*w2 (:138) from the stored w sequence (same ops, same order)
#include "gac_kernels.h"

namespace gac {

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


// ---------------------------------------------------------------------------------------------------------------------
// K3a: the hysteresis walk (:121-134).  CTA = 128 consecutive quanta of one job; thread = one quantum, which walks
// channel 0 then channel 1.  f / Q / idx travel through shared-memory tiles of [128 quanta][32 frames] so that global
// traffic is coalesced although every thread walks its own 128 frames.
//   idx[c][n]  = frame whose (f, Q) produced the coefficients in force at (channel c, frame n); -1 = the set the quantum
//                started with (resolved by k_biquad_entry)
//   last[c][b] = frame of channel c's last recompute in quantum b, or -1
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kSelQ = 64;  // quanta (= threads) per CTA
__global__ void __launch_bounds__(kSelQ) k_biquad_select(const BiquadJob* __restrict__ jobs, int sample_rate, int64_t n_quanta,
                                                         int64_t n_frames, int32_t* __restrict__ last_base) {
  __shared__ float tf[kSelQ][33];
  __shared__ float tq[kSelQ][33];
  __shared__ int32_t ti[kSelQ][33];
  const int jid = blockIdx.y;
  const BiquadJob job = jobs[jid];
  const int64_t b0 = (int64_t)blockIdx.x * kSelQ;
  const int64_t b = b0 + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t* last = last_base + (size_t)jid * 2 * n_quanta;
  const int64_t n0 = b * 128;
  const bool active = b < n_quanta && n0 >= job.lo && n0 < job.hi;  // silent-flagged block: state untouched (:103-108)
  const float nyq = (float)sample_rate / 2.f;
  float usedF = 1000.f, usedQ = 1.0f;
  bool dirty = (n0 == job.lo);  // _coefficientsDirty is still set when the first non-silent block arrives
  for (int c = 0; c < 2; c++) {
    int32_t cur = -1;
    for (int ci = 0; ci < 4; ci++) {
      // cooperative tile load: row r = quantum b0 + r, 32 consecutive frames
      for (int r = warp; r < kSelQ; r += kSelQ / 32) {
        const int64_t n = (b0 + r) * 128 + ci * 32 + lane;
        float f = job.freq_const, q = job.q_const;
        if (n < n_frames) {
          if (job.freq) f = job.freq[n];
          if (job.q) q = job.q[n];
        }
        tf[r][lane] = f < 1.f ? 1.f : (f > nyq ? nyq : f);  // Math.Clamp(freq, 1, fs/2)   :123
        tq[r][lane] = q > 0.001f ? q : 0.001f;              // Math.Max(0.001f, q)         :124
      }
      __syncthreads();
      if (active) {
#pragma unroll 8
        for (int i = 0; i < 32; i++) {
          const float f = tf[threadIdx.x][i], q = tq[threadIdx.x][i];
          const bool re = dirty || fabsf(f - usedF) > 0.001f || fabsf(q - usedQ) > 0.0001f;  // :126 (gain never differs inside a block)
          if (re) {
            usedF = f;
            usedQ = q;
            dirty = false;
            cur = (int32_t)(n0 + ci * 32 + i);
          }
          ti[threadIdx.x][i] = cur;
        }
      }
      __syncthreads();
      for (int r = warp; r < kSelQ; r += kSelQ / 32) {
        const int64_t qb = b0 + r;
        const int64_t n = qb * 128 + ci * 32 + lane;
        if (qb < n_quanta && n >= job.lo && n < job.hi) job.idx[(size_t)c * n_frames + n] = ti[r][lane];
      }
      __syncthreads();
    }
    if (b < n_quanta) last[(size_t)c * n_quanta + b] = active ? cur : -1;
  }
}

// K3b: which coefficient set the reference's fields (_b0.._a2) hold when channel c of quantum b starts.
// ent[c][b] = frame index whose (f, Q) applies, or -1 (only before the very first recompute).  One thread per job.
__global__ void __launch_bounds__(64) k_biquad_entry(int n_jobs, int64_t n_quanta, const int32_t* __restrict__ last_base,
                                                     int32_t* __restrict__ ent_base) {
  int jid = blockIdx.x * 64 + threadIdx.x;
  if (jid >= n_jobs) return;
  const int32_t* last0 = last_base + (size_t)jid * 2 * n_quanta;
  const int32_t* last1 = last0 + n_quanta;
  int32_t* ent0 = ent_base + (size_t)jid * 2 * n_quanta;
  int32_t* ent1 = ent0 + n_quanta;
  int32_t cur = -1;
  for (int64_t b = 0; b < n_quanta; b++) {
    ent0[b] = cur;                      // channel 0 starts from the fields as the previous block left them
    int32_t l0 = last0[b];
    int32_t c0 = l0 >= 0 ? l0 : cur;
    ent1[b] = c0;                       // channel 1 starts from channel 0's end state (:110-115)
    int32_t l1 = last1[b];
    cur = l1 >= 0 ? l1 : c0;
  }
}

// K3c: coefficients in force at every (channel, frame): RBJ of the frame the walk selected, with the k-rate gain of
// that frame's block.  Writes the two streams the lanes / the output pass read.
__global__ void __launch_bounds__(256) k_biquad_resolve(const BiquadJob* __restrict__ jobs, int sample_rate, int64_t n_quanta,
                                                        int64_t n_frames, const int32_t* __restrict__ ent_base) {
  const int jid = blockIdx.y;
  const BiquadJob job = jobs[jid];
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;  // over 2 * n_frames
  if (t >= 2 * n_frames) return;
  const int c = t >= n_frames ? 1 : 0;
  const int64_t n = t - (int64_t)c * n_frames;
  if (n < job.lo || n >= job.hi) return;
  int32_t i = job.idx[t];
  if (i < 0) i = ent_base[((size_t)jid * 2 + c) * n_quanta + (n >> 7)];
  Coef k;
  if (i >= 0) {
    const float nyq = (float)sample_rate / 2.f;
    k = rbj(job.type, clamped_freq(job, i, nyq), clamped_q(job, i), job.gain ? job.gain[i >> 7] : job.gain_const, sample_rate);
  } else {
    k.b0 = k.b1 = k.b2 = k.a1 = k.a2 = 0.f;  // unreachable for active frames: the first one always recomputes (dirty)
  }
  job.s1[t] = make_float4(job.sig[c][n], k.a1, k.a2, k.b0);
  job.s2[t] = make_float2(k.b1, k.b2);
}

// K3d: the recursion.  One lane per (job, channel); the only sequential arithmetic left is
//     w = x - a1*w1 - a2*w2        (:137, left to right, unfused)
// A warp owns 32 lanes (16 voices x 2 channels) = 32 consecutive rows of the batch-wide (x, a1, a2, b0) stream
// s1_all[row][n_frames].  Slabs of 32 frames travel global -> shared through a 3-stage cp.async pipeline (one warp
// instruction copies one 512 B row; rows that are silent in the slab are zero-filled by the src-size form, so the
// producer code is branch-free); a lane reads its row with conflict-free LDS.128, collects four w's in registers and
// writes them to a small tile that is stored coalesced.
constexpr int kSlab = 32;
constexpr int kBqStages = 3;
constexpr int kSStride = kSlab * 4 + 4;   // floats per stream row (528 B)
constexpr int kWStride = 36;              // floats per w row (144 B)
constexpr int kStageFloats = 32 * kSStride;
constexpr size_t kBqSmem = (size_t)(kBqStages * kStageFloats) * sizeof(float);

// 16-byte async copy; copies src_bytes (0 or 16) and zero-fills the rest
__device__ __forceinline__ void cp_async16_zfill(uint32_t smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(32) k_biquad_lanes(const BiquadJob* __restrict__ jobs, int n_jobs, int64_t n_frames,
                                                     const float4* __restrict__ s1_all, float* __restrict__ w_all) {
  extern __shared__ __align__(16) float bq_smem[];
  __shared__ __align__(16) float wt[32 * kWStride];
  const int lane = threadIdx.x;
  const int j = blockIdx.x * 16 + (lane >> 1);
  const bool valid = j < n_jobs;
  const int64_t my_lo = valid ? jobs[j].lo : 0, my_hi = valid ? jobs[j].hi : 0;
  int64_t lo = my_hi > my_lo ? my_lo : INT64_MAX, hi = my_hi > my_lo ? my_hi : 0;
  for (int o = 16; o > 0; o >>= 1) {
    int64_t lo2 = __shfl_xor_sync(0xffffffffu, lo, o), hi2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = lo2 < lo ? lo2 : lo;
    hi = hi2 > hi ? hi2 : hi;
  }
  if (hi <= lo) return;
  const int n_slabs = (int)((hi - lo) / kSlab);  // ranges are multiples of 128
  const size_t row0 = (size_t)blockIdx.x * 32;   // first row of this warp in the batch-wide streams
  const float4* __restrict__ src0 = s1_all + row0 * (size_t)n_frames + lane;
  float* __restrict__ dst0 = w_all + row0 * (size_t)n_frames;
  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(bq_smem) + lane * 16;

  auto row_mask = [&](int64_t base) -> unsigned { return __ballot_sync(0xffffffffu, base >= my_lo && base < my_hi); };
  auto issue = [&](int s) {
    if (s < n_slabs) {
      const int64_t base = lo + (int64_t)s * kSlab;
      const unsigned mask = row_mask(base);
      const uint32_t st = smem0 + (uint32_t)((s % kBqStages) * kStageFloats * 4);
      const float4* src = src0 + base;
#pragma unroll
      for (int i = 0; i < 32; i++)
        cp_async16_zfill(st + i * (kSStride * 4), src + (size_t)i * n_frames, ((mask >> i) & 1u) ? 16 : 0);
    }
    cp_async_commit();
  };

  for (int s = 0; s < kBqStages - 1; s++) issue(s);
  float w1 = 0.f, w2 = 0.f;
  for (int s = 0; s < n_slabs; s++) {
    issue(s + kBqStages - 1);
    cp_async_wait<kBqStages - 1>();
    __syncwarp();
    const int64_t base = lo + (int64_t)s * kSlab;
    const float* __restrict__ row = bq_smem + (s % kBqStages) * kStageFloats + lane * kSStride;
    float* __restrict__ wrow = wt + lane * kWStride;
    // silent rows read zeros (zero-filled), so their w stays 0 and nothing of them is stored: no branch needed
#pragma unroll
    for (int i4 = 0; i4 < kSlab / 4; i4++) {
      float wo[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const float4 r = *reinterpret_cast<const float4*>(row + (i4 * 4 + u) * 4);  // (x, a1, a2, b0)
        const float w = r.x - r.y * w1 - r.z * w2;
        w2 = w1;
        w1 = w;
        wo[u] = w;
      }
      *reinterpret_cast<float4*>(wrow + i4 * 4) = make_float4(wo[0], wo[1], wo[2], wo[3]);
    }
    __syncwarp();
    const unsigned mask = row_mask(base);
#pragma unroll
    for (int i = 0; i < 8; i++) {  // coalesced store of the w tile: 32 rows x 8 chunks of 16 B
      const int q = lane + 32 * i, r = q >> 3, ch = q & 7;
      if ((mask >> r) & 1u)
        *reinterpret_cast<float4*>(dst0 + (size_t)r * n_frames + base + ch * 4) = *reinterpret_cast<const float4*>(wt + r * kWStride + ch * 4);
    }
    __syncwarp();
  }
}

// K3e: y = b0*w + b1*w1 + b2*w2 (:138) from the stored w sequence; frames outside the non-silent range are cleared (:103-108)
__global__ void __launch_bounds__(256) k_biquad_output(const BiquadJob* __restrict__ jobs, int64_t n_frames) {
  const BiquadJob job = jobs[blockIdx.y];
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= 2 * n_frames) return;
  const int c = t >= n_frames ? 1 : 0;
  const int64_t n = t - (int64_t)c * n_frames;
  float y = 0.f;
  if (n >= job.lo && n < job.hi) {
    const float* w = job.w + (size_t)c * n_frames;
    const float w0 = w[n];
    const float w1 = n - 1 >= job.lo ? w[n - 1] : 0.f;
    const float w2 = n - 2 >= job.lo ? w[n - 2] : 0.f;
    const float b0 = job.s1[t].w;
    const float2 b12 = job.s2[t];
    y = b0 * w0 + b12.x * w1 + b12.y * w2;
  }
  job.sig[c][n] = y;
}

void launch_biquad(const BiquadJob* d_jobs, int n_jobs, int64_t n_frames, int64_t n_quanta, int sample_rate, int32_t* d_last,
                   int32_t* d_ent, const float4* d_s1_all, float* d_w_all, cudaStream_t s) {
  if (n_jobs <= 0 || n_frames <= 0) return;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_biquad_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBqSmem);
    attr = true;
  }
  for (int j0 = 0; j0 < n_jobs; j0 += 32768) {
    int nj = n_jobs - j0 < 32768 ? n_jobs - j0 : 32768;
    int32_t* last = d_last + (size_t)j0 * 2 * n_quanta;
    int32_t* ent = d_ent + (size_t)j0 * 2 * n_quanta;
    k_biquad_select<<<dim3((unsigned)((n_quanta + kSelQ - 1) / kSelQ), (unsigned)nj), kSelQ, 0, s>>>(d_jobs + j0, sample_rate, n_quanta, n_frames, last);
    k_biquad_entry<<<(unsigned)((nj + 63) / 64), 64, 0, s>>>(nj, n_quanta, last, ent);
    k_biquad_resolve<<<dim3((unsigned)((2 * n_frames + 255) / 256), (unsigned)nj), 256, 0, s>>>(d_jobs + j0, sample_rate, n_quanta, n_frames, ent);
    k_biquad_lanes<<<(unsigned)((nj + 15) / 16), 32, kBqSmem, s>>>(d_jobs + j0, nj, n_frames, d_s1_all + (size_t)j0 * 2 * n_frames,
                                                                     d_w_all + (size_t)j0 * 2 * n_frames);
    k_biquad_output<<<dim3((unsigned)((2 * n_frames + 255) / 256), (unsigned)nj), 256, 0, s>>>(d_jobs + j0, n_frames);
  }
}

}  // namespace gac
