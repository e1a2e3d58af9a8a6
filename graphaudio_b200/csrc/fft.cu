// fft.cu — batched real FFTs for the partitioned convolver (K5 forward, K7 inverse + overlap-add).
//
// Replaces RealFourierTransform.Forward/Inverse (FftFlat/RealFourierTransform.cs:62-131, Ooura rdft in
// fftsg.cs:30) and the surrounding casts / zero-padding / overlap-add of PartitionedConvolver.Process
// (PartitionedConvolver.cs:106-124 and :134-150).  The reference computes the transform in double and
// rounds the spectra to float32; here the transform itself is float32 (measured difference to the
// double restatement ~1.2e-7 of full scale) — the only sanctioned precision change.
//
// Shape: a real FFT of N = 2B points whose upper half is zero (forward) / whose lower half is kept and
// upper half carried to the next block (inverse).  Done as an H = B point complex FFT of the even/odd
// packing plus a split step.  One WARP per transform: each lane holds R = H/32 complex points in
// registers; radix-2 DIF stages over the register index first, then five warp-shuffle butterfly stages
// over the lane index; the bit-reversed result is exchanged through a per-warp shared-memory tile, which
// also serves the k <-> H-k pairing of the split step.  HBM-bound by design: 4B bytes in, 8B bytes out.
#include "gac_kernels.h"

namespace gac {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

template <int LOG2H>
struct WarpFft {
  static constexpr int H = 1 << LOG2H;
  static constexpr int R = H / 32;
  static constexpr int LOG2R = LOG2H - 5;
  // number of register-stage twiddles: R/2 + R/4 + ... + 1 = R - 1
  float2 twj[R > 1 ? R - 1 : 1];
  float2 twl[5];

  // tw = e^{-2 pi i k / (2H)} table, k in [0, H).  W_H^t = tw[2t].
  __device__ __forceinline__ void init(const float2* __restrict__ tw, int lane) {
    int idx = 0;
#pragma unroll
    for (int mj = R / 2; mj >= 1; mj >>= 1) {
      // stage half-size m = 32*mj points; exponent (n mod m) * H/(2m), n = jj*32 + lane, jj in [0, mj)
#pragma unroll
      for (int jj = 0; jj < mj; jj++) {
        int e = (jj * 32 + lane) * (H / (64 * mj));
        twj[idx++] = tw[2 * e];
      }
    }
#pragma unroll
    for (int s = 0; s < 5; s++) {
      int m = 16 >> s;
      int e = (lane & (m - 1)) * (H / (2 * m));
      twl[s] = tw[2 * e];
    }
  }

  // forward (INV=false) uses e^{-i..}, inverse uses the conjugate.  v[j] <-> n = j*32 + lane.
  // On return v[j] holds the transform at index bitrev_LOG2H(j*32 + lane).
  template <bool INV>
  __device__ __forceinline__ void run(float2 (&v)[R], int lane) const {
    int idx = 0;
#pragma unroll
    for (int mj = R / 2; mj >= 1; mj >>= 1) {
#pragma unroll
      for (int j = 0; j < R; j++) {
        if ((j & mj) == 0) {
          float2 a = v[j], b = v[j + mj];
          float2 w = twj[idx + (j & (mj - 1))];
          float2 d = make_float2(a.x - b.x, a.y - b.y);
          v[j] = make_float2(a.x + b.x, a.y + b.y);
          v[j + mj] = INV ? cmul_conj(d, w) : cmul(d, w);
        }
      }
      idx += mj;
    }
#pragma unroll
    for (int s = 0; s < 5; s++) {
      const int m = 16 >> s;
      const bool upper = (lane & m) != 0;
      const float2 w = twl[s];
#pragma unroll
      for (int j = 0; j < R; j++) {
        float2 o;
        o.x = __shfl_xor_sync(0xffffffffu, v[j].x, m);
        o.y = __shfl_xor_sync(0xffffffffu, v[j].y, m);
        if (upper) {
          float2 d = make_float2(o.x - v[j].x, o.y - v[j].y);
          v[j] = INV ? cmul_conj(d, w) : cmul(d, w);
        } else {
          v[j] = make_float2(v[j].x + o.x, v[j].y + o.y);
        }
      }
    }
  }

  // index (in the natural-order array) of the value held in v[j]
  __device__ __forceinline__ static int out_index(int j, int lane) {
    unsigned n = (unsigned)(j * 32 + lane);
    return (int)(__brev(n) >> (32 - LOG2H));
  }
};

constexpr int kFftWarps = 4;  // warps (= concurrent transforms) per CTA

// --------------------------------------------------------------------------------------------
// K5 forward: block b of job -> packed spectrum row b.
// grid = (ceil(max_blocks / (kFftWarps*BLOCKS_PER_WARP)), n_jobs)
// --------------------------------------------------------------------------------------------
template <int LOG2H, int BPW>
__device__ __forceinline__ void rfft_fwd_body(const FftFwdJob& job, const float2* __restrict__ tw);

template <int LOG2H, int BPW>
__global__ void __launch_bounds__(kFftWarps * 32) k_rfft_fwd(const FftFwdJob* __restrict__ jobs, const float2* __restrict__ tw) {
  const FftFwdJob job = jobs[blockIdx.y];
  rfft_fwd_body<LOG2H, BPW>(job, tw);
}
// same transform, jobs described arithmetically (no job array in HBM): channel y of one buffer.  Used by IR preparation.
template <int LOG2H, int BPW>
__global__ void __launch_bounds__(kFftWarps * 32) k_rfft_fwd_uniform(FftFwdUniform u, const float2* __restrict__ tw) {
  FftFwdJob job;
  job.in = u.in_base + (int64_t)blockIdx.y * u.in_stride;
  job.out = u.out_base + (int64_t)blockIdx.y * u.out_stride;
  job.scale = u.scale_base ? u.scale_base + blockIdx.y : nullptr;
  job.gain = nullptr;
  job.gain_const = 1.0f;
  job.in2 = nullptr;
  job.mix_scale = 1.0f;
  job.n_valid = u.n_valid;
  job.n_blocks = u.n_blocks;
  job.gate_lo = 0;
  job.gate_hi = INT64_MAX;
  rfft_fwd_body<LOG2H, BPW>(job, tw);
}

template <int LOG2H, int BPW>
__device__ __forceinline__ void rfft_fwd_body(const FftFwdJob& job, const float2* __restrict__ tw) {
  using F = WarpFft<LOG2H>;
  constexpr int H = F::H, R = F::R;
  __shared__ float2 tile[kFftWarps][H + 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  F fft;
  fft.init(tw, lane);
  float2 twk[R];  // split twiddles e^{-2 pi i k / N}, k = lane + 32 j
#pragma unroll
  for (int j = 0; j < R; j++) twk[j] = tw[lane + 32 * j];
  const float sc = job.scale ? *job.scale : 1.0f;

  int64_t b_first = ((int64_t)blockIdx.x * kFftWarps + warp) * BPW;
  for (int bi = 0; bi < BPW; bi++) {
    int64_t b = b_first + bi;
    if (b >= job.n_blocks) break;
    const float* in = job.in + b * H;
    const int64_t base = b * H;
    float2 v[R];
    // z[n] = x[2n] + i x[2n+1] for n < H/2 ; the zero-padded upper half (PartitionedConvolver.cs:107) gives z[n >= H/2] = 0
#pragma unroll
    for (int j = 0; j < R; j++) {
      if (j < R / 2 || R == 1) {
        int n = j * 32 + lane;
        float2 x = make_float2(0.f, 0.f);
        if (R > 1 || n < H / 2) {
          int64_t g = base + 2 * n;
          if (g + 1 < job.n_valid) {
            x = *reinterpret_cast<const float2*>(in + 2 * n);
          } else if (g < job.n_valid) {
            x.x = in[2 * n];
          }
          // silent-flagged quanta read as zero; fused GainNode multiply (Nodes/GainNode.cs:49-58)
          float g0 = job.gain ? job.gain[g] : job.gain_const;
          float g1 = job.gain ? job.gain[g + 1 < job.n_valid ? g + 1 : g] : job.gain_const;
          x.x = (g >= job.gate_lo && g < job.gate_hi) ? __fmul_rn(x.x, g0) : 0.f;
          x.y = (g + 1 >= job.gate_lo && g + 1 < job.gate_hi) ? __fmul_rn(x.y, g1) : 0.f;
          if (job.in2) {
            // stereo -> mono down-mix at the input of a mono-IR convolver: dst = (L + R) * (1/sqrt(2))  (AudioNodeInput.cs:214-228)
            const float* in_b = job.in2 + b * H;
            float2 y = make_float2(0.f, 0.f);
            if (g + 1 < job.n_valid) y = *reinterpret_cast<const float2*>(in_b + 2 * n);
            else if (g < job.n_valid) y.x = in_b[2 * n];
            y.x = (g >= job.gate_lo && g < job.gate_hi) ? __fmul_rn(y.x, g0) : 0.f;
            y.y = (g + 1 >= job.gate_lo && g + 1 < job.gate_hi) ? __fmul_rn(y.y, g1) : 0.f;
            x.x = __fmul_rn(__fadd_rn(x.x, y.x), job.mix_scale);
            x.y = __fmul_rn(__fadd_rn(x.y, y.y), job.mix_scale);
          }
          x.x = x.x * sc;  // float * float, as `sourceIr[offset + i] * scale` (PartitionedConvolver.cs:80)
          x.y = x.y * sc;
        }
        v[j] = x;
      } else {
        v[j] = make_float2(0.f, 0.f);
      }
    }
    fft.template run<false>(v, lane);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < R; j++) tile[warp][F::out_index(j, lane)] = v[j];
    __syncwarp();
    float2* out = job.out + b * H;
#pragma unroll
    for (int j = 0; j < R; j++) {
      int k = lane + 32 * j;
      float2 a = tile[warp][k];
      float2 c = tile[warp][(H - k) & (H - 1)];
      float2 r;
      if (k == 0) {
        r = make_float2(a.x + a.y, a.x - a.y);  // (DC, Nyquist), both purely real
      } else {
        // E = (Z[k] + conj Z[H-k]) / 2 ; O = (Z[k] - conj Z[H-k]) / (2i) ; X[k] = E + e^{-2 pi i k/N} O
        float2 e = make_float2(0.5f * (a.x + c.x), 0.5f * (a.y - c.y));
        float2 o = make_float2(0.5f * (a.y + c.y), -0.5f * (a.x - c.x));
        float2 wo = cmul(twk[j], o);
        r = make_float2(e.x + wo.x, e.y + wo.y);
      }
      out[k] = r;
    }
  }
}

// --------------------------------------------------------------------------------------------
// K7 inverse + overlap-add.  A warp walks BPW consecutive blocks and carries the upper half of each
// inverse transform in registers into the next block (PartitionedConvolver.cs:146-150:
// out[i] = (float)r[i] + overlap[i]; overlap[i] = (float)r[i+B]).  The first block of a walk re-derives
// its predecessor's upper half (one extra transform per BPW).
// --------------------------------------------------------------------------------------------
template <int LOG2H, int BPW>
__global__ void __launch_bounds__(kFftWarps * 32) k_irfft_ola(const FftInvJob* __restrict__ jobs, const float2* __restrict__ tw) {
  using F = WarpFft<LOG2H>;
  constexpr int H = F::H, R = F::R;
  __shared__ float2 tile[kFftWarps][H + 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const FftInvJob job = jobs[blockIdx.y];
  F fft;
  fft.init(tw, lane);
  float2 twk[R];
#pragma unroll
  for (int j = 0; j < R; j++) twk[j] = tw[lane + 32 * j];
  constexpr int HALF = (R >= 2) ? R / 2 : 1;
  float2 ov[HALF], ov2[HALF];
#pragma unroll
  for (int j = 0; j < HALF; j++) ov[j] = ov2[j] = make_float2(0.f, 0.f);

  const int64_t b_first = ((int64_t)blockIdx.x * kFftWarps + warp) * BPW;
  if (b_first >= job.n_blocks) return;
  const float inv_h = 1.0f / (float)H;
  // One inverse transform of block b of spectrogram `spec`: lo[] = float32(r[0:B]) + overlap, overlap <- float32(r[B:2B]).
  auto pass = [&](const float2* __restrict__ spec, int64_t b, float2 (&ovr)[HALF], float2 (&lo)[HALF]) {
    const float2* in = spec + b * H;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < R; j++) tile[warp][lane + 32 * j] = in[lane + 32 * j];
    __syncwarp();
    float2 v[R];
#pragma unroll
    for (int j = 0; j < R; j++) {
      int k = lane + 32 * j;
      float2 a = tile[warp][k];
      float2 c = tile[warp][(H - k) & (H - 1)];
      float2 z;
      if (k == 0) {
        // X[0] = a.x (DC), X[H] = a.y (Nyquist): E = (X0 + XH)/2, O = (X0 - XH)/2, Z[0] = E + iO
        z = make_float2(0.5f * (a.x + a.y), 0.5f * (a.x - a.y));
      } else {
        // conj X[H-k] = (c.x, -c.y)
        float2 e = make_float2(0.5f * (a.x + c.x), 0.5f * (a.y - c.y));
        float2 d = make_float2(0.5f * (a.x - c.x), 0.5f * (a.y + c.y));
        float2 o = cmul_conj(d, twk[j]);  // d * e^{+2 pi i k/N}
        z = make_float2(e.x - o.y, e.y + o.x);
      }
      v[j] = z;
    }
    fft.template run<true>(v, lane);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < R; j++) tile[warp][F::out_index(j, lane)] = make_float2(v[j].x * inv_h, v[j].y * inv_h);
    __syncwarp();
    // tile now holds r[0..2H) as floats (r[2m] = Re z[m], r[2m+1] = Im z[m])
#pragma unroll
    for (int j = 0; j < HALF; j++) {
      const float2 r = tile[warp][lane + 32 * j];  // float2 index in the lower half
      lo[j] = make_float2(r.x + ovr[j].x, r.y + ovr[j].y);
      ovr[j] = tile[warp][H / 2 + lane + 32 * j];
    }
  };
  for (int bi = (b_first > 0 ? -1 : 0); bi < BPW; bi++) {
    const int64_t b = b_first + bi;
    if (b >= job.n_blocks) break;
    float2 lo[HALF];
    pass(job.in, b, ov, lo);
    if (job.in2) {
      // true-stereo: out = c_a(in) + c_b(in) as ConvolverNode.Sum adds the two convolver outputs (ConvolverNode.cs:137-143,158-164)
      float2 lo2[HALF];
      pass(job.in2, b, ov2, lo2);
#pragma unroll
      for (int j = 0; j < HALF; j++) lo[j] = make_float2(lo[j].x + lo2[j].x, lo[j].y + lo2[j].y);
    }
    if (bi >= 0) {
      float* out = job.out + b * H;
#pragma unroll
      for (int j = 0; j < HALF; j++) *reinterpret_cast<float2*>(out + 2 * (lane + 32 * j)) = lo[j];
      if (job.out2) {  // mono result duplicated into the second row (1 -> 2 up-mix copy at the next input)
        float* out2 = job.out2 + b * H;
#pragma unroll
        for (int j = 0; j < HALF; j++) *reinterpret_cast<float2*>(out2 + 2 * (lane + 32 * j)) = lo[j];
      }
    }
  }
}

template <int LOG2H>
static void launch_fwd_t(const FftFwdJob* jobs, int n_jobs, int64_t max_blocks, const float2* tw, cudaStream_t s) {
  constexpr int BPW = 4;
  dim3 grid((unsigned)((max_blocks + kFftWarps * BPW - 1) / (kFftWarps * BPW)), (unsigned)n_jobs);
  k_rfft_fwd<LOG2H, BPW><<<grid, kFftWarps * 32, 0, s>>>(jobs, tw);
}
template <int LOG2H>
static void launch_inv_t(const FftInvJob* jobs, int n_jobs, int64_t max_blocks, const float2* tw, cudaStream_t s) {
  constexpr int BPW = 16;
  dim3 grid((unsigned)((max_blocks + kFftWarps * BPW - 1) / (kFftWarps * BPW)), (unsigned)n_jobs);
  k_irfft_ola<LOG2H, BPW><<<grid, kFftWarps * 32, 0, s>>>(jobs, tw);
}

template <int LOG2H>
static void launch_fwd_uniform_t(const FftFwdUniform& u, int n_jobs, const float2* tw, cudaStream_t s) {
  constexpr int BPW = 4;
  dim3 grid((unsigned)((u.n_blocks + kFftWarps * BPW - 1) / (kFftWarps * BPW)), (unsigned)n_jobs);
  k_rfft_fwd_uniform<LOG2H, BPW><<<grid, kFftWarps * 32, 0, s>>>(u, tw);
}
void launch_rfft_fwd_uniform(const FftFwdUniform& u, int n_jobs, int B, const float2* d_tw, cudaStream_t s) {
  if (n_jobs <= 0 || u.n_blocks <= 0) return;
  switch (B) {
    case 128: launch_fwd_uniform_t<7>(u, n_jobs, d_tw, s); break;
    case 256: launch_fwd_uniform_t<8>(u, n_jobs, d_tw, s); break;
    case 512: launch_fwd_uniform_t<9>(u, n_jobs, d_tw, s); break;
  }
}

void launch_rfft_fwd(const FftFwdJob* d_jobs, int n_jobs, int64_t max_blocks, int B, const float2* d_tw, cudaStream_t s) {
  if (n_jobs <= 0 || max_blocks <= 0) return;
  // gridDim.y limit is 65535: split the job list
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
    switch (B) {
      case 128: launch_fwd_t<7>(d_jobs + j0, nj, max_blocks, d_tw, s); break;
      case 256: launch_fwd_t<8>(d_jobs + j0, nj, max_blocks, d_tw, s); break;
      case 512: launch_fwd_t<9>(d_jobs + j0, nj, max_blocks, d_tw, s); break;
    }
  }
}
void launch_irfft_ola(const FftInvJob* d_jobs, int n_jobs, int64_t max_blocks, int B, const float2* d_tw, cudaStream_t s) {
  if (n_jobs <= 0 || max_blocks <= 0) return;
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
    switch (B) {
      case 128: launch_inv_t<7>(d_jobs + j0, nj, max_blocks, d_tw, s); break;
      case 256: launch_inv_t<8>(d_jobs + j0, nj, max_blocks, d_tw, s); break;
      case 512: launch_inv_t<9>(d_jobs + j0, nj, max_blocks, d_tw, s); break;
    }
  }
}

}  // namespace gac
