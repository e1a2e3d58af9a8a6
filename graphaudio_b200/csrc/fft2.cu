// fft2.cu — K6 as a fast convolution along BLOCK TIME ("second-level FFT").
//
// Replaces PartitionedConvolver.ProcessSpectralConvolution (PartitionedConvolver.cs:154-223) together with the
// frequency-domain delay line it walks (:115-128).  For one channel-convolver and one frequency bin k the reference
// computes, quantum after quantum,
//     Y[b][k] = sum_{p=0}^{P-1} X[b-p][k] * H[p][k]                       (complex, X[<0] = 0)
// i.e. a causal linear convolution of the bin's input-spectrum sequence X[.][k] with the bin's partition sequence
// H[.][k] along the block index b.  Offline, the whole sequence X[0..Q) is known, so that convolution is done with a
// second FFT of length M along b (overlap-save: every segment of M input spectra yields V = M - Lh valid outputs,
// Lh >= P-1):  Q*P complex MACs per bin become (Q/V) * (2 * 5 M log2 M + 6 M) flops.  For the BASELINE shapes
// (P = 750, Q = 4500, M = 2048) that is 27x fewer flops than the direct sum, and the kernel is no longer bound by
// the FP32 pipe but by moving X, H2 and Y through HBM once.
//
// Layout: "transposed" spectrograms  XT[chan][k][b]  (k = 0..B, C = B+1 rows as in the reference, :41; row 0 = DC and
// row B = Nyquist, both real sequences stored with zero imaginary part; b contiguous), written by the transposing
// variant of K5 and read by the transposing variant of K7 (fft.cu).  The prepared second-level IR spectra are
// H2[irchan][k][M] (digit-reversed along M, scaled by 1/M), built once per impulse response by k_fft2_prep.
//
// One CTA = one (channel-convolver, bin, segment): in-place FFT through shared memory (fft2_core.cuh), pointwise
// product, mirrored inverse, store of the V valid outputs.  M <= 4096: radix-16 plan, M/16 threads x 16 points, four
// shared-memory crossings (k_fft2_conv16); M = 8192: radix-8 plan, M/8 threads x 8 points (k_fft2_conv).  The two plans
// order the spectrum differently, so H2 is always prepared by the plan that consumes it.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "fft2_core.cuh"
#include "gac_kernels.h"

namespace gac {

using namespace f2;

int fft2_pick_m(int P, int* Lh_out) {
  int Lh = ((std::max(P - 1, 0) + 15) / 16) * 16;
  if (Lh_out) *Lh_out = Lh;
  for (int M = 512; M <= 8192; M *= 2)
    if (M - Lh >= M / 2) return M;
  // Longer impulse responses still win with short segments: V = 256 valid outputs per 8192-point segment is ~25 kFLOP per output
  // block and bin against 8 P = 63 kFLOP for the direct sum at P = 7900 (and the direct sum re-reads X P/16 times from L2).
  if (8192 - Lh >= 256) return 8192;
  return 0;  // impulse response too long for one second-level transform: the caller falls back to the direct MAC
}

template <int M, int S>
__device__ __forceinline__ void fwd_rest(float2 (&v)[8], float2* sm, const float2* __restrict__ tw, int t) {
  if constexpr (S > 8) {
    fwd_stage<S, true>(v, sm, tw, t);
    __syncthreads();
    fwd_rest<M, S / 8>(v, sm, tw, t);
  }
}
template <int M, int S>
__device__ __forceinline__ void inv_rest(float2 (&v)[8], float2* sm, const float2* __restrict__ tw, int t) {
  // S = span of the stage to run now; stages below M store and hand over to the next larger span
  if constexpr (S < M) {
    inv_stage<S, true>(v, sm, tw, t);
    __syncthreads();
    inv_rest<M, S * 8>(v, sm, tw, t);
  } else {
    inv_stage<M, false>(v, sm, tw, t);
  }
}

template <int M>
__global__ void __launch_bounds__(M / 8) k_fft2_conv(const Fft2Job* __restrict__ jobs, const float2* __restrict__ tw, int64_t n_blocks,
                                                     int64_t xs, int64_t ys) {
  using P = Plan<M>;
  constexpr int T = P::T;
  extern __shared__ __align__(128) float2 sm[];
  const Fft2Job job = jobs[blockIdx.z];
  const int seg = blockIdx.x;
  if (seg >= job.nseg) return;
  const int k = blockIdx.y;
  const int t = threadIdx.x;
  const int V = M - job.Lh;
  const int64_t b_first = job.b0 + (int64_t)seg * V - job.Lh;  // block index of window element 0
  const float2* __restrict__ xrow = job.X + (int64_t)k * xs;
  float2 v[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int64_t b = b_first + t + T * j;
    v[j] = (b >= 0 && b < n_blocks) ? xrow[b] : make_float2(0.f, 0.f);
  }
  // this bin's IR spectrum at positions 8t .. 8t+7 (issued early: consumed after the forward transform)
  float4 h4[4];
  {
    const float4* __restrict__ hp = reinterpret_cast<const float4*>(job.H2 + (int64_t)k * M + 8 * t);
#pragma unroll
    for (int q = 0; q < 4; q++) h4[q] = hp[q];
  }
  fwd_stage<M, false>(v, sm, tw, t);
  __syncthreads();
  fwd_rest<M, M / 8>(v, sm, tw, t);
  float2 u[8];
  to_points<P::TAIL>(u, sm, t);
#pragma unroll
  for (int q = 0; q < 4; q++) {
    u[2 * q] = cmulf(u[2 * q], make_float2(h4[q].x, h4[q].y));
    u[2 * q + 1] = cmulf(u[2 * q + 1], make_float2(h4[q].z, h4[q].w));
  }
  from_points<P::TAIL>(u, sm, t);
  __syncthreads();
  inv_rest<M, (P::TAIL == 1 ? 64 : P::SL)>(v, sm, tw, t);
  float2* __restrict__ yrow = job.Y + (int64_t)k * ys;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int n = t + T * j;
    const int64_t b = b_first + n;
    if (n >= job.Lh && b < n_blocks) yrow[b] = v[j];
  }
}

// H2[ch][k][.] = forward second-level FFT of the bin-k partition sequence H[ch][0..P)[k], zero-padded to M, times 1/M.
// H is the packed first-level layout [ch][P16][B] (bin 0 = (DC, Nyquist)); rows k = 0 and k = B unpack it.
template <int M>
__global__ void __launch_bounds__(M / 8) k_fft2_prep(const float2* __restrict__ H, int64_t h_ch_stride, int B, int P, float2* __restrict__ H2,
                                                     const float2* __restrict__ tw) {
  using Pl = Plan<M>;
  constexpr int T = Pl::T;
  extern __shared__ __align__(128) float2 sm[];
  const int k = blockIdx.x, ch = blockIdx.y, t = threadIdx.x;
  const float2* __restrict__ Hc = H + (int64_t)ch * h_ch_stride;
  const int col = (k == B) ? 0 : k;
  float2 v[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int p = t + T * j;
    float2 h = make_float2(0.f, 0.f);
    if (p < P) {
      h = Hc[(int64_t)p * B + col];
      if (k == 0) h = make_float2(h.x, 0.f);
      else if (k == B) h = make_float2(h.y, 0.f);
    }
    v[j] = h;
  }
  fwd_stage<M, false>(v, sm, tw, t);
  __syncthreads();
  fwd_rest<M, M / 8>(v, sm, tw, t);
  float2 u[8];
  to_points<Pl::TAIL>(u, sm, t);
  const float sc = 1.0f / (float)M;
  float4* __restrict__ out = reinterpret_cast<float4*>(H2 + ((int64_t)ch * (B + 1) + k) * M + 8 * t);
#pragma unroll
  for (int q = 0; q < 4; q++) out[q] = make_float4(u[2 * q].x * sc, u[2 * q].y * sc, u[2 * q + 1].x * sc, u[2 * q + 1].y * sc);
}

// ---------------------------------------------------------------------------------------------------------------------
// Radix-16 plan (M = 512 .. 4096): T = M/16 threads, 16 points per thread, four shared-memory crossings per item.
// tab = this M's twiddle table (r16::table_elems(M) entries, layout in fft2_core.cuh).
// TMA plumbing (cp.async.bulk + mbarrier; SASS: UBLKCP)
__device__ __forceinline__ uint32_t f2_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void f2_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void f2_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void f2_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void f2_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void f2_bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

// Shared memory of one CTA: [ FFT buffer (padded) | X window / Y staging (M float2) | H2 row (M float2, chunk-swizzled) | 2 mbarriers ]
// H2 rows are stored in global memory in the swizzled order of r16::load16_swz, M float2 per row, so that one contiguous bulk
// copy lands them in a layout whose per-thread 128-byte reads are bank-conflict free.
template <int M>
constexpr size_t conv16_smem() { return sizeof(float2) * (size_t)(r16::smem_elems(M) + 2 * M) + 16; }

template <int M>
__global__ void __launch_bounds__(M / 16, 8192 / M) k_fft2_conv16(const Fft2Job* __restrict__ jobs, const float2* __restrict__ tab, int64_t n_blocks, int64_t xs,
                                                        int64_t ys) {
  using P = r16::Plan<M>;
  constexpr int T = P::T;
  extern __shared__ __align__(128) float2 sm[];
  constexpr int SE = P::SE;
  float2* stage = sm + SE;                   // X window in, Y segment out (unpadded: lanes touch consecutive elements)
  float2* hrow = stage + M;                  // this bin's H2 row (swizzled order)
  uint64_t* bars = reinterpret_cast<uint64_t*>(hrow + M);
  const Fft2Job job = jobs[blockIdx.z];
  const int seg = blockIdx.x;
  if (seg >= job.nseg) return;
  const int k = blockIdx.y;
  const int t = threadIdx.x;
  const int V = M - job.Lh;
  const int64_t b_first = job.b0 + (int64_t)seg * V - job.Lh;  // block index of window element 0 (multiple of 16)
  const float2* __restrict__ xrow = job.X + (int64_t)k * xs;
  // window elements [n_lo, n_hi) exist in the spectrogram; the rest of the window is zero (blocks < 0, blocks >= n_blocks)
  const int n_lo = b_first < 0 ? (int)(-b_first) : 0;
  const int64_t avail = n_blocks - b_first;
  const int n_hi = avail < M ? (int)avail : M;
  const uint32_t bar_x = f2_smem_u32(bars), bar_h = bar_x + 8;
  if (t == 0) {
    f2_mbar_init(bar_x, 1);
    f2_mbar_init(bar_h, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // one bulk copy each: the window (rounded up to an even element count: the row stride keeps that inside the row)
    const uint32_t xbytes = (uint32_t)(((n_hi - n_lo + 1) & ~1) * (int)sizeof(float2));
    f2_mbar_expect_tx(bar_x, xbytes);
    f2_bulk_g2s(f2_smem_u32(stage + n_lo), xrow + b_first + n_lo, xbytes, bar_x);
    f2_mbar_expect_tx(bar_h, (uint32_t)(M * sizeof(float2)));
    f2_bulk_g2s(f2_smem_u32(hrow), job.H2 + (int64_t)k * M, (uint32_t)(M * sizeof(float2)), bar_h);
  }
  __syncthreads();  // barrier initialisation visible to every waiter
  f2_mbar_wait(bar_x, 0);
  float2 v[16];
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const int n = t + T * j;
    const float2 x = stage[n];
    v[j] = (n >= n_lo && n < n_hi) ? x : make_float2(0.f, 0.f);
  }
  r16::fwd_a<M>(v, sm, tab, t);
  __syncthreads();
  r16::fwd_b<M>(sm, tab, t);
  __syncthreads();
  {
    float2 u[16];
    r16::load16(u, sm, t);
    r16::stage_c<P::L, false>(u);
    f2_mbar_wait(bar_h, 0);
    float2 h[16];
    r16::load16_swz(h, hrow, t);
#pragma unroll
    for (int q = 0; q < 16; q++) u[q] = cmulf(u[q], h[q]);
    r16::stage_c<P::L, true>(u);
    r16::store16(u, sm, t);
  }
  __syncthreads();
  r16::inv_b<M>(sm, tab, t);
  __syncthreads();
  r16::inv_a<M>(v, sm, tab, t);
  // the V valid outputs leave through the staging buffer with one bulk store (everyone has read the X window long ago)
#pragma unroll
  for (int j = 0; j < 16; j++) stage[t + T * j] = v[j];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (t == 0) {
    const int64_t b0 = b_first + job.Lh;            // first output block of the segment
    int64_t len = n_blocks - b0;
    if (len > V) len = V;
    if (len > 0) {
      const uint32_t ybytes = (uint32_t)(((len + 1) & ~(int64_t)1) * (int64_t)sizeof(float2));
      f2_bulk_s2g(job.Y + (int64_t)k * ys + b0, f2_smem_u32(stage + job.Lh), ybytes);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Fan-in fusion (Fft2SumJob): one CTA = (output channel of a voice chunk, bin, segment).  The CTA walks the chunk's members:
// X window and H2 row of member m + 1 are in flight (bulk copies) while member m is transformed, the products accumulate in
// REGISTERS (the thread's 16 spectrum positions are the same for every member), and only the sum is transformed back.
// Shared memory as k_fft2_conv16 (one window, one row): the window buffer is free again as soon as every thread holds its 16
// points, the row buffer as soon as the products are formed — each is re-filled right then.
// ---------------------------------------------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(M / 16, 8192 / M) k_fft2_sum16(const Fft2SumJob* __restrict__ jobs, const float2* __restrict__ tab, int64_t n_blocks,
                                                                 int64_t xs, int64_t ys) {
  using P = r16::Plan<M>;
  constexpr int T = P::T;
  extern __shared__ __align__(128) float2 sm[];
  constexpr int SE = P::SE;
  float2* stage = sm + SE;
  float2* hrow = stage + M;
  uint64_t* bars = reinterpret_cast<uint64_t*>(hrow + M);
  const Fft2SumJob job = jobs[blockIdx.z];
  const int seg = blockIdx.x;
  if (seg >= job.nseg) return;
  const int k = blockIdx.y;
  const int t = threadIdx.x;
  const int V = M - job.Lh;
  const int64_t b_first = job.b0 + (int64_t)seg * V - job.Lh;
  const int n_lo = b_first < 0 ? (int)(-b_first) : 0;
  const int64_t avail = n_blocks - b_first;
  const int n_hi = avail < M ? (int)avail : M;
  const uint32_t bar_x = f2_smem_u32(bars), bar_h = bar_x + 8;
  const uint32_t xbytes = (uint32_t)(((n_hi - n_lo + 1) & ~1) * (int)sizeof(float2));
  constexpr uint32_t hbytes = (uint32_t)(M * sizeof(float2));
  const int64_t xoff = (int64_t)k * xs + b_first + n_lo, hoff = (int64_t)k * M;
  const uint32_t x_dst = f2_smem_u32(stage + n_lo), h_dst = f2_smem_u32(hrow);
  if (t == 0) {
    f2_mbar_init(bar_x, 1);
    f2_mbar_init(bar_h, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const Fft2SumMember m0 = job.members[0];
    f2_mbar_expect_tx(bar_x, xbytes);
    f2_bulk_g2s(x_dst, m0.X + xoff, xbytes, bar_x);
    f2_mbar_expect_tx(bar_h, hbytes);
    f2_bulk_g2s(h_dst, m0.H2 + hoff, hbytes, bar_h);
  }
  __syncthreads();
  float2 acc[16];
#pragma unroll
  for (int q = 0; q < 16; q++) acc[q] = make_float2(0.f, 0.f);
  for (int m = 0; m < job.n_members; m++) {
    const uint32_t ph = (uint32_t)(m & 1);
    Fft2SumMember nxt{nullptr, nullptr};
    if (t == 0 && m + 1 < job.n_members) nxt = job.members[m + 1];
    f2_mbar_wait(bar_x, ph);
    {
      float2 v[16];
#pragma unroll
      for (int j = 0; j < 16; j++) {
        const int n = t + T * j;
        const float2 x = stage[n];
        v[j] = (n >= n_lo && n < n_hi) ? x : make_float2(0.f, 0.f);
      }
      r16::fwd_a<M>(v, sm, tab, t);
    }
    __syncthreads();  // (everyone holds its window points: the window buffer is free)
    if (nxt.X) {
      f2_mbar_expect_tx(bar_x, xbytes);
      f2_bulk_g2s(x_dst, nxt.X + xoff, xbytes, bar_x);
    }
    r16::fwd_b<M>(sm, tab, t);
    __syncthreads();
    {
      float2 u[16];
      r16::load16(u, sm, t);
      r16::stage_c<P::L, false>(u);
      f2_mbar_wait(bar_h, ph);
      const float4* hp = reinterpret_cast<const float4*>(hrow);
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const float4 h = hp[r16::swz_chunk(t, q)];
        const float2 a = u[2 * q], b = u[2 * q + 1];
        acc[2 * q].x += a.x * h.x - a.y * h.y;
        acc[2 * q].y += a.x * h.y + a.y * h.x;
        acc[2 * q + 1].x += b.x * h.z - b.y * h.w;
        acc[2 * q + 1].y += b.x * h.w + b.y * h.z;
      }
    }
    __syncthreads();  // (the row has been consumed, and the FFT buffer may be overwritten by the next member)
    if (nxt.H2) {
      f2_mbar_expect_tx(bar_h, hbytes);
      f2_bulk_g2s(h_dst, nxt.H2 + hoff, hbytes, bar_h);
    }
  }
  r16::stage_c<P::L, true>(acc);
  r16::store16(acc, sm, t);
  __syncthreads();
  r16::inv_b<M>(sm, tab, t);
  __syncthreads();
  float2 v[16];
  r16::inv_a<M>(v, sm, tab, t);
#pragma unroll
  for (int j = 0; j < 16; j++) stage[t + T * j] = v[j];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (t == 0) {
    const int64_t b0 = b_first + job.Lh;
    int64_t len = n_blocks - b0;
    if (len > V) len = V;
    if (len > 0) {
      const uint32_t ybytes = (uint32_t)(((len + 1) & ~(int64_t)1) * (int64_t)sizeof(float2));
      f2_bulk_s2g(job.Y + (int64_t)k * ys + b0, f2_smem_u32(stage + job.Lh), ybytes);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }
}

template <int M>
__global__ void __launch_bounds__(M / 16) k_fft2_prep16(const float2* __restrict__ H, int64_t h_ch_stride, int B, int P, float2* __restrict__ H2,
                                                        const float2* __restrict__ tab) {
  using Pl = r16::Plan<M>;
  constexpr int T = Pl::T;
  extern __shared__ __align__(128) float2 sm[];
  const int k = blockIdx.x, ch = blockIdx.y, t = threadIdx.x;
  const float2* __restrict__ Hc = H + (int64_t)ch * h_ch_stride;
  const int col = (k == B) ? 0 : k;
  float2 v[16];
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const int p = t + T * j;
    float2 h = make_float2(0.f, 0.f);
    if (p < P) {
      h = Hc[(int64_t)p * B + col];
      if (k == 0) h = make_float2(h.x, 0.f);
      else if (k == B) h = make_float2(h.y, 0.f);
    }
    v[j] = h;
  }
  r16::fwd_a<M>(v, sm, tab, t);
  __syncthreads();
  r16::fwd_b<M>(sm, tab, t);
  __syncthreads();
  float2 u[16];
  r16::load16(u, sm, t);
  r16::stage_c<Pl::L, false>(u);
  r16::store16_swz(u, 1.0f / (float)M, H2 + ((int64_t)ch * (B + 1) + k) * M, t);
}

// the same for a batch of channels described by a job array (deferred IR preparation)
template <int M>
__global__ void __launch_bounds__(M / 16) k_fft2_prep16_batch(const IrChanJob* __restrict__ jobs, int B, const float2* __restrict__ tab) {
  using Pl = r16::Plan<M>;
  constexpr int T = Pl::T;
  extern __shared__ __align__(128) float2 sm[];
  const IrChanJob job = jobs[blockIdx.y];
  const int k = blockIdx.x, t = threadIdx.x;
  const int col = (k == B) ? 0 : k;
  float2 v[16];
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const int p = t + T * j;
    float2 h = make_float2(0.f, 0.f);
    if (p < job.P) {
      h = job.H[(int64_t)p * B + col];
      if (k == 0) h = make_float2(h.x, 0.f);
      else if (k == B) h = make_float2(h.y, 0.f);
    }
    v[j] = h;
  }
  r16::fwd_a<M>(v, sm, tab, t);
  __syncthreads();
  r16::fwd_b<M>(sm, tab, t);
  __syncthreads();
  float2 u[16];
  r16::load16(u, sm, t);
  r16::stage_c<Pl::L, false>(u);
  r16::store16_swz(u, 1.0f / (float)M, job.H2 + (int64_t)k * M, t);
}
template <int M>
static void prep16_batch_t(const IrChanJob* d_jobs, int n_jobs, int B, const float2* d_tab, cudaStream_t s) {
  constexpr size_t smem = sizeof(float2) * r16::smem_elems(M);
  k_fft2_prep16_batch<M><<<dim3((unsigned)(B + 1), (unsigned)n_jobs), M / 16, smem, s>>>(d_jobs, B, d_tab + fft2_table_offset(M));
}
void launch_fft2_prep_batch(const IrChanJob* d_jobs, int n_jobs, int B, int M, const float2* d_tab16, cudaStream_t s) {
  if (n_jobs <= 0) return;
  switch (M) {
    case 512: prep16_batch_t<512>(d_jobs, n_jobs, B, d_tab16, s); break;
    case 1024: prep16_batch_t<1024>(d_jobs, n_jobs, B, d_tab16, s); break;
    case 2048: prep16_batch_t<2048>(d_jobs, n_jobs, B, d_tab16, s); break;
    case 4096: prep16_batch_t<4096>(d_jobs, n_jobs, B, d_tab16, s); break;
  }
}

// Which plan runs a transform length: radix 16 up to 4096 points, radix 8 (M/8 threads x 8 points) above.  GAC_FFT2_R16_MAX=2048
// moves the 4096-point transforms to the radix-8 plan as well (A/B measurements: 512 threads and 32 KB per CTA instead of 256
// threads and 110 KB).  H2 is always prepared by the plan that consumes it.
int fft2_r16_max() {
  static const int v = [] {
    const char* e = getenv("GAC_FFT2_R16_MAX");
    const int x = e ? atoi(e) : 4096;
    return (x == 512 || x == 1024 || x == 2048 || x == 4096) ? x : 4096;
  }();
  return v;
}
int fft2_h2_row_elems(int M) { return M; }  // (both plans: M float2 per row; the radix-16 plan swizzles the row's 16-byte chunks)

int fft2_table_offset(int M) {  // offset of M's radix-16 twiddle table inside the concatenated table buffer
  // order: 512, 1024, 2048, 4096 (second-level transforms), then 128, 256 (first-level transforms of fft_r16.cu)
  int off = 0;
  if (M >= 512) {
    for (int m = 512; m < M; m *= 2) off += r16::table_elems(m);
    return off;
  }
  for (int m = 512; m <= 4096; m *= 2) off += r16::table_elems(m);
  for (int m = 128; m < M; m *= 2) off += r16::table_elems(m);
  return off;
}
int fft2_table_total() { return fft2_table_offset(256) + r16::table_elems(256); }
void fft2_fill_tables(float2* host) {
  const double pi = 3.14159265358979323846;
  const int sizes[6] = {512, 1024, 2048, 4096, 128, 256};
  for (int si = 0; si < 6; si++) {
    const int M = sizes[si];
    float2* tab = host + fft2_table_offset(M);
    const int T = M / 16, L = M >= 512 ? M / 256 : 0;  // L = 0: no stage B (first-level plans)
    for (int q = 0; q < 4; q++) {
      for (int t = 0; t < T; t++) {
        const double a = -2.0 * pi * (double)(t << q) / (double)M;
        tab[q * T + t] = make_float2((float)std::cos(a), (float)std::sin(a));
      }
      for (int i = 0; i < L; i++) {
        const double a = -2.0 * pi * (double)(i << q) / (double)(16 * L);
        tab[4 * T + q * L + i] = make_float2((float)std::cos(a), (float)std::sin(a));
      }
    }
  }
}

template <int M>
static void conv16_t(const Fft2Job* d_jobs, dim3 grid, const float2* d_tab, int64_t n_blocks, int64_t xs, int64_t ys, cudaStream_t s) {
  constexpr size_t smem = conv16_smem<M>();
  GAC_SMEM_OPT_IN(k_fft2_conv16<M>, smem);
  k_fft2_conv16<M><<<grid, M / 16, smem, s>>>(d_jobs, d_tab + fft2_table_offset(M), n_blocks, xs, ys);
}
template <int M>
static void sum16_t(const Fft2SumJob* d_jobs, dim3 grid, const float2* d_tab, int64_t n_blocks, int64_t xs, int64_t ys, cudaStream_t s) {
  constexpr size_t smem = conv16_smem<M>();
  GAC_SMEM_OPT_IN(k_fft2_sum16<M>, smem);
  k_fft2_sum16<M><<<grid, M / 16, smem, s>>>(d_jobs, d_tab + fft2_table_offset(M), n_blocks, xs, ys);
}
template <int M>
static void prep16_t(const float2* d_H, int64_t h_ch_stride, dim3 grid, int B, int P, float2* d_H2, const float2* d_tab, cudaStream_t s) {
  constexpr size_t smem = sizeof(float2) * r16::smem_elems(M);
  k_fft2_prep16<M><<<grid, M / 16, smem, s>>>(d_H, h_ch_stride, B, P, d_H2, d_tab + fft2_table_offset(M));
}

template <int M>
static void conv_t(const Fft2Job* d_jobs, dim3 grid, const float2* d_tw2, int64_t n_blocks, int64_t xs, int64_t ys, cudaStream_t s) {
  constexpr size_t smem = sizeof(float2) * smem_elems(M);
  GAC_SMEM_OPT_IN(k_fft2_conv<M>, smem);
  k_fft2_conv<M><<<grid, M / 8, smem, s>>>(d_jobs, d_tw2, n_blocks, xs, ys);
}
template <int M>
static void prep_t(const float2* d_H, int64_t h_ch_stride, dim3 grid, int B, int P, float2* d_H2, const float2* d_tw2, cudaStream_t s) {
  constexpr size_t smem = sizeof(float2) * smem_elems(M);
  GAC_SMEM_OPT_IN(k_fft2_prep<M>, smem);
  k_fft2_prep<M><<<grid, M / 8, smem, s>>>(d_H, h_ch_stride, B, P, d_H2, d_tw2);
}

void launch_fft2_conv(const Fft2Job* d_jobs, int n_jobs, int max_seg, int C, int M, const float2* d_tw2, const float2* d_tab16, int64_t n_blocks,
                      int64_t xs, int64_t ys, cudaStream_t s) {
  if (n_jobs <= 0 || max_seg <= 0) return;
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    const int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
    dim3 grid((unsigned)max_seg, (unsigned)C, (unsigned)nj);
    if (M == 4096 && fft2_r16_max() < 4096) {
      conv_t<4096>(d_jobs + j0, grid, d_tw2, n_blocks, xs, ys, s);
      continue;
    }
    switch (M) {
      case 512: conv16_t<512>(d_jobs + j0, grid, d_tab16, n_blocks, xs, ys, s); break;
      case 1024: conv16_t<1024>(d_jobs + j0, grid, d_tab16, n_blocks, xs, ys, s); break;
      case 2048: conv16_t<2048>(d_jobs + j0, grid, d_tab16, n_blocks, xs, ys, s); break;
      case 4096: conv16_t<4096>(d_jobs + j0, grid, d_tab16, n_blocks, xs, ys, s); break;
      case 8192: conv_t<8192>(d_jobs + j0, grid, d_tw2, n_blocks, xs, ys, s); break;
    }
  }
}

void launch_fft2_sum(const Fft2SumJob* d_jobs, int n_jobs, int max_seg, int C, int M, const float2* d_tab16, int64_t n_blocks, int64_t xs, int64_t ys,
                     cudaStream_t s) {
  if (n_jobs <= 0 || max_seg <= 0) return;
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    const int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
    dim3 grid((unsigned)max_seg, (unsigned)C, (unsigned)nj);
    switch (M) {
      case 512: sum16_t<512>(d_jobs + j0, grid, d_tab16, n_blocks, xs, ys, s); break;
      case 1024: sum16_t<1024>(d_jobs + j0, grid, d_tab16, n_blocks, xs, ys, s); break;
      case 2048: sum16_t<2048>(d_jobs + j0, grid, d_tab16, n_blocks, xs, ys, s); break;
      case 4096: sum16_t<4096>(d_jobs + j0, grid, d_tab16, n_blocks, xs, ys, s); break;
    }
  }
}
int fft2_sum_ctas_per_sm(int M) {
  size_t smem = 0;
  switch (M) {
    case 512: smem = conv16_smem<512>(); break;
    case 1024: smem = conv16_smem<1024>(); break;
    case 2048: smem = conv16_smem<2048>(); break;
    case 4096: smem = conv16_smem<4096>(); break;
    default: return 1;
  }
  const int by_smem = (int)((size_t)227 * 1024 / (smem + 1024)), by_regs = 8192 / M;  // (launch bounds: 128 registers per thread)
  return by_smem < by_regs ? (by_smem < 1 ? 1 : by_smem) : by_regs;
}

void launch_fft2_prep(const float2* d_H, int64_t h_ch_stride, int n_ch, int B, int P, int M, float2* d_H2, const float2* d_tw2, const float2* d_tab16,
                      cudaStream_t s) {
  if (n_ch <= 0) return;
  dim3 grid((unsigned)(B + 1), (unsigned)n_ch);
  if (M == 4096 && fft2_r16_max() < 4096) {
    prep_t<4096>(d_H, h_ch_stride, grid, B, P, d_H2, d_tw2, s);
    return;
  }
  switch (M) {
    case 512: prep16_t<512>(d_H, h_ch_stride, grid, B, P, d_H2, d_tab16, s); break;
    case 1024: prep16_t<1024>(d_H, h_ch_stride, grid, B, P, d_H2, d_tab16, s); break;
    case 2048: prep16_t<2048>(d_H, h_ch_stride, grid, B, P, d_H2, d_tab16, s); break;
    case 4096: prep16_t<4096>(d_H, h_ch_stride, grid, B, P, d_H2, d_tab16, s); break;
    case 8192: prep_t<8192>(d_H, h_ch_stride, grid, B, P, d_H2, d_tw2, s); break;
  }
}

}  // namespace gac
