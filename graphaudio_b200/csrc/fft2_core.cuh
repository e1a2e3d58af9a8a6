// fft2_core.cuh — per-thread phases of the second-level FFT (the FFT along BLOCK TIME that turns the spectral
// multiply-accumulate of PartitionedConvolver.ProcessSpectralConvolution, PartitionedConvolver.cs:154-223, into a
// per-bin fast convolution; see fft2.cu).
//
// One CTA of T = M/8 threads transforms one length-M complex sequence held in shared memory; every thread owns 8
// points.  In-place decimation-in-frequency, radix 8 with a radix-2/4 tail:
//     M = 512  : 8·8·8        M = 1024 : 8·8·8·2      M = 2048 : 8·8·8·4
//     M = 4096 : 8·8·8·8      M = 8192 : 8·8·8·8·2
// The forward transform leaves the spectrum in digit-reversed order; the pointwise product does not care about the
// order (the IR spectra are produced by the same forward code), and the inverse is the exact mirror of the forward
// (decimation-in-time, conjugate twiddles), so no reordering pass exists.  The inverse is unnormalised (scale M); the
// factor 1/M is folded into the prepared IR spectra.
//
// The code is written as per-thread PHASES with no barrier inside: a kernel calls the phases with __syncthreads()
// between them, and a host harness (tools/fft2_host_test.cu) replays them thread by thread to validate the index
// algebra without a GPU.
#pragma once
#include <cuda_runtime.h>

namespace gac {
namespace f2 {

constexpr int kTwLen = 8192;  // twiddle table tw[e] = exp(-2 pi i e / 8192), e in [0, 8192)

#define F2_HD __host__ __device__ __forceinline__

// shared-memory index of element e: two float2 of padding per 16 elements keep every radix-8 stage at the two
// wavefronts a 64-bit warp access needs anyway, and the contiguous 8-element (4 x 128-bit) accesses conflict-free
F2_HD int pad(int e) { return e + ((e >> 4) << 1); }
constexpr int smem_elems(int M) { return M + (M >> 4) * 2; }

F2_HD float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
F2_HD float2 cmulcf(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a * conj(b)
// complex add / sub: one packed instruction on the device (Blackwell FADD2, PTX add/sub.rn.f32x2) — the butterflies are
// dominated by them (680 of the 1700 instructions of k_fft2_conv16 were scalar FADDs)
#ifdef __CUDA_ARCH__
__device__ __forceinline__ unsigned long long pk2(float2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ float2 upk2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
#endif
F2_HD float2 addf(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
  return upk2(r);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
F2_HD float2 subf(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
  return upk2(r);
#else
  return make_float2(a.x - b.x, a.y - b.y);
#endif
}

F2_HD constexpr int rev3(int r) { return ((r & 1) << 2) | (r & 2) | ((r >> 2) & 1); }

// Radix-8 butterfly.  Forward: in v[j] = x_j, out v[rev3(r)] = sum_j x_j e^{-2 pi i j r / 8}.
// Inverse: in v[rev3(r)] = X_r, out v[j] = sum_r X_r e^{+2 pi i j r / 8}  (= 8 x_j).
template <bool INV>
F2_HD void bf8(float2 (&v)[8]) {
  const float c = 0.70710678118654752440f;
  if (!INV) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float2 a = v[j], b = v[j + 4];
      v[j] = addf(a, b);
      const float2 d = subf(a, b);
      if (j == 0) v[4] = d;
      else if (j == 1) v[5] = make_float2((d.x + d.y) * c, (d.y - d.x) * c);   // d * (1 - i)/sqrt2
      else if (j == 2) v[6] = make_float2(d.y, -d.x);                           // d * (-i)
      else v[7] = make_float2((d.y - d.x) * c, -(d.x + d.y) * c);              // d * (-1 - i)/sqrt2
    }
#pragma unroll
    for (int h = 0; h < 8; h += 4) {
      const float2 a0 = v[h], b0 = v[h + 2], a1 = v[h + 1], b1 = v[h + 3];
      v[h] = addf(a0, b0);
      v[h + 2] = subf(a0, b0);
      v[h + 1] = addf(a1, b1);
      const float2 d = subf(a1, b1);
      v[h + 3] = make_float2(d.y, -d.x);
    }
#pragma unroll
    for (int h = 0; h < 8; h += 2) {
      const float2 a = v[h], b = v[h + 1];
      v[h] = addf(a, b);
      v[h + 1] = subf(a, b);
    }
  } else {
#pragma unroll
    for (int h = 0; h < 8; h += 2) {
      const float2 a = v[h], b = v[h + 1];
      v[h] = addf(a, b);
      v[h + 1] = subf(a, b);
    }
#pragma unroll
    for (int h = 0; h < 8; h += 4) {
      const float2 a0 = v[h], b0 = v[h + 2], a1 = v[h + 1];
      const float2 b1 = make_float2(-v[h + 3].y, v[h + 3].x);  // * (+i)
      v[h] = addf(a0, b0);
      v[h + 2] = subf(a0, b0);
      v[h + 1] = addf(a1, b1);
      v[h + 3] = subf(a1, b1);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float2 a = v[j], b = v[j + 4];
      float2 bp;
      if (j == 0) bp = b;
      else if (j == 1) bp = make_float2((b.x - b.y) * c, (b.x + b.y) * c);     // b * (1 + i)/sqrt2
      else if (j == 2) bp = make_float2(-b.y, b.x);                             // b * (+i)
      else bp = make_float2(-(b.x + b.y) * c, (b.x - b.y) * c);                // b * (-1 + i)/sqrt2
      v[j] = addf(a, bp);
      v[j + 4] = subf(a, bp);
    }
  }
}

// radix-2 / radix-4 tail on 8 contiguous points (8/TAIL independent groups), no twiddles; in place, mirror pair
template <int TAIL, bool INV>
F2_HD void tail(float2 (&u)[8]) {
  if (TAIL == 2) {
#pragma unroll
    for (int g = 0; g < 8; g += 2) {
      const float2 a = u[g], b = u[g + 1];
      u[g] = addf(a, b);
      u[g + 1] = subf(a, b);
    }
  } else if (TAIL == 4) {
#pragma unroll
    for (int g = 0; g < 8; g += 4) {
      if (!INV) {
        const float2 a0 = addf(u[g], u[g + 2]), a1 = addf(u[g + 1], u[g + 3]);
        const float2 d0 = subf(u[g], u[g + 2]), e = subf(u[g + 1], u[g + 3]);
        const float2 d1 = make_float2(e.y, -e.x);
        u[g] = addf(a0, a1);
        u[g + 1] = subf(a0, a1);
        u[g + 2] = addf(d0, d1);
        u[g + 3] = subf(d0, d1);
      } else {
        const float2 a0 = addf(u[g], u[g + 1]), a1 = subf(u[g], u[g + 1]);
        const float2 d0 = addf(u[g + 2], u[g + 3]), e = subf(u[g + 2], u[g + 3]);
        const float2 d1 = make_float2(-e.y, e.x);
        u[g] = addf(a0, d0);
        u[g + 2] = subf(a0, d0);
        u[g + 1] = addf(a1, d1);
        u[g + 3] = subf(a1, d1);
      }
    }
  }
}

template <int M>
struct Plan {
  static constexpr int T = M / 8;                          // threads per transform
  static constexpr int NST = (M >= 4096) ? 4 : 3;          // radix-8 stages
  static constexpr int TAIL = M >> (3 * NST);              // 1 (none), 2 or 4
  static constexpr int SL = 8 * TAIL;                      // span of the last radix-8 stage
  static_assert(M == 512 || M == 1024 || M == 2048 || M == 4096 || M == 8192, "unsupported second-level FFT size");
};

// twiddles w^r, r = 1..7, of a stage of span S for in-block index i: w = exp(-2 pi i * i / S)
template <int S>
F2_HD void stage_twiddles(const float2* __restrict__ tw, int i, float2 (&w)[8]) {
  const int e = i * (kTwLen / S);
  w[1] = tw[e];
  w[2] = tw[2 * e];
  w[4] = tw[4 * e];
  w[3] = cmulf(w[1], w[2]);
  w[5] = cmulf(w[1], w[4]);
  w[6] = cmulf(w[2], w[4]);
  w[7] = cmulf(w[3], w[4]);
}

// ---- forward radix-8 stage of span S (S > 8): butterfly + twiddle + in-place store.  LOAD = false: v already holds
// the inputs x[base + s*j] (first stage, fed from global memory).
template <int S, bool LOAD>
F2_HD void fwd_stage(float2 (&v)[8], float2* sm, const float2* __restrict__ tw, int t) {
  constexpr int s = S / 8;
  const int blk = t / s, i = t % s;
  const int base = blk * S + i;
  if (LOAD) {
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = sm[pad(base + s * j)];
  }
  bf8<false>(v);
  float2 w[8];
  stage_twiddles<S>(tw, i, w);
  sm[pad(base)] = v[0];
#pragma unroll
  for (int r = 1; r < 8; r++) sm[pad(base + s * r)] = cmulf(v[rev3(r)], w[r]);
}

// ---- the turn-around: the last forward step(s) down to single points, then (optionally) the pointwise product and
// the first inverse step(s); works on the 8 contiguous points 8t .. 8t+7.
// After `to_points`, u[q] is the spectrum value stored at position 8t + q (digit-reversed order overall).
template <int TAIL>
F2_HD void load8(float2 (&u)[8], const float2* sm, int t) {
  const float4* p = reinterpret_cast<const float4*>(sm + pad(8 * t));
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const float4 x = p[q];
    u[2 * q] = make_float2(x.x, x.y);
    u[2 * q + 1] = make_float2(x.z, x.w);
  }
}
F2_HD void store8(const float2 (&u)[8], float2* sm, int t) {
  float4* p = reinterpret_cast<float4*>(sm + pad(8 * t));
#pragma unroll
  for (int q = 0; q < 4; q++) p[q] = make_float4(u[2 * q].x, u[2 * q].y, u[2 * q + 1].x, u[2 * q + 1].y);
}
template <int TAIL>
F2_HD void to_points(float2 (&u)[8], const float2* sm, int t) {
  load8<TAIL>(u, sm, t);
  if (TAIL == 1) {  // the span-8 radix-8 stage (no twiddles): position 8t + r <- X_r
    bf8<false>(u);
    float2 y[8];
#pragma unroll
    for (int r = 0; r < 8; r++) y[r] = u[rev3(r)];
#pragma unroll
    for (int r = 0; r < 8; r++) u[r] = y[r];
  } else {
    tail<TAIL, false>(u);
  }
}
template <int TAIL>
F2_HD void from_points(float2 (&u)[8], float2* sm, int t) {
  if (TAIL == 1) {
    float2 y[8];
#pragma unroll
    for (int r = 0; r < 8; r++) y[rev3(r)] = u[r];
    bf8<true>(y);
#pragma unroll
    for (int j = 0; j < 8; j++) u[j] = y[j];
  } else {
    tail<TAIL, true>(u);
  }
  store8(u, sm, t);
}

// ---- inverse radix-8 stage of span S (S > 8): load, conjugate twiddle, butterfly; STORE = false leaves the
// results x[base + s*j] (times the accumulated scale) in v (last stage, written to global memory by the caller).
template <int S, bool STORE>
F2_HD void inv_stage(float2 (&v)[8], float2* sm, const float2* __restrict__ tw, int t) {
  constexpr int s = S / 8;
  const int blk = t / s, i = t % s;
  const int base = blk * S + i;
  float2 w[8];
  stage_twiddles<S>(tw, i, w);
  v[0] = sm[pad(base)];
#pragma unroll
  for (int r = 1; r < 8; r++) v[rev3(r)] = cmulcf(sm[pad(base + s * r)], w[r]);
  bf8<true>(v);
  if (STORE) {
#pragma unroll
    for (int j = 0; j < 8; j++) sm[pad(base + s * j)] = v[j];
  }
}

#undef F2_HD
}  // namespace f2
// =====================================================================================================================
// Radix-16 plan (production path for M = 512 .. 4096):  M = 16 * 16 * L,  L in {2, 4, 8, 16},  T = M/16 threads, 16 points
// per thread.  Three stages
//     A: span M,    radix 16, stride T          (fed from global memory, twiddles w_M^(t r))
//     B: span 16 L, radix 16, stride L          (twiddles w_16L^(i r))
//     C: 16/L contiguous L-point transforms on the thread's 16 contiguous points 16 t .. 16 t + 15, no twiddles
// so a forward + inverse pair crosses shared memory four times (A->B, B->C, C'->B', B'->A') instead of six, which is
// what bounds the radix-8 plan (ncu: l1tex data pipe 87 % busy).  The register slot q of a thread after stage C IS the
// spectrum position 16 t + q by definition: the IR spectra are produced by the same code, so no order is ever undone.
// Twiddles come from per-M tables laid out [4][T] (+ [4][L]) holding w^(i 2^q), q = 0..3, so that a warp reads
// consecutive entries; the other powers are composed with one or two complex products.
namespace r16 {

#define F2_HD __host__ __device__ __forceinline__
using f2::addf;
using f2::cmulcf;
using f2::cmulf;
using f2::subf;

F2_HD int pad(int e) { return e + ((e >> 4) << 1) + ((e >> 7) << 3); }
constexpr int smem_elems(int M) { return M + (M >> 4) * 2 + (M >> 7) * 8; }
constexpr int table_elems(int M) { return 4 * (M / 16) + (M >= 512 ? 4 * (M / 256) : 0); }  // [4][T] stage A, [4][L] stage B

F2_HD constexpr int rev4(int r) { return ((r & 1) << 3) | ((r & 2) << 1) | ((r & 4) >> 1) | ((r >> 3) & 1); }

// radix-8 butterfly on v[O .. O+7] (same conventions as f2::bf8)
template <bool INV, int O, int N>
F2_HD void bf8_at(float2 (&v)[N]) {
  float2 x[8];
#pragma unroll
  for (int j = 0; j < 8; j++) x[j] = v[O + j];
  f2::bf8<INV>(x);
#pragma unroll
  for (int j = 0; j < 8; j++) v[O + j] = x[j];
}

// Radix-16 butterfly.  Forward: in v[j] = x_j, out v[rev4(r)] = sum_j x_j e^{-2 pi i j r / 16}.
// Inverse: in v[rev4(r)] = X_r, out v[j] = sum_r X_r e^{+2 pi i j r / 16}.
template <bool INV>
F2_HD void bf16(float2 (&v)[16]) {
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, c2 = 0.70710678118654752440f;
  // w16^j = (wr[j], -wi[j]) for the forward transform
  const float wr[8] = {1.f, c1, c2, s1, 0.f, -s1, -c2, -c1};
  const float wi[8] = {0.f, s1, c2, c1, 1.f, c1, c2, s1};
  if (!INV) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float2 a = v[j], b = v[j + 8];
      v[j] = addf(a, b);
      const float2 d = subf(a, b);
      if (j == 0) v[8] = d;
      else if (j == 4) v[12] = make_float2(d.y, -d.x);
      else v[j + 8] = make_float2(d.x * wr[j] + d.y * wi[j], d.y * wr[j] - d.x * wi[j]);  // d * (wr - i wi)
    }
    bf8_at<false, 0>(v);
    bf8_at<false, 8>(v);
  } else {
    bf8_at<true, 0>(v);
    bf8_at<true, 8>(v);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float2 a = v[j], b = v[j + 8];
      float2 bp;
      if (j == 0) bp = b;
      else if (j == 4) bp = make_float2(-b.y, b.x);
      else bp = make_float2(b.x * wr[j] - b.y * wi[j], b.y * wr[j] + b.x * wi[j]);          // b * (wr + i wi)
      v[j] = addf(a, bp);
      v[j + 8] = subf(a, bp);
    }
  }
}

// w[r] = w^r, r = 1..15, from the four table entries w^1, w^2, w^4, w^8
F2_HD void compose(float2 w1, float2 w2, float2 w4, float2 w8, float2 (&w)[16]) {
  w[1] = w1; w[2] = w2; w[4] = w4; w[8] = w8;
  w[3] = cmulf(w1, w2);
  w[5] = cmulf(w1, w4);
  w[6] = cmulf(w2, w4);
  w[7] = cmulf(w[3], w4);
  w[9] = cmulf(w1, w8);
  w[10] = cmulf(w2, w8);
  w[11] = cmulf(w[3], w8);
  w[12] = cmulf(w4, w8);
  w[13] = cmulf(w[5], w8);
  w[14] = cmulf(w[6], w8);
  w[15] = cmulf(w[7], w8);
}

template <int M>
struct Plan {
  static constexpr int T = M / 16;
  static constexpr bool HAS_B = M >= 512;           // M = 128, 256 (the first-level transforms of fft_r16.cu): stages A and C only
  static constexpr int L = HAS_B ? M / 256 : M / 16;
  static constexpr int SE = M + (M >> 4) * 2 + (M >> 7) * 8;  // == smem_elems(M), usable in device code
  static_assert(M == 128 || M == 256 || M == 512 || M == 1024 || M == 2048 || M == 4096, "radix-16 plan: M = 16 L or 256 L, L in {2,4,8,16}");
};

// tab: [4][T] then [4][L]
template <int M>
F2_HD void twiddles_a(const float2* __restrict__ tab, int t, float2 (&w)[16]) {
  constexpr int T = Plan<M>::T;
  compose(tab[t], tab[T + t], tab[2 * T + t], tab[3 * T + t], w);
}
template <int M>
F2_HD void twiddles_b(const float2* __restrict__ tab, int i, float2 (&w)[16]) {
  constexpr int T = Plan<M>::T, L = Plan<M>::L;
  const float2* tb = tab + 4 * T;
  compose(tb[i], tb[L + i], tb[2 * L + i], tb[3 * L + i], w);
}

// ---- forward
template <int M>
F2_HD void fwd_a(float2 (&v)[16], float2* sm, const float2* __restrict__ tab, int t) {  // v[j] = x[t + T j]
  constexpr int T = Plan<M>::T;
  bf16<false>(v);
  float2 w[16];
  twiddles_a<M>(tab, t, w);
  sm[pad(t)] = v[0];
#pragma unroll
  for (int r = 1; r < 16; r++) sm[pad(t + T * r)] = cmulf(v[rev4(r)], w[r]);
}
template <int M>
F2_HD void fwd_b(float2* sm, const float2* __restrict__ tab, int t) {
  constexpr int L = Plan<M>::L;
  const int blk = t / L, i = t % L, base = blk * 16 * L + i;
  float2 v[16];
#pragma unroll
  for (int j = 0; j < 16; j++) v[j] = sm[pad(base + L * j)];
  bf16<false>(v);
  float2 w[16];
  twiddles_b<M>(tab, i, w);
  sm[pad(base)] = v[0];
#pragma unroll
  for (int r = 1; r < 16; r++) sm[pad(base + L * r)] = cmulf(v[rev4(r)], w[r]);
}
// stage C on the 16 contiguous points of the thread (in registers)
template <int L, bool INV>
F2_HD void stage_c(float2 (&u)[16]) {
  if (L == 16) {
    bf16<INV>(u);
  } else if (L == 8) {
    bf8_at<INV, 0>(u);
    bf8_at<INV, 8>(u);
  } else {
    float2 a[8], b[8];
#pragma unroll
    for (int q = 0; q < 8; q++) { a[q] = u[q]; b[q] = u[8 + q]; }
    f2::tail<L, INV>(a);
    f2::tail<L, INV>(b);
#pragma unroll
    for (int q = 0; q < 8; q++) { u[q] = a[q]; u[8 + q] = b[q]; }
  }
}
F2_HD void load16(float2 (&u)[16], const float2* sm, int t) {
  const float4* p = reinterpret_cast<const float4*>(sm + pad(16 * t));
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const float4 x = p[q];
    u[2 * q] = make_float2(x.x, x.y);
    u[2 * q + 1] = make_float2(x.z, x.w);
  }
}
F2_HD void store16(const float2 (&u)[16], float2* sm, int t) {
  float4* p = reinterpret_cast<float4*>(sm + pad(16 * t));
#pragma unroll
  for (int q = 0; q < 8; q++) p[q] = make_float4(u[2 * q].x, u[2 * q].y, u[2 * q + 1].x, u[2 * q + 1].y);
}
// Rows of second-level IR spectra (H2) are kept UNPADDED, M float2 per row, with their 16-byte chunks XOR-swizzled: spectrum
// positions 16 t + 2 q, 16 t + 2 q + 1 (register slots 2q, 2q+1 of thread t after stage C) live in chunk 8 t + (q ^ (t & 7)).
// One bulk copy lands a row in shared memory as it lies in global memory, and the eight threads of a quarter-warp read eight
// different bank groups with every 128-bit load (a padded row did the same with 19 % more bytes to move).
F2_HD int swz_chunk(int t, int q) { return 8 * t + (q ^ (t & 7)); }
F2_HD void load16_swz(float2 (&u)[16], const float2* row, int t) {
  const float4* p = reinterpret_cast<const float4*>(row);
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const float4 x = p[swz_chunk(t, q)];
    u[2 * q] = make_float2(x.x, x.y);
    u[2 * q + 1] = make_float2(x.z, x.w);
  }
}
F2_HD void store16_swz(const float2 (&u)[16], float sc, float2* row, int t) {
  float4* p = reinterpret_cast<float4*>(row);
#pragma unroll
  for (int q = 0; q < 8; q++) p[swz_chunk(t, q)] = make_float4(u[2 * q].x * sc, u[2 * q].y * sc, u[2 * q + 1].x * sc, u[2 * q + 1].y * sc);
}
// ---- inverse
template <int M>
F2_HD void inv_b(float2* sm, const float2* __restrict__ tab, int t) {
  constexpr int L = Plan<M>::L;
  const int blk = t / L, i = t % L, base = blk * 16 * L + i;
  float2 w[16];
  twiddles_b<M>(tab, i, w);
  float2 v[16];
  v[0] = sm[pad(base)];
#pragma unroll
  for (int r = 1; r < 16; r++) v[rev4(r)] = cmulcf(sm[pad(base + L * r)], w[r]);
  bf16<true>(v);
#pragma unroll
  for (int j = 0; j < 16; j++) sm[pad(base + L * j)] = v[j];
}
template <int M>
F2_HD void inv_a(float2 (&v)[16], const float2* sm, const float2* __restrict__ tab, int t) {  // out v[j] = M x[t + T j]
  constexpr int T = Plan<M>::T;
  float2 w[16];
  twiddles_a<M>(tab, t, w);
  v[0] = sm[pad(t)];
#pragma unroll
  for (int r = 1; r < 16; r++) v[rev4(r)] = cmulcf(sm[pad(t + T * r)], w[r]);
  bf16<true>(v);
}

#undef F2_HD
}  // namespace r16
}  // namespace gac
