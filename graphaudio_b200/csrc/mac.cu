// mac.cu — K6: the spectral multiply-accumulate over the frequency-domain delay line.
//
// Replaces PartitionedConvolver.ProcessSpectralConvolution (PartitionedConvolver.cs:154-223):
//     acc[k] = sum_{p=0}^{P-1} D[(w+p) mod P][k] * H[p][k]        (complex, k = 0..B)
// The reference keeps a ring of the last P input spectra and walks it once per 128-frame quantum.
// Offline, every input block is known up front, so the ring becomes a spectrogram X[b][k] and the
// time axis b is a parallel axis:
//     Y[b][k] = sum_{p=0}^{min(P-1, b)} X[b-p][k] * H[p][k]
//
// Kernels:
//  * k_mac_stream — one pass over X and H per output block, separate mul/sub/add in the reference's
//    order (p ascending, :195-204).  Bit-identical to the CPU oracle for identical spectra.  This is the
//    "T = 1" algorithm whose bytes SURVEY.md §8(d) defines as the roofline contract.
//  * k_mac_tiled — the production kernel.  Each thread owns one frequency bin and T = 16 consecutive
//    output blocks; it keeps a sliding window of 16 input-spectrum values in registers, so every H[p][k]
//    and X[j][k] fetched from shared memory feeds 16 complex MACs.  A CTA covers 128 bins x TB output
//    blocks (TB = 32 or 64); H chunks and X rows are staged into shared memory by TMA bulk copies
//    (cp.async.bulk + mbarrier), double buffered.  Per-CTA traffic is 2 KB per 16*TB*128 complex MACs,
//    i.e. the kernel is bound by the FP32 FMA pipe, not by HBM or L2.  Two accumulation flavours:
//    packed FFMA2 (default) and scalar FFMA.
//  * k_mac_dc — bin 0 of the packed layout holds two purely real bins (DC, Nyquist), whose products are
//    real products, not complex ones.  The tiled kernel treats bin 0 like any other bin (no divergence in
//    the hot loop) and this small kernel overwrites Y[b][0] with the two real convolutions (0.4 % extra work).
#include <cstdio>

#include "gac_kernels.h"

namespace gac {

// --------------------------------------------------------------------------------------------
// k_mac_stream : grid (n_blocks, n_jobs), 128 threads, thread = bin.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_mac_stream(const MacJob* __restrict__ jobs, int stride) {
  const MacJob job = jobs[blockIdx.y];
  const int b = blockIdx.x;
  const int k = threadIdx.x;
  const bool dc = job.has_dc && k == 0;
  float ar = 0.f, ai = 0.f;
  const int pmax = b < job.P - 1 ? b : job.P - 1;  // X[b-p] = 0 for p > b: the delay line starts cleared
  const float2* __restrict__ X = job.X + (int64_t)b * stride + k;
  const float2* __restrict__ H = job.H + k;
#pragma unroll 4
  for (int p = 0; p <= pmax; p++) {
    float2 x = X[-(int64_t)p * stride];
    float2 h = H[(int64_t)p * stride];
    if (dc) {
      // bin 0 packs two purely real bins: DC in .x, Nyquist in .y (imaginary parts are exactly zero in the
      // reference, so its complex product reduces to the real product; the 0*0 terms vanish)
      ar = __fadd_rn(ar, __fmul_rn(x.x, h.x));
      ai = __fadd_rn(ai, __fmul_rn(x.y, h.y));
    } else {
      float re = __fsub_rn(__fmul_rn(x.x, h.x), __fmul_rn(x.y, h.y));  // (dr*ir) - (di*ii)   :197,218
      float im = __fadd_rn(__fmul_rn(x.x, h.y), __fmul_rn(x.y, h.x));  // (dr*ii) + (di*ir)   :201,219
      ar = __fadd_rn(ar, re);
      ai = __fadd_rn(ai, im);
    }
  }
  job.Y[(int64_t)b * stride + k] = make_float2(ar, ai);
}

void launch_mac_stream(const MacJob* d_jobs, int n_jobs, int64_t n_blocks, int stride, cudaStream_t s) {
  if (n_jobs <= 0 || n_blocks <= 0) return;
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
    dim3 grid((unsigned)n_blocks, (unsigned)nj);
    k_mac_stream<<<grid, 128, 0, s>>>(d_jobs + j0, stride);
  }
}

// --------------------------------------------------------------------------------------------
// k_mac_ring : ONE block of a streaming convolver (the per-quantum plugin seam, gac_convolver_process_block):
// acc[k] = sum_{p < P} X[row - p][k] * H[p][k], p ascending, unfused — PartitionedConvolver.cs:154-223 term by term.
// The delay line is a LINEAR history (rows behind `row` are readable and start zeroed: the reference's cleared
// FDL contributes x = 0 products too), so there is no modulo in the loop.
// grid (B / 128, n_jobs, n_blocks), 128 threads, thread = bin; block j of the call reads rows row + j - p and writes Y + j*B.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_mac_ring(const RingMacJob* __restrict__ jobs, int64_t row, int stride) {
  const RingMacJob job = jobs[blockIdx.y];
  const int k = blockIdx.x * 128 + threadIdx.x;
  const bool dc = k == 0;
  float ar = 0.f, ai = 0.f;
  const float2* __restrict__ X = job.X + (row + blockIdx.z) * stride + k;
  const float2* __restrict__ H = job.H + k;
#pragma unroll 8
  for (int p = 0; p < job.P; p++) {
    const float2 x = X[-(int64_t)p * stride];
    const float2 h = H[(int64_t)p * stride];
    if (dc) {  // packed (DC, Nyquist): two real products
      ar = __fadd_rn(ar, __fmul_rn(x.x, h.x));
      ai = __fadd_rn(ai, __fmul_rn(x.y, h.y));
    } else {
      const float re = __fsub_rn(__fmul_rn(x.x, h.x), __fmul_rn(x.y, h.y));
      const float im = __fadd_rn(__fmul_rn(x.x, h.y), __fmul_rn(x.y, h.x));
      ar = __fadd_rn(ar, re);
      ai = __fadd_rn(ai, im);
    }
  }
  job.Y[(int64_t)blockIdx.z * stride + k] = make_float2(ar, ai);
}
void launch_mac_ring(const RingMacJob* d_jobs, int n_jobs, int64_t row, int n_blocks, int B, cudaStream_t s) {
  if (n_jobs <= 0 || n_blocks <= 0) return;
  k_mac_ring<<<dim3((unsigned)(B / 128), (unsigned)n_jobs, (unsigned)n_blocks), 128, 0, s>>>(d_jobs, row, B);
}

// --------------------------------------------------------------------------------------------
// k_mac_dc : Y[b][0] = (sum_p X[b-p][0].x * H[p][0].x , sum_p X[b-p][0].y * H[p][0].y), p ascending.
// grid (ceil(n_blocks / 256), n_jobs), 256 threads, thread = output block.  X column 0 and H column 0 of the job
// are gathered into shared memory once per CTA.
// --------------------------------------------------------------------------------------------
constexpr int kDcThreads = 256;
__global__ void __launch_bounds__(kDcThreads) k_mac_dc(const MacJob* __restrict__ jobs, int64_t n_blocks, int stride, int p_max) {
  const MacJob job = jobs[blockIdx.y];
  if (!job.has_dc) return;
  extern __shared__ float2 dsm[];
  float2* hs = dsm;           // [p_max]
  float2* xs = dsm + p_max;   // [kDcThreads + p_max - 1] : xs[i] = X[b0 - (P-1) + i][0]
  const int P = job.P;
  const int64_t b0 = (int64_t)blockIdx.x * kDcThreads;
  for (int p = threadIdx.x; p < P; p += kDcThreads) hs[p] = job.H[(int64_t)p * stride];
  for (int i = threadIdx.x; i < kDcThreads + P - 1; i += kDcThreads) {
    int64_t j = b0 - (P - 1) + i;
    xs[i] = (j >= 0 && j < n_blocks) ? job.X[j * stride] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  const int64_t b = b0 + threadIdx.x;
  if (b >= n_blocks) return;
  float ar = 0.f, ai = 0.f;
  const float2* xw = xs + threadIdx.x + (P - 1);  // xw[-p] = X[b - p][0]
#pragma unroll 4
  for (int p = 0; p < P; p++) {
    float2 x = xw[-p];
    float2 h = hs[p];
    ar = fmaf(x.x, h.x, ar);
    ai = fmaf(x.y, h.y, ai);
  }
  job.Y[b * stride] = make_float2(ar, ai);
}

// --------------------------------------------------------------------------------------------
// k_mac_tiled
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
// TMA bulk copy global -> shared, completion reported to an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// NS = time sub-tiles per CTA (threads = 128 * NS, TB = 16 * NS output blocks)
template <int NS>
struct MacCfg {
  static constexpr int T = kMacT;
  static constexpr int TB = NS * T;
  static constexpr int GROUPS = TB / 16 + 2;  // ring of 16-row groups: window (TB/16 + 1) + one in flight
  static constexpr int RING = GROUPS * 16;
  static constexpr int THREADS = 128 * NS;
  static constexpr size_t SMEM = (size_t)(RING + 2 * kMacChunk) * 128 * sizeof(float2) + 64;
};

// One pipeline stage for one thread: 16 partitions p = pbase..pbase+15, 16 output blocks each.
// xa points at ring row of block (B0 - pbase) [row 0 of its 16-row group], xb at row 0 of the group below;
// because tiles and stages are 16-aligned, X[B0 - pbase - j] is xa[0] for j = 0 and xb[(16 - j) * 128] for j >= 1.
// Sliding window: win[(t - j) & 15] holds X[B0 + t - pbase - j]; sub-step j loads the one new element into the slot
// that fell out of the window.  All indices are compile-time after unrolling, so the window lives in registers.
__device__ __forceinline__ void mac_stage(float2 (&acc)[kMacT], float2 (&win)[kMacT], const float2* __restrict__ hst,
                                          const float2* __restrict__ xa, const float2* __restrict__ xb) {
#pragma unroll
  for (int j = 0; j < kMacChunk; j++) {
    const float2 h = hst[j * 128];
    win[(16 - j) & 15] = (j == 0) ? xa[0] : xb[(16 - j) * 128];
#pragma unroll
    for (int t = 0; t < kMacT; t++) {
      const float2 x = win[(t - j) & 15];
      acc[t].x = fmaf(x.x, h.x, acc[t].x);
      acc[t].x = fmaf(-x.y, h.y, acc[t].x);
      acc[t].y = fmaf(x.x, h.y, acc[t].y);
      acc[t].y = fmaf(x.y, h.x, acc[t].y);
    }
  }
}

// ---- packed-FP32 flavour (Blackwell FFMA2, PTX fma.rn.f32x2): one instruction = two FMAs on an aligned register pair.
// SASS FFMA2 can broadcast a scalar register to both halves (R.F32) and swap the halves of a pair (.LO_HI), so
//   A += h.re * (x.re, x.im)          -> (sum h.re*x.re, sum h.re*x.im)
//   B += h.im * swap(x.re, x.im)      -> (sum h.im*x.im, sum h.im*x.re)
// need no operand shuffling at all; the complex sum is re = A.lo - B.lo, im = A.hi + B.hi at the end.  Every operand is
// an aligned even/odd pair, so the register-bank conflicts that cap the scalar FFMA form at ~63 % of the FMA pipe
// (acc.x/x.x vs acc.x/x.y parities cannot all differ while (x.x, x.y) arrive as an LDS.64 pair) do not arise.
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__device__ __forceinline__ void mac_stage_f2(uint64_t (&A)[kMacT], uint64_t (&Bq)[kMacT], float2 (&win)[kMacT],
                                             const float2* __restrict__ hst, const float2* __restrict__ xa,
                                             const float2* __restrict__ xb) {
#pragma unroll
  for (int j = 0; j < kMacChunk; j++) {
    const float2 h = hst[j * 128];
    win[(16 - j) & 15] = (j == 0) ? xa[0] : xb[(16 - j) * 128];
    const uint64_t h1 = pk2(h.x, h.x);
    const uint64_t h2 = pk2(h.y, h.y);
#pragma unroll
    for (int t = 0; t < kMacT; t++) {
      const float2 x = win[(t - j) & 15];
      A[t] = ffma2(h1, pk2(x.x, x.y), A[t]);
      Bq[t] = ffma2(h2, pk2(x.y, x.x), Bq[t]);
    }
  }
}

template <int NS, bool F2>
__global__ void __launch_bounds__(MacCfg<NS>::THREADS, (NS <= 2 ? 2 : 1))
    k_mac_tiled(const MacJob* __restrict__ jobs, const MacTile* __restrict__ tiles, int stride) {
  using Cfg = MacCfg<NS>;
  constexpr int T = Cfg::T;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* xs = reinterpret_cast<float2*>(smem_raw);                        // [RING][128]
  float2* hs = xs + (size_t)Cfg::RING * 128;                              // [2][16][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(hs + 2 * kMacChunk * 128);  // [2]
  const uint32_t bar_base = smem_u32(bars);

  const MacTile tile = tiles[blockIdx.x];
  const MacJob job = jobs[tile.job];
  const int b0 = tile.b0;  // multiple of TB (hence of 16)
  const int k = threadIdx.x & 127;
  const int sub = threadIdx.x >> 7;
  // stages: chunks of 16 partitions, skipping chunks that only meet X rows before block 0 (all zero)
  const int p16 = (job.P + kMacChunk - 1) / kMacChunk;
  const int causal = (b0 + Cfg::TB) / 16;
  const int n_stages = p16 < causal ? p16 : causal;
  // 16-row group g (absolute index (j + TB) / 16 for block j >= -TB) lives in ring slot g % GROUPS
  const int G0 = (b0 + Cfg::TB) / 16;  // group of block b0
  const float2* __restrict__ gX = job.X;
  const float2* __restrict__ gH = job.H;

  if (threadIdx.x == 0) {
    mbar_init(bar_base, 1);
    mbar_init(bar_base + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // stage 0: H chunk 0 and the initial X window = groups G0-1 .. G0+TB/16-1 (blocks [b0-16, b0+TB))
    const uint32_t bar = bar_base;
    constexpr int NG0 = Cfg::TB / 16 + 1;
    mbar_expect_tx(bar, (uint32_t)(kMacChunk + 16 * NG0) * 1024u);
    const uint32_t hdst = smem_u32(hs);
    if (stride == 128) {
      tma_bulk_g2s(hdst, gH, kMacChunk * 1024u, bar);
      for (int g = 0; g < NG0; g++) {
        const int grp = G0 - 1 + g;
        tma_bulk_g2s(smem_u32(xs + (size_t)(grp % Cfg::GROUPS) * 16 * 128), gX + (int64_t)(b0 - 16 + 16 * g) * 128, 16u * 1024u, bar);
      }
    } else {
      for (int r = 0; r < kMacChunk; r++) tma_bulk_g2s(hdst + r * 1024u, gH + (int64_t)r * stride, 1024u, bar);
      for (int g = 0; g < NG0; g++) {
        const int grp = G0 - 1 + g;
        const uint32_t xdst = smem_u32(xs + (size_t)(grp % Cfg::GROUPS) * 16 * 128);
        for (int r = 0; r < 16; r++) tma_bulk_g2s(xdst + r * 1024u, gX + (int64_t)(b0 - 16 + 16 * g + r) * stride, 1024u, bar);
      }
    }
  }

  float2 acc[F2 ? 1 : T];
  uint64_t accA[F2 ? T : 1], accB[F2 ? T : 1];
  float2 win[T];
#pragma unroll
  for (int t = 0; t < (F2 ? 1 : T); t++) acc[t] = make_float2(0.f, 0.f);
#pragma unroll
  for (int t = 0; t < (F2 ? T : 1); t++) accA[t] = accB[t] = 0ull;
  const int B0 = b0 + sub * T;  // first output block of this thread; its group is G0 + sub

  uint32_t phase0 = 0u, phase1 = 0u;
  for (int st = 0; st < n_stages; st++) {
    const int buf = st & 1;
    if (buf == 0) { mbar_wait(bar_base, phase0); phase0 ^= 1u; } else { mbar_wait(bar_base + 8, phase1); phase1 ^= 1u; }
    __syncthreads();  // everyone is done with stage st-1: its H buffer and the X group that left the window are free
    if (threadIdx.x == 0 && st + 1 < n_stages) {
      const int nb = (st + 1) & 1;
      const uint32_t bar = bar_base + 8 * nb;
      const int p1 = (st + 1) * kMacChunk;
      const int jx = b0 - p1 - 16;          // new X group: blocks [b0 - 16(st+1) - 16, +16)
      const int grp = G0 - (st + 1) - 1;    // its absolute group index (>= 0)
      const uint32_t hdst = smem_u32(hs + (size_t)nb * kMacChunk * 128);
      const uint32_t xdst = smem_u32(xs + (size_t)(grp % Cfg::GROUPS) * 16 * 128);
      mbar_expect_tx(bar, 2u * kMacChunk * 1024u);
      if (stride == 128) {
        tma_bulk_g2s(hdst, gH + (int64_t)p1 * 128, kMacChunk * 1024u, bar);
        tma_bulk_g2s(xdst, gX + (int64_t)jx * 128, 16u * 1024u, bar);
      } else {
        for (int r = 0; r < kMacChunk; r++) tma_bulk_g2s(hdst + r * 1024u, gH + (int64_t)(p1 + r) * stride, 1024u, bar);
        for (int r = 0; r < 16; r++) tma_bulk_g2s(xdst + r * 1024u, gX + (int64_t)(jx + r) * stride, 1024u, bar);
      }
    }
    // this thread's rows for the stage: block B0 - 16*st sits in group G0 + sub - st (row 0); older rows in the group below
    const int ga = G0 + sub - st;
    const float2* __restrict__ xa = xs + (size_t)(ga % Cfg::GROUPS) * 16 * 128 + k;
    const float2* __restrict__ xb = xs + (size_t)((ga - 1) % Cfg::GROUPS) * 16 * 128 + k;
    if (st == 0) {
      // initial window: slots 1..15 <- X[B0 + t], rows t of group `ga`
#pragma unroll
      for (int t = 1; t < T; t++) win[t] = xa[t * 128];
    }
    const float2* __restrict__ hst = hs + (size_t)buf * kMacChunk * 128 + k;
    if constexpr (F2) {
      mac_stage_f2(accA, accB, win, hst, xa, xb);
    } else {
      mac_stage(reinterpret_cast<float2(&)[kMacT]>(acc), win, hst, xa, xb);
    }
  }
  float2* __restrict__ Y = job.Y + (int64_t)B0 * stride + k;
  if constexpr (F2) {
#pragma unroll
    for (int t = 0; t < T; t++) {
      float al, ah, bl, bh;
      upk2(accA[t], al, ah);
      upk2(accB[t], bl, bh);
      Y[(int64_t)t * stride] = make_float2(al - bl, ah + bh);
    }
  } else {
#pragma unroll
    for (int t = 0; t < T; t++) Y[(int64_t)t * stride] = acc[t];
  }
}

int mac_tile_blocks(int variant) { return variant == 64 ? 64 : 32; }

template <int NS, bool F2>
static void launch_mac_tiled_t(const MacJob* d_jobs, const MacTile* d_tiles, int n_tiles, int stride, cudaStream_t s) {
  using Cfg = MacCfg<NS>;
  GAC_SMEM_OPT_IN((k_mac_tiled<NS, F2>), Cfg::SMEM);
  k_mac_tiled<NS, F2><<<n_tiles, Cfg::THREADS, Cfg::SMEM, s>>>(d_jobs, d_tiles, stride);
}

void launch_mac_tiled(const MacJob* d_jobs, int n_jobs, const MacTile* d_tiles, int n_tiles, int64_t n_blocks, int p_max, int stride,
                      int tile_blocks, int flavour, cudaStream_t s) {
  if (n_tiles <= 0) return;
  const bool f2 = flavour != 2;  // 0 = FFMA2 (default), 2 = scalar FFMA
  if (tile_blocks == 64) {
    if (f2) launch_mac_tiled_t<4, true>(d_jobs, d_tiles, n_tiles, stride, s);
    else launch_mac_tiled_t<4, false>(d_jobs, d_tiles, n_tiles, stride, s);
  } else {
    if (f2) launch_mac_tiled_t<2, true>(d_jobs, d_tiles, n_tiles, stride, s);
    else launch_mac_tiled_t<2, false>(d_jobs, d_tiles, n_tiles, stride, s);
  }
  // bin 0 = (DC, Nyquist): two real convolutions, written over the generic result
  const size_t smem = (size_t)(2 * p_max + kDcThreads) * sizeof(float2);
  GAC_SMEM_OPT_IN(k_mac_dc, smem);
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
    dim3 grid((unsigned)((n_blocks + kDcThreads - 1) / kDcThreads), (unsigned)nj);
    k_mac_dc<<<grid, kDcThreads, smem, s>>>(d_jobs + j0, n_blocks, stride, p_max);
  }
}

}  // namespace gac
