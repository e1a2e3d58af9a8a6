// mac.cu — K6: the spectral multiply-accumulate over the frequency-domain delay line.
//
// Replaces PartitionedConvolver.ProcessSpectralConvolution (PartitionedConvolver.cs:154-223):
//     acc[k] = sum_{p=0}^{P-1} D[(w+p) mod P][k] * H[p][k]        (complex, k = 0..B)
// The reference keeps a ring of the last P input spectra and walks it once per 128-frame quantum.
// Offline, every input block is known up front, so the ring becomes a spectrogram X[b][k] and the
// time axis b is a parallel axis:
//     Y[b][k] = sum_{p=0}^{min(P-1, b)} X[b-p][k] * H[p][k]
//
// Two kernels:
//  * k_mac_stream — one pass over X and H per output block, separate mul/sub/add in the reference's
//    order (p ascending, :195-204).  Bit-identical to the CPU oracle for identical spectra.  This is the
//    "T = 1" algorithm whose bytes SURVEY.md §8(d) defines as the roofline contract.
//  * k_mac_tiled — the production kernel.  Each thread owns one frequency bin and T = 16 consecutive
//    output blocks; it keeps a sliding window of 16 input spectra values in registers, so every H[p][k]
//    and X[j][k] fetched from shared memory feeds 16 complex MACs.  A CTA covers 128 bins x TB output
//    blocks (TB = 32 or 64); H chunks and X rows are staged into shared memory by TMA bulk copies
//    (cp.async.bulk + mbarrier), double buffered.  Per-CTA traffic is 2 KB per 16*TB*128 complex MACs,
//    i.e. the kernel is bound by the FP32 FMA pipe, not by HBM or L2.
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
template <bool DC>
__device__ __forceinline__ void mac_stage(float2 (&acc)[kMacT], float2 (&win)[kMacT], const float2* __restrict__ hst,
                                          const float2* __restrict__ xa, const float2* __restrict__ xb, bool dc) {
#pragma unroll
  for (int j = 0; j < kMacChunk; j++) {
    const float2 h = hst[j * 128];
    win[(16 - j) & 15] = (j == 0) ? xa[0] : xb[(16 - j) * 128];
    float ha = h.x, hb = h.y, hc = h.y, hd = h.x;
    if (DC) {
      // bin 0 = two independent real bins (DC, Nyquist): re += x.re*h.re ; im += x.im*h.im
      hb = dc ? 0.f : h.y;
      hc = dc ? 0.f : h.y;
      hd = dc ? h.y : h.x;
    }
#pragma unroll
    for (int t = 0; t < kMacT; t++) {
      const float2 x = win[(t - j) & 15];
      acc[t].x = fmaf(x.x, ha, acc[t].x);
      acc[t].x = fmaf(-x.y, hb, acc[t].x);
      acc[t].y = fmaf(x.x, hc, acc[t].y);
      acc[t].y = fmaf(x.y, hd, acc[t].y);
    }
  }
}

template <int NS>
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
  const bool warp_dc = job.has_dc && k < 32;  // warp-uniform
  const bool dc = job.has_dc && k == 0;
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

  float2 acc[T];
  float2 win[T];
#pragma unroll
  for (int t = 0; t < T; t++) acc[t] = make_float2(0.f, 0.f);
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
    if (warp_dc) {
      mac_stage<true>(acc, win, hst, xa, xb, dc);
    } else {
      mac_stage<false>(acc, win, hst, xa, xb, false);
    }
  }
  float2* __restrict__ Y = job.Y + (int64_t)B0 * stride + k;
#pragma unroll
  for (int t = 0; t < T; t++) Y[(int64_t)t * stride] = acc[t];
}

int mac_tile_blocks(int variant) { return variant == 64 ? 64 : 32; }

void launch_mac_tiled(const MacJob* d_jobs, const MacTile* d_tiles, int n_tiles, int stride, int tile_blocks, cudaStream_t s) {
  if (n_tiles <= 0) return;
  if (tile_blocks == 64) {
    using Cfg = MacCfg<4>;
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(k_mac_tiled<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
      attr = true;
    }
    k_mac_tiled<4><<<n_tiles, Cfg::THREADS, Cfg::SMEM, s>>>(d_jobs, d_tiles, stride);
  } else {
    using Cfg = MacCfg<2>;
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(k_mac_tiled<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
      attr = true;
    }
    k_mac_tiled<2><<<n_tiles, Cfg::THREADS, Cfg::SMEM, s>>>(d_jobs, d_tiles, stride);
  }
}

}  // namespace gac
