// biquad_lanes.cu — K3d: the only sequential part of BiQuadFilterNode (Nodes/BiQuadFilterNode.cs:136-141):
//     w[n] = x[n] - a1[n]*w[n-1] - a2[n]*w[n-2]          (float32, left to right, unfused; compiled with --fmad=false)
// One lane per (voice, channel); a warp owns one 32-row group of the slab-transposed stream (layout: biquad.cu header).
//
// The recursion is latency-bound (three dependent FP32 ops per frame), so everything else is kept out of the warp's
// instruction stream: 128 frames of the whole group are ONE contiguous 64 KB block that lane 0 fetches with a single
// TMA bulk copy (cp.async.bulk + mbarrier, 3-stage ring); lane r reads (x, a1, a2) of frame i at [i][r] — consecutive lanes,
// consecutive 16 bytes, conflict-free — and the 16 KB tile of w leaves with one bulk store.
#include "gac_kernels.h"

namespace gac {

__device__ __forceinline__ uint32_t bq_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bq_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bq_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bq_mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void bq_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bq_bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

constexpr int kSlab = 128;   // frames per pipeline stage = 4 consecutive 32-frame layout slabs (contiguous in HBM)
constexpr int kStages = 3;   // stages in flight
constexpr int kStageBytes = kSlab * 32 * 16;  // 64 KB
constexpr int kWtBytes = kSlab * 32 * 4;      // 16 KB
constexpr size_t kLanesSmem = (size_t)kStages * kStageBytes + 2 * kWtBytes + 64;

__global__ void __launch_bounds__(32) k_biquad_lanes(const BiquadJob* __restrict__ jobs, int n_jobs, int64_t n_frames,
                                                     const float4* __restrict__ s1t, float* __restrict__ wt_all) {
  extern __shared__ __align__(128) unsigned char lanes_smem[];
  float* wt = reinterpret_cast<float*>(lanes_smem + kStages * kStageBytes);                    // [2][32 frames][32 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(lanes_smem + kStages * kStageBytes + 2 * kWtBytes);  // [kStages]
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
  const int n_slabs = (int)((hi - lo) / kSlab);        // ranges are multiples of 128 frames
  // element (frame n, row r) of group g lives at (g * n_frames + n) * 32 + r in both streams
  const float4* __restrict__ src = s1t + ((size_t)blockIdx.x * (size_t)n_frames + (size_t)lo) * 32;
  float* __restrict__ dst = wt_all + ((size_t)blockIdx.x * (size_t)n_frames + (size_t)lo) * 32;
  const uint32_t bar0 = bq_smem_u32(bars);
  const uint32_t stage0 = bq_smem_u32(lanes_smem);
  const uint32_t wt0 = bq_smem_u32(wt);

  if (lane == 0) {
    for (int s = 0; s < kStages; s++) bq_mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  auto issue = [&](int s) {
    if (s < n_slabs && lane == 0) {
      const int st = s % kStages;
      bq_mbar_expect_tx(bar0 + 8 * st, kStageBytes);
      bq_bulk_g2s(stage0 + st * kStageBytes, src + (size_t)s * (kSlab * 32), kStageBytes, bar0 + 8 * st);
    }
  };

  for (int s = 0; s < kStages - 1; s++) issue(s);
  float w1 = 0.f, w2 = 0.f;
  uint32_t phases = 0u;  // bit st = parity to wait for on stage st
  for (int s = 0; s < n_slabs; s++) {
    __syncwarp();                // every lane is done reading the stage that is refilled next
    issue(s + kStages - 1);
    const int st = s % kStages;
    bq_mbar_wait(bar0 + 8 * st, (phases >> st) & 1u);
    phases ^= 1u << st;
    const int64_t base = lo + (int64_t)s * kSlab;
    const bool act = base >= my_lo && base < my_hi;  // silent-flagged quanta: state untouched (:103-108)
    // the w tile written two iterations ago must have been read out by its bulk store before it is overwritten
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    const float4* __restrict__ rows = reinterpret_cast<const float4*>(lanes_smem + st * kStageBytes) + lane;  // [i][lane]
    float* __restrict__ wrow = wt + (s & 1) * (kSlab * 32) + lane;
    if (act) {
      // 32 frames at a time: all loads first, then the dependent chain from registers, then all stores, so that the
      // 30-cycle LDS latency is paid once per 32 frames (the warp is alone on its SM sub-partition; nothing else hides it)
#pragma unroll 1
      for (int q = 0; q < kSlab / 32; q++) {
        float x[32], p1[32], p2[32], wo[32];
#pragma unroll
        for (int i = 0; i < 32; i++) {
          const float4 r = rows[(q * 32 + i) * 32];        // (x, a1, a2, -)
          x[i] = r.x;
          p1[i] = r.y;
          p2[i] = r.z;
        }
#pragma unroll
        for (int i = 0; i < 32; i++) {
          const float w = x[i] - p1[i] * w1 - p2[i] * w2;  // :137
          w2 = w1;
          w1 = w;
          wo[i] = w;
        }
#pragma unroll
        for (int i = 0; i < 32; i++) wrow[(q * 32 + i) * 32] = wo[i];
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk store
    __syncwarp();
    if (lane == 0) {
      bq_bulk_s2g(dst + (size_t)s * (kSlab * 32), wt0 + (uint32_t)((s & 1) * kWtBytes), kWtBytes);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

void launch_biquad_lanes(const BiquadJob* d_jobs, int n_jobs, int64_t n_frames, const float4* d_s1t, float* d_wt, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_biquad_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLanesSmem);
    attr = true;
  }
  k_biquad_lanes<<<(unsigned)((n_jobs + 15) / 16), 32, kLanesSmem, s>>>(d_jobs, n_jobs, n_frames, d_s1t, d_wt);
}

}  // namespace gac
