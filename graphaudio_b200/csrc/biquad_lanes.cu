// biquad_lanes.cu — K3d: the only sequential part of BiQuadFilterNode (Nodes/BiQuadFilterNode.cs:136-141):
//     w[n] = x[n] - a1[n]*w[n-1] - a2[n]*w[n-2]          (float32, left to right, unfused; compiled with --fmad=false)
// One lane per (voice, channel); a warp owns one 32-row group of the slab-transposed stream (layout: biquad.cu header).
//
// Time is cut into segments that run CONCURRENTLY, speculate-and-verify, still bit-exact:
//   * segment k > 0 starts kWarm frames early from the state (0, 0) and discards what it computes before its own first
//     frame.  A stable biquad forgets its state, and in float32 two trajectories fed the same input do not merely get
//     close, they become bit-identical (measured: after ~300 frames for the BASELINE lowpass sweep, a few thousand for
//     Q = 8 at 1 kHz), after which they stay identical forever because the recursion is deterministic;
//   * every segment records the state it started its own frames with and the state it ended with; k_biquad_verify checks
//     end[k-1] == start[k] BITWISE along the chain.  The first segment is exact by construction, so if every link matches,
//     every sample equals the sequential result bit for bit;
//   * if a link does not match (very low cutoff / very high Q never merge), the repair launch recomputes that 32-row group
//     sequentially from the last verified state: the result is always exact, only the speed-up is lost for that group.
//
// The recursion is latency-bound (three dependent FP32 ops per frame), so everything else is kept out of the dependent
// chain: 64 frames of the whole group are two contiguous 32 KB blocks — (x, a1, a2) and (b0, b1, b2) — that lane 0 fetches with
// TMA bulk copies (cp.async.bulk + mbarrier, 3-stage ring); lane r reads frame i at [i][r] — consecutive lanes, consecutive
// 16 bytes, conflict-free — and y = b0*w + b1*w[n-1] + b2*w[n-2] (:138) is formed from the chain's results while the next
// loads are in flight; each lane stores its row's 32 results as one 128-byte line.
#include <algorithm>
#include <cstdlib>

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

constexpr int kSlab = 32;    // frames per pipeline stage = one 32-frame layout slab
constexpr int kStages = 3;   // stages in flight (96 KB per CTA: two CTAs, i.e. two concurrent segments, per SM)
constexpr int kWarmSlabs = 256;  // warm-up of a speculative segment: 256 * 32 = 8192 frames
constexpr int kLaneSlots = 2 * 148;  // concurrent single-warp CTAs
constexpr int kTileBytes = kSlab * 32 * 16;   // one stream's stage in the per-row layout: 16 KB (8 KB per voice)
constexpr int kStageBytes = 2 * kTileBytes;   // (x, a1, a2) tile + (b0, b1, b2) tile
constexpr size_t kLanesSmem = (size_t)kStages * kStageBytes + 64 + kStages * 256;  // stages, barriers, staged speculative states (repair)

// states: float2 [groups][n_seg][2 (start, end)][32 rows], then slab_states: float2 [groups][n_frames / 32][32 rows] (the state
// after every 32-frame slab of the speculative pass).  flags: int [groups] first_bad, [groups] layout, [groups][n_seg] link_bad.
//
// REPAIR = false: grid (groups, n_seg), segment blockIdx.y runs speculatively.
// REPAIR = true : grid (groups).  For every link the verification found broken, the group re-runs from the last true state —
//   the end state of the previous segment — and compares its state with the speculative pass's after every slab: the moment
//   all 32 rows coincide bitwise it has re-joined the speculative trajectory, everything behind that point is already right,
//   and it moves on to the next broken link.  A segment that never re-joins hands its state on to the next one.
template <bool REPAIR, bool WIDE>
__global__ void __launch_bounds__(32) k_biquad_lanes(const BiquadJob* __restrict__ jobs, int n_jobs, int64_t n_frames,
                                                     const float4* __restrict__ s1t, const float4* __restrict__ s2t, int seg_slabs, int n_seg, int warm_slabs,
                                                     float2* __restrict__ states, float2* __restrict__ slab_states, const int* __restrict__ first_bad,
                                                     const int* __restrict__ wide_flags, const int* __restrict__ link_bad) {
  extern __shared__ __align__(128) unsigned char lanes_smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(lanes_smem + kStages * kStageBytes);  // [kStages]
  float2* spec_sm = reinterpret_cast<float2*>(lanes_smem + kStages * kStageBytes + 64);  // [kStages][32]: the speculative pass's state after the slab
  if ((wide_flags[blockIdx.x] != 0) != WIDE) return;  // the other instantiation serves this group
  const int lane = threadIdx.x;
  const int j = blockIdx.x * 16 + (lane >> 1);
  const bool valid = j < n_jobs;
  const int64_t my_lo = valid ? jobs[j].lo : 0, my_hi = valid ? jobs[j].hi : 0;
  float* __restrict__ my_sig = valid ? jobs[j].sig[lane & 1] : nullptr;
  int64_t lo = my_hi > my_lo ? my_lo : INT64_MAX, hi = my_hi > my_lo ? my_hi : 0;
  for (int o = 16; o > 0; o >>= 1) {
    int64_t lo2 = __shfl_xor_sync(0xffffffffu, lo, o), hi2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = lo2 < lo ? lo2 : lo;
    hi = hi2 > hi ? hi2 : hi;
  }
  if (hi <= lo) return;
  const int total_slabs = (int)((hi - lo + kSlab - 1) / kSlab);  // ranges are multiples of 128 frames
  float2* st_group = states + (size_t)blockIdx.x * n_seg * 64;
  float2* slab_group = slab_states + ((size_t)blockIdx.x * (size_t)(n_frames / kSlab) + (size_t)(lo / kSlab)) * 32;  // [slab relative to lo][row]
  // stream layout of this group (biquad.cu header): per row — element (frame n, row r) at n * 32 + r — or per voice —
  // element (frame n, voice v) at n * 16 + v, carrying (a1, a2, xL, xR); both relative to the group base g * n_frames * 32
  constexpr int RS = WIDE ? 32 : 16;                  // elements per frame
  const int ridx = WIDE ? lane : (lane >> 1);         // this lane's element within a frame
  const bool right = (lane & 1) != 0;
  constexpr uint32_t tile_bytes = (uint32_t)(kSlab * RS * 16);
  const size_t group_elem = (size_t)blockIdx.x * (size_t)n_frames * 32 + (size_t)lo * RS;
  const uint32_t bar0 = bq_smem_u32(bars);
  const uint32_t stage0 = bq_smem_u32(lanes_smem);

  if (lane == 0) {
    for (int s = 0; s < kStages; s++) bq_mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t phases = 0u;  // bit st = parity to wait for on stage st (persists across runs)
  float w1 = 0.f, w2 = 0.f;

  // One run: slabs [s_first, s_end) relative to lo, the first n_warm of them recursion only (nothing written).
  // REJOIN: stop as soon as the state after a slab equals the one the speculative pass recorded there; returns true then.
  auto run = [&](const int s_first, const int n_warm, const int s_end, const bool rejoin) -> bool {
    const int n_slabs = s_end - s_first;
    const float4* __restrict__ src1 = s1t + group_elem + (size_t)s_first * kSlab * RS;
    const float4* __restrict__ src2 = s2t + group_elem + (size_t)s_first * kSlab * RS;
    auto issue = [&](int s) {
      if (s < n_slabs && lane == 0) {
        const int st = s % kStages;
        const bool own = s >= n_warm;  // the output coefficients are not needed while warming up
        const bool spec = REPAIR && rejoin && own;  // the state to compare with travels with the slab (a dependent global load
                                                    // per slab would cost more than the slab's recursion)
        bq_mbar_expect_tx(bar0 + 8 * st, (own ? 2 * tile_bytes : tile_bytes) + (spec ? 256u : 0u));
        bq_bulk_g2s(stage0 + st * kStageBytes, src1 + (size_t)s * (kSlab * RS), tile_bytes, bar0 + 8 * st);
        if (own) bq_bulk_g2s(stage0 + st * kStageBytes + kTileBytes, src2 + (size_t)s * (kSlab * RS), tile_bytes, bar0 + 8 * st);
        if (spec) bq_bulk_g2s(bq_smem_u32(spec_sm + st * 32), slab_group + (size_t)(s_first + s) * 32, 256u, bar0 + 8 * st);
      }
    };
    for (int s = 0; s < kStages - 1; s++) issue(s);
    for (int s = 0; s < n_slabs; s++) {
      if (s == n_warm && !REPAIR) st_group[((size_t)blockIdx.y * 2 + 0) * 32 + lane] = make_float2(w1, w2);  // state at the segment's first own frame
      __syncwarp();                // every lane is done reading the stage that is refilled next
      issue(s + kStages - 1);
      const int st = s % kStages;
      bq_mbar_wait(bar0 + 8 * st, (phases >> st) & 1u);
      phases ^= 1u << st;
      const int64_t base = lo + (int64_t)(s_first + s) * kSlab;
      const bool act = base >= my_lo && base < my_hi;  // silent-flagged quanta: state untouched (:103-108)
      const bool own = s >= n_warm;
      const float4* __restrict__ rows = reinterpret_cast<const float4*>(lanes_smem + st * kStageBytes) + ridx;               // [i][element]
      const float4* __restrict__ outs = reinterpret_cast<const float4*>(lanes_smem + st * kStageBytes + kTileBytes) + ridx;  // (b0, b1, b2, -)
      if (act) {
        // 32 frames: all loads first, then the dependent chain from registers, so that the 30-cycle LDS latency is paid once per
        // slab (the warp shares its SM sub-partition with at most one other); then y, off the critical path
        float x[32], p1[32], p2[32], wo[34];
#pragma unroll
        for (int i = 0; i < 32; i++) {
          const float4 r = rows[i * RS];
          x[i] = WIDE ? r.x : (right ? r.w : r.z);  // per row: (x, a1, a2, -); per voice: (a1, a2, xL, xR)
          p1[i] = WIDE ? r.y : r.x;
          p2[i] = WIDE ? r.z : r.y;
        }
        wo[0] = w2;
        wo[1] = w1;
#pragma unroll
        for (int i = 0; i < 32; i++) {
          const float w = x[i] - p1[i] * w1 - p2[i] * w2;  // :137
          w2 = w1;
          w1 = w;
          wo[i + 2] = w;
        }
        if (own) {
          float* __restrict__ dst = my_sig + base;
#pragma unroll
          for (int i4 = 0; i4 < 8; i4++) {
            float y[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const int i = 4 * i4 + e;
              const float4 b = outs[i * RS];
              y[e] = b.x * wo[i + 2] + b.y * wo[i + 1] + b.z * wo[i];  // :138  y = b0*w + b1*w1 + b2*w2
            }
            *reinterpret_cast<float4*>(dst + 4 * i4) = make_float4(y[0], y[1], y[2], y[3]);
          }
        }
      }
      if (own) {
        float2* rec = slab_group + (size_t)(s_first + s) * 32 + lane;
        if (REPAIR && rejoin) {
          const float2 spec = spec_sm[st * 32 + lane];
          const bool same = __float_as_uint(spec.x) == __float_as_uint(w1) && __float_as_uint(spec.y) == __float_as_uint(w2);
          if (__all_sync(0xffffffffu, same)) {
            // re-joined: drain the stages already requested, so that the next run starts from quiescent barriers
            for (int o = s + 1; o < n_slabs && o <= s + kStages - 1; o++) {
              const int so = o % kStages;
              bq_mbar_wait(bar0 + 8 * so, (phases >> so) & 1u);
              phases ^= 1u << so;
            }
            __syncwarp();
            return true;
          }
        }
        *rec = make_float2(w1, w2);
      }
    }
    return false;
  };

  if (!REPAIR) {
    const int seg = (int)blockIdx.y;
    const int s_own = seg * seg_slabs;
    if (s_own >= total_slabs) return;
    const int s_end = s_own + seg_slabs < total_slabs ? s_own + seg_slabs : total_slabs;
    const int s_first = seg == 0 ? s_own : (s_own - warm_slabs > 0 ? s_own - warm_slabs : 0);
    run(s_first, s_own - s_first, s_end, false);
    st_group[((size_t)seg * 2 + 1) * 32 + lane] = make_float2(w1, w2);
  } else {
    bool carry = false;
    for (int seg = first_bad[blockIdx.x]; seg < n_seg; seg++) {
      const int s_own = seg * seg_slabs;
      if (s_own >= total_slabs) break;
      if (!carry && !link_bad[(size_t)blockIdx.x * n_seg + seg]) continue;
      if (!carry) {  // the last true state: what segment seg-1 ended with (verified, or repaired and re-joined before its end)
        const float2 e = st_group[((size_t)(seg - 1) * 2 + 1) * 32 + lane];
        w1 = e.x;
        w2 = e.y;
      }
      const int s_end = s_own + seg_slabs < total_slabs ? s_own + seg_slabs : total_slabs;
      carry = !run(s_own, 0, s_end, true);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Shared-coefficient variant (biquad.cu: "Shared-coefficient path"): the 16 voices of a group run the same filter over the same
// non-silent range, so the coefficients come from ONE class stream and the samples are read where they lie (the source buffer
// itself for a rate-1 source, BiquadJob::in) — 16 bytes per voice and frame instead of ~90.
//
// CTA = TWO warps with different jobs, because the recursion w = x - a1*w1 - a2*w2 (:137) is a 12-cycle dependent chain per frame and
// everything a warp does besides it lands in the same in-order pipe (measured: 20 cycles per frame for the bare chain, 60-80 with
// the loads, the output sums and the stores in the same warp):
//   warp 0, the CHAIN warp: one lane per (voice, channel) row; per 64-frame stage it loads its operands into registers one 16-frame
//     step ahead (across the loop's back edge: ptxas sinks loads that are consumed in the same trip down to their use), runs the
//     chain and writes w over the samples it consumed; it also records / compares the slab states of the speculate-verify-repair
//     protocol (same protocol, same records, same verification kernel as the stream-fed variant above);
//   warp 1, the MOVER warp: brings the stages in (cp.async, completion signalled to the chain warp through an mbarrier), and, for
//     chunks whose output counts, forms y = b0*w + b1*w[n-1] + b2*w[n-2] (:138) — embarrassingly parallel once w exists — and
//     stores it.  Both directions are transpositions (rows are contiguous in time, lanes own rows); done naively every memory
//     instruction is 32 separate accesses, and per-lane bulk copies are worse (cp.async.bulk takes uniform operands: 32 row
//     addresses become a 32-trip serialisation loop).  Here instruction (m, k) lets the four lanes of quad q = lane / 4 move the
//     m-th 32-byte sector of row q + 8k, 8 bytes each: a lane only ever needs the pointers of its quad's four rows.
// Stage row: [8 bytes unused | w[-2], w[-1] of the previous chunk | 64 samples] = 272 bytes (the 16-byte lead keeps every lane's
// LDS.128 / STS.128 aligned and conflict-free, and gives the mover warp its two samples of history contiguously).
constexpr int kChunk = kBqChunkFrames;  // 64
constexpr int kShStages = 4;  // the chain warp works on chunk s, chunks s + 1, s + 2 have landed or are landing, chunk s - 1's
                              // output sums are being formed: its stage is refilled only behind them
constexpr int kShRowBytes = kChunk * 4 + 16;
constexpr int kShXBytes = 32 * kShRowBytes;
constexpr int kShStageBytes = kShXBytes + kBqChunkBytes + (kChunk / 32) * 256;
constexpr size_t kShSmem = (size_t)kShStages * kShStageBytes + 128;  // + 2 x 4 mbarriers + the per-stage control words
constexpr int kShSlots = 4 * 148;     // 49.5 KB per CTA: four per SM = one chain warp and one mover warp per scheduler
static_assert(kShStageBytes % 16 == 0, "stage alignment");

__device__ __forceinline__ void bq_cp8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void bq_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void bq_cp_arrive(uint32_t bar) {  // arrives on `bar` once this thread's earlier cp.async have landed
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bq_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bq_cta_sync() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

template <bool REPAIR>
__global__ void __launch_bounds__(64) k_biquad_lanes_shared(const BiquadJob* __restrict__ jobs, const BqGroup* __restrict__ groups,
                                                            const unsigned char* __restrict__ cs, size_t cs_stride, int64_t n_frames, int seg_chunks,
                                                            int n_seg, int warm_chunks, float2* __restrict__ states, float2* __restrict__ slab_states,
                                                            const int* __restrict__ first_bad, const int* __restrict__ link_bad) {
  extern __shared__ __align__(128) unsigned char sh_smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sh_smem + kShStages * kShStageBytes);   // full[3], done[3]
  volatile int* ctrl = reinterpret_cast<volatile int*>(sh_smem + kShStages * kShStageBytes + 16 * kShStages);  // [st]: frames done | rejoined << 16
  const int lane = threadIdx.x & 31;
  const bool mover = threadIdx.x >= 32;
  const int g = blockIdx.x;
  const BqGroup grp = groups[g];
  const bool valid = (lane >> 1) < grp.count;
  const int j = grp.first + (valid ? (lane >> 1) : 0);  // lanes beyond the group's last job shadow its first one (nothing is stored for them)
  const int ch = lane & 1;
  const int64_t lo = jobs[j].lo, hi = jobs[j].hi;  // the same for every job of the group
  if (hi <= lo) return;
  const float* xin = jobs[j].in[ch] ? jobs[j].in[ch] : jobs[j].sig[ch];
  float* yout = jobs[j].sig[ch];
  const unsigned char* __restrict__ cls = cs + (size_t)grp.cls * cs_stride + (size_t)(lo / kChunk) * kBqChunkBytes;
  const int total_chunks = (int)((hi - lo) / kChunk);  // ranges are multiples of 128 frames
  float2* st_group = states + (size_t)g * n_seg * 64;
  float2* slab_group = slab_states + ((size_t)g * (size_t)(n_frames / kSlab) + (size_t)(lo / kSlab)) * 32;  // [32-frame slab relative to lo][row]
  const uint32_t stage0 = bq_smem_u32(sh_smem);
  const uint32_t bar_full = bq_smem_u32(bars), bar_done = bar_full + 8 * kShStages;
  // the mover's rows: the eight lanes of octet o = lane / 8 move one 128-byte line of row o + 4k per instruction, 16 bytes each
  // (lane r owns row r: the pointers come from there).  Whole lines, not sectors: what limits a latency-bound stream per SM is the
  // number of requests in flight, and a request is a line whether one or four of its sectors are wanted.
  const int oct = lane >> 3, t8 = lane & 7;
  const float* xrow[8];
  float* yrow[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    xrow[k] = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(xin), oct + 4 * k)) + 4 * t8;
    yrow[k] = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(yout), oct + 4 * k)) + 4 * t8;
  }
  const int n_rows = 2 * grp.count;  // rows that are stored
  float w1 = 0.f, w2 = 0.f;         // (chain warp) the recursion's state

  // One run: chunks [c_first, c_end) relative to lo, the first n_warm of them recursion only (nothing written).
  // rejoin: stop as soon as the state after a slab equals the one the speculative pass recorded there; returns true then.
  auto run = [&](const int c_first, const int n_warm, const int c_end, const bool rejoin) -> bool {
    const int n_chunks = c_end - c_first;
    // both warps enter a run with quiescent barriers: nothing in flight, every phase at 0
    bq_cta_sync();
    if (threadIdx.x == 0) {
      for (int s = 0; s < 2 * kShStages; s++) bq_mbar_init(bar_full + 8 * s, 32);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    bq_cta_sync();
    bool rejoined = false;
    if (mover) {
      auto issue = [&](int s) {
        if (s < n_chunks) {
          const uint32_t st = stage0 + (uint32_t)(s % kShStages) * kShStageBytes;
          const bool own = s >= n_warm;  // the output coefficients are not needed while warming up
          const int64_t off = lo + (int64_t)(c_first + s) * kChunk;
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const float* src = xrow[k] + off;
            const uint32_t dst = st + (uint32_t)((oct + 4 * k) * kShRowBytes + 16 + t8 * 16);
#pragma unroll
            for (int m = 0; m < kChunk / 32; m++) bq_cp16(dst + m * 128, src + m * 32);
          }
          const unsigned char* cc = cls + (size_t)(c_first + s) * kBqChunkBytes;
          const int n16 = (own ? kBqChunkBytes : kBqChunkABytes) / 16;
          for (int q = lane; q < n16; q += 32) bq_cp16(st + kShXBytes + (uint32_t)q * 16u, cc + (size_t)q * 16);
          if (REPAIR && rejoin && own)  // the speculative pass's slab states of this chunk: (kChunk / 32) * 256 bytes
            for (int q = lane; q < (kChunk / 32) * 16; q += 32)
              bq_cp16(st + kShXBytes + kBqChunkBytes + (uint32_t)q * 16u,
                      reinterpret_cast<const unsigned char*>(slab_group + (size_t)(c_first + s) * (kChunk / 32) * 32) + (size_t)q * 16);
          bq_cp_arrive(bar_full + 8 * (s % kShStages));
        }
      };
      for (int s = 0; s < kShStages - 1; s++) issue(s);
      for (int s = 0; s < n_chunks && !rejoined; s++) {
        const int sti = s % kShStages;
        bq_mbar_wait(bar_done + 8 * sti, (uint32_t)((s / kShStages) & 1));  // the chain warp has written w over this stage's samples
        const int c = ctrl[sti];
        const int frames_done = c & 0xffff;
        rejoined = (c >> 16) != 0;
        // the next request goes out BEFORE this chunk's output sums: into the stage of chunk s - 1, whose sums are done.  (Behind
        // them, a chunk's samples had one chunk of chain time to arrive and the chain warp spent half its time waiting.)
        if (!rejoined) issue(s + kShStages - 1);
        if (s >= n_warm && frames_done > 0) {
          const unsigned char* stp = sh_smem + (size_t)sti * kShStageBytes;
          const float4* __restrict__ Bc = reinterpret_cast<const float4*>(stp + kShXBytes + kBqChunkABytes) + (oct & 1) * (kChunk + 1);
          const int64_t base = lo + (int64_t)(c_first + s) * kChunk;
          // this lane's frames 32m + 4t .. 32m + 4t + 3 (m = 0, 1) belong to channel oct & 1 in every one of its rows: 8 coefficient sets
          float4 b[kChunk / 8];
#pragma unroll
          for (int m = 0; m < kChunk / 32; m++)
#pragma unroll
            for (int e = 0; e < 4; e++) b[4 * m + e] = Bc[32 * m + 4 * t8 + e];
          // (a row's sums are independent of each other: all of a row's loads first, then the arithmetic, then the stores — with
          // per-sum conditions in the loop the compiler keeps every LDS -> FMUL -> FADD -> FADD -> ST sequence to itself, ~100
          // cycles each, and the mover warp became the bottleneck of the pair)
          const bool whole = frames_done == kChunk;
#pragma unroll
          for (int k = 0; k < 8; k++) {
            if (oct + 4 * k >= n_rows) continue;
            const unsigned char* row = stp + (oct + 4 * k) * kShRowBytes + 16 + t8 * 16;
            float* dst = yrow[k] + base;
            float4 wc[kChunk / 32], y[kChunk / 32];
            float2 wp[kChunk / 32];
#pragma unroll
            for (int m = 0; m < kChunk / 32; m++) {
              wc[m] = *reinterpret_cast<const float4*>(row + m * 128);      // w[n .. n + 3]
              wp[m] = *reinterpret_cast<const float2*>(row + m * 128 - 8);  // w[n - 2], w[n - 1]
            }
#pragma unroll
            for (int m = 0; m < kChunk / 32; m++) {  // :138  y = b0*w + b1*w1 + b2*w2
              y[m].x = b[4 * m].x * wc[m].x + b[4 * m].y * wp[m].y + b[4 * m].z * wp[m].x;
              y[m].y = b[4 * m + 1].x * wc[m].y + b[4 * m + 1].y * wc[m].x + b[4 * m + 1].z * wp[m].y;
              y[m].z = b[4 * m + 2].x * wc[m].z + b[4 * m + 2].y * wc[m].y + b[4 * m + 2].z * wc[m].x;
              y[m].w = b[4 * m + 3].x * wc[m].w + b[4 * m + 3].y * wc[m].z + b[4 * m + 3].z * wc[m].y;
            }
#pragma unroll
            for (int m = 0; m < kChunk / 32; m++)
              if (whole || 32 * m < frames_done) *reinterpret_cast<float4*>(dst + m * 32) = y[m];
          }
        }
        __syncwarp();
        if (rejoined)  // the chunks already requested but never consumed: their arrivals must be in before the barriers are initialised again
          for (int o = s + 1; o < n_chunks && o < s + kShStages - 1; o++) bq_mbar_wait(bar_full + 8 * (o % kShStages), (uint32_t)((o / kShStages) & 1));
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
      for (int s = 0; s < n_chunks && !rejoined; s++) {
        if (s == n_warm && !REPAIR) st_group[((size_t)blockIdx.y * 2 + 0) * 32 + lane] = make_float2(w1, w2);  // state at the segment's first own frame
        const int sti = s % kShStages;
        bq_mbar_wait(bar_full + 8 * sti, (uint32_t)((s / kShStages) & 1));
        unsigned char* stp = sh_smem + (size_t)sti * kShStageBytes;
        float4* __restrict__ xs = reinterpret_cast<float4*>(stp + (size_t)lane * kShRowBytes + 16);
        const float4* __restrict__ A4 = reinterpret_cast<const float4*>(stp + kShXBytes + ch * (kBqChunkABytes / 2));  // two frames per element
        const float2* __restrict__ specs = reinterpret_cast<const float2*>(stp + kShXBytes + kBqChunkBytes);
        const bool own = s >= n_warm;
        // the two samples of history the mover's output sums need in front of this chunk
        *reinterpret_cast<float2*>(stp + (size_t)lane * kShRowBytes + 8) = make_float2(w2, w1);
        constexpr int NS = kChunk / 16;
        auto load16 = [&](float (&vx)[16], float (&v1)[16], float (&v2)[16], int step) {
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const float4 v = xs[step * 4 + q];
            vx[4 * q] = v.x;
            vx[4 * q + 1] = v.y;
            vx[4 * q + 2] = v.z;
            vx[4 * q + 3] = v.w;
          }
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const float4 a = A4[step * 8 + q];  // (a1, a2) of two consecutive frames
            v1[2 * q] = a.x;
            v2[2 * q] = a.y;
            v1[2 * q + 1] = a.z;
            v2[2 * q + 1] = a.w;
          }
        };
        float cx[16], c1[16], c2[16];
        load16(cx, c1, c2, 0);
        int done_steps = NS;
#pragma unroll 1
        for (int step = 0; step < NS; step++) {
          float nx[16], n1[16], n2[16];
          load16(nx, n1, n2, step + 1 < NS ? step + 1 : step);  // consumed by the NEXT trip (the last trip's load is a dummy)
          float wv[16];
#pragma unroll
          for (int i = 0; i < 16; i++) {
            const float w = cx[i] - c1[i] * w1 - c2[i] * w2;  // :137
            w2 = w1;
            w1 = w;
            wv[i] = w;
          }
          if (own) {
#pragma unroll
            for (int q = 0; q < 4; q++) xs[step * 4 + q] = make_float4(wv[4 * q], wv[4 * q + 1], wv[4 * q + 2], wv[4 * q + 3]);
          }
#pragma unroll
          for (int i = 0; i < 16; i++) {
            cx[i] = nx[i];
            c1[i] = n1[i];
            c2[i] = n2[i];
          }
          if (own && (step & 1)) {  // a 32-frame slab is complete: its end state is recorded / compared
            float2* rec = slab_group + ((size_t)(c_first + s) * (kChunk / 32) + (step >> 1)) * 32 + lane;
            if (REPAIR && rejoin) {
              const float2 spec = specs[(step >> 1) * 32 + lane];
              const bool same = __float_as_uint(spec.x) == __float_as_uint(w1) && __float_as_uint(spec.y) == __float_as_uint(w2);
              if (__all_sync(0xffffffffu, same)) {
                // re-joined the speculative trajectory: everything behind this slab is already right
                rejoined = true;
                done_steps = step + 1;
                break;
              }
            }
            *rec = make_float2(w1, w2);
          }
        }
        if (lane == 0) ctrl[sti] = (own ? done_steps * 16 : 0) | (rejoined ? 1 << 16 : 0);
        __syncwarp();
        bq_mbar_arrive(bar_done + 8 * sti);  // (release: w and the control word are visible to the mover behind its wait)
      }
    }
    // the mover learns about a re-join from the control word; both leave the run together
    // both warps leave the run together and with the same answer (each learnt of a re-join on its own: the chain warp found it, the
    // mover read it from the chunk's control word)
    bq_cta_sync();
    return rejoined;
  };

  if (!REPAIR) {
    const int seg = (int)blockIdx.y;
    const int c_own = seg * seg_chunks;
    if (c_own >= total_chunks) return;
    const int c_end = c_own + seg_chunks < total_chunks ? c_own + seg_chunks : total_chunks;
    const int c_first = seg == 0 ? c_own : (c_own - warm_chunks > 0 ? c_own - warm_chunks : 0);
    run(c_first, c_own - c_first, c_end, false);
    if (!mover) st_group[((size_t)seg * 2 + 1) * 32 + lane] = make_float2(w1, w2);
  } else {
    bool carry = false;
    for (int seg = first_bad[g]; seg < n_seg; seg++) {
      const int c_own = seg * seg_chunks;
      if (c_own >= total_chunks) break;
      if (!carry && !link_bad[(size_t)g * n_seg + seg]) continue;
      if (!carry && !mover) {  // the last true state: what segment seg-1 ended with
        const float2 e = st_group[((size_t)(seg - 1) * 2 + 1) * 32 + lane];
        w1 = e.x;
        w2 = e.y;
      }
      const int c_end = c_own + seg_chunks < total_chunks ? c_own + seg_chunks : total_chunks;
      carry = !run(c_own, 0, c_end, true);
    }
  }
}

// first_bad[g] = first segment k >= 1 whose speculative start state differs (bitwise, any of the 32 rows) from the state
// segment k-1 ended with (n_seg if the whole chain verifies); link_bad[g][k] says so for every link.  One CTA per group, the links
// dealt out to its warps.
constexpr int kVerifyWarps = 8;
__global__ void __launch_bounds__(kVerifyWarps * 32) k_biquad_verify(int n_seg, const float2* __restrict__ states, int* __restrict__ first_bad,
                                                                     int* __restrict__ link_bad) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint2* st = reinterpret_cast<const uint2*>(states) + (size_t)blockIdx.x * n_seg * 64;
  __shared__ int s_bad[kVerifyWarps];
  int bad = n_seg;
  if (threadIdx.x == 0) link_bad[(size_t)blockIdx.x * n_seg] = 0;
  for (int k = 1 + warp; k < n_seg; k += kVerifyWarps) {
    const uint2 e = st[((size_t)(k - 1) * 2 + 1) * 32 + lane];
    const uint2 b = st[((size_t)k * 2 + 0) * 32 + lane];
    const bool differ = __any_sync(0xffffffffu, e.x != b.x || e.y != b.y);
    if (differ && bad == n_seg) bad = k;  // (k ascends inside a warp: the first one found is the warp's smallest)
    if (lane == 0) link_bad[(size_t)blockIdx.x * n_seg + k] = differ ? 1 : 0;
  }
  if (lane == 0) s_bad[warp] = bad;
  __syncthreads();
  if (threadIdx.x == 0) {
    int m = n_seg;
    for (int w = 0; w < kVerifyWarps; w++) m = s_bad[w] < m ? s_bad[w] : m;
    first_bad[blockIdx.x] = m;
  }
}

// warm-up of a speculative segment in 32-frame slabs (GAC_BIQUAD_WARM_SLABS overrides the default for measurements)
static int warm_slabs() {
  static const int v = [] {
    const char* e = getenv("GAC_BIQUAD_WARM_SLABS");
    const int x = e ? atoi(e) : kWarmSlabs;
    return x >= 8 && x <= 65536 ? x : kWarmSlabs;
  }();
  return v;
}

int biquad_lane_segments(int n_jobs, int64_t n_frames, int* seg_slabs_out) {
  // one CTA = one warp with 96 KB of staging, two per SM: as many concurrent segments as there are slots per group, but
  // segments no shorter than the warm-up (below that the redundant work outweighs the concurrency)
  const int groups = (n_jobs + 15) / 16;
  const int total_slabs = (int)((n_frames + kSlab - 1) / kSlab);
  int n_seg = kLaneSlots / (groups > 0 ? groups : 1);
  if (n_seg < 1) n_seg = 1;
  int seg_slabs = (total_slabs + n_seg - 1) / n_seg;
  if (seg_slabs < warm_slabs()) seg_slabs = warm_slabs();
  n_seg = (total_slabs + seg_slabs - 1) / seg_slabs;
  if (n_seg < 1) n_seg = 1;
  if (seg_slabs_out) *seg_slabs_out = seg_slabs;
  return n_seg;
}

// scratch sizes for launch_biquad: float2 elements (segment states + per-slab states) and ints (flags)
void biquad_scratch_sizes(int n_jobs, int64_t n_frames, size_t* n_float2, size_t* n_int) {
  const size_t groups = (size_t)((n_jobs + 15) / 16);
  const int n_seg = biquad_lane_segments(n_jobs, n_frames, nullptr);
  *n_float2 = groups * (size_t)n_seg * 64 + groups * (size_t)(n_frames / kSlab + 1) * 32;
  *n_int = groups * 2 + groups * (size_t)n_seg;
}

static int warm_slabs();
// warm-up of the shared path in chunks: the caller's hint (32-frame slabs; 0 = none) unless GAC_BIQUAD_WARM_SLABS forces a value
static int shared_warm_chunks(int warm_hint_slabs) {
  const int slabs = (warm_hint_slabs > 0 && !getenv("GAC_BIQUAD_WARM_SLABS")) ? warm_hint_slabs : warm_slabs();
  return std::max(2, slabs / (kChunk / 32));
}
int biquad_shared_segments(int n_groups, int64_t n_frames, int* seg_chunks_out, int warm_hint_slabs) {
  const int groups = n_groups;
  const int total_chunks = (int)((n_frames + kChunk - 1) / kChunk);
  const int warm = shared_warm_chunks(warm_hint_slabs);
  int n_seg = kShSlots / (groups > 0 ? groups : 1);
  if (n_seg < 1) n_seg = 1;
  int seg_chunks = (total_chunks + n_seg - 1) / n_seg;
  if (seg_chunks < warm) seg_chunks = warm;
  n_seg = (total_chunks + seg_chunks - 1) / seg_chunks;
  if (n_seg < 1) n_seg = 1;
  if (seg_chunks_out) *seg_chunks_out = seg_chunks;
  return n_seg;
}
void biquad_shared_scratch_sizes(int n_groups, int64_t n_frames, size_t* n_float2, size_t* n_int, int warm_hint_slabs) {
  const size_t groups = (size_t)n_groups;
  const int n_seg = biquad_shared_segments(n_groups, n_frames, nullptr, warm_hint_slabs);
  *n_float2 = groups * (size_t)n_seg * 64 + groups * (size_t)(n_frames / kSlab + 4) * 32;
  *n_int = groups + groups * (size_t)n_seg;
}
void launch_biquad_lanes_shared(const BiquadJob* d_jobs, const BqGroup* d_groups, int n_groups, const unsigned char* d_cs, size_t cs_stride,
                                int64_t n_frames, float2* d_states, int* d_flags, bool sequential, cudaStream_t s, int warm_hint_slabs) {
  if (n_groups <= 0) return;
  if (sequential) {
    // filters that forget too slowly for speculative segments to re-join (constant low cutoff / high Q): one segment per group from
    // the true initial state — exact by construction, nothing to verify or repair, and no speculative pass thrown away
    GAC_SMEM_OPT_IN((k_biquad_lanes_shared<false>), kShSmem);
    const int total_chunks = (int)((n_frames + kChunk - 1) / kChunk);
    k_biquad_lanes_shared<false><<<dim3((unsigned)n_groups, 1), 64, kShSmem, s>>>(d_jobs, d_groups, d_cs, cs_stride, n_frames, total_chunks, 1, 0, d_states,
                                                                              d_states + (size_t)n_groups * 64, nullptr, nullptr);
    return;
  }
  GAC_SMEM_OPT_IN((k_biquad_lanes_shared<false>), kShSmem);
  GAC_SMEM_OPT_IN((k_biquad_lanes_shared<true>), kShSmem);
  const unsigned groups = (unsigned)n_groups;
  int seg_chunks = 0;
  const int n_seg = biquad_shared_segments(n_groups, n_frames, &seg_chunks, warm_hint_slabs);
  const int warm = shared_warm_chunks(warm_hint_slabs);
  float2* d_slab = d_states + (size_t)groups * n_seg * 64;
  int* d_first_bad = d_flags;
  int* d_link = d_flags + groups;
  k_biquad_lanes_shared<false><<<dim3(groups, (unsigned)n_seg), 64, kShSmem, s>>>(d_jobs, d_groups, d_cs, cs_stride, n_frames, seg_chunks, n_seg,
                                                                                 warm, d_states, d_slab, nullptr, nullptr);
  if (n_seg > 1) {
    k_biquad_verify<<<groups, kVerifyWarps * 32, 0, s>>>(n_seg, d_states, d_first_bad, d_link);
    k_biquad_lanes_shared<true><<<groups, 64, kShSmem, s>>>(d_jobs, d_groups, d_cs, cs_stride, n_frames, seg_chunks, n_seg, warm, d_states, d_slab,
                                                            d_first_bad, d_link);
  }
}

void launch_biquad_lanes(const BiquadJob* d_jobs, int n_jobs, int64_t n_frames, const float4* d_s1t, const float4* d_s2t, float2* d_states,
                         int* d_flags, cudaStream_t s) {
  GAC_SMEM_OPT_IN((k_biquad_lanes<false, false>), kLanesSmem);
  GAC_SMEM_OPT_IN((k_biquad_lanes<false, true>), kLanesSmem);
  GAC_SMEM_OPT_IN((k_biquad_lanes<true, false>), kLanesSmem);
  GAC_SMEM_OPT_IN((k_biquad_lanes<true, true>), kLanesSmem);
  const unsigned groups = (unsigned)((n_jobs + 15) / 16);
  int seg_slabs = 0;
  const int n_seg = biquad_lane_segments(n_jobs, n_frames, &seg_slabs);
  float2* d_slab = d_states + (size_t)groups * n_seg * 64;
  int* d_first_bad = d_flags;
  const int* d_wide = d_flags + groups;
  int* d_link = d_flags + 2 * groups;
  // both stream layouts are launched; a CTA whose group uses the other layout exits at once
  k_biquad_lanes<false, false><<<dim3(groups, (unsigned)n_seg), 32, kLanesSmem, s>>>(d_jobs, n_jobs, n_frames, d_s1t, d_s2t, seg_slabs, n_seg, warm_slabs(), d_states, d_slab, nullptr, d_wide, nullptr);
  k_biquad_lanes<false, true><<<dim3(groups, (unsigned)n_seg), 32, kLanesSmem, s>>>(d_jobs, n_jobs, n_frames, d_s1t, d_s2t, seg_slabs, n_seg, warm_slabs(), d_states, d_slab, nullptr, d_wide, nullptr);
  if (n_seg > 1) {
    k_biquad_verify<<<groups, kVerifyWarps * 32, 0, s>>>(n_seg, d_states, d_first_bad, d_link);
    k_biquad_lanes<true, false><<<groups, 32, kLanesSmem, s>>>(d_jobs, n_jobs, n_frames, d_s1t, d_s2t, seg_slabs, n_seg, warm_slabs(), d_states, d_slab, d_first_bad, d_wide, d_link);
    k_biquad_lanes<true, true><<<groups, 32, kLanesSmem, s>>>(d_jobs, n_jobs, n_frames, d_s1t, d_s2t, seg_slabs, n_seg, warm_slabs(), d_states, d_slab, d_first_bad, d_wide, d_link);
  }
}

}  // namespace gac
