// biquad.cu — K3: BiQuadFilterNode (Nodes/BiQuadFilterNode.cs:87-258), batched over voices.
// Compiled with --fmad=false: the RBJ formulas and the Direct-Form-II recursion round like the reference's scalar float32 code.
//
// The reference recomputes coefficients inside the per-sample loop (sin, cos, five divisions) whenever the a-rate
// frequency/Q moved beyond a hysteresis (:126).  Here that is split so that only the unavoidable recursion is sequential:
//   k_biquad_select   parallel over quanta   the hysteresis walk (usedFreq restarts at 1000 Hz in EVERY block, :13-14,111-112;
//                                            channel 1 inherits channel 0's end state): for every (channel, frame) the frame
//                                            whose (f, Q) produced the coefficients in force, or -1 = "what the block started with"
//   k_biquad_entry    one CTA per voice      which coefficient set the reference's fields hold when a quantum starts
//   k_biquad_resolve  parallel over frames   RBJ (glibc-exact sinf/cosf, :149-258) of the frame in force; writes (b0, b1, b2) per
//                                            (channel, frame) and the slab-transposed stream (x, a1, a2) the lanes read
//   k_biquad_lanes    one lane per (voice, channel): w = x - a1*w1 - a2*w2 (:137) is the sequential chain, y = b0*w + b1*w1 +
//                     b2*w2 (:138) rides along off the critical path; TMA-fed, concurrent verified time segments (biquad_lanes.cu)
//
// Slab-transposed layout (shared by resolve and lanes): rows are grouped 32 at a time (16 voices x 2 channels = one
// warp of lanes), time is cut into slabs of 32 frames, and inside a slab the FRAME index is major and the row index minor
// (per-row layout, below) — or, when the select pass proved that both channels of every voice of the group share one
// coefficient set per frame (always, except for drifts slower than the hysteresis), the VOICE index is minor and an element
// carries both channels' samples: [frame][voice 0..15] (a1, a2, xL, xR) / (b0, b1, b2, -), half the bytes:
//     S1T[group][slab][frame 0..31][row 0..31]  float4 (x, a1, a2, -)         S2T[group][slab][frame][row]  float4 (b0, b1, b2, -)
// so that a slab is one contiguous 16 KB (4 KB) block for TMA and lane r's LDS.128 / STS.32 at frame i are conflict-free.
#include "biquad_math.cuh"

namespace gac {

// ---------------------------------------------------------------------------------------------------------------------
// K3a: the hysteresis walk (:121-134).  One WARP per quantum of one job: the lanes load the quantum's 128 clamped (f, Q)
// pairs (float4 each, coalesced) and decide the two cases that cover practically every quantum without walking it:
//   * every frame differs from its predecessor beyond the hysteresis (an a-rate sweep): every frame recomputes, provided
//     frame 0 does (channel 0: against the 1000 Hz / Q 1 the block restarts with, or the dirty flag; channel 1: against
//     channel 0's last frame);
//   * every frame equals frame 0 (constant parameters): at most frame 0 of channel 0 recomputes, channel 1 never does.
// Anything else (drifts slower than the hysteresis) is walked serially by lane 0 from the staged values, exactly as the
// reference loop does.
//   idx[c][n]  = frame whose (f, Q) produced the coefficients in force at (channel c, frame n); -1 = the set the quantum
//                started with (resolved by k_biquad_entry)
//   last[c][b] = frame of channel c's last recompute in quantum b, or -1
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kSelWarps = 8;
__device__ __forceinline__ bool bq_differs(float f, float q, float uf, float uq) {
  return fabsf(f - uf) > 0.001f || fabsf(q - uq) > 0.0001f;  // :126 (gain never differs inside a block)
}
__global__ void __launch_bounds__(kSelWarps * 32) k_biquad_select(const BiquadJob* __restrict__ jobs, int sample_rate, int64_t n_quanta,
                                                                  int64_t n_frames, int32_t* __restrict__ last_base, int* __restrict__ wide) {
  __shared__ float sf[kSelWarps][128];
  __shared__ float sq[kSelWarps][128];
  __shared__ int32_t si[kSelWarps][2][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int jid = blockIdx.y;
  const int64_t b = (int64_t)blockIdx.x * kSelWarps + warp;
  if (b >= n_quanta) return;
  const BiquadJob job = jobs[jid];
  int32_t* last = last_base + (size_t)jid * 2 * n_quanta;
  const int64_t n0 = b * 128;
  if (!(n0 >= job.lo && n0 < job.hi)) {  // silent-flagged block: state untouched (:103-108)
    if (lane == 0) last[b] = last[n_quanta + b] = -1;
    return;
  }
  const float nyq = (float)sample_rate / 2.f;
  float f[4], q[4];
  {
    float4 f4 = make_float4(job.freq_const, job.freq_const, job.freq_const, job.freq_const);
    float4 q4 = make_float4(job.q_const, job.q_const, job.q_const, job.q_const);
    if (job.freq) f4 = *reinterpret_cast<const float4*>(job.freq + n0 + 4 * lane);
    if (job.q) q4 = *reinterpret_cast<const float4*>(job.q + n0 + 4 * lane);
    const float fe[4] = {f4.x, f4.y, f4.z, f4.w}, qe[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
    for (int e = 0; e < 4; e++) {
      f[e] = fe[e] < 1.f ? 1.f : (fe[e] > nyq ? nyq : fe[e]);  // Math.Clamp(freq, 1, fs/2)   :123
      q[e] = qe[e] > 0.001f ? qe[e] : 0.001f;                   // Math.Max(0.001f, q)         :124
    }
  }
  // predecessor of this lane's first frame, frame 0 and frame 127 of the quantum
  const float pf = __shfl_up_sync(0xffffffffu, f[3], 1), pq = __shfl_up_sync(0xffffffffu, q[3], 1);
  const float f0 = __shfl_sync(0xffffffffu, f[0], 0), q0 = __shfl_sync(0xffffffffu, q[0], 0);
  const float fl = __shfl_sync(0xffffffffu, f[3], 31), ql = __shfl_sync(0xffffffffu, q[3], 31);
  bool all_diff = true, all_same = true;
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const float uf = e == 0 ? pf : f[e - 1], uq = e == 0 ? pq : q[e - 1];
    if (!(lane == 0 && e == 0)) all_diff = all_diff && bq_differs(f[e], q[e], uf, uq);
    all_same = all_same && f[e] == f0 && q[e] == q0;
  }
  all_diff = __all_sync(0xffffffffu, all_diff);
  all_same = __all_sync(0xffffffffu, all_same);
  const bool dirty0 = (n0 == job.lo);  // _coefficientsDirty is still set when the first non-silent block arrives
  const bool r0 = dirty0 || bq_differs(f0, q0, 1000.f, 1.0f);
  int32_t* idx0 = job.idx + n0 + 4 * lane;
  int32_t* idx1 = job.idx + n_frames + n0 + 4 * lane;
  if (all_same) {
    const int32_t k = r0 ? (int32_t)n0 : -1;
    *reinterpret_cast<int4*>(idx0) = make_int4(k, k, k, k);
    *reinterpret_cast<int4*>(idx1) = make_int4(-1, -1, -1, -1);
    if (lane == 0) {
      last[b] = k;
      last[n_quanta + b] = -1;
    }
    return;
  }
  const bool r1 = bq_differs(f0, q0, fl, ql);  // channel 1 starts from channel 0's end state (:110-115)
  if (all_diff && r0 && r1) {
    const int32_t k = (int32_t)(n0 + 4 * lane);
    *reinterpret_cast<int4*>(idx0) = make_int4(k, k + 1, k + 2, k + 3);
    *reinterpret_cast<int4*>(idx1) = make_int4(k, k + 1, k + 2, k + 3);
    if (lane == 0) last[b] = last[n_quanta + b] = (int32_t)(n0 + 127);
    return;
  }
  // general case: the reference's walk, channel 0 then channel 1, by one lane over the staged values
#pragma unroll
  for (int e = 0; e < 4; e++) {
    sf[warp][4 * lane + e] = f[e];
    sq[warp][4 * lane + e] = q[e];
  }
  __syncwarp();
  if (lane == 0) {
    float usedF = 1000.f, usedQ = 1.0f;
    bool dirty = dirty0;
    for (int c = 0; c < 2; c++) {
      int32_t cur = -1;
      for (int i = 0; i < 128; i++) {
        const float fi = sf[warp][i], qi = sq[warp][i];
        if (dirty || bq_differs(fi, qi, usedF, usedQ)) {
          usedF = fi;
          usedQ = qi;
          dirty = false;
          cur = (int32_t)(n0 + i);
        }
        si[warp][c][i] = cur;
      }
      last[(size_t)c * n_quanta + b] = cur;
    }
    // the two channels of this voice may now be governed by different coefficient sets somewhere in the quantum: its 32-row
    // group takes the per-row stream layout (both closed-form cases above provably share one set per frame)
    // (frames before a channel's first recompute use the entry set: channel 0's is what the previous quantum left, channel 1's
    // is channel 0's last recompute of THIS quantum if there was one — k_biquad_entry — so "-1 in both" is not always equal)
    const int32_t l0 = last[b];
    bool differ = false;
    for (int i = 0; i < 128; i++) {
      const int32_t a = si[warp][0][i], c = si[warp][1][i];
      differ = differ || a != c || (c < 0 && l0 >= 0);
    }
    if (differ) atomicOr(&wide[jid >> 4], 1);
  }
  __syncwarp();
  *reinterpret_cast<int4*>(idx0) = *reinterpret_cast<const int4*>(&si[warp][0][4 * lane]);
  *reinterpret_cast<int4*>(idx1) = *reinterpret_cast<const int4*>(&si[warp][1][4 * lane]);
}

// K3b: which coefficient set the reference's fields (_b0.._a2) hold when channel c of quantum b starts.
// ent[c][b] = frame index whose (f, Q) applies, or -1 (only before the very first recompute).  A "last recompute so far"
// scan over the quanta: one CTA per job, every thread owns a short contiguous run of quanta (run summary, exclusive scan of the
// summaries over the CTA with the operator "the right-most one that recomputed at all", replay).
constexpr int kEntThreads = 512;
__global__ void __launch_bounds__(kEntThreads) k_biquad_entry(int n_jobs, int64_t n_quanta, const int32_t* __restrict__ last_base,
                                                              int32_t* __restrict__ ent_base) {
  const int jid = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (jid >= n_jobs) return;
  const int32_t* last0 = last_base + (size_t)jid * 2 * n_quanta;
  const int32_t* last1 = last0 + n_quanta;
  int32_t* ent0 = ent_base + (size_t)jid * 2 * n_quanta;
  int32_t* ent1 = ent0 + n_quanta;
  const int64_t chunk = (n_quanta + kEntThreads - 1) / kEntThreads;
  const int64_t b_lo = tid * chunk < n_quanta ? tid * chunk : n_quanta, b_hi = (b_lo + chunk < n_quanta) ? b_lo + chunk : n_quanta;
  // summary of the run: the last recompute inside it (channel 1's wins over channel 0's within a quantum), or -2 = none
  int32_t tail = -2;
  for (int64_t b = b_lo; b < b_hi; b++) {
    const int32_t l0 = last0[b], l1 = last1[b];
    if (l0 >= 0) tail = l0;
    if (l1 >= 0) tail = l1;
  }
  // inclusive scan inside the warp, then across the warps
  int32_t inc = tail;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d && inc == -2) inc = t;
  }
  __shared__ int32_t warp_tail[kEntThreads / 32];
  if (lane == 31) warp_tail[warp] = inc;
  __syncthreads();
  // carry into this thread = summary of the nearest lower thread that recomputed at all, else -1
  int32_t cur = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) cur = -2;
  if (cur == -2) {
    for (int w = warp - 1; w >= 0; w--)
      if (warp_tail[w] != -2) {
        cur = warp_tail[w];
        break;
      }
  }
  if (cur == -2) cur = -1;
  for (int64_t b = b_lo; b < b_hi; b++) {
    ent0[b] = cur;                      // channel 0 starts from the fields as the previous block left them
    const int32_t l0 = last0[b];
    const int32_t c0 = l0 >= 0 ? l0 : cur;
    ent1[b] = c0;                       // channel 1 starts from channel 0's end state (:110-115)
    const int32_t l1 = last1[b];
    cur = l1 >= 0 ? l1 : c0;
  }
}

// K3c: coefficients in force at every (channel, frame): RBJ of the frame the walk selected, with the k-rate gain of that
// frame's block.  CTA = one slab of one 32-row group, 16 warps: warp = voice, lane = frame, and the thread serves BOTH
// channels: they almost always selected the same frame, so one RBJ evaluation (glibc-exact sinf/cosf, five divisions) feeds
// both rows; and lanes that selected the SAME frame (constant parameters: one recompute per quantum) elect one of them to
// evaluate it (__match_any_sync) and take its result by shuffle.  Reads run along time (coalesced); the (x, a1, a2) and
// (b0, b1, b2) tiles are transposed through shared memory and written as contiguous 16 KB blocks; frames outside the
// voice's non-silent range are cleared here (:103-108) because the recursion kernel only writes inside it.
__device__ __forceinline__ Coef rbj_shared(const BiquadJob& job, int32_t k, float nyq, int sample_rate) {
  // one evaluation per distinct k in the warp
  const unsigned peers = __match_any_sync(0xffffffffu, k);
  const int leader = __ffs(peers) - 1;
  Coef c;
  c.b0 = c.b1 = c.b2 = c.a1 = c.a2 = 0.f;  // k < 0 is unreachable for active frames: the first one always recomputes (dirty)
  if ((int)(threadIdx.x & 31) == leader && k >= 0)
    c = rbj(job.type, clamped_freq(job, k, nyq), clamped_q(job, k), job.gain ? job.gain[k >> 7] : job.gain_const, sample_rate);
  c.b0 = __shfl_sync(0xffffffffu, c.b0, leader);
  c.b1 = __shfl_sync(0xffffffffu, c.b1, leader);
  c.b2 = __shfl_sync(0xffffffffu, c.b2, leader);
  c.a1 = __shfl_sync(0xffffffffu, c.a1, leader);
  c.a2 = __shfl_sync(0xffffffffu, c.a2, leader);
  return c;
}

constexpr int kResSlabs = 16;  // 32-frame slabs per CTA of k_biquad_resolve
__global__ void __launch_bounds__(512, 2) k_biquad_resolve(const BiquadJob* __restrict__ jobs, int n_jobs, int sample_rate, int64_t n_quanta,
                                                        int64_t n_frames, const int32_t* __restrict__ ent_base, float4* __restrict__ s1t,
                                                        float4* __restrict__ s2t, const int* __restrict__ wide_flags) {
  __shared__ float4 tile[32][33];
  __shared__ float4 tile2[32][33];
  // Voices of a group very often run the SAME filter automation (one table, deduplicated by the engine, or the same constants):
  // then the coefficient set of (frame, selected frame k) is the same for all of them.  Phase 1: warp w evaluates RBJ for the
  // group's FIRST voice over slab w of the CTA's 16 slabs (all 16 warps busy with the expensive part at once); phase 2: slab by
  // slab, warp = voice copies the first voice's set when its parameters and its selected frame coincide, and evaluates its own
  // otherwise.
  __shared__ int32_t lead_k[kResSlabs][32];
  __shared__ float lead_c[kResSlabs][5][32];
  const int w = threadIdx.x >> 5, i = threadIdx.x & 31;
  const int g = blockIdx.y;
  const int64_t slab0 = (int64_t)blockIdx.x * kResSlabs;
  const int64_t n_slabs = n_frames / 32;
  const float nyq = (float)sample_rate / 2.f;
  {
    const int64_t n = (slab0 + w) * 32 + i;
    const BiquadJob& lead = jobs[g * 16];
    int32_t k0 = INT32_MIN;
    Coef c;
    c.b0 = c.b1 = c.b2 = c.a1 = c.a2 = 0.f;
    if (slab0 + w < n_slabs && n >= lead.lo && n < lead.hi) {  // warp-uniform: ranges are multiples of 128 frames
      k0 = lead.idx[n];
      if (k0 < 0) k0 = ent_base[((size_t)(g * 16) * 2 + 0) * n_quanta + (n >> 7)];
      c = rbj_shared(lead, k0, nyq, sample_rate);
    }
    lead_k[w][i] = k0;
    lead_c[w][0][i] = c.b0;
    lead_c[w][1][i] = c.b1;
    lead_c[w][2][i] = c.b2;
    lead_c[w][3][i] = c.a1;
    lead_c[w][4][i] = c.a2;
  }
  __syncthreads();
  const int jv = w;
  const int jid = g * 16 + jv;
  const bool wide = wide_flags[g] != 0;
  const size_t group_base = (size_t)g * (size_t)n_frames * 32;
  bool same_params = false;
  if (jid < n_jobs) {
    const BiquadJob& job = jobs[jid];
    const BiquadJob& lead = jobs[g * 16];
    same_params = job.type == lead.type && job.freq == lead.freq && job.q == lead.q && job.gain == lead.gain &&
                  job.freq_const == lead.freq_const && job.q_const == lead.q_const && job.gain_const == lead.gain_const;
  }
  // The selected frames and the two input samples of slab sl + 1 are requested before slab sl is worked on: the kernel was 60 %
  // long-scoreboard (every warp waited for these loads, then for the CTA's barriers) at 35 % of the DRAM peak.
  const bool have_job = jid < n_jobs;
  const int64_t my_lo = have_job ? jobs[jid].lo : 0, my_hi = have_job ? jobs[jid].hi : 0;
  const int32_t* __restrict__ my_idx = have_job ? jobs[jid].idx : nullptr;
  float* __restrict__ sig0 = have_job ? jobs[jid].sig[0] : nullptr;
  float* __restrict__ sig1 = have_job ? jobs[jid].sig[1] : nullptr;
  const float* __restrict__ in0 = have_job ? (jobs[jid].in[0] ? jobs[jid].in[0] : jobs[jid].sig[0]) : nullptr;
  const float* __restrict__ in1 = have_job ? (jobs[jid].in[1] ? jobs[jid].in[1] : jobs[jid].sig[1]) : nullptr;
  int32_t pk0 = 0, pk1 = 0;
  float px0 = 0.f, px1 = 0.f;
  auto prefetch = [&](int64_t slab) {
    const int64_t n = slab * 32 + i;
    if (slab < n_slabs && have_job && n >= my_lo && n < my_hi) {
      pk0 = my_idx[n];
      pk1 = my_idx[n_frames + n];
      px0 = in0[n];
      px1 = in1[n];
    }
  };
  prefetch(slab0);
  for (int sl = 0; sl < kResSlabs; sl++) {
    const int64_t slab = slab0 + sl;
    if (slab >= n_slabs) break;  // uniform for the CTA
    const int64_t n = slab * 32 + i;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0, u0 = v0, u1 = v0;
    const bool in_range = have_job && n >= my_lo && n < my_hi;  // warp-uniform
    int32_t k0 = pk0, k1 = pk1;
    const float x0 = px0, x1 = px1;
    if (sl + 1 < kResSlabs) prefetch(slab + 1);
    if (in_range) {
      const BiquadJob& job = jobs[jid];
      if (k0 < 0) k0 = ent_base[((size_t)jid * 2 + 0) * n_quanta + (n >> 7)];
      if (k1 < 0) k1 = ent_base[((size_t)jid * 2 + 1) * n_quanta + (n >> 7)];
      Coef c0;
      const bool reuse = same_params && k0 == lead_k[sl][i];
      if (__all_sync(0xffffffffu, reuse)) {
        c0.b0 = lead_c[sl][0][i];
        c0.b1 = lead_c[sl][1][i];
        c0.b2 = lead_c[sl][2][i];
        c0.a1 = lead_c[sl][3][i];
        c0.a2 = lead_c[sl][4][i];
      } else {
        c0 = rbj_shared(job, k0, nyq, sample_rate);
      }
      Coef c1 = c0;
      if (__any_sync(0xffffffffu, k1 != k0)) {
        const Coef alt = rbj_shared(job, k1, nyq, sample_rate);
        if (k1 != k0) c1 = alt;
      }
      v0 = make_float4(x0, c0.a1, c0.a2, 0.f);
      v1 = make_float4(x1, c1.a1, c1.a2, 0.f);
      u0 = make_float4(c0.b0, c0.b1, c0.b2, 0.f);
      u1 = make_float4(c1.b0, c1.b1, c1.b2, 0.f);
    } else if (have_job && n < n_frames) {
      sig0[n] = 0.f;
      sig1[n] = 0.f;
    }
    if (wide) {
      // per-row layout: tile[frame][row] -> S1T / S2T [g][slab][frame][row 0..31], element e = frame * 32 + row
      tile[i][2 * jv] = v0;
      tile[i][2 * jv + 1] = v1;
      tile2[i][2 * jv] = u0;
      tile2[i][2 * jv + 1] = u1;
      __syncthreads();
      float4* dst = s1t + group_base + (size_t)slab * 1024;
      float4* dst2 = s2t + group_base + (size_t)slab * 1024;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int e = threadIdx.x + 512 * h;
        dst[e] = tile[e >> 5][e & 31];
        dst2[e] = tile2[e >> 5][e & 31];
      }
    } else {
      // per-voice layout (both channels share the coefficient set at every frame of this group):
      // S1T [g][slab][frame][voice 0..15] = (a1, a2, xL, xR), S2T = (b0, b1, b2, -): half the bytes of the per-row layout
      tile[i][jv] = make_float4(v0.y, v0.z, v0.x, v1.x);
      tile2[i][jv] = u0;
      __syncthreads();
      const int e = threadIdx.x;  // element e = frame * 16 + voice
      s1t[group_base + (size_t)slab * 512 + e] = tile[e >> 4][e & 15];
      s2t[group_base + (size_t)slab * 512 + e] = tile2[e >> 4][e & 15];
    }
    __syncthreads();  // the tiles are rewritten by the next slab
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Shared-coefficient path.  Voices very often run the SAME filter: one type, one automation table (the engine evaluates
// identical automation once and hands every voice the same table), the same non-silent range.  The coefficients in force at
// (channel, frame) then depend on nothing voice-specific, so select / entry / RBJ run ONCE for a representative of the class
// and the recursion kernel (k_biquad_lanes_shared, biquad_lanes.cu) reads every voice's samples where they lie: the
// per-voice coefficient streams — 64 of the ~90 bytes the general path moves per voice and frame — are not written at all.
//
// Class stream, per chunk of kBqChunkFrames = 64 frames (one contiguous block a stage of the recursion kernel copies):
//   float2 A[2 channels][66]  (a1, a2)       float4 B[2 channels][65]  (b0, b1, b2, -)        (66 / 65: channel 1 sits four banks
//   behind channel 0, so that the two addresses a warp reads at one frame never collide; A is read two frames at a time)
__global__ void __launch_bounds__(kBqChunkFrames) k_biquad_class_coef(const BiquadJob* __restrict__ reps, int sample_rate, int64_t n_quanta,
                                                                      int64_t n_frames, const int32_t* __restrict__ ent_base,
                                                                      unsigned char* __restrict__ cs, size_t cs_stride) {
  const int cls = blockIdx.y;
  const int64_t chunk = blockIdx.x;
  const int i = threadIdx.x;
  const int64_t n = chunk * kBqChunkFrames + i;
  const BiquadJob& job = reps[cls];
  const float nyq = (float)sample_rate / 2.f;
  Coef c0, c1;
  c0.b0 = c0.b1 = c0.b2 = c0.a1 = c0.a2 = 0.f;
  c1 = c0;
  if (n >= job.lo && n < job.hi) {
    int32_t k0 = job.idx[n], k1 = job.idx[n_frames + n];
    if (k0 < 0) k0 = ent_base[((size_t)cls * 2 + 0) * n_quanta + (n >> 7)];
    if (k1 < 0) k1 = ent_base[((size_t)cls * 2 + 1) * n_quanta + (n >> 7)];
    if (k0 >= 0) c0 = rbj(job.type, clamped_freq(job, k0, nyq), clamped_q(job, k0), job.gain ? job.gain[k0 >> 7] : job.gain_const, sample_rate);
    c1 = c0;
    if (k1 != k0 && k1 >= 0) c1 = rbj(job.type, clamped_freq(job, k1, nyq), clamped_q(job, k1), job.gain ? job.gain[k1 >> 7] : job.gain_const, sample_rate);
  }
  unsigned char* blk = cs + (size_t)cls * cs_stride + (size_t)chunk * kBqChunkBytes;
  float2* A = reinterpret_cast<float2*>(blk);
  float4* B = reinterpret_cast<float4*>(blk + kBqChunkABytes);
  A[i] = make_float2(c0.a1, c0.a2);
  A[kBqChunkFrames + 2 + i] = make_float2(c1.a1, c1.a2);
  B[i] = make_float4(c0.b0, c0.b1, c0.b2, 0.f);
  B[kBqChunkFrames + 1 + i] = make_float4(c1.b0, c1.b1, c1.b2, 0.f);
}

// frames outside a voice's non-silent range: the node's output block is cleared there (:103-108); the recursion kernel only
// writes inside the range
__global__ void __launch_bounds__(256) k_biquad_zero_outside(const BiquadJob* __restrict__ jobs, int64_t n_frames) {
  const BiquadJob& job = jobs[blockIdx.y >> 1];
  float* __restrict__ row = job.sig[blockIdx.y & 1];
  // the CTAs of a row walk the frames OUTSIDE [lo, hi) only: [0, lo) then [hi, n_frames)  (ranges are multiples of 128 frames)
  const int64_t n_out = job.lo + (n_frames - job.hi);
  for (int64_t e = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4; e < n_out; e += (int64_t)gridDim.x * 1024) {
    const int64_t n4 = e < job.lo ? e : job.hi + (e - job.lo);
    *reinterpret_cast<float4*>(row + n4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

void launch_biquad_classes(const BiquadJob* d_reps, int n_classes, int64_t n_frames, int64_t n_quanta, int sample_rate, int32_t* d_last, int32_t* d_ent,
                           int* d_wide_scratch, unsigned char* d_cs, size_t cs_stride, cudaStream_t s) {
  if (n_classes <= 0 || n_frames <= 0) return;
  cudaMemsetAsync(d_wide_scratch, 0, sizeof(int) * ((n_classes + 15) / 16), s);
  k_biquad_select<<<dim3((unsigned)((n_quanta + kSelWarps - 1) / kSelWarps), (unsigned)n_classes), kSelWarps * 32, 0, s>>>(d_reps, sample_rate, n_quanta, n_frames, d_last, d_wide_scratch);
  k_biquad_entry<<<(unsigned)n_classes, kEntThreads, 0, s>>>(n_classes, n_quanta, d_last, d_ent);
  k_biquad_class_coef<<<dim3((unsigned)(n_frames / kBqChunkFrames), (unsigned)n_classes), kBqChunkFrames, 0, s>>>(d_reps, sample_rate, n_quanta, n_frames, d_ent, d_cs, cs_stride);
}
void launch_biquad_zero_outside(const BiquadJob* d_jobs, int n_jobs, int64_t n_frames, cudaStream_t s) {
  if (n_jobs <= 0) return;
  k_biquad_zero_outside<<<dim3(16, (unsigned)(2 * n_jobs)), 256, 0, s>>>(d_jobs, n_frames);
}

void launch_biquad(const BiquadJob* d_jobs, int n_jobs, int64_t n_frames, int64_t n_quanta, int sample_rate, int32_t* d_last,
                   int32_t* d_ent, float4* d_s1t, float4* d_s2t, float2* d_states, int* d_flags, cudaStream_t s) {
  if (n_jobs <= 0 || n_frames <= 0) return;
  const int groups = (n_jobs + 15) / 16;
  const unsigned n_slabs = (unsigned)(n_frames / 32);
  int* d_wide = d_flags + groups;  // [groups] stream layout flags, set by the select pass (layout of d_flags: biquad_lanes.cu)
  cudaMemsetAsync(d_wide, 0, sizeof(int) * groups, s);
  k_biquad_select<<<dim3((unsigned)((n_quanta + kSelWarps - 1) / kSelWarps), (unsigned)n_jobs), kSelWarps * 32, 0, s>>>(d_jobs, sample_rate, n_quanta, n_frames, d_last, d_wide);
  k_biquad_entry<<<(unsigned)n_jobs, kEntThreads, 0, s>>>(n_jobs, n_quanta, d_last, d_ent);
  k_biquad_resolve<<<dim3((n_slabs + kResSlabs - 1) / kResSlabs, (unsigned)groups), 512, 0, s>>>(d_jobs, n_jobs, sample_rate, n_quanta, n_frames, d_ent, d_s1t, d_s2t, d_wide);
  launch_biquad_lanes(d_jobs, n_jobs, n_frames, d_s1t, d_s2t, d_states, d_flags, s);
}

}  // namespace gac
