// fft_r16.cu — K5 / K7 with a radix-16 first-level transform: H/16 threads per H-point complex transform, 16 points per thread
// (fft2_core.cuh).  H = 128 frames per partition (what ConvolverNode uses, Nodes/ConvolverNode.cs:55): EIGHT threads per
// transform, stage A = one radix-16 butterfly over stride 8 + twiddles, one exchange through shared memory, stage C = two 8-point
// transforms on the thread's 16 contiguous points.  H = 256: 16 threads, stage C = one 16-point transform.  H = 512 (BASELINE
// config 5): 32 threads, stages A and B (two exchanges) and a radix-2 stage C.  The warp-per-transform kernels of fft.cu spend ~620 warp instructions
// per transform (radix-2 + five shuffle stages, ncu: 45 % issue utilisation, instruction-bound); this core needs ~5x fewer.
//
// Same arithmetic contract as fft.cu: Forward == numpy.fft.rfft of the zero-padded 256-frame block, Inverse == irfft +
// overlap-add (FftFlat/RealFourierTransform.cs:62-131, PartitionedConvolver.cs:106-124 and :134-150), float32.
// Both kernels work on the TRANSPOSED spectrograms of fft2.cu (row k = 0..128 of XT / YT, block time contiguous).
#include <cstdlib>

#include "fft2_core.cuh"
#include "gac_kernels.h"

namespace gac {

namespace {
constexpr int NTHR = 128;         // threads per CTA
constexpr int K5_DEFAULT_ROUNDS = 1;

__device__ __forceinline__ float2 cmul1(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cmulc1(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }

// Geometry of the first-level transform of H complex points (H = frames per partition): T = H/16 threads per transform,
// NTHR/T transforms in flight per CTA, two rounds per CTA -> COLS tile columns (blocks).
template <int H>
struct Geo {
  static constexpr int T = H / 16;
  static constexpr int GROUPS = NTHR / T;          // 16 / 8 / 4
  static constexpr int COLS = 2 * GROUPS;          // 32 / 16 / 8
  static constexpr int LD = COLS + 1;              // odd leading dimension of the [bin][block] tile: conflict-free column access
  static constexpr int TILE = ((H + 1) * LD + 1) & ~1;  // float2, even count (16-byte alignment of what follows)
  // float2 per transform buffer: the padded extent r16::pad(H - 1) + 1 rounded up to 16, plus 8 so that neighbouring buffers start
  // 64 bytes apart modulo 128 (H = 128: a half-warp is two transforms, their 64-bit accesses must not share banks): 152 / 312 / 616
  static constexpr int ZS = (((H - 1) + 2 * ((H - 1) >> 4) + 8 * ((H - 1) >> 7) + 1 + 15) & ~15) + 8;
  // frequency index of register slot q of thread t after stage C (validated against a DFT on the host: tools/fft2_host_test.cu)
  __device__ __forceinline__ static constexpr int slot_k(int t, int q) {
    return H == 128 ? 2 * t + (q >> 3) + 16 * f2::rev3(q & 7)
           : H == 256 ? t + 16 * r16::rev4(q)
                      : (t >> 1) + 16 * (8 * (t & 1) + (q >> 1)) + 256 * (q & 1);
  }
  // tile column of (group, round): a group's two columns are neighbours (K7 hands the even column's upper half to the odd one
  // in registers).  H = 128: 8 gi + 2 w + rnd, chosen so that a half-warp's two transforms write 16 different bank pairs;
  // H >= 256: a half-warp is one transform, any assignment is conflict-free.
  __device__ __forceinline__ static int tile_col(int tid, int rnd) {
    if (H == 128) return 8 * ((tid >> 3) & 3) + 2 * (tid >> 5) + rnd;
    return 2 * (tid / T) + rnd;
  }
};
}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// K5: COLS consecutive blocks of one channel -> H+1 rows x COLS columns of XT.
// tab = radix-16 twiddle table of M = H; tw = split twiddles e^{-2 pi i k / (2H)}, k < H.
// ---------------------------------------------------------------------------------------------------------------------
// R = rounds per CTA (transforms per thread group): 2 -> COLS = 2 GROUPS columns and 256-byte runs of XT per tile row; 1 -> half the
// tile, half the prefetch registers, more CTAs per SM (H = 128: 37 KB and <= 96 registers: 5 CTAs instead of 4).
template <int H, int R>
struct GeoF {
  using G = Geo<H>;
  static constexpr int COLS = R * G::GROUPS;
  static constexpr int LD = COLS + 1;
  static constexpr int TILE = (((H + 1) * LD + 1) & ~1) + 32;
  // tile index of (row k, column c): rows 128 apart would meet in the same bank pair (H = 512: a half-warp writes rows r and r + 128
  // of one column), so every block of 128 rows is shifted by eight entries
  __device__ __forceinline__ static int tix(int k, int c) { return k * LD + c + 8 * (k >> 7); }
  __device__ __forceinline__ static int tile_col(int tid, int rnd) {
    if (R == 2) return G::tile_col(tid, rnd);
    // one column per group.  H = 128: a thread writes the rows 2 t + const of its column into the tile (leading dimension COLS + 1:
    // bank pair = row + column), so the two groups of a half-warp take NEIGHBOURING columns: 16 different bank pairs per 64-bit access
    return tid / G::T;
  }
};
template <int H, int R>
__global__ void __launch_bounds__(NTHR, (H == 128 ? (R == 1 ? 6 : 4) : 3)) k_rfft_fwd_t8(const FftFwdJob* __restrict__ jobs, const float2* __restrict__ tab,
                                                                                       const float2* __restrict__ tw, int64_t ts) {
  using G = Geo<H>;
  using F = GeoF<H, R>;
  constexpr int T = G::T, COLS = F::COLS;
  extern __shared__ __align__(16) float2 smem[];
  float2* tileT = smem;            // Z[k] of every column: F::tix(k, column)
  float2* zb = smem + F::TILE;     // [GROUPS][ZS]
  const FftFwdJob job = jobs[blockIdx.y];
  const int tid = threadIdx.x;
  const float sc = job.scale ? *job.scale : 1.0f;
  const int64_t b0 = (int64_t)blockIdx.x * COLS;
  const int t = tid % T;
  float2* zg = zb + (tid / T) * G::ZS;

  // Both rounds' input frames (and gain-table entries) are requested before anything is computed: 32 independent 8-byte
  // loads in flight per thread.  Thread t of a group owns z[n] = x[2n] + i x[2n+1] for n = t + T j, j < 8.
  // Gate: frames outside the non-silent range [gate_lo, gate_hi) read as zero and are NOT loaded (`in` may alias a source buffer
  // that only covers that range).  The bounds are multiples of 128 and n_valid is a multiple of the partition size on this path,
  // so the test is made once per block: [lo_b, hi_b) = the open frames of the block, relative to its first frame.
  float2 xin[R][8], gin[R][8];
  int lo_b[R], hi_b[R];
#pragma unroll
  for (int rnd = 0; rnd < R; rnd++) {
    const int64_t f0 = (b0 + F::tile_col(tid, rnd)) * H;
    const int64_t lo = job.gate_lo - f0, hi = (job.gate_hi < job.n_valid ? job.gate_hi : job.n_valid) - f0;
    lo_b[rnd] = lo < 0 ? 0 : (lo > H ? H : (int)lo);
    hi_b[rnd] = hi < 0 ? 0 : (hi > H ? H : (int)hi);
    const float* __restrict__ src = job.in + f0;
    const float* __restrict__ gsrc = job.gain ? job.gain + f0 : nullptr;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int fi = 2 * (t + T * j);
      xin[rnd][j] = make_float2(0.f, 0.f);
      gin[rnd][j] = make_float2(job.gain_const, job.gain_const);
      if (fi >= lo_b[rnd] && fi < hi_b[rnd]) {  // (the bounds are even: a pair is open or closed as a whole)
        xin[rnd][j] = *reinterpret_cast<const float2*>(src + fi);
        if (gsrc) gin[rnd][j] = *reinterpret_cast<const float2*>(gsrc + fi);
      }
    }
  }
#pragma unroll
  for (int rnd = 0; rnd < R; rnd++) {
    const int bl = F::tile_col(tid, rnd);
    const int64_t f0 = (b0 + bl) * H;
    // fused GainNode multiply (Nodes/GainNode.cs:49-58), silent-quantum gate, stereo -> mono down-mix (AudioNodeInput.cs:214-228),
    // IR scale; the zero-padded upper half of the block (PartitionedConvolver.cs:107) gives z[n >= H/2] = 0
    float2 v[16];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int fi = 2 * (t + T * j);
      const bool open = fi >= lo_b[rnd] && fi < hi_b[rnd];
      float a0 = open ? __fmul_rn(xin[rnd][j].x, gin[rnd][j].x) : 0.f;
      float a1 = open ? __fmul_rn(xin[rnd][j].y, gin[rnd][j].y) : 0.f;
      if (job.in2) {
        float2 y = make_float2(0.f, 0.f);
        if (open) y = *reinterpret_cast<const float2*>(job.in2 + f0 + fi);
        const float c0 = open ? __fmul_rn(y.x, gin[rnd][j].x) : 0.f;
        const float c1 = open ? __fmul_rn(y.y, gin[rnd][j].y) : 0.f;
        a0 = __fmul_rn(__fadd_rn(a0, c0), job.mix_scale);
        a1 = __fmul_rn(__fadd_rn(a1, c1), job.mix_scale);
      }
      v[j] = job.scale ? make_float2(a0 * sc, a1 * sc) : make_float2(a0, a1);
      v[j + 8] = make_float2(0.f, 0.f);
    }
    r16::fwd_a<H>(v, zg, tab, t);
    __syncwarp();
    if constexpr (r16::Plan<H>::HAS_B) {
      r16::fwd_b<H>(zg, tab, t);
      __syncwarp();
    }
    float2 u[16];
    r16::load16(u, zg, t);
    r16::stage_c<r16::Plan<H>::L, false>(u);
    // Z[k] goes straight into the tile, column bl, natural order (slot_k): the split step is done by the store phase below, on the way
    // out — two shared-memory crossings fewer per transform (the L1 data pipe was this kernel's busiest unit: ncu, 78 %)
#pragma unroll
    for (int q = 0; q < 16; q++) tileT[F::tix(G::slot_k(t, q), bl)] = u[q];
    __syncwarp();  // (zg is re-used by the group's next round)
  }
  __syncthreads();
  {
    // Split step + store.  E = (Z[k] + conj Z[H-k]) / 2 ; O = (Z[k] - conj Z[H-k]) / (2i) ; X[k] = E + w_k O, w_k = e^{-2 pi i k / 2H}.
    // Bins k and H - k come from the same pair: E[H-k] = conj E[k], O[H-k] = conj O[k], w_{H-k} = -conj w_k, so
    // X[H-k] = conj(E - w_k O): one complex product and two shared-memory reads for two bins.  A thread takes column c of the
    // pairs k = tid / COLS, + NTHR / COLS, ... <= H/2 and stores rows k and H - k: every warp-wide store is COLS * 8-byte runs of XT rows.
    constexpr int RSTEP = NTHR / COLS;
    const int c = tid % COLS;
    if (b0 + c < job.n_blocks) {
      float2* __restrict__ dst = job.out + b0 + c;
      constexpr int ITERS = (H / 2 + RSTEP) / RSTEP;  // pairs per thread, the last sweep partly idle
      // every pair's operands are requested before the first product (the loop is unrolled in full, predicated on k <= H/2)
      float2 av[ITERS], zv[ITERS], wv[ITERS];
#pragma unroll
      for (int it = 0; it < ITERS; it++) {
        const int k = tid / COLS + it * RSTEP;
        const bool ok = k <= H / 2;
        av[it] = ok ? tileT[F::tix(k, c)] : make_float2(0.f, 0.f);
        zv[it] = (ok && k > 0) ? tileT[F::tix(H - k, c)] : make_float2(0.f, 0.f);
        wv[it] = (ok && k > 0 && k < H / 2) ? tw[k] : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int it = 0; it < ITERS; it++) {
        const int k = tid / COLS + it * RSTEP;
        if (k > H / 2) break;
        const float2 a = av[it];
        if (k == 0) {
          dst[0] = make_float2(a.x + a.y, 0.f);                   // row 0: DC
          dst[(int64_t)H * ts] = make_float2(a.x - a.y, 0.f);     // row H: Nyquist
        } else if (k == H / 2) {  // pairs with itself: E = (Re Z, 0), O = (Im Z, 0), w = -i
          dst[(int64_t)k * ts] = make_float2(a.x, -a.y);
        } else {
          const float2 z = zv[it];
          const float2 e = make_float2(0.5f * (a.x + z.x), 0.5f * (a.y - z.y));
          const float2 o = make_float2(0.5f * (a.y + z.y), -0.5f * (a.x - z.x));
          const float2 wo = cmul1(wv[it], o);
          dst[(int64_t)k * ts] = make_float2(e.x + wo.x, e.y + wo.y);
          dst[(int64_t)(H - k) * ts] = make_float2(e.x - wo.x, -(e.y - wo.y));
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// K7: COLS - 1 consecutive blocks of one channel (tile columns 1..COLS-1) plus their predecessor (column 0) from YT; inverse
// transforms, overlap-add (PartitionedConvolver.cs:146-150: out[i] = (float)r[i] + overlap[i]; overlap[i] = (float)r[i+B]).
// Every column's upper half goes to shared memory; after one barrier each column adds its left neighbour's.
// ---------------------------------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(NTHR) k_irfft_ola_t8(const FftInvJob* __restrict__ jobs, const float2* __restrict__ tab, const float2* __restrict__ tw,
                                                       int64_t ts) {
  using G = Geo<H>;
  constexpr int T = G::T, LD = G::LD, COLS = G::COLS;
  extern __shared__ __align__(16) float2 smem[];
  float2* tileT = smem;                                // [H + 1][LD], column c <-> block b0 - 1 + c
  float2* zb = smem + G::TILE;                         // [GROUPS][ZS]
  float2* hi = zb + G::GROUPS * G::ZS;                 // [GROUPS][HS]: upper halves r[B:2B] of the odd columns, as float2 pairs
  constexpr int HS = H / 2 + 8;                        // (+8: two groups of a half-warp 64 bytes apart modulo 128)
  const FftInvJob job = jobs[blockIdx.y];
  const int tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * (COLS - 1);
  if (b0 >= job.n_blocks) return;
  {
    // tile load: thread = column cc of rows rr, rr + 4, ...  The whole 33 KB tile is put in flight at once with 8-byte
    // cp.async copies (LDGSTS, zero-filled outside the spectrogram): the kernel is latency-bound otherwise (ncu: 75 % of the
    // samples on long_scoreboard with register-staged loads).  True-stereo pairs need an add and take the register path.
    constexpr int RSTEP = NTHR / COLS;  // rows covered per sweep
    const int cc = tid % COLS, rr = tid / COLS;
    const int64_t bcol = b0 - 1 + cc;
    const bool col_ok = bcol >= 0 && bcol < job.n_blocks;
    if (!job.in2 && job.n_parts <= 1) {
      const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(tileT + cc);
      const float2* src0 = job.in + (col_ok ? bcol : 0);
      const uint32_t nbytes = col_ok ? 8u : 0u;
#pragma unroll 4
      for (int row = rr; row <= H; row += RSTEP)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst0 + (uint32_t)(row * LD * 8)), "l"(src0 + (int64_t)row * ts), "r"(nbytes)
                     : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
      // spectrograms that are summed before the inverse transform (linear): the pair of a true-stereo output, or the partial
      // sums of a fan-in group
      constexpr int U = 8;
      const int np = job.in2 ? 2 : job.n_parts;
      const int64_t pstride = job.in2 ? (int64_t)(job.in2 - job.in) : job.part_stride;
      for (int r0 = rr; r0 <= H; r0 += RSTEP * U) {
        float2 y[U];
#pragma unroll
        for (int u = 0; u < U; u++) y[u] = make_float2(0.f, 0.f);
        if (col_ok) {
          for (int p = 0; p < np; p++) {
            const float2* src = job.in + (int64_t)p * pstride + bcol;
            float2 a[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
              const int row = r0 + RSTEP * u;
              a[u] = row <= H ? src[(int64_t)row * ts] : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U; u++) y[u] = make_float2(y[u].x + a[u].x, y[u].y + a[u].y);
          }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int row = r0 + RSTEP * u;
          if (row <= H) tileT[row * LD + cc] = y[u];
        }
      }
    }
  }
  __syncthreads();
  const int t = tid % T;
  float2* zg = zb + (tid / T) * G::ZS;
  const float inv_h = 1.0f / (float)H;
  float2 lo[2][8], carry[8];
#pragma unroll
  for (int rnd = 0; rnd < 2; rnd++) {
    const int c = G::tile_col(tid, rnd);
    // inverse split step: Z[k] = E + i O with E = (X[k] + conj X[H-k]) / 2, O = (X[k] - conj X[H-k]) / 2 * e^{+2 pi i k / 2H}
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const int k = t + T * j;
      float2 z;
      if (k == 0) {
        const float x0 = tileT[c].x, xh = tileT[H * LD + c].x;  // DC, Nyquist
        z = make_float2(0.5f * (x0 + xh), 0.5f * (x0 - xh));
      } else {
        const float2 a = tileT[k * LD + c];
        const float2 cc2 = tileT[(H - k) * LD + c];
        const float2 e = make_float2(0.5f * (a.x + cc2.x), 0.5f * (a.y - cc2.y));
        const float2 d = make_float2(0.5f * (a.x - cc2.x), 0.5f * (a.y + cc2.y));
        const float2 o = cmulc1(d, tw[k]);
        z = make_float2(e.x - o.y, e.y + o.x);
      }
      zg[k] = z;
    }
    __syncwarp();
    float2 u[16];
#pragma unroll
    for (int q = 0; q < 16; q++) u[q] = zg[G::slot_k(t, q)];
    __syncwarp();
    r16::stage_c<r16::Plan<H>::L, true>(u);
    r16::store16(u, zg, t);
    __syncwarp();
    if constexpr (r16::Plan<H>::HAS_B) {
      r16::inv_b<H>(zg, tab, t);
      __syncwarp();
    }
    float2 v[16];
    r16::inv_a<H>(v, zg, tab, t);
    __syncwarp();
    // v[j] = H * z[t + T j]; r[2n] = Re z[n], r[2n+1] = Im z[n]: n < H/2 is the block's own half, n >= H/2 the carried one.
    // A group's two columns are neighbours (even c in round 0, c + 1 in round 1): the even column's upper half is handed over
    // in registers, only the odd column's goes through shared memory to the group that owns column c + 2.
#pragma unroll
    for (int j = 0; j < 8; j++) {
      lo[rnd][j] = make_float2(v[j].x * inv_h, v[j].y * inv_h);
      const float2 up = make_float2(v[j + 8].x * inv_h, v[j + 8].y * inv_h);
      if (rnd == 0) carry[j] = up;
      else hi[(c >> 1) * HS + t + T * j] = up;
    }
  }
  __syncthreads();
#pragma unroll
  for (int rnd = 0; rnd < 2; rnd++) {
    const int c = G::tile_col(tid, rnd);
    const int64_t b = b0 - 1 + c;
    if (c >= 1 && b < job.n_blocks) {
      float* out = job.out + b * H;
      float* out2 = job.out2 ? job.out2 + b * H : nullptr;  // mono result duplicated (1 -> 2 up-mix copy at the next input)
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float2 o = rnd == 1 ? carry[j] : hi[((c - 1) >> 1) * HS + t + T * j];
        const float2 r = make_float2(lo[rnd][j].x + o.x, lo[rnd][j].y + o.y);
        *reinterpret_cast<float2*>(out + 2 * (t + T * j)) = r;
        if (out2) *reinterpret_cast<float2*>(out2 + 2 * (t + T * j)) = r;
      }
    }
  }
}

template <int H, int R>
static void fwd_launch_r(const FftFwdJob* d_jobs, int n_jobs, int64_t max_blocks, int64_t t_stride, const float2* d_tab, const float2* d_tw, cudaStream_t s) {
  using G = Geo<H>;
  using F = GeoF<H, R>;
  constexpr size_t smem = sizeof(float2) * (size_t)(F::TILE + G::GROUPS * G::ZS);
  GAC_SMEM_OPT_IN((k_rfft_fwd_t8<H, R>), smem);
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    const int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
    dim3 grid((unsigned)((max_blocks + F::COLS - 1) / F::COLS), (unsigned)nj);
    k_rfft_fwd_t8<H, R><<<grid, NTHR, smem, s>>>(d_jobs + j0, d_tab, d_tw, t_stride);
  }
}
// rounds per CTA of the 128-frame K5: 1 by default (72 registers, 37 KB: 6 CTAs per SM; measured 0.50 vs 0.52 ms on the C3 shard);
// GAC_K5_ROUNDS=2 for A/B measurements
static int k5_rounds() {
  static const int v = [] {
    const char* e = getenv("GAC_K5_ROUNDS");
    return e && atoi(e) == 1 ? 1 : (e && atoi(e) == 2 ? 2 : K5_DEFAULT_ROUNDS);
  }();
  return v;
}
template <int H>
static void fwd_launch(const FftFwdJob* d_jobs, int n_jobs, int64_t max_blocks, int64_t t_stride, const float2* d_tab, const float2* d_tw, cudaStream_t s) {
  if (H == 128 && k5_rounds() == 1) fwd_launch_r<128, 1>(d_jobs, n_jobs, max_blocks, t_stride, d_tab, d_tw, s);
  else fwd_launch_r<H, 2>(d_jobs, n_jobs, max_blocks, t_stride, d_tab, d_tw, s);
}
template <int H>
static void inv_launch(const FftInvJob* d_jobs, int n_jobs, int64_t max_blocks, int64_t t_stride, const float2* d_tab, const float2* d_tw, cudaStream_t s) {
  using G = Geo<H>;
  constexpr size_t smem = sizeof(float2) * (size_t)(G::TILE + G::GROUPS * G::ZS + G::GROUPS * (H / 2 + 8));
  GAC_SMEM_OPT_IN(k_irfft_ola_t8<H>, smem);
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    const int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
    dim3 grid((unsigned)((max_blocks + G::COLS - 2) / (G::COLS - 1)), (unsigned)nj);
    k_irfft_ola_t8<H><<<grid, NTHR, smem, s>>>(d_jobs + j0, d_tab, d_tw, t_stride);
  }
}

// d_tab16 = the concatenated radix-16 twiddle tables (fft2.cu); B = 128, 256 or 512 frames per partition
void launch_rfft_fwd_t8(const FftFwdJob* d_jobs, int n_jobs, int64_t max_blocks, int B, int64_t t_stride, const float2* d_tab16, const float2* d_tw,
                        cudaStream_t s) {
  if (n_jobs <= 0 || max_blocks <= 0) return;
  const float2* tab = d_tab16 + fft2_table_offset(B);
  switch (B) {
    case 128: fwd_launch<128>(d_jobs, n_jobs, max_blocks, t_stride, tab, d_tw, s); break;
    case 256: fwd_launch<256>(d_jobs, n_jobs, max_blocks, t_stride, tab, d_tw, s); break;
    case 512: fwd_launch<512>(d_jobs, n_jobs, max_blocks, t_stride, tab, d_tw, s); break;
  }
}
void launch_irfft_ola_t8(const FftInvJob* d_jobs, int n_jobs, int64_t max_blocks, int B, int64_t t_stride, const float2* d_tab16, const float2* d_tw,
                         cudaStream_t s) {
  if (n_jobs <= 0 || max_blocks <= 0) return;
  const float2* tab = d_tab16 + fft2_table_offset(B);
  switch (B) {
    case 128: inv_launch<128>(d_jobs, n_jobs, max_blocks, t_stride, tab, d_tw, s); break;
    case 256: inv_launch<256>(d_jobs, n_jobs, max_blocks, t_stride, tab, d_tw, s); break;
    case 512: inv_launch<512>(d_jobs, n_jobs, max_blocks, t_stride, tab, d_tw, s); break;
  }
}

}  // namespace gac
