// fft_r16.cu — K5 / K7 for 128-frame partitions (what ConvolverNode uses, Nodes/ConvolverNode.cs:55) with a radix-16
// first-level transform: EIGHT threads per 128-point complex transform, 16 points per thread (fft2_core.cuh, plan M = 128:
// stage A = one radix-16 butterfly over stride 8 + twiddles, one exchange through shared memory, stage C = two 8-point
// transforms on the thread's 16 contiguous points).  The warp-per-transform kernels of fft.cu spend ~620 warp instructions
// per transform (radix-2 + five shuffle stages, ncu: 45 % issue utilisation, instruction-bound); this core needs ~5x fewer.
//
// Same arithmetic contract as fft.cu: Forward == numpy.fft.rfft of the zero-padded 256-frame block, Inverse == irfft +
// overlap-add (FftFlat/RealFourierTransform.cs:62-131, PartitionedConvolver.cs:106-124 and :134-150), float32.
// Both kernels work on the TRANSPOSED spectrograms of fft2.cu (row k = 0..128 of XT / YT, block time contiguous).
#include "fft2_core.cuh"
#include "gac_kernels.h"

namespace gac {

namespace {
constexpr int H = 128;            // complex points per transform = frames per partition
constexpr int TPT = 8;            // threads per transform
constexpr int NTHR = 128;         // threads per CTA -> 16 transforms in flight, 4 per warp
constexpr int ZS = 152;           // float2 per transform buffer: r16::smem_elems(128) = 144, +8 so that the four buffers of a warp
                                  // start 64 bytes apart modulo 128 (half-warp = two transforms: conflict-free 64-bit accesses)
constexpr int LD = 33;            // leading dimension of the [bin][block] tile (odd: conflict-free column access)
constexpr int TILE = (H + 1) * LD + 1;  // float2, rounded to an even count (16-byte alignment of what follows)

__device__ __forceinline__ float2 cmul1(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cmulc1(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }

// frequency index of register slot q of thread t after stage C (validated on the host: scratch/fft2_host_test.cu)
__device__ __forceinline__ constexpr int slot_k(int t, int q) { return 2 * t + (q >> 3) + 16 * f2::rev3(q & 7); }
// tile column of (warp w, group-in-warp gi, round): a bijection onto 0..31 chosen so that a half-warp's writes of one
// bin-row pair land in 16 different bank pairs
__device__ __forceinline__ int tile_col(int w, int gi, int rnd) { return 8 * gi + 2 * w + rnd; }
}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// K5: 32 consecutive blocks of one channel -> 129 rows x 32 columns of XT.
// tab = radix-16 twiddle table of M = 128 ([4][8]); tw = split twiddles e^{-2 pi i k / 256}, k < 128.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHR, 4) k_rfft_fwd_t8(const FftFwdJob* __restrict__ jobs, const float2* __restrict__ tab,
                                                         const float2* __restrict__ tw, int64_t ts) {
  extern __shared__ __align__(16) float2 smem[];
  float2* tileT = smem;         // [129][33]
  float2* zb = smem + TILE;     // [16][ZS]
  const FftFwdJob job = jobs[blockIdx.y];
  const int tid = threadIdx.x;
  const float sc = job.scale ? *job.scale : 1.0f;
  const int64_t b0 = (int64_t)blockIdx.x * 32;
  const int t = tid & 7, w = tid >> 5, gi = (tid >> 3) & 3;
  float2* zg = zb + (tid >> 3) * ZS;

  // Both rounds' input frames (and gain-table entries) are requested before anything is computed: 32 independent 8-byte
  // loads in flight per thread.  Thread t of a group owns z[n] = x[2n] + i x[2n+1] for n = t + 8 j, j < 8.
  float2 xin[2][8], gin[2][8];
#pragma unroll
  for (int rnd = 0; rnd < 2; rnd++) {
    const int64_t f0 = (b0 + tile_col(w, gi, rnd)) * H;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int64_t g = f0 + 2 * (t + 8 * j);
      xin[rnd][j] = make_float2(0.f, 0.f);
      gin[rnd][j] = make_float2(job.gain_const, job.gain_const);
      if (g + 1 < job.n_valid) {  // n_valid is a multiple of the partition size on this path
        xin[rnd][j] = *reinterpret_cast<const float2*>(job.in + g);
        if (job.gain) gin[rnd][j] = *reinterpret_cast<const float2*>(job.gain + g);
      }
    }
  }
#pragma unroll
  for (int rnd = 0; rnd < 2; rnd++) {
    const int bl = tile_col(w, gi, rnd);
    const int64_t f0 = (b0 + bl) * H;
    // fused GainNode multiply (Nodes/GainNode.cs:49-58), silent-quantum gate, stereo -> mono down-mix (AudioNodeInput.cs:214-228),
    // IR scale; the zero-padded upper half of the block (PartitionedConvolver.cs:107) gives z[n >= 64] = 0
    float2 v[16];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int64_t g = f0 + 2 * (t + 8 * j);
      const bool open0 = g >= job.gate_lo && g < job.gate_hi, open1 = g + 1 >= job.gate_lo && g + 1 < job.gate_hi;
      float a0 = open0 ? __fmul_rn(xin[rnd][j].x, gin[rnd][j].x) : 0.f;
      float a1 = open1 ? __fmul_rn(xin[rnd][j].y, gin[rnd][j].y) : 0.f;
      if (job.in2) {
        float2 y = make_float2(0.f, 0.f);
        if (g + 1 < job.n_valid) y = *reinterpret_cast<const float2*>(job.in2 + g);
        const float c0 = open0 ? __fmul_rn(y.x, gin[rnd][j].x) : 0.f;
        const float c1 = open1 ? __fmul_rn(y.y, gin[rnd][j].y) : 0.f;
        a0 = __fmul_rn(__fadd_rn(a0, c0), job.mix_scale);
        a1 = __fmul_rn(__fadd_rn(a1, c1), job.mix_scale);
      }
      v[j] = make_float2(a0 * sc, a1 * sc);
      v[j + 8] = make_float2(0.f, 0.f);
    }
    r16::fwd_a<H>(v, zg, tab, t);
    __syncwarp();
    float2 u[16];
    r16::load16(u, zg, t);
    r16::stage_c<8, false>(u);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 16; q++) zg[slot_k(t, q)] = u[q];  // natural order Z[k]
    __syncwarp();
    // split step: E = (Z[k] + conj Z[H-k]) / 2 ; O = (Z[k] - conj Z[H-k]) / (2i) ; X[k] = E + e^{-2 pi i k / 256} O
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const int k = t + 8 * j;
      const float2 a = zg[k];
      const float2 c = zg[(H - k) & (H - 1)];
      if (k == 0) {
        tileT[bl] = make_float2(a.x + a.y, 0.f);           // row 0: DC
        tileT[H * LD + bl] = make_float2(a.x - a.y, 0.f);  // row 128: Nyquist
      } else {
        const float2 e = make_float2(0.5f * (a.x + c.x), 0.5f * (a.y - c.y));
        const float2 o = make_float2(0.5f * (a.y + c.y), -0.5f * (a.x - c.x));
        const float2 wo = cmul1(tw[k], o);
        tileT[k * LD + bl] = make_float2(e.x + wo.x, e.y + wo.y);
      }
    }
    __syncwarp();
  }
  __syncthreads();
  for (int idx = tid; idx < (H + 1) * 32; idx += NTHR) {
    const int row = idx >> 5, c = idx & 31;
    if (b0 + c < job.n_blocks) job.out[(int64_t)row * ts + b0 + c] = tileT[row * LD + c];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// K7: 31 consecutive blocks of one channel (tile columns 1..31) plus their predecessor (column 0) from YT; inverse
// transforms, overlap-add (PartitionedConvolver.cs:146-150: out[i] = (float)r[i] + overlap[i]; overlap[i] = (float)r[i+B]).
// Every column's upper half goes to shared memory; after one barrier each column adds its left neighbour's.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kInvCols = 31;  // new blocks per CTA
__global__ void __launch_bounds__(NTHR) k_irfft_ola_t8(const FftInvJob* __restrict__ jobs, const float2* __restrict__ tab, const float2* __restrict__ tw,
                                                       int64_t ts) {
  extern __shared__ __align__(16) float2 smem[];
  float2* tileT = smem;                                // [129][33], column c <-> block b0 - 1 + c
  float2* zb = smem + TILE;                            // [16][ZS]
  float2* hi = zb + 16 * ZS;                           // [16][64]: upper halves r[B:2B] of the odd columns, as float2 pairs
  const FftInvJob job = jobs[blockIdx.y];
  const int tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * kInvCols;
  if (b0 >= job.n_blocks) return;
  {
    // tile load: thread = column cc of rows rr, rr + 4, ...  The whole 33 KB tile is put in flight at once with 8-byte
    // cp.async copies (LDGSTS, zero-filled outside the spectrogram): the kernel is latency-bound otherwise (ncu: 75 % of the
    // samples on long_scoreboard with register-staged loads).  True-stereo pairs need an add and take the register path.
    const int cc = tid & 31, rr = tid >> 5;
    const int64_t bcol = b0 - 1 + cc;
    const bool col_ok = bcol >= 0 && bcol < job.n_blocks;
    if (!job.in2) {
      const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(tileT + cc);
      const float2* src0 = job.in + (col_ok ? bcol : 0);
      const uint32_t nbytes = col_ok ? 8u : 0u;
#pragma unroll 4
      for (int row = rr; row <= H; row += 4)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst0 + (uint32_t)(row * LD * 8)), "l"(src0 + (int64_t)row * ts), "r"(nbytes)
                     : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
      constexpr int U = 8;
      for (int r0 = rr; r0 <= H; r0 += 4 * U) {
        float2 y[U], y2[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int row = r0 + 4 * u;
          y[u] = y2[u] = make_float2(0.f, 0.f);
          if (row <= H && col_ok) {
            y[u] = job.in[(int64_t)row * ts + bcol];
            y2[u] = job.in2[(int64_t)row * ts + bcol];  // true stereo: the pair is summed as spectra (linear)
          }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int row = r0 + 4 * u;
          if (row <= H) tileT[row * LD + cc] = make_float2(y[u].x + y2[u].x, y[u].y + y2[u].y);
        }
      }
    }
  }
  __syncthreads();
  const int t = tid & 7, w = tid >> 5, gi = (tid >> 3) & 3;
  float2* zg = zb + (tid >> 3) * ZS;
  const float inv_h = 1.0f / (float)H;
  float2 lo[2][8], carry[8];
#pragma unroll
  for (int rnd = 0; rnd < 2; rnd++) {
    const int c = tile_col(w, gi, rnd);
    // inverse split step: Z[k] = E + i O with E = (X[k] + conj X[H-k]) / 2, O = (X[k] - conj X[H-k]) / 2 * e^{+2 pi i k / 256}
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const int k = t + 8 * j;
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
    for (int q = 0; q < 16; q++) u[q] = zg[slot_k(t, q)];
    __syncwarp();
    r16::stage_c<8, true>(u);
    r16::store16(u, zg, t);
    __syncwarp();
    float2 v[16];
    r16::inv_a<H>(v, zg, tab, t);
    __syncwarp();
    // v[j] = H * z[t + 8 j]; r[2n] = Re z[n], r[2n+1] = Im z[n]: n < 64 is the block's own half, n >= 64 the carried one.
    // A group's two columns are neighbours (even c in round 0, c + 1 in round 1): the even column's upper half is handed over
    // in registers, only the odd column's goes through shared memory to the group that owns column c + 2.
#pragma unroll
    for (int j = 0; j < 8; j++) {
      lo[rnd][j] = make_float2(v[j].x * inv_h, v[j].y * inv_h);
      const float2 up = make_float2(v[j + 8].x * inv_h, v[j + 8].y * inv_h);
      if (rnd == 0) carry[j] = up;
      else hi[(c >> 1) * 64 + t + 8 * j] = up;
    }
  }
  __syncthreads();
#pragma unroll
  for (int rnd = 0; rnd < 2; rnd++) {
    const int c = tile_col(w, gi, rnd);
    const int64_t b = b0 - 1 + c;
    if (c >= 1 && b < job.n_blocks) {
      float* out = job.out + b * H;
      float* out2 = job.out2 ? job.out2 + b * H : nullptr;  // mono result duplicated (1 -> 2 up-mix copy at the next input)
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float2 o = rnd == 1 ? carry[j] : hi[((c - 1) >> 1) * 64 + t + 8 * j];
        const float2 r = make_float2(lo[rnd][j].x + o.x, lo[rnd][j].y + o.y);
        *reinterpret_cast<float2*>(out + 2 * (t + 8 * j)) = r;
        if (out2) *reinterpret_cast<float2*>(out2 + 2 * (t + 8 * j)) = r;
      }
    }
  }
}

constexpr size_t kFwdSmem = sizeof(float2) * (size_t)(TILE + 16 * ZS);
constexpr size_t kInvSmem = sizeof(float2) * (size_t)(TILE + 16 * ZS + 16 * 64);

void launch_rfft_fwd_t8(const FftFwdJob* d_jobs, int n_jobs, int64_t max_blocks, int64_t t_stride, const float2* d_tab128, const float2* d_tw,
                        cudaStream_t s) {
  if (n_jobs <= 0 || max_blocks <= 0) return;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_rfft_fwd_t8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem);
    attr = true;
  }
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    const int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
    dim3 grid((unsigned)((max_blocks + 31) / 32), (unsigned)nj);
    k_rfft_fwd_t8<<<grid, NTHR, kFwdSmem, s>>>(d_jobs + j0, d_tab128, d_tw, t_stride);
  }
}
void launch_irfft_ola_t8(const FftInvJob* d_jobs, int n_jobs, int64_t max_blocks, int64_t t_stride, const float2* d_tab128, const float2* d_tw,
                         cudaStream_t s) {
  if (n_jobs <= 0 || max_blocks <= 0) return;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_irfft_ola_t8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kInvSmem);
    attr = true;
  }
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    const int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
    dim3 grid((unsigned)((max_blocks + kInvCols - 1) / kInvCols), (unsigned)nj);
    k_irfft_ola_t8<<<grid, NTHR, kInvSmem, s>>>(d_jobs + j0, d_tab128, d_tw, t_stride);
  }
}

}  // namespace gac
