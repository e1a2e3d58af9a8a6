// gac_kernels.h — launcher declarations for the sm_100a kernels (internal; not part of the C ABI).
//
// Data layout in HBM:
//   time-domain signals   float  sig[S][2][Npad]           planar, Npad multiple of the partition B
//   spectra ("packed")    float2 X[chan][rows][B]          row = one partition/block; bin 0 holds
//                                                           (DC.re, Nyquist.re), bins 1..B-1 (re, im).
//                                                           The reference stores C = B+1 planar re/im
//                                                           floats (PartitionedConvolver.cs:41,46-51);
//                                                           DC and Nyquist are purely real
//                                                           (FftFlat/RealFourierTransform.cs:76-78).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: a process-wide "done" flag leaves every
// device but the first without the opt-in (launches with more than 48 KB of dynamic shared memory then fail with "invalid
// argument"), and an unsynchronised flag races when two render threads launch the kernel for the first time.  One mutex and one
// bit per device at each call site; `bytes` may grow over time (the largest value seen per device is kept).
#define GAC_SMEM_OPT_IN(kernel, bytes)                                                                   \
  do {                                                                                                    \
    static std::mutex mu_;                                                                                \
    static size_t set_[64] = {0};                                                                         \
    int dev_ = 0;                                                                                         \
    cudaGetDevice(&dev_);                                                                                 \
    std::lock_guard<std::mutex> lk_(mu_);                                                                 \
    if (dev_ >= 0 && dev_ < 64 && (size_t)(bytes) > set_[dev_]) {                                         \
      if ((size_t)(bytes) > 48 * 1024)                                                                    \
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));          \
      set_[dev_] = (size_t)(bytes);                                                                       \
    }                                                                                                     \
  } while (0)

namespace gac {

// ------------------------------------------------------------------ FFT (fft.cu)
// Twiddle table e^{-2*pi*i*k/N}, k in [0, N/2), N = 2B, as float2, computed on the host in double.
struct FftFwdJob {
  const float* in;       // B-sample blocks, contiguous: block b at in + b*B
  float2* out;           // packed spectra: block b at out + b*B
  const float* scale;    // optional device scalar multiplied into every sample (IR normalisation), or nullptr
  const float* gain;     // optional per-sample gain table (fused GainNode), or nullptr
  float gain_const;      // constant gain used when gain == nullptr (1.0f = none)
  int64_t n_valid;       // samples of `in` that exist; the rest of the last block reads as zero
  int64_t n_blocks;
  int64_t gate_lo, gate_hi;  // samples outside [gate_lo, gate_hi) read as zero (silent-flagged quanta)
  const float* in2 = nullptr;  // optional second channel: input = (in + in2) * mix_scale (stereo -> mono down-mix)
  float mix_scale = 1.0f;
};
void launch_rfft_fwd(const FftFwdJob* d_jobs, int n_jobs, int64_t max_blocks, int B, const float2* d_tw, cudaStream_t s);
// the same transform for the channels of ONE buffer, described arithmetically and passed by value (no job array upload)
struct FftFwdUniform {
  const float* in_base;     // channel y at in_base + y*in_stride
  int64_t in_stride;
  float2* out_base;         // channel y at out_base + y*out_stride
  int64_t out_stride;
  const float* scale_base;  // channel y's scale at scale_base[y], or nullptr
  int64_t n_valid, n_blocks;
};
void launch_rfft_fwd_uniform(const FftFwdUniform& u, int n_jobs, int B, const float2* d_tw, cudaStream_t s);

struct FftInvJob {
  const float2* in;  // packed spectra, block b at in + b*B
  float* out;        // B-sample blocks, block b at out + b*B ; out[b] = r_b[0:B] + r_{b-1}[B:2B]
  int64_t n_blocks;
  const float2* in2 = nullptr;  // optional second spectrogram: out = ola(in) + ola(in2)  (true-stereo Sum)
  float* out2 = nullptr;        // optional second destination receiving the same samples (mono -> both rows)
  // k_irfft_ola_t8 only: `in` is the first of n_parts partial spectrograms, part_stride float2 apart, summed BEFORE the inverse
  // transform (the partial sums of a fan-in group, Fft2SumJob)
  int n_parts = 1;
  int64_t part_stride = 0;
};
void launch_irfft_ola(const FftInvJob* d_jobs, int n_jobs, int64_t max_blocks, int B, const float2* d_tw, cudaStream_t s);

// ------------------------------------------------------------------ spectral MAC (mac.cu)
// One job = one channel-convolver bin-group of 128 bins:  Y[b][k] = sum_p X[b-p][k] * H[p][k].
struct MacJob {
  const float2* X;  // points at row 0 (block 0), bin-group offset applied; rows [-pad_rows, rows_alloc) readable
  const float2* H;  // row p at H + p*stride; rows [0, P16) readable (zero beyond P)
  float2* Y;        // row b at Y + b*stride
  int P;            // partitions
  int has_dc;       // bin-group 0 carries the packed (DC, Nyquist) pair in bin 0
};
struct MacTile {
  int job;
  int b0;  // first output block of the tile
};
constexpr int kMacT = 16;         // output blocks per thread (register tile)
constexpr int kMacChunk = 16;     // partitions per pipeline stage
int mac_tile_blocks(int variant);  // output blocks per CTA for the tiled kernel (32 or 64)
// flavour: 0 = packed FFMA2 accumulation (default), 2 = scalar FFMA
// (two launches: the tiled kernel, then k_mac_dc for the packed DC/Nyquist bin; p_max = largest P among the jobs)
void launch_mac_tiled(const MacJob* d_jobs, int n_jobs, const MacTile* d_tiles, int n_tiles, int64_t n_blocks, int p_max, int stride,
                      int tile_blocks, int flavour, cudaStream_t s);
void launch_mac_stream(const MacJob* d_jobs, int n_jobs, int64_t n_blocks, int stride, cudaStream_t s);

// one convolver of a streaming (per-quantum) ConvolverNode: Y[k] = sum_p X[row - p][k] * H[p][k], exact reference order
struct RingMacJob {
  const float2* X;  // linear history of the input channel's spectra: row r at X + r*B; rows [row - P + 1, row] are read
  const float2* H;  // IR spectra, row p at H + p*B
  float2* Y;        // block j of a call at Y + j*B
  int P;
};
void launch_mac_ring(const RingMacJob* d_jobs, int n_jobs, int64_t row, int n_blocks, int B, cudaStream_t s);

// ------------------------------------------------------------------ second-level FFT MAC (fft2.cu)
// One job = one channel-convolver over transposed spectrograms: row k (k = 0..B) of XT/YT at X + k*xs / Y + k*ys.
struct Fft2Job {
  const float2* X;   // XT channel base
  const float2* H2;  // prepared second-level IR spectra of the IR channel: [B+1][fft2_h2_row_elems(M)]
  float2* Y;         // YT channel base
  int Lh;            // history blocks per segment (>= P-1, multiple of 16); V = M - Lh valid outputs per segment
  int nseg;          // segments of this launch
  int64_t b0 = 0;    // first output block of segment 0 (multiple of 16): lets two launches with different M share one spectrogram
};
// Fan-in fusion: the convolvers of several voices that end in the SAME fan-in (AudioNodeInput.MixBuffer, AudioNodeInput.cs:118-137)
// are summed where the sum is cheapest — as second-level spectra.  One job = one output channel of a chunk of such voices:
//     Y[k][.] = IFFT_M( sum_v FFT_M(X_v[k][window]) * H2_v[k][.] )
// one forward transform + multiply-accumulate per member, ONE inverse transform and one Y segment per job (linear: the same
// samples as transforming every member back and adding the results, up to float32 rounding order).
struct Fft2SumMember {
  const float2* X;   // the member's XT channel base
  const float2* H2;  // the member's second-level IR spectra [B+1][M]
};
struct Fft2SumJob {
  const Fft2SumMember* members;  // device array
  int n_members;
  float2* Y;         // partial YT channel base of this chunk
  int Lh, nseg;
  int64_t b0 = 0;
};
void launch_fft2_sum(const Fft2SumJob* d_jobs, int n_jobs, int max_seg, int C, int M, const float2* d_tab16, int64_t n_blocks, int64_t xs, int64_t ys,
                     cudaStream_t s);
int fft2_sum_ctas_per_sm(int M);  // resident CTAs per SM of k_fft2_sum16<M> (shared-memory bound): the planner sizes chunks with it
constexpr int kFft2TwLen = 8192;  // twiddle table exp(-2 pi i e / 8192)
// second-level transform length for P partitions (512..8192), 0 if the IR is too long; *Lh = history length
int fft2_pick_m(int P, int* Lh);
// float2 elements of one H2 row (one bin of one IR channel): M (the radix-16 plan keeps the row's 16-byte chunks XOR-swizzled)
int fft2_h2_row_elems(int M);
int fft2_r16_max();                  // largest transform length on the radix-16 plan (4096; GAC_FFT2_R16_MAX overrides for A/B runs)
// d_tw2: the 8192-entry table (radix-8 plan, M = 8192); d_tab16: the concatenated radix-16 tables (M = 512 .. 4096)
int fft2_table_total();              // float2 entries of the concatenated radix-16 tables
int fft2_table_offset(int M);        // where M's table starts (M = 128, 256: the first-level transforms of fft_r16.cu)
// K5 / K7 over the TRANSPOSED spectrograms of fft2.cu, radix-16 first-level core (fft_r16.cu; B = 128, 256, 512).  Same job
// structs as above, but job.out / job.in / job.in2 are XT / YT channel bases: bin k of block b lives at base[k*t_stride + b],
// k = 0..B (row 0 = (DC, 0), row B = (Nyquist, 0)); with in2 the two spectrograms are summed BEFORE the inverse transform (the
// transform is linear; ConvolverNode.Sum adds afterwards)
void launch_rfft_fwd_t8(const FftFwdJob* d_jobs, int n_jobs, int64_t max_blocks, int B, int64_t t_stride, const float2* d_tab16, const float2* d_tw,
                        cudaStream_t s);
void launch_irfft_ola_t8(const FftInvJob* d_jobs, int n_jobs, int64_t max_blocks, int B, int64_t t_stride, const float2* d_tab16, const float2* d_tw,
                         cudaStream_t s);
void fft2_fill_tables(float2* host);  // fills them (double precision, rounded once)
void launch_fft2_conv(const Fft2Job* d_jobs, int n_jobs, int max_seg, int C, int M, const float2* d_tw2, const float2* d_tab16, int64_t n_blocks,
                      int64_t xs, int64_t ys, cudaStream_t s);
// H: packed first-level IR spectra [n_ch][...h_ch_stride...] rows of B -> H2 [n_ch][B+1][M]
void launch_fft2_prep(const float2* d_H, int64_t h_ch_stride, int n_ch, int B, int P, int M, float2* d_H2, const float2* d_tw2, const float2* d_tab16,
                      cudaStream_t s);

// ------------------------------------------------------------------ node kernels (nodes.cu)
struct DevEvent {  // bit-compatible with gac_event / AudioParam.cs:360-367
  int32_t type;
  float value;
  float target;
  double time;
  double time_constant;
};
struct ParamJob {
  float value;
  int n_events;
  const DevEvent* events;
  float* out;       // [n_frames] for a-rate, [n_blocks] for k-rate
  int a_rate;
  int64_t q_lo = 0, q_hi = INT64_MAX;  // quanta this job writes (one job per epoch of a parameter edited between Render calls)
};
void launch_param_eval(const ParamJob* d_jobs, int n_jobs, const double* d_block_time, int64_t n_quanta, int sample_rate, cudaStream_t s);

// AudioParam with a modulation input: table[n] = clamp(table[n] + mod[n], min, max) on the frames where the modulator is non-silent
// (a-rate; k-rate tables hold one value per quantum and take the modulator's first frame of the quantum) — AudioParam.cs:114-166
struct ModJob {
  float* table;
  const float* mod;
  int64_t lo, hi;  // non-silent frames of the modulator (multiples of 128)
  float minv, maxv;
  int a_rate;
};
void launch_param_modulate(const ModJob* d_jobs, int n_jobs, int64_t n_frames, cudaStream_t s);

// Scheduled one-channel sources (Nodes/ConstantSourceNode.cs, Nodes/OscillatorNode.cs): frames [s0, s1) carry the source, the rest of
// [lo, hi) (the quanta the node plays in) zeros; rows outside [lo, hi) are not written (flagged silent)
struct SchedJob {
  float* dst[2];       // both rows receive the (mono) signal
  const float* table;  // a-rate Offset / Frequency values, or nullptr (constant)
  float value;
  int64_t lo, hi;      // quanta the node plays in, as frames
  int64_t s0, s1;      // sample-accurate start / stop
  int osc_type;        // -1: ConstantSourceNode; 0..3: OscillatorNode type
  double* chunk_sum;   // oscillator scratch: [ceil((s1 - s0) / 1024) + 1] phase-increment sums of 1024-frame chunks from s0
};
void launch_scheduled_sources(const SchedJob* d_jobs, int n_jobs, int64_t n_frames, int sample_rate, bool any_oscillator, cudaStream_t s);

struct SourceJob {
  const float* src[2];  // channel pointers (both the same for a mono buffer: the 1->2 up-mix copy of AudioNodeInput.cs:201-213)
  float* dst[2];        // sig rows
  int64_t pos0;         // buffer frame that lands on out frame out0
  int64_t out0;         // first emitted output frame (multiple of 128)
  int64_t n_emit;       // emitted frames (multiple of 128)
  // looping playback at rate 1 (AudioBufferSourceNode.cs:197-234): loop_len > 0 switches it on.  Emitted frame j reads buffer frame
  //   k = pos0 + j;  k < loop_end ? k : loop_start + (k - loop_end) % loop_len
  // except that a start position at or behind loop_end restarts the FIRST block at loop_start (:197-200): j < 128 -> loop_start + j % loop_len
  int64_t loop_start = 0, loop_end = 0, loop_len = 0;
};
void launch_source_copy(const SourceJob* d_jobs, int n_jobs, int64_t n_frames, cudaStream_t s);

constexpr int32_t kResampleCleared = INT32_MIN;
struct ResampleJob {
  const float* src[2];
  float* dst[2];
  const int32_t* k;   // per output frame: k >= 0: index of S0 in the source (S0..S3 = src[k..k+3]); k == kResampleCleared: the frame
                      // is zero; other k < 0: the window is not contiguous (a looping source at a loop seam), its four frame
                      // indices are x[4 * (-k - 1) ..]
  const float* t;     // per output frame: fractional position
  const int32_t* x;   // window exceptions, four source frame indices each (looping sources only; else nullptr)
  int64_t out0;       // first output frame
  int64_t n_emit;     // frames for which (k,t) exist and the block is kept
  int64_t n_zero_from;  // frames >= this (relative to out0) inside kept blocks are zero (stall point)
};
void launch_resample(const ResampleJob* d_jobs, int n_jobs, int64_t n_frames, cudaStream_t s);

struct GainJob {
  float* sig[2];
  const float* gain;  // a-rate table or nullptr
  float gain_const;
  int64_t lo, hi;     // active frame range; outside -> zeros (silent-flagged blocks, GainNode.cs:41-46)
};
void launch_gain(const GainJob* d_jobs, int n_jobs, int64_t n_frames, cudaStream_t s);

struct DelayJob {       // DelayNode: out[n] = d[n] >= 1 ? in[n - d[n]] : 0
  const float* in[2];
  float* out[2];        // distinct from `in`
  const float* dt;      // a-rate DelayTime table (seconds) or nullptr
  float dt_const;
  int max_delay;        // (int)(maxDelayTime * sampleRate)
  int64_t in_lo, in_hi; // frames where the input is non-silent
};
void launch_delay(const DelayJob* d_jobs, int n_jobs, int64_t n_frames, int sample_rate, cudaStream_t s);

struct PannerJob {      // StereoPannerNode, in place
  float* sig[2];
  const float* pan;     // a-rate table or nullptr
  float pan_const;
  int mono;             // the input block has one channel (ProcessMono) — rows duplicated
  int64_t lo, hi;
  int64_t sp_block;     // first frame of the one quantum whose channel count differs (see k_panner), or a negative value
  int sp_mode;          // 1: mono up-mixed to two equal channels, stereo formula; 2: stereo mixed down, mono formula
  unsigned long long* first_change;  // device scalar (preset huge): first frame behind that quantum whose pan differs from the
                                     // previous frame's — until then the gain pair of the odd quantum stays cached; nullptr = constant pan
};
// scan_changes: some job has sp_mode != 0 and a pan table (runs k_pan_first_change first)
void launch_panner(const PannerJob* d_jobs, int n_jobs, int64_t n_frames, bool scan_changes, cudaStream_t s);

struct BiquadJob {
  float* sig[2];       // output rows (and the input, unless `in` says otherwise)
  const float* in[2] = {nullptr, nullptr};  // input rows when they are not `sig`: in[c][n] for n in [lo, hi) — a rate-1 source read in place
  const float* freq;   // a-rate table or nullptr
  const float* q;      // a-rate table or nullptr
  const float* gain;   // k-rate table [n_quanta] or nullptr
  float freq_const, q_const, gain_const;
  int type;
  int64_t lo, hi;      // active frame range (multiples of 128)
  // scratch, per job (c = channel, n = frame):
  int32_t* idx;        // [2][n_frames] frame whose (f, Q) gave the coefficients in force, -1 = the quantum's entry set
};
// scratch: d_last, d_ent int32 [n_jobs][2][n_quanta]; the slab-transposed batch-wide streams (layout: biquad.cu header)
//   d_s1t float4 [ceil(n_jobs/16)][n_frames/32][32 frames][32 rows]  (x, a1, a2, -)
//   d_s2t float4 [ceil(n_jobs/16)][n_frames/32][32 frames][32 rows]  (b0, b1, b2, -)
// n_jobs <= 65535 per call.
// d_states (float2) and d_flags (int): scratch of the sizes biquad_scratch_sizes() reports (segment / slab states of the
// speculative recursion, verification flags, stream-layout flags; layout: biquad_lanes.cu)
int biquad_lane_segments(int n_jobs, int64_t n_frames, int* seg_slabs);
void biquad_scratch_sizes(int n_jobs, int64_t n_frames, size_t* n_float2, size_t* n_int);
void launch_biquad(const BiquadJob* d_jobs, int n_jobs, int64_t n_frames, int64_t n_quanta, int sample_rate, int32_t* d_last,
                   int32_t* d_ent, float4* d_s1t, float4* d_s2t, float2* d_states, int* d_flags, cudaStream_t s);

// ---- shared-coefficient path (voices of a class run the same filter automation over the same non-silent range: biquad.cu)
constexpr int kBqChunkFrames = 64;                                               // frames per class-stream chunk
constexpr int kBqChunkABytes = 2 * (kBqChunkFrames + 2) * 8;                     // (a1, a2) of both channels (channel stride 66: 16-byte aligned, 4 banks apart)
constexpr int kBqChunkBytes = kBqChunkABytes + 2 * (kBqChunkFrames + 1) * 16;    // + (b0, b1, b2, -): one chunk (3120 bytes)
// select / entry / RBJ for one representative job per class -> class streams d_cs[class * cs_stride + chunk * kBqChunkBytes]
// (n_frames multiple of 128; chunk = kBqChunkFrames frames).  d_last, d_ent: int32 [n_classes][2][n_quanta]; reps[c].idx: int32 [2][n_frames];
// d_wide_scratch: int [ceil(n_classes / 16)]
void launch_biquad_classes(const BiquadJob* d_reps, int n_classes, int64_t n_frames, int64_t n_quanta, int sample_rate, int32_t* d_last, int32_t* d_ent,
                           int* d_wide_scratch, unsigned char* d_cs, size_t cs_stride, cudaStream_t s);
void launch_biquad_zero_outside(const BiquadJob* d_jobs, int n_jobs, int64_t n_frames, cudaStream_t s);
// the recursion over groups of up to 16 jobs that all belong to one class; input rows 16-byte aligned, read in place.
// d_states / d_flags sized by biquad_shared_scratch_sizes
struct BqGroup {
  int first, count, cls;  // jobs [first, first + count) of the job array, class stream index
};
// warm_hint_slabs: warm-up of a speculative segment in 32-frame slabs (0 = the default of 256 = 8192 frames); the same value must be
// passed to all three.  Whatever the warm-up, the result is verified bit for bit and repaired where a segment did not re-join.
int biquad_shared_segments(int n_groups, int64_t n_frames, int* seg_chunks, int warm_hint_slabs = 0);
void biquad_shared_scratch_sizes(int n_groups, int64_t n_frames, size_t* n_float2, size_t* n_int, int warm_hint_slabs = 0);
// sequential: one segment per group (filters known to forget too slowly for the speculative segments); scratch sized the same way
void launch_biquad_lanes_shared(const BiquadJob* d_jobs, const BqGroup* d_groups, int n_groups, const unsigned char* d_cs, size_t cs_stride,
                                int64_t n_frames, float2* d_states, int* d_flags, bool sequential, cudaStream_t s, int warm_hint_slabs = 0);

// K3d alone (biquad_lanes.cu): the recursion over the slab-transposed streams, TMA-fed
void launch_biquad_lanes(const BiquadJob* d_jobs, int n_jobs, int64_t n_frames, const float4* d_s1t, const float4* d_s2t, float2* d_states,
                         int* d_flags, cudaStream_t s);

struct MixJob {       // dst[c][n] = (((0 + src0) + src1) + ...) over active ranges, AudioNodeInput.cs:118-137
  float* dst[2];
  int first_input;    // index into the MixInput array
  int n_inputs;
};
struct MixInput {
  const float* src[2];
  int64_t lo, hi;     // frames where the input is non-silent
  float downmix = 0.f;  // != 0: the fan-in has ONE channel and this input two: (L + R) * downmix lands in both rows
  int pad_ = 0;         //       (AudioNodeInput.cs:214-228; 1 / sqrt(2))
};
void launch_mix(const MixJob* d_jobs, int n_jobs, const MixInput* d_inputs, int64_t n_frames, cudaStream_t s);

// IR preparation (PartitionedConvolver.cs:93-102): scale[ch] from the channel's RMS
// channel c at d_base + c*stride; normalize == 0 writes 1.0
void launch_ir_scale(const float* d_base, int64_t stride, int n_channels, int64_t n_frames, int normalize, float calibration, float* d_scale,
                     cudaStream_t s);

// one impulse-response channel of a batched (deferred) preparation
struct IrChanJob {
  const float* ir;     // the channel's frames
  int64_t n_frames;
  int normalize;
  float* scale;        // the channel's normalisation scale (device scalar)
  float2* H;           // first-level spectra [P16][B]
  int P, P16;
  float2* H2;          // second-level spectra [B+1][row] or nullptr
};
void launch_ir_scale_batch(const IrChanJob* d_jobs, int n_jobs, float calibration, int B, cudaStream_t s);
// second-level spectra of a batch of channels that share the transform length M (<= 4096: radix-16 plan)
void launch_fft2_prep_batch(const IrChanJob* d_jobs, int n_jobs, int B, int M, const float2* d_tab16, cudaStream_t s);

// interleaved samples (fmt 0 s16, 1 s24, 2 s32, 3 f32) -> planar float32 rows d_out[c*stride + f], normalised like sf_readf_float
void launch_deinterleave(const void* d_raw, int fmt, int n_channels, int64_t n_frames, float* d_out, int64_t stride, cudaStream_t s);
// out[f*channels + c] = c == 0 ? c0[f] : c == 1 ? c1[f] : 0   (c1 may be null: one destination channel requested)
void launch_interleave(const float* c0, const float* c1, float* out, int64_t n_frames, int channels, cudaStream_t s);
void launch_fill_zero(void* p, size_t bytes, cudaStream_t s);
// d_dst <- h_src (page-locked, 16-byte aligned, bytes a multiple of 16) by a copy kernel instead of the DMA engine
void launch_copy_from_host(void* d_dst, const void* h_src, size_t bytes, cudaStream_t s);
constexpr int kCopySegs = 24;
struct CopySegments {
  void* dst[kCopySegs];
  const void* src[kCopySegs];
  size_t n16[kCopySegs];  // 16-byte units
  int n = 0;
};
void launch_copy_from_host_multi(const CopySegments& segs, cudaStream_t s);

}  // namespace gac
