/* graphaudio_cuda.h — C ABI of libgraphaudio_cuda.so
 *
 * B200-native (sm_100a) batched implementation of GraphAudio's offline graph-render hot path:
 * AudioBufferSourceNode (+CubicResampler) -> BiQuadFilterNode* -> GainNode -> ConvolverNode
 * (uniformly partitioned FFT convolution) -> fan-in bus -> destination, for many voices at once.
 *
 * This is the boundary a `GraphAudio.Cuda` C# package P/Invokes ([LibraryImport("graphaudio_cuda")],
 * cdecl — the idiom of GraphAudio.IO/Libsndfile.cs:36-68 and GraphAudio.Realtime/Miniaudio.cs:305-349);
 * see INTEGRATION.md for the binding.  Plain pointers and sizes only.
 *
 * Every function returns GAC_OK (0) or a negative gac_status; the message is available from
 * gac_last_error() (thread-local, owned by the library — cf. sf_strerror, Libsndfile.cs:48-56).
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with
 * GAC_ERR_NO_DEVICE.
 *
 * All reference citations are relative to /root/reference/GraphAudio.Core/.
 */
#ifndef GRAPHAUDIO_CUDA_H_
#define GRAPHAUDIO_CUDA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* v7.  Compatible additions since: gac_voice_desc.source_param is read for buffer sources too (PlaybackRate with events; a
 * zero-initialised field keeps the static playback_rate), looping sources at any effective rate. */
#define GAC_ABI_VERSION 7

/* ---- status codes; the C# layer maps them onto the exception types the reference throws ---- */
typedef enum gac_status {
  GAC_OK = 0,
  GAC_ERR_INVALID_ARGUMENT = -1,  /* ArgumentException            (OfflineAudioContext.cs:32-51)            */
  GAC_ERR_OUT_OF_RANGE = -2,      /* ArgumentOutOfRangeException  (OfflineAudioContext.cs:35-39)            */
  GAC_ERR_INVALID_OPERATION = -3, /* InvalidOperationException    (Nodes/ConvolverNode.cs:45-49)            */
  GAC_ERR_DISPOSED = -4,          /* ObjectDisposedException      (AudioContextBase.cs:54-55)               */
  GAC_ERR_NO_DEVICE = -5,         /* no CUDA device / wrong architecture: there is no CPU fallback          */
  GAC_ERR_CUDA = -6,              /* a CUDA runtime call failed (message carries cudaGetErrorString)        */
  GAC_ERR_OUT_OF_MEMORY = -7,
  GAC_ERR_NCCL = -8,
  GAC_ERR_UNSUPPORTED = -9        /* graph shape outside the accelerated path (SURVEY.md §8f "next")        */
} gac_status;

typedef struct gac_context gac_context; /* ≙ OfflineAudioContext                                  */
typedef struct gac_buffer gac_buffer;   /* ≙ PlayableAudioBuffer (device-resident copy)            */
typedef struct gac_ir gac_ir;           /* ≙ the PartitionedConvolver[] a ConvolverNode owns       */
typedef struct gac_graph gac_graph;     /* a flattened, immutable render graph                     */
typedef struct gac_convolver gac_convolver; /* one ConvolverNode processed quantum by quantum       */

/* ---- library ---- */
int gac_version(void);
const char* gac_last_error(void);
int gac_device_count(int* count);

/* ---- context ≙ OfflineAudioContext(int sampleRate = 48000)  (OfflineAudioContext.cs:18) ---- */
typedef struct gac_context_desc {
  int sample_rate; /* > 0                                                                            */
  int quantum;     /* AudioBuffer.FramesPerBlock; must be 128 (AudioBuffer.cs:10)                    */
  int partition;   /* PartitionedConvolver blockSize (PartitionedConvolver.cs:37): 128 (what
                      ConvolverNode uses, ConvolverNode.cs:55) or 256/512 (offline-only option;
                      same linear convolution, different rounding points).  0 = 128.                 */
  int device_id;   /* CUDA ordinal; -1 = current device                                              */
  int mac_variant; /* spectral MAC (K6) algorithm:
                      0 = default: fast convolution along block time (a second FFT over the partition
                          axis, csrc/fft2.cu) for impulse responses of >= 64 partitions, register-tiled
                          direct sum (packed FFMA2) below;
                      1 = streaming one-pass-per-quantum direct sum (reference op order, unfused: the
                          T=1 roofline contract of SURVEY.md §8d; bit-exact against the oracle);
                      2 = register-tiled direct sum, scalar FFMA;   4 = register-tiled, packed FFMA2;
                      3 = second-level FFT for every impulse response                                */
  int tile_blocks; /* output blocks per CTA of the tiled MAC: 32 (default, 0) or 64                  */
  int flags;       /* GAC_FLAG_*                                                                     */
  int reserved;
} gac_context_desc;

/* Asynchronous uploads: when set AND the arrays passed to gac_buffer_create are page-locked (cudaHostAlloc /
 * cudaHostRegister), the host->device copy is queued on a copy stream and gac_buffer_create returns at once; the
 * arrays must stay valid and unmodified until the next gac_render* / gac_synchronize on the context returns.  IR
 * preparation and the first voice batches of the render overlap the remaining copies.  Default (flag clear): the
 * reference's semantics — the data is copied during the call (PlayableAudioBuffer.cs:84-93). */
#define GAC_FLAG_ASYNC_UPLOAD 1
/* K6 (second-level FFT) runs one or more overlap-save segments of TWICE the transform length in front when that saves work (a 12 s
 * render at P = 750: one 4096- and one 2048-point segment instead of four 2048-point ones, -25 % transform points, K6 0.51 -> 0.46 ms
 * on the bench workload; the double-length spectra of an impulse response are prepared on first use).  This flag keeps every segment
 * at the length picked for the impulse response (A/B measurements, tests); results agree to 2e-6. */
#define GAC_FLAG_UNIFORM_SEGMENTS 2
/* Fan-in fusion (default ON): ConvolverNodes that are the last node of their chain and whose outputs meet in the same fan-in
 * (AudioNodeInput.MixBuffer, AudioNodeInput.cs:118-137: a bus head or the destination) are summed as second-level spectra, so
 * that a whole group of voices pays for ONE inverse transform pair and one fan-in input (csrc/fft2.cu, k_fft2_sum16).  The sum
 * is the same linear combination in another float32 rounding order (measured <= 3e-7 of the bus peak).  This flag keeps one
 * inverse transform and one fan-in input per voice: the reference's order of additions (A/B measurements, tests). */
#define GAC_FLAG_NO_FANIN_FUSION 4

int gac_context_create(const gac_context_desc* desc, gac_context** out);
int gac_context_destroy(gac_context* ctx);
/* Blocks until every queued upload and render of the context has finished. */
int gac_synchronize(gac_context* ctx);

/* ---- buffers ≙ PlayableAudioBuffer.FromChannelArrays (PlayableAudioBuffer.cs:122-143) ----
 * `channels` are caller-owned host arrays, copied (to HBM) during the call, as CopyToChannel
 * copies (PlayableAudioBuffer.cs:84-93).  1..32 channels, equal lengths. */
int gac_buffer_create(gac_context* ctx, const float* const* channels, int n_channels, int64_t n_frames,
                      int sample_rate, gac_buffer** out);

/* ≙ GraphAudio.IO's AudioDecoder.LoadFromStream (GraphAudio.IO/LibsndfileDecoder.cs:195-220) behind the container parser: the
 * file's INTERLEAVED samples are uploaded as they are and converted + de-interleaved on the device, with the normalisation
 * libsndfile's sf_readf_float applies (16-bit * 2^-15, 24-bit * 2^-23, 32-bit (float)x * 2^-31, float32 unchanged).  Little-endian. */
typedef enum gac_sample_format { GAC_SAMPLE_S16 = 0, GAC_SAMPLE_S24 = 1, GAC_SAMPLE_S32 = 2, GAC_SAMPLE_F32 = 3 } gac_sample_format;
int gac_buffer_create_interleaved(gac_context* ctx, const void* samples, int sample_format, int n_channels,
                                  int64_t n_frames, int sample_rate, gac_buffer** out);
int gac_buffer_destroy(gac_buffer* buf);

/* ---- ConvolverNode.Buffer = ir  (Nodes/ConvolverNode.cs:25-79 -> PartitionedConvolver ctor
 * PartitionedConvolver.cs:37-102): per-channel RMS normalisation, partition, zero-pad, rFFT.
 * Fails with GAC_ERR_INVALID_OPERATION if the buffer's rate differs from the context's (:48-49).
 * Supported IR channel counts on the accelerated path: 1, 2 (discrete) and 4 with true_stereo. */
int gac_ir_prepare(gac_context* ctx, const gac_buffer* buf, int normalize, int true_stereo, gac_ir** out);
int gac_ir_destroy(gac_ir* ir);

/* ---- automation ≙ AudioParam (AudioParam.cs) ----
 * gac_event is bit-compatible with the reference's private AutomationEvent struct
 * (AudioParam.cs:360-367: enum(int) Type; float Value; float Target; double Time; double TimeConstant). */
typedef enum gac_event_type {
  GAC_EVENT_SET_VALUE = 0,        /* SetValueAtTime                 AudioParam.cs:252 */
  GAC_EVENT_LINEAR_RAMP = 1,      /* LinearRampToValueAtTime        AudioParam.cs:266 */
  GAC_EVENT_EXPONENTIAL_RAMP = 2, /* ExponentialRampToValueAtTime   AudioParam.cs:280 */
  GAC_EVENT_SET_TARGET = 3,       /* SetTargetAtTime                AudioParam.cs:297 */
  GAC_EVENT_EPOCH = 4             /* not an AudioParam event: a marker that splits the list into EPOCHS for parameters that were edited
                                     between successive Render calls of one context (OfflineAudioContext.cs:55-100: the timeline
                                     continues; AudioParam.Value / scheduling calls made in between act from the next unprocessed
                                     quantum on).  value = the static value from then on, time_constant = index of the first
                                     quantum of the epoch (> the previous epoch's), time / target unused; the events behind the
                                     marker (up to the next one) are the parameter's whole event list during that epoch.  Frames of
                                     quantum q are evaluated with the (value, events) of the last epoch that starts at or before q */
} gac_event_type;

typedef struct gac_event {
  int32_t type;
  float value;
  float target;
  double time;
  double time_constant;
} gac_event;

/* A parameter as the render thread sees it: the static `_value` plus the time-sorted event list
 * exactly as AddEvent leaves it (stable upper-bound insert, AudioParam.cs:333-352).  Values must
 * already be clamped to the param's range (the reference clamps at schedule time, :254,268,282,299). */
typedef struct gac_param {
  float value;
  int32_t n_events;
  const gac_event* events;
  /* Modulation input ≙ AudioNode.Connect(AudioParam) (Nodes/AudioNode.cs:86-92): the parameter's own AudioNodeInput (Explicit, ONE
   * channel: AudioParam.cs:60-62) sums the connected outputs; while that block is non-silent the value is
   * clamp(intrinsic + modulation, min, max) (AudioParam.cs:114-166; a-rate: per frame, k-rate: frame 0 of the quantum).
   * mod_bus = 0: none; k > 0: the fan-in is bus k-1 of the graph, which must carry GAC_BUS_MONO_INPUT. */
  int32_t mod_bus;
  float min_value, max_value; /* AudioParam.MinValue / MaxValue; used with modulation only */
  int32_t reserved;
} gac_param;

/* ---- per-voice processing chain ---- */
typedef enum gac_op_kind {
  GAC_OP_BIQUAD = 1,   /* BiQuadFilterNode  Nodes/BiQuadFilterNode.cs:87-258 */
  GAC_OP_GAIN = 2,     /* GainNode          Nodes/GainNode.cs:29-61          */
  GAC_OP_CONVOLVER = 3,/* ConvolverNode     Nodes/ConvolverNode.cs:102-155   */
  GAC_OP_DELAY = 4,    /* DelayNode         Nodes/DelayNode.cs:43-149        */
  GAC_OP_PANNER = 5,   /* StereoPannerNode  Nodes/StereoPannerNode.cs:36-153 */
  GAC_OP_GATE = 7,     /* a CONNECTION that was made or removed between two Render calls (AudioNode.Connect / Disconnect posted to
                          the render thread, Nodes/AudioNode.cs:109-147, act from the next unprocessed quantum on): the signal
                          reaches the next node only from quantum `aux` on (filter_type 0) / only before quantum `aux`
                          (filter_type 1); outside that window the next node's input sees no connection, i.e. a cleared,
                          silent-flagged block (AudioNodeInput.cs:102-107)                                                    */
  GAC_OP_CHANNEL = 6   /* one output of a ChannelSplitterNode (Nodes/ChannelSplitterNode.cs): the signal becomes channel `aux` of
                          its input as a ONE-channel signal (silence if the input has no such channel).  Only valid as the first
                          op of a chain fed by a bus (the splitter's input is that bus)                                       */
} gac_op_kind;

typedef enum gac_filter_type { /* FilterType, Nodes/BiQuadFilterNode.cs:288-298 */
  GAC_FILTER_LOWPASS = 0,
  GAC_FILTER_HIGHPASS = 1,
  GAC_FILTER_BANDPASS = 2,
  GAC_FILTER_NOTCH = 3,
  GAC_FILTER_ALLPASS = 4,
  GAC_FILTER_PEAKING = 5,
  GAC_FILTER_LOWSHELF = 6,
  GAC_FILTER_HIGHSHELF = 7
} gac_filter_type;

typedef struct gac_op_desc {
  int32_t kind;        /* gac_op_kind                                                         */
  int32_t filter_type; /* BIQUAD: gac_filter_type.  GATE: 0 "from", 1 "until".  CONVOLVER: 0, or 1 = this op is a LATER
                          EPOCH of the ConvolverNode described by the preceding CONVOLVER op: ConvolverNode.Buffer was set
                          again between two Render calls, so from quantum `aux` on the node runs new PartitionedConvolvers
                          (cleared delay lines and overlap, Nodes/ConvolverNode.cs:51-77) built from this op's `ir`; the
                          epochs of a node must share the channel layout                                              */
  gac_param p0;        /* BIQUAD: Frequency (a-rate)   GAIN: Gain (a-rate)   DELAY: DelayTime in seconds
                          (a-rate)   PANNER: Pan (a-rate, -1 .. 1)                             */
  gac_param p1;        /* BIQUAD: Q (a-rate)                                                   */
  gac_param p2;        /* BIQUAD: Gain in dB (k-rate)                                          */
  const gac_ir* ir;    /* CONVOLVER: prepared impulse response; NULL ≙ ConvolverNode without a
                          Buffer, which outputs silence (ConvolverNode.cs:107-119)             */
  double aux;          /* GATE / later CONVOLVER epoch: the quantum index (above).
                          DELAY: maxDelayTime in seconds, (0, 10] (DelayNode.cs:22-29).  PANNER: index of the first quantum the
                          node processes (0; later for a node created between two Render calls): its ClampedMax input has
                          no upstream block to count channels from in that quantum (AudioNodeInput.cs:109,140-168).  Else 0 */
} gac_op_desc;

/* One voice = AudioBufferSourceNode -> ops[0] -> ops[1] -> ... -> bus (or destination).
 * Source semantics: Nodes/AudioBufferSourceNode.cs:79-143 (Start/Stop are block-granular),
 * :186-235 (rate == 1 copy path), :236-358 (CubicResampler path), :360-372 (final block dropped). */
typedef struct gac_voice_desc {
  const gac_buffer* source; /* 1 or 2 channels                                                 */
  double start_when;        /* Start(when, offset, duration); duration = +inf for "whole buffer" */
  double start_offset;
  double start_duration;
  double stop_when;         /* Stop(when); NaN = never called                                   */
  float playback_rate;      /* PlaybackRate.Value (k-rate).  A PlaybackRate with automation events, or one whose Value was edited
                               between Render calls (epochs), is passed in full as `source_param` (n_events > 0), which then takes
                               precedence: evaluated per quantum on the host (Nodes/AudioBufferSourceNode.cs:165-169), the path
                               (copy / CubicResampler) is chosen per quantum as the reference does.  With a modulation input
                               (source_param.mod_bus > 0) the k-rate values are evaluated on the device behind the modulator and
                               read back before the positions are replayed */
  int32_t n_ops;
  const gac_op_desc* ops;
  int32_t bus;              /* index into buses, or -1: connected straight to the destination   */
  int32_t input;            /* 0: the chain is fed by `source` (an AudioBufferSourceNode).
                               k > 0: the chain is fed by the OUTPUT of bus k-1 (fan-out of a mixed or
                               processed signal, e.g. the dry / wet branches of GraphAudio.Kit's
                               ReverbEffect, Effects/ReverbEffect.cs:63-91); `source` and the Start/Stop
                               fields are then ignored                                           */
  int32_t loop;             /* AudioBufferSourceNode.Loop (:40-44): at effective rate 1 the copy path (:186-235), otherwise the
                               wrap-buffer CubicResampler path (:236-358).  LoopStart >= LoopEnd: GAC_ERR_UNSUPPORTED at render  */
  int32_t source_kind;      /* gac_source_kind (input == 0 only): what feeds the chain                                */
  double loop_start;        /* LoopStart in seconds (:49-53)                                     */
  double loop_end;          /* LoopEnd in seconds, 0 = end of the buffer (:58-62)                */
  gac_param source_param;   /* CONSTANT: Offset (a-rate, Nodes/ConstantSourceNode.cs); OSCILLATOR: Frequency (a-rate, Hz);
                               BUFFER: PlaybackRate when n_events > 0 (k-rate; see playback_rate)                                 */
  int32_t oscillator_type;  /* OSCILLATOR: 0 sine, 1 square, 2 sawtooth, 3 triangle (Nodes/OscillatorNode.cs:207-213)            */
  int32_t reserved;
} gac_voice_desc;

/* Scheduled sources (IAudioScheduledSourceNode): one channel, sample-accurate start and stop inside a quantum
 * (Nodes/OscillatorNode.cs:97-118, Nodes/ConstantSourceNode.cs:82-110).  They use start_when (NaN: never started),
 * start_duration (NaN: none) and stop_when of the voice; `source`, start_offset, playback_rate and the loop fields are ignored.
 * The oscillator's phase is the running sum of 2*pi*f[n]/fs (:150-156); the device forms it as a blocked prefix sum in double,
 * which agrees with the reference's sample-by-sample sum to ~1e-10 rad: within 1e-7 of the reference's samples for sine,
 * sawtooth and triangle, while a square wave may switch one frame early or late at an edge. */
typedef enum gac_source_kind { GAC_SOURCE_BUFFER = 0, GAC_SOURCE_CONSTANT = 1, GAC_SOURCE_OSCILLATOR = 2 } gac_source_kind;

/* A bus = fan-in AudioNodeInput (AudioNodeInput.cs:100-138) followed by a chain of ops
 * (typically one GainNode), connected to the destination, to another bus, or read by further chains.
 * Inputs are summed in connection order, float32, skipping silent-flagged blocks (:121-132).
 * Voices, buses and chains fed by buses together express any acyclic graph of the supported nodes:
 * a node with several inputs is a bus, a node whose output feeds several nodes ends a bus (or is a
 * source, which may feed several voices). */
typedef struct gac_bus_desc {
  int32_t n_ops;
  const gac_op_desc* ops;
  int32_t target;        /* where the bus output is connected: 0 = the destination, k > 0 = an input of bus
                            k-1 (bus hierarchies, GraphAudio.Kit/AudioBus.cs:76-114), -1 = nowhere directly
                            (it is only consumed by chains with input = this bus + 1, or by parameters)    */
  int32_t n_inputs;      /* connection order at the bus fan-in (AudioNodeInput.cs:118-137): entries >= 0  */
  const int32_t* inputs; /* are bus indices, entries < 0 are ~voice_index.  NULL = the voices routed here
                            in index order, then the buses targeting this one in index order              */
  int32_t flags;         /* GAC_BUS_*                                                                      */
  int32_t reserved;
  const int32_t* input_slots; /* ChannelMergerNode (Nodes/ChannelMergerNode.cs): one entry per `inputs` entry, or NULL.
                            0 = an ordinary fan-in connection; 1 / 2 = the connection goes to merger input 0 / 1: channel 0
                            of what that input mixes becomes the bus's left / right channel.  (Mergers of more than two
                            inputs are outside the accelerated path.)                                      */
} gac_bus_desc;
/* the fan-in has ONE channel (Explicit 1: the input of an AudioParam, AudioParam.cs:60-62): mono inputs are added as they are,
 * stereo inputs as (L + R) / sqrt(2) (AudioNodeInput.cs:214-228) */
#define GAC_BUS_MONO_INPUT 1

typedef struct gac_graph_desc {
  int32_t n_voices;
  const gac_voice_desc* voices;
  int32_t n_buses;
  const gac_bus_desc* buses;
  /* Connection order at the destination's input: entries >= 0 are bus indices, entries < 0 are
   * ~voice_index for voices with bus == -1.  NULL = buses in index order, then direct voices. */
  int32_t n_dest_inputs;
  const int32_t* dest_inputs;
} gac_graph_desc;

int gac_graph_create(gac_context* ctx, const gac_graph_desc* desc, gac_graph** out);
int gac_graph_destroy(gac_graph* graph);

/* ---- render ≙ OfflineAudioContext.Render(float[][] output, int frameCount, int startIndex = 0)
 * (OfflineAudioContext.cs:30-102).  Renders frames [first_frame, first_frame + n_frames) of the
 * graph's timeline (first_frame lets a caller reproduce successive Render calls, :55-100) into
 * caller-allocated host arrays out_channels[c][start_index ...]; n_out_channels is 1 or 2 (the
 * destination is stereo, Nodes/AudioDestinationNode.cs:17).  Synchronous: returns after the
 * device->host copy. */
int gac_render(gac_context* ctx, const gac_graph* graph, int64_t first_frame, int64_t n_frames,
               float* const* out_channels, int n_out_channels, int64_t start_index);

/* As gac_render, but the result is written INTERLEAVED: interleaved[(start_index + f) * channels + c], channels in 1..32,
 * the destination's channels first and zeros in the others — what a caller gets from driving
 * AudioContextBase.ProcessBlockInterleaved (AudioContextBase.cs:88-161) block after block (the output side of a file writer
 * or a device callback).  The interleaving runs on the device; one device->host copy. */
int gac_render_interleaved(gac_context* ctx, const gac_graph* graph, int64_t first_frame, int64_t n_frames,
                           float* interleaved, int channels, int64_t start_index);

/* As gac_render, but the result stays in HBM: d_out is a device pointer to [n_out_channels][n_frames]
 * float32 (row stride n_frames).  The call returns when the device has finished (the job tables of a render live on the
 * host until then); `sync` is kept for ABI compatibility and ignored. */
int gac_render_device(gac_context* ctx, const gac_graph* graph, int64_t first_frame, int64_t n_frames,
                      float* d_out, int n_out_channels, int sync);

/* Batch of independent renders (BASELINE config 4: one OfflineAudioContext per render): graph g is
 * rendered into out_channels[g * n_out_channels + c].  All graphs share n_frames. */
int gac_render_batch(gac_context* ctx, const gac_graph* const* graphs, int n_graphs, int64_t n_frames,
                     float* const* out_channels, int n_out_channels);

/* ---- multi-GPU bus mix (one process per GPU) ----
 * The voices of one logical graph are sharded over ranks; each rank renders its shard with
 * gac_render_device(…) up to the bus fan-in, then a single ncclReduce(sum, float32) over
 * NVLink/NVSwitch delivers the bus to the root, which applies the bus ops and returns the result. */
#define GAC_NCCL_UNIQUE_ID_BYTES 128
int gac_comm_unique_id(void* id128);
int gac_comm_init(gac_context* ctx, const void* id128, int rank, int n_ranks);
int gac_comm_destroy(gac_context* ctx);
/* Collective render: every rank passes its shard graph (whose voices feed bus 0 / the destination
 * directly; bus ops must be identical on all ranks and are applied on the root AFTER the reduce).
 * Only the root (rank `root`) writes out_channels. */
int gac_render_sharded(gac_context* ctx, const gac_graph* shard, int64_t n_frames, int root,
                       float* const* out_channels, int n_out_channels);

/* ---- one process, several GPUs ----
 * The reference's caller is one process with one OfflineAudioContext (OfflineAudioContext.cs:18).  A gac_group is that context
 * spread over `n_devices` GPUs of the box: one member gac_context per device (gac_group_context), created from `desc` (its
 * device_id is ignored), and one NCCL communicator over the device set (ncclCommInitAll).  Buffers, impulse responses and the
 * shard graphs are created against the member that renders them; gac_group_render runs the members' shards concurrently (one
 * host thread per device), reduces the buses onto member 0 with the single ncclReduce of gac_render_sharded and fills
 * out_channels[c][start_index ...] with frames [first_frame, first_frame + n_frames) as gac_render does.  shards[i] belongs to
 * member i; a member without voices passes a graph that holds the bus only. */
typedef struct gac_group gac_group;
int gac_group_create(const gac_context_desc* desc, const int* device_ids, int n_devices, gac_group** out);
int gac_group_destroy(gac_group* group);
int gac_group_size(gac_group* group, int* n_members);
int gac_group_context(gac_group* group, int index, gac_context** member);
int gac_group_render(gac_group* group, const gac_graph* const* shards, int64_t first_frame, int64_t n_frames,
                     float* const* out_channels, int n_out_channels, int64_t start_index);

/* ---- statistics of the last render on this context ---- */
typedef struct gac_stats {
  double ms_total;      /* device time of the whole render (CUDA events on the context stream)      */
  double ms_source;     /* K1 source copy / cubic resample                                          */
  double ms_automation; /* K2 a-rate parameter evaluation                                           */
  double ms_biquad;     /* K3 coefficient tables + recursive lanes                                  */
  double ms_gain;       /* K4                                                                       */
  double ms_fft_fwd;    /* K5                                                                       */
  double ms_mac;        /* K6 spectral multiply-accumulate                                          */
  double ms_fft_inv;    /* K7 inverse FFT + overlap-add                                             */
  double ms_mix;        /* fan-in sums + bus ops                                                    */
  double ms_d2h;
  int64_t conv_units;          /* channel-convolver blocks processed (SURVEY.md §8d "unit")         */
  double algorithmic_bytes;    /* Σ units · (16·P·C + 8·C + 8·B), the T=1 contract                  */
  double mac_complex_macs;     /* complex MACs of the direct sum (after causal skipping)            */
  int64_t kernel_launches;     /* kernels launched by the last render                               */
  int64_t voices;
  int64_t frames;
  double mac_flops;            /* flops K6 actually issued (direct: 8/cMAC; second-level FFT: 2 FFTs + product per segment) */
  double mac_bytes_moved;      /* bytes the K6 variant in use has to move through HBM (X, H, Y once) */
  int32_t mac_variant_used;    /* 1 stream, 2/4 register-tiled, 3 second-level FFT (last convolver batch) */
  int32_t mac_big_segments;    /* variant 3: double-length overlap-save segments per channel-convolver in front (0: one length) */
  double ms_delay;             /* DelayNode gather                                                  */
  double ms_panner;            /* StereoPannerNode                                                  */
  double mac_h2_bytes_single;  /* bytes of ONE set of IR spectra over the channel-convolvers of the render: with the
                                  spectrograms read and written once, the compulsory traffic of K6   */
  int32_t fanin_groups;        /* fan-in groups whose convolvers were summed as spectra (GAC_FLAG_NO_FANIN_FUSION: 0)  */
  int32_t fanin_members;       /* convolver nodes in those groups                                                    */
} gac_stats;
int gac_get_stats(gac_context* ctx, gac_stats* out);

/* ---- kernel-level entry points (used by the parity tests; host pointers in, host pointers out) ----
 * x: [n_signals][n_blocks*B] float32 time-domain blocks -> spectra [n_signals][n_blocks][B] complex64,
 * packed: bin 0 = (DC.re, Nyquist.re), bins 1..B-1 = (re, im).          ≙ PartitionedConvolver.cs:106-124 */
int gac_rfft_fwd_batch(gac_context* ctx, const float* x, int n_signals, int64_t n_blocks, float* spectra);
/* Y[s][b][k] = Σ_p X[s][b-p][k]·H[s][p][k] (X[<0] = 0), packed layout as above.   ≙ :154-223      */
int gac_spectral_mac(gac_context* ctx, const float* X, const float* H, int n_signals, int64_t n_blocks,
                     int n_partitions, int variant, float* Y);
/* inverse rFFT of every block + overlap-add -> y [n_signals][n_blocks*B]           ≙ :134-150      */
int gac_irfft_ola_batch(gac_context* ctx, const float* Y, int n_signals, int64_t n_blocks, float* y);
/* whole PartitionedConvolver for n_signals independent (x, ir) pairs               ≙ :37-152       */
int gac_convolve_batch(gac_context* ctx, const float* x, int n_signals, int64_t n_frames, const float* ir,
                       int64_t ir_frames, int normalize, float* y);
/* a-rate evaluation of one parameter for frames [0, n_frames)                      ≙ AudioParam.cs:114-141 */
int gac_automation_eval(gac_context* ctx, const gac_param* param, int a_rate, int64_t n_frames, float* values);
/* CubicResampler over one input, one Process call                                  ≙ CubicResampler.cs:26-63 */
int gac_resample_cubic(gac_context* ctx, const float* in, int64_t n_in, double rate, int64_t n_out, float* out,
                       int64_t* produced, int64_t* consumed);

/* BiQuadFilterNode.Process (Nodes/BiQuadFilterNode.cs:87-258) for n_signals independent stereo signals through the production
 * biquad kernels: x, y = [n_signals][2][n_frames]; one Frequency (a-rate), Q (a-rate) and Gain (k-rate, dB) parameter per signal,
 * evaluated on the device as in a render; filter_types[n_signals] = gac_filter_type. */
int gac_biquad_batch(gac_context* ctx, const float* x, int n_signals, int64_t n_frames, const int* filter_types,
                     const gac_param* frequency, const gac_param* q, const gac_param* gain_db, float* y);
/* The fan-in sum of AudioNodeInput.Pull / MixBuffer (AudioNodeInput.cs:118-137,182-244) over n_inputs stereo blocks sequences:
 * inputs[i] -> [2][n_frames]; input i is non-silent on frames [lo[i], hi[i]) (whole quanta) and skipped elsewhere; summed in
 * the given (connection) order in float32.  downmix (may be NULL): downmix[i] != 0 mixes input i down to one channel first,
 * (L + R) * downmix[i] in both output rows — the N -> 1 rule of :214-228 with downmix = 1/sqrt(2).  out = [2][n_frames]. */
int gac_mix(gac_context* ctx, const float* const* inputs, const int64_t* lo, const int64_t* hi, const float* downmix,
            int n_inputs, int64_t n_frames, float* out);

/* Host-only planning hook (no device needed): how K6's second-level FFT covers n_blocks output blocks of an impulse response of
 * n_partitions partitions — *m = transform length (0: direct sum), *n_big = overlap-save segments of length 2m in front,
 * *n_small = segments of length m behind them (uniform != 0: the single-length plan of GAC_FLAG_UNIFORM_SEGMENTS). */
int gac_plan_segments(int64_t n_blocks, int n_partitions, int uniform, int* m, int* n_big, int* n_small);

/* ---- the literal plugin seam: ONE ConvolverNode inside an ordinary reference graph ----
 * A `CudaConvolverNode : AudioNode` overrides Process(), pins Inputs[0].Buffer and its pooled output block and calls
 * gac_convolver_process_block once per render quantum — the pattern of GraphAudio.SteamAudio's nodes
 * (GraphAudio.SteamAudio/Nodes/SteamAudioNodeBase.cs:50-135).  The delay line, the overlap and the IR spectra stay on the
 * device between calls.  Channel routing is ConvolverNode's (Nodes/ConvolverNode.cs:58-77,121-151): a mono IR takes 1 channel
 * and produces 1, a stereo IR 2 -> 2, a 4-channel IR prepared with true_stereo 2 -> 2 (L = c0(inL) + c2(inR),
 * R = c1(inL) + c3(inR)); the caller's AudioNodeInput mixes to that count (SetChannelCount / Explicit, AudioNodeInput.cs:41-58).
 * The arithmetic of one block is PartitionedConvolver.Process (PartitionedConvolver.cs:104-152) with the multiply-accumulate in
 * the reference's order (p ascending, unfused, :154-223); the FFTs are float32.
 * One call costs a few launches and two small copies: it serves mixed and realtime graphs, it is NOT the throughput path
 * (that is gac_render*).  `ir` must outlive the convolver. */
int gac_convolver_create(gac_context* ctx, gac_ir* ir, gac_convolver** out);
int gac_convolver_destroy(gac_convolver* conv);
int gac_convolver_channels(gac_convolver* conv, int* n_in, int* n_out);
/* clears the delay line and the overlap (≙ constructing the PartitionedConvolvers again, PartitionedConvolver.cs:37-63) */
int gac_convolver_reset(gac_convolver* conv);
/* one render quantum: in[c] / out[c] are blocks of `partition` frames (128 = AudioBuffer.FramesPerBlock)  ≙ ConvolverNode.Process */
int gac_convolver_process_block(gac_convolver* conv, const float* const* in, int n_in_channels, float* const* out,
                                int n_out_channels);
/* n_frames / partition consecutive quanta in one call (same result as calling process_block for each) */
int gac_convolver_process(gac_convolver* conv, const float* const* in, int n_in_channels, float* const* out,
                          int n_out_channels, int64_t n_frames);

#ifdef __cplusplus
}
#endif
#endif /* GRAPHAUDIO_CUDA_H_ */
