// engine_stream.inl — included at the bottom of engine.cu.
//
// The literal plugin seam ("CudaConvolverNode : AudioNode"): one ConvolverNode processed quantum by quantum
// inside an ordinary reference graph, the way GraphAudio.SteamAudio's nodes P/Invoke their native effect once per block
// (GraphAudio.SteamAudio/Nodes/SteamAudioNodeBase.cs:50-135).  State lives on the device between calls:
//   X   float2 [n_in][hist + window][B]   linear history of the input spectra (the reference's FDL ring, PartitionedConvolver.cs:115-128,
//                                          unrolled: row r is the block fed r - hist calls ago ... no modulo in the MAC)
//   Y   float2 [n_conv][1 + chunk][B]     accumulated spectra; row 0 = the previous call's last block (its upper half is the overlap,
//                                          PartitionedConvolver.cs:146-150)
// Per call: H2D of the blocks, K5 (forward rFFT into the next rows), k_mac_ring (exact reference order), K7 (inverse + overlap-add),
// D2H.  It is latency-bound by design — the batched renderer (gac_render*) is the throughput path.  (The block copies run on the
// context stream's copy-engine queue: in a GAC_FLAG_ASYNC_UPLOAD context that orders later renders behind pending uploads, see
// gac_context::h_stage — use a context of its own for per-quantum work.)

struct gac_convolver {
  uint32_t magic = 0x47414343;  // "GACC"
  gac_context* ctx = nullptr;
  int B = 128;
  int P = 0;
  int n_in = 0, n_conv = 0, n_out = 0;
  int64_t hist = 0, window = 0, row = 0;  // rows [0, hist) start as the cleared delay line; `row` = where the next block goes
  int chunk = 0;                          // blocks per internal pass
  int inv_nb = -1;                        // block count the device copy of the inverse jobs was built for
  float2* d_X = nullptr;
  float2* d_Y = nullptr;
  float* d_in = nullptr;   // [n_in][chunk*B]
  float* d_out = nullptr;  // [n_out][(1 + chunk)*B]
  RingMacJob* d_mac = nullptr;
  FftInvJob* d_inv = nullptr;
  void* block = nullptr;  // the one allocation behind all of the above
};

static bool conv_ok(gac_convolver* c) { return c && c->magic == 0x47414343 && ctx_ok(c->ctx); }

static int convolver_clear(gac_convolver* c) {
  cudaStream_t st = c->ctx->stream;
  CU(cudaMemsetAsync(c->d_X, 0, sizeof(float2) * (size_t)c->n_in * (c->hist + c->window) * c->B, st));
  CU(cudaMemsetAsync(c->d_Y, 0, sizeof(float2) * (size_t)c->n_conv * (1 + c->chunk) * c->B, st));
  c->row = c->hist;
  return GAC_OK;
}

extern "C" int gac_convolver_create(gac_context* ctx, gac_ir* ir, gac_convolver** out) {
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or destroyed");
  if (!ir || !out) return fail(GAC_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  if (ir->ctx != ctx) return fail(GAC_ERR_INVALID_ARGUMENT, "the impulse response belongs to another context");
  CU(cudaSetDevice(ctx->device));
  int rc;
  if (!ir->prepared) {  // asynchronous-upload contexts defer the preparation to the first use: this is it
    wait_ready(ctx, ir->src);
    if ((rc = ir_prepare_run(ctx, ir->src->d, ir->src->stride, ir->nch, ir->frames, ir->normalize, ir))) return rc;
    buffer_unref(ir->src);
    ir->src = nullptr;
  }
  auto c = std::make_unique<gac_convolver>();
  c->ctx = ctx;
  c->B = ctx->B;
  c->P = ir->P;
  // routing: one convolver per IR channel on the matching input channel; 4 channels + true stereo = 2 in, 4 convolvers, 2 out
  // (Nodes/ConvolverNode.cs:58-77,127-151)
  c->n_in = ir->true_stereo ? 2 : ir->nch;
  c->n_conv = ir->nch;
  c->n_out = ir->true_stereo ? 2 : ir->nch;
  c->chunk = 32;
  c->hist = ir->P16;
  c->window = std::max<int64_t>(ir->P16, 256) + c->chunk;
  const int B = c->B;
  const size_t xb = sizeof(float2) * (size_t)c->n_in * (c->hist + c->window) * B;
  const size_t yb = sizeof(float2) * (size_t)c->n_conv * (1 + c->chunk) * B;
  const size_t ib = sizeof(float) * (size_t)c->n_in * c->chunk * B;
  const size_t ob = sizeof(float) * (size_t)c->n_out * (1 + c->chunk) * B;
  const size_t mb = (sizeof(RingMacJob) * c->n_conv + 15) & ~(size_t)15;
  const size_t vb = (sizeof(FftInvJob) * c->n_out + 15) & ~(size_t)15;
  CU(cudaMallocAsync(&c->block, xb + yb + ib + ob + mb + vb, ctx->stream));
  char* p = reinterpret_cast<char*>(c->block);
  c->d_X = reinterpret_cast<float2*>(p);
  p += xb;
  c->d_Y = reinterpret_cast<float2*>(p);
  p += yb;
  c->d_in = reinterpret_cast<float*>(p);
  p += ib;
  c->d_out = reinterpret_cast<float*>(p);
  p += ob;
  c->d_mac = reinterpret_cast<RingMacJob*>(p);
  p += mb;
  c->d_inv = reinterpret_cast<FftInvJob*>(p);
  const int64_t xrows = c->hist + c->window, yrows = 1 + c->chunk;
  std::vector<RingMacJob> mj(c->n_conv);
  for (int k = 0; k < c->n_conv; k++) {
    // true stereo: convolvers 0, 1 read the left input, 2, 3 the right one (ConvolverNode.cs:137-143)
    const int in_ch = ir->true_stereo ? k / 2 : k;
    mj[k].X = c->d_X + (size_t)in_ch * xrows * B;
    mj[k].H = ir->d_H + (size_t)k * ir->P16 * B;
    mj[k].Y = c->d_Y + ((size_t)k * yrows + 1) * B;
    mj[k].P = ir->P;
  }
  // (the inverse-transform jobs carry the block count of a call: sent by gac_convolver_process when it changes)
  if ((rc = table_h2d(ctx, c->d_mac, mj.data(), sizeof(RingMacJob) * mj.size())) == GAC_OK) rc = convolver_clear(c.get());
  if (rc == GAC_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(GAC_ERR_CUDA, "convolver setup failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (rc) {
    cudaFreeAsync(c->block, ctx->stream);
    return rc;
  }
  *out = c.release();
  return GAC_OK;
}

extern "C" int gac_convolver_reset(gac_convolver* c) {
  if (!conv_ok(c)) return fail(GAC_ERR_DISPOSED, "convolver is null or destroyed");
  CU(cudaSetDevice(c->ctx->device));
  return convolver_clear(c);
}

extern "C" int gac_convolver_destroy(gac_convolver* c) {
  if (!c || c->magic != 0x47414343) return fail(GAC_ERR_INVALID_ARGUMENT, "convolver is null or already destroyed");
  if (ctx_ok(c->ctx)) {
    cudaSetDevice(c->ctx->device);
    cudaFreeAsync(c->block, c->ctx->stream);
  }
  c->magic = 0;
  delete c;
  return GAC_OK;
}

extern "C" int gac_convolver_channels(gac_convolver* c, int* n_in, int* n_out) {
  if (!conv_ok(c)) return fail(GAC_ERR_DISPOSED, "convolver is null or destroyed");
  if (n_in) *n_in = c->n_in;
  if (n_out) *n_out = c->n_out;
  return GAC_OK;
}

// n_frames consecutive frames (a multiple of the partition) through ConvolverNode.Process (ConvolverNode.cs:102-155)
extern "C" int gac_convolver_process(gac_convolver* c, const float* const* in, int n_in_channels, float* const* out, int n_out_channels,
                                     int64_t n_frames) {
  if (!conv_ok(c)) return fail(GAC_ERR_DISPOSED, "convolver is null or destroyed");
  if (!in || !out) return fail(GAC_ERR_INVALID_ARGUMENT, "null argument");
  gac_context* ctx = c->ctx;
  const int B = c->B;
  if (n_in_channels != c->n_in)  // the node's input is Explicit with the IR's channel count (ConvolverNode.cs:67-76): the caller mixes to it
    return fail(GAC_ERR_INVALID_ARGUMENT, "the convolver takes %d input channel(s), got %d", c->n_in, n_in_channels);
  if (n_out_channels != c->n_out) return fail(GAC_ERR_INVALID_ARGUMENT, "the convolver produces %d output channel(s), got %d", c->n_out, n_out_channels);
  if (n_frames <= 0 || n_frames % B) return fail(GAC_ERR_OUT_OF_RANGE, "frame count must be a positive multiple of %d", B);
  for (int i = 0; i < n_in_channels; i++)
    if (!in[i]) return fail(GAC_ERR_INVALID_ARGUMENT, "input channel %d is null", i);
  for (int i = 0; i < n_out_channels; i++)
    if (!out[i]) return fail(GAC_ERR_INVALID_ARGUMENT, "output channel %d is null", i);
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int64_t xrows = c->hist + c->window, yrows = 1 + c->chunk;
  for (int64_t b0 = 0; b0 < n_frames / B; b0 += c->chunk) {
    const int nb = (int)std::min<int64_t>(c->chunk, n_frames / B - b0);
    if (c->row + nb > xrows) {
      // the linear history is full: its last P - 1 rows move to the front (window >= P16: source and destination are disjoint)
      const int64_t keep = c->hist;
      for (int i = 0; i < c->n_in; i++)
        CU(cudaMemcpyAsync(c->d_X + (size_t)i * xrows * B, c->d_X + ((size_t)i * xrows + (c->row - keep)) * B, sizeof(float2) * (size_t)keep * B,
                           cudaMemcpyDeviceToDevice, st));
      c->row = keep;
    }
    for (int i = 0; i < c->n_in; i++)
      CU(cudaMemcpyAsync(c->d_in + (size_t)i * c->chunk * B, in[i] + b0 * B, sizeof(float) * (size_t)nb * B, cudaMemcpyHostToDevice, st));
    FftFwdUniform u;  // x -> zero-padded 2B -> rFFT -> FDL slot (PartitionedConvolver.cs:106-124)
    u.in_base = c->d_in;
    u.in_stride = (int64_t)c->chunk * B;
    u.out_base = c->d_X + (size_t)c->row * B;
    u.out_stride = xrows * B;
    u.scale_base = nullptr;
    u.n_valid = (int64_t)nb * B;
    u.n_blocks = nb;
    launch_rfft_fwd_uniform(u, c->n_in, B, ctx->d_tw, st);
    launch_mac_ring(c->d_mac, c->n_conv, c->row, nb, B, st);  // :125,154-223
    if (nb != c->inv_nb) {
      std::vector<FftInvJob> vj(c->n_out);
      for (int o = 0; o < c->n_out; o++) {
        FftInvJob j;
        j.in = c->d_Y + (size_t)o * yrows * B;  // row 0 = the previous block: the inverse pass re-derives its upper half (the overlap)
        j.in2 = c->n_conv == 4 ? c->d_Y + (size_t)(o + 2) * yrows * B : nullptr;  // true stereo: L = c0 + c2, R = c1 + c3 (ConvolverNode.cs:137-143)
        j.out = c->d_out + (size_t)o * yrows * B;
        j.out2 = nullptr;
        j.n_blocks = 1 + nb;
        vj[o] = j;
      }
      CU(cudaMemcpyAsync(c->d_inv, vj.data(), sizeof(FftInvJob) * vj.size(), cudaMemcpyHostToDevice, st));  // pageable: staged before return
      c->inv_nb = nb;
    }
    launch_irfft_ola(c->d_inv, c->n_out, 1 + nb, B, ctx->d_tw, st);  // :134-150; block 0 of the pass is the previous call's last block
    CU(cudaGetLastError());
    for (int o = 0; o < c->n_out; o++)
      CU(cudaMemcpyAsync(out[o] + b0 * B, c->d_out + ((size_t)o * yrows + 1) * B, sizeof(float) * (size_t)nb * B, cudaMemcpyDeviceToHost, st));
    // the last block's spectrum becomes row 0 for the next call
    CU(cudaMemcpy2DAsync(c->d_Y, sizeof(float2) * (size_t)yrows * B, c->d_Y + (size_t)nb * B, sizeof(float2) * (size_t)yrows * B, sizeof(float2) * B,
                         (size_t)c->n_conv, cudaMemcpyDeviceToDevice, st));
    c->row += nb;
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return fail(GAC_ERR_CUDA, "convolver block failed: %s", cudaGetErrorString(e));
  return GAC_OK;
}

// one render quantum: exactly what CudaConvolverNode.Process() P/Invokes (cf. SteamAudioNodeBase.cs:74-135)
extern "C" int gac_convolver_process_block(gac_convolver* c, const float* const* in, int n_in_channels, float* const* out, int n_out_channels) {
  if (!conv_ok(c)) return fail(GAC_ERR_DISPOSED, "convolver is null or destroyed");
  return gac_convolver_process(c, in, n_in_channels, out, n_out_channels, c->B);
}
