// engine_render.inl — included at the bottom of engine.cu.
// Source planning (AudioBufferSourceNode semantics), the render drivers, the NCCL bus reduce and the
// kernel-level test entry points.

// ------------------------------------------------------------------------------------------ source stage (K1)
// (struct ResampleTable lives in engine.cu: the tables are cached in the context across renders)

// Nodes/AudioBufferSourceNode.cs:131-376, all four paths (rate 1 / CubicResampler, with and without Loop).  Decides, on the host, which quanta the source emits
// (block-granular start/stop :137-143; final block dropped :360-368) and, for the CubicResampler path, replays
// the data-independent phase recurrence of CubicResampler.cs:40-60 exactly (double Pos, (int)Pos consumes).
static int plan_sources(RenderEnv& env, const std::vector<const VoiceH*>& voices, std::vector<Sig>& sigs) {
  gac_context* ctx = env.ctx;
  const double inc = (double)128 / (double)ctx->fs;
  const std::vector<double>& bt = ctx->bt->h;
  auto& cj = env.keep->make<SourceJob>();
  auto& rj = env.keep->make<ResampleJob>();
  auto& tables = ctx->resample_cache;  // the phase recurrence depends on (rate, offset, end, length) only: replayed once per context
  // tables evicted from the cache while this batch was planned: released (stream-ordered) when the function returns, i.e. behind
  // the resample launch that may still read them; the host vectors stay alive in `keep` (a queued copy may still read them)
  struct Evicted {
    gac_context* ctx;
    HostKeep* keep;
    std::vector<std::shared_ptr<ResampleTable>> v;
    void push_back(const std::shared_ptr<ResampleTable>& t) { v.push_back(t); }
    ~Evicted() {
      for (auto& t : v) {
        if (t->d_k) cudaFreeAsync(t->d_k, ctx->stream);
        if (t->d_t) cudaFreeAsync(t->d_t, ctx->stream);
        if (t->d_x) cudaFreeAsync(t->d_x, ctx->stream);
        t->d_k = nullptr;
        t->d_t = nullptr;
        t->d_x = nullptr;
        keep->items.push_back(t);
      }
    }
  } evicted{ctx, env.keep, {}};
  const double inf = std::numeric_limits<double>::infinity();
  auto make_room = [&]() {
    if (tables.size() >= kResampleCacheMax) {
      // bounded: drop everything.  Jobs already planned for THIS batch still point at the device tables, so the frees are
      // queued behind launch_resample (`evicted`), never here.
      for (auto& kv : tables) evicted.push_back(kv.second);
      tables.clear();
    }
  };
  // context-owned device copies (not render scratch): later renders of the same source geometry reuse them
  auto upload_table = [&](ResampleTable& tab) -> int {
    if (tab.k.empty()) return GAC_OK;
    const size_t tb = (tab.k.size() * sizeof(int32_t) + 15) & ~(size_t)15;
    CU(cudaMallocAsync(&tab.d_k, tb, ctx->stream));
    CU(cudaMallocAsync(&tab.d_t, tb, ctx->stream));
    int rc;
    if ((rc = table_h2d(ctx, tab.d_k, tab.k.data(), tab.k.size() * sizeof(int32_t)))) return rc;
    if ((rc = table_h2d(ctx, tab.d_t, tab.t.data(), tab.t.size() * sizeof(float)))) return rc;
    if (!tab.x.empty()) {
      CU(cudaMallocAsync(&tab.d_x, tab.x.size() * sizeof(int32_t), ctx->stream));
      if ((rc = table_h2d(ctx, tab.d_x, tab.x.data(), tab.x.size() * sizeof(int32_t)))) return rc;
    }
    return GAC_OK;
  };

  for (size_t i = 0; i < voices.size(); i++) {
    const VoiceH& v = *voices[i];
    Sig& s = sigs[i];
    const gac_buffer* buf = v.src;
    wait_ready(ctx, buf);  // asynchronous upload still in flight: order this voice batch behind it
    s.lo = s.hi = 0;
    s.from_source = true;
    s.ch = buf->nch;  // the source emits a block with the buffer's channel count (AudioBufferSourceNode.cs:157-163)
    const float* src0 = buf->d;
    const float* src1 = buf->nch > 1 ? buf->d + buf->stride : buf->d;  // mono: 1 -> 2 up-mix copies the channel (AudioNodeInput.cs:201-213)
    SourceJob job{};
    job.src[0] = src0;
    job.src[1] = src1;
    job.dst[0] = s.p[0];
    job.dst[1] = s.p[1];
    job.pos0 = 0;
    job.out0 = 0;
    job.n_emit = 0;
    if (std::isnan(v.when)) {  // Start() never called: ProduceSilence forever
      cj.push_back(job);
      continue;
    }
    const double startTime = std::max(0.0, v.when);            // :93
    const double offset = std::max(0.0, v.offset);             // :94
    const int64_t pos = (int64_t)(offset * (double)buf->rate);  // :96
    double stopTime = std::numeric_limits<double>::quiet_NaN();
    if (!std::isinf(v.duration) && v.duration >= 0) stopTime = startTime + v.duration;  // :106-110 (and Stop() is then ignored, :120-121)
    else if (!std::isnan(v.stop_when)) stopTime = std::max(0.0, v.stop_when);           // :123-125
    // first playing block: t1 = t0 + 128/fs > startTime ; playing while t0 < stopTime  (:133-143)
    int64_t b_start = 0;
    {
      int64_t lo = 0, hi = env.NQ;
      while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (bt[mid] + inc > startTime) hi = mid; else lo = mid + 1;
      }
      b_start = lo;
    }
    int64_t b_stop = env.NQ;
    if (!std::isnan(stopTime)) {
      int64_t lo = b_start, hi = env.NQ;
      while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (!(bt[mid] < stopTime)) hi = mid; else lo = mid + 1;
      }
      b_stop = lo;
    }
    int64_t durEnd = v.duration < inf ? (int64_t)(offset * (double)buf->rate) + (int64_t)(v.duration * (double)buf->rate) : buf->n;  // :179-182
    durEnd = std::min(durEnd, buf->n);
    const double ratio = (double)buf->rate / (double)ctx->fs;  // :168
    const double eff = ratio * (double)v.rate;                 // :169
    const int64_t max_blocks = std::max<int64_t>(0, b_stop - b_start);
    if (v.rate_events || (v.loop && eff != 1.0)) {
      // The general case: a looping source on the CubicResampler path (:236-358 with _loop set) and / or a PlaybackRate that changes from
      // quantum to quantum (k-rate automation, evaluated on the host: host_param_value).  Which buffer frames are shifted into the
      // resampler and the phase of every output depend on positions only, so Process() is replayed here quantum by quantum WITH FRAME
      // INDICES — the copy path (:186-235) when the quantum's effective rate is exactly 1, else the resampler path, whose Process calls
      // are fed from the 512-float wrap buffer when looping (:296-314: `pos + available` IS loopEndFrame, so the test against
      // loopEndFrame - 4 always holds; frames pos .. loopEnd-1, then ONE pass over the loop region, cut at min(128 - outIdx + 4, 512)).
      // The device evaluates the float32 polynomial: a window of four consecutive frames is `k`, any other window (loop seams, buffer
      // edges) goes to `x`, a copied frame is the window around it with t = 0 (S1 + 0 * (...) == S1), a frame the reference clears
      // is kResampleCleared.  A quantum without output — or, not looping, after which every input frame is consumed — ends the source
      // (:360-368).
      int64_t loopEnd = 0, loopStart = 0;
      if (v.loop) {
        loopEnd = v.loop_end > 0 ? (int64_t)(v.loop_end * (double)buf->rate) : buf->n;  // :171-177
        loopEnd = std::min(loopEnd, buf->n);
        loopStart = std::min((int64_t)(v.loop_start * (double)buf->rate), loopEnd);
        if (loopEnd - loopStart <= 0) return fail(GAC_ERR_UNSUPPORTED, "looping source with an empty loop region");
      }
      if (max_blocks == 0 || buf->n <= 0) {
        cj.push_back(job);
        continue;
      }
      const int64_t n_out_max = max_blocks * 128;
      // a modulated PlaybackRate (AudioNode.Connect(AudioParam), AudioParam.cs:144-158: clamp(intrinsic + the modulator's first frame of
      // the quantum) while the modulator's block is non-silent): the k-rate table is evaluated on the device — the voice was scheduled
      // behind the bus that carries the modulator — and read back; the one point where a render waits for the device in its middle
      std::vector<float> rate_q;
      if (v.rate_events && v.src_param.mod_bus >= 0) {
        std::vector<ParamJob> pj;
        float* d_rate = nullptr;
        int rc = param_table(env, v.src_param, false, pj, &d_rate, s.bus_base);
        if (rc) return rc;
        if ((rc = run_param_jobs(env, pj))) return rc;
        rate_q.resize((size_t)env.NQ);
        CU(cudaMemcpyAsync(rate_q.data(), d_rate, sizeof(float) * (size_t)env.NQ, cudaMemcpyDeviceToHost, kstream(ctx)));
        CU(cudaStreamSynchronize(ctx->stream));
      }
      std::shared_ptr<ResampleTable> tab;
      auto key = std::make_tuple(eff, pos, loopEnd, n_out_max, loopStart);
      if (!v.rate_events) {
        auto it = tables.find(key);
        if (it != tables.end()) tab = it->second;
      }
      if (!tab) {
        tab = std::make_shared<ResampleTable>();
        tab->k.reserve((size_t)n_out_max);
        tab->t.reserve((size_t)n_out_max);
        const bool loop = v.loop;
        const int64_t len = buf->n, loopLen = loopEnd - loopStart;
        int64_t win[4] = {0, 0, 0, 0};
        int ready = 0;
        double Pos = 0.0;
        int64_t position = pos, active = 0;
        std::vector<int64_t> wrap;
        wrap.reserve(512);
        auto shift = [&](int64_t f) { win[0] = win[1]; win[1] = win[2]; win[2] = win[3]; win[3] = f; };
        auto emit = [&](const int64_t* w, float t) {
          if (w[1] == w[0] + 1 && w[2] == w[0] + 2 && w[3] == w[0] + 3) {
            tab->k.push_back((int32_t)w[0]);
          } else {
            tab->k.push_back(-(int32_t)(tab->x.size() / 4) - 1);
            for (int q = 0; q < 4; q++) tab->x.push_back((int32_t)w[q]);
          }
          tab->t.push_back(t);
        };
        for (int64_t blk = 0; blk < max_blocks; blk++) {
          double eff_b = eff;
          if (v.rate_events)  // :165-169
            eff_b = ratio * (double)(rate_q.empty() ? host_param_value(v.src_param, b_start + blk, bt[b_start + blk]) : rate_q[(size_t)(b_start + blk)]);
          int oi = 0;
          bool more = false;
          if (eff_b == 1.0) {  // :186-235
            int64_t p = position;
            while (oi < 128) {
              if (loop && p >= loopEnd) p = loopStart;
              if (p >= durEnd && !loop) break;
              const int64_t endFrame = loop ? loopEnd : std::min(durEnd, len);
              const int64_t avail = std::min<int64_t>(endFrame - p, 128 - oi);
              if (avail <= 0) break;
              for (int64_t q = 0; q < avail; q++) {
                const int64_t f = p + q;
                const int64_t w[4] = {std::max<int64_t>(f - 1, 0), f, std::min(f + 1, len - 1), std::min(f + 2, len - 1)};
                emit(w, 0.f);
              }
              p += avail;
              oi += (int)avail;
              more = true;
            }
            position += 128;  // :224
          } else {             // :236-358
            int64_t p = position, consumedCh = 0;
            while (oi < 128) {
              if (loop && p >= loopEnd) p = loopStart;                             // :265-268
              if (p >= durEnd && !loop) break;                                      // :270-274
              const int64_t endFrame = loop ? loopEnd : std::min(durEnd, len);
              const int64_t avail = std::min(endFrame - p, len - p);                // :277
              if (avail <= 0) break;  // (looping: only an empty loop region gets here, refused above)
              int64_t n_in = avail;
              if (loop) {
                const int64_t needed = std::min<int64_t>(128 - oi + 4, 512);        // :301
                wrap.clear();
                for (int64_t q = 0; q < loopEnd - p && (int64_t)wrap.size() < needed; q++) wrap.push_back(p + q);          // :303-306
                for (int64_t q = 0; (int64_t)wrap.size() < needed && q < loopLen; q++) wrap.push_back(loopStart + q);      // :308-311
                n_in = (int64_t)wrap.size();
              }
              auto in_frame = [&](int64_t i) { return loop ? wrap[(size_t)i] : p + i; };
              int64_t ip = 0;
              int op = 0;
              while (ready < 4 && ip < n_in) {   // CubicResampler.cs:31-35
                shift(in_frame(ip++));
                ready++;
              }
              if (ready == 4) {
                while (oi + op < 128) {          // :40-60
                  const int consume = (int)Pos;
                  if (ip + consume > n_in) break;
                  for (int q = 0; q < consume; q++) shift(in_frame(ip++));
                  Pos -= consume;
                  emit(win, (float)Pos);
                  op++;
                  Pos += eff_b;
                }
              }
              more = more || op > 0;
              int64_t np = p + ip;
              if (loop && np >= loopEnd) np = loopStart + (np - loopEnd);                                  // :323-328 (no modulo here)
              consumedCh += (np >= p) ? (np - p) : (loopEnd - p + np - loopStart);                         // :330
              p = np;
              oi += op;
              if (ip == 0 && op == 0) break;                                                               // :334-338
            }
            position += consumedCh;  // :347
          }
          for (; oi < 128; oi++) {  // the cleared rest of the quantum
            tab->k.push_back(kResampleCleared);
            tab->t.push_back(0.f);
          }
          if (loop && position >= loopEnd) position = loopStart + (position - loopEnd) % loopLen;          // :226-234, :349-357
          if (!more || (!loop && position >= durEnd)) break;  // the quantum is cleared and the source ends (:360-368)
          active = blk + 1;
        }
        tab->n_active_blocks = active;
        tab->k.resize((size_t)(active * 128));
        tab->t.resize(tab->k.size());
        tab->n_zero_from = (int64_t)tab->k.size();
        int rc = upload_table(*tab);
        if (rc) return rc;
        if (v.rate_events) {
          evicted.push_back(tab);  // not cached (the key would be the whole event list): released behind the resample launch
        } else {
          make_room();
          tables[key] = tab;
        }
      }
      if (tab->n_active_blocks == 0) {
        cj.push_back(job);
        continue;
      }
      ResampleJob r{};
      r.src[0] = src0;  // the table holds absolute frame indices
      r.src[1] = src1;
      r.dst[0] = s.p[0];
      r.dst[1] = s.p[1];
      r.k = tab->d_k;
      r.t = tab->d_t;
      r.x = tab->d_x;
      r.out0 = b_start * 128;
      r.n_emit = tab->n_active_blocks * 128;
      r.n_zero_from = tab->n_zero_from;
      s.lo = r.out0;
      s.hi = r.out0 + r.n_emit;
      rj.push_back(r);
      continue;
    }
    if (v.loop) {
      // Looping playback (:171-177, :197-234): runs until the stop time, duration only sets that stop time (:106-110).
      int64_t loopEnd = v.loop_end > 0 ? (int64_t)(v.loop_end * (double)buf->rate) : buf->n;
      loopEnd = std::min(loopEnd, buf->n);
      const int64_t loopStart = std::min((int64_t)(v.loop_start * (double)buf->rate), loopEnd);
      if (loopEnd - loopStart <= 0) return fail(GAC_ERR_UNSUPPORTED, "looping source with an empty loop region");
      job.pos0 = pos;
      job.out0 = b_start * 128;
      job.n_emit = max_blocks * 128;
      job.loop_start = loopStart;
      job.loop_end = loopEnd;
      job.loop_len = loopEnd - loopStart;
      s.lo = job.out0;
      s.hi = job.out0 + job.n_emit;
      cj.push_back(job);
      continue;
    }
    if (max_blocks == 0 || pos >= durEnd) {
      cj.push_back(job);
      continue;
    }
    if (eff == 1.0) {
      // block k (k-th playing block) is emitted iff pos + 128(k+1) < durEnd (:224,:360), and is then a full block
      int64_t n_data = (durEnd - pos + 127) / 128 - 1;
      int64_t nb = std::max<int64_t>(0, std::min(n_data, max_blocks));
      job.pos0 = pos;
      job.out0 = b_start * 128;
      job.n_emit = nb * 128;
      s.lo = job.out0;
      s.hi = job.out0 + job.n_emit;
      // A convolver (second-level-FFT path) that follows directly, or behind a GainNode that is fused into its forward
      // transform, reads the source buffer in place: K5 only touches frames inside [lo, hi), which map into the buffer.
      const std::vector<OpH>& ops = v.ops;
      const OpH* conv = nullptr;
      if (!ops.empty() && ops[0].kind == GAC_OP_CONVOLVER) conv = &ops[0];
      else if (ops.size() > 1 && ops[0].kind == GAC_OP_GAIN && ops[1].kind == GAC_OP_CONVOLVER) conv = &ops[1];
      const bool in_place = conv && conv->ir && conv->ir->d_H2 && nb > 0 && ((job.pos0 - job.out0) % 2 == 0) &&
                            (reinterpret_cast<uintptr_t>(src0) % 8 == 0) && (reinterpret_cast<uintptr_t>(src1) % 8 == 0);
      // ... and so does a BiQuadFilterNode that follows directly: k_biquad_resolve reads scalars (no alignment demands) inside
      // [lo, hi) only and clears the signal's rows outside it
      const bool in_place_biquad = !ops.empty() && ops[0].kind == GAC_OP_BIQUAD && nb > 0;
      if (in_place || in_place_biquad) {
        s.lazy[0] = src0 + (job.pos0 - job.out0);
        s.lazy[1] = src1 + (job.pos0 - job.out0);
      } else {
        cj.push_back(job);
      }
    } else {
      const int64_t avail = durEnd - pos;
      const int64_t n_out_max = max_blocks * 128;
      auto key = std::make_tuple(eff, pos, durEnd, n_out_max, (int64_t)-1);
      auto it = tables.find(key);
      std::shared_ptr<ResampleTable> tab;
      if (it != tables.end()) {
        tab = it->second;
      } else {
        make_room();
        tab = std::make_shared<ResampleTable>();
        if (avail >= 4) {
          tab->k.reserve((size_t)n_out_max);
          tab->t.reserve((size_t)n_out_max);
          int64_t inPos = 4;  // priming: four Shift()s (CubicResampler.cs:31-35)
          double Pos = 0.0;
          int64_t m = 0, M = -1;
          int64_t active = 0;
          for (int64_t blk = 0; blk < max_blocks; blk++) {
            int64_t produced = 0;
            for (int i = 0; i < 128; i++) {
              int consume = (int)Pos;                  // :42
              if (inPos + consume > avail) { M = m; break; }  // :43-44
              inPos += consume;
              Pos -= consume;                          // :49
              tab->k.push_back((int32_t)(inPos - 4));
              tab->t.push_back((float)Pos);            // :51
              Pos += eff;                              // :59
              m++;
              produced++;
            }
            // :360-368: a block that produced nothing, or after which every input frame is consumed, is cleared
            if (produced == 0 || pos + inPos >= durEnd) break;
            active = blk + 1;
            if (M >= 0) break;  // stalled inside a kept block: the next block produces nothing
          }
          tab->n_active_blocks = active;
          tab->n_zero_from = (M >= 0) ? M : m;
          tab->k.resize((size_t)std::min<int64_t>(m, active * 128));
          tab->t.resize(tab->k.size());
          int rc = upload_table(*tab);
          if (rc) return rc;
        }
        tables[key] = tab;
      }
      if (tab->n_active_blocks == 0) {
        cj.push_back(job);
        continue;
      }
      ResampleJob r{};
      r.src[0] = src0 + pos;
      r.src[1] = src1 + pos;
      r.dst[0] = s.p[0];
      r.dst[1] = s.p[1];
      r.k = tab->d_k;
      r.t = tab->d_t;
      r.out0 = b_start * 128;
      r.n_emit = tab->n_active_blocks * 128;
      r.n_zero_from = std::min<int64_t>(tab->n_zero_from, (int64_t)tab->k.size());
      s.lo = r.out0;
      s.hi = r.out0 + r.n_emit;
      rj.push_back(r);
    }
  }
  int t = env.timer->begin(C_SOURCE);
  if (!cj.empty()) {
    SourceJob* d = nullptr;
    int rc = env.scratch->upload(&d, cj);
    if (rc) return rc;
    launch_source_copy(d, (int)cj.size(), env.Npad, kstream(ctx));
    env.launches++;
  }
  if (!rj.empty()) {
    ResampleJob* d = nullptr;
    int rc = env.scratch->upload(&d, rj);
    if (rc) return rc;
    launch_resample(d, (int)rj.size(), env.Npad, kstream(ctx));
    env.launches++;
  }
  env.timer->end(t);
  CU(cudaGetLastError());
  return GAC_OK;
}

// Scheduled one-channel sources: ConstantSourceNode (Nodes/ConstantSourceNode.cs:75-141) and OscillatorNode
// (Nodes/OscillatorNode.cs:91-160).  The host decides the quanta the node plays in and its sample-accurate first / last frame from
// the accumulated block times, exactly as Process() does block by block; the device fills the rows.
static int plan_scheduled(RenderEnv& env, const std::vector<const VoiceH*>& voices, std::vector<Sig>& sigs) {
  gac_context* ctx = env.ctx;
  const double inc = (double)128 / (double)ctx->fs;
  const std::vector<double>& bt = ctx->bt->h;
  auto& jobs = env.keep->make<SchedJob>();
  std::vector<ParamJob> pj;
  bool any_osc = false;
  for (size_t i = 0; i < voices.size(); i++) {
    const VoiceH& v = *voices[i];
    Sig& s = sigs[i];
    s.lo = s.hi = 0;
    s.ch = 1;  // the node rents a one-channel block (:77-80)
    s.from_source = false;
    if (std::isnan(v.when)) continue;  // never started: cleared output, flagged silent
    const double startTime = std::max(0.0, v.when);  // Start(): _startTime = Math.Max(0, when)
    double stopTime = std::numeric_limits<double>::quiet_NaN();
    if (!std::isnan(v.duration) && !std::isinf(v.duration) && v.duration >= 0) stopTime = startTime + v.duration;  // (and Stop() is then ignored)
    else if (!std::isnan(v.stop_when)) stopTime = std::max(0.0, v.stop_when);
    // shouldPlay: t1 > _startTime && (NaN(_stopTime) || t0 < _stopTime), t1 = t0 + 128 / fs
    int64_t b_start = 0;
    {
      int64_t lo = 0, hi = env.NQ;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (bt[mid] + inc > startTime) hi = mid; else lo = mid + 1;
      }
      b_start = lo;
    }
    int64_t b_stop = env.NQ;
    if (!std::isnan(stopTime)) {
      int64_t lo = b_start, hi = env.NQ;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (!(bt[mid] < stopTime)) hi = mid; else lo = mid + 1;
      }
      b_stop = lo;
    }
    if (b_stop <= b_start) continue;
    int64_t s0 = b_start * 128, s1 = b_stop * 128;
    {
      const double t0 = bt[b_start], t1 = t0 + inc;
      if (t0 < startTime && startTime < t1) {  // startFrame = clamp(ceil((start - t0) * fs), 0, 128)
        double f = std::ceil((startTime - t0) * (double)ctx->fs);
        f = f < 0 ? 0 : (f > 128 ? 128 : f);
        s0 += (int64_t)(int)f;
      }
    }
    if (!std::isnan(stopTime)) {
      const double t0 = bt[b_stop - 1], t1 = t0 + inc;
      if (t0 < stopTime && stopTime < t1) {  // endFrame = clamp(floor((stop - t0) * fs), 0, 128)
        double f = std::floor((stopTime - t0) * (double)ctx->fs);
        f = f < 0 ? 0 : (f > 128 ? 128 : f);
        s1 = (b_stop - 1) * 128 + (int64_t)(int)f;
      }
    }
    if (s1 < s0) s1 = s0;  // (a stop inside the start quantum, before the start frame)
    float* tab = nullptr;
    int rc = param_table(env, v.src_param, true, pj, &tab, s.bus_base);
    if (rc) return rc;
    SchedJob j{};
    j.dst[0] = s.p[0];
    j.dst[1] = s.p[1];
    j.table = tab;
    j.value = v.src_param.value;
    j.lo = b_start * 128;
    j.hi = b_stop * 128;
    j.s0 = s0;
    j.s1 = s1;
    j.osc_type = v.kind == GAC_SOURCE_OSCILLATOR ? v.osc_type : -1;
    j.chunk_sum = nullptr;
    if (v.kind == GAC_SOURCE_OSCILLATOR) {
      any_osc = true;
      if ((rc = env.scratch->alloc(&j.chunk_sum, (size_t)((s1 - s0) / 1024 + 2)))) return rc;
    }
    jobs.push_back(j);
    s.lo = j.lo;  // every quantum the node plays in is marked non-silent, also its zero-filled frames (:158-159)
    s.hi = j.hi;
  }
  int rc = run_param_jobs(env, pj);
  if (rc) return rc;
  if (jobs.empty()) return GAC_OK;
  SchedJob* dj = nullptr;
  if ((rc = env.scratch->upload(&dj, jobs))) return rc;
  int t = env.timer->begin(C_SOURCE);
  launch_scheduled_sources(dj, (int)jobs.size(), env.Npad, ctx->fs, any_osc, kstream(ctx));
  env.timer->end(t);
  env.launches += any_osc ? 3 : 1;
  CU(cudaGetLastError());
  return GAC_OK;
}

// ------------------------------------------------------------------------------------------ NCCL (dlopen'ed)
struct NcclUid {
  char internal[128];
};
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUid*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*CommInitAll)(void**, int, const int*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
};
static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (api.lib) {
      api.GetUniqueId = (int (*)(NcclUid*))dlsym(api.lib, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(void**, int, NcclUid, int))dlsym(api.lib, "ncclCommInitRank");
      api.Reduce = (int (*)(const void*, void*, size_t, int, int, int, void*, cudaStream_t))dlsym(api.lib, "ncclReduce");
      api.CommDestroy = (int (*)(void*))dlsym(api.lib, "ncclCommDestroy");
      api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
      api.CommInitAll = (int (*)(void**, int, const int*))dlsym(api.lib, "ncclCommInitAll");
      api.GroupStart = (int (*)())dlsym(api.lib, "ncclGroupStart");
      api.GroupEnd = (int (*)())dlsym(api.lib, "ncclGroupEnd");
    }
  }
  if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.Reduce || !api.CommDestroy) return nullptr;
  return &api;
}
static int nccl_fail(NcclApi* a, int r, const char* what) {
  return fail(GAC_ERR_NCCL, "%s failed: %s", what, (a && a->GetErrorString) ? a->GetErrorString(r) : "nccl error");
}

extern "C" int gac_comm_unique_id(void* id128) {
  if (!id128) return fail(GAC_ERR_INVALID_ARGUMENT, "id is null");
  NcclApi* a = nccl_api();
  if (!a) return fail(GAC_ERR_NCCL, "libnccl.so.2 could not be loaded");
  NcclUid uid;
  int r = a->GetUniqueId(&uid);
  if (r != 0) return nccl_fail(a, r, "ncclGetUniqueId");
  memcpy(id128, uid.internal, 128);
  return GAC_OK;
}
extern "C" int gac_comm_init(gac_context* ctx, const void* id128, int rank, int n_ranks) {
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or destroyed");
  if (!id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(GAC_ERR_INVALID_ARGUMENT, "bad communicator arguments");
  NcclApi* a = nccl_api();
  if (!a) return fail(GAC_ERR_NCCL, "libnccl.so.2 could not be loaded");
  CU(cudaSetDevice(ctx->device));
  if (ctx->comm) gac_comm_destroy(ctx);
  NcclUid uid;
  memcpy(uid.internal, id128, 128);
  void* comm = nullptr;
  int r = a->CommInitRank(&comm, n_ranks, uid, rank);
  if (r != 0) return nccl_fail(a, r, "ncclCommInitRank");
  ctx->comm = comm;
  ctx->rank = rank;
  ctx->n_ranks = n_ranks;
  // NCCL connects its channels lazily on the first collective: pay that here, not inside the first render
  float* d_warm = nullptr;
  CU(cudaMallocAsync(&d_warm, 256 * sizeof(float), ctx->stream));
  CU(cudaMemsetAsync(d_warm, 0, 256 * sizeof(float), ctx->stream));
  r = a->Reduce(d_warm, d_warm, 256, /*ncclFloat32*/ 7, /*ncclSum*/ 0, 0, ctx->comm, ctx->stream);
  if (r != 0) return nccl_fail(a, r, "ncclReduce (warm-up)");
  CU(cudaFreeAsync(d_warm, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return GAC_OK;
}
extern "C" int gac_comm_destroy(gac_context* ctx) {
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or destroyed");
  if (ctx->comm) {
    NcclApi* a = nccl_api();
    cudaStreamSynchronize(ctx->stream);
    if (a) a->CommDestroy(ctx->comm);
    ctx->comm = nullptr;
  }
  ctx->rank = 0;
  ctx->n_ranks = 1;
  return GAC_OK;
}

// ------------------------------------------------------------------------------------------ render core
struct RenderArgs {
  const gac_graph* const* graphs = nullptr;
  int n_graphs = 0;
  int64_t first_frame = 0, n_frames = 0;
  float* const* h_out = nullptr;  // [n_graphs * n_out_channels] host rows, or null
  float* d_out = nullptr;         // [n_graphs][n_out_channels][n_frames] device, or null
  int n_out = 2;
  int64_t start_index = 0;
  bool sharded = false;
  int root = 0;
  bool sync = true;
  float* h_inter = nullptr;  // interleaved host output [start_index + n_frames][inter_channels], or null (one graph)
  int inter_channels = 0;
  bool overlap_d2h = false;  // batch renders: the results leave through ctx->d2h_stream behind a staging copy, so that the copy of
                             // one sub-batch runs while the next one computes (the caller synchronises ctx->d2h_stream at the end)
};

static int mix_into(RenderEnv& env, std::vector<MixJob>& jobs, std::vector<MixInput>& inputs) {
  if (jobs.empty()) return GAC_OK;
  auto& hj = env.keep->make<MixJob>();
  auto& hi = env.keep->make<MixInput>();
  hj = jobs;
  hi = inputs;
  MixJob* dj = nullptr;
  MixInput* di = nullptr;
  int rc;
  if ((rc = env.scratch->upload(&dj, hj))) return rc;
  if ((rc = env.scratch->upload(&di, hi))) return rc;
  int t = env.timer->begin(C_MIX);
  for (size_t j0 = 0; j0 < hj.size(); j0 += 65535) {
    size_t nj = std::min<size_t>(65535, hj.size() - j0);
    launch_mix(dj + j0, (int)nj, di, env.Npad, kstream(env.ctx));
    env.launches++;
  }
  env.timer->end(t);
  CU(cudaGetLastError());
  return GAC_OK;
}

static int render_core(gac_context* ctx, const RenderArgs& a) {
  HostTrace trace;
  struct TraceScope {
    explicit TraceScope(HostTrace* t) { g_trace = t->on ? t : nullptr; }
    ~TraceScope() { g_trace = nullptr; }
  } trace_scope(&trace);
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or destroyed");  // ObjectDisposedException (AudioContextBase.cs:54-55)
  if (!a.graphs || a.n_graphs <= 0) return fail(GAC_ERR_INVALID_ARGUMENT, "no graph");
  if (a.n_frames <= 0) return fail(GAC_ERR_OUT_OF_RANGE, "Frame count must be positive.");         // OfflineAudioContext.cs:35-36
  if (a.first_frame < 0 || a.start_index < 0) return fail(GAC_ERR_OUT_OF_RANGE, "Start index must be non-negative.");  // :38-39
  if (a.n_out < 1) return fail(GAC_ERR_INVALID_ARGUMENT, "Output buffer must have at least one channel.");             // :32-33
  if (a.n_out > 2) return fail(GAC_ERR_OUT_OF_RANGE, "the destination has 2 channels (AudioDestinationNode.cs:17)");
  for (int g = 0; g < a.n_graphs; g++) {
    if (!a.graphs[g]) return fail(GAC_ERR_INVALID_ARGUMENT, "graph %d is null", g);
    if (a.graphs[g]->ctx != ctx) return fail(GAC_ERR_INVALID_ARGUMENT, "graph %d belongs to another context", g);
  }
  if (a.h_out)
    for (int i = 0; i < a.n_graphs * a.n_out; i++)
      if (!a.h_out[i]) return fail(GAC_ERR_INVALID_ARGUMENT, "Channel %d buffer is null.", i % a.n_out);  // :47-48
  if (a.sharded && !ctx->comm && ctx->n_ranks > 1) return fail(GAC_ERR_INVALID_OPERATION, "gac_comm_init has not been called");
  CU(cudaSetDevice(ctx->device));

  const int B = ctx->B;
  const int64_t total = a.first_frame + a.n_frames;
  if (ctx->kick_pending) {  // impulse responses prepared ahead of this render: their staged job tables must have been read
    cudaEventSynchronize(ctx->kick_done);
    ctx->kick_pending = false;
  }
  ctx->stage_block = 0;
  ctx->stage_used = 0;  // the previous render has synchronised: its staged job tables are dead
  const int64_t copy_launches_before = ctx->copy_launches;
  ctx->pending.n = 0;
  struct DeferScope {  // job tables staged during the render travel with the next kernel launch, several per copy launch
    gac_context* c;
    explicit DeferScope(gac_context* x) : c(x) { c->defer_copies = true; }
    ~DeferScope() {
      flush_copies(c);
      c->defer_copies = false;
    }
  } defer_scope(ctx);
  RenderEnv env;
  Scratch scratch(ctx);
  HostKeep keep;
  Timer timer(ctx);
  env.ctx = ctx;
  env.scratch = &scratch;
  env.keep = &keep;
  env.timer = &timer;
  env.Npad = ((total + B - 1) / B) * B;
  env.NQ = env.Npad / 128;
  env.QB = env.Npad / B;
  int rc = ensure_block_times(ctx, env.NQ + 1);
  if (rc) return rc;
  cudaEvent_t trace_copy_done = nullptr;
  if (trace.on && ctx->copy_stream) {
    cudaEventCreate(&trace_copy_done);
    cudaEventRecord(trace_copy_done, ctx->copy_stream);
  }

  // ---- voices
  std::vector<const VoiceH*> voices;
  std::vector<int> voice_graph;
  for (int g = 0; g < a.n_graphs; g++)
    for (auto& v : a.graphs[g]->voices) {
      voices.push_back(&v);
      voice_graph.push_back(g);
    }
  const size_t S = voices.size();
  float* d_sig = nullptr;
  if ((rc = scratch.alloc(&d_sig, std::max<size_t>(1, S) * 2 * (size_t)env.Npad))) return rc;
  std::vector<Sig> sigs(S);
  for (size_t i = 0; i < S; i++) {
    sigs[i].p[0] = d_sig + (i * 2 + 0) * (size_t)env.Npad;
    sigs[i].p[1] = d_sig + (i * 2 + 1) * (size_t)env.Npad;
    sigs[i].ops = &voices[i]->ops;
  }
  // ---- fan-in fusion: a voice whose ONLY consumer is a plain two-channel fan-in (a bus head that takes its input as two channels, or
  // the destination) carries that fan-in's id; run_chains sums the ConvolverNodes that end such chains as spectra
  if (ctx->fuse_fanin) {
    std::vector<int> consumers(S, 0);
    size_t vb = 0, nb_total = 0;
    for (int g = 0; g < a.n_graphs; g++) {
      const gac_graph* gr = a.graphs[g];
      for (auto& b : gr->buses)
        for (int x : b.inputs)
          if (x < 0) consumers[vb + (size_t)(~x)]++;
      for (int x : gr->dest_inputs)
        if (x < 0) consumers[vb + (size_t)(~x)]++;
      vb += gr->voices.size();
      nb_total += gr->buses.size();
    }
    vb = 0;
    size_t bb = 0;
    for (int g = 0; g < a.n_graphs; g++) {
      const gac_graph* gr = a.graphs[g];
      for (size_t b = 0; b < gr->buses.size(); b++) {
        const BusH& bh = gr->buses[b];
        bool plain = !bh.mono && bh.slots.empty();
        if (plain && !bh.ops.empty()) {
          const OpH& h = bh.ops[0];
          plain = h.kind == GAC_OP_GAIN || h.kind == GAC_OP_BIQUAD || h.kind == GAC_OP_DELAY ||
                  (h.kind == GAC_OP_CONVOLVER && h.ftype == 0 && h.ir && h.ir->nch != 1);
        }
        if (!plain) continue;
        for (int x : bh.inputs)
          if (x < 0 && consumers[vb + (size_t)(~x)] == 1) sigs[vb + (size_t)(~x)].sum_key = (int64_t)(bb + b);
      }
      if (gr->dest_inputs.size() > 1)
        for (int x : gr->dest_inputs)
          if (x < 0 && consumers[vb + (size_t)(~x)] == 1) sigs[vb + (size_t)(~x)].sum_key = (int64_t)(nb_total + (size_t)g);
      vb += gr->voices.size();
      bb += gr->buses.size();
    }
  }
  // ---- stage 1: the chains fed by source buffers.
  // With asynchronous uploads some buffers may still be in flight (and their impulse responses not yet prepared).  Uploads
  // are queued in creation order, so the voices whose data has landed form a prefix: that prefix runs at once as one batch,
  // the rest follows in small batches so that little work is left when the copy engine delivers the last buffer.
  auto landed = [&](const VoiceH* v) {
    if (v->src && v->src->ready && cudaEventQuery(v->src->ready) == cudaErrorNotReady) return false;
    for (const OpH& op : v->ops)
      if (op.kind == GAC_OP_CONVOLVER && op.ir && !op.ir->prepared && op.ir->src && op.ir->src->ready &&
          cudaEventQuery(op.ir->src->ready) == cudaErrorNotReady)
        return false;
    return true;
  };
  // parameters with a modulation input make their owner wait for the bus that carries the modulator (graph-local bus indices)
  auto ops_deps = [](const std::vector<OpH>& ops, std::vector<int>& out) {
    for (const OpH& o : ops)
      for (const ParamH* p : {&o.p0, &o.p1, &o.p2})
        if (p->mod_bus >= 0) out.push_back(p->mod_bus);
  };
  std::vector<std::vector<int>> voice_deps(S);
  for (size_t i = 0; i < S; i++) {
    ops_deps(voices[i]->ops, voice_deps[i]);
    if (voices[i]->src_param.mod_bus >= 0) voice_deps[i].push_back(voices[i]->src_param.mod_bus);
  }
  std::vector<size_t> src_voices;  // the chains fed by buffer sources that wait for nothing: stage 1
  size_t n_bus_fed = 0;
  for (size_t i = 0; i < S; i++) {
    if (voices[i]->input_bus >= 0) n_bus_fed++;
    if (voices[i]->input_bus < 0 && voices[i]->kind == GAC_SOURCE_BUFFER && voice_deps[i].empty()) src_voices.push_back(i);
  }
  const size_t S0 = src_voices.size();
  std::vector<char> voice_done(S, 0);
  {
    size_t v_ready = 0;
    while (v_ready < S0 && landed(voices[src_voices[v_ready]])) v_ready++;
    cudaGetLastError();
    std::vector<size_t> cuts;
    cuts.push_back(0);
    if (v_ready < S0 && S0 >= 16) {
      if (v_ready >= 4) cuts.push_back(v_ready);
      // the rest in up to three batches: each one is queued while its uploads are still in flight.  (Halving batches down to two or
      // three voices were measured too: the fixed cost of six more launches outweighs the shorter last batch, 0.74 vs 0.62 ms behind
      // the last upload on the 64-voice bench workload.)  A small rest is ONE batch: the landed prefix keeps the device busy longer
      // than the rest needs to arrive, and every further batch adds its fixed cost behind the last upload (C3 shard, 21 of 128 voices
      // in flight at the call: 3.65 ms of device time in four batches).
      const size_t rest = S0 - cuts.back();
      size_t n_rest = rest * 3 <= S0 ? 1 : (rest * 2 <= S0 ? 2 : 3);
      if (const char* e = getenv("GAC_REST_BATCHES")) {  // (measurements)
        const int x = atoi(e);
        if (x >= 1 && x <= 8) n_rest = (size_t)x;
      }
      const size_t chunk = std::max<size_t>(4, (rest + n_rest - 1) / n_rest);
      for (size_t v = cuts.back() + chunk; v < S0; v += chunk) cuts.push_back(v);
    }
    cuts.push_back(S0);
    if (trace.on) fprintf(stderr, "[gac_trace] voices %zu landed %zu batches %zu\n", S0, v_ready, cuts.size() - 1);
    for (size_t bi = 0; bi + 1 < cuts.size(); bi++) {
      const size_t v0 = cuts[bi], v1 = cuts[bi + 1];
      if (v1 <= v0) continue;
      std::vector<const VoiceH*> vsub;
      std::vector<Sig> ssub;
      for (size_t k = v0; k < v1; k++) {
        vsub.push_back(voices[src_voices[k]]);
        ssub.push_back(sigs[src_voices[k]]);
      }
      if ((rc = plan_sources(env, vsub, ssub))) return rc;
      trace.mark("sources planned");
      if ((rc = run_chains(env, ssub))) return rc;
      for (size_t k = v0; k < v1; k++) {
        sigs[src_voices[k]] = ssub[k - v0];
        voice_done[src_voices[k]] = 1;
      }
      trace.mark("voice batch queued");
    }
  }

  // ---- stage 2: buses (fan-in in connection order, AudioNodeInput.cs:118-137, then the bus ops) and the chains fed by bus
  // outputs, in dependency order: every round mixes the buses whose inputs are complete, then runs the chains they feed
  size_t NB = 0;
  std::vector<size_t> bus_base(a.n_graphs), voice_base(a.n_graphs);
  {
    size_t vb = 0;
    for (int g = 0; g < a.n_graphs; g++) {
      bus_base[g] = NB;
      voice_base[g] = vb;
      NB += a.graphs[g]->buses.size();
      vb += a.graphs[g]->voices.size();
    }
  }
  bool hierarchy = n_bus_fed != 0;
  for (size_t i = 0; i < S; i++) hierarchy = hierarchy || !voice_deps[i].empty();
  for (int g = 0; g < a.n_graphs; g++)
    for (auto& b : a.graphs[g]->buses) hierarchy = hierarchy || b.target != -1;
  if (a.sharded && ctx->n_ranks > 1 && hierarchy)
    return fail(GAC_ERR_UNSUPPORTED, "sharded renders need a flat graph: voices -> buses -> destination (the buses are what is reduced)");
  const bool is_root = !a.sharded || ctx->n_ranks <= 1 || ctx->rank == a.root;
  float* d_bus = nullptr;
  std::vector<Sig> buses(NB);
  std::vector<char> bus_done(NB, 0);
  if (NB) {
    if ((rc = scratch.alloc(&d_bus, NB * 2 * (size_t)env.Npad))) return rc;
    for (int g = 0; g < a.n_graphs; g++)
      for (size_t b = 0; b < a.graphs[g]->buses.size(); b++) {
        Sig& bs = buses[bus_base[g] + b];
        bs.p[0] = d_bus + ((bus_base[g] + b) * 2 + 0) * (size_t)env.Npad;
        bs.p[1] = d_bus + ((bus_base[g] + b) * 2 + 1) * (size_t)env.Npad;
        bs.ops = &a.graphs[g]->buses[b].ops;
        bs.bus_base = bus_base[g];
      }
  }
  env.buses = &buses;
  for (size_t i = 0; i < S; i++) sigs[i].bus_base = bus_base[voice_graph[i]];
  std::vector<std::vector<int>> bus_deps(NB);
  for (int g = 0; g < a.n_graphs; g++)
    for (size_t b = 0; b < a.graphs[g]->buses.size(); b++) ops_deps(a.graphs[g]->buses[b].ops, bus_deps[bus_base[g] + b]);
  float* d_zero_row = nullptr;  // what a ChannelMergerNode connection adds to the channel it does not feed
  auto zero_row = [&]() -> int {
    if (d_zero_row) return GAC_OK;
    int rc0 = scratch.alloc(&d_zero_row, (size_t)env.Npad);
    if (rc0) return rc0;
    CU(cudaMemsetAsync(d_zero_row, 0, sizeof(float) * (size_t)env.Npad, ctx->stream));
    return GAC_OK;
  };
  for (size_t round = 0;; round++) {
    // buses whose inputs are all available
    std::vector<size_t> ready;
    for (int g = 0; g < a.n_graphs; g++) {
      const gac_graph* gr = a.graphs[g];
      for (size_t b = 0; b < gr->buses.size(); b++) {
        if (bus_done[bus_base[g] + b]) continue;
        bool ok = true;
        for (int x : gr->buses[b].inputs) ok = ok && (x >= 0 ? bus_done[bus_base[g] + (size_t)x] : voice_done[voice_base[g] + (size_t)(~x)]);
        for (int d : bus_deps[bus_base[g] + b]) ok = ok && bus_done[bus_base[g] + (size_t)d];
        if (ok) ready.push_back(bus_base[g] + b);
      }
    }
    std::vector<size_t> fed;  // chains fed by a finished bus
    if (ready.empty()) {
      for (size_t i = 0; i < S; i++) {
        if (voice_done[i]) continue;
        bool ok = voices[i]->input_bus < 0 || bus_done[bus_base[voice_graph[i]] + (size_t)voices[i]->input_bus];
        for (int d : voice_deps[i]) ok = ok && bus_done[bus_base[voice_graph[i]] + (size_t)d];
        if (ok) fed.push_back(i);
      }
      if (fed.empty()) break;
    }
    std::vector<MixJob> mjobs;
    std::vector<MixInput> minputs;
    if (!ready.empty()) {
      for (size_t gb : ready) {
        int g = 0;
        while (g + 1 < a.n_graphs && bus_base[g + 1] <= gb) g++;
        const gac_graph* gr = a.graphs[g];
        const BusH& bh = gr->buses[gb - bus_base[g]];
        Sig& bs = buses[gb];
        MixJob mj;
        mj.dst[0] = bs.p[0];
        mj.dst[1] = bs.p[1];
        mj.first_input = (int)minputs.size();
        int64_t lo = std::numeric_limits<int64_t>::max(), hi = 0;
        // Channel count of the fan-in block: the head node's input mode decides (AudioNodeInput.ComputeOutputChannelCount :139-168).
        //   Max with channelCount 2 (GainNode, BiQuadFilterNode, DelayNode) and Explicit 2 (stereo / true-stereo ConvolverNode): 2,
        //     a mono input is copied into both channels (:201-213);
        //   Explicit 1 (ConvolverNode with a mono impulse response, ConvolverNode.cs:72-76): 1 — a mono input is added as it is, a
        //     stereo input as (L + R) * 1/sqrt(2), each input on its own (:214-228);
        //   ClampedMax 2 (StereoPannerNode.cs:24-26): min(max over the inputs' blocks, 2), i.e. 1 when every input is mono.
        auto sig_in = [&](int x) -> const Sig& { return x >= 0 ? buses[bus_base[g] + (size_t)x] : sigs[voice_base[g] + (size_t)(~x)]; };
        int head_ch = 2;
        if (bh.mono) {
          head_ch = 1;  // the input of an AudioParam: Explicit, one channel (AudioParam.cs:60-62)
        } else if (!bh.slots.empty()) {
          head_ch = 2;  // ChannelMergerNode with two inputs: its output has two channels
        } else if (bh.ops.empty()) {
          for (int x : bh.inputs) head_ch = sig_in(x).ch;  // a materialised fan-out point: no node, the signal passes through
        } else if (bh.ops[0].kind == GAC_OP_CONVOLVER && bh.ops[0].ir && bh.ops[0].ir->nch == 1) {
          head_ch = 1;
        } else if (bh.ops[0].kind == GAC_OP_PANNER) {
          bool all_mono = true, steady2 = false;
          for (int x : bh.inputs) {
            const Sig& vs = sig_in(x);
            if (vs.hi <= vs.lo || vs.ch == 1) continue;
            all_mono = false;
            // a source connected directly carries two channels only while it plays (its idle block has one, AudioBufferSourceNode.cs:391-402)
            const bool direct = x < 0 && vs.from_source && (!vs.ops || vs.ops->empty());
            if (!direct || (vs.lo == 0 && vs.hi == env.Npad)) steady2 = true;
          }
          if (all_mono) head_ch = 1;
          else if (!steady2)
            return fail(GAC_ERR_UNSUPPORTED, "a StereoPannerNode fed by several inputs whose channel count changes during the render (stereo sources "
                                             "connected directly that start late or end early) is outside the accelerated path");
        }
        for (size_t xi = 0; xi < bh.inputs.size(); xi++) {
          const int x = bh.inputs[xi];
          const Sig& vs = sig_in(x);
          if (vs.hi <= vs.lo) continue;  // silent throughout: never mixed (:127)
          MixInput in;
          in.src[0] = vs.p[0];
          in.src[1] = vs.p[1];
          in.lo = vs.lo;
          in.hi = vs.hi;
          if (head_ch == 1 && vs.ch == 2) in.downmix = 1.0f / sqrtf(2.0f);  // 1.0f / MathF.Sqrt(srcChannels)  (:217)
          const int slot = bh.slots.empty() ? 0 : bh.slots[xi];
          if (slot) {
            // ChannelMergerNode (Nodes/ChannelMergerNode.cs:40-62): channel 0 of what merger input slot-1 mixes becomes output
            // channel slot-1; the other channel receives + 0 from this connection
            if ((rc = zero_row())) return rc;
            in.src[slot == 1 ? 0 : 1] = vs.p[0];
            in.src[slot == 1 ? 1 : 0] = d_zero_row;
          }
          minputs.push_back(in);
          lo = std::min(lo, vs.lo);
          hi = std::max(hi, vs.hi);
        }
        bs.ch = head_ch;
        mj.n_inputs = (int)minputs.size() - mj.first_input;
        bs.lo = mj.n_inputs ? lo : 0;  // non-silent where any input was mixed (convex hull)
        bs.hi = mj.n_inputs ? hi : 0;
        mjobs.push_back(mj);
      }
      if ((rc = mix_into(env, mjobs, minputs))) return rc;
      if (a.sharded && ctx->n_ranks > 1) {
        // the one exchange step: per-rank partial bus sums -> root, float32 sum over NVLink/NVSwitch (flat graphs: every bus is
        // ready in the first round, so the whole bus block is reduced at once)
        NcclApi* api = nccl_api();
        if (!api) return fail(GAC_ERR_NCCL, "libnccl.so.2 could not be loaded");
        int t = timer.begin(C_MIX);
        int r = api->Reduce(d_bus, d_bus, NB * 2 * (size_t)env.Npad, /*ncclFloat32*/ 7, /*ncclSum*/ 0, a.root, ctx->comm, ctx->stream);
        timer.end(t);
        if (r != 0) return nccl_fail(api, r, "ncclReduce");
        for (auto& bs : buses) {  // other ranks' voices may be audible anywhere
          bs.lo = 0;
          bs.hi = env.Npad;
        }
      }
      if (is_root) {
        std::vector<Sig> sub;
        for (size_t gb : ready) sub.push_back(buses[gb]);
        if ((rc = run_chains(env, sub))) return rc;
        for (size_t k = 0; k < ready.size(); k++) buses[ready[k]] = sub[k];
      }
      for (size_t gb : ready) bus_done[gb] = 1;
      continue;
    }
    // chains that became runnable: (a) fed by a bus output — the chain works on its own copy of the bus signal ((0 + x) == x, silent
    // quanta stay silent), or on ONE channel of it behind a ChannelSplitterNode; (b) buffer sources whose chain waited for a
    // modulator; (c) scheduled sources (ConstantSourceNode / OscillatorNode)
    std::vector<Sig> sub;
    std::vector<const VoiceH*> late_src;
    std::vector<size_t> late_src_pos, sched_pos;
    for (size_t i : fed) {
      Sig& s = sigs[i];
      if (voices[i]->input_bus >= 0) {
        const Sig& src = buses[bus_base[voice_graph[i]] + (size_t)voices[i]->input_bus];
        MixJob mj;
        mj.dst[0] = s.p[0];
        mj.dst[1] = s.p[1];
        mj.first_input = (int)minputs.size();
        s.lo = src.lo;
        s.hi = src.hi;
        s.ch = src.ch;
        int pick = -1;  // ChannelSplitterNode output (Nodes/ChannelSplitterNode.cs:29-62): channel `pick` of the input as one channel
        if (!voices[i]->ops.empty() && voices[i]->ops[0].kind == GAC_OP_CHANNEL) pick = (int)voices[i]->ops[0].aux;
        if (pick >= 2) s.lo = s.hi = 0;  // the splitter's input has two channels (Max mode, channelCount 2): outputs beyond are cleared
        if (s.hi > s.lo) {
          MixInput in;
          in.src[0] = pick == 1 ? src.p[1] : src.p[0];  // (a mono signal keeps its channel in both rows: the 1 -> 2 up-mix of the input)
          in.src[1] = pick == 0 ? src.p[0] : src.p[1];
          in.lo = src.lo;
          in.hi = src.hi;
          minputs.push_back(in);
        }
        if (pick >= 0) s.ch = 1;
        mj.n_inputs = (int)minputs.size() - mj.first_input;
        mjobs.push_back(mj);
      } else if (voices[i]->kind == GAC_SOURCE_BUFFER) {
        late_src.push_back(voices[i]);
        late_src_pos.push_back(sub.size());
      } else {
        sched_pos.push_back(sub.size());
      }
      sub.push_back(s);
    }
    if (is_root || !a.sharded) {
      if ((rc = mix_into(env, mjobs, minputs))) return rc;
      if (!late_src.empty()) {
        std::vector<Sig> ls;
        for (size_t k : late_src_pos) ls.push_back(sub[k]);
        if ((rc = plan_sources(env, late_src, ls))) return rc;
        for (size_t k = 0; k < late_src_pos.size(); k++) sub[late_src_pos[k]] = ls[k];
      }
      if (!sched_pos.empty()) {
        std::vector<const VoiceH*> sv;
        std::vector<Sig> ss;
        for (size_t k : sched_pos) {
          sv.push_back(voices[fed[k]]);
          ss.push_back(sub[k]);
        }
        if ((rc = plan_scheduled(env, sv, ss))) return rc;
        for (size_t k = 0; k < sched_pos.size(); k++) sub[sched_pos[k]] = ss[k];
      }
      if ((rc = run_chains(env, sub))) return rc;
    }
    for (size_t k = 0; k < fed.size(); k++) {
      sigs[fed[k]] = sub[k];
      voice_done[fed[k]] = 1;
    }
  }
  for (size_t i = 0; i < S; i++)
    if (!voice_done[i]) return fail(GAC_ERR_INVALID_ARGUMENT, "the graph has a cycle or a chain fed by a bus that never completes");
  for (size_t b = 0; b < NB; b++)
    if (!bus_done[b]) return fail(GAC_ERR_INVALID_ARGUMENT, "the graph has a cycle through bus %zu", b);
  if (is_root) {
    // ---- destination: fan-in of buses and direct voices in connection order; alias when there is exactly one input
    std::vector<const float*> dest0(a.n_graphs), dest1(a.n_graphs);
    float* d_dest = nullptr;
    std::vector<MixJob> mjobs;
    std::vector<MixInput> minputs;
    size_t vbase = 0;
    size_t need = 0;
    for (int g = 0; g < a.n_graphs; g++)
      if (a.graphs[g]->dest_inputs.size() != 1) need++;
    if (need && (rc = scratch.alloc(&d_dest, need * 2 * (size_t)env.Npad))) return rc;
    size_t used = 0;
    for (int g = 0; g < a.n_graphs; g++) {
      const gac_graph* gr = a.graphs[g];
      auto sig_of = [&](int x) -> const Sig& { return x >= 0 ? buses[bus_base[g] + x] : sigs[vbase + (size_t)(~x)]; };
      if (gr->dest_inputs.size() == 1 && sig_of(gr->dest_inputs[0]).lo == 0 && sig_of(gr->dest_inputs[0]).hi == env.Npad) {
        const Sig& s = sig_of(gr->dest_inputs[0]);  // (0 + x) == x exactly
        dest0[g] = s.p[0];
        dest1[g] = s.p[1];
      } else {
        if (gr->dest_inputs.size() == 1) {  // rare: single input with silent-flagged quanta; needs its own buffer
          float* extra = nullptr;
          if ((rc = scratch.alloc(&extra, 2 * (size_t)env.Npad))) return rc;
          dest0[g] = extra;
          dest1[g] = extra + env.Npad;
        } else {
          dest0[g] = d_dest + (used * 2 + 0) * (size_t)env.Npad;
          dest1[g] = d_dest + (used * 2 + 1) * (size_t)env.Npad;
          used++;
        }
        MixJob mj;
        mj.dst[0] = const_cast<float*>(dest0[g]);
        mj.dst[1] = const_cast<float*>(dest1[g]);
        mj.first_input = (int)minputs.size();
        for (int x : gr->dest_inputs) {
          const Sig& s = sig_of(x);
          if (s.hi <= s.lo) continue;
          MixInput in;
          in.src[0] = s.p[0];
          in.src[1] = s.p[1];
          in.lo = s.lo;
          in.hi = s.hi;
          minputs.push_back(in);
        }
        mj.n_inputs = (int)minputs.size() - mj.first_input;
        mjobs.push_back(mj);
      }
      vbase += gr->voices.size();
    }
    if ((rc = mix_into(env, mjobs, minputs))) return rc;

    // ---- output: frames [first_frame, first_frame + n_frames) (OfflineAudioContext.cs:77-101)
    int t = timer.begin(C_D2H);
    for (int g = 0; g < a.n_graphs; g++) {
      for (int c = 0; c < a.n_out; c++) {
        const float* src = (c == 0 ? dest0[g] : dest1[g]) + a.first_frame;
        if (a.h_inter) {
          // (handled below: one interleaving pass over both rows)
        } else if (a.h_out && a.overlap_d2h) {
          // (handled below: staged, then copied out on the second stream)
        } else if (a.h_out) {
          CU(cudaMemcpyAsync(a.h_out[(size_t)g * a.n_out + c] + a.start_index, src, sizeof(float) * (size_t)a.n_frames, cudaMemcpyDeviceToHost, ctx->stream));
        } else if (a.d_out) {
          CU(cudaMemcpyAsync(a.d_out + ((size_t)g * a.n_out + c) * (size_t)a.n_frames, src, sizeof(float) * (size_t)a.n_frames, cudaMemcpyDeviceToDevice,
                             ctx->stream));
        }
      }
    }
    if (a.h_out && a.overlap_d2h && !a.h_inter) {
      // the destination rows live in the render's scratch arena, which the next sub-batch reuses: they are copied (device to device,
      // ~0.3 ms per GB) into a block of their own that the copy stream frees when it is done with it
      float* stage = nullptr;
      const size_t rows = (size_t)a.n_graphs * a.n_out;
      CU(cudaMallocAsync(&stage, sizeof(float) * rows * (size_t)a.n_frames, ctx->stream));
      for (int g = 0; g < a.n_graphs; g++)
        for (int c = 0; c < a.n_out; c++)
          CU(cudaMemcpyAsync(stage + ((size_t)g * a.n_out + c) * (size_t)a.n_frames, (c == 0 ? dest0[g] : dest1[g]) + a.first_frame,
                             sizeof(float) * (size_t)a.n_frames, cudaMemcpyDeviceToDevice, ctx->stream));
      cudaEvent_t staged = take_event(ctx);
      CU(cudaEventRecord(staged, ctx->stream));
      CU(cudaStreamWaitEvent(ctx->d2h_stream, staged, 0));
      ctx->event_pool.push_back(staged);
      for (int g = 0; g < a.n_graphs; g++)
        for (int c = 0; c < a.n_out; c++)
          CU(cudaMemcpyAsync(a.h_out[(size_t)g * a.n_out + c] + a.start_index, stage + ((size_t)g * a.n_out + c) * (size_t)a.n_frames,
                             sizeof(float) * (size_t)a.n_frames, cudaMemcpyDeviceToHost, ctx->d2h_stream));
      CU(cudaFreeAsync(stage, ctx->d2h_stream));
    }
    if (a.h_inter) {  // ≙ the interleaving loop of ProcessBlockInterleaved (AudioContextBase.cs:127-160), for the whole render
      float* d_inter = nullptr;
      if ((rc = scratch.alloc(&d_inter, (size_t)a.n_frames * a.inter_channels))) return rc;
      launch_interleave(dest0[0] + a.first_frame, a.n_out > 1 ? dest1[0] + a.first_frame : nullptr, d_inter, a.n_frames, a.inter_channels, kstream(ctx));
      env.launches++;
      CU(cudaMemcpyAsync(a.h_inter + (size_t)a.start_index * a.inter_channels, d_inter, sizeof(float) * (size_t)a.n_frames * a.inter_channels,
                         cudaMemcpyDeviceToHost, ctx->stream));
    }
    timer.end(t);
  }
  // ---- finish: the host job arrays in `keep` must outlive the stream work, so every render synchronises here
  flush_copies(ctx);
  trace.mark("everything queued");
  gac_stats st{};
  timer.finish(&st);
  trace.mark("device finished");
  if (trace.on && trace_copy_done) {  // diagnostics: when did the last queued upload land, relative to the render's first / last event?
    float a = 0.f, b = 0.f;
    cudaEventSynchronize(trace_copy_done);
    cudaEventElapsedTime(&a, timer.t0, trace_copy_done);
    cudaEventElapsedTime(&b, trace_copy_done, timer.t1);
    fprintf(stderr, "[gac_trace] uploads landed %.3f ms after the render's first event, %.3f ms before its last\n", a, b);
    cudaEventDestroy(trace_copy_done);
  }
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return fail(GAC_ERR_CUDA, "render failed: %s", cudaGetErrorString(e));
  // the async-upload contract releases the caller's arrays when a render returns: buffers this render did not use too
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(GAC_ERR_CUDA, "render failed: %s", cudaGetErrorString(e));
  st.conv_units = env.conv_units;
  st.algorithmic_bytes = env.alg_bytes;
  st.mac_complex_macs = env.macs;
  st.mac_flops = env.mac_flops;
  st.mac_bytes_moved = env.mac_bytes;
  st.mac_h2_bytes_single = env.mac_h2_single;
  st.mac_variant_used = env.mac_used;
  st.mac_big_segments = env.mac_big;
  st.fanin_groups = env.sum_groups;
  st.fanin_members = env.sum_members;
  flush_copies(ctx);
  st.kernel_launches = env.launches + (ctx->copy_launches - copy_launches_before);  // (the table-copy kernels included)
  st.voices = (int64_t)S;
  st.frames = a.n_frames;
  ctx->stats = st;
  trace.mark("streams synchronised");
  if (trace.on) {  // (what the destructors do on return, made visible)
    scratch.release();
    trace.mark("scratch released");
  }
  return GAC_OK;
}

extern "C" int gac_render(gac_context* ctx, const gac_graph* graph, int64_t first_frame, int64_t n_frames, float* const* out_channels,
                          int n_out_channels, int64_t start_index) {
  if (!out_channels) return fail(GAC_ERR_INVALID_ARGUMENT, "Output buffer must have at least one channel.");
  RenderArgs a;
  a.graphs = &graph;
  a.n_graphs = 1;
  a.first_frame = first_frame;
  a.n_frames = n_frames;
  a.h_out = out_channels;
  a.n_out = n_out_channels;
  a.start_index = start_index;
  return render_core(ctx, a);
}
// ≙ rendering with ProcessBlockInterleaved (AudioContextBase.cs:88-161) block after block into one interleaved array
extern "C" int gac_render_interleaved(gac_context* ctx, const gac_graph* graph, int64_t first_frame, int64_t n_frames, float* interleaved,
                                      int channels, int64_t start_index) {
  if (!interleaved) return fail(GAC_ERR_INVALID_ARGUMENT, "interleaved buffer is null");
  if (channels < 1 || channels > 32) return fail(GAC_ERR_OUT_OF_RANGE, "channels must be between 1 and 32");  // AudioContextBase.cs:93
  RenderArgs a;
  a.graphs = &graph;
  a.n_graphs = 1;
  a.first_frame = first_frame;
  a.n_frames = n_frames;
  a.h_inter = interleaved;
  a.inter_channels = channels;
  a.n_out = channels < 2 ? 1 : 2;  // usedChannels = min(channels, destination channels) (:125)
  a.start_index = start_index;
  return render_core(ctx, a);
}
extern "C" int gac_render_device(gac_context* ctx, const gac_graph* graph, int64_t first_frame, int64_t n_frames, float* d_out, int n_out_channels,
                                 int sync) {
  if (!d_out) return fail(GAC_ERR_INVALID_ARGUMENT, "d_out is null");
  RenderArgs a;
  a.graphs = &graph;
  a.n_graphs = 1;
  a.first_frame = first_frame;
  a.n_frames = n_frames;
  a.d_out = d_out;
  a.n_out = n_out_channels;
  a.sync = sync != 0;
  return render_core(ctx, a);
}
extern "C" int gac_render_batch(gac_context* ctx, const gac_graph* const* graphs, int n_graphs, int64_t n_frames, float* const* out_channels,
                                int n_out_channels) {
  if (!out_channels) return fail(GAC_ERR_INVALID_ARGUMENT, "Output buffer must have at least one channel.");
  RenderArgs a;
  a.graphs = graphs;
  a.n_graphs = n_graphs;
  a.n_frames = n_frames;
  a.h_out = out_channels;
  a.n_out = n_out_channels;
  // Opt-in (GAC_BATCH_SPLIT=k): a large batch is cut into k sub-batches whose results leave on a second stream while the next
  // sub-batch computes.  Measured on a 512-render shard of BASELINE config 4 (983 MB of results, 21 ms over PCIe, 9.1 ms of
  // compute): 1 batch 31.0 ms, k = 2: 30.9 ms, 3: 31.7 ms, 4: 34.9 ms, 8: 53.8 ms — a render has latency-bound stages (the
  // sequential biquad pass of a slow filter costs ~5 ms whatever the batch size), so every cut adds about as much compute as it
  // hides copy time.  Hence off by default; the copy is what bounds this configuration end to end.
  static const int forced = getenv("GAC_BATCH_SPLIT") ? atoi(getenv("GAC_BATCH_SPLIT")) : 0;
  const double out_bytes = (double)n_graphs * n_out_channels * (double)n_frames * 4.0;
  if (forced < 2 || !ctx_ok(ctx) || n_graphs < 32 || out_bytes < 64e6 || n_frames <= 0 || n_out_channels < 1 || n_out_channels > 2) return render_core(ctx, a);
  CU(cudaSetDevice(ctx->device));
  if (!ctx->d2h_stream) CU(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
  const int n_sub = std::max(1, std::min(forced, n_graphs / 16));
  gac_stats total{};
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  CU(cudaEventRecord(e0, ctx->stream));
  int rc = GAC_OK;
  for (int k = 0; k < n_sub && rc == GAC_OK; k++) {
    const int g0 = (int)((int64_t)n_graphs * k / n_sub), g1 = (int)((int64_t)n_graphs * (k + 1) / n_sub);
    RenderArgs sub = a;
    sub.graphs = graphs + g0;
    sub.n_graphs = g1 - g0;
    sub.h_out = out_channels + (size_t)g0 * n_out_channels;
    sub.overlap_d2h = true;
    rc = render_core(ctx, sub);
    if (rc) break;
    const gac_stats& st = ctx->stats;
    total.ms_source += st.ms_source; total.ms_automation += st.ms_automation; total.ms_biquad += st.ms_biquad; total.ms_gain += st.ms_gain;
    total.ms_fft_fwd += st.ms_fft_fwd; total.ms_mac += st.ms_mac; total.ms_fft_inv += st.ms_fft_inv; total.ms_mix += st.ms_mix;
    total.ms_delay += st.ms_delay; total.ms_panner += st.ms_panner;
    total.conv_units += st.conv_units; total.algorithmic_bytes += st.algorithmic_bytes; total.mac_complex_macs += st.mac_complex_macs;
    total.kernel_launches += st.kernel_launches; total.voices += st.voices; total.mac_flops += st.mac_flops;
    total.mac_bytes_moved += st.mac_bytes_moved; total.mac_h2_bytes_single += st.mac_h2_bytes_single;
    total.mac_variant_used = st.mac_variant_used; total.mac_big_segments = st.mac_big_segments; total.frames = st.frames;
    total.fanin_groups += st.fanin_groups; total.fanin_members += st.fanin_members;
  }
  cudaError_t e = cudaStreamSynchronize(ctx->d2h_stream);
  cudaEventRecord(e1, ctx->stream);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(GAC_ERR_CUDA, "batch render failed: %s", cudaGetErrorString(e));
  total.ms_total = ms;  // first kernel to last byte on the host
  const double compute = total.ms_source + total.ms_automation + total.ms_biquad + total.ms_gain + total.ms_fft_fwd + total.ms_mac + total.ms_fft_inv +
                         total.ms_mix + total.ms_delay + total.ms_panner;
  total.ms_d2h = std::max(0.0, (double)ms - compute);  // what the copies add behind the kernels they could not hide under
  ctx->stats = total;
  return GAC_OK;
}
extern "C" int gac_render_sharded(gac_context* ctx, const gac_graph* shard, int64_t n_frames, int root, float* const* out_channels,
                                  int n_out_channels) {
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or destroyed");
  if (!shard) return fail(GAC_ERR_INVALID_ARGUMENT, "graph is null");
  for (auto& v : shard->voices)
    if (v.bus < 0) return fail(GAC_ERR_UNSUPPORTED, "sharded renders need every voice routed through a bus (the bus is what is reduced)");
  if (shard->buses.empty()) return fail(GAC_ERR_UNSUPPORTED, "sharded renders need at least one bus");
  const bool is_root = ctx->n_ranks <= 1 || ctx->rank == root;
  if (is_root && !out_channels) return fail(GAC_ERR_INVALID_ARGUMENT, "Output buffer must have at least one channel.");
  RenderArgs a;
  a.graphs = &shard;
  a.n_graphs = 1;
  a.n_frames = n_frames;
  a.h_out = is_root ? out_channels : nullptr;
  a.n_out = n_out_channels;
  a.sharded = true;
  a.root = root;
  return render_core(ctx, a);
}

// ------------------------------------------------------------------------------------------ one process, several GPUs
// The reference's caller is ONE process holding ONE OfflineAudioContext (OfflineAudioContext.cs:18).  A gac_group is that context
// spread over several GPUs of the box: one member gac_context (own stream, own scratch arena) per device, one NCCL communicator
// over the device set (ncclCommInitAll: SURVEY.md §8e), voices sharded over the members by the host mirror.  gac_group_render
// runs the members' shards concurrently — one host thread per device, because a render is synchronous and the single
// ncclReduce of the bus needs every rank inside it at the same time — and the root member (index 0) delivers the result.
struct gac_group {
  uint32_t magic = 0x47414347;  // "GACG"
  std::vector<gac_context*> members;
};
static bool group_ok(gac_group* g) { return g && g->magic == 0x47414347; }

extern "C" int gac_group_destroy(gac_group* grp);
extern "C" int gac_group_create(const gac_context_desc* desc, const int* device_ids, int n_devices, gac_group** out) {
  if (!desc || !out || !device_ids) return fail(GAC_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  if (n_devices < 1 || n_devices > 64) return fail(GAC_ERR_OUT_OF_RANGE, "n_devices must be in 1 .. 64");
  for (int i = 0; i < n_devices; i++)
    for (int k = 0; k < i; k++)
      if (device_ids[i] == device_ids[k]) return fail(GAC_ERR_INVALID_ARGUMENT, "device %d listed twice", device_ids[i]);
  auto grp = std::make_unique<gac_group>();
  for (int i = 0; i < n_devices; i++) {
    gac_context_desc d = *desc;
    d.device_id = device_ids[i];
    gac_context* c = nullptr;
    int rc = gac_context_create(&d, &c);
    if (rc) {
      gac_group* partial = grp.release();
      gac_group_destroy(partial);
      return rc;
    }
    grp->members.push_back(c);
  }
  if (n_devices > 1) {
    NcclApi* a = nccl_api();
    if (!a || !a->CommInitAll || !a->GroupStart || !a->GroupEnd) {
      gac_group_destroy(grp.release());
      return fail(GAC_ERR_NCCL, "libnccl.so.2 could not be loaded (ncclCommInitAll)");
    }
    std::vector<void*> comms((size_t)n_devices, nullptr);
    int r = a->CommInitAll(comms.data(), n_devices, device_ids);
    if (r != 0) {
      gac_group_destroy(grp.release());
      return nccl_fail(a, r, "ncclCommInitAll");
    }
    for (int i = 0; i < n_devices; i++) {
      grp->members[i]->comm = comms[i];
      grp->members[i]->rank = i;
      grp->members[i]->n_ranks = n_devices;
    }
    // NCCL connects its channels lazily on the first collective: pay that here, not inside the first render
    std::vector<float*> warm((size_t)n_devices, nullptr);
    for (int i = 0; i < n_devices; i++) {
      gac_context* c = grp->members[i];
      cudaSetDevice(c->device);
      cudaMallocAsync(&warm[i], 256 * sizeof(float), c->stream);
      cudaMemsetAsync(warm[i], 0, 256 * sizeof(float), c->stream);
    }
    a->GroupStart();
    for (int i = 0; i < n_devices; i++) {
      gac_context* c = grp->members[i];
      cudaSetDevice(c->device);
      r = a->Reduce(warm[i], warm[i], 256, /*ncclFloat32*/ 7, /*ncclSum*/ 0, 0, c->comm, c->stream);
      if (r != 0) break;
    }
    const int r2 = a->GroupEnd();
    for (int i = 0; i < n_devices; i++) {
      gac_context* c = grp->members[i];
      cudaSetDevice(c->device);
      cudaFreeAsync(warm[i], c->stream);
      cudaStreamSynchronize(c->stream);
    }
    if (r != 0 || r2 != 0) {
      gac_group_destroy(grp.release());
      return nccl_fail(a, r != 0 ? r : r2, "ncclReduce (warm-up)");
    }
  }
  *out = grp.release();
  return GAC_OK;
}
extern "C" int gac_group_destroy(gac_group* grp) {
  if (!group_ok(grp)) return fail(GAC_ERR_INVALID_ARGUMENT, "group is null or already destroyed");
  for (gac_context* c : grp->members) gac_context_destroy(c);  // (destroys the member's communicator as well)
  grp->magic = 0;
  delete grp;
  return GAC_OK;
}
extern "C" int gac_group_size(gac_group* grp, int* n) {
  if (!group_ok(grp) || !n) return fail(GAC_ERR_INVALID_ARGUMENT, "bad arguments");
  *n = (int)grp->members.size();
  return GAC_OK;
}
extern "C" int gac_group_context(gac_group* grp, int index, gac_context** ctx) {
  if (!group_ok(grp) || !ctx) return fail(GAC_ERR_INVALID_ARGUMENT, "bad arguments");
  if (index < 0 || index >= (int)grp->members.size()) return fail(GAC_ERR_OUT_OF_RANGE, "member index %d out of range", index);
  *ctx = grp->members[(size_t)index];
  return GAC_OK;
}
extern "C" int gac_group_render(gac_group* grp, const gac_graph* const* shards, int64_t first_frame, int64_t n_frames, float* const* out_channels,
                                int n_out_channels, int64_t start_index) {
  if (!group_ok(grp)) return fail(GAC_ERR_DISPOSED, "group is null or destroyed");
  if (!shards || !out_channels) return fail(GAC_ERR_INVALID_ARGUMENT, "null argument");
  const size_t n = grp->members.size();
  for (size_t i = 0; i < n; i++) {
    if (!shards[i]) return fail(GAC_ERR_INVALID_ARGUMENT, "shard %zu is null (a member without voices still takes part in the reduce: pass a graph with the bus and no voices)", i);
    if (shards[i]->ctx != grp->members[i]) return fail(GAC_ERR_INVALID_ARGUMENT, "shard %zu was not created against member %zu", i, i);
    for (auto& v : shards[i]->voices)
      if (v.bus < 0) return fail(GAC_ERR_UNSUPPORTED, "sharded renders need every voice routed through a bus (the bus is what is reduced)");
    if (shards[i]->buses.empty()) return fail(GAC_ERR_UNSUPPORTED, "sharded renders need at least one bus");
  }
  auto member_render = [&](size_t i) {
    RenderArgs a;
    a.graphs = &shards[i];
    a.n_graphs = 1;
    a.first_frame = first_frame;
    a.n_frames = n_frames;
    a.h_out = i == 0 ? out_channels : nullptr;
    a.n_out = n_out_channels;
    a.start_index = start_index;
    a.sharded = true;
    a.root = 0;
    return render_core(grp->members[i], a);
  };
  if (n == 1) return member_render(0);
  std::vector<int> rcs(n, GAC_OK);
  std::vector<std::string> msgs(n);
  std::vector<std::thread> workers;
  for (size_t i = 0; i < n; i++)
    workers.emplace_back([&, i] {
      rcs[i] = member_render(i);
      if (rcs[i]) msgs[i] = g_err;  // (thread-local: carried over to the caller below)
    });
  for (auto& w : workers) w.join();
  for (size_t i = 0; i < n; i++)
    if (rcs[i]) return fail(rcs[i], "member %zu (device %d): %s", i, grp->members[i]->device, msgs[i].c_str());
  return GAC_OK;
}

extern "C" int gac_get_stats(gac_context* ctx, gac_stats* out) {
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or destroyed");
  if (!out) return fail(GAC_ERR_INVALID_ARGUMENT, "out is null");
  *out = ctx->stats;
  return GAC_OK;
}

// ------------------------------------------------------------------------------------------ kernel-level entry points
// Host pointers in, host pointers out; used by the parity tests to pin each kernel against the oracle.
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  template <typename T> T* as() { return (T*)p; }
};
static int dev_alloc(DevBuf& b, size_t bytes) {
  CU(cudaMalloc(&b.p, std::max<size_t>(bytes, 16)));
  return GAC_OK;
}

extern "C" int gac_rfft_fwd_batch(gac_context* ctx, const float* x, int n_signals, int64_t n_blocks, float* spectra) {
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or destroyed");
  if (!x || !spectra || n_signals <= 0 || n_blocks <= 0) return fail(GAC_ERR_INVALID_ARGUMENT, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  const int B = ctx->B;
  const size_t n = (size_t)n_signals * n_blocks * B;
  DevBuf dx, dX, dj;
  int rc;
  if ((rc = dev_alloc(dx, n * 4)) || (rc = dev_alloc(dX, n * 8)) || (rc = dev_alloc(dj, sizeof(FftFwdJob) * n_signals))) return rc;
  CU(cudaMemcpyAsync(dx.p, x, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  std::vector<FftFwdJob> jobs(n_signals);
  for (int s = 0; s < n_signals; s++) {
    jobs[s] = FftFwdJob{dx.as<float>() + (size_t)s * n_blocks * B, dX.as<float2>() + (size_t)s * n_blocks * B, nullptr, nullptr, 1.0f,
                        n_blocks * B, n_blocks, 0, std::numeric_limits<int64_t>::max()};
  }
  CU(cudaMemcpyAsync(dj.p, jobs.data(), sizeof(FftFwdJob) * n_signals, cudaMemcpyHostToDevice, ctx->stream));
  launch_rfft_fwd(dj.as<FftFwdJob>(), n_signals, n_blocks, B, ctx->d_tw, kstream(ctx));
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(spectra, dX.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return GAC_OK;
}

extern "C" int gac_irfft_ola_batch(gac_context* ctx, const float* Y, int n_signals, int64_t n_blocks, float* y) {
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or destroyed");
  if (!Y || !y || n_signals <= 0 || n_blocks <= 0) return fail(GAC_ERR_INVALID_ARGUMENT, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  const int B = ctx->B;
  const size_t n = (size_t)n_signals * n_blocks * B;
  DevBuf dY, dy, dj;
  int rc;
  if ((rc = dev_alloc(dY, n * 8)) || (rc = dev_alloc(dy, n * 4)) || (rc = dev_alloc(dj, sizeof(FftInvJob) * n_signals))) return rc;
  CU(cudaMemcpyAsync(dY.p, Y, n * 8, cudaMemcpyHostToDevice, ctx->stream));
  std::vector<FftInvJob> jobs(n_signals);
  for (int s = 0; s < n_signals; s++)
    jobs[s] = FftInvJob{dY.as<float2>() + (size_t)s * n_blocks * B, dy.as<float>() + (size_t)s * n_blocks * B, n_blocks};
  CU(cudaMemcpyAsync(dj.p, jobs.data(), sizeof(FftInvJob) * n_signals, cudaMemcpyHostToDevice, ctx->stream));
  launch_irfft_ola(dj.as<FftInvJob>(), n_signals, n_blocks, B, ctx->d_tw, kstream(ctx));
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(y, dy.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return GAC_OK;
}

extern "C" int gac_spectral_mac(gac_context* ctx, const float* X, const float* H, int n_signals, int64_t n_blocks, int n_partitions, int variant,
                                float* Y) {
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or destroyed");
  if (!X || !H || !Y || n_signals <= 0 || n_blocks <= 0 || n_partitions <= 0) return fail(GAC_ERR_INVALID_ARGUMENT, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  const int B = ctx->B;
  if (variant == 3) {
    // second-level FFT path (fft2.cu): transposed spectrograms XT/YT[s][B+1][Qs]; the packed (DC, Nyquist) bin is
    // unpacked into rows 0 and B on the way in and packed again on the way out (host side: this is a test entry point)
    int Lh = 0;
    const int M = fft2_pick_m(n_partitions, &Lh);
    if (M <= 0) return fail(GAC_ERR_UNSUPPORTED, "%d partitions exceed the second-level transform (use variant 0/2)", n_partitions);
    const int C = B + 1;
    const int64_t Qs = ((n_blocks + 15) / 16) * 16;
    const int P16 = std::max(16, ((n_partitions + 15) / 16) * 16);
    std::vector<float2> xt((size_t)n_signals * C * Qs, make_float2(0.f, 0.f));
    const float2* Xh = reinterpret_cast<const float2*>(X);
    for (int s = 0; s < n_signals; s++)
      for (int64_t b = 0; b < n_blocks; b++)
        for (int k = 0; k < B; k++) {
          const float2 v = Xh[((size_t)s * n_blocks + b) * B + k];
          if (k == 0) {
            xt[((size_t)s * C + 0) * Qs + b] = make_float2(v.x, 0.f);
            xt[((size_t)s * C + B) * Qs + b] = make_float2(v.y, 0.f);
          } else {
            xt[((size_t)s * C + k) * Qs + b] = v;
          }
        }
    DevBuf dXT, dYT, dH, dH2, dj;
    int rc;
    if ((rc = dev_alloc(dXT, xt.size() * 8)) || (rc = dev_alloc(dYT, xt.size() * 8)) || (rc = dev_alloc(dH, (size_t)n_signals * P16 * B * 8)) ||
        (rc = dev_alloc(dH2, (size_t)n_signals * C * fft2_h2_row_elems(M) * 8)) || (rc = dev_alloc(dj, sizeof(Fft2Job) * n_signals)))
      return rc;
    CU(cudaMemcpyAsync(dXT.p, xt.data(), xt.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(dH.p, 0, (size_t)n_signals * P16 * B * 8, ctx->stream));
    for (int s = 0; s < n_signals; s++)
      CU(cudaMemcpyAsync(dH.as<float2>() + (size_t)s * P16 * B, H + (size_t)s * n_partitions * B * 2, (size_t)n_partitions * B * 8,
                         cudaMemcpyHostToDevice, ctx->stream));
    launch_fft2_prep(dH.as<float2>(), (int64_t)P16 * B, n_signals, B, n_partitions, M, dH2.as<float2>(), ctx->d_tw2, ctx->d_tab16, kstream(ctx));
    const int V = M - Lh;
    const int nseg = (int)((n_blocks + V - 1) / V);
    std::vector<Fft2Job> jobs(n_signals);
    for (int s = 0; s < n_signals; s++)
      jobs[s] = Fft2Job{dXT.as<float2>() + (size_t)s * C * Qs, dH2.as<float2>() + (size_t)s * C * fft2_h2_row_elems(M), dYT.as<float2>() + (size_t)s * C * Qs, Lh, nseg};
    CU(cudaMemcpyAsync(dj.p, jobs.data(), sizeof(Fft2Job) * n_signals, cudaMemcpyHostToDevice, ctx->stream));
    launch_fft2_conv(dj.as<Fft2Job>(), n_signals, nseg, C, M, ctx->d_tw2, ctx->d_tab16, n_blocks, Qs, Qs, kstream(ctx));
    CU(cudaGetLastError());
    std::vector<float2> yt(xt.size());
    CU(cudaMemcpyAsync(yt.data(), dYT.p, yt.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float2* Yh = reinterpret_cast<float2*>(Y);
    for (int s = 0; s < n_signals; s++)
      for (int64_t b = 0; b < n_blocks; b++)
        for (int k = 0; k < B; k++) {
          float2 v = yt[((size_t)s * C + k) * Qs + b];
          if (k == 0) v.y = yt[((size_t)s * C + B) * Qs + b].x;
          Yh[((size_t)s * n_blocks + b) * B + k] = v;
        }
    return GAC_OK;
  }
  const int TB = ctx->tile_blocks;
  const int64_t QBpad = ((n_blocks + TB - 1) / TB) * TB;
  const int64_t rowsX = TB + QBpad;
  const int P16 = std::max(16, ((n_partitions + 15) / 16) * 16);
  const int groups = B / 128;
  DevBuf dX, dH, dY, dj, dt;
  int rc;
  if ((rc = dev_alloc(dX, (size_t)n_signals * rowsX * B * 8)) || (rc = dev_alloc(dH, (size_t)n_signals * P16 * B * 8)) ||
      (rc = dev_alloc(dY, (size_t)n_signals * QBpad * B * 8)))
    return rc;
  CU(cudaMemsetAsync(dX.p, 0, (size_t)n_signals * rowsX * B * 8, ctx->stream));
  CU(cudaMemsetAsync(dH.p, 0, (size_t)n_signals * P16 * B * 8, ctx->stream));
  std::vector<MacJob> jobs;
  std::vector<MacTile> tiles;
  for (int s = 0; s < n_signals; s++) {
    float2* Xs = dX.as<float2>() + (size_t)s * rowsX * B + (size_t)TB * B;
    float2* Hs = dH.as<float2>() + (size_t)s * P16 * B;
    float2* Ys = dY.as<float2>() + (size_t)s * QBpad * B;
    CU(cudaMemcpyAsync(Xs, X + (size_t)s * n_blocks * B * 2, (size_t)n_blocks * B * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(Hs, H + (size_t)s * n_partitions * B * 2, (size_t)n_partitions * B * 8, cudaMemcpyHostToDevice, ctx->stream));
    for (int g = 0; g < groups; g++) jobs.push_back(MacJob{Xs + g * 128, Hs + g * 128, Ys + g * 128, n_partitions, g == 0});
  }
  for (int j = 0; j < (int)jobs.size(); j++)
    for (int64_t b0 = 0; b0 < n_blocks; b0 += TB) tiles.push_back(MacTile{j, (int)b0});
  if ((rc = dev_alloc(dj, sizeof(MacJob) * jobs.size())) || (rc = dev_alloc(dt, sizeof(MacTile) * tiles.size()))) return rc;
  CU(cudaMemcpyAsync(dj.p, jobs.data(), sizeof(MacJob) * jobs.size(), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dt.p, tiles.data(), sizeof(MacTile) * tiles.size(), cudaMemcpyHostToDevice, ctx->stream));
  if (variant == 1)
    launch_mac_stream(dj.as<MacJob>(), (int)jobs.size(), n_blocks, B, kstream(ctx));
  else
    launch_mac_tiled(dj.as<MacJob>(), (int)jobs.size(), dt.as<MacTile>(), (int)tiles.size(), n_blocks, n_partitions, B, TB, variant == 2 ? 2 : 0, kstream(ctx));
  CU(cudaGetLastError());
  for (int s = 0; s < n_signals; s++)
    CU(cudaMemcpyAsync(Y + (size_t)s * n_blocks * B * 2, dY.as<float2>() + (size_t)s * QBpad * B, (size_t)n_blocks * B * 8, cudaMemcpyDeviceToHost,
                       ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return GAC_OK;
}

extern "C" int gac_convolve_batch(gac_context* ctx, const float* x, int n_signals, int64_t n_frames, const float* ir, int64_t ir_frames,
                                  int normalize, float* y) {
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or destroyed");
  if (!x || !ir || !y || n_signals <= 0 || n_frames <= 0 || ir_frames <= 0) return fail(GAC_ERR_INVALID_ARGUMENT, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  const int B = ctx->B;
  const int64_t Npad = ((n_frames + B - 1) / B) * B;
  // impulse responses: one n_signals-channel "buffer"
  DevBuf dir;
  int rc;
  const int64_t stride = ((ir_frames + 8 + 63) / 64) * 64;
  if ((rc = dev_alloc(dir, (size_t)n_signals * stride * 4))) return rc;
  CU(cudaMemsetAsync(dir.p, 0, (size_t)n_signals * stride * 4, ctx->stream));
  for (int s = 0; s < n_signals; s++)
    CU(cudaMemcpyAsync(dir.as<float>() + (size_t)s * stride, ir + (size_t)s * ir_frames, (size_t)ir_frames * 4, cudaMemcpyHostToDevice, ctx->stream));
  gac_ir irh{};
  irh.ctx = ctx;
  irh.nch = n_signals;
  rc = ir_prepare_device(ctx, dir.as<float>(), stride, n_signals, ir_frames, normalize != 0, &irh);
  DevBuf holdH;
  holdH.p = irh.d_H;  // (the second-level spectra and the scales live in the same allocation)
  if (rc) return rc;
  DevBuf dx;
  if ((rc = dev_alloc(dx, (size_t)n_signals * Npad * 4))) return rc;
  CU(cudaMemsetAsync(dx.p, 0, (size_t)n_signals * Npad * 4, ctx->stream));
  for (int s = 0; s < n_signals; s++)
    CU(cudaMemcpyAsync(dx.as<float>() + (size_t)s * Npad, x + (size_t)s * n_frames, (size_t)n_frames * 4, cudaMemcpyHostToDevice, ctx->stream));
  {
    RenderEnv env;
    Scratch scratch(ctx);
    HostKeep keep;
    Timer timer(ctx);
    env.ctx = ctx;
    env.scratch = &scratch;
    env.keep = &keep;
    env.timer = &timer;
    env.Npad = Npad;
    env.NQ = Npad / 128;
    env.QB = Npad / B;
    std::vector<ConvItem> items(n_signals);
    for (int s = 0; s < n_signals; s++)
    {
      ConvItem& it = items[s];
      float* xs = dx.as<float>() + (size_t)s * Npad;
      it.n_fwd = 1;
      it.fwd[0] = {xs, nullptr, 1.0f};
      it.lo = 0;
      it.hi = Npad;
      it.n_mac = 1;
      it.mac[0] = {0, irh.d_H + (size_t)s * irh.P16 * B, irh.d_H2 ? irh.d_H2 + (size_t)s * (B + 1) * fft2_h2_row_elems(irh.M2) : nullptr};
      it.P = irh.P;
      it.M2 = irh.d_H2 ? irh.M2 : 0;
      it.Lh = irh.Lh;
      it.n_inv = 1;
      it.inv[0] = {0, -1, xs, nullptr};
    }
    rc = conv_batch(env, items);
    gac_stats st{};
    timer.finish(&st);
    cudaStreamSynchronize(ctx->stream);
    if (rc) return rc;
    st.conv_units = env.conv_units;
    st.algorithmic_bytes = env.alg_bytes;
    st.mac_complex_macs = env.macs;
    st.mac_flops = env.mac_flops;
    st.mac_bytes_moved = env.mac_bytes;
    st.mac_variant_used = env.mac_used;
    st.kernel_launches = env.launches;
    st.voices = n_signals;
    st.frames = n_frames;
    ctx->stats = st;
  }
  for (int s = 0; s < n_signals; s++)
    CU(cudaMemcpyAsync(y + (size_t)s * n_frames, dx.as<float>() + (size_t)s * Npad, (size_t)n_frames * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return GAC_OK;
}

extern "C" int gac_automation_eval(gac_context* ctx, const gac_param* param, int a_rate, int64_t n_frames, float* values) {
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or destroyed");
  if (!param || !values || n_frames <= 0) return fail(GAC_ERR_INVALID_ARGUMENT, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  ParamH p;
  int rc = copy_param(*param, &p, "param");
  if (rc) return rc;
  const int64_t NQ = (n_frames + 127) / 128;
  if ((rc = ensure_block_times(ctx, NQ + 1))) return rc;
  DevBuf dev, dout, dj;
  if ((rc = dev_alloc(dev, sizeof(DevEvent) * std::max<size_t>(1, p.ev.size()))) || (rc = dev_alloc(dout, (size_t)NQ * 128 * 4)) ||
      (rc = dev_alloc(dj, sizeof(ParamJob))))
    return rc;
  if (!p.ev.empty()) CU(cudaMemcpyAsync(dev.p, p.ev.data(), sizeof(gac_event) * p.ev.size(), cudaMemcpyHostToDevice, ctx->stream));
  ParamJob j{p.value, (int)p.ev.size(), dev.as<DevEvent>(), dout.as<float>(), a_rate ? 1 : 0};
  CU(cudaMemcpyAsync(dj.p, &j, sizeof(j), cudaMemcpyHostToDevice, ctx->stream));
  launch_param_eval(dj.as<ParamJob>(), 1, ctx->d_bt, NQ, ctx->fs, kstream(ctx));
  CU(cudaGetLastError());
  if (a_rate) {
    CU(cudaMemcpyAsync(values, dout.p, (size_t)n_frames * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  } else {
    // k-rate: one value per quantum, repeated 128 times by the reference (AudioParam.cs:164)
    std::vector<float> q((size_t)NQ);
    CU(cudaMemcpyAsync(q.data(), dout.p, (size_t)NQ * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int64_t n = 0; n < n_frames; n++) values[n] = q[(size_t)(n >> 7)];
  }
  return GAC_OK;
}

extern "C" int gac_resample_cubic(gac_context* ctx, const float* in, int64_t n_in, double rate, int64_t n_out, float* out, int64_t* produced,
                                  int64_t* consumed) {
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or destroyed");
  if (!in || !out || n_in < 0 || n_out <= 0 || !(rate > 0)) return fail(GAC_ERR_INVALID_ARGUMENT, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  // phase recurrence on the host (CubicResampler.cs:40-60), interpolation on the device
  std::vector<int32_t> k;
  std::vector<float> t;
  int64_t inPos = 0, m = 0;
  if (n_in >= 4) {
    inPos = 4;
    double Pos = 0.0;
    while (m < n_out) {
      int consume = (int)Pos;
      if (inPos + consume > n_in) break;
      inPos += consume;
      Pos -= consume;
      k.push_back((int32_t)(inPos - 4));
      t.push_back((float)Pos);
      Pos += rate;
      m++;
    }
  } else {
    inPos = n_in;
  }
  if (produced) *produced = m;
  if (consumed) *consumed = inPos;
  for (int64_t i = 0; i < n_out; i++) out[i] = 0.f;
  if (m == 0) return GAC_OK;
  const int64_t n_pad = ((m + 3) / 4) * 4;
  DevBuf din, dk, dt, dout, dj;
  int rc;
  if ((rc = dev_alloc(din, (size_t)(n_in + 8) * 4)) || (rc = dev_alloc(dk, (size_t)m * 4)) || (rc = dev_alloc(dt, (size_t)m * 4)) ||
      (rc = dev_alloc(dout, (size_t)n_pad * 2 * 4)) || (rc = dev_alloc(dj, sizeof(ResampleJob))))
    return rc;
  CU(cudaMemcpyAsync(din.p, in, (size_t)n_in * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dk.p, k.data(), (size_t)m * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dt.p, t.data(), (size_t)m * 4, cudaMemcpyHostToDevice, ctx->stream));
  ResampleJob j{};
  j.src[0] = j.src[1] = din.as<float>();
  j.dst[0] = dout.as<float>();
  j.dst[1] = dout.as<float>() + n_pad;
  j.k = dk.as<int32_t>();
  j.t = dt.as<float>();
  j.out0 = 0;
  j.n_emit = m;
  j.n_zero_from = m;
  CU(cudaMemcpyAsync(dj.p, &j, sizeof(j), cudaMemcpyHostToDevice, ctx->stream));
  launch_resample(dj.as<ResampleJob>(), 1, m, kstream(ctx));
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, dout.p, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return GAC_OK;
}

// BiQuadFilterNode.Process for n_signals independent stereo signals through the production kernels (k_biquad_select / _entry /
// _resolve / _lanes / _verify behind run_chains): x, y are [n_signals][2][n_frames] host arrays; frequency / q / gain are one
// gac_param per signal (a-rate, a-rate, k-rate; events evaluated by k_param_eval exactly as in a render)
extern "C" int gac_biquad_batch(gac_context* ctx, const float* x, int n_signals, int64_t n_frames, const int* filter_types,
                                const gac_param* frequency, const gac_param* q, const gac_param* gain_db, float* y) {
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or destroyed");
  if (!x || !y || !filter_types || !frequency || !q || !gain_db || n_signals <= 0 || n_frames <= 0) return fail(GAC_ERR_INVALID_ARGUMENT, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  const int64_t Npad = ((n_frames + 127) / 128) * 128;
  std::vector<std::vector<OpH>> ops((size_t)n_signals);
  int rc;
  for (int s = 0; s < n_signals; s++) {
    ops[s].resize(1);
    OpH& o = ops[s][0];
    o.kind = GAC_OP_BIQUAD;
    o.ftype = filter_types[s];
    if (o.ftype < 0 || o.ftype > 7) return fail(GAC_ERR_INVALID_ARGUMENT, "bad filter type %d", o.ftype);
    if ((rc = copy_param(frequency[s], &o.p0, "biquad.frequency")) || (rc = copy_param(q[s], &o.p1, "biquad.Q")) ||
        (rc = copy_param(gain_db[s], &o.p2, "biquad.gain")))
      return rc;
  }
  if ((rc = ensure_block_times(ctx, Npad / 128 + 1))) return rc;
  DevBuf dx;
  if ((rc = dev_alloc(dx, (size_t)n_signals * 2 * Npad * 4))) return rc;
  CU(cudaMemsetAsync(dx.p, 0, (size_t)n_signals * 2 * Npad * 4, ctx->stream));
  CU(cudaMemcpy2DAsync(dx.p, (size_t)Npad * 4, x, (size_t)n_frames * 4, (size_t)n_frames * 4, (size_t)n_signals * 2, cudaMemcpyHostToDevice, ctx->stream));
  {
    if (ctx->kick_pending) {
      cudaEventSynchronize(ctx->kick_done);
      ctx->kick_pending = false;
    }
    ctx->stage_block = 0;
    ctx->stage_used = 0;
    RenderEnv env;
    Scratch scratch(ctx);
    HostKeep keep;
    Timer timer(ctx);
    env.ctx = ctx;
    env.scratch = &scratch;
    env.keep = &keep;
    env.timer = &timer;
    env.Npad = Npad;
    env.NQ = Npad / 128;
    env.QB = Npad / ctx->B;
    std::vector<Sig> sigs((size_t)n_signals);
    for (int s = 0; s < n_signals; s++) {
      sigs[s].p[0] = dx.as<float>() + ((size_t)s * 2 + 0) * Npad;
      sigs[s].p[1] = dx.as<float>() + ((size_t)s * 2 + 1) * Npad;
      sigs[s].lo = 0;
      sigs[s].hi = Npad;
      sigs[s].ops = &ops[s];
    }
    rc = run_chains(env, sigs);
    gac_stats st{};
    timer.finish(&st);
    cudaStreamSynchronize(ctx->stream);
    if (rc) return rc;
    st.kernel_launches = env.launches;
    st.voices = n_signals;
    st.frames = n_frames;
    ctx->stats = st;
  }
  CU(cudaMemcpy2DAsync(y, (size_t)n_frames * 4, dx.p, (size_t)Npad * 4, (size_t)n_frames * 4, (size_t)n_signals * 2, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return GAC_OK;
}

// AudioNodeInput.Pull's fan-in sum (AudioNodeInput.cs:118-137) through k_mix: out[c][n] = ((0 + in_0[c][n]) + in_1[c][n]) + ... over the
// inputs that are non-silent at n (frames [lo[i], hi[i]), multiples of 128), in the given order.  inputs[i] -> [2][n_frames] host
// rows; downmix[i] != 0 mixes input i to one channel first ((L + R) * downmix[i] into both rows, :214-228).  downmix may be null.
extern "C" int gac_mix(gac_context* ctx, const float* const* inputs, const int64_t* lo, const int64_t* hi, const float* downmix, int n_inputs,
                       int64_t n_frames, float* out) {
  if (!ctx_ok(ctx)) return fail(GAC_ERR_DISPOSED, "context is null or destroyed");
  if (!inputs || !lo || !hi || !out || n_inputs < 0 || n_frames <= 0) return fail(GAC_ERR_INVALID_ARGUMENT, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  const int64_t Npad = ((n_frames + 127) / 128) * 128;
  DevBuf din, dout, dj, di;
  int rc;
  if ((rc = dev_alloc(din, (size_t)std::max(1, n_inputs) * 2 * Npad * 4)) || (rc = dev_alloc(dout, (size_t)2 * Npad * 4)) ||
      (rc = dev_alloc(dj, sizeof(MixJob))) || (rc = dev_alloc(di, sizeof(MixInput) * std::max(1, n_inputs))))
    return rc;
  CU(cudaMemsetAsync(din.p, 0, (size_t)std::max(1, n_inputs) * 2 * Npad * 4, ctx->stream));
  std::vector<MixInput> mi((size_t)n_inputs);
  for (int i = 0; i < n_inputs; i++) {
    if (!inputs[i]) return fail(GAC_ERR_INVALID_ARGUMENT, "input %d is null", i);
    if (lo[i] < 0 || hi[i] > Npad || lo[i] % 128 || hi[i] % 128) return fail(GAC_ERR_INVALID_ARGUMENT, "input %d: the non-silent range must consist of whole quanta", i);
    float* d = din.as<float>() + (size_t)i * 2 * Npad;
    CU(cudaMemcpy2DAsync(d, (size_t)Npad * 4, inputs[i], (size_t)n_frames * 4, (size_t)n_frames * 4, 2, cudaMemcpyHostToDevice, ctx->stream));
    mi[i].src[0] = d;
    mi[i].src[1] = d + Npad;
    mi[i].lo = lo[i];
    mi[i].hi = hi[i];
    mi[i].downmix = downmix ? downmix[i] : 0.f;
  }
  MixJob mj;
  mj.dst[0] = dout.as<float>();
  mj.dst[1] = dout.as<float>() + Npad;
  mj.first_input = 0;
  mj.n_inputs = n_inputs;
  if (n_inputs) CU(cudaMemcpyAsync(di.p, mi.data(), sizeof(MixInput) * mi.size(), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dj.p, &mj, sizeof(mj), cudaMemcpyHostToDevice, ctx->stream));
  launch_mix(dj.as<MixJob>(), 1, di.as<MixInput>(), Npad, kstream(ctx));
  CU(cudaGetLastError());
  CU(cudaMemcpy2DAsync(out, (size_t)n_frames * 4, dout.p, (size_t)Npad * 4, (size_t)n_frames * 4, 2, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return GAC_OK;
}
