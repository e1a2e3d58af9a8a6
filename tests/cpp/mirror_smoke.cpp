// mirror_smoke.cpp — drives libgraphaudio_cuda.so through the C++ mirror of the reference API
// (graphaudio_b200/host/graphaudio_cuda.hpp).  Usage: mirror_smoke <dir> <n_voices> <src_frames> <ir_frames> <n_frames>
// Reads <dir>/src_<v>_<c>.f32 and <dir>/ir_<v>_<c>.f32 (little-endian float32), builds the C2-shaped graph
//   AudioBufferSourceNode -> GainNode(automation) -> ConvolverNode -> bus GainNode(0.25) -> Destination
// renders n_frames in two Render calls and writes <dir>/out_<c>.f32.  Exit code 3 = no device (no CPU fallback).
#include <cstdio>
#include <string>
#include <vector>

#include "../../graphaudio_b200/host/graphaudio_cuda.hpp"

using namespace GraphAudio::Cuda;

static std::vector<float> read_f32(const std::string& path, size_t n) {
  std::vector<float> v(n);
  FILE* f = fopen(path.c_str(), "rb");
  if (!f || fread(v.data(), sizeof(float), n, f) != n) { fprintf(stderr, "cannot read %s\n", path.c_str()); exit(2); }
  fclose(f);
  return v;
}

int main(int argc, char** argv) {
  if (argc < 6) return 2;
  const std::string dir = argv[1];
  const int nv = atoi(argv[2]);
  const size_t ns = (size_t)atol(argv[3]), ni = (size_t)atol(argv[4]);
  const int n = atoi(argv[5]);
  try {
    OfflineAudioContext ctx(48000);
    auto bus = ctx.CreateGain();
    bus->Gain.SetValue(0.25f);
    bus->Connect(ctx.Destination());
    for (int v = 0; v < nv; v++) {
      auto src = ctx.CreateBufferSource();
      src->Buffer = PlayableAudioBuffer::FromStereoArrays(read_f32(dir + "/src_" + std::to_string(v) + "_0.f32", ns),
                                                          read_f32(dir + "/src_" + std::to_string(v) + "_1.f32", ns), 48000);
      auto gain = ctx.CreateGain();
      gain->Gain.SetValueAtTime(0.9f, 0.0);
      gain->Gain.LinearRampToValueAtTime(0.3f, 0.05 * (v + 1));
      gain->Gain.ExponentialRampToValueAtTime(0.8f, 0.2);
      gain->Gain.SetTargetAtTime(0.0f, 0.25, 0.05);
      auto conv = ctx.CreateConvolver();
      conv->SetBuffer(PlayableAudioBuffer::FromStereoArrays(read_f32(dir + "/ir_" + std::to_string(v) + "_0.f32", ni),
                                                            read_f32(dir + "/ir_" + std::to_string(v) + "_1.f32", ni), 48000));
      src->Connect(gain)->Connect(conv)->Connect(bus);
      src->Start();
    }
    std::vector<std::vector<float>> out(2, std::vector<float>((size_t)n));
    float* rows[2] = {out[0].data(), out[1].data()};
    const int n1 = n / 3 + 7;
    ctx.Render(rows, 2, n1, 0);          // successive Render calls continue the timeline (OfflineAudioContext.cs:55-100)
    ctx.Render(rows, 2, n - n1, n1);
    for (int c = 0; c < 2; c++) {
      FILE* f = fopen((dir + "/out_" + std::to_string(c) + ".f32").c_str(), "wb");
      fwrite(out[c].data(), sizeof(float), (size_t)n, f);
      fclose(f);
    }
    gac_stats st = ctx.LastStats();
    printf("mirror_smoke: %d voices, %d frames, %.3f ms, %lld launches\n", nv, n, st.ms_total, (long long)st.kernel_launches);
    try {
      ctx.Render(rows, 2, 0, 0);
      return 4;
    } catch (const ArgumentOutOfRangeException&) {  // "Frame count must be positive."
    }
  } catch (const CudaException& e) {
    fprintf(stderr, "CudaException: %s\n", e.what());
    return 3;
  }
  return 0;
}
