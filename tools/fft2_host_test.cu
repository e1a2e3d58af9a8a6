// Host replay of the per-thread phases in fft2_core.cuh: validates the index algebra / butterflies without a GPU.
// nvcc -O2 -std=c++17 -Igraphaudio_b200/csrc tools/fft2_host_test.cu -o /tmp/fft2_host_test && /tmp/fft2_host_test
#include <cmath>
#include <complex>
#include <cstdio>
#include <random>
#include <vector>
#include "fft2_core.cuh"
using namespace gac::f2;
typedef std::complex<double> cd;

template <int M, int S> void fwd_all(std::vector<float2>& sm, const float2* tw) {
  if constexpr (S > 8) {
    for (int t = 0; t < M / 8; t++) { float2 v[8]; fwd_stage<S, true>(v, sm.data(), tw, t); }
    fwd_all<M, S / 8>(sm, tw);
  }
}
template <int M, int S> void inv_all(std::vector<float2>& sm, const float2* tw, std::vector<float2>& out) {
  if constexpr (S < M) {
    for (int t = 0; t < M / 8; t++) { float2 v[8]; inv_stage<S, true>(v, sm.data(), tw, t); }
    inv_all<M, S * 8>(sm, tw, out);
  } else {
    for (int t = 0; t < M / 8; t++) { float2 v[8]; inv_stage<M, false>(v, sm.data(), tw, t); for (int j = 0; j < 8; j++) out[t + (M / 8) * j] = v[j]; }
  }
}
template <int M> std::vector<float2> forward_points(const std::vector<float2>& x, const float2* tw) {
  using P = Plan<M>;
  std::vector<float2> sm(smem_elems(M)), pts(M);
  for (int t = 0; t < M / 8; t++) { float2 v[8]; for (int j = 0; j < 8; j++) v[j] = x[t + (M / 8) * j]; fwd_stage<M, false>(v, sm.data(), tw, t); }
  fwd_all<M, M / 8>(sm, tw);
  for (int t = 0; t < M / 8; t++) { float2 u[8]; to_points<P::TAIL>(u, sm.data(), t); for (int q = 0; q < 8; q++) pts[8 * t + q] = u[q]; }
  return pts;
}
template <int M> double run() {
  using P = Plan<M>;
  std::vector<float2> tw(kTwLen);
  for (int e = 0; e < kTwLen; e++) { double a = -2.0 * M_PI * e / kTwLen; tw[e] = make_float2((float)cos(a), (float)sin(a)); }
  std::mt19937 rng(M);
  std::uniform_real_distribution<float> U(-1.f, 1.f);
  std::vector<float2> x(M), h(M);
  for (auto& v : x) v = make_float2(U(rng), U(rng));
  int Pn = M / 3;
  for (int i = 0; i < M; i++) h[i] = i < Pn ? make_float2(U(rng), U(rng)) : make_float2(0, 0);
  auto X = forward_points<M>(x, tw.data());
  auto H = forward_points<M>(h, tw.data());
  // check the forward transform is a permutation of the DFT: compare sorted energy + sum
  std::vector<float2> sm(smem_elems(M)), out(M);
  for (int t = 0; t < M / 8; t++) {
    float2 u[8];
    for (int q = 0; q < 8; q++) u[q] = cmulf(X[8 * t + q], make_float2(H[8 * t + q].x / M, H[8 * t + q].y / M));
    from_points<P::TAIL>(u, sm.data(), t);
  }
  inv_all<M, (P::TAIL == 1 ? 64 : P::SL)>(sm, tw.data(), out);
  // reference circular convolution in double
  double err = 0, mag = 0;
  for (int n = 0; n < M; n++) {
    cd acc = 0;
    for (int p = 0; p < Pn; p++) { int m = (n - p + M) % M; acc += cd(x[m].x, x[m].y) * cd(h[p].x, h[p].y); }
    err = std::max(err, std::abs(acc - cd(out[n].x, out[n].y)));
    mag = std::max(mag, std::abs(acc));
  }
  printf("M=%d  max|err| %.3e  (max|y| %.2f, rel %.2e)\n", M, err, mag, err / mag);
  return err / mag;
}
namespace t16 {
namespace R = gac::r16;
using R::pad; using R::smem_elems; using R::table_elems; using R::fwd_a; using R::fwd_b; using R::inv_a; using R::inv_b; using R::load16; using R::store16; using R::stage_c;
template <int M> std::vector<float2> make_table() {
  constexpr int T = R::Plan<M>::T, L = R::Plan<M>::L;
  std::vector<float2> tab(table_elems(M));
  for (int q = 0; q < 4; q++) {
    for (int t = 0; t < T; t++) { double a = -2.0 * M_PI * (double)(t << q) / M; tab[q * T + t] = make_float2((float)cos(a), (float)sin(a)); }
    if (R::Plan<M>::HAS_B) for (int i = 0; i < L; i++) { double a = -2.0 * M_PI * (double)(i << q) / (16 * L); tab[4 * T + q * L + i] = make_float2((float)cos(a), (float)sin(a)); }
  }
  return tab;
}
template <int M> std::vector<float2> forward_points(const std::vector<float2>& x, const float2* tab) {
  constexpr int T = R::Plan<M>::T, L = R::Plan<M>::L;
  std::vector<float2> sm(smem_elems(M)), pts(M);
  for (int t = 0; t < T; t++) { float2 v[16]; for (int j = 0; j < 16; j++) v[j] = x[t + T * j]; fwd_a<M>(v, sm.data(), tab, t); }
  if constexpr (R::Plan<M>::HAS_B) for (int t = 0; t < T; t++) fwd_b<M>(sm.data(), tab, t);
  for (int t = 0; t < T; t++) { float2 u[16]; load16(u, sm.data(), t); stage_c<L, false>(u); for (int q = 0; q < 16; q++) pts[16 * t + q] = u[q]; }
  return pts;
}
template <int M> double run() {
  constexpr int T = R::Plan<M>::T, L = R::Plan<M>::L;
  auto tab = make_table<M>();
  std::mt19937 rng(M + 7);
  std::uniform_real_distribution<float> U(-1.f, 1.f);
  std::vector<float2> x(M), h(M);
  for (auto& v : x) v = make_float2(U(rng), U(rng));
  int Pn = M / 3;
  for (int i = 0; i < M; i++) h[i] = i < Pn ? make_float2(U(rng), U(rng)) : make_float2(0, 0);
  auto X = t16::forward_points<M>(x, tab.data());
  auto H = t16::forward_points<M>(h, tab.data());
  std::vector<float2> sm(smem_elems(M)), out(M);
  for (int t = 0; t < T; t++) {
    float2 u[16];
    for (int q = 0; q < 16; q++) u[q] = cmulf(X[16 * t + q], make_float2(H[16 * t + q].x / M, H[16 * t + q].y / M));
    stage_c<L, true>(u);
    store16(u, sm.data(), t);
  }
  if constexpr (R::Plan<M>::HAS_B) for (int t = 0; t < T; t++) inv_b<M>(sm.data(), tab.data(), t);
  for (int t = 0; t < T; t++) { float2 v[16]; inv_a<M>(v, sm.data(), tab.data(), t); for (int j = 0; j < 16; j++) out[t + T * j] = v[j]; }
  double err = 0, mag = 0;
  for (int n = 0; n < M; n++) {
    cd acc = 0;
    for (int p = 0; p < Pn; p++) { int m = (n - p + M) % M; acc += cd(x[m].x, x[m].y) * cd(h[p].x, h[p].y); }
    err = std::max(err, std::abs(acc - cd(out[n].x, out[n].y)));
    mag = std::max(mag, std::abs(acc));
  }
  printf("r16 M=%d  max|err| %.3e  (max|y| %.2f, rel %.2e)\n", M, err, mag, err / mag);
  return err / mag;
}
}  // namespace t16
// frequency index held by (thread t, slot q) after stage C of the M = 128 plan: k = 2t + h + 16 rev3(s), q = 8h + s
static int check_map128() {
  using namespace t16;
  auto tab = make_table<128>();
  std::mt19937 rng(5);
  std::uniform_real_distribution<float> U(-1.f, 1.f);
  std::vector<float2> x(128);
  for (auto& v : x) v = make_float2(U(rng), U(rng));
  auto X = t16::forward_points<128>(x, tab.data());
  double worst = 0;
  for (int t = 0; t < 8; t++)
    for (int q = 0; q < 16; q++) {
      int h = q / 8, s = q % 8, k = 2 * t + h + 16 * gac::f2::rev3(s);
      cd acc = 0;
      for (int n = 0; n < 128; n++) acc += cd(x[n].x, x[n].y) * std::polar(1.0, -2.0 * M_PI * k * n / 128.0);
      worst = std::max(worst, std::abs(acc - cd(X[16 * t + q].x, X[16 * t + q].y)));
    }
  printf("map128: worst |X[k(t,q)] - DFT| = %.3e\n", worst);
  return worst < 1e-4 ? 0 : 1;
}
// the same for the M = 256 plan (k = t + 16 rev4(q)) and the M = 512 plan (k = (t >> 1) + 16 (8 (t & 1) + (q >> 1)) + 256 (q & 1))
template <int M> static int check_map_generic() {
  using namespace t16;
  auto tab = make_table<M>();
  std::mt19937 rng(M);
  std::uniform_real_distribution<float> U(-1.f, 1.f);
  std::vector<float2> x(M);
  for (auto& v : x) v = make_float2(U(rng), U(rng));
  auto X = t16::forward_points<M>(x, tab.data());
  double worst = 0;
  for (int t = 0; t < M / 16; t++)
    for (int q = 0; q < 16; q++) {
      int k = M == 256 ? t + 16 * gac::r16::rev4(q) : (t >> 1) + 16 * (8 * (t & 1) + (q >> 1)) + 256 * (q & 1);
      cd acc = 0;
      for (int n = 0; n < M; n++) acc += cd(x[n].x, x[n].y) * std::polar(1.0, -2.0 * M_PI * (double)k * n / M);
      worst = std::max(worst, std::abs(acc - cd(X[16 * t + q].x, X[16 * t + q].y)));
    }
  printf("map%d: worst |X[k(t,q)] - DFT| = %.3e\n", M, worst);
  return worst < 1e-3 ? 0 : 1;
}
int main() {
  if (check_map128()) return 3;
  if (check_map_generic<256>()) return 4;
  if (check_map_generic<512>()) return 5;
  {
    double e = 0;
    e = std::max(e, t16::run<128>());
    e = std::max(e, t16::run<256>());
    e = std::max(e, t16::run<512>());
    e = std::max(e, t16::run<1024>());
    e = std::max(e, t16::run<2048>());
    e = std::max(e, t16::run<4096>());
    if (!(e < 2e-6)) return 2;
  }
  double e = 0;
  e = std::max(e, run<512>());
  e = std::max(e, run<1024>());
  e = std::max(e, run<2048>());
  e = std::max(e, run<4096>());
  e = std::max(e, run<8192>());
  return e < 2e-6 ? 0 : 1;
}
