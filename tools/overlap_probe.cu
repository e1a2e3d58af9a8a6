// Does work on a second stream start while a long queue of pinned H2D copies is in flight on the first?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstring>
#include <chrono>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void k_touch(float* p, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] += 1.f; }
__global__ void k_copy(float* d, const float* s, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) d[i] = s[i]; }
int main() {
  const int NB = 64; const size_t bytes = 4608000;
  cudaMemPool_t pool; CK(cudaDeviceGetDefaultMemPool(&pool, 0)); uint64_t thr = UINT64_MAX; CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  char* h; CK(cudaHostAlloc((void**)&h, NB * bytes, cudaHostAllocDefault)); memset(h, 1, NB * bytes);
  std::vector<float> pageable(300000, 1.f);
  for (int variant = 1; variant < 10; variant++) {
    for (int rep = 0; rep < 5; rep++) {
      cudaStream_t A, B; CK(cudaStreamCreateWithFlags(&B, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&A, cudaStreamNonBlocking));
      cudaEvent_t start, t0, t1, t2, done, ev[NB]; CK(cudaEventCreate(&start)); CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1)); CK(cudaEventCreate(&t2)); CK(cudaEventCreate(&done));
      for (int i = 0; i < NB; i++) CK(cudaEventCreateWithFlags(&ev[i], variant == 5 ? cudaEventDefault : cudaEventDisableTiming));
      float* d[NB]; float* small = nullptr; float* tw = nullptr;
      if (variant == 2 || variant == 4) {  // a pageable copy on B first (context creation)
        CK(cudaMallocAsync(&tw, pageable.size() * 4, B));
        CK(cudaMemcpyAsync(tw, pageable.data(), pageable.size() * 4, cudaMemcpyHostToDevice, B));
      }
      if (variant == 6) {  // pageable copy on B, then synchronise B
        CK(cudaMallocAsync(&tw, pageable.size() * 4, B));
        CK(cudaMemcpyAsync(tw, pageable.data(), pageable.size() * 4, cudaMemcpyHostToDevice, B));
        CK(cudaStreamSynchronize(B));
      }
      if (variant == 7) {  // pinned copy on B
        CK(cudaMallocAsync(&tw, pageable.size() * 4, B));
        CK(cudaMemcpyAsync(tw, h, pageable.size() * 4, cudaMemcpyHostToDevice, B));
      }
      if (variant == 8) {  // kernel reading pinned memory on B
        CK(cudaMallocAsync(&tw, pageable.size() * 4, B));
        k_copy<<<(pageable.size() + 255) / 256, 256, 0, B>>>(tw, (const float*)h, (int)pageable.size());
      }
      if (variant == 9) {  // pageable copy on a third stream, synchronised; B only waits for its event
        cudaStream_t C; CK(cudaStreamCreateWithFlags(&C, cudaStreamNonBlocking));
        CK(cudaMallocAsync(&tw, pageable.size() * 4, C));
        CK(cudaMemcpyAsync(tw, pageable.data(), pageable.size() * 4, cudaMemcpyHostToDevice, C));
        CK(cudaStreamSynchronize(C));
        CK(cudaStreamDestroy(C));
      }
      CK(cudaEventRecord(start, A));
      auto h0 = std::chrono::steady_clock::now();
      for (int i = 0; i < NB; i++) {
        if (variant == 0) CK(cudaMalloc(&d[i], bytes)); else CK(cudaMallocAsync(&d[i], bytes, A));
        CK(cudaMemcpyAsync(d[i], h + i * bytes, bytes, cudaMemcpyHostToDevice, A));
        CK(cudaEventRecord(ev[i], A));
        if (variant >= 3) { float* x; CK(cudaMallocAsync(&x, 1500000, B)); small = x; }  // interleaved allocations on B (IR spectra)
      }
      CK(cudaEventRecord(done, A));
      // "render": B
      double host_t0 = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
      CK(cudaEventRecord(t0, B));
      CK(cudaStreamWaitEvent(B, ev[8], 0));
      k_touch<<<100, 256, 0, B>>>(d[8], 25600);
      CK(cudaEventRecord(t1, B));
      CK(cudaStreamWaitEvent(B, ev[NB - 1], 0));
      k_touch<<<100, 256, 0, B>>>(d[NB - 1], 25600);
      CK(cudaEventRecord(t2, B));
      CK(cudaStreamSynchronize(B)); CK(cudaStreamSynchronize(A));
      float a, b, c, e; cudaEventElapsedTime(&a, start, t0); cudaEventElapsedTime(&b, start, t1); cudaEventElapsedTime(&c, start, t2); cudaEventElapsedTime(&e, start, done);
      if (rep == 4) printf("variant %d: t0 issued by the host at %+.3f ms, fired %+.3f ms, first kernel done %+.3f (wants buffer 8 of 64), last kernel done %+.3f, copies done %+.3f\n", variant, host_t0, a, b, c, e);
      for (int i = 0; i < NB; i++) { if (variant == 0) cudaFree(d[i]); else cudaFreeAsync(d[i], B); cudaEventDestroy(ev[i]); }
      if (tw) cudaFreeAsync(tw, B);
      cudaStreamSynchronize(B);
      cudaStreamDestroy(A); cudaStreamDestroy(B);
    }
  }
  return 0;
}
