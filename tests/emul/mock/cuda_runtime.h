// TEST INFRASTRUCTURE.  A stand-in for <cuda_runtime.h> so the HOST side of the C ABI units that launch only
// non-cooperative kernels (capi_fr.cu, capi_h2c.cu) can be compiled with g++ and run in the CPU suite: "device memory"
// is host memory, a stream is a no-op (every call completes before it returns), a kernel launch is the (block, thread)
// loop of cuda_host_shim.h.  It exercises the staging / offset / caching / error-code logic of those entry points;
// it says nothing about real asynchrony, alignment faults or launch limits.
#pragma once
#include <cstdlib>
#include <cstring>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };

static inline cudaError_t cudaMalloc(void **p, size_t n) {
  n = (n + 255) & ~(size_t)255;
  *p = std::aligned_alloc(256, n ? n : 256);
  if (*p) std::memset(*p, 0xA5, n ? n : 256);  // poison: reads of never-written scratch show up as wrong results
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void *p) {
  std::free(p);
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) {
  std::memmove(d, s, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) {
  std::memset(d, v, n);
  return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) {
  *d = 0;
  return cudaSuccess;
}
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
struct emul_event_rec {
  double t_ms;
};
double emul_event_now_ms();
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) {
  *e = std::calloc(1, sizeof(emul_event_rec));
  return cudaSuccess;
}
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
  std::free(e);
  return cudaSuccess;
}
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
  if (e) static_cast<emul_event_rec *>(e)->t_ms = emul_event_now_ms();
  return cudaSuccess;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = (float)(static_cast<emul_event_rec *>(b)->t_ms - static_cast<emul_event_rec *>(a)->t_ms);
  return cudaSuccess;
}
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "mock CUDA error"; }
// ---- what capi_basic.cu needs on top: one "device", streams / events as opaque tokens, a wall clock for event times
#include <chrono>
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
struct cudaDeviceProp {
  int multiProcessorCount;
};
static inline cudaError_t cudaGetDeviceCount(int *n) {
  *n = 1;
  return cudaSuccess;
}
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
  p->multiProcessorCount = 1;   // keeps the "batch larger than one wave" thresholds small on the CPU
  return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) {
  *s = std::malloc(1);
  return cudaSuccess;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) {
  std::free(s);
  return cudaSuccess;
}
inline double emul_event_now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
