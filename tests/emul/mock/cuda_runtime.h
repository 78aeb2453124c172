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
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) {
  *d = 0;
  return cudaSuccess;
}
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) {
  *e = nullptr;
  return cudaSuccess;
}
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "mock CUDA error"; }
