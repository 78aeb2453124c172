// TEST INFRASTRUCTURE.  Lets g++ compile the device-side headers of bls12_381_b200/csrc/ (fp.cuh ... pairing.cuh)
// as plain C++ so the CPU suite (-m "not gpu") can run the SAME source the GPU runs — limb algorithms, tower,
// curve formulas, Miller loop, final exponentiation — against the oracle.  What it does NOT cover: the PTX -> SASS
// path, launch geometry, warp/CTA cooperation and memory staging; those are the job of the -m gpu tests.
//
//   * CUDA qualifiers become no-ops, threadIdx/blockIdx/... are thread-local variables a host loop steps through
//     (only kernels without intra-block cooperation can be run that way);
//   * the PTX carry-chain primitives of fp.cuh are replaced by bit-exact C models (emul_ptx.h) with an explicit
//     carry flag: add.cc/addc/sub.cc/subc/mad{c}.{lo,hi}{.cc} as defined in the PTX ISA ("Extended-Precision
//     Integer Arithmetic": CC.CF is the carry-out of add/mad and the borrow-out of sub).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <algorithm>   // all standard headers BEFORE the qualifier macros (libstdc++ uses __noinline__ itself)
#include <vector>
#include <thread>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>

#define B200_HOST_EMUL 1
#define __device__
#define __host__
#define __global__
#define __constant__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static   /* static shared arrays: one block runs at a time; dynamic ones: B200_DYN_SMEM (ctx.cuh) */
#define __restrict__

struct uint4 {
  uint32_t x, y, z, w;
};
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct emul_dim3 {
  unsigned x = 1, y = 1, z = 1;
};
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
// `inline`: ONE instance per thread across all translation units of a library (inline device functions with external
// linkage are merged by the linker and must see the same state as the kernel that calls them)
inline thread_local emul_dim3 threadIdx, blockIdx, blockDim, gridDim;

template <class T>
static inline T __ldg(const T *p) { return *p; }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) {
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return (uint32_t)((v << (s & 31)) >> 32);
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return (uint32_t)(v >> (s & 31));
}
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
// PRMT: result byte i = byte (selector nibble i & 7) of the 8 bytes {y, x}; bit 3 of a nibble (sign replication) unused here
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t sel) {
  uint64_t v = ((uint64_t)y << 32) | x;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) {
    uint32_t n = (sel >> (4 * i)) & 0xf;
    uint32_t b = (uint32_t)(v >> (8 * (n & 7))) & 0xff;
    if (n & 8) b = (b & 0x80) ? 0xff : 0x00;
    r |= b << (8 * i);
  }
  return r;
}
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
using std::min;
using std::max;

// EMUL_TRACE=1 in the environment prints every kernel launch of the mock-runtime build to stderr and installs a SIGSEGV
// handler that prints a backtrace (addresses resolvable with addr2line against the -g build of the library)
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static inline void emul_segv_handler(int) {
  void *bt[64];
  int n = backtrace(bt, 64);
  backtrace_symbols_fd(bt, n, 2);
  _exit(139);
}
static inline void emul_trace_launch(const char *name) {
  static const bool on = getenv("EMUL_TRACE") != nullptr;
  static bool installed = false;
  if (on && !installed) {
    installed = true;
    static char altstack[1 << 16];
    stack_t ss;
    ss.ss_sp = altstack;
    ss.ss_size = sizeof(altstack);
    ss.ss_flags = 0;
    sigaltstack(&ss, nullptr);
    struct sigaction sa;
    memset(&sa, 0, sizeof(sa));
    sa.sa_handler = emul_segv_handler;
    sa.sa_flags = SA_ONSTACK;
    sigaction(SIGSEGV, &sa, nullptr);
  }
  if (on) fprintf(stderr, "[emul] launch %s\n", name);
}
#ifndef EMUL_DYNAMIC_SMEM_BYTES
#define EMUL_DYNAMIC_SMEM_BYTES (256 * 1024)
#endif
alignas(16) inline char emul_dyn_smem[EMUL_DYNAMIC_SMEM_BYTES];   // what B200_DYN_SMEM points at
#ifdef EMUL_LAUNCH_COOPERATIVE
// every launch goes through the fiber scheduler of fiber_warp.h (kernels with __syncthreads / shuffles / shared memory)
#include "fiber_warp.h"
template <class K, class... A>
static inline void emul_kernel_launch(K k, dim3 grid, dim3 block, A... a) {
  emul_cooperative_launch(k, grid, block.x, a...);
}
#else
// a kernel "launch": every (block, thread) in turn on the calling host thread.  Only valid for kernels without
// intra-block cooperation (no __syncthreads / shared memory / shuffles).
template <class K, class... A>
static inline void emul_kernel_launch(K k, unsigned grid, unsigned block, A... a) {
  blockDim = emul_dim3{block, 1, 1};
  gridDim = emul_dim3{grid, 1, 1};
  for (unsigned b = 0; b < grid; b++) {
    blockIdx = emul_dim3{b, 0, 0};
    for (unsigned t = 0; t < block; t++) {
      threadIdx = emul_dim3{t, 0, 0};
      k(a...);
    }
  }
}
#endif
