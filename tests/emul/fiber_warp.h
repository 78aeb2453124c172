// TEST INFRASTRUCTURE.  Warp- and block-cooperative kernels on the host: every CUDA thread of a block is a user-level
// fiber (ucontext) on ONE OS thread, scheduled round-robin; a collective (__shfl_sync, __ballot_sync, __syncwarp,
// __syncthreads) is a counting barrier the fibers yield on.  Shared memory = `static` storage (one block runs at a
// time per OS thread), atomics = plain operations (fibers never run concurrently).  Deterministic, no data races by
// construction — it checks the ALGORITHM of a cooperative kernel (who exchanges what with whom), not its timing, memory
// model or divergence behaviour on real hardware.  Requires cuda_host_shim.h to be included first.
#pragma once
#include <ucontext.h>

#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

struct emul_fiber_block {
  static constexpr size_t STACK = 256 * 1024;
  unsigned nthreads = 0;
  std::vector<ucontext_t> ctx;
  std::vector<char *> stacks;
  std::vector<char> done;
  ucontext_t sched;
  unsigned cur = 0;
  std::function<void()> body;
  // block barrier
  unsigned bar_count = 0, bar_gen = 0, live = 0;
  // per-warp barrier + exchange slots
  std::vector<unsigned> wcount, wgen, wlive;
  std::vector<unsigned long long> xchg;  // [warp][32]
  unsigned long long progress = 0;
};
inline thread_local emul_fiber_block *emul_blk = nullptr;   // ONE instance per thread across translation units

static inline void emul_yield() {
  emul_fiber_block *b = emul_blk;
  swapcontext(&b->ctx[b->cur], &b->sched);
}
static inline void emul_block_barrier() {
  emul_fiber_block *b = emul_blk;
  unsigned g = b->bar_gen;
  if (++b->bar_count == b->live) {
    b->bar_count = 0;
    b->bar_gen++;
    b->progress++;
  } else {
    while (b->bar_gen == g) emul_yield();
  }
}
static inline void emul_warp_barrier() {
  emul_fiber_block *b = emul_blk;
  unsigned w = b->cur >> 5, g = b->wgen[w];
  if (++b->wcount[w] == b->wlive[w]) {
    b->wcount[w] = 0;
    b->wgen[w]++;
    b->progress++;
  } else {
    while (b->wgen[w] == g) emul_yield();
  }
}
static inline void __syncthreads() { emul_block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emul_warp_barrier(); }
template <class T>
static inline T emul_warp_exchange(T v, int src_lane) {
  emul_fiber_block *b = emul_blk;
  unsigned w = b->cur >> 5, lane = b->cur & 31;
  unsigned long long bits = 0;
  static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
  memcpy(&bits, &v, sizeof(T));
  b->xchg[32 * w + lane] = bits;
  emul_warp_barrier();                       // everybody has deposited
  unsigned long long got = b->xchg[32 * w + (src_lane & 31)];
  emul_warp_barrier();                       // everybody has read: the slots may be overwritten
  T r;
  memcpy(&r, &got, sizeof(T));
  return r;
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emul_warp_exchange(v, src); }
template <class T>
static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
  int lane = (int)(emul_blk->cur & 31);
  return emul_warp_exchange(v, lane + (int)d < 32 ? lane + (int)d : lane);
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
  int lane = (int)(emul_blk->cur & 31);
  return emul_warp_exchange(v, lane >= (int)d ? lane - (int)d : lane);
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emul_warp_exchange(v, (int)(emul_blk->cur & 31) ^ m); }
static inline unsigned __ballot_sync(unsigned, int pred) {
  emul_fiber_block *b = emul_blk;
  unsigned w = b->cur >> 5, lane = b->cur & 31;
  b->xchg[32 * w + lane] = pred ? 1 : 0;
  emul_warp_barrier();
  unsigned r = 0;
  for (unsigned l = 0; l < 32; l++)
    if (32 * w + l < b->nthreads && !b->done[32 * w + l] && b->xchg[32 * w + l]) r |= 1u << l;
  emul_warp_barrier();
  return r;
}
// lanes of the calling warp that have not exited
static inline unsigned __activemask() {
  emul_fiber_block *b = emul_blk;
  unsigned w = b->cur >> 5, m = 0;
  for (unsigned l = 0; l < 32; l++)
    if (32 * w + l < b->nthreads && !b->done[32 * w + l]) m |= 1u << l;
  return m;
}
// mask of the live lanes whose value equals the caller's (all live lanes of the warp must call)
template <class T>
static inline unsigned __match_any_sync(unsigned, T v) {
  emul_fiber_block *b = emul_blk;
  unsigned w = b->cur >> 5, lane = b->cur & 31;
  unsigned long long bits = 0;
  memcpy(&bits, &v, sizeof(T));
  b->xchg[32 * w + lane] = bits;
  emul_warp_barrier();
  unsigned r = 0;
  for (unsigned l = 0; l < 32; l++)
    if (32 * w + l < b->nthreads && !b->done[32 * w + l] && b->xchg[32 * w + l] == bits) r |= 1u << l;
  emul_warp_barrier();
  return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) {
  emul_fiber_block *b = emul_blk;
  unsigned w = b->cur >> 5, livemask = 0;
  for (unsigned l = 0; l < 32; l++)
    if (32 * w + l < b->nthreads && !b->done[32 * w + l]) livemask |= 1u << l;
  return __ballot_sync(m, pred) == livemask;
}
template <class T>
static inline T atomicAdd(T *p, T v) {
  T old = *p;
  *p = old + v;
  return old;
}

static void emul_fiber_main() {
  emul_fiber_block *b = emul_blk;
  b->body();
  unsigned t = b->cur;
  b->done[t] = 1;
  b->live--;
  b->wlive[t >> 5]--;
  b->progress++;
  // a thread that exits no longer takes part in barriers: release a barrier that was only waiting for it
  if (b->live && b->bar_count == b->live) {
    b->bar_count = 0;
    b->bar_gen++;
  }
  unsigned w = t >> 5;
  if (b->wlive[w] && b->wcount[w] == b->wlive[w]) {
    b->wcount[w] = 0;
    b->wgen[w]++;
  }
  swapcontext(&b->ctx[t], &b->sched);
}

// run ONE block of `block` threads cooperatively; `call()` invokes the kernel body with its arguments
static inline void emul_run_block_cooperative(unsigned block, emul_dim3 block_index, emul_dim3 grid, std::function<void()> call,
                                              std::vector<char *> *stack_pool = nullptr) {
  emul_fiber_block b;
  b.nthreads = block;
  b.live = block;
  b.ctx.resize(block);
  b.stacks.resize(block);
  b.done.assign(block, 0);
  unsigned nw = (block + 31) / 32;
  b.wcount.assign(nw, 0);
  b.wgen.assign(nw, 0);
  b.wlive.assign(nw, 0);
  for (unsigned t = 0; t < block; t++) b.wlive[t >> 5]++;
  b.xchg.assign(32 * nw, 0);
  b.body = call;
  emul_blk = &b;
  for (unsigned t = 0; t < block; t++) {
    b.stacks[t] = stack_pool ? (*stack_pool)[t] : (char *)malloc(emul_fiber_block::STACK);
    getcontext(&b.ctx[t]);
    b.ctx[t].uc_stack.ss_sp = b.stacks[t];
    b.ctx[t].uc_stack.ss_size = emul_fiber_block::STACK;
    b.ctx[t].uc_link = &b.sched;
    makecontext(&b.ctx[t], emul_fiber_main, 0);
  }
  blockDim = emul_dim3{block, 1, 1};
  gridDim = grid;
  blockIdx = block_index;
  while (b.live) {
    unsigned long long before = b.progress;
    for (unsigned t = 0; t < block; t++) {
      if (b.done[t]) continue;
      b.cur = t;
      threadIdx = emul_dim3{t, 0, 0};
      swapcontext(&b.sched, &b.ctx[t]);
    }
    if (b.live && b.progress == before) {
      fprintf(stderr, "emul_run_block_cooperative: deadlock (threads wait at different collectives)\n");
      abort();
    }
  }
  if (!stack_pool)
    for (unsigned t = 0; t < block; t++) free(b.stacks[t]);
  emul_blk = nullptr;
}
template <class K, class... A>
static inline void emul_cooperative_launch(K k, dim3 grid, unsigned block, A... a) {
  std::vector<char *> pool(block);
  for (unsigned t = 0; t < block; t++) pool[t] = (char *)malloc(emul_fiber_block::STACK);
  for (unsigned by = 0; by < grid.y; by++)
    for (unsigned bx = 0; bx < grid.x; bx++)
      emul_run_block_cooperative(block, emul_dim3{bx, by, 0}, emul_dim3{grid.x, grid.y, 1}, [=]() { k(a...); }, &pool);
  for (unsigned t = 0; t < block; t++) free(pool[t]);
}
