// simt_emu.h — minimal single-block SIMT emulator (TEST INFRASTRUCTURE ONLY).
//
// There is no GPU in the development container, so the CUDA kernels in
// flygym_b200/csrc/*.cuh are additionally compiled with g++ against this header
// and executed one thread block at a time: every CUDA thread is a ucontext
// coroutine; __syncthreads / __syncwarp / __shfl_*_sync are rendezvous points
// keyed by (warp, mask).  This catches indexing, masking and barrier-placement
// bugs on the CPU; it is never part of the product library (the product path
// fails loudly without the CUDA extension).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <map>
#include <vector>

namespace simt {

struct Thread {
  ucontext_t ctx;
  int tid = 0;
  bool done = false;
  std::vector<char> stack;
};

struct Barrier { int arrived = 0; unsigned long gen = 0; };

struct Block {
  std::vector<Thread> threads;
  ucontext_t sched;
  int cur = 0, nthreads = 0, block_id = 0, grid = 1;
  std::map<uint64_t, Barrier> barriers;
  float slot_f[1024];
  int or_acc = 0;
  std::function<void()> body;
};

inline Block*& blk() { static Block* b = nullptr; return b; }

inline void yield() {
  Block* b = blk();
  swapcontext(&b->threads[b->cur].ctx, &b->sched);
}

inline void rendezvous(uint64_t key, int count) {
  Block* b = blk();
  Barrier& br = b->barriers[key];
  unsigned long gen = br.gen;
  if (++br.arrived == count) { br.arrived = 0; br.gen++; return; }
  while (b->barriers[key].gen == gen) yield();
}

inline void trampoline() {
  Block* b = blk();
  b->body();
  b->threads[b->cur].done = true;
  swapcontext(&b->threads[b->cur].ctx, &b->sched);
}

// Run `body` as one block of `nthreads` CUDA threads.
inline void run_block(int nthreads, int block_id, int grid, std::function<void()> body) {
  Block b;
  blk() = &b;
  b.nthreads = nthreads; b.block_id = block_id; b.grid = grid; b.body = body;
  b.threads.resize(nthreads);
  for (int t = 0; t < nthreads; t++) {
    Thread& th = b.threads[t];
    th.tid = t;
    th.stack.resize(1 << 18);
    getcontext(&th.ctx);
    th.ctx.uc_stack.ss_sp = th.stack.data();
    th.ctx.uc_stack.ss_size = th.stack.size();
    th.ctx.uc_link = &b.sched;
    makecontext(&th.ctx, (void (*)())trampoline, 0);
  }
  int remaining = nthreads;
  long spins = 0;
  while (remaining > 0) {
    bool progressed = false;
    for (int t = 0; t < nthreads; t++) {
      if (b.threads[t].done) continue;
      b.cur = t;
      swapcontext(&b.sched, &b.threads[t].ctx);
      if (b.threads[t].done) { remaining--; progressed = true; }
    }
    if (!progressed && ++spins > 50000000L) { fprintf(stderr, "simt_emu: deadlock\n"); abort(); }
  }
  blk() = nullptr;
}

struct Idx { unsigned x, y, z; };
inline Idx thread_idx() { return Idx{(unsigned)blk()->cur, 0, 0}; }
inline Idx block_idx() { return Idx{(unsigned)blk()->block_id, 0, 0}; }
inline Idx block_dim() { return Idx{(unsigned)blk()->nthreads, 1, 1}; }
inline Idx grid_dim() { return Idx{(unsigned)blk()->grid, 1, 1}; }

inline int popc(unsigned m) { return __builtin_popcount(m); }

inline float shfl_from(unsigned mask, float v, int src_lane) {
  Block* b = blk();
  int tid = b->cur, warp = tid / 32, lane = tid % 32;
  if (!((mask >> lane) & 1u)) { fprintf(stderr, "simt_emu: lane %d not in its own shuffle mask %08x\n", lane, mask); abort(); }
  // a partial last warp only has the existing lanes
  int wthreads = b->nthreads - warp * 32; if (wthreads > 32) wthreads = 32;
  unsigned present = wthreads == 32 ? 0xffffffffu : ((1u << wthreads) - 1);
  int cnt = popc(mask & present);
  uint64_t key = ((uint64_t)(warp + 1) << 40) | ((uint64_t)mask << 4);
  b->slot_f[tid] = v;
  rendezvous(key | 1, cnt);
  float r = v;
  if (src_lane >= 0 && src_lane < 32 && ((mask >> src_lane) & 1u)) r = b->slot_f[warp * 32 + src_lane];
  rendezvous(key | 2, cnt);
  return r;
}

}  // namespace simt

// ------------------------------------------------------------------ CUDA spellings
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
#define threadIdx (simt::thread_idx())
#define blockIdx (simt::block_idx())
#define blockDim (simt::block_dim())
#define gridDim (simt::grid_dim())

inline void __syncthreads() { simt::rendezvous(1, simt::blk()->nthreads); }
// named barrier over `count` threads (PTX bar.sync id, count)
inline void simt_named_barrier(int id, int count) { simt::rendezvous(0x100 + (uint64_t)id, count); }
inline int __syncthreads_or(int pred) {
  simt::Block* b = simt::blk();
  if (pred) b->or_acc = 1;
  simt::rendezvous(5, b->nthreads);
  const int r = b->or_acc;
  simt::rendezvous(6, b->nthreads);
  b->or_acc = 0;                      // every thread clears it; nobody sets it again before passing rendezvous 5 of the next call
  simt::rendezvous(7, b->nthreads);
  return r;
}
inline void __syncwarp(unsigned mask = 0xffffffffu) {
  simt::Block* b = simt::blk();
  int warp = b->cur / 32;
  int wthreads = b->nthreads - warp * 32; if (wthreads > 32) wthreads = 32;
  unsigned present = wthreads == 32 ? 0xffffffffu : ((1u << wthreads) - 1);
  simt::rendezvous(((uint64_t)(warp + 1) << 40) | ((uint64_t)mask << 4) | 3, simt::popc(mask & present));
}
inline float __shfl_sync(unsigned mask, float v, int src, int width = 32) {
  int lane = simt::blk()->cur % 32;
  int base = lane / width * width;
  return simt::shfl_from(mask, v, base + (src % width));
}
inline float __shfl_up_sync(unsigned mask, float v, unsigned delta, int width = 32) {
  int lane = simt::blk()->cur % 32;
  int src = (lane % width) >= (int)delta ? lane - (int)delta : lane;
  return simt::shfl_from(mask, v, src);
}
inline float __shfl_down_sync(unsigned mask, float v, unsigned delta, int width = 32) {
  int lane = simt::blk()->cur % 32;
  int src = (lane % width) + (int)delta < width ? lane + (int)delta : lane;
  return simt::shfl_from(mask, v, src);
}
inline float __shfl_xor_sync(unsigned mask, float v, int lanemask, int width = 32) {
  int lane = simt::blk()->cur % 32;
  int src = lane ^ lanemask;
  if (src / width != lane / width) src = lane;
  return simt::shfl_from(mask, v, src);
}
// 64-bit shuffles move the two halves one after the other, as the hardware does
#define SIMT_SHFL_F64(NAME, ARGT)                                                         \
  inline double NAME(unsigned mask, double v, ARGT a, int width = 32) {                   \
    float h[2]; memcpy(h, &v, 8);                                                         \
    h[0] = NAME(mask, h[0], a, width); h[1] = NAME(mask, h[1], a, width);                 \
    memcpy(&v, h, 8); return v;                                                           \
  }
SIMT_SHFL_F64(__shfl_sync, int)
SIMT_SHFL_F64(__shfl_up_sync, unsigned)
SIMT_SHFL_F64(__shfl_down_sync, unsigned)
SIMT_SHFL_F64(__shfl_xor_sync, int)
inline int __shfl_sync(unsigned mask, int v, int src, int width = 32) {
  float f; memcpy(&f, &v, 4); f = __shfl_sync(mask, f, src, width); memcpy(&v, &f, 4); return v;
}
inline int __shfl_xor_sync(unsigned mask, int v, int lanemask, int width = 32) {
  float f; memcpy(&f, &v, 4); f = __shfl_xor_sync(mask, f, lanemask, width); memcpy(&v, &f, 4); return v;
}
inline int __any_sync(unsigned mask, int pred) {
  // emulate with an OR-reduction over the mask via repeated shuffles
  int lane = simt::blk()->cur % 32; int acc = pred != 0;
  for (int l = 0; l < 32; l++) { if (!((mask >> l) & 1u)) continue; int v = __shfl_sync(mask, pred != 0, l); acc |= v; }
  (void)lane; return acc;
}
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcs(const T* p) { return *p; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline void sincos(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }
inline float __fdividef(float a, float b) { return a / b; }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline void __threadfence_block() {}
