// nmf_step_common.cuh — one NeuroMechFly physics step per thread block (sm_100a): overview + precision-independent helpers.
//
// Replaces, for the reference benchmark model, the whole of
//   GPUSimulation.step -> mujoco_warp.step   (reference src/flygym/warp/simulation.py:260-263)
//   Simulation.step    -> mujoco.mj_step     (reference src/flygym/simulation.py:74-76)
// with ONE fused kernel: forward kinematics, composite inertias, bias forces,
// position + adhesion actuators, geom-plane collision, soft-contact Newton solve
// and semi-implicit Euler, state staying in shared memory / registers for
// `nsteps` consecutive steps.
//
// Mapping (B200-first, not a port): a block of 64 threads owns one fly.
//   tid  0..47 : leg-body lanes, (leg = tid/8, link = tid%8); 8-lane shuffle
//                segments = one kinematic chain, so every chain recursion of the
//                classical algorithms becomes a 3-step warp-shuffle scan:
//                  FK           = inclusive scan of rigid transforms
//                  velocities   = prefix sums of spatial vectors (common c-frame)
//                  CRBA / RNE   = suffix sums of spatial inertias / wrenches
//   tid 48..63 : hub lanes (free body, its 6 DoFs, its contact geoms)
// Newton Hessian: M + J'DJ is assembled as a CRBA over *contact-augmented*
// spatial inertias (each contact adds X'WX to its body), so it keeps M's
// arrowhead sparsity (hub 6x6 + six 11x11 chains); each chain block is factorised
// L'DL in registers, one matrix column per lane, the hub block by Schur complement.
//
// The same source is compiled by g++ against tests/simt_emu/simt_emu.h
// (NMF_SIMT_EMU) so it can be exercised without a GPU; that is test
// infrastructure, not a fallback: the shipped library only contains the nvcc build.
#pragma once
#include <type_traits>

#include "nmf_layout.h"

// Precision-independent part of the step kernel source.  nmf_step.cuh (the kernel body) is written in terms of `real` and is
// included twice by nmf_step_all.cuh: as namespace nmf::f32 (real = float, the product path) and nmf::f64 (real = double, a
// validation instantiation of the SAME source that shadows the fp64 oracle over long horizons).
namespace nmf {

#define NMF_FULL 0xffffffffu

// Block barrier that first reconverges each warp: __syncthreads() is the *aligned* barrier and is undefined when a warp
// reaches it diverged (ptxas may leave lanes diverged after predicated stores; compute-sanitizer synccheck flags it).
// `bar` selects the barrier: 0 = the whole block (one fly per block); b > 0 = named barrier b over the CTA (= 64) threads of ONE
// fly, used when a block steps several flies side by side (step_entry<WORLD, FPB > 1>).
#ifndef NMF_SIMT_EMU
__device__ __forceinline__ void block_sync(int bar) {
  __syncwarp(NMF_FULL);
  if (bar == 0) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(CTA) : "memory");
}
#else
inline void block_sync(int bar) { __syncwarp(NMF_FULL); if (bar == 0) __syncthreads(); else simt_named_barrier(bar, CTA); }
#endif

// work-queue primitives (device: gpu-scope acquire / release; emulator: one block at a time, plain accesses)
#ifndef NMF_SIMT_EMU
#define NMF_ATOMIC_ADD(ptr, v) atomicAdd((ptr), (v))
__device__ __forceinline__ int nmf_ld_acquire(const int* ptr) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory"); return v; }
__device__ __forceinline__ void nmf_st_release(int* ptr, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory"); }
#define NMF_LD_ACQUIRE(ptr) nmf_ld_acquire(ptr)
#define NMF_ST_RELEASE(ptr, v) nmf_st_release((ptr), (v))
#define NMF_FENCE_PROXY_ASYNC() asm volatile("fence.proxy.async;" ::: "memory")
#define NMF_THREADFENCE() __threadfence()
#else
inline int nmf_emu_atomic_add(int* ptr, int v) { const int old = *ptr; *ptr = old + v; return old; }
#define NMF_ATOMIC_ADD(ptr, v) nmf_emu_atomic_add((ptr), (v))
#define NMF_LD_ACQUIRE(ptr) (*(ptr))
#define NMF_ST_RELEASE(ptr, v) (*(ptr) = (v))
#define NMF_FENCE_PROXY_ASYNC() ((void)0)
#define NMF_THREADFENCE() ((void)0)
#endif

// Several flies per block (step_slot<WORLD, FPB>): which fly slot a thread serves and its thread index inside that fly.
// NMF_FPB_INTERLEAVE = 1 gives fly s the warps s and s + FPB of the block (warp w runs on SM sub-partition w % 4, so with
// FPB = 4 the two warps of a fly share one sub-partition's L0 instruction cache); 0 = consecutive warps.
#ifndef NMF_FPB_INTERLEAVE
#define NMF_FPB_INTERLEAVE 0
#endif
// value of lane 0, which tells the compiler that v is the same in every lane of the warp: warp-uniform values (fly slot,
// work item, shared-memory base of the slot) can then live in uniform registers instead of the 64 scarce vector registers
__device__ __forceinline__ int warp_uniform(int v) { return __shfl_sync(NMF_FULL, v, 0); }
template <int FPB> __device__ __forceinline__ int fly_slot() {
  if (FPB == 1) return 0;
  return warp_uniform(NMF_FPB_INTERLEAVE ? (int)(threadIdx.x >> 5) % FPB : (int)(threadIdx.x / CTA));
}
template <int FPB> __device__ __forceinline__ int fly_tid() {
  if (FPB == 1) return (int)threadIdx.x;
  return NMF_FPB_INTERLEAVE ? (int)((threadIdx.x >> 5) / FPB) * 32 + (int)(threadIdx.x & 31) : (int)(threadIdx.x & (CTA - 1));
}

// ------------------------------------------------------------------ TMA (bulk async copy) of the float32 state record
// One elected thread moves the whole 1216-byte record HBM <-> shared memory with cp.async.bulk (SASS: UBLKCP); the block
// waits on an mbarrier.  Under the SIMT emulator the same copies are plain loops.
#ifndef NMF_SIMT_EMU
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_load_f32(float* dst_smem, const float* src_gmem, unsigned long long* mbar, int tid, int bar) {
  const unsigned mb = smem_u32(mbar), dst = smem_u32(dst_smem);
  constexpr unsigned bytes = S_STRIDE * sizeof(float);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  block_sync(bar);
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src_gmem), "r"(bytes), "r"(mb) : "memory");
  }
  unsigned done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mb) : "memory");
  }
  block_sync(bar);
  if (tid == 0) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(mb) : "memory");   // the slot is re-initialised by the next work item
}
// `published`: the caller hands the record to another block afterwards (work-queue scheduling), so wait until the
// global writes have completed, not only until shared memory has been read.
__device__ __forceinline__ void tma_store_f32(float* dst_gmem, const float* src_smem, int tid, bool published, int bar) {
  block_sync(bar);                                                      // all generic-proxy writes to the record are done
  if (tid == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // make them visible to the async (TMA) proxy
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"((unsigned)(S_STRIDE * sizeof(float))) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (published) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must stay valid until the copy has read it
  }
}
#else
__device__ __forceinline__ void tma_load_f32(float* dst_smem, const float* src_gmem, unsigned long long*, int tid, int bar) {
  for (int i = tid; i < S_STRIDE; i += CTA) dst_smem[i] = src_gmem[i];
  block_sync(bar);
}
__device__ __forceinline__ void tma_store_f32(float* dst_gmem, const float* src_smem, int tid, bool, int bar) {
  block_sync(bar);
  for (int i = tid; i < S_STRIDE; i += CTA) dst_gmem[i] = src_smem[i];
}
#endif

// The same movers for a record of `nfloats` floats (a multiple of 4) and a block of any size synchronised with __syncthreads
// (the general-topology kernels of nmf_tree.cuh: the record length depends on the model).
#ifndef NMF_SIMT_EMU
__device__ __forceinline__ void tma_load_n(float* dst_smem, const float* src_gmem, int nfloats, unsigned long long* mbar, int tid) {
  const unsigned mb = smem_u32(mbar), dst = smem_u32(dst_smem), bytes = (unsigned)nfloats * 4u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  block_sync(0);
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src_gmem), "r"(bytes), "r"(mb) : "memory");
  }
  unsigned done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mb) : "memory");
  }
  block_sync(0);
  if (tid == 0) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(mb) : "memory");
}
__device__ __forceinline__ void tma_store_n(float* dst_gmem, const float* src_smem, int nfloats, int tid) {
  block_sync(0);
  if (tid == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"((unsigned)nfloats * 4u) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}
#else
__device__ __forceinline__ void tma_load_n(float* dst_smem, const float* src_gmem, int nfloats, unsigned long long*, int tid) {
  for (int i = tid; i < nfloats; i += (int)blockDim.x) dst_smem[i] = src_gmem[i];
  block_sync(0);
}
__device__ __forceinline__ void tma_store_n(float* dst_gmem, const float* src_smem, int nfloats, int tid) {
  block_sync(0);
  for (int i = tid; i < nfloats; i += (int)blockDim.x) dst_gmem[i] = src_smem[i];
}
#endif

// ------------------------------------------------------------------ float / double spellings of the math used by the body
__device__ __forceinline__ float m_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double m_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float m_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double m_min(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float m_abs(float a) { return fabsf(a); }
__device__ __forceinline__ double m_abs(double a) { return fabs(a); }
__device__ __forceinline__ float m_rsqrt(float a) { return rsqrtf(a); }
__device__ __forceinline__ double m_rsqrt(double a) { return 1.0 / sqrt(a); }
__device__ __forceinline__ float m_sqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ double m_sqrt(double a) { return sqrt(a); }
__device__ __forceinline__ float m_rint(float a) { return rintf(a); }
__device__ __forceinline__ double m_rint(double a) { return rint(a); }
__device__ __forceinline__ float m_pow(float a, float b) { return powf(a, b); }
__device__ __forceinline__ double m_pow(double a, double b) { return pow(a, b); }
__device__ __forceinline__ float m_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double m_fma(double a, double b, double c) { return fma(a, b, c); }
// integer fields of the role table: bit patterns in the float table, plain numbers in the double one
__device__ __forceinline__ int role_int(float v) { return __float_as_int(v); }
__device__ __forceinline__ int role_int(double v) { return (int)v; }
// line-search stopping thresholds (relative derivative, relative bracket): at the resolution of the arithmetic
template <class real> struct Prec;
template <> struct Prec<float> { static constexpr float ls_rel = 2e-6f, ls_bracket = 1e-6f, ls_amin = 1e-3f; };
template <> struct Prec<double> { static constexpr double ls_rel = 1e-14, ls_bracket = 1e-15, ls_amin = 1e-3; };

// sin/cos with a 2-term Cody-Waite reduction and cephes-style minimax polynomials (|err| ~ 1 ulp for |x| < ~1e3):
// replaces sincosf, whose inlined slow path bloated the instruction footprint of an I-cache-bound kernel.
__device__ __forceinline__ void sincos_small(float x, float* sn, float* cs) {
  const float kf = rintf(x * 0.63661977236758134f);
  float r = fmaf(-kf, 1.5707962512969971f, x);
  r = fmaf(-kf, 7.5497894158615964e-8f, r);
  const int q = (int)kf;
  const float r2 = r * r;
  const float ps = fmaf(fmaf(fmaf(-1.9515295891e-4f, r2, 8.3321608736e-3f), r2, -1.6666654611e-1f) * r2, r, r);
  const float pc = fmaf(fmaf(fmaf(2.443315711809948e-5f, r2, -1.388731625493765e-3f), r2, 4.166664568298827e-2f) * r2, r2, fmaf(-0.5f, r2, 1.0f));
  const float s0 = (q & 1) ? pc : ps, c0 = (q & 1) ? ps : pc;
  *sn = (q & 2) ? -s0 : s0;
  *cs = ((q + 1) & 2) ? -c0 : c0;
}
__device__ __forceinline__ void sincos_small(double x, double* sn, double* cs) { sincos(x, sn, cs); }

// packed upper-triangular index of a symmetric 6x6, a <= b
__device__ __forceinline__ constexpr int s6(int a, int b) { return a * 6 - a * (a - 1) / 2 + (b - a); }
#ifdef NMF_SIMT_EMU
#define NMF_COLD
#else
#define NMF_COLD __noinline__
#endif

}  // namespace nmf
