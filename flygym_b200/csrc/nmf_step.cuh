// nmf_step.cuh — one NeuroMechFly physics step per thread block (sm_100a).
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

namespace nmf {

#define NMF_FULL 0xffffffffu
#define NMF_MINVAL 1e-15f

// Block barrier that first reconverges each warp: __syncthreads() is the *aligned* barrier and is undefined when a warp
// reaches it diverged (ptxas may leave lanes diverged after predicated stores; compute-sanitizer synccheck flags it).
__device__ __forceinline__ void block_sync() { __syncwarp(NMF_FULL); __syncthreads(); }

// ------------------------------------------------------------------ small math
__device__ __forceinline__ void qmul(const float* a, const float* b, float* r) {
  float w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  float x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  float y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  float z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
__device__ __forceinline__ void qrot(const float* q, const float* v, float* r) {
  // r = v + 2 w (u x v) + 2 u x (u x v)
  float tx = 2.f * (q[2] * v[2] - q[3] * v[1]), ty = 2.f * (q[3] * v[0] - q[1] * v[2]), tz = 2.f * (q[1] * v[1] - q[2] * v[0]);
  float rx = v[0] + q[0] * tx + (q[2] * tz - q[3] * ty);
  float ry = v[1] + q[0] * ty + (q[3] * tx - q[1] * tz);
  float rz = v[2] + q[0] * tz + (q[1] * ty - q[2] * tx);
  r[0] = rx; r[1] = ry; r[2] = rz;
}
__device__ __forceinline__ void q2mat(const float* q, float* m) {
  float w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = 1.f - 2.f * (y * y + z * z); m[1] = 2.f * (x * y - w * z); m[2] = 2.f * (x * z + w * y);
  m[3] = 2.f * (x * y + w * z); m[4] = 1.f - 2.f * (x * x + z * z); m[5] = 2.f * (y * z - w * x);
  m[6] = 2.f * (x * z - w * y); m[7] = 2.f * (y * z + w * x); m[8] = 1.f - 2.f * (x * x + y * y);
}
__device__ __forceinline__ void qnormalize(float* q) {
  float n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (n2 < 1e-30f) { q[0] = 1.f; q[1] = q[2] = q[3] = 0.f; return; }
  float s = rsqrtf(n2);
  q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
}
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
__device__ __forceinline__ void cross3(const float* a, const float* b, float* r) {
  float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
__device__ __forceinline__ float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ float dot6(const float* a, const float* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
// spatial inertia (Ixx Iyy Izz Ixy Ixz Iyz | hx hy hz | m) times motion vector (ang, lin)
__device__ __forceinline__ void mul_inert(const float* i, const float* v, float* r) {
  float r0 = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  float r1 = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  float r2 = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  float r3 = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  float r4 = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  float r5 = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
  r[0] = r0; r[1] = r1; r[2] = r2; r[3] = r3; r[4] = r4; r[5] = r5;
}
__device__ __forceinline__ void cross_motion(const float* vel, const float* v, float* r) {
  float a[3], b[3], c[3];
  cross3(vel, v, a); cross3(vel, v + 3, b); cross3(vel + 3, v, c);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
__device__ __forceinline__ void cross_force(const float* vel, const float* f, float* r) {
  float a[3], b[3], c[3];
  cross3(vel, f, a); cross3(vel + 3, f + 3, b); cross3(vel, f + 3, c);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
// packed upper-triangular index of a symmetric 6x6, a <= b
__device__ __forceinline__ constexpr int s6(int a, int b) { return a * 6 - a * (a - 1) / 2 + (b - a); }
// expand a 10-float spatial inertia into packed symmetric 6x6 (order: wx wy wz vx vy vz)
__device__ __forceinline__ void expand_inert(const float* i, float* P) {
  P[s6(0, 0)] = i[0]; P[s6(0, 1)] = i[3]; P[s6(0, 2)] = i[4]; P[s6(0, 3)] = 0.f;   P[s6(0, 4)] = -i[8]; P[s6(0, 5)] = i[7];
  P[s6(1, 1)] = i[1]; P[s6(1, 2)] = i[5]; P[s6(1, 3)] = i[8];  P[s6(1, 4)] = 0.f;  P[s6(1, 5)] = -i[6];
  P[s6(2, 2)] = i[2]; P[s6(2, 3)] = -i[7]; P[s6(2, 4)] = i[6]; P[s6(2, 5)] = 0.f;
  P[s6(3, 3)] = i[9]; P[s6(3, 4)] = 0.f; P[s6(3, 5)] = 0.f; P[s6(4, 4)] = i[9]; P[s6(4, 5)] = 0.f; P[s6(5, 5)] = i[9];
}
__device__ __forceinline__ void sym6_mul(const float* P, const float* v, float* r) {
#pragma unroll
  for (int a = 0; a < 6; a++) {
    float s = 0.f;
#pragma unroll
    for (int b = 0; b < 6; b++) s += P[a <= b ? s6(a, b) : s6(b, a)] * v[b];
    r[a] = s;
  }
}

// ------------------------------------------------------------------ 8-lane chain scans
template <int N>
__device__ __forceinline__ void chain_prefix(float* v, unsigned mask, int k) {  // inclusive, root -> tip
#pragma unroll
  for (int off = 1; off < 8; off <<= 1) {
#pragma unroll
    for (int n = 0; n < N; n++) { float t = __shfl_up_sync(mask, v[n], off, 8); if (k >= off) v[n] += t; }
  }
}
template <int N>
__device__ __forceinline__ void chain_suffix(float* v, unsigned mask, int k) {  // inclusive, tip -> root
#pragma unroll
  for (int off = 1; off < 8; off <<= 1) {
#pragma unroll
    for (int n = 0; n < N; n++) { float t = __shfl_down_sync(mask, v[n], off, 8); if (k + off < 8) v[n] += t; }
  }
}

// block-wide sum of N values; call from converged code only (all 64 threads)
template <int N>
__device__ __forceinline__ void cta_reduce(float* v, float* s_red, int& parity, int tid) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int n = 0; n < N; n++) v[n] += __shfl_xor_sync(NMF_FULL, v[n], off);
  }
  float* buf = s_red + parity * 16;
  if ((tid & 31) == 0) {
#pragma unroll
    for (int n = 0; n < N; n++) buf[(tid >> 5) * 8 + n] = v[n];
  }
  block_sync();
#pragma unroll
  for (int n = 0; n < N; n++) v[n] = buf[n] + buf[8 + n];
  parity ^= 1;
}

// ------------------------------------------------------------------ contact slot (kept in registers)
// Branch-free convention: an inactive slot has D = c0 = w = s = 0, so it contributes nothing anywhere.
struct Contact {     // 10 registers per slot; "active" <=> D > 0; signed distance = 2 (r[2] + com_z) (the point sits midway)
  float r[3];        // contact position relative to the subtree COM
  float cx, cy;      // first tangent (cx, cy, 0); second = (-cy, cx, 0); normal = +z
  float D;           // 1/R of the four pyramid rows (0 when the slot is inactive)
  float c0;          // K * imp * (dist - margin)
  float w[3];        // (n, mu t1, mu t2) . (a_p + B v_p)  for the current qacc
};
__device__ __forceinline__ float con_on(const Contact& c) { return c.D > 0.f ? 1.f : 0.f; }

// general-exponent branch of the impedance sigmoid: kept out of line (four inlined powf bodies per call site are ~18 KB of
// SASS that the reference's power-2 / power-1 settings never execute, in an instruction-fetch-bound kernel)
#ifdef NMF_SIMT_EMU
#define NMF_COLD
#else
#define NMF_COLD __noinline__
#endif
__device__ NMF_COLD float impedance_general(float x, float mid, float power) {
  return (x <= mid) ? powf(x, power) / powf(mid, power - 1.f) : 1.f - powf(1.f - x, power) / powf(1.f - mid, power - 1.f);
}
__device__ __forceinline__ float impedance_of(const float* solimp, float x_abs) {
  const float d0 = solimp[0], d1 = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  if (d0 == d1 || width <= NMF_MINVAL) return 0.5f * (d0 + d1);   // (uniform branch)
  float x = fminf(x_abs / width, 1.f);
  float y;
  if (power == 1.f) y = x;
  else if (power == 2.f) y = (x <= mid) ? x * x / mid : 1.f - (1.f - x) * (1.f - x) / (1.f - mid);
  else y = impedance_general(x, mid, power);
  return d0 + y * (d1 - d0);
}

__device__ __forceinline__ float impedance(const StepParams& p, float x_abs) { return impedance_of(p.solimp, x_abs); }

// finishes a candidate contact: solver parameters and the B*velocity part of the rows
__device__ __forceinline__ void finish_contact(const StepParams& p, Contact& c, float active, float dist, const float* pos, float hx, float hy,
                                               const float* com, const float* cvel, float invw) {
  c.r[0] = pos[0] - com[0]; c.r[1] = pos[1] - com[1]; c.r[2] = pos[2] - com[2];
  c.cx = hx; c.cy = hy;
  float imp = impedance(p, fabsf(dist - p.margin));
  float R0 = fmaxf(NMF_MINVAL, (1.f - imp) * invw * (1.f + p.mu * p.mu) / imp);
  c.D = active / (2.f * (p.mu * p.mu / p.impratio) * R0);
  c.c0 = active * p.cK * imp * (dist - p.margin);
  float vp[3] = {cvel[3] + cvel[1] * c.r[2] - cvel[2] * c.r[1], cvel[4] + cvel[2] * c.r[0] - cvel[0] * c.r[2],
                 cvel[5] + cvel[0] * c.r[1] - cvel[1] * c.r[0]};
  c.w[0] = active * p.cB * vp[2];
  c.w[1] = active * p.cB * p.mu * (c.cx * vp[0] + c.cy * vp[1]);
  c.w[2] = active * p.cB * p.mu * (-c.cy * vp[0] + c.cx * vp[1]);
}

// geom-vs-ground-plane narrow phase for the geom carried by this lane's body
// (plane z = 0, normal +z: reference world.py:251-260).  Fills two slots.
__device__ __forceinline__ float collide(const StepParams& p, const float* role, int tid, const float* xpos, const float* R,
                                         const float* com, const float* cvel, float invw, Contact* con, int& hullv) {
  const int gtype = __float_as_int(role[RF_GTYPE * CTA + tid]);
  float pos0[3] = {0.f, 0.f, 0.f}, pos1[3] = {0.f, 0.f, 0.f}, d0 = 1.f, d1 = 1.f, a0 = 0.f, a1 = 0.f, hx = 0.f, hy = 1.f;
  {  // capsule: two sphere-plane tests, frame aligned with the capsule axis (evaluated on every lane, masked by type)
    float gp[3] = {role[(RF_GPOS + 0) * CTA + tid], role[(RF_GPOS + 1) * CTA + tid], role[(RF_GPOS + 2) * CTA + tid]};
    float ga[3] = {role[(RF_GAXIS + 0) * CTA + tid], role[(RF_GAXIS + 1) * CTA + tid], role[(RF_GAXIS + 2) * CTA + tid]};
    float rad = role[RF_GRAD * CTA + tid], half = role[RF_GHALF * CTA + tid];
    float c[3], a[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      c[i] = xpos[i] + R[3 * i] * gp[0] + R[3 * i + 1] * gp[1] + R[3 * i + 2] * gp[2];
      a[i] = R[3 * i] * ga[0] + R[3 * i + 1] * ga[1] + R[3 * i + 2] * ga[2];
    }
    float hn2 = a[0] * a[0] + a[1] * a[1];
    float inv = rsqrtf(fmaxf(hn2, 1e-24f));
    const bool cap = gtype == 0;
    if (cap) { hx = hn2 < 1e-24f ? 1.f : a[0] * inv; hy = hn2 < 1e-24f ? 0.f : a[1] * inv; }
    float e0 = c[2] + half * a[2], e1 = c[2] - half * a[2];
    if (cap) {
      d0 = e0 - rad; d1 = e1 - rad;
      pos0[0] = c[0] + half * a[0]; pos0[1] = c[1] + half * a[1]; pos0[2] = 0.5f * d0;
      pos1[0] = c[0] - half * a[0]; pos1[1] = c[1] - half * a[1]; pos1[2] = 0.5f * d1;
      a0 = (e0 <= p.margin + rad) ? 1.f : 0.f; a1 = (e1 <= p.margin + rad) ? 1.f : 0.f;
    }
  }
  if (gtype == 1) {
    // convex hull: deepest vertex = support point along -z.  Steepest-descent walk on the hull's vertex graph, warm-started
    // from the previous step's support vertex (exact: on a convex polytope a vertex with no lower neighbour is the global
    // minimum of a linear function).  Lane-dependent trip count: the warp reconverges below.
    const int adr = __float_as_int(role[RF_GVADR * CTA + tid]), num = __float_as_int(role[RF_GVNUM * CTA + tid]);
    int bi = hullv < num ? hullv : 0;
    float best = 3.0e38f;
    if (num > 0) { const float* hv = p.hull + 3 * (adr + bi); best = R[6] * __ldg(hv) + R[7] * __ldg(hv + 1) + R[8] * __ldg(hv + 2); }
    for (int moved = num > 0; moved;) {
      moved = 0;
      const int n0 = __ldg(p.hull_nbr_adr + adr + bi), n1 = __ldg(p.hull_nbr_adr + adr + bi + 1);
      int cand = bi;
      for (int e = n0; e < n1; e++) {
        const int v = __ldg(p.hull_nbr + e);
        const float* hv = p.hull + 3 * (adr + v);
        float z = R[6] * __ldg(hv) + R[7] * __ldg(hv + 1) + R[8] * __ldg(hv + 2);
        if (z < best) { best = z; cand = v; moved = 1; }
      }
      bi = cand;
    }
    hullv = bi;
    const float* hv = p.hull + 3 * (adr + bi);
    float h0 = __ldg(hv), h1 = __ldg(hv + 1), h2 = __ldg(hv + 2);
    d0 = best + xpos[2];
    pos0[0] = xpos[0] + R[0] * h0 + R[1] * h1 + R[2] * h2;
    pos0[1] = xpos[1] + R[3] * h0 + R[4] * h1 + R[5] * h2;
    pos0[2] = 0.5f * d0;
    a0 = (num > 0 && d0 <= p.margin) ? 1.f : 0.f;
  }
  __syncwarp(NMF_FULL);
  finish_contact(p, con[0], a0, d0, pos0, hx, hy, com, cvel, invw);
  finish_contact(p, con[1], a1, d1, pos1, hx, hy, com, cvel, invw);
  return a0 + a1;   // number of contacts of this geom (adhesion transmission divides by it)
}

// point "acceleration" of a contact for a body spatial vector S (ang, lin), projected on (n, mu t1, mu t2)
__device__ __forceinline__ void project_point(const Contact& c, const float* S, float mu, float* out) {
  float ax = S[3] + S[1] * c.r[2] - S[2] * c.r[1];
  float ay = S[4] + S[2] * c.r[0] - S[0] * c.r[2];
  float az = S[5] + S[0] * c.r[1] - S[1] * c.r[0];
  const float on = con_on(c);
  out[0] = on * az; out[1] = on * mu * (c.cx * ax + c.cy * ay); out[2] = on * mu * (-c.cy * ax + c.cx * ay);
}

// pyramid rows of one contact: jar_r = base +- w1 / w2
__device__ __forceinline__ void rows4(const float* w, float c0, float* jar) {
  float b = w[0] + c0;
  jar[0] = b + w[1]; jar[1] = b - w[1]; jar[2] = b + w[2]; jar[3] = b - w[2];
}

// contact forces for the current jar: accumulates the world wrench about the COM (ang, lin) into Wc and, when WITH_A,
// the contact augmentation A += X' W X (21 packed).  Branch-free (inactive slots have D = 0).
template <bool WITH_A>
__device__ __forceinline__ void contact_forces(const Contact& c, float mu, float* Wc, float* A, float* fn_out) {
  float jar[4]; rows4(c.w, c.c0, jar);
  float a[4], f[4];
#pragma unroll
  for (int r = 0; r < 4; r++) { a[r] = jar[r] < 0.f ? 1.f : 0.f; f[r] = -c.D * fminf(jar[r], 0.f); }
  float fn = f[0] + f[1] + f[2] + f[3], f1 = mu * (f[0] - f[1]), f2 = mu * (f[2] - f[3]);
  float F[3] = {f1 * c.cx - f2 * c.cy, f1 * c.cy + f2 * c.cx, fn};
  float T[3]; cross3(c.r, F, T);
  Wc[0] += T[0]; Wc[1] += T[1]; Wc[2] += T[2]; Wc[3] += F[0]; Wc[4] += F[1]; Wc[5] += F[2];
  if (fn_out) *fn_out = fn;
  if (WITH_A) {
    float s1 = a[0] + a[1], s2 = a[2] + a[3], d1 = a[0] - a[1], d2 = a[2] - a[3], m2 = mu * mu;
    float W[9];
    W[0] = c.D * m2 * (s1 * c.cx * c.cx + s2 * c.cy * c.cy);
    W[4] = c.D * m2 * (s1 * c.cy * c.cy + s2 * c.cx * c.cx);
    W[1] = W[3] = c.D * m2 * (s1 - s2) * c.cx * c.cy;
    W[8] = c.D * (s1 + s2);
    W[2] = W[6] = c.D * mu * (d1 * c.cx - d2 * c.cy);
    W[5] = W[7] = c.D * mu * (d1 * c.cy + d2 * c.cx);
    // T = [r]x W  (columns: r x W[:,j]);  A_ww = T (-[r]x) -> row i: r x T[i,:]
    float Tm[9];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      float col[3] = {W[j], W[3 + j], W[6 + j]}, t[3]; cross3(c.r, col, t);
      Tm[j] = t[0]; Tm[3 + j] = t[1]; Tm[6 + j] = t[2];
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      float row[3] = {Tm[3 * i], Tm[3 * i + 1], Tm[3 * i + 2]}, t[3]; cross3(c.r, row, t);
#pragma unroll
      for (int j = i; j < 3; j++) A[s6(i, j)] += t[j];
#pragma unroll
      for (int j = 0; j < 3; j++) A[s6(i, 3 + j)] += Tm[3 * i + j];
    }
    A[s6(3, 3)] += W[0]; A[s6(3, 4)] += W[1]; A[s6(3, 5)] += W[2]; A[s6(4, 4)] += W[4]; A[s6(4, 5)] += W[5]; A[s6(5, 5)] += W[8];
  }
}

// line-search partial sums of one contact at step alpha: d0 += D x jv, d1 += D jv^2 over rows with x < 0
template <class Con>
__device__ __forceinline__ void ls_eval(const Con& c, const float* sv, float alpha, float& d0, float& d1, float& nchanged) {
  float jar[4], jv[4]; rows4(c.w, c.c0, jar); rows4(sv, 0.f, jv);
#pragma unroll
  for (int r = 0; r < 4; r++) {
    float x = jar[r] + alpha * jv[r];
    float on = x < 0.f ? c.D : 0.f;
    d0 += on * x * jv[r]; d1 += on * jv[r] * jv[r];
    nchanged += ((x < 0.f) != (jar[r] < 0.f)) ? 1.f : 0.f;   // rows whose state differs from the one the Hessian was built for
  }
}

// ------------------------------------------------------------------ general-frame contact slot (terrain worlds)
// Same conventions as Contact, but the contact normal is arbitrary (box-column terrain has walls and edges):
// frame = (n, t1, t2 = n x t1).  Used by the TERRAIN instantiation of the step only; the flat-ground kernel keeps the
// cheaper z-normal slot above.
struct ContactG {    // 14 registers per slot
  float r[3];        // contact position relative to the subtree COM
  float n[3], t[3];  // normal, first tangent
  float D, c0;
  float w[3];
};
__device__ __forceinline__ float con_on(const ContactG& c) { return c.D > 0.f ? 1.f : 0.f; }

__device__ __forceinline__ void finish_contact(const StepParams& p, ContactG& c, float active, float dist, const float* pos, const float* nrm,
                                               const float* hint, const float* com, const float* cvel, float invw) {
  c.r[0] = pos[0] - com[0]; c.r[1] = pos[1] - com[1]; c.r[2] = pos[2] - com[2];
  c.n[0] = nrm[0]; c.n[1] = nrm[1]; c.n[2] = nrm[2];
  {  // first tangent: the hint (capsule axis) orthogonalised against the normal; world x (or y) when they are parallel
    float hn = dot3(hint, nrm), t[3] = {hint[0] - hn * nrm[0], hint[1] - hn * nrm[1], hint[2] - hn * nrm[2]};
    float t2 = dot3(t, t);
    if (t2 < 1e-12f) {
      const bool usex = fabsf(nrm[0]) < 0.9f;
      const float e[3] = {usex ? 1.f : 0.f, usex ? 0.f : 1.f, 0.f};
      hn = dot3(e, nrm); t[0] = e[0] - hn * nrm[0]; t[1] = e[1] - hn * nrm[1]; t[2] = e[2] - hn * nrm[2]; t2 = dot3(t, t);
    }
    const float inv = rsqrtf(t2);
    c.t[0] = t[0] * inv; c.t[1] = t[1] * inv; c.t[2] = t[2] * inv;
  }
  float imp = impedance(p, fabsf(dist - p.margin));
  float R0 = fmaxf(NMF_MINVAL, (1.f - imp) * invw * (1.f + p.mu * p.mu) / imp);
  c.D = active / (2.f * (p.mu * p.mu / p.impratio) * R0);
  c.c0 = active * p.cK * imp * (dist - p.margin);
  float vp[3] = {cvel[3] + cvel[1] * c.r[2] - cvel[2] * c.r[1], cvel[4] + cvel[2] * c.r[0] - cvel[0] * c.r[2],
                 cvel[5] + cvel[0] * c.r[1] - cvel[1] * c.r[0]};
  float t2v[3]; cross3(c.n, c.t, t2v);
  c.w[0] = active * p.cB * dot3(c.n, vp);
  c.w[1] = active * p.cB * p.mu * dot3(c.t, vp);
  c.w[2] = active * p.cB * p.mu * dot3(t2v, vp);
}

// one sphere (centre c, radius rad) against the terrain solid = floor plane + grid of box columns: the closest point of
// the solid decides normal and distance (ONE contact per sphere).  Column (i, j) covers |x - i Px| <= hx, |y - j Py| <= hy,
// z <= top(i, j), top = terr[4 + ((i + j) & 1)].  A centre inside a column is pushed out through the top face.
__device__ __forceinline__ void sphere_terrain(const StepParams& p, const float* c, float rad, float* nrm, float& dist) {
  const float Px = p.terr[0], Py = p.terr[1], hx = p.terr[2], hy = p.terr[3];
  nrm[0] = 0.f; nrm[1] = 0.f; nrm[2] = 1.f; dist = c[2] - p.terr[6] - rad;      // floor plane
  const float fi = rintf(c[0] / Px), fj = rintf(c[1] / Py);
  const int i0 = (int)fi, j0 = (int)fj;
  const int sx = c[0] >= fi * Px ? 1 : -1, sy = c[1] >= fj * Py ? 1 : -1;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int i = i0 + ((q & 1) ? sx : 0), j = j0 + ((q & 2) ? sy : 0);
    const float cx = (float)i * Px, cy = (float)j * Py, top = ((i + j) & 1) ? p.terr[5] : p.terr[4];
    const float qx = fminf(fmaxf(c[0], cx - hx), cx + hx), qy = fminf(fmaxf(c[1], cy - hy), cy + hy), qz = fminf(c[2], top);
    const float dx = c[0] - qx, dy = c[1] - qy, dz = c[2] - qz, d2 = dx * dx + dy * dy + dz * dz;
    float d, n0, n1, n2;
    if (d2 > 0.f) { const float inv = rsqrtf(d2); d = d2 * inv - rad; n0 = dx * inv; n1 = dy * inv; n2 = dz * inv; }
    else { d = c[2] - top - rad; n0 = 0.f; n1 = 0.f; n2 = 1.f; }
    if (d < dist) { dist = d; nrm[0] = n0; nrm[1] = n1; nrm[2] = n2; }
  }
}

// capsule-vs-terrain narrow phase for the geom carried by this lane's body: the two end spheres, one contact each
__device__ __forceinline__ float collide(const StepParams& p, const float* role, int tid, const float* xpos, const float* R,
                                         const float* com, const float* cvel, float invw, ContactG* con, int&) {
  const int gtype = __float_as_int(role[RF_GTYPE * CTA + tid]);
  float gp[3] = {role[(RF_GPOS + 0) * CTA + tid], role[(RF_GPOS + 1) * CTA + tid], role[(RF_GPOS + 2) * CTA + tid]};
  float ga[3] = {role[(RF_GAXIS + 0) * CTA + tid], role[(RF_GAXIS + 1) * CTA + tid], role[(RF_GAXIS + 2) * CTA + tid]};
  const float rad = role[RF_GRAD * CTA + tid], half = role[RF_GHALF * CTA + tid];
  float c[3], a[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    c[i] = xpos[i] + R[3 * i] * gp[0] + R[3 * i + 1] * gp[1] + R[3 * i + 2] * gp[2];
    a[i] = R[3 * i] * ga[0] + R[3 * i + 1] * ga[1] + R[3 * i + 2] * ga[2];
  }
  float total = 0.f;
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const float sg = s == 0 ? half : -half;
    float e[3] = {c[0] + sg * a[0], c[1] + sg * a[1], c[2] + sg * a[2]}, nrm[3], dist;
    sphere_terrain(p, e, rad, nrm, dist);
    const float act = (gtype == 0 && dist <= p.margin) ? 1.f : 0.f;
    const float back = rad + 0.5f * dist;
    float pos[3] = {e[0] - back * nrm[0], e[1] - back * nrm[1], e[2] - back * nrm[2]};
    finish_contact(p, con[s], act, dist, pos, nrm, a, com, cvel, invw);
    total += act;
  }
  return total;
}

__device__ __forceinline__ void project_point(const ContactG& c, const float* S, float mu, float* out) {
  float a[3] = {S[3] + S[1] * c.r[2] - S[2] * c.r[1], S[4] + S[2] * c.r[0] - S[0] * c.r[2], S[5] + S[0] * c.r[1] - S[1] * c.r[0]};
  float t2[3]; cross3(c.n, c.t, t2);
  const float on = con_on(c);
  out[0] = on * dot3(c.n, a); out[1] = on * mu * dot3(c.t, a); out[2] = on * mu * dot3(t2, a);
}

template <bool WITH_A>
__device__ __forceinline__ void contact_forces(const ContactG& c, float mu, float* Wc, float* A, float* fn_out) {
  float jar[4]; rows4(c.w, c.c0, jar);
  float a[4], f[4];
#pragma unroll
  for (int r = 0; r < 4; r++) { a[r] = jar[r] < 0.f ? 1.f : 0.f; f[r] = -c.D * fminf(jar[r], 0.f); }
  float fn = f[0] + f[1] + f[2] + f[3], f1 = mu * (f[0] - f[1]), f2 = mu * (f[2] - f[3]);
  float t2[3]; cross3(c.n, c.t, t2);
  float F[3] = {fn * c.n[0] + f1 * c.t[0] + f2 * t2[0], fn * c.n[1] + f1 * c.t[1] + f2 * t2[1], fn * c.n[2] + f1 * c.t[2] + f2 * t2[2]};
  float T[3]; cross3(c.r, F, T);
  Wc[0] += T[0]; Wc[1] += T[1]; Wc[2] += T[2]; Wc[3] += F[0]; Wc[4] += F[1]; Wc[5] += F[2];
  if (fn_out) *fn_out = fn;
  if (WITH_A) {
    // W = D sum_r a_r d_r d_r',  d = n +- mu t1 | n +- mu t2
    const float s1 = a[0] + a[1], s2 = a[2] + a[3], d1 = mu * (a[0] - a[1]), d2 = mu * (a[2] - a[3]), m2 = mu * mu;
    float u[3] = {d1 * c.t[0] + d2 * t2[0], d1 * c.t[1] + d2 * t2[1], d1 * c.t[2] + d2 * t2[2]};   // n u' + u n' carries the cross terms
    float W[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = i; j < 3; j++) {
        float v = (s1 + s2) * c.n[i] * c.n[j] + c.n[i] * u[j] + u[i] * c.n[j] + m2 * (s1 * c.t[i] * c.t[j] + s2 * t2[i] * t2[j]);
        W[3 * i + j] = W[3 * j + i] = c.D * v;
      }
    float Tm[9];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      float col[3] = {W[j], W[3 + j], W[6 + j]}, t[3]; cross3(c.r, col, t);
      Tm[j] = t[0]; Tm[3 + j] = t[1]; Tm[6 + j] = t[2];
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      float row[3] = {Tm[3 * i], Tm[3 * i + 1], Tm[3 * i + 2]}, t[3]; cross3(c.r, row, t);
#pragma unroll
      for (int j = i; j < 3; j++) A[s6(i, j)] += t[j];
#pragma unroll
      for (int j = 0; j < 3; j++) A[s6(i, 3 + j)] += Tm[3 * i + j];
    }
    A[s6(3, 3)] += W[0]; A[s6(3, 4)] += W[1]; A[s6(3, 5)] += W[2]; A[s6(4, 4)] += W[4]; A[s6(4, 5)] += W[5]; A[s6(5, 5)] += W[8];
  }
}

// adhesion pull f (>= 0 towards the surface) of one contact: wrench of -f n at the contact point
__device__ __forceinline__ void adhesion_wrench(const ContactG& c, float f, float* W) {
  const float s = -con_on(c) * f;
  float F[3] = {s * c.n[0], s * c.n[1], s * c.n[2]}, T[3]; cross3(c.r, F, T);
  W[0] += T[0]; W[1] += T[1]; W[2] += T[2]; W[3] += F[0]; W[4] += F[1]; W[5] += F[2];
}
// signed distance of an active slot (debug dump only)
__device__ __forceinline__ float con_dist(const Contact& c, const float* com) { return 2.f * (c.r[2] + com[2]); }
__device__ __forceinline__ float con_dist(const ContactG&, const float*) { return 0.f; }

// ------------------------------------------------------------------ weld equality of the TetheredWorld (hub lane only)
// Six always-active rows on the free body: rows 0-2 = world position of the hub-frame point weld_a, rows 3-5 =
// torquescale * vec(q_hub * weld_q) with Jacobian G w = 0.5 ts vec((0, w) * q) (w = world angular velocity).  They only touch
// the hub's 6x6 block, so they enter exactly like a contact on the hub: a wrench in the gradient and an X'WX augmentation
// of the hub's spatial inertia.  State lives in shared memory (one lane uses it): r[3] G[9] D[6] c0[6] w[6] sv[6].
constexpr int WL_R = 0, WL_G = 3, WL_D = 12, WL_C0 = 18, WL_W = 24, WL_SV = 30, WL_COUNT = 36;

__device__ __forceinline__ void point_and_rot(const float* sw, const float* S, float* out) {   // J * (spatial vector of the hub)
  const float* r = sw + WL_R; const float* G = sw + WL_G;
  out[0] = S[3] + S[1] * r[2] - S[2] * r[1]; out[1] = S[4] + S[2] * r[0] - S[0] * r[2]; out[2] = S[5] + S[0] * r[1] - S[1] * r[0];
#pragma unroll
  for (int i = 0; i < 3; i++) out[3 + i] = G[3 * i] * S[0] + G[3 * i + 1] * S[1] + G[3 * i + 2] * S[2];
}
__device__ __forceinline__ void weld_setup(const StepParams& p, float* sw, const float* qh, const float* xh, const float* com,
                                           const float* cvel, const float* Sa) {
  float a[3], q[4], pos[6];
  qrot(qh, p.weld_a, a); qmul(qh, p.weld_q, q);
  const float h = 0.5f * p.weld_ts;
#pragma unroll
  for (int i = 0; i < 3; i++) { pos[i] = xh[i] + a[i]; sw[WL_R + i] = pos[i] - com[i]; pos[3 + i] = p.weld_ts * q[1 + i]; }
  float* G = sw + WL_G;
  G[0] = h * q[0]; G[1] = h * q[3]; G[2] = -h * q[2];
  G[3] = -h * q[3]; G[4] = h * q[0]; G[5] = h * q[1];
  G[6] = h * q[2]; G[7] = -h * q[1]; G[8] = h * q[0];
  float jv[6], ja[6];
  point_and_rot(sw, cvel, jv); point_and_rot(sw, Sa, ja);
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const float imp = impedance_of(p.weld_imp, fabsf(pos[i]));
    const float Rr = fmaxf(NMF_MINVAL, (1.f - imp) * p.weld_invw[i >= 3 ? 1 : 0] / imp);
    sw[WL_D + i] = 1.f / Rr;
    sw[WL_C0 + i] = p.weld_K * imp * pos[i];
    sw[WL_W + i] = p.weld_B * jv[i] + ja[i];
  }
}
// forces of the six rows for the current acceleration: wrench about the COM into Wc, Hessian augmentation into A
__device__ __forceinline__ void weld_forces(const float* sw, float* Wc, float* A) {
  const float* r = sw + WL_R; const float* G = sw + WL_G; const float* D = sw + WL_D;
  float f[6];
#pragma unroll
  for (int i = 0; i < 6; i++) f[i] = -D[i] * (sw[WL_W + i] + sw[WL_C0 + i]);
  float T[3]; cross3(r, f, T);
#pragma unroll
  for (int i = 0; i < 3; i++) { Wc[i] += T[i] + G[i] * f[3] + G[3 + i] * f[4] + G[6 + i] * f[5]; Wc[3 + i] += f[i]; }
  // position rows: W = diag(D0, D1, D2) at the point r  ->  X' W X ; rotation rows: G' diag(D3..5) G on the angular block
  float Tm[9];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    float col[3] = {j == 0 ? D[0] : 0.f, j == 1 ? D[1] : 0.f, j == 2 ? D[2] : 0.f}, t[3]; cross3(r, col, t);
    Tm[j] = t[0]; Tm[3 + j] = t[1]; Tm[6 + j] = t[2];
  }
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float row[3] = {Tm[3 * i], Tm[3 * i + 1], Tm[3 * i + 2]}, t[3]; cross3(r, row, t);
#pragma unroll
    for (int j = i; j < 3; j++) A[s6(i, j)] += t[j] + G[i] * D[3] * G[j] + G[3 + i] * D[4] * G[3 + j] + G[6 + i] * D[5] * G[6 + j];
#pragma unroll
    for (int j = 0; j < 3; j++) A[s6(i, 3 + j)] += Tm[3 * i + j];
  }
  A[s6(3, 3)] += D[0]; A[s6(4, 4)] += D[1]; A[s6(5, 5)] += D[2];
}
// line-search sums of the (always active) rows at step alpha
__device__ __forceinline__ void weld_ls(const float* sw, float alpha, float& d0, float& d1) {
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const float jv = sw[WL_SV + i], x = sw[WL_W + i] + sw[WL_C0 + i] + alpha * jv;
    d0 += sw[WL_D + i] * x * jv; d1 += sw[WL_D + i] * jv * jv;
  }
}

// ------------------------------------------------------------------ shared-memory plan (floats)
// The 64 lanes form 8 shuffle groups of 8: groups 0..5 are the leg chains, groups 6..7 are
// "hub chains": massless, joint-less bodies welded to the hub that carry the hub's contact
// geoms (lane 48 additionally carries the hub's inertia).  Every chain scan therefore runs
// convergently on all lanes with the full warp mask; hub-specific work (the six free-joint
// DoFs and the 6x6 Schur block) is smem-only code executed by the hub lanes afterwards.
// Lane-dependent conditionals in the chain code are written as selects / predicated stores so
// that warps do not diverge (divergence doubles the issue cost and slows every *_sync collective).
constexpr int NGROUP = 8;
constexpr int SM_STATE = 0;                         // S_STRIDE
constexpr int CDS = 12;                             // floats per cdof slot: 48-byte stride keeps the 8 lanes of a chain on distinct banks for 16-byte loads
constexpr int SM_CDOF = SM_STATE + S_STRIDE;        // NV * CDS
constexpr int SM_FS = SM_CDOF + NV * CDS;           // qfrc_smooth
constexpr int SM_GRAD = SM_FS + NV;                 // gradient / rhs
constexpr int SM_X = SM_GRAD + NV;                  // search direction / solve result
constexpr int SM_U = SM_X + NV;                     // u_i = P cdof_i of every chain DoF: NGROUP * 11 * 8 (reused as pose buffer)
constexpr int U_STRIDE = NLEGDOF * 8;
constexpr int SM_ROOT = SM_U + NGROUP * U_STRIDE;   // per-group root publications
constexpr int ROOT_STRIDE = 40;                     // [0..5] wrench, [6..15] crb / fc wrench, [16..36] A-hat
constexpr int SM_BASE = SM_ROOT + NGROUP * ROOT_STRIDE;  // NLEG*21 Schur contributions, then NLEG*6 rhs contributions
constexpr int SM_HBB = SM_BASE + 168;               // 21 hub block + 6 xb + 6 S_h
constexpr int SM_HUB = SM_HBB + 112;                // hub uniforms
constexpr int SM_RED = SM_HUB + 64;                 // 32
constexpr int SM_MBAR = SM_RED + 32;                // 8-byte mbarrier for the TMA record load (16-byte slot)
constexpr int SM_WELD = SM_MBAR + 4;                // weld rows of the tethered world (WL_COUNT)
constexpr int SM_TOTAL = SM_WELD + WL_COUNT;
constexpr int HU_CVEL = 0, HU_CACC = 6;
constexpr int HB_S = 0, HB_XB = 21, HB_SH = 27, HB_TOT = 33, HB_SR = 72;   // TOT: 37 root totals; SR: assembled Schur block (21) + rhs (6)

// Lane-constant description of the two matrix columns (of 16) + the shared last one a lane holds.
struct Cols {
  const float *cd0, *cd1, *cd10;   // cdof of column t, column 8+t (leg dof t+2), leg dof 10
  float add0, add1, add10;         // diagonal additions (armature [+ dt damping]) of those DoFs
};

// H columns of this lane from the staged u_i = P_i cdof_i :  H[i][c] = cdof_c . u_i   (c <= 6 + i)
__device__ __forceinline__ void load_columns(const float* su, const Cols& cl, int t, float* hk0, float* hk1, float& d10) {
  float c0[6], c1[6];
#pragma unroll
  for (int i = 0; i < 6; i++) { c0[i] = cl.cd0[i]; c1[i] = cl.cd1[i]; }
#pragma unroll
  for (int i = 0; i < NLEGDOF; i++) {
    const float* u = su + 8 * i;
    float u6[6] = {u[0], u[1], u[2], u[3], u[4], u[5]};
    float v0 = dot6(c0, u6), v1 = dot6(c1, u6);
    hk0[i] = (t <= 6 + i) ? v0 : 0.f;
    hk1[i] = (t + 2 <= i) ? v1 : 0.f;
    if (6 + i == t) hk0[i] += cl.add0;
    if (i == t + 2) hk1[i] += cl.add1;
  }
  d10 = dot6(cl.cd10, su + 8 * 10) + cl.add10;
}

// L'DL of the 11x11 chain block + its 11x6 border, one matrix column per lane (columns t and 8+t of 16; the last
// diagonal entry d10 is held by every lane).  Leaves L (unit lower, scaled rows) in hk0/hk1, the inverse pivots of
// the lane's own DoFs in i0own/i1own/i10 and this chain's Schur contribution to the hub block in contrib[3].
__device__ __forceinline__ void chain_factor(float* hk0, float* hk1, float d10, int t, const int* pb, const int* pc, float& i0own, float& i1own,
                                             float& i10, float* contrib) {
  contrib[0] = contrib[1] = contrib[2] = 0.f;
#pragma unroll
  for (int kk = NLEGDOF - 1; kk >= 0; kk--) {
    float dk;
    if (kk == 10) dk = d10; else { const int c = 6 + kk; dk = __shfl_sync(NMF_FULL, c < 8 ? hk0[kk] : hk1[kk], c & 7, 8); }
    float ik = 1.0f / dk;
    if (kk == 10) i10 = ik;
    if (6 + kk == t) i0own = ik;
    if (kk == t + 2) i1own = ik;
    float l0 = hk0[kk] * ik, l1 = hk1[kk] * ik;
#pragma unroll
    for (int s = 0; s < 3; s++) {
      float lb = __shfl_sync(NMF_FULL, l0, pb[s], 8), hc = __shfl_sync(NMF_FULL, hk0[kk], pc[s], 8);
      contrib[s] += lb * hc;
    }
#pragma unroll
    for (int j = 0; j < kk; j++) {
      const int cj = 6 + j;
      float l = __shfl_sync(NMF_FULL, cj < 8 ? l0 : l1, cj & 7, 8);
      hk0[j] -= (t <= cj ? l : 0.f) * hk0[kk];
      hk1[j] -= (t + 2 <= j ? l : 0.f) * hk1[kk];
    }
    hk0[kk] = l0; hk1[kk] = l1;
  }
}
// x <- L^-T x on the chain; lanes t < 6 return (in x0) minus the chain's contribution to the hub right-hand side
__device__ __forceinline__ void chain_solve_up(const float* hk0, const float* hk1, int t, float& x0, float& x1, float x10) {
#pragma unroll
  for (int kk = NLEGDOF - 1; kk >= 0; kk--) {
    float xk;
    if (kk == 10) xk = x10; else { const int c = 6 + kk; xk = __shfl_sync(NMF_FULL, c < 8 ? x0 : x1, c & 7, 8); }
    x0 -= (t < 6 + kk ? hk0[kk] : 0.f) * xk;
    x1 -= (t + 2 < kk ? hk1[kk] : 0.f) * xk;
  }
}
// x <- L^-1 D^-1 x given the hub solution xb (lanes t < 6)
__device__ __forceinline__ void chain_solve_down(const float* hk0, const float* hk1, int t, float xb, float i0own, float i1own,
                                                 float i10, float& x0, float& x1, float& x10) {
  x0 = t >= 6 ? x0 * i0own : x0;
  x1 *= i1own; x10 *= i10;
#pragma unroll
  for (int kk = 0; kk < NLEGDOF; kk++) {
    float part = (t < 6 ? hk0[kk] * xb : (t - 6 < kk ? hk0[kk] * x0 : 0.f)) + (t + 2 < kk ? hk1[kk] * x1 : 0.f);
    part += __shfl_xor_sync(NMF_FULL, part, 1, 8); part += __shfl_xor_sync(NMF_FULL, part, 2, 8); part += __shfl_xor_sync(NMF_FULL, part, 4, 8);
    x0 = (6 + kk == t) ? x0 - part : x0;
    x1 = (kk == t + 2) ? x1 - part : x1;
    if (kk == 10) x10 -= part;
  }
}
// sum of entries [lo, lo+n) of the 8 chain-root records, computed cooperatively by the 16 hub lanes into SM_HBB + HB_TOT
__device__ __forceinline__ void hub_root_totals(float* sm, int hl, int lo, int n) {
  for (int i = hl; i < n; i += NHUBLANE) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < NGROUP; g++) s += sm[SM_ROOT + g * ROOT_STRIDE + lo + i];
    sm[SM_HBB + HB_TOT + i] = s;
  }
}
// hub 6x6 block: S = Hbb - sum(chain contributions); solve S xb = rhs (dense L'DL, serial, one lane)
__device__ __forceinline__ void hub_solve(float* sm, float* xb) {
  float S[21];
#pragma unroll
  for (int i = 0; i < 21; i++) S[i] = sm[SM_HBB + HB_SR + i];
#pragma unroll
  for (int b = 0; b < 6; b++) xb[b] = sm[SM_HBB + HB_SR + 21 + b];
  float dinv[6];
#pragma unroll
  for (int kk = 5; kk >= 0; kk--) {
    dinv[kk] = 1.0f / S[kk * (kk + 1) / 2 + kk];
#pragma unroll
    for (int j = 0; j < kk; j++) {
      float l = S[kk * (kk + 1) / 2 + j] * dinv[kk];
#pragma unroll
      for (int c = 0; c <= j; c++) S[j * (j + 1) / 2 + c] -= l * S[kk * (kk + 1) / 2 + c];
    }
#pragma unroll
    for (int j = 0; j < kk; j++) S[kk * (kk + 1) / 2 + j] *= dinv[kk];
  }
#pragma unroll
  for (int kk = 5; kk >= 0; kk--)
#pragma unroll
    for (int j = 0; j < kk; j++) xb[j] -= S[kk * (kk + 1) / 2 + j] * xb[kk];
#pragma unroll
  for (int kk = 0; kk < 6; kk++) xb[kk] *= dinv[kk];
#pragma unroll
  for (int kk = 0; kk < 6; kk++)
#pragma unroll
    for (int j = 0; j < kk; j++) xb[kk] -= S[kk * (kk + 1) / 2 + j] * xb[j];
}

// factor + solve of the arrowhead system  H x = -rhs(SM_GRAD), H given by the staged u vectors (chains) and SM_HBB (hub);
// the result is written to SM_X.  WITH_SH: lane 48 also publishes the hub part of the spatial acceleration of x.
__device__ __forceinline__ void arrowhead_solve(float* sm, const float* s_cdof, const Cols& cl, int grp, int t, bool is_leg, int hl, int lbase,
                                                const int* pb, const int* pc, float* dbg_rows) {
  float hk0[NLEGDOF], hk1[NLEGDOF], d10, i0own = 0.f, i1own = 0.f, i10 = 0.f, contrib[3];
  load_columns(sm + SM_U + grp * U_STRIDE, cl, t, hk0, hk1, d10);
  if (dbg_rows) {
#pragma unroll
    for (int i = 0; i < NLEGDOF; i++) { dbg_rows[i * 16 + t] = hk0[i]; dbg_rows[i * 16 + 8 + t] = hk1[i]; }
    dbg_rows[176] = d10;
  }
  chain_factor(hk0, hk1, d10, t, pb, pc, i0own, i1own, i10, contrib);
#pragma unroll
  for (int s = 0; s < 3; s++) if (is_leg && t + 8 * s < 21) sm[SM_BASE + grp * 21 + t + 8 * s] = contrib[s];
  float x0 = (is_leg && t >= 6) ? -sm[SM_GRAD + lbase + t - 6] : 0.f, x1 = is_leg ? -sm[SM_GRAD + lbase + t + 2] : 0.f,
        x10 = is_leg ? -sm[SM_GRAD + lbase + 10] : 0.f;
  chain_solve_up(hk0, hk1, t, x0, x1, x10);
  if (is_leg && t < 6) sm[SM_BASE + NLEG * 21 + grp * 6 + t] = x0;
  block_sync();
  if (!is_leg) {   // Schur block and hub right-hand side assembled by the 16 hub lanes, then solved by lane 48
    for (int i = hl; i < 27; i += NHUBLANE) {
      float v;
      if (i < 21) { v = sm[SM_HBB + HB_S + i]; for (int l = 0; l < NLEG; l++) v -= sm[SM_BASE + l * 21 + i]; }
      else { const int b = i - 21; v = -sm[SM_GRAD + b]; for (int l = 0; l < NLEG; l++) v += sm[SM_BASE + NLEG * 21 + l * 6 + b]; }
      sm[SM_HBB + HB_SR + i] = v;
    }
  }
  __syncwarp(NMF_FULL);
  if (!is_leg && hl == 0) {
    float xb[6]; hub_solve(sm, xb);
    float Sh[6] = {0, 0, 0, 0, 0, 0};
    for (int b = 0; b < 6; b++) {
      sm[SM_HBB + HB_XB + b] = xb[b]; sm[SM_X + b] = xb[b];
      const float* cd = s_cdof + CDS * b; for (int i = 0; i < 6; i++) Sh[i] += cd[i] * xb[b];
    }
    for (int i = 0; i < 6; i++) sm[SM_HBB + HB_SH + i] = Sh[i];
  }
  block_sync();
  chain_solve_down(hk0, hk1, t, t < 6 ? sm[SM_HBB + HB_XB + t] : 0.f, i0own, i1own, i10, x0, x1, x10);
  float* sx = sm + SM_X + lbase;
  if (is_leg && t >= 6) sx[t - 6] = x0;
  if (is_leg) sx[t + 2] = x1;
  if (is_leg && t == 0) sx[10] = x10;
}

// ------------------------------------------------------------------ TMA (bulk async copy) staging of the state record
// One elected thread moves the whole 1216-byte record HBM <-> shared memory with cp.async.bulk (SASS: UBLKCP); the block
// waits on an mbarrier.  Under the SIMT emulator the same copies are plain loops.
#ifndef NMF_SIMT_EMU
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_load_record(float* dst_smem, const float* src_gmem, unsigned long long* mbar, int tid) {
  const unsigned bar = smem_u32(mbar), dst = smem_u32(dst_smem);
  constexpr unsigned bytes = S_STRIDE * sizeof(float);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  block_sync();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src_gmem), "r"(bytes), "r"(bar) : "memory");
  }
  unsigned done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
  }
  block_sync();
  if (tid == 0) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory");   // the slot is re-initialised by the next work item
}
// `published`: the caller hands the record to another block afterwards (work-queue scheduling), so wait until the
// global writes have completed, not only until shared memory has been read.
__device__ __forceinline__ void tma_store_record(float* dst_gmem, const float* src_smem, int tid, bool published) {
  block_sync();                                                      // all generic-proxy writes to the record are done
  if (tid == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // make them visible to the async (TMA) proxy
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"((unsigned)(S_STRIDE * sizeof(float))) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (published) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must stay valid until the copy has read it
  }
}
#else
__device__ __forceinline__ void tma_load_record(float* dst_smem, const float* src_gmem, unsigned long long*, int tid) {
  for (int i = tid; i < S_STRIDE; i += CTA) dst_smem[i] = src_gmem[i];
  block_sync();
}
__device__ __forceinline__ void tma_store_record(float* dst_gmem, const float* src_smem, int tid, bool) {
  block_sync();
  for (int i = tid; i < S_STRIDE; i += CTA) dst_gmem[i] = src_smem[i];
}
#endif

// ------------------------------------------------------------------ the step
// Advances fly `fly` by steps [step0, step0 + nsub) of the launch's p.nsteps (one work item of the launch: the whole
// launch when flies map 1:1 to blocks, a sub-chunk under work-queue scheduling).
// WORLD selects the kernel instantiation: W_FLAT = the reference's FlatGroundWorld; W_TERRAIN = general-frame contact
// slots + capsule-vs-box-column narrow phase; W_TETHER = TetheredWorld (no ground, weld equality on the hub).
constexpr int W_FLAT = 0, W_TERRAIN = 1, W_TETHER = 2;
template <int WORLD>
__device__ __forceinline__ void step_block(const StepParams& p, float* sm, const int fly, const int step0, const int nsub, const bool published) {
  using Con = typename std::conditional<WORLD == W_TERRAIN, ContactG, Contact>::type;
  constexpr bool TETHER = WORLD == W_TETHER;
  float* sw = sm + SM_WELD;
  const int tid = threadIdx.x;
  const bool weld_lane = TETHER && tid == NLEG * NLINK;
  const int grp = tid >> 3, k = tid & 7, t = k;
  const bool is_leg = grp < NLEG;
  const int hl = tid - NLEG * NLINK;          // hub lane index (valid when !is_leg)
  const bool hubdof = !is_leg && hl < 6;      // hub lanes 0..5 own the six free-joint DoFs
  const float* role = p.role;
  float* st = sm + SM_STATE;
  float* s_cdof = sm + SM_CDOF;
  float* s_hub = sm + SM_HUB;
  float* s_red = sm + SM_RED;
  float* su = sm + SM_U + grp * U_STRIDE;
  float* rt = sm + SM_ROOT + grp * ROOT_STRIDE;
  int parity = 0;

  // ---- load the state record (one TMA bulk copy of 1216 B), clear the u staging (hub chains keep u = 0)
  for (int i = tid; i < NGROUP * U_STRIDE; i += CTA) sm[SM_U + i] = 0.f;
  tma_load_record(st, p.state + (size_t)fly * S_STRIDE, reinterpret_cast<unsigned long long*>(sm + SM_MBAR), tid);
  block_sync();

  // per-lane constants that stay in registers for the whole launch
  const int ndof = __float_as_int(role[RF_NDOF * CTA + tid]);                 // 0 on hub lanes
  const int dof0 = is_leg ? __float_as_int(role[RF_DOF0 * CTA + tid]) : (hl < 6 ? hl : 0);
  const int ldof0 = is_leg ? dof0 - 6 - NLEGDOF * grp : 0;                    // first dof index inside the leg
  const int lbase = is_leg ? 6 + NLEGDOF * grp : 6;                           // first global dof of this chain
  const float mass = role[RF_MASS * CTA + tid];
  const float invw = role[RF_INVW * CTA + tid];
  float armv[3], msk[3];          // per own-dof constants; msk[j] = 1 if the lane owns a j-th dof
  int dj[3];                                // global dof index of own dof j (clamped to a valid one when masked)
#pragma unroll
  for (int j = 0; j < 3; j++) {
    armv[j] = role[(RF_ARM + j) * CTA + tid];
    msk[j] = j < ndof ? 1.f : 0.f; dj[j] = j < ndof ? dof0 + j : dof0;
  }
#define CDO(j) (s_cdof + CDS * dj[j])   /* cdof of own dof j (a valid, masked address when the lane has fewer dofs) */
  Cols cl, cle;   // Newton (armature) and Euler (armature + dt damping) column descriptions
  {
    const int g0 = t < 6 ? t : lbase + t - 6;
    cl.cd0 = s_cdof + CDS * g0; cl.cd1 = s_cdof + CDS * (lbase + t + 2); cl.cd10 = s_cdof + CDS * (lbase + 10);
    cl.add0 = role[(RF_CARM + 0) * CTA + tid]; cl.add1 = role[(RF_CARM + 1) * CTA + tid]; cl.add10 = role[(RF_CARM + 2) * CTA + tid];
    cle = cl;   // (only the three diagonal additions differ; the compiler keeps one copy of the pointers)
    cle.add0 += p.dt * role[(RF_CDMP + 0) * CTA + tid]; cle.add1 += p.dt * role[(RF_CDMP + 1) * CTA + tid]; cle.add10 += p.dt * role[(RF_CDMP + 2) * CTA + tid];
  }
  int hullv = 0;      // support vertex of this lane's hull geom, carried from step to step as the warm start of the hill climb
  int pb[3], pc[3];   // (b >= c) pairs number t, t+8, t+16 of the 21 lower-triangular hub entries
#pragma unroll
  for (int s = 0; s < 3; s++) {
    int idx = t + 8 * s, b = 0;
#pragma unroll
    for (int bb = 1; bb < 6; bb++) b += (bb * (bb + 1) / 2 <= idx) ? 1 : 0;
    pb[s] = idx < 21 ? b : 0; pc[s] = idx < 21 ? idx - b * (b + 1) / 2 : 0;
  }

  for (int step = step0; step < step0 + nsub; step++) {
    // ---- controls for this step
    if (p.act_table) {
      const float* row = p.act_table + ((size_t)fly * p.table_T + (size_t)((p.table_t0 + step) % p.table_T)) * p.table_cols;
      for (int i = tid; i < p.table_cols; i += CTA) st[S_CTRL + i] = row[i];
      block_sync();
    }

    // =====================================================================
    // A. kinematics: scan of rigid transforms along each chain
    // =====================================================================
    float qh[4] = {st[S_QPOS + 3], st[S_QPOS + 4], st[S_QPOS + 5], st[S_QPOS + 6]};
    qnormalize(qh);
    const float xh[3] = {st[S_QPOS], st[S_QPOS + 1], st[S_QPOS + 2]};
    float xpos[3], xq[4], R[9], laxis[9];   // laxis: hinge axes in the parent frame, later world
    {
      float q[4] = {role[(RF_BQUAT + 0) * CTA + tid], role[(RF_BQUAT + 1) * CTA + tid], role[(RF_BQUAT + 2) * CTA + tid], role[(RF_BQUAT + 3) * CTA + tid]};
      float pp[3] = {role[(RF_BPOS + 0) * CTA + tid], role[(RF_BPOS + 1) * CTA + tid], role[(RF_BPOS + 2) * CTA + tid]};
#pragma unroll
      for (int j = 0; j < 3; j++) {
        float ax[3] = {role[(RF_AXIS + 3 * j) * CTA + tid], role[(RF_AXIS + 3 * j + 1) * CTA + tid], role[(RF_AXIS + 3 * j + 2) * CTA + tid]};
        qrot(q, ax, laxis + 3 * j);
        float ang = msk[j] * st[S_QPOS + 1 + dj[j]], sn, cs; sincos_small(0.5f * ang, &sn, &cs);   // masked dof: identity rotation
        float ql[4] = {cs, ax[0] * sn, ax[1] * sn, ax[2] * sn}, nq[4];
        qmul(q, ql, nq); q[0] = nq[0]; q[1] = nq[1]; q[2] = nq[2]; q[3] = nq[3];
      }
      {  // seed the chain root with the hub pose
        float t3[3], nq[4]; qrot(qh, pp, t3); qmul(qh, q, nq);
        const bool root = k == 0;
#pragma unroll
        for (int i = 0; i < 3; i++) pp[i] = root ? xh[i] + t3[i] : pp[i];
#pragma unroll
        for (int i = 0; i < 4; i++) q[i] = root ? nq[i] : q[i];
      }
#pragma unroll
      for (int off = 1; off < 8; off <<= 1) {
        float uq[4], up[3], t3[3], nq[4];
#pragma unroll
        for (int i = 0; i < 4; i++) uq[i] = __shfl_up_sync(NMF_FULL, q[i], off, 8);
#pragma unroll
        for (int i = 0; i < 3; i++) up[i] = __shfl_up_sync(NMF_FULL, pp[i], off, 8);
        qrot(uq, pp, t3); qmul(uq, q, nq);
        const bool on = k >= off;
#pragma unroll
        for (int i = 0; i < 3; i++) pp[i] = on ? up[i] + t3[i] : pp[i];
#pragma unroll
        for (int i = 0; i < 4; i++) q[i] = on ? nq[i] : q[i];
      }
      qnormalize(q);
      {  // parent world orientation -> world hinge axes
        float qpar[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { float pq = __shfl_up_sync(NMF_FULL, q[i], 1, 8); qpar[i] = k > 0 ? pq : qh[i]; }
#pragma unroll
        for (int j = 0; j < 3; j++) { float w[3]; qrot(qpar, laxis + 3 * j, w); laxis[3 * j] = w[0]; laxis[3 * j + 1] = w[1]; laxis[3 * j + 2] = w[2]; }
      }
      xpos[0] = pp[0]; xpos[1] = pp[1]; xpos[2] = pp[2]; xq[0] = q[0]; xq[1] = q[1]; xq[2] = q[2]; xq[3] = q[3];
    }
    q2mat(xq, R);

    // ---- subtree COM (block reduction), inertial quantities about it
    float xipos[3];
    {
      float ip[3] = {role[(RF_IPOS + 0) * CTA + tid], role[(RF_IPOS + 1) * CTA + tid], role[(RF_IPOS + 2) * CTA + tid]};
#pragma unroll
      for (int i = 0; i < 3; i++) xipos[i] = xpos[i] + R[3 * i] * ip[0] + R[3 * i + 1] * ip[1] + R[3 * i + 2] * ip[2];
    }
    float com[3];
    {
      float v[3] = {mass * xipos[0], mass * xipos[1], mass * xipos[2]};
      cta_reduce<3>(v, s_red, parity, tid);
      com[0] = v[0] * p.inv_total_mass; com[1] = v[1] * p.inv_total_mass; com[2] = v[2] * p.inv_total_mass;
    }
    float cinert[10];
    {
      float ib[6];
#pragma unroll
      for (int i = 0; i < 6; i++) ib[i] = role[(RF_IB + i) * CTA + tid];
      float Ib[9] = {ib[0], ib[3], ib[4], ib[3], ib[1], ib[5], ib[4], ib[5], ib[2]}, T[9], G[9];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) T[3 * i + j] = R[3 * i] * Ib[j] + R[3 * i + 1] * Ib[3 + j] + R[3 * i + 2] * Ib[6 + j];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = i; j < 3; j++) G[3 * i + j] = T[3 * i] * R[3 * j] + T[3 * i + 1] * R[3 * j + 1] + T[3 * i + 2] * R[3 * j + 2];
      float off[3] = {xipos[0] - com[0], xipos[1] - com[1], xipos[2] - com[2]};
      float o2 = dot3(off, off);
      cinert[0] = G[0] + mass * (o2 - off[0] * off[0]); cinert[1] = G[4] + mass * (o2 - off[1] * off[1]); cinert[2] = G[8] + mass * (o2 - off[2] * off[2]);
      cinert[3] = G[1] - mass * off[0] * off[1]; cinert[4] = G[2] - mass * off[0] * off[2]; cinert[5] = G[5] - mass * off[1] * off[2];
      cinert[6] = mass * off[0]; cinert[7] = mass * off[1]; cinert[8] = mass * off[2]; cinert[9] = mass;
    }
    // cdof of own dofs -> shared (predicated stores); hub lanes 0..5 own the free-joint dofs
    // own cdof live in shared memory (s_cdof + 8*dj[j]); hub-dof lanes use s_cdof + 8*hl
    {
      float off[3] = {com[0] - xpos[0], com[1] - xpos[1], com[2] - xpos[2]};
#pragma unroll
      for (int j = 0; j < 3; j++) {
        float l[3]; cross3(laxis + 3 * j, off, l);
        if (j < ndof) {
          float* cd = s_cdof + CDS * dj[j];
          cd[0] = laxis[3 * j]; cd[1] = laxis[3 * j + 1]; cd[2] = laxis[3 * j + 2]; cd[3] = l[0]; cd[4] = l[1]; cd[5] = l[2];
        }
      }
      if (hubdof) {
        float* cd = s_cdof + CDS * hl;
        const int a = hl < 3 ? 0 : hl - 3;
        float ax[3] = {R[a], R[3 + a], R[6 + a]}, l[3]; cross3(ax, off, l);
        const bool tr = hl < 3;
        cd[0] = tr ? 0.f : ax[0]; cd[1] = tr ? 0.f : ax[1]; cd[2] = tr ? 0.f : ax[2];
        cd[3] = tr ? (hl == 0 ? 1.f : 0.f) : l[0]; cd[4] = tr ? (hl == 1 ? 1.f : 0.f) : l[1]; cd[5] = tr ? (hl == 2 ? 1.f : 0.f) : l[2];
      }
      if (!is_leg && hl == 0) {
        // hub velocity / bias acceleration (free joint: translations first, rotations against the updated velocity)
        float wl[3] = {st[S_QVEL + 3], st[S_QVEL + 4], st[S_QVEL + 5]};
        float cv0[6] = {0.f, 0.f, 0.f, st[S_QVEL], st[S_QVEL + 1], st[S_QVEL + 2]};
        float cacc[6] = {0.f, 0.f, 0.f, -p.gx, -p.gy, -p.gz};
        float cvel[6] = {cv0[0], cv0[1], cv0[2], cv0[3], cv0[4], cv0[5]};
        float Sh[6] = {0, 0, 0, st[S_WARM], st[S_WARM + 1], st[S_WARM + 2]};
#pragma unroll
        for (int a = 0; a < 3; a++) {
          float ax[3] = {R[a], R[3 + a], R[6 + a]}, l[3]; cross3(ax, off, l);
          float cd[6] = {ax[0], ax[1], ax[2], l[0], l[1], l[2]}, cdd[6];
          cross_motion(cv0, cd, cdd);
          float qa = st[S_WARM + 3 + a];
#pragma unroll
          for (int i = 0; i < 6; i++) { cacc[i] += cdd[i] * wl[a]; cvel[i] += cd[i] * wl[a]; Sh[i] += cd[i] * qa; }
        }
#pragma unroll
        for (int i = 0; i < 6; i++) { s_hub[HU_CVEL + i] = cvel[i]; s_hub[HU_CACC + i] = cacc[i]; sm[SM_HBB + HB_SH + i] = Sh[i]; }
      }
    }
    block_sync();   // cdof, hub cvel/cacc, S_h visible

    // =====================================================================
    // B. velocities, composite inertia, collision, bias + actuator forces (all lanes, convergent)
    // =====================================================================
    float crb[10], cvel[6], Sa[6];
    Con con[2];
    float fs_own[3] = {0.f, 0.f, 0.f};
    float actf[3] = {0.f, 0.f, 0.f}, adhf = 0.f;
    {
      float qv[3], loc[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < 3; j++) {
        qv[j] = msk[j] * st[S_QVEL + dj[j]];
#pragma unroll
        for (int i = 0; i < 6; i++) loc[i] += CDO(j)[i] * qv[j];
      }
      float pre[6] = {loc[0], loc[1], loc[2], loc[3], loc[4], loc[5]};
      chain_prefix<6>(pre, NMF_FULL, k);
      float cv[6], ad[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int i = 0; i < 6; i++) cv[i] = pre[i] - loc[i] + s_hub[HU_CVEL + i];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        float cdd[6]; cross_motion(cv, CDO(j), cdd);
#pragma unroll
        for (int i = 0; i < 6; i++) { ad[i] += cdd[i] * qv[j]; cv[i] += CDO(j)[i] * qv[j]; }
      }
#pragma unroll
      for (int i = 0; i < 6; i++) cvel[i] = cv[i];
      chain_prefix<6>(ad, NMF_FULL, k);
      float cacc[6];
#pragma unroll
      for (int i = 0; i < 6; i++) cacc[i] = ad[i] + s_hub[HU_CACC + i];
      // body wrench  W = -(I a + v x* I v)  (+ adhesion below)
      float t1[6], t2[6], t3[6], W[6];
      mul_inert(cinert, cacc, t1); mul_inert(cinert, cvel, t2); cross_force(cvel, t2, t3);
#pragma unroll
      for (int i = 0; i < 6; i++) W[i] = -(t1[i] + t3[i]);
      // composite inertia
#pragma unroll
      for (int i = 0; i < 10; i++) crb[i] = cinert[i];
      chain_suffix<10>(crb, NMF_FULL, k);
      // collision for this body's geom
      const float ncon_lane = collide(p, role, tid, xpos, R, com, cvel, invw, con, hullv);
      // adhesion (body transmission): force pulls the body onto the plane along each contact normal
      {
        const int acidx = __float_as_int(role[RF_ADH_CIDX * CTA + tid]);
        float c = fminf(role[RF_ADH_HI * CTA + tid], fmaxf(role[RF_ADH_LO * CTA + tid], st[S_CTRL + (acidx >= 0 ? acidx : 0)]));
        adhf = role[RF_ADH_GAIN * CTA + tid] * c;     // gain = 0 on lanes without an adhesion actuator
        if constexpr (WORLD == W_TERRAIN) {
          const float pull = ncon_lane > 0.f ? adhf / ncon_lane : 0.f;
#pragma unroll
          for (int s = 0; s < 2; s++) adhesion_wrench(con[s], pull, W);
        } else {   // z-normal slots: written out in place (routing this through a helper cost 3 % on B200: register allocation)
          float fz = ncon_lane > 0.f ? -adhf / ncon_lane : 0.f;
#pragma unroll
          for (int s = 0; s < 2; s++) { float f = con_on(con[s]) * fz; W[0] += con[s].r[1] * f; W[1] -= con[s].r[0] * f; W[5] += f; }
        }
      }
      chain_suffix<6>(W, NMF_FULL, k);
      // joint-space smooth force of own dofs: passive + actuator + C'W
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int d = dj[j];
        float q = st[S_QPOS + 1 + d], qvj = st[S_QVEL + d];
        float f = -role[(RF_STIFF + j) * CTA + tid] * (q - role[(RF_SREF + j) * CTA + tid]) - role[(RF_DAMP + j) * CTA + tid] * qvj;
        const int ci = __float_as_int(role[(RF_CIDX + j) * CTA + tid]);
        float kp = role[(RF_KP + j) * CTA + tid], kv = role[(RF_KV + j) * CTA + tid];     // 0 without an actuator
        float af = kp * st[S_CTRL + (ci >= 0 ? ci : 0)] - kp * q - kv * qvj;
        af = fminf(role[(RF_FHI + j) * CTA + tid], fmaxf(role[(RF_FLO + j) * CTA + tid], af));
        actf[j] = af;
        f += af + dot6(CDO(j), W);
        fs_own[j] = f;
        if (j < ndof) sm[SM_FS + d] = f;
      }
      if (k == 0) {
#pragma unroll
        for (int i = 0; i < 6; i++) rt[i] = W[i];
#pragma unroll
        for (int i = 0; i < 10; i++) rt[6 + i] = crb[i];
      }
      // spatial acceleration of this body generated by the warm-start qacc
      float sl[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < 3; j++) {
        float qa = msk[j] * st[S_WARM + dj[j]];
#pragma unroll
        for (int i = 0; i < 6; i++) sl[i] += CDO(j)[i] * qa;
      }
      chain_prefix<6>(sl, NMF_FULL, k);
#pragma unroll
      for (int i = 0; i < 6; i++) Sa[i] = sl[i] + sm[SM_HBB + HB_SH + i];
      // contact rows at the warm-start acceleration
#pragma unroll
      for (int s = 0; s < 2; s++) {
        float ap[3]; project_point(con[s], Sa, p.mu, ap);
        con[s].w[0] += ap[0]; con[s].w[1] += ap[1]; con[s].w[2] += ap[2];
      }
      if (TETHER && weld_lane) weld_setup(p, sw, qh, xh, com, s_hub + HU_CVEL, Sa);
    }
    const bool any0 = __any_sync(NMF_FULL, con[0].D > 0.f), any1 = __any_sync(NMF_FULL, con[1].D > 0.f);
    block_sync();   // chain roots (wrench, crb) visible to the hub lanes
    float crbh[10];    // hub-dof lanes: composite inertia of the whole fly
    if (!is_leg) hub_root_totals(sm, hl, 0, 16);
    __syncwarp(NMF_FULL);
    if (hubdof) {
      float W[6];
#pragma unroll
      for (int i = 0; i < 10; i++) crbh[i] = sm[SM_HBB + HB_TOT + 6 + i];
#pragma unroll
      for (int i = 0; i < 6; i++) W[i] = sm[SM_HBB + HB_TOT + i];
      fs_own[0] = dot6(s_cdof + CDS * hl, W); sm[SM_FS + hl] = fs_own[0];
    }
    block_sync();   // roots consumed before the solver overwrites them

    // =====================================================================
    // C. soft-contact solve: primal Newton on  1/2 (a-a0)'M(a-a0) + s(Ja - aref)
    //    started from qacc_warmstart (unique minimiser => same result as the
    //    reference's mj_solNewton; solver=Newton in mujoco_globals.yaml:12)
    // =====================================================================
    float* qacc = st + S_WARM;      // qacc lives in the warm-start slot of the record
    int niter = 0, nls_total = 0, nchanged_last = 0;
    // One loop body serves every Newton iteration AND the final implicit-damping (Euler) solve, so the large unrolled
    // factorisation exists once in the instruction stream (the kernel is I-cache sensitive):
    //   pass `iter`:  forces(qacc) -> gradient/fc -> [converged? euler : newton] system -> arrowhead solve -> (line search, move)
    for (int iter = 0;; iter++) {
      const bool euler = iter > 0 && (nchanged_last == 0 || iter >= p.max_newton);
      // ---- forces, active set, contact augmentation
      float Wc[6] = {0, 0, 0, 0, 0, 0}, A[21];
#pragma unroll
      for (int i = 0; i < 21; i++) A[i] = 0.f;
      if (any0) contact_forces<true>(con[0], p.mu, Wc, A, nullptr);    // warp-uniform skips: most lanes have no contact
      if (any1) contact_forces<true>(con[1], p.mu, Wc, A, nullptr);
      if (TETHER && weld_lane) weld_forces(sw, Wc, A);
      // ---- gradient  g = C' suffix(I S - Wc) + armature a - fs ;  fc = C' suffix(Wc)
      float y[12];
      {
        float t6[6]; mul_inert(cinert, Sa, t6);
#pragma unroll
        for (int i = 0; i < 6; i++) { y[i] = t6[i] - Wc[i]; y[6 + i] = Wc[i]; }
      }
      float gown[3] = {0.f, 0.f, 0.f};
      chain_suffix<12>(y, NMF_FULL, k);
      if (!euler) chain_suffix<21>(A, NMF_FULL, k);
      const float am = euler ? 0.f : 1.f;     // the Euler system uses the plain inertia (no contact augmentation)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int d = dj[j];
        float fc = dot6(CDO(j), y + 6);
        float g = dot6(CDO(j), y) + armv[j] * qacc[d] - fs_own[j];
        gown[j] = msk[j] * g;
        if (j < ndof) sm[SM_GRAD + d] = euler ? -(fs_own[j] + fc) : g;   // right-hand side is -(this slot)
      }
      if (k == 0) {
#pragma unroll
        for (int i = 0; i < 12; i++) rt[i] = y[i];
#pragma unroll
        for (int i = 0; i < 21; i++) rt[16 + i] = am * A[i];
      }
      block_sync();
      {
        float P[21];
        if (!is_leg) hub_root_totals(sm, hl, 0, 37);
        __syncwarp(NMF_FULL);
        if (hubdof) {
          float yh[12];
#pragma unroll
          for (int i = 0; i < 12; i++) yh[i] = sm[SM_HBB + HB_TOT + i];
          float fc = dot6(s_cdof + CDS * hl, yh + 6), g = dot6(s_cdof + CDS * hl, yh) - fs_own[0];
          gown[0] = g; sm[SM_GRAD + hl] = euler ? -(fs_own[0] + fc) : g;
          expand_inert(crbh, P);
#pragma unroll
          for (int i = 0; i < 21; i++) P[i] += sm[SM_HBB + HB_TOT + 16 + i];
          float u[6]; sym6_mul(P, s_cdof + CDS * hl, u);
          for (int c = 0; c <= hl; c++) sm[SM_HBB + HB_S + hl * (hl + 1) / 2 + c] = dot6(s_cdof + CDS * c, u);
        }
        __syncwarp(NMF_FULL);
        // ---- system matrix  H = C'(crb + A-hat)C + diag : each DoF owner stages u = P cdof, columns are formed by the factoriser
        expand_inert(crb, P);
#pragma unroll
        for (int i = 0; i < 21; i++) P[i] += am * A[i];
#pragma unroll
        for (int j = 0; j < 3; j++) {
          float u[6]; sym6_mul(P, CDO(j), u);
          if (j < ndof) {
            float* up = su + 8 * (ldof0 + j);
#pragma unroll
            for (int i = 0; i < 6; i++) up[i] = u[i];
          }
        }
      }
      __syncwarp(NMF_FULL);
      {
        Cols c = cl;
        c.add0 = euler ? cle.add0 : cl.add0; c.add1 = euler ? cle.add1 : cl.add1; c.add10 = euler ? cle.add10 : cl.add10;
        float* dbg_rows = (euler && p.dbg && is_leg) ? p.dbg + (size_t)fly * DBG_STRIDE + DBG_HROWS + grp * 177 : nullptr;
        arrowhead_solve(sm, s_cdof, c, grp, t, is_leg, hl, lbase, pb, pc, dbg_rows);
      }
      if (euler) { niter = iter; break; }
      __syncwarp(NMF_FULL);
      float sown[3];
#pragma unroll
      for (int j = 0; j < 3; j++) sown[j] = msk[j] * sm[SM_X + dj[j]];
      if (hubdof) sown[0] = sm[SM_X + hl];

      // ---- spatial acceleration of the search direction, row directions, quadratic terms
      float Ss[6];
      {
        float sl[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
          for (int i = 0; i < 6; i++) sl[i] += (is_leg ? CDO(j)[i] : 0.f) * sown[j];
        chain_prefix<6>(sl, NMF_FULL, k);
#pragma unroll
        for (int i = 0; i < 6; i++) Ss[i] = sl[i] + sm[SM_HBB + HB_SH + i];
      }
      float red[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // s.g , s'Ms , d0 rows(0), d1 rows(0), |s|^2
      float sv0[3] = {0.f, 0.f, 0.f}, sv1[3] = {0.f, 0.f, 0.f};   // row directions of the two contact slots along the search vector
      {
        float t6[6]; mul_inert(cinert, Ss, t6);
        red[1] += dot6(Ss, t6);
#pragma unroll
        for (int j = 0; j < 3; j++) { red[0] += sown[j] * gown[j]; red[1] += (is_leg ? armv[j] : 0.f) * sown[j] * sown[j]; red[4] += sown[j] * sown[j]; }
        float dummy = 0.f;
        if (any0) { project_point(con[0], Ss, p.mu, sv0); ls_eval(con[0], sv0, 0.f, red[2], red[3], dummy); }
        if (any1) { project_point(con[1], Ss, p.mu, sv1); ls_eval(con[1], sv1, 0.f, red[2], red[3], dummy); }
        if (TETHER && weld_lane) { point_and_rot(sw, Ss, sw + WL_SV); weld_ls(sw, 0.f, red[2], red[3]); }
      }
      cta_reduce<5>(red, s_red, parity, tid);
      // ---- exact line search along the Newton direction (safeguarded Newton on the derivative)
      float alpha = 0.f;
      nchanged_last = 0;
      {
        const float q1 = red[0] - red[2], q2 = red[1];
        float d0 = red[0], d1 = q2 + red[3], lo = 0.f, hi = 3.0e38f;
        const int nls = (red[4] > 1e-30f && d1 > 0.f) ? p.max_ls : 0;   // zero direction: nothing to search
        for (int it = 0; it < nls; it++) {
          if (it > 0 && (fabsf(d0) <= 2e-6f * d1 * fmaxf(fabsf(alpha), 1e-3f) || (hi < 1.0e38f && hi - lo <= 1e-6f * hi))) break;
          if (d0 < 0.f) lo = alpha; else hi = alpha;
          float nx = alpha - d0 / d1;
          if (nx <= lo || nx >= hi) nx = (hi > 1.0e38f) ? 2.f * fmaxf(alpha, 1.f) : 0.5f * (lo + hi);
          alpha = nx;
          float e[3] = {0.f, 0.f, 0.f};
          if (any0) ls_eval(con[0], sv0, alpha, e[0], e[1], e[2]);
          if (any1) ls_eval(con[1], sv1, alpha, e[0], e[1], e[2]);
          if (TETHER && weld_lane) weld_ls(sw, alpha, e[0], e[1]);
          cta_reduce<3>(e, s_red, parity, tid);
          d0 = q1 + alpha * q2 + e[0]; d1 = q2 + e[1];
          nchanged_last = (int)e[2];
          nls_total++;
        }
      }
      // ---- move
#pragma unroll
      for (int j = 0; j < 3; j++) if (j < ndof || (j == 0 && hubdof)) qacc[dj[j]] += alpha * sown[j];
#pragma unroll
      for (int i = 0; i < 6; i++) Sa[i] += alpha * Ss[i];
#pragma unroll
      for (int i = 0; i < 3; i++) { con[0].w[i] += alpha * sv0[i]; con[1].w[i] += alpha * sv1[i]; }
      if (TETHER && weld_lane) for (int i = 0; i < 6; i++) sw[WL_W + i] += alpha * sw[WL_SV + i];
    }
    // SM_X now holds the implicit-damping (Euler) acceleration  (M + dt diag(damping))^-1 (qfrc_smooth + qfrc_constraint)
    block_sync();

    // ---- optional outputs of this step (derived quantities belong to the pre-integration state, as in mj_step)
    const bool last_step = step == p.nsteps - 1;
    if (last_step) {
      if (p.dbg) {
        float* dg = p.dbg + (size_t)fly * DBG_STRIDE;
        if (tid == 0) { dg[DBG_NITER] = (float)niter; dg[DBG_NLS] = (float)nls_total; dg[DBG_NCHG] = (float)nchanged_last; }
        for (int i = tid; i < NV; i += CTA) { dg[DBG_FS + i] = sm[SM_FS + i]; dg[DBG_QACC + i] = qacc[i]; dg[DBG_FC + i] = -sm[SM_GRAD + i] - sm[SM_FS + i]; dg[DBG_QACCE + i] = sm[SM_X + i]; }
        if (tid < 21) dg[DBG_HROWS + NLEG * 177 + tid] = sm[SM_HBB + HB_S + tid];
        for (int s = 0; s < 2; s++) {
          float* c = dg + DBG_CON + (tid * 2 + s) * 6; float fn = 0.f, Wt[6] = {0, 0, 0, 0, 0, 0};
          contact_forces<false>(con[s], p.mu, Wt, nullptr, &fn);
          const float on = con_on(con[s]);
          c[0] = on; c[1] = on * con_dist(con[s], com);
          c[2] = on * (con[s].r[0] + com[0]); c[3] = on * (con[s].r[1] + com[1]);
          c[4] = on * (con[s].r[2] + com[2]); c[5] = fn;
        }
        for (int i = 0; i < 3; i++) dg[DBG_XPOS + tid * 3 + i] = xpos[i];
        for (int i = tid; i < NV * 6; i += CTA) dg[DBG_CDOF + i] = s_cdof[CDS * (i / 6) + i % 6];
        float nc[1] = {con_on(con[0]) + con_on(con[1])};
        cta_reduce<1>(nc, s_red, parity, tid);
        if (tid == 0) dg[DBG_NCON] = nc[0];
      }
      if (p.out_actf) {
        float* o = p.out_actf + (size_t)fly * (p.nu_pos + p.nu_adh);
#pragma unroll
        for (int j = 0; j < 3; j++) { int ci = __float_as_int(role[(RF_CIDX + j) * CTA + tid]); if (j < ndof && ci >= 0) o[ci] = actf[j]; }
        int ai = __float_as_int(role[RF_ADH_CIDX * CTA + tid]); if (ai >= 0) o[ai] = adhf;
      }
      if (p.out_sensor) {
        // per-leg contact sensor (world.py:311-331), reduce="netforce": found, force, torque, pos, normal, tangent
        const float sens = (is_leg && __float_as_int(role[RF_LEGSENSOR * CTA + tid]) != 0) ? 1.f : 0.f;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // F(3), fn-weighted pos(3), fn sum, count
        float Fc[2][3], plain[3] = {0, 0, 0};
#pragma unroll
        for (int s = 0; s < 2; s++) {
          float Wt[6] = {0, 0, 0, 0, 0, 0}, fn = 0.f; contact_forces<false>(con[s], p.mu, Wt, nullptr, &fn);
          const float on = sens * con_on(con[s]);
          Fc[s][0] = on * Wt[3]; Fc[s][1] = on * Wt[4]; Fc[s][2] = on * Wt[5];
#pragma unroll
          for (int i = 0; i < 3; i++) { acc[i] += Fc[s][i]; acc[3 + i] += on * fn * (con[s].r[i] + com[i]); plain[i] += on * (con[s].r[i] + com[i]); }
          acc[6] += on * fn; acc[7] += on;
        }
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) {
#pragma unroll
          for (int i = 0; i < 8; i++) acc[i] += __shfl_xor_sync(NMF_FULL, acc[i], off, 8);
#pragma unroll
          for (int i = 0; i < 3; i++) plain[i] += __shfl_xor_sync(NMF_FULL, plain[i], off, 8);
        }
        float P3[3] = {0, 0, 0};
        if (acc[7] > 0.f) for (int i = 0; i < 3; i++) P3[i] = acc[6] > NMF_MINVAL ? acc[3 + i] / acc[6] : plain[i] / acc[7];
        float T[3] = {0, 0, 0};
#pragma unroll
        for (int s = 0; s < 2; s++) {
          const float on = sens * con_on(con[s]);
          float rr[3] = {on * (con[s].r[0] + com[0] - P3[0]), on * (con[s].r[1] + com[1] - P3[1]), on * (con[s].r[2] + com[2] - P3[2])}, tt[3];
          cross3(rr, Fc[s], tt); T[0] += tt[0]; T[1] += tt[1]; T[2] += tt[2];
        }
#pragma unroll
        for (int off = 1; off < 8; off <<= 1)
#pragma unroll
          for (int i = 0; i < 3; i++) T[i] += __shfl_xor_sync(NMF_FULL, T[i], off, 8);
        if (is_leg && k == 0) {
          float* o = p.out_sensor + ((size_t)fly * NLEG + grp) * 16;
          o[0] = acc[7];
          for (int i = 0; i < 3; i++) { o[1 + i] = acc[7] > 0.f ? -acc[i] : 0.f; o[4 + i] = acc[7] > 0.f ? -T[i] : 0.f; o[7 + i] = P3[i]; }
          o[10] = acc[7] > 0.f ? 1.f : 0.f; o[11] = 0.f; o[12] = 0.f; o[13] = 0.f; o[14] = acc[7] > 0.f ? 1.f : 0.f; o[15] = 0.f;
        }
      }
      if (p.out_xpos || p.out_xquat) {
        block_sync();   // the u staging is free again: reuse it as the pose exchange buffer
        float* ps = sm + SM_U + tid * 8;
        ps[0] = xpos[0]; ps[1] = xpos[1]; ps[2] = xpos[2]; ps[3] = xq[0]; ps[4] = xq[1]; ps[5] = xq[2]; ps[6] = xq[3];
        block_sync();
        for (int sgi = tid; sgi < p.nseg; sgi += CTA) {
          const float* tb = p.seg_tab + sgi * 8; const float* bp = sm + SM_U + __float_as_int(tb[0]) * 8;
          float lp[3] = {tb[1], tb[2], tb[3]}, lq[4] = {tb[4], tb[5], tb[6], tb[7]}, w[3], wq[4];
          qrot(bp + 3, lp, w); qmul(bp + 3, lq, wq);
          if (p.out_xpos) { float* o = p.out_xpos + ((size_t)fly * p.nseg + sgi) * 3; o[0] = bp[0] + w[0]; o[1] = bp[1] + w[1]; o[2] = bp[2] + w[2]; }
          if (p.out_xquat) { float* o = p.out_xquat + ((size_t)fly * p.nseg + sgi) * 4; o[0] = wq[0]; o[1] = wq[1]; o[2] = wq[2]; o[3] = wq[3]; }
        }
        block_sync();
        for (int i = tid; i < NGROUP * U_STRIDE; i += CTA) sm[SM_U + i] = 0.f;   // restore the staging invariant (hub chains: u = 0)
      }
    }

    // ---- advance: qvel += dt a' ; positions integrate with the NEW velocity ; qacc stays as next warm start
    block_sync();
    for (int i = tid; i < NV; i += CTA) st[S_QVEL + i] += p.dt * sm[SM_X + i];
    block_sync();
    for (int i = tid + 6; i < NV; i += CTA) st[S_QPOS + 1 + i] += p.dt * st[S_QVEL + i];
    if (tid == 0) {
      for (int i = 0; i < 3; i++) st[S_QPOS + i] += p.dt * st[S_QVEL + i];
      float w[3] = {st[S_QVEL + 3], st[S_QVEL + 4], st[S_QVEL + 5]};
      float n = sqrtf(dot3(w, w));
      float q[4] = {st[S_QPOS + 3], st[S_QPOS + 4], st[S_QPOS + 5], st[S_QPOS + 6]};
      if (n > NMF_MINVAL) {
        float sn, cs; sincos_small(0.5f * p.dt * n, &sn, &cs);
        float dq[4] = {cs, w[0] / n * sn, w[1] / n * sn, w[2] / n * sn}, nq[4];
        qmul(q, dq, nq); q[0] = nq[0]; q[1] = nq[1]; q[2] = nq[2]; q[3] = nq[3];
      }
      qnormalize(q);
      st[S_QPOS + 3] = q[0]; st[S_QPOS + 4] = q[1]; st[S_QPOS + 5] = q[2]; st[S_QPOS + 6] = q[3];
      st[S_TIME] += p.dt;
    }
    block_sync();
  }

  // ---- write the record back (TMA bulk store)
  if (!p.forward_only) tma_store_record(p.state + (size_t)fly * S_STRIDE, st, tid, published);
}

}  // namespace nmf
