// nmf_step.cuh — body of the fused step kernel, written in terms of `real`.  Do not include directly: nmf_step_all.cuh defines
// `real` / `NMF_NS` and includes this file once per precision (see nmf_step_common.cuh for the overview).
namespace nmf {
namespace NMF_NS {

typedef StepParamsT<real> SP;


// ------------------------------------------------------------------ small math
__device__ __forceinline__ void qmul(const real* a, const real* b, real* r) {
  real w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  real x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  real y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  real z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
__device__ __forceinline__ void qrot(const real* q, const real* v, real* r) {
  // r = v + 2 w (u x v) + 2 u x (u x v)
  real tx = real(2.) * (q[2] * v[2] - q[3] * v[1]), ty = real(2.) * (q[3] * v[0] - q[1] * v[2]), tz = real(2.) * (q[1] * v[1] - q[2] * v[0]);
  real rx = v[0] + q[0] * tx + (q[2] * tz - q[3] * ty);
  real ry = v[1] + q[0] * ty + (q[3] * tx - q[1] * tz);
  real rz = v[2] + q[0] * tz + (q[1] * ty - q[2] * tx);
  r[0] = rx; r[1] = ry; r[2] = rz;
}
__device__ __forceinline__ void q2mat(const real* q, real* m) {
  real w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = real(1.) - real(2.) * (y * y + z * z); m[1] = real(2.) * (x * y - w * z); m[2] = real(2.) * (x * z + w * y);
  m[3] = real(2.) * (x * y + w * z); m[4] = real(1.) - real(2.) * (x * x + z * z); m[5] = real(2.) * (y * z - w * x);
  m[6] = real(2.) * (x * z - w * y); m[7] = real(2.) * (y * z + w * x); m[8] = real(1.) - real(2.) * (x * x + y * y);
}
__device__ __forceinline__ void qnormalize(real* q) {
  real n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (n2 < real(1e-30)) { q[0] = real(1.); q[1] = q[2] = q[3] = real(0.); return; }
  real s = m_rsqrt(n2);
  q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
}
__device__ __forceinline__ void cross3(const real* a, const real* b, real* r) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
__device__ __forceinline__ real dot3(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ real dot6(const real* a, const real* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
// spatial inertia (Ixx Iyy Izz Ixy Ixz Iyz | hx hy hz | m) times motion vector (ang, lin)
__device__ __forceinline__ void mul_inert(const real* i, const real* v, real* r) {
  real r0 = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  real r1 = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  real r2 = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  real r3 = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  real r4 = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  real r5 = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
  r[0] = r0; r[1] = r1; r[2] = r2; r[3] = r3; r[4] = r4; r[5] = r5;
}
__device__ __forceinline__ void cross_motion(const real* vel, const real* v, real* r) {
  real a[3], b[3], c[3];
  cross3(vel, v, a); cross3(vel, v + 3, b); cross3(vel + 3, v, c);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
__device__ __forceinline__ void cross_force(const real* vel, const real* f, real* r) {
  real a[3], b[3], c[3];
  cross3(vel, f, a); cross3(vel + 3, f + 3, b); cross3(vel, f + 3, c);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
// expand a 10-real spatial inertia into packed symmetric 6x6 (order: wx wy wz vx vy vz)
__device__ __forceinline__ void expand_inert(const real* i, real* P) {
  P[s6(0, 0)] = i[0]; P[s6(0, 1)] = i[3]; P[s6(0, 2)] = i[4]; P[s6(0, 3)] = real(0.);   P[s6(0, 4)] = -i[8]; P[s6(0, 5)] = i[7];
  P[s6(1, 1)] = i[1]; P[s6(1, 2)] = i[5]; P[s6(1, 3)] = i[8];  P[s6(1, 4)] = real(0.);  P[s6(1, 5)] = -i[6];
  P[s6(2, 2)] = i[2]; P[s6(2, 3)] = -i[7]; P[s6(2, 4)] = i[6]; P[s6(2, 5)] = real(0.);
  P[s6(3, 3)] = i[9]; P[s6(3, 4)] = real(0.); P[s6(3, 5)] = real(0.); P[s6(4, 4)] = i[9]; P[s6(4, 5)] = real(0.); P[s6(5, 5)] = i[9];
}
__device__ __forceinline__ void sym6_mul(const real* P, const real* v, real* r) {
#pragma unroll
  for (int a = 0; a < 6; a++) {
    real s = real(0.);
#pragma unroll
    for (int b = 0; b < 6; b++) s += P[a <= b ? s6(a, b) : s6(b, a)] * v[b];
    r[a] = s;
  }
}

// ------------------------------------------------------------------ 8-lane chain scans
template <int N>
__device__ __forceinline__ void chain_prefix(real* v, unsigned mask, int k) {  // inclusive, root -> tip
#pragma unroll
  for (int off = 1; off < 8; off <<= 1) {
#pragma unroll
    for (int n = 0; n < N; n++) { real t = __shfl_up_sync(mask, v[n], off, 8); if (k >= off) v[n] += t; }
  }
}
template <int N>
__device__ __forceinline__ void chain_suffix(real* v, unsigned mask, int k) {  // inclusive, tip -> root
#pragma unroll
  for (int off = 1; off < 8; off <<= 1) {
#pragma unroll
    for (int n = 0; n < N; n++) { real t = __shfl_down_sync(mask, v[n], off, 8); if (k + off < 8) v[n] += t; }
  }
}

// block-wide sum of N values; call from converged code only (all 64 threads)
template <int N>
__device__ __forceinline__ void cta_reduce(real* v, real* s_red, int& parity, int tid, int bar) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int n = 0; n < N; n++) v[n] += __shfl_xor_sync(NMF_FULL, v[n], off);
  }
  real* buf = s_red + parity * 16;
  if ((tid & 31) == 0) {
#pragma unroll
    for (int n = 0; n < N; n++) buf[(tid >> 5) * 8 + n] = v[n];
  }
  block_sync(bar);
#pragma unroll
  for (int n = 0; n < N; n++) v[n] = buf[n] + buf[8 + n];
  parity ^= 1;
}

// ------------------------------------------------------------------ contact slot (kept in registers)
// Branch-free convention: an inactive slot has D = c0 = w = s = 0, so it contributes nothing anywhere.
struct Contact {     // 10 registers per slot; "active" <=> D > 0; signed distance = 2 (r[2] + com_z) (the point sits midway)
  real r[3];        // contact position relative to the subtree COM
  real cx, cy;      // first tangent (cx, cy, 0); second = (-cy, cx, 0); normal = +z
  real D;           // 1/R of the four pyramid rows (0 when the slot is inactive)
  real c0;          // K * imp * (dist - margin)
  real w[3];        // (n, mu t1, mu t2) . (a_p + B v_p)  for the current qacc
};
__device__ __forceinline__ real con_on(const Contact& c) { return c.D > real(0.) ? real(1.) : real(0.); }

// general-exponent branch of the impedance sigmoid: kept out of line (four inlined powf bodies per call site are ~18 KB of
// SASS that the reference's power-2 / power-1 settings never execute, in an instruction-fetch-bound kernel)
__device__ NMF_COLD real impedance_general(real x, real mid, real power) {
  return (x <= mid) ? m_pow(x, power) / m_pow(mid, power - real(1.)) : real(1.) - m_pow(real(1.) - x, power) / m_pow(real(1.) - mid, power - real(1.));
}
__device__ __forceinline__ real impedance_of(const real* solimp, real x_abs) {
  const real d0 = solimp[0], d1 = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  if (d0 == d1 || width <= real(1e-15)) return real(0.5) * (d0 + d1);   // (uniform branch)
  real x = m_min(x_abs / width, real(1.));
  real y;
  if (power == real(1.)) y = x;
  else if (power == real(2.)) y = (x <= mid) ? x * x / mid : real(1.) - (real(1.) - x) * (real(1.) - x) / (real(1.) - mid);
  else y = impedance_general(x, mid, power);
  return d0 + y * (d1 - d0);
}

__device__ __forceinline__ real impedance(const SP& p, real x_abs) { return impedance_of(p.solimp, x_abs); }

// finishes a candidate contact: solver parameters and the B*velocity part of the rows
__device__ __forceinline__ void finish_contact(const SP& p, Contact& c, real active, real dist, const real* pos, real hx, real hy,
                                               const real* com, const real* cvel, real invw) {
  c.r[0] = pos[0] - com[0]; c.r[1] = pos[1] - com[1]; c.r[2] = pos[2] - com[2];
  c.cx = hx; c.cy = hy;
  real imp = impedance(p, m_abs(dist - p.margin));
  real R0 = m_max(real(1e-15), (real(1.) - imp) * invw * (real(1.) + p.mu * p.mu) / imp);
  c.D = active / (real(2.) * (p.mu * p.mu / p.impratio) * R0);
  c.c0 = active * p.cK * imp * (dist - p.margin);
  real vp[3] = {cvel[3] + cvel[1] * c.r[2] - cvel[2] * c.r[1], cvel[4] + cvel[2] * c.r[0] - cvel[0] * c.r[2],
                 cvel[5] + cvel[0] * c.r[1] - cvel[1] * c.r[0]};
  c.w[0] = active * p.cB * vp[2];
  c.w[1] = active * p.cB * p.mu * (c.cx * vp[0] + c.cy * vp[1]);
  c.w[2] = active * p.cB * p.mu * (-c.cy * vp[0] + c.cx * vp[1]);
}

// geom-vs-ground-plane narrow phase for the geom carried by this lane's body
// (plane z = 0, normal +z: reference world.py:251-260).  Fills NSLOT slots: a capsule uses two (its end spheres); a convex
// hull uses one (the support vertex) when NSLOT = 2 and up to four when NSLOT = 4 -- [PRIOR] mjc_PlaneConvex on a mesh with the
// `multiccd` flag (mujoco_globals.yaml:18): the neighbours of the support vertex in the hull's vertex graph that are also
// within the margin become contacts too, in graph order, up to 4 per geom.
template <int NSLOT>
__device__ __forceinline__ real collide(const SP& p, const real* role, int tid, const real* xpos, const real* R,
                                         const real* com, const real* cvel, real invw, Contact* con, int& hullv) {
  const int gtype = role_int(role[RF_GTYPE * CTA + tid]);
  real pos[NSLOT][3], dist[NSLOT], act[NSLOT], hx = real(0.), hy = real(1.);
#pragma unroll
  for (int s = 0; s < NSLOT; s++) { pos[s][0] = pos[s][1] = pos[s][2] = real(0.); dist[s] = real(1.); act[s] = real(0.); }
  {  // capsule: two sphere-plane tests, frame aligned with the capsule axis (evaluated on every lane, masked by type)
    real gp[3] = {role[(RF_GPOS + 0) * CTA + tid], role[(RF_GPOS + 1) * CTA + tid], role[(RF_GPOS + 2) * CTA + tid]};
    real ga[3] = {role[(RF_GAXIS + 0) * CTA + tid], role[(RF_GAXIS + 1) * CTA + tid], role[(RF_GAXIS + 2) * CTA + tid]};
    real rad = role[RF_GRAD * CTA + tid], half = role[RF_GHALF * CTA + tid];
    real c[3], a[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      c[i] = xpos[i] + R[3 * i] * gp[0] + R[3 * i + 1] * gp[1] + R[3 * i + 2] * gp[2];
      a[i] = R[3 * i] * ga[0] + R[3 * i + 1] * ga[1] + R[3 * i + 2] * ga[2];
    }
    real hn2 = a[0] * a[0] + a[1] * a[1];
    real inv = m_rsqrt(m_max(hn2, real(1e-24)));
    const bool cap = gtype == 0;
    if (cap) { hx = hn2 < real(1e-24) ? real(1.) : a[0] * inv; hy = hn2 < real(1e-24) ? real(0.) : a[1] * inv; }
    real e0 = c[2] + half * a[2], e1 = c[2] - half * a[2];
    if (cap) {
      dist[0] = e0 - rad; dist[1] = e1 - rad;
      pos[0][0] = c[0] + half * a[0]; pos[0][1] = c[1] + half * a[1]; pos[0][2] = real(0.5) * dist[0];
      pos[1][0] = c[0] - half * a[0]; pos[1][1] = c[1] - half * a[1]; pos[1][2] = real(0.5) * dist[1];
      act[0] = (e0 <= p.margin + rad) ? real(1.) : real(0.); act[1] = (e1 <= p.margin + rad) ? real(1.) : real(0.);
    }
  }
  if (gtype == 1) {
    // convex hull: deepest vertex = support point along -z.  Steepest-descent walk on the hull's vertex graph, warm-started
    // from the previous step's support vertex (exact: on a convex polytope a vertex with no lower neighbour is the global
    // minimum of a linear function).  Lane-dependent trip count: the warp reconverges below.  The last sweep of the walk has
    // looked at every neighbour of the support vertex, which is where the extra (multiccd) contacts are picked up.
    const int adr = role_int(role[RF_GVADR * CTA + tid]), num = role_int(role[RF_GVNUM * CTA + tid]);
    int bi = hullv < num ? hullv : 0;
    real best = real(3.0e38);
    if (num > 0) { const real* hv = p.hull + 3 * (adr + bi); best = R[6] * __ldg(hv) + R[7] * __ldg(hv + 1) + R[8] * __ldg(hv + 2); }
    int extra[NSLOT > 2 ? NSLOT - 1 : 1], nextra = 0;
    const real zlim = p.margin - xpos[2];       // a vertex with R[2,:] . v <= zlim lies within the margin of the plane
    for (int moved = num > 0; moved;) {
      moved = 0; nextra = 0;
      const int n0 = __ldg(p.hull_nbr_adr + adr + bi), n1 = __ldg(p.hull_nbr_adr + adr + bi + 1);
      int cand = bi;
      for (int e = n0; e < n1; e++) {
        const int v = __ldg(p.hull_nbr + e);
        const real* hv = p.hull + 3 * (adr + v);
        real z = R[6] * __ldg(hv) + R[7] * __ldg(hv + 1) + R[8] * __ldg(hv + 2);
        if (z < best) { best = z; cand = v; moved = 1; }
        if (NSLOT > 2 && z <= zlim && nextra < NSLOT - 1) {
#pragma unroll
          for (int q = 0; q < NSLOT - 1; q++) if (q == nextra) extra[q] = v;
          nextra++;
        }
      }
      bi = cand;
    }
    hullv = bi;
    if (!(NSLOT > 2 && p.multiccd)) nextra = 0;
#pragma unroll
    for (int s = 0; s < (NSLOT > 2 ? NSLOT : 1); s++) {
      const int v = s == 0 ? bi : extra[s - 1];
      const bool on = num > 0 && (s == 0 || s - 1 < nextra);
      const real* hv = p.hull + 3 * (adr + (on ? v : 0));
      real h0 = __ldg(hv), h1 = __ldg(hv + 1), h2 = __ldg(hv + 2);
      const real d = xpos[2] + R[6] * h0 + R[7] * h1 + R[8] * h2;
      dist[s] = on ? d : real(1.);
      pos[s][0] = xpos[0] + R[0] * h0 + R[1] * h1 + R[2] * h2;
      pos[s][1] = xpos[1] + R[3] * h0 + R[4] * h1 + R[5] * h2;
      pos[s][2] = real(0.5) * dist[s];
      act[s] = (on && dist[s] <= p.margin) ? real(1.) : real(0.);
    }
  }
  __syncwarp(NMF_FULL);
  real total = real(0.);
#pragma unroll
  for (int s = 0; s < NSLOT; s++) { finish_contact(p, con[s], act[s], dist[s], pos[s], hx, hy, com, cvel, invw); total += act[s]; }
  return total;   // number of contacts of this geom (adhesion transmission divides by it)
}

// point "acceleration" of a contact for a body spatial vector S (ang, lin), projected on (n, mu t1, mu t2)
__device__ __forceinline__ void project_point(const Contact& c, const real* S, real mu, real* out) {
  real ax = S[3] + S[1] * c.r[2] - S[2] * c.r[1];
  real ay = S[4] + S[2] * c.r[0] - S[0] * c.r[2];
  real az = S[5] + S[0] * c.r[1] - S[1] * c.r[0];
  const real on = con_on(c);
  out[0] = on * az; out[1] = on * mu * (c.cx * ax + c.cy * ay); out[2] = on * mu * (-c.cy * ax + c.cx * ay);
}

// pyramid rows of one contact: jar_r = base +- w1 / w2
__device__ __forceinline__ void rows4(const real* w, real c0, real* jar) {
  real b = w[0] + c0;
  jar[0] = b + w[1]; jar[1] = b - w[1]; jar[2] = b + w[2]; jar[3] = b - w[2];
}

// contact forces for the current jar: accumulates the world wrench about the COM (ang, lin) into Wc and, when WITH_A,
// the contact augmentation A += X' W X (21 packed).  Branch-free (inactive slots have D = 0).
template <bool WITH_A>
__device__ __forceinline__ void contact_forces(const Contact& c, real mu, real* Wc, real* A, real* fn_out) {
  real jar[4]; rows4(c.w, c.c0, jar);
  real a[4], f[4];
#pragma unroll
  for (int r = 0; r < 4; r++) { a[r] = jar[r] < real(0.) ? real(1.) : real(0.); f[r] = -c.D * m_min(jar[r], real(0.)); }
  real fn = f[0] + f[1] + f[2] + f[3], f1 = mu * (f[0] - f[1]), f2 = mu * (f[2] - f[3]);
  real F[3] = {f1 * c.cx - f2 * c.cy, f1 * c.cy + f2 * c.cx, fn};
  real T[3]; cross3(c.r, F, T);
  Wc[0] += T[0]; Wc[1] += T[1]; Wc[2] += T[2]; Wc[3] += F[0]; Wc[4] += F[1]; Wc[5] += F[2];
  if (fn_out) *fn_out = fn;
  if (WITH_A) {
    real s1 = a[0] + a[1], s2 = a[2] + a[3], d1 = a[0] - a[1], d2 = a[2] - a[3], m2 = mu * mu;
    real W[9];
    W[0] = c.D * m2 * (s1 * c.cx * c.cx + s2 * c.cy * c.cy);
    W[4] = c.D * m2 * (s1 * c.cy * c.cy + s2 * c.cx * c.cx);
    W[1] = W[3] = c.D * m2 * (s1 - s2) * c.cx * c.cy;
    W[8] = c.D * (s1 + s2);
    W[2] = W[6] = c.D * mu * (d1 * c.cx - d2 * c.cy);
    W[5] = W[7] = c.D * mu * (d1 * c.cy + d2 * c.cx);
    // T = [r]x W  (columns: r x W[:,j]);  A_ww = T (-[r]x) -> row i: r x T[i,:]
    real Tm[9];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      real col[3] = {W[j], W[3 + j], W[6 + j]}, t[3]; cross3(c.r, col, t);
      Tm[j] = t[0]; Tm[3 + j] = t[1]; Tm[6 + j] = t[2];
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      real row[3] = {Tm[3 * i], Tm[3 * i + 1], Tm[3 * i + 2]}, t[3]; cross3(c.r, row, t);
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
__device__ __forceinline__ void ls_eval(const Con& c, const real* sv, real alpha, real& d0, real& d1, real& nchanged) {
  real jar[4], jv[4]; rows4(c.w, c.c0, jar); rows4(sv, real(0.), jv);
#pragma unroll
  for (int r = 0; r < 4; r++) {
    real x = jar[r] + alpha * jv[r];
    real on = x < real(0.) ? c.D : real(0.);
    d0 += on * x * jv[r]; d1 += on * jv[r] * jv[r];
    nchanged += ((x < real(0.)) != (jar[r] < real(0.))) ? real(1.) : real(0.);   // rows whose state differs from the one the Hessian was built for
  }
}

// ------------------------------------------------------------------ general-frame contact slot (terrain worlds)
// Same conventions as Contact, but the contact normal is arbitrary (box-column terrain has walls and edges):
// frame = (n, t1, t2 = n x t1).  Used by the TERRAIN instantiation of the step only; the flat-ground kernel keeps the
// cheaper z-normal slot above.
struct ContactG {    // 14 registers per slot
  real r[3];        // contact position relative to the subtree COM
  real n[3], t[3];  // normal, first tangent
  real D, c0;
  real w[3];
};
__device__ __forceinline__ real con_on(const ContactG& c) { return c.D > real(0.) ? real(1.) : real(0.); }

__device__ __forceinline__ void finish_contact(const SP& p, ContactG& c, real active, real dist, const real* pos, const real* nrm,
                                               const real* hint, const real* com, const real* cvel, real invw) {
  c.r[0] = pos[0] - com[0]; c.r[1] = pos[1] - com[1]; c.r[2] = pos[2] - com[2];
  c.n[0] = nrm[0]; c.n[1] = nrm[1]; c.n[2] = nrm[2];
  {  // first tangent: the hint (capsule axis) orthogonalised against the normal; world x (or y) when they are parallel
    real hn = dot3(hint, nrm), t[3] = {hint[0] - hn * nrm[0], hint[1] - hn * nrm[1], hint[2] - hn * nrm[2]};
    real t2 = dot3(t, t);
    if (t2 < real(1e-12)) {
      const bool usex = m_abs(nrm[0]) < real(0.9);
      const real e[3] = {usex ? real(1.) : real(0.), usex ? real(0.) : real(1.), real(0.)};
      hn = dot3(e, nrm); t[0] = e[0] - hn * nrm[0]; t[1] = e[1] - hn * nrm[1]; t[2] = e[2] - hn * nrm[2]; t2 = dot3(t, t);
    }
    const real inv = m_rsqrt(t2);
    c.t[0] = t[0] * inv; c.t[1] = t[1] * inv; c.t[2] = t[2] * inv;
  }
  real imp = impedance(p, m_abs(dist - p.margin));
  real R0 = m_max(real(1e-15), (real(1.) - imp) * invw * (real(1.) + p.mu * p.mu) / imp);
  c.D = active / (real(2.) * (p.mu * p.mu / p.impratio) * R0);
  c.c0 = active * p.cK * imp * (dist - p.margin);
  real vp[3] = {cvel[3] + cvel[1] * c.r[2] - cvel[2] * c.r[1], cvel[4] + cvel[2] * c.r[0] - cvel[0] * c.r[2],
                 cvel[5] + cvel[0] * c.r[1] - cvel[1] * c.r[0]};
  real t2v[3]; cross3(c.n, c.t, t2v);
  c.w[0] = active * p.cB * dot3(c.n, vp);
  c.w[1] = active * p.cB * p.mu * dot3(c.t, vp);
  c.w[2] = active * p.cB * p.mu * dot3(t2v, vp);
}

// one sphere (centre c, radius rad) against the terrain solid = floor plane + grid of box columns: the closest point of
// the solid decides normal and distance (ONE contact per sphere).  Column (i, j) covers |x - i Px| <= hx, |y - j Py| <= hy,
// z <= top(i, j), top = terr[4 + ((i + j) & 1)].  A centre inside a column is pushed out through the top face.
__device__ __forceinline__ void sphere_terrain(const SP& p, const real* c, real rad, real* nrm, real& dist) {
  const real Px = p.terr[0], Py = p.terr[1], hx = p.terr[2], hy = p.terr[3];
  nrm[0] = real(0.); nrm[1] = real(0.); nrm[2] = real(1.); dist = c[2] - p.terr[6] - rad;      // floor plane
  const real fi = m_rint(c[0] / Px), fj = m_rint(c[1] / Py);
  const int i0 = (int)fi, j0 = (int)fj;
  const int sx = c[0] >= fi * Px ? 1 : -1, sy = c[1] >= fj * Py ? 1 : -1;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int i = i0 + ((q & 1) ? sx : 0), j = j0 + ((q & 2) ? sy : 0);
    const real cx = (real)i * Px, cy = (real)j * Py, top = ((i + j) & 1) ? p.terr[5] : p.terr[4];
    const real qx = m_min(m_max(c[0], cx - hx), cx + hx), qy = m_min(m_max(c[1], cy - hy), cy + hy), qz = m_min(c[2], top);
    const real dx = c[0] - qx, dy = c[1] - qy, dz = c[2] - qz, d2 = dx * dx + dy * dy + dz * dz;
    real d, n0, n1, n2;
    if (d2 > real(0.)) { const real inv = m_rsqrt(d2); d = d2 * inv - rad; n0 = dx * inv; n1 = dy * inv; n2 = dz * inv; }
    else { d = c[2] - top - rad; n0 = real(0.); n1 = real(0.); n2 = real(1.); }
    if (d < dist) { dist = d; nrm[0] = n0; nrm[1] = n1; nrm[2] = n2; }
  }
}

// capsule-vs-terrain narrow phase for the geom carried by this lane's body: the two end spheres, one contact each
template <int NSLOT>
__device__ __forceinline__ real collide(const SP& p, const real* role, int tid, const real* xpos, const real* R,
                                         const real* com, const real* cvel, real invw, ContactG* con, int&) {
  const int gtype = role_int(role[RF_GTYPE * CTA + tid]);
  real gp[3] = {role[(RF_GPOS + 0) * CTA + tid], role[(RF_GPOS + 1) * CTA + tid], role[(RF_GPOS + 2) * CTA + tid]};
  real ga[3] = {role[(RF_GAXIS + 0) * CTA + tid], role[(RF_GAXIS + 1) * CTA + tid], role[(RF_GAXIS + 2) * CTA + tid]};
  const real rad = role[RF_GRAD * CTA + tid], half = role[RF_GHALF * CTA + tid];
  real c[3], a[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    c[i] = xpos[i] + R[3 * i] * gp[0] + R[3 * i + 1] * gp[1] + R[3 * i + 2] * gp[2];
    a[i] = R[3 * i] * ga[0] + R[3 * i + 1] * ga[1] + R[3 * i + 2] * ga[2];
  }
  real total = real(0.);
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const real sg = s == 0 ? half : -half;
    real e[3] = {c[0] + sg * a[0], c[1] + sg * a[1], c[2] + sg * a[2]}, nrm[3], dist;
    sphere_terrain(p, e, rad, nrm, dist);
    const real act = (gtype == 0 && dist <= p.margin) ? real(1.) : real(0.);
    const real back = rad + real(0.5) * dist;
    real pos[3] = {e[0] - back * nrm[0], e[1] - back * nrm[1], e[2] - back * nrm[2]};
    finish_contact(p, con[s], act, dist, pos, nrm, a, com, cvel, invw);
    total += act;
  }
  return total;
}

__device__ __forceinline__ void project_point(const ContactG& c, const real* S, real mu, real* out) {
  real a[3] = {S[3] + S[1] * c.r[2] - S[2] * c.r[1], S[4] + S[2] * c.r[0] - S[0] * c.r[2], S[5] + S[0] * c.r[1] - S[1] * c.r[0]};
  real t2[3]; cross3(c.n, c.t, t2);
  const real on = con_on(c);
  out[0] = on * dot3(c.n, a); out[1] = on * mu * dot3(c.t, a); out[2] = on * mu * dot3(t2, a);
}

template <bool WITH_A>
__device__ __forceinline__ void contact_forces(const ContactG& c, real mu, real* Wc, real* A, real* fn_out) {
  real jar[4]; rows4(c.w, c.c0, jar);
  real a[4], f[4];
#pragma unroll
  for (int r = 0; r < 4; r++) { a[r] = jar[r] < real(0.) ? real(1.) : real(0.); f[r] = -c.D * m_min(jar[r], real(0.)); }
  real fn = f[0] + f[1] + f[2] + f[3], f1 = mu * (f[0] - f[1]), f2 = mu * (f[2] - f[3]);
  real t2[3]; cross3(c.n, c.t, t2);
  real F[3] = {fn * c.n[0] + f1 * c.t[0] + f2 * t2[0], fn * c.n[1] + f1 * c.t[1] + f2 * t2[1], fn * c.n[2] + f1 * c.t[2] + f2 * t2[2]};
  real T[3]; cross3(c.r, F, T);
  Wc[0] += T[0]; Wc[1] += T[1]; Wc[2] += T[2]; Wc[3] += F[0]; Wc[4] += F[1]; Wc[5] += F[2];
  if (fn_out) *fn_out = fn;
  if (WITH_A) {
    // W = D sum_r a_r d_r d_r',  d = n +- mu t1 | n +- mu t2
    const real s1 = a[0] + a[1], s2 = a[2] + a[3], d1 = mu * (a[0] - a[1]), d2 = mu * (a[2] - a[3]), m2 = mu * mu;
    real u[3] = {d1 * c.t[0] + d2 * t2[0], d1 * c.t[1] + d2 * t2[1], d1 * c.t[2] + d2 * t2[2]};   // n u' + u n' carries the cross terms
    real W[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = i; j < 3; j++) {
        real v = (s1 + s2) * c.n[i] * c.n[j] + c.n[i] * u[j] + u[i] * c.n[j] + m2 * (s1 * c.t[i] * c.t[j] + s2 * t2[i] * t2[j]);
        W[3 * i + j] = W[3 * j + i] = c.D * v;
      }
    real Tm[9];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      real col[3] = {W[j], W[3 + j], W[6 + j]}, t[3]; cross3(c.r, col, t);
      Tm[j] = t[0]; Tm[3 + j] = t[1]; Tm[6 + j] = t[2];
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      real row[3] = {Tm[3 * i], Tm[3 * i + 1], Tm[3 * i + 2]}, t[3]; cross3(c.r, row, t);
#pragma unroll
      for (int j = i; j < 3; j++) A[s6(i, j)] += t[j];
#pragma unroll
      for (int j = 0; j < 3; j++) A[s6(i, 3 + j)] += Tm[3 * i + j];
    }
    A[s6(3, 3)] += W[0]; A[s6(3, 4)] += W[1]; A[s6(3, 5)] += W[2]; A[s6(4, 4)] += W[4]; A[s6(4, 5)] += W[5]; A[s6(5, 5)] += W[8];
  }
}

// adhesion pull f (>= 0 towards the surface) of one contact: wrench of -f n at the contact point
__device__ __forceinline__ void adhesion_wrench(const ContactG& c, real f, real* W) {
  const real s = -con_on(c) * f;
  real F[3] = {s * c.n[0], s * c.n[1], s * c.n[2]}, T[3]; cross3(c.r, F, T);
  W[0] += T[0]; W[1] += T[1]; W[2] += T[2]; W[3] += F[0]; W[4] += F[1]; W[5] += F[2];
}
// signed distance of an active slot (debug dump only)
__device__ __forceinline__ real con_dist(const Contact& c, const real* com) { return real(2.) * (c.r[2] + com[2]); }
__device__ __forceinline__ real con_dist(const ContactG&, const real*) { return real(0.); }

// ------------------------------------------------------------------ contact basis and explicit contact forces (noslip)
// Basis of a contact's force space: e0 = n, e1 = mu t1, e2 = mu t2 (what project_point projects on).  The four pyramid rows are
// e0 +- e1, e0 +- e2, so row forces f map to basis forces G = (f0+f1+f2+f3, f0-f1, f2-f3) and the contact force is sum G_k e_k.
__device__ __forceinline__ void contact_normal(const Contact&, real* n) { n[0] = real(0.); n[1] = real(0.); n[2] = real(1.); }
__device__ __forceinline__ void contact_normal(const ContactG& c, real* n) { n[0] = c.n[0]; n[1] = c.n[1]; n[2] = c.n[2]; }
__device__ __forceinline__ void contact_tangent(const Contact& c, int k, real mu, real* d) {
  d[0] = mu * (k == 0 ? c.cx : -c.cy); d[1] = mu * (k == 0 ? c.cy : c.cx); d[2] = real(0.);
}
__device__ __forceinline__ void contact_tangent(const ContactG& c, int k, real mu, real* d) {
  real t2[3]; cross3(c.n, c.t, t2);
#pragma unroll
  for (int i = 0; i < 3; i++) d[i] = mu * (k == 0 ? c.t[i] : t2[i]);
}
// Newton-form row forces of a slot in basis coordinates, the two pair sums (the normal force each pair of opposing edges
// carries; noslip keeps them) and 1/2 sum f_r^2 R_r
template <class Con>
__device__ __forceinline__ void basis_forces(const Con& c, real* G, real* lim, real& cost_r) {
  real jar[4]; rows4(c.w, c.c0, jar);
  real f[4];
#pragma unroll
  for (int r = 0; r < 4; r++) f[r] = -c.D * m_min(jar[r], real(0.));
  G[0] = f[0] + f[1] + f[2] + f[3]; G[1] = f[0] - f[1]; G[2] = f[2] - f[3];
  lim[0] = f[0] + f[1]; lim[1] = f[2] + f[3];
  cost_r = c.D > real(0.) ? real(0.5) * (f[0] * f[0] + f[1] * f[1] + f[2] * f[2] + f[3] * f[3]) / c.D : real(0.);
}
// wrench about the COM (ang, lin) of the basis force G at the contact point, accumulated into W
template <class Con>
__device__ __forceinline__ void basis_wrench(const Con& c, const real* G, real mu, real* W, real* Fout) {
  real n[3], d1[3], d2[3]; contact_normal(c, n); contact_tangent(c, 0, mu, d1); contact_tangent(c, 1, mu, d2);
  real F[3] = {G[0] * n[0] + G[1] * d1[0] + G[2] * d2[0], G[0] * n[1] + G[1] * d1[1] + G[2] * d2[1], G[0] * n[2] + G[1] * d1[2] + G[2] * d2[2]};
  real T[3]; cross3(c.r, F, T);
  W[0] += T[0]; W[1] += T[1]; W[2] += T[2]; W[3] += F[0]; W[4] += F[1]; W[5] += F[2];
  if (Fout) { Fout[0] = F[0]; Fout[1] = F[1]; Fout[2] = F[2]; }
}
constexpr int NS_MAXC = 48;             // contacts the noslip pass handles (more: the pass is skipped and ST_NOSLIP_SKIP is raised); a mesh-hull fly dropped
                                        // flat on the ground has 48 with multiccd.  The region (B alone is 96 x 96) puts the noslip kernels into dynamic shared memory
constexpr int NS_LD = 2 * NS_MAXC;      // two friction dimensions per contact

// ------------------------------------------------------------------ weld equality of the TetheredWorld (hub lane only)
// Six always-active rows on the free body: rows 0-2 = world position of the hub-frame point weld_a, rows 3-5 =
// torquescale * vec(q_hub * weld_q) with Jacobian G w = 0.5 ts vec((0, w) * q) (w = world angular velocity).  They only touch
// the hub's 6x6 block, so they enter exactly like a contact on the hub: a wrench in the gradient and an X'WX augmentation
// of the hub's spatial inertia.  State lives in shared memory (one lane uses it): r[3] G[9] D[6] c0[6] w[6] sv[6].
constexpr int WL_R = 0, WL_G = 3, WL_D = 12, WL_C0 = 18, WL_W = 24, WL_SV = 30, WL_F = 36, WL_COUNT = 42;   // WL_F: explicit row forces after noslip

__device__ __forceinline__ void point_and_rot(const real* sw, const real* S, real* out) {   // J * (spatial vector of the hub)
  const real* r = sw + WL_R; const real* G = sw + WL_G;
  out[0] = S[3] + S[1] * r[2] - S[2] * r[1]; out[1] = S[4] + S[2] * r[0] - S[0] * r[2]; out[2] = S[5] + S[0] * r[1] - S[1] * r[0];
#pragma unroll
  for (int i = 0; i < 3; i++) out[3 + i] = G[3 * i] * S[0] + G[3 * i + 1] * S[1] + G[3 * i + 2] * S[2];
}
template <class PT>      // StepParamsT or TreeParamsT: both carry the weld_* fields
__device__ __forceinline__ void weld_setup(const PT& p, real* sw, const real* qh, const real* xh, const real* com,
                                           const real* cvel, const real* Sa) {
  real a[3], q[4], pos[6];
  qrot(qh, p.weld_a, a); qmul(qh, p.weld_q, q);
  const real h = real(0.5) * p.weld_ts;
#pragma unroll
  for (int i = 0; i < 3; i++) { pos[i] = xh[i] + a[i]; sw[WL_R + i] = pos[i] - com[i]; pos[3 + i] = p.weld_ts * q[1 + i]; }
  real* G = sw + WL_G;
  G[0] = h * q[0]; G[1] = h * q[3]; G[2] = -h * q[2];
  G[3] = -h * q[3]; G[4] = h * q[0]; G[5] = h * q[1];
  G[6] = h * q[2]; G[7] = -h * q[1]; G[8] = h * q[0];
  real jv[6], ja[6];
  point_and_rot(sw, cvel, jv); point_and_rot(sw, Sa, ja);
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const real imp = impedance_of(p.weld_imp, m_abs(pos[i]));
    const real Rr = m_max(real(1e-15), (real(1.) - imp) * p.weld_invw[i >= 3 ? 1 : 0] / imp);
    sw[WL_D + i] = real(1.) / Rr;
    sw[WL_C0 + i] = p.weld_K * imp * pos[i];
    sw[WL_W + i] = p.weld_B * jv[i] + ja[i];
  }
}
// forces of the six rows for the current acceleration: wrench about the COM into Wc, Hessian augmentation into A
__device__ __forceinline__ void weld_forces(const real* sw, real* Wc, real* A, bool explicit_f = false) {
  const real* r = sw + WL_R; const real* G = sw + WL_G; const real* D = sw + WL_D;
  real f[6];
#pragma unroll
  for (int i = 0; i < 6; i++) f[i] = explicit_f ? sw[WL_F + i] : -D[i] * (sw[WL_W + i] + sw[WL_C0 + i]);
  real T[3]; cross3(r, f, T);
#pragma unroll
  for (int i = 0; i < 3; i++) { Wc[i] += T[i] + G[i] * f[3] + G[3 + i] * f[4] + G[6 + i] * f[5]; Wc[3 + i] += f[i]; }
  // position rows: W = diag(D0, D1, D2) at the point r  ->  X' W X ; rotation rows: G' diag(D3..5) G on the angular block
  real Tm[9];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    real col[3] = {j == 0 ? D[0] : real(0.), j == 1 ? D[1] : real(0.), j == 2 ? D[2] : real(0.)}, t[3]; cross3(r, col, t);
    Tm[j] = t[0]; Tm[3 + j] = t[1]; Tm[6 + j] = t[2];
  }
#pragma unroll
  for (int i = 0; i < 3; i++) {
    real row[3] = {Tm[3 * i], Tm[3 * i + 1], Tm[3 * i + 2]}, t[3]; cross3(r, row, t);
#pragma unroll
    for (int j = i; j < 3; j++) A[s6(i, j)] += t[j] + G[i] * D[3] * G[j] + G[3 + i] * D[4] * G[3 + j] + G[6 + i] * D[5] * G[6 + j];
#pragma unroll
    for (int j = 0; j < 3; j++) A[s6(i, 3 + j)] += Tm[3 * i + j];
  }
  A[s6(3, 3)] += D[0]; A[s6(4, 4)] += D[1]; A[s6(5, 5)] += D[2];
}
// line-search sums of the (always active) rows at step alpha
__device__ __forceinline__ void weld_ls(const real* sw, real alpha, real& d0, real& d1) {
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const real jv = sw[WL_SV + i], x = sw[WL_W + i] + sw[WL_C0 + i] + alpha * jv;
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
constexpr int SM_WORK = SM_MBAR + 4;                // 4 words of per-fly scratch flags (non-finite detector)
constexpr int SM_STAGE = SM_WORK + 4;               // f64 only: the float32 record as it travels (S_STRIDE floats)
constexpr int SM_WELD = SM_STAGE + (sizeof(real) == 8 ? S_STRIDE / 2 : 0);   // weld rows of the tethered world (WL_COUNT)
constexpr int SM_TOTAL = SM_WELD + WL_COUNT;
// noslip instantiations only (appended after the slot's regular region): B (NS_LD x NS_LD), current / Newton friction forces,
// pair sums, tangential residuals; the ranking keys of the contacts alias B before it is filled
constexpr int NS_B = 0, NS_G = NS_B + NS_LD * NS_LD, NS_G0 = NS_G + NS_LD, NS_LIM = NS_G0 + NS_LD, NS_JT = NS_LIM + NS_LD, NS_COUNT = NS_JT + NS_LD + 8;
constexpr int HU_CVEL = 0, HU_CACC = 6;
constexpr int HB_S = 0, HB_XB = 21, HB_SH = 27, HB_TOT = 33, HB_SR = 72;   // TOT: 37 root totals; SR: assembled Schur block (21) + rhs (6)

// Lane-constant description of the two matrix columns (of 16) + the shared last one a lane holds.
struct Cols {
  const real *cd0, *cd1, *cd10;   // cdof of column t, column 8+t (leg dof t+2), leg dof 10
  real add0, add1, add10;         // diagonal additions (armature [+ dt damping]) of those DoFs
};

// H columns of this lane from the staged u_i = P_i cdof_i :  H[i][c] = cdof_c . u_i   (c <= 6 + i)
__device__ __forceinline__ void load_columns(const real* su, const Cols& cl, int t, real* hk0, real* hk1, real& d10) {
  real c0[6], c1[6];
#pragma unroll
  for (int i = 0; i < 6; i++) { c0[i] = cl.cd0[i]; c1[i] = cl.cd1[i]; }
#pragma unroll
  for (int i = 0; i < NLEGDOF; i++) {
    const real* u = su + 8 * i;
    real u6[6] = {u[0], u[1], u[2], u[3], u[4], u[5]};
    real v0 = dot6(c0, u6), v1 = dot6(c1, u6);
    hk0[i] = (t <= 6 + i) ? v0 : real(0.);
    hk1[i] = (t + 2 <= i) ? v1 : real(0.);
    if (6 + i == t) hk0[i] += cl.add0;
    if (i == t + 2) hk1[i] += cl.add1;
  }
  d10 = dot6(cl.cd10, su + 8 * 10) + cl.add10;
}

// L'DL of the 11x11 chain block + its 11x6 border, one matrix column per lane (columns t and 8+t of 16; the last
// diagonal entry d10 is held by every lane).  Leaves L (unit lower, scaled rows) in hk0/hk1, the inverse pivots of
// the lane's own DoFs in i0own/i1own/i10 and this chain's Schur contribution to the hub block in contrib[3].
__device__ __forceinline__ void chain_factor(real* hk0, real* hk1, real d10, int t, const int* pb, const int* pc, real& i0own, real& i1own,
                                             real& i10, real* contrib) {
  contrib[0] = contrib[1] = contrib[2] = real(0.);
#pragma unroll
  for (int kk = NLEGDOF - 1; kk >= 0; kk--) {
    real dk;
    if (kk == 10) dk = d10; else { const int c = 6 + kk; dk = __shfl_sync(NMF_FULL, c < 8 ? hk0[kk] : hk1[kk], c & 7, 8); }
    real ik = real(1.0) / dk;
    if (kk == 10) i10 = ik;
    if (6 + kk == t) i0own = ik;
    if (kk == t + 2) i1own = ik;
    real l0 = hk0[kk] * ik, l1 = hk1[kk] * ik;
#pragma unroll
    for (int s = 0; s < 3; s++) {
      real lb = __shfl_sync(NMF_FULL, l0, pb[s], 8), hc = __shfl_sync(NMF_FULL, hk0[kk], pc[s], 8);
      contrib[s] += lb * hc;
    }
#pragma unroll
    for (int j = 0; j < kk; j++) {
      const int cj = 6 + j;
      real l = __shfl_sync(NMF_FULL, cj < 8 ? l0 : l1, cj & 7, 8);
      hk0[j] -= (t <= cj ? l : real(0.)) * hk0[kk];
      hk1[j] -= (t + 2 <= j ? l : real(0.)) * hk1[kk];
    }
    hk0[kk] = l0; hk1[kk] = l1;
  }
}
// x <- L^-T x on the chain; lanes t < 6 return (in x0) minus the chain's contribution to the hub right-hand side
__device__ __forceinline__ void chain_solve_up(const real* hk0, const real* hk1, int t, real& x0, real& x1, real x10) {
#pragma unroll
  for (int kk = NLEGDOF - 1; kk >= 0; kk--) {
    real xk;
    if (kk == 10) xk = x10; else { const int c = 6 + kk; xk = __shfl_sync(NMF_FULL, c < 8 ? x0 : x1, c & 7, 8); }
    x0 -= (t < 6 + kk ? hk0[kk] : real(0.)) * xk;
    x1 -= (t + 2 < kk ? hk1[kk] : real(0.)) * xk;
  }
}
// x <- L^-1 D^-1 x given the hub solution xb (lanes t < 6)
__device__ __forceinline__ void chain_solve_down(const real* hk0, const real* hk1, int t, real xb, real i0own, real i1own,
                                                 real i10, real& x0, real& x1, real& x10) {
  x0 = t >= 6 ? x0 * i0own : x0;
  x1 *= i1own; x10 *= i10;
#pragma unroll
  for (int kk = 0; kk < NLEGDOF; kk++) {
    real part = (t < 6 ? hk0[kk] * xb : (t - 6 < kk ? hk0[kk] * x0 : real(0.))) + (t + 2 < kk ? hk1[kk] * x1 : real(0.));
    part += __shfl_xor_sync(NMF_FULL, part, 1, 8); part += __shfl_xor_sync(NMF_FULL, part, 2, 8); part += __shfl_xor_sync(NMF_FULL, part, 4, 8);
    x0 = (6 + kk == t) ? x0 - part : x0;
    x1 = (kk == t + 2) ? x1 - part : x1;
    if (kk == 10) x10 -= part;
  }
}
// sum of entries [lo, lo+n) of the 8 chain-root records, computed cooperatively by the 16 hub lanes into SM_HBB + HB_TOT
__device__ __forceinline__ void hub_root_totals(real* sm, int hl, int lo, int n) {
  for (int i = hl; i < n; i += NHUBLANE) {
    real s = real(0.);
#pragma unroll
    for (int g = 0; g < NGROUP; g++) s += sm[SM_ROOT + g * ROOT_STRIDE + lo + i];
    sm[SM_HBB + HB_TOT + i] = s;
  }
}
// hub 6x6 block: S = Hbb - sum(chain contributions); solve S xb = rhs (dense L'DL, serial, one lane)
__device__ __forceinline__ void hub_solve(real* sm, real* xb) {
  real S[21];
#pragma unroll
  for (int i = 0; i < 21; i++) S[i] = sm[SM_HBB + HB_SR + i];
#pragma unroll
  for (int b = 0; b < 6; b++) xb[b] = sm[SM_HBB + HB_SR + 21 + b];
  real dinv[6];
#pragma unroll
  for (int kk = 5; kk >= 0; kk--) {
    dinv[kk] = real(1.0) / S[kk * (kk + 1) / 2 + kk];
#pragma unroll
    for (int j = 0; j < kk; j++) {
      real l = S[kk * (kk + 1) / 2 + j] * dinv[kk];
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
__device__ __forceinline__ void arrowhead_solve(real* sm, const real* s_cdof, const Cols& cl, int grp, int t, bool is_leg, int hl, int lbase,
                                                const int* pb, const int* pc, float* dbg_rows, int bar) {
  real hk0[NLEGDOF], hk1[NLEGDOF], d10, i0own = real(0.), i1own = real(0.), i10 = real(0.), contrib[3];
  load_columns(sm + SM_U + grp * U_STRIDE, cl, t, hk0, hk1, d10);
  if (dbg_rows) {
#pragma unroll
    for (int i = 0; i < NLEGDOF; i++) { dbg_rows[i * 16 + t] = hk0[i]; dbg_rows[i * 16 + 8 + t] = hk1[i]; }
    dbg_rows[176] = d10;
  }
  chain_factor(hk0, hk1, d10, t, pb, pc, i0own, i1own, i10, contrib);
#pragma unroll
  for (int s = 0; s < 3; s++) if (is_leg && t + 8 * s < 21) sm[SM_BASE + grp * 21 + t + 8 * s] = contrib[s];
  real x0 = (is_leg && t >= 6) ? -sm[SM_GRAD + lbase + t - 6] : real(0.), x1 = is_leg ? -sm[SM_GRAD + lbase + t + 2] : real(0.),
        x10 = is_leg ? -sm[SM_GRAD + lbase + 10] : real(0.);
  chain_solve_up(hk0, hk1, t, x0, x1, x10);
  if (is_leg && t < 6) sm[SM_BASE + NLEG * 21 + grp * 6 + t] = x0;
  block_sync(bar);
  if (!is_leg) {   // Schur block and hub right-hand side assembled by the 16 hub lanes, then solved by lane 48
    for (int i = hl; i < 27; i += NHUBLANE) {
      real v;
      if (i < 21) { v = sm[SM_HBB + HB_S + i]; for (int l = 0; l < NLEG; l++) v -= sm[SM_BASE + l * 21 + i]; }
      else { const int b = i - 21; v = -sm[SM_GRAD + b]; for (int l = 0; l < NLEG; l++) v += sm[SM_BASE + NLEG * 21 + l * 6 + b]; }
      sm[SM_HBB + HB_SR + i] = v;
    }
  }
  __syncwarp(NMF_FULL);
  if (!is_leg && hl == 0) {
    real xb[6]; hub_solve(sm, xb);
    real Sh[6] = {0, 0, 0, 0, 0, 0};
    for (int b = 0; b < 6; b++) {
      sm[SM_HBB + HB_XB + b] = xb[b]; sm[SM_X + b] = xb[b];
      const real* cd = s_cdof + CDS * b; for (int i = 0; i < 6; i++) Sh[i] += cd[i] * xb[b];
    }
    for (int i = 0; i < 6; i++) sm[SM_HBB + HB_SH + i] = Sh[i];
  }
  block_sync(bar);
  chain_solve_down(hk0, hk1, t, t < 6 ? sm[SM_HBB + HB_XB + t] : real(0.), i0own, i1own, i10, x0, x1, x10);
  real* sx = sm + SM_X + lbase;
  if (is_leg && t >= 6) sx[t - 6] = x0;
  if (is_leg) sx[t + 2] = x1;
  if (is_leg && t == 0) sx[10] = x10;
}

// ------------------------------------------------------------------ state record <-> shared memory
// The record is float32 in HBM and moves with one TMA bulk copy (nmf_step_common.cuh).  The f32 instantiation copies straight
// into / out of its working state; the f64 one stages the floats next to it and widens / narrows them.
__device__ __forceinline__ void load_record(const SP& p, real* st, real* sm, int fly, int tid, int bar) {
  const float* src_gmem = p.state + (size_t)fly * S_STRIDE;
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(sm + SM_MBAR);
  if (std::is_same<real, float>::value) {
    tma_load_f32(reinterpret_cast<float*>(st), src_gmem, mbar, tid, bar);
  } else {
    float* stage = reinterpret_cast<float*>(sm + SM_STAGE);
    tma_load_f32(stage, src_gmem, mbar, tid, bar);
    block_sync(bar);
    if (p.state64) {
      // full-precision records persist between launches.  An entry whose float32 image still equals what the last f64 launch
      // wrote is taken from the double record; anything else was edited through the API (reset, setters, direct writes to the
      // state tensor) and is taken from the float record.
      const double* s64 = p.state64 + (size_t)fly * S_STRIDE; const float* sh = p.shadow + (size_t)fly * S_STRIDE;
      for (int i = tid; i < S_STRIDE; i += CTA) st[i] = (stage[i] == sh[i]) ? (real)s64[i] : (real)stage[i];
    } else {
      for (int i = tid; i < S_STRIDE; i += CTA) st[i] = (real)stage[i];
    }
  }
}
__device__ __forceinline__ void store_record(const SP& p, const real* st, real* sm, int fly, int tid, bool published, int bar) {
  float* dst_gmem = p.state + (size_t)fly * S_STRIDE;
  if (std::is_same<real, float>::value) {
    tma_store_f32(dst_gmem, reinterpret_cast<const float*>(st), tid, published, bar);
  } else {
    float* stage = reinterpret_cast<float*>(sm + SM_STAGE);
    block_sync(bar);
    for (int i = tid; i < S_STRIDE; i += CTA) {
      stage[i] = (float)st[i];
      if (p.state64) { p.state64[(size_t)fly * S_STRIDE + i] = (double)st[i]; p.shadow[(size_t)fly * S_STRIDE + i] = stage[i]; }
    }
    tma_store_f32(dst_gmem, stage, tid, published, bar);
  }
}

// ------------------------------------------------------------------ the step
// Advances fly `fly` by steps [step0, step0 + nsub) of the launch's p.nsteps (one work item of the launch: the whole
// launch when flies map 1:1 to blocks, a sub-chunk under work-queue scheduling).
// WORLD selects the kernel instantiation: W_FLAT = the reference's FlatGroundWorld; W_TERRAIN = general-frame contact
// slots + capsule-vs-box-column narrow phase; W_TETHER = TetheredWorld (no ground, weld equality on the hub).
// W_MESH = W_FLAT with four contact slots per lane: convex-hull geoms with the `multiccd` flag (up to 4 plane-hull contacts).
constexpr int W_FLAT = 0, W_TERRAIN = 1, W_TETHER = 2, W_MESH = 3;
// FPB > 1: the block steps FPB flies side by side (64 threads and a private shared-memory region each; `fly` < 0 = an empty
// slot).  The kernel is bound by instruction fetch, not by issue slots: the flies of a block meet at a block-wide barrier at the
// top of every solver pass, so that its warps stream the same stretch of code at the same time and share the fetches (L0 /
// L1.5 instruction caches); a fly that needs fewer Newton iterations than its neighbours waits for them there.  Everything else
// synchronises over the fly's own named barrier.  Results do not depend on FPB (bit-identical records).  Measured on B200
// (profiles/fpb_sweep_r02.txt, 4096 flies): 21.5 / 24.4 / 25.3 M env-steps/s at FPB = 1 / 4 / 8; letting the flies of a block
// drift apart by whole stages instead of waiting (no idle slots, but two code streams per block) was slower: 21.7 M;
// one or two more alignment barriers inside a pass changed nothing (25.4 / 24.8 M at FPB = 4 / 8).
template <int WORLD, int FPB = 1, bool NOSLIP = false>
__device__ __forceinline__ void step_block(const SP& p, real* sm, const int fly, const int step0, const int nsub, const bool published) {
  using Con = typename std::conditional<WORLD == W_TERRAIN, ContactG, Contact>::type;
  constexpr bool TETHER = WORLD == W_TETHER;
  constexpr int NSLOT = WORLD == W_MESH ? 4 : 2;      // contact slots per lane
  real* sw = sm + SM_WELD;
  const int tid = fly_tid<FPB>();
  const int bar = FPB == 1 ? 0 : 1 + fly_slot<FPB>();
  if (FPB > 1 && fly < 0) {   // empty slot: only keeps the block-wide pass barriers of the other flies company
    for (int step = 0; step < nsub; step++) while (__syncthreads_or(0)) {}
    return;
  }
  const bool weld_lane = TETHER && tid == NLEG * NLINK;
  const int grp = tid >> 3, k = tid & 7, t = k;
  const bool is_leg = grp < NLEG;
  const int hl = tid - NLEG * NLINK;          // hub lane index (valid when !is_leg)
  const bool hubdof = !is_leg && hl < 6;      // hub lanes 0..5 own the six free-joint DoFs
  const real* role = p.role;
  real* st = sm + SM_STATE;
  real* s_cdof = sm + SM_CDOF;
  real* s_hub = sm + SM_HUB;
  real* s_red = sm + SM_RED;
  real* su = sm + SM_U + grp * U_STRIDE;
  real* rt = sm + SM_ROOT + grp * ROOT_STRIDE;
  int parity = 0;

  // ---- load the state record (one TMA bulk copy of 1216 B), clear the u staging (hub chains keep u = 0)
  for (int i = tid; i < NGROUP * U_STRIDE; i += CTA) sm[SM_U + i] = real(0.);
  load_record(p, st, sm, fly, tid, bar);
  block_sync(bar);

  // per-lane constants that stay in registers for the whole launch
  const int ndof = role_int(role[RF_NDOF * CTA + tid]);                 // 0 on hub lanes
  const int dof0 = is_leg ? role_int(role[RF_DOF0 * CTA + tid]) : (hl < 6 ? hl : 0);
  const int ldof0 = is_leg ? dof0 - 6 - NLEGDOF * grp : 0;                    // first dof index inside the leg
  const int lbase = is_leg ? 6 + NLEGDOF * grp : 6;                           // first global dof of this chain
  const real mass = role[RF_MASS * CTA + tid];
  const real invw = role[RF_INVW * CTA + tid];
  real armv[3], msk[3];          // per own-dof constants; msk[j] = 1 if the lane owns a j-th dof
  int dj[3];                                // global dof index of own dof j (clamped to a valid one when masked)
#pragma unroll
  for (int j = 0; j < 3; j++) {
    armv[j] = role[(RF_ARM + j) * CTA + tid];
    msk[j] = j < ndof ? real(1.) : real(0.); dj[j] = j < ndof ? dof0 + j : dof0;
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
      block_sync(bar);
    }

    // =====================================================================
    // A. kinematics: scan of rigid transforms along each chain
    // =====================================================================
    real qh[4] = {st[S_QPOS + 3], st[S_QPOS + 4], st[S_QPOS + 5], st[S_QPOS + 6]};
    qnormalize(qh);
    const real xh[3] = {st[S_QPOS], st[S_QPOS + 1], st[S_QPOS + 2]};
    real xpos[3], xq[4], R[9], laxis[9];   // laxis: hinge axes in the parent frame, later world
    {
      real q[4] = {role[(RF_BQUAT + 0) * CTA + tid], role[(RF_BQUAT + 1) * CTA + tid], role[(RF_BQUAT + 2) * CTA + tid], role[(RF_BQUAT + 3) * CTA + tid]};
      real pp[3] = {role[(RF_BPOS + 0) * CTA + tid], role[(RF_BPOS + 1) * CTA + tid], role[(RF_BPOS + 2) * CTA + tid]};
#pragma unroll
      for (int j = 0; j < 3; j++) {
        real ax[3] = {role[(RF_AXIS + 3 * j) * CTA + tid], role[(RF_AXIS + 3 * j + 1) * CTA + tid], role[(RF_AXIS + 3 * j + 2) * CTA + tid]};
        qrot(q, ax, laxis + 3 * j);
        real ang = msk[j] * st[S_QPOS + 1 + dj[j]], sn, cs; sincos_small(real(0.5) * ang, &sn, &cs);   // masked dof: identity rotation
        real ql[4] = {cs, ax[0] * sn, ax[1] * sn, ax[2] * sn}, nq[4];
        qmul(q, ql, nq); q[0] = nq[0]; q[1] = nq[1]; q[2] = nq[2]; q[3] = nq[3];
      }
      {  // seed the chain root with the hub pose
        real t3[3], nq[4]; qrot(qh, pp, t3); qmul(qh, q, nq);
        const bool root = k == 0;
#pragma unroll
        for (int i = 0; i < 3; i++) pp[i] = root ? xh[i] + t3[i] : pp[i];
#pragma unroll
        for (int i = 0; i < 4; i++) q[i] = root ? nq[i] : q[i];
      }
#pragma unroll
      for (int off = 1; off < 8; off <<= 1) {
        real uq[4], up[3], t3[3], nq[4];
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
        real qpar[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { real pq = __shfl_up_sync(NMF_FULL, q[i], 1, 8); qpar[i] = k > 0 ? pq : qh[i]; }
#pragma unroll
        for (int j = 0; j < 3; j++) { real w[3]; qrot(qpar, laxis + 3 * j, w); laxis[3 * j] = w[0]; laxis[3 * j + 1] = w[1]; laxis[3 * j + 2] = w[2]; }
      }
      xpos[0] = pp[0]; xpos[1] = pp[1]; xpos[2] = pp[2]; xq[0] = q[0]; xq[1] = q[1]; xq[2] = q[2]; xq[3] = q[3];
    }
    q2mat(xq, R);

    // ---- subtree COM (block reduction), inertial quantities about it
    real xipos[3];
    {
      real ip[3] = {role[(RF_IPOS + 0) * CTA + tid], role[(RF_IPOS + 1) * CTA + tid], role[(RF_IPOS + 2) * CTA + tid]};
#pragma unroll
      for (int i = 0; i < 3; i++) xipos[i] = xpos[i] + R[3 * i] * ip[0] + R[3 * i + 1] * ip[1] + R[3 * i + 2] * ip[2];
    }
    real com[3];
    {
      real v[3] = {mass * xipos[0], mass * xipos[1], mass * xipos[2]};
      cta_reduce<3>(v, s_red, parity, tid, bar);
      com[0] = v[0] * p.inv_total_mass; com[1] = v[1] * p.inv_total_mass; com[2] = v[2] * p.inv_total_mass;
    }
    real cinert[10];
    {
      real ib[6];
#pragma unroll
      for (int i = 0; i < 6; i++) ib[i] = role[(RF_IB + i) * CTA + tid];
      real Ib[9] = {ib[0], ib[3], ib[4], ib[3], ib[1], ib[5], ib[4], ib[5], ib[2]}, T[9], G[9];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) T[3 * i + j] = R[3 * i] * Ib[j] + R[3 * i + 1] * Ib[3 + j] + R[3 * i + 2] * Ib[6 + j];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = i; j < 3; j++) G[3 * i + j] = T[3 * i] * R[3 * j] + T[3 * i + 1] * R[3 * j + 1] + T[3 * i + 2] * R[3 * j + 2];
      real off[3] = {xipos[0] - com[0], xipos[1] - com[1], xipos[2] - com[2]};
      real o2 = dot3(off, off);
      cinert[0] = G[0] + mass * (o2 - off[0] * off[0]); cinert[1] = G[4] + mass * (o2 - off[1] * off[1]); cinert[2] = G[8] + mass * (o2 - off[2] * off[2]);
      cinert[3] = G[1] - mass * off[0] * off[1]; cinert[4] = G[2] - mass * off[0] * off[2]; cinert[5] = G[5] - mass * off[1] * off[2];
      cinert[6] = mass * off[0]; cinert[7] = mass * off[1]; cinert[8] = mass * off[2]; cinert[9] = mass;
    }
    // cdof of own dofs -> shared (predicated stores); hub lanes 0..5 own the free-joint dofs
    // own cdof live in shared memory (s_cdof + 8*dj[j]); hub-dof lanes use s_cdof + 8*hl
    {
      real off[3] = {com[0] - xpos[0], com[1] - xpos[1], com[2] - xpos[2]};
#pragma unroll
      for (int j = 0; j < 3; j++) {
        real l[3]; cross3(laxis + 3 * j, off, l);
        if (j < ndof) {
          real* cd = s_cdof + CDS * dj[j];
          cd[0] = laxis[3 * j]; cd[1] = laxis[3 * j + 1]; cd[2] = laxis[3 * j + 2]; cd[3] = l[0]; cd[4] = l[1]; cd[5] = l[2];
        }
      }
      if (hubdof) {
        real* cd = s_cdof + CDS * hl;
        const int a = hl < 3 ? 0 : hl - 3;
        real ax[3] = {R[a], R[3 + a], R[6 + a]}, l[3]; cross3(ax, off, l);
        const bool tr = hl < 3;
        cd[0] = tr ? real(0.) : ax[0]; cd[1] = tr ? real(0.) : ax[1]; cd[2] = tr ? real(0.) : ax[2];
        cd[3] = tr ? (hl == 0 ? real(1.) : real(0.)) : l[0]; cd[4] = tr ? (hl == 1 ? real(1.) : real(0.)) : l[1]; cd[5] = tr ? (hl == 2 ? real(1.) : real(0.)) : l[2];
      }
      if (!is_leg && hl == 0) {
        // hub velocity / bias acceleration (free joint: translations first, rotations against the updated velocity)
        real wl[3] = {st[S_QVEL + 3], st[S_QVEL + 4], st[S_QVEL + 5]};
        real cv0[6] = {real(0.), real(0.), real(0.), st[S_QVEL], st[S_QVEL + 1], st[S_QVEL + 2]};
        real cacc[6] = {real(0.), real(0.), real(0.), -p.gx, -p.gy, -p.gz};
        real cvel[6] = {cv0[0], cv0[1], cv0[2], cv0[3], cv0[4], cv0[5]};
        real Sh[6] = {0, 0, 0, st[S_WARM], st[S_WARM + 1], st[S_WARM + 2]};
#pragma unroll
        for (int a = 0; a < 3; a++) {
          real ax[3] = {R[a], R[3 + a], R[6 + a]}, l[3]; cross3(ax, off, l);
          real cd[6] = {ax[0], ax[1], ax[2], l[0], l[1], l[2]}, cdd[6];
          cross_motion(cv0, cd, cdd);
          real qa = st[S_WARM + 3 + a];
#pragma unroll
          for (int i = 0; i < 6; i++) { cacc[i] += cdd[i] * wl[a]; cvel[i] += cd[i] * wl[a]; Sh[i] += cd[i] * qa; }
        }
#pragma unroll
        for (int i = 0; i < 6; i++) { s_hub[HU_CVEL + i] = cvel[i]; s_hub[HU_CACC + i] = cacc[i]; sm[SM_HBB + HB_SH + i] = Sh[i]; }
      }
    }
    block_sync(bar);   // cdof, hub cvel/cacc, S_h visible

    // =====================================================================
    // B. velocities, composite inertia, collision, bias + actuator forces (all lanes, convergent)
    // =====================================================================
    real crb[10], cvel[6], Sa[6];
    Con con[NSLOT];
    real fs_own[3] = {real(0.), real(0.), real(0.)};
    real actf[3] = {real(0.), real(0.), real(0.)}, adhf = real(0.);
    {
      real qv[3], loc[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < 3; j++) {
        qv[j] = msk[j] * st[S_QVEL + dj[j]];
#pragma unroll
        for (int i = 0; i < 6; i++) loc[i] += CDO(j)[i] * qv[j];
      }
      real pre[6] = {loc[0], loc[1], loc[2], loc[3], loc[4], loc[5]};
      chain_prefix<6>(pre, NMF_FULL, k);
      real cv[6], ad[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int i = 0; i < 6; i++) cv[i] = pre[i] - loc[i] + s_hub[HU_CVEL + i];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        real cdd[6]; cross_motion(cv, CDO(j), cdd);
#pragma unroll
        for (int i = 0; i < 6; i++) { ad[i] += cdd[i] * qv[j]; cv[i] += CDO(j)[i] * qv[j]; }
      }
#pragma unroll
      for (int i = 0; i < 6; i++) cvel[i] = cv[i];
      chain_prefix<6>(ad, NMF_FULL, k);
      real cacc[6];
#pragma unroll
      for (int i = 0; i < 6; i++) cacc[i] = ad[i] + s_hub[HU_CACC + i];
      // body wrench  W = -(I a + v x* I v)  (+ adhesion below)
      real t1[6], t2[6], t3[6], W[6];
      mul_inert(cinert, cacc, t1); mul_inert(cinert, cvel, t2); cross_force(cvel, t2, t3);
      if (p.out_energy && step == p.nsteps - 1) {
        // `energy` flag of the reference model (mujoco_globals.yaml:19): potential = -sum m g.x + joint springs, kinetic = 1/2 v'Mv
        real e[2] = {-mass * (p.gx * xipos[0] + p.gy * xipos[1] + p.gz * xipos[2]), real(0.5) * dot6(cvel, t2)};
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const real dq = msk[j] * (st[S_QPOS + 1 + dj[j]] - role[(RF_SREF + j) * CTA + tid]);
          e[0] += real(0.5) * role[(RF_STIFF + j) * CTA + tid] * dq * dq;
          e[1] += real(0.5) * armv[j] * qv[j] * qv[j];
        }
        cta_reduce<2>(e, s_red, parity, tid, bar);
        if (tid == 0) { p.out_energy[2 * (size_t)fly] = (float)e[0]; p.out_energy[2 * (size_t)fly + 1] = (float)e[1]; }
      }
#pragma unroll
      for (int i = 0; i < 6; i++) W[i] = -(t1[i] + t3[i]);
      // composite inertia
#pragma unroll
      for (int i = 0; i < 10; i++) crb[i] = cinert[i];
      chain_suffix<10>(crb, NMF_FULL, k);
      // collision for this body's geom
      const real ncon_lane = collide<NSLOT>(p, role, tid, xpos, R, com, cvel, invw, con, hullv);
      // adhesion (body transmission): force pulls the body onto the plane along each contact normal
      {
        const int acidx = role_int(role[RF_ADH_CIDX * CTA + tid]);
        real c = m_min(role[RF_ADH_HI * CTA + tid], m_max(role[RF_ADH_LO * CTA + tid], st[S_CTRL + (acidx >= 0 ? acidx : 0)]));
        adhf = role[RF_ADH_GAIN * CTA + tid] * c;     // gain = 0 on lanes without an adhesion actuator
        if constexpr (WORLD == W_TERRAIN) {
          const real pull = ncon_lane > real(0.) ? adhf / ncon_lane : real(0.);
#pragma unroll
          for (int s = 0; s < NSLOT; s++) adhesion_wrench(con[s], pull, W);
        } else {   // z-normal slots: written out in place (routing this through a helper cost 3 % on B200: register allocation)
          real fz = ncon_lane > real(0.) ? -adhf / ncon_lane : real(0.);
#pragma unroll
          for (int s = 0; s < NSLOT; s++) { real f = con_on(con[s]) * fz; W[0] += con[s].r[1] * f; W[1] -= con[s].r[0] * f; W[5] += f; }
        }
      }
      chain_suffix<6>(W, NMF_FULL, k);
      // joint-space smooth force of own dofs: passive + actuator + C'W
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int d = dj[j];
        real q = st[S_QPOS + 1 + d], qvj = st[S_QVEL + d];
        real f = -role[(RF_STIFF + j) * CTA + tid] * (q - role[(RF_SREF + j) * CTA + tid]) - role[(RF_DAMP + j) * CTA + tid] * qvj;
        const int ci = role_int(role[(RF_CIDX + j) * CTA + tid]);
        real kp = role[(RF_KP + j) * CTA + tid], kv = role[(RF_KV + j) * CTA + tid];     // 0 without an actuator
        real af = kp * st[S_CTRL + (ci >= 0 ? ci : 0)] - kp * q - kv * qvj;
        af = m_min(role[(RF_FHI + j) * CTA + tid], m_max(role[(RF_FLO + j) * CTA + tid], af));
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
      real sl[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < 3; j++) {
        real qa = msk[j] * st[S_WARM + dj[j]];
#pragma unroll
        for (int i = 0; i < 6; i++) sl[i] += CDO(j)[i] * qa;
      }
      chain_prefix<6>(sl, NMF_FULL, k);
#pragma unroll
      for (int i = 0; i < 6; i++) Sa[i] = sl[i] + sm[SM_HBB + HB_SH + i];
      // contact rows at the warm-start acceleration
#pragma unroll
      for (int s = 0; s < NSLOT; s++) {
        real ap[3]; project_point(con[s], Sa, p.mu, ap);
        con[s].w[0] += ap[0]; con[s].w[1] += ap[1]; con[s].w[2] += ap[2];
      }
      if (TETHER && weld_lane) weld_setup(p, sw, qh, xh, com, s_hub + HU_CVEL, Sa);
    }
    bool any[NSLOT];     // warp-uniform: some lane of the warp uses slot s (most lanes have no contact)
#pragma unroll
    for (int s = 0; s < NSLOT; s++) any[s] = __any_sync(NMF_FULL, con[s].D > real(0.));
    block_sync(bar);   // chain roots (wrench, crb) visible to the hub lanes
    real crbh[10];    // hub-dof lanes: composite inertia of the whole fly
    if (!is_leg) hub_root_totals(sm, hl, 0, 16);
    __syncwarp(NMF_FULL);
    if (hubdof) {
      real W[6];
#pragma unroll
      for (int i = 0; i < 10; i++) crbh[i] = sm[SM_HBB + HB_TOT + 6 + i];
#pragma unroll
      for (int i = 0; i < 6; i++) W[i] = sm[SM_HBB + HB_TOT + i];
      fs_own[0] = dot6(s_cdof + CDS * hl, W); sm[SM_FS + hl] = fs_own[0];
    }
    block_sync(bar);   // roots consumed before the solver overwrites them

    // =====================================================================
    // C. soft-contact solve: primal Newton on  1/2 (a-a0)'M(a-a0) + s(Ja - aref)
    //    started from qacc_warmstart (unique minimiser => same result as the
    //    reference's mj_solNewton; solver=Newton in mujoco_globals.yaml:12)
    // =====================================================================
    real* qacc = st + S_WARM;      // qacc lives in the warm-start slot of the record
    int niter = 0, nls_total = 0, nchanged_last = 0;
    int fault = 0;                 // bits of the per-fly status word raised by this step (nmf_layout.h ST_*)
    real G[NSLOT][3];              // noslip: explicit contact forces in basis coordinates (normal, mu t1, mu t2), valid when explicit_f
    bool explicit_f = false;
    // One loop body serves every Newton iteration AND the final implicit-damping (Euler) solve, so the large unrolled
    // factorisation exists once in the instruction stream (the kernel is I-cache sensitive):
    //   pass `iter`:  forces(qacc) -> gradient/fc -> [converged? euler : newton] system -> arrowhead solve -> (line search, move)
    bool running = true;
    for (int iter = 0;; iter++) {
      if (FPB > 1) {   // pass boundary: the flies of the block realign; a fly that is through waits here for the others
        __syncwarp(NMF_FULL);
        if (!__syncthreads_or(running ? 1 : 0)) break;
        if (!running) continue;
      }
      const bool euler = iter > 0 && (nchanged_last == 0 || iter >= p.max_newton);
      if (euler && nchanged_last != 0) fault |= ST_NEWTON_CAP;      // iteration cap reached with the active set still changing
      if (NOSLIP && euler && p.noslip_iterations > 0) {
        // =================================================================
        // noslip post-solver of the reference's CPU path (mujoco_globals.yaml:15; [PRIOR] mj_solNoSlip): projected Gauss-Seidel
        // on the friction dimensions of the DUAL problem without the regulariser.  In the basis (n, mu t1, mu t2) of every
        // contact the update of a pair of opposing pyramid edges is a plain Gauss-Seidel step on one tangential component,
        //     g <- clamp(g - h / B_gg, -lim, lim),   h = jar_t(qacc) + B_tt (g - g_newton),   B_tt = E_t M^-1 E_t',
        // with lim = the normal force the pair carries (kept).  B_tt is built column by column with the arrowhead solver on
        // the plain inertia matrix (the matrix-free J / J' of the rest of the kernel), the sweeps run on it in shared memory
        // in the oracle's contact order (geom order, then slot), and qacc moves by M^-1 E_t' (g - g_newton) at the end.
        // =================================================================
        real* ns = sm + SM_TOTAL;
        int* keys = reinterpret_cast<int*>(ns + NS_B);
        const int gidx = role_int(role[RF_GIDX * CTA + tid]);
        real lim[NSLOT][2], costr = real(0.);
#pragma unroll
        for (int s = 0; s < NSLOT; s++) {
          real c1; basis_forces(con[s], G[s], lim[s], c1); costr += c1;
          keys[tid * NSLOT + s] = con[s].D > real(0.) ? gidx * NSLOT + s : 0x7fffffff;
        }
        real c0r[1] = {costr};
        cta_reduce<1>(c0r, s_red, parity, tid, bar);            // (also makes the keys visible)
        int rank[NSLOT], C = 0;
#pragma unroll
        for (int s = 0; s < NSLOT; s++) rank[s] = 0;
        for (int i = 0; i < CTA * NSLOT; i++) {
          const int ki = keys[i];
          C += ki != 0x7fffffff ? 1 : 0;
#pragma unroll
          for (int s = 0; s < NSLOT; s++) rank[s] += ki < keys[tid * NSLOT + s] ? 1 : 0;
        }
        block_sync(bar);                                        // keys alias B: everybody has ranked before B is written
        if (C > NS_MAXC) fault |= ST_NOSLIP_SKIP;
        if (TETHER || (C > 0 && C <= NS_MAXC)) {
          // ---- stage the plain inertia matrix (the same staging the Euler pass repeats with its damping diagonal)
          {
            real P[21];
            if (hubdof) {
              expand_inert(crbh, P);
              real u[6]; sym6_mul(P, s_cdof + CDS * hl, u);
              for (int c = 0; c <= hl; c++) sm[SM_HBB + HB_S + hl * (hl + 1) / 2 + c] = dot6(s_cdof + CDS * c, u);
            }
            expand_inert(crb, P);
#pragma unroll
            for (int j = 0; j < 3; j++) {
              real u[6]; sym6_mul(P, CDO(j), u);
              if (j < ndof) {
                real* up = su + 8 * (ldof0 + j);
#pragma unroll
                for (int i = 0; i < 6; i++) up[i] = u[i];
              }
            }
          }
          block_sync(bar);
          // x = M^-1 J'(lane wrenches W): on return W holds the spatial acceleration of x at this lane's body, xo its own DoFs
          auto m_solve = [&](real* W, real* xo) {
            chain_suffix<6>(W, NMF_FULL, k);
#pragma unroll
            for (int j = 0; j < 3; j++) if (j < ndof) sm[SM_GRAD + dj[j]] = -dot6(CDO(j), W);
            if (k == 0) {
#pragma unroll
              for (int i = 0; i < 6; i++) rt[i] = W[i];
            }
            block_sync(bar);
            if (!is_leg) hub_root_totals(sm, hl, 0, 6);
            __syncwarp(NMF_FULL);
            if (hubdof) sm[SM_GRAD + hl] = -dot6(s_cdof + CDS * hl, sm + SM_HBB + HB_TOT);
            block_sync(bar);
            arrowhead_solve(sm, s_cdof, cl, grp, t, is_leg, hl, lbase, pb, pc, nullptr, bar);
            __syncwarp(NMF_FULL);
#pragma unroll
            for (int j = 0; j < 3; j++) xo[j] = msk[j] * sm[SM_X + dj[j]];
            if (hubdof) xo[0] = sm[SM_X + hl];
            real sl[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
              for (int i = 0; i < 6; i++) sl[i] += (is_leg ? CDO(j)[i] : real(0.)) * xo[j];
            chain_prefix<6>(sl, NMF_FULL, k);
#pragma unroll
            for (int i = 0; i < 6; i++) W[i] = sl[i] + sm[SM_HBB + HB_SH + i];
            block_sync(bar);                                    // SM_GRAD / SM_X / the root records are free again
          };
          if (TETHER) {
            // TetheredWorld: the six weld rows are equality rows, which noslip sweeps unclamped ( f_i -= residual_i / A_ii ).
            // A = J_w M^-1 J_w' (6 x 6) column by column; everything lives on the weld lane.
            for (int j = 0; j < 6; j++) {
              real W[6] = {0, 0, 0, 0, 0, 0}, xo[3];
              if (weld_lane) {
                const real* r = sw + WL_R; const real* Gm = sw + WL_G;
                if (j < 3) { real e3[3] = {j == 0 ? real(1.) : real(0.), j == 1 ? real(1.) : real(0.), j == 2 ? real(1.) : real(0.)}, T[3]; cross3(r, e3, T);
                             W[0] = T[0]; W[1] = T[1]; W[2] = T[2]; W[3] = e3[0]; W[4] = e3[1]; W[5] = e3[2]; }
                else { W[0] = Gm[3 * (j - 3)]; W[1] = Gm[3 * (j - 3) + 1]; W[2] = Gm[3 * (j - 3) + 2]; }
              }
              m_solve(W, xo);
              if (weld_lane) { real o6[6]; point_and_rot(sw, W, o6); for (int i = 0; i < 6; i++) ns[NS_B + i * NS_LD + j] = o6[i]; }
            }
            block_sync(bar);
            if (weld_lane) {
              real f[6], f0[6], jar[6], improvement0 = real(0.);
              for (int i = 0; i < 6; i++) {
                jar[i] = sw[WL_W + i] + sw[WL_C0 + i]; f0[i] = f[i] = -sw[WL_D + i] * jar[i];
                improvement0 += real(0.5) * f0[i] * f0[i] / sw[WL_D + i];
              }
              for (int it = 0; it < p.noslip_iterations; it++) {
                real improvement = it == 0 ? improvement0 : real(0.);
                for (int i = 0; i < 6; i++) {
                  real res = jar[i];
                  for (int q = 0; q < 6; q++) res += ns[NS_B + i * NS_LD + q] * (f[q] - f0[q]);
                  const real Aii = ns[NS_B + i * NS_LD + i], old = f[i];
                  f[i] = old - res / m_max(real(1e-15), Aii);
                  const real dd = f[i] - old;
                  real change = real(0.5) * dd * dd * Aii + dd * res;
                  if (change > real(1e-10)) { f[i] = old; change = real(0.); }
                  improvement -= change;
                }
                if (improvement * p.noslip_scale < p.noslip_tol) break;
              }
              for (int i = 0; i < 6; i++) { sw[WL_F + i] = f[i]; ns[NS_G + i] = f[i] - f0[i]; }
            }
            block_sync(bar);
            {
              real W[6] = {0, 0, 0, 0, 0, 0}, xo[3];
              if (weld_lane) {
                const real* r = sw + WL_R; const real* Gm = sw + WL_G; const real* df = ns + NS_G;
                real T[3]; cross3(r, df, T);
                for (int i = 0; i < 3; i++) { W[i] = T[i] + Gm[i] * df[3] + Gm[3 + i] * df[4] + Gm[6 + i] * df[5]; W[3 + i] = df[i]; }
              }
              m_solve(W, xo);
#pragma unroll
              for (int j = 0; j < 3; j++) if (j < ndof || (j == 0 && hubdof)) qacc[dj[j]] += xo[j];
#pragma unroll
              for (int i = 0; i < 6; i++) Sa[i] += W[i];
              if (weld_lane) { real o6[6]; point_and_rot(sw, W, o6); for (int i = 0; i < 6; i++) sw[WL_W + i] += o6[i]; }
              block_sync(bar);
            }
          } else {
          // ---- columns of B_tt
          for (int j = 0; j < 2 * C; j++) {
            real W[6] = {0, 0, 0, 0, 0, 0}, xo[3];
#pragma unroll
            for (int s = 0; s < NSLOT; s++)
              if (con[s].D > real(0.) && rank[s] == (j >> 1)) {
                real d[3], T[3]; contact_tangent(con[s], j & 1, p.mu, d); cross3(con[s].r, d, T);
                W[0] = T[0]; W[1] = T[1]; W[2] = T[2]; W[3] = d[0]; W[4] = d[1]; W[5] = d[2];
              }
            m_solve(W, xo);
#pragma unroll
            for (int s = 0; s < NSLOT; s++)
              if (con[s].D > real(0.)) {
                real o3[3]; project_point(con[s], W, p.mu, o3);
                ns[NS_B + (2 * rank[s]) * NS_LD + j] = o3[1]; ns[NS_B + (2 * rank[s] + 1) * NS_LD + j] = o3[2];
              }
          }
#pragma unroll
          for (int s = 0; s < NSLOT; s++)
            if (con[s].D > real(0.)) {
#pragma unroll
              for (int q = 0; q < 2; q++) {
                const int i = 2 * rank[s] + q;
                ns[NS_G + i] = G[s][1 + q]; ns[NS_G0 + i] = G[s][1 + q]; ns[NS_LIM + i] = lim[s][q]; ns[NS_JT + i] = con[s].w[1 + q];
              }
            }
          block_sync(bar);
          // ---- the sweeps (serial Gauss-Seidel, one thread; the problem is 2C <= 48 unknowns)
          if (tid == 0) {
            const int n2 = 2 * C;
            for (int it = 0; it < p.noslip_iterations; it++) {
              real improvement = it == 0 ? c0r[0] : real(0.);
              for (int i = 0; i < n2; i++) {
                real h = ns[NS_JT + i];
                for (int q = 0; q < n2; q++) h += ns[NS_B + i * NS_LD + q] * (ns[NS_G + q] - ns[NS_G0 + q]);
                const real Bii = ns[NS_B + i * NS_LD + i], gold = ns[NS_G + i], l = ns[NS_LIM + i];
                real gnew = real(0.);                           // K1 = 4 B_ii below MuJoCo's mjMINVAL: both edges get the mean
                if (real(4.) * Bii >= real(1e-15)) gnew = m_min(l, m_max(-l, gold - h / Bii));
                const real dy = real(0.5) * (gnew - gold);      // y = (f_j - f_j+1) / 2
                real change = real(2.) * Bii * dy * dy + real(2.) * h * dy;
                if (change > real(1e-10)) { gnew = gold; change = real(0.); }
                ns[NS_G + i] = gnew; improvement -= change;
              }
              if (improvement * p.noslip_scale < p.noslip_tol) break;
            }
          }
          block_sync(bar);
          // ---- move: qacc += M^-1 E_t' (g - g_newton); rows and explicit forces follow
          {
            real W[6] = {0, 0, 0, 0, 0, 0}, xo[3];
#pragma unroll
            for (int s = 0; s < NSLOT; s++)
              if (con[s].D > real(0.)) {
                real dG[3] = {real(0.), ns[NS_G + 2 * rank[s]] - G[s][1], ns[NS_G + 2 * rank[s] + 1] - G[s][2]};
                basis_wrench(con[s], dG, p.mu, W, nullptr);
                G[s][1] += dG[1]; G[s][2] += dG[2];
              }
            m_solve(W, xo);
#pragma unroll
            for (int j = 0; j < 3; j++) if (j < ndof || (j == 0 && hubdof)) qacc[dj[j]] += xo[j];
#pragma unroll
            for (int i = 0; i < 6; i++) Sa[i] += W[i];
#pragma unroll
            for (int s = 0; s < NSLOT; s++) {
              real o3[3]; project_point(con[s], W, p.mu, o3);
              con[s].w[0] += o3[0]; con[s].w[1] += o3[1]; con[s].w[2] += o3[2];
            }
            block_sync(bar);
          }
          }
          explicit_f = true;
        }
      }
      // ---- forces, active set, contact augmentation
      real Wc[6] = {0, 0, 0, 0, 0, 0}, A[21];
#pragma unroll
      for (int i = 0; i < 21; i++) A[i] = real(0.);
      if (NOSLIP && explicit_f) {      // (Euler pass after noslip: the forces are no longer a function of the rows)
#pragma unroll
        for (int s = 0; s < NSLOT; s++) if (any[s] && con[s].D > real(0.)) basis_wrench(con[s], G[s], p.mu, Wc, nullptr);
      } else {
#pragma unroll
        for (int s = 0; s < NSLOT; s++) if (any[s]) contact_forces<true>(con[s], p.mu, Wc, A, nullptr);    // warp-uniform skips
      }
      if (TETHER && weld_lane) weld_forces(sw, Wc, A, NOSLIP && explicit_f);
      // ---- gradient  g = C' suffix(I S - Wc) + armature a - fs ;  fc = C' suffix(Wc)
      real y[12];
      {
        real t6[6]; mul_inert(cinert, Sa, t6);
#pragma unroll
        for (int i = 0; i < 6; i++) { y[i] = t6[i] - Wc[i]; y[6 + i] = Wc[i]; }
      }
      real gown[3] = {real(0.), real(0.), real(0.)};
      chain_suffix<12>(y, NMF_FULL, k);
      if (!euler) chain_suffix<21>(A, NMF_FULL, k);
      const real am = euler ? real(0.) : real(1.);     // the Euler system uses the plain inertia (no contact augmentation)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int d = dj[j];
        real fc = dot6(CDO(j), y + 6);
        real g = dot6(CDO(j), y) + armv[j] * qacc[d] - fs_own[j];
        gown[j] = msk[j] * g;
        if (j < ndof) sm[SM_GRAD + d] = euler ? -(fs_own[j] + fc) : g;   // right-hand side is -(this slot)
      }
      if (k == 0) {
#pragma unroll
        for (int i = 0; i < 12; i++) rt[i] = y[i];
#pragma unroll
        for (int i = 0; i < 21; i++) rt[16 + i] = am * A[i];
      }
      block_sync(bar);
      {
        real P[21];
        if (!is_leg) hub_root_totals(sm, hl, 0, 37);
        __syncwarp(NMF_FULL);
        if (hubdof) {
          real yh[12];
#pragma unroll
          for (int i = 0; i < 12; i++) yh[i] = sm[SM_HBB + HB_TOT + i];
          real fc = dot6(s_cdof + CDS * hl, yh + 6), g = dot6(s_cdof + CDS * hl, yh) - fs_own[0];
          gown[0] = g; sm[SM_GRAD + hl] = euler ? -(fs_own[0] + fc) : g;
          expand_inert(crbh, P);
#pragma unroll
          for (int i = 0; i < 21; i++) P[i] += sm[SM_HBB + HB_TOT + 16 + i];
          real u[6]; sym6_mul(P, s_cdof + CDS * hl, u);
          for (int c = 0; c <= hl; c++) sm[SM_HBB + HB_S + hl * (hl + 1) / 2 + c] = dot6(s_cdof + CDS * c, u);
        }
        __syncwarp(NMF_FULL);
        // ---- system matrix  H = C'(crb + A-hat)C + diag : each DoF owner stages u = P cdof, columns are formed by the factoriser
        expand_inert(crb, P);
#pragma unroll
        for (int i = 0; i < 21; i++) P[i] += am * A[i];
#pragma unroll
        for (int j = 0; j < 3; j++) {
          real u[6]; sym6_mul(P, CDO(j), u);
          if (j < ndof) {
            real* up = su + 8 * (ldof0 + j);
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
        arrowhead_solve(sm, s_cdof, c, grp, t, is_leg, hl, lbase, pb, pc, dbg_rows, bar);
      }
      if (euler) { niter = iter; if (FPB == 1) break; running = false; continue; }
      __syncwarp(NMF_FULL);
      real sown[3];
#pragma unroll
      for (int j = 0; j < 3; j++) sown[j] = msk[j] * sm[SM_X + dj[j]];
      if (hubdof) sown[0] = sm[SM_X + hl];

      // ---- spatial acceleration of the search direction, row directions, quadratic terms
      real Ss[6];
      {
        real sl[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
          for (int i = 0; i < 6; i++) sl[i] += (is_leg ? CDO(j)[i] : real(0.)) * sown[j];
        chain_prefix<6>(sl, NMF_FULL, k);
#pragma unroll
        for (int i = 0; i < 6; i++) Ss[i] = sl[i] + sm[SM_HBB + HB_SH + i];
      }
      real red[5] = {real(0.), real(0.), real(0.), real(0.), real(0.)};   // s.g , s'Ms , d0 rows(0), d1 rows(0), |s|^2
      real sv[NSLOT][3];   // row directions of the contact slots along the search vector
#pragma unroll
      for (int s = 0; s < NSLOT; s++) sv[s][0] = sv[s][1] = sv[s][2] = real(0.);
      {
        real t6[6]; mul_inert(cinert, Ss, t6);
        red[1] += dot6(Ss, t6);
#pragma unroll
        for (int j = 0; j < 3; j++) { red[0] += sown[j] * gown[j]; red[1] += (is_leg ? armv[j] : real(0.)) * sown[j] * sown[j]; red[4] += sown[j] * sown[j]; }
        real dummy = real(0.);
#pragma unroll
        for (int s = 0; s < NSLOT; s++) if (any[s]) { project_point(con[s], Ss, p.mu, sv[s]); ls_eval(con[s], sv[s], real(0.), red[2], red[3], dummy); }
        if (TETHER && weld_lane) { point_and_rot(sw, Ss, sw + WL_SV); weld_ls(sw, real(0.), red[2], red[3]); }
      }
      cta_reduce<5>(red, s_red, parity, tid, bar);
      // ---- exact line search along the Newton direction (safeguarded Newton on the derivative)
      real alpha = real(0.);
      nchanged_last = 0;
      {
        const real q1 = red[0] - red[2], q2 = red[1];
        real d0 = red[0], d1 = q2 + red[3], lo = real(0.), hi = real(3.0e38);
        const int nls = (red[4] > real(1e-30) && d1 > real(0.)) ? p.max_ls : 0;   // zero direction: nothing to search
        for (int it = 0; it < nls; it++) {
          if (it > 0 && (m_abs(d0) <= Prec<real>::ls_rel * d1 * m_max(m_abs(alpha), Prec<real>::ls_amin) || (hi < real(1.0e38) && hi - lo <= Prec<real>::ls_bracket * hi))) break;
          if (it == nls - 1) fault |= ST_LS_CAP;                     // the last allowed evaluation is about to be spent
          if (d0 < real(0.)) lo = alpha; else hi = alpha;
          real nx = alpha - d0 / d1;
          if (nx <= lo || nx >= hi) nx = (hi > real(1.0e38)) ? real(2.) * m_max(alpha, real(1.)) : real(0.5) * (lo + hi);
          alpha = nx;
          real e[3] = {real(0.), real(0.), real(0.)};
#pragma unroll
          for (int s = 0; s < NSLOT; s++) if (any[s]) ls_eval(con[s], sv[s], alpha, e[0], e[1], e[2]);
          if (TETHER && weld_lane) weld_ls(sw, alpha, e[0], e[1]);
          cta_reduce<3>(e, s_red, parity, tid, bar);
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
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int s = 0; s < NSLOT; s++) con[s].w[i] += alpha * sv[s][i];
      if (TETHER && weld_lane) for (int i = 0; i < 6; i++) sw[WL_W + i] += alpha * sw[WL_SV + i];
    }
    // SM_X now holds the implicit-damping (Euler) acceleration  (M + dt diag(damping))^-1 (qfrc_smooth + qfrc_constraint)
    block_sync(bar);

    // ---- optional outputs of this step (derived quantities belong to the pre-integration state, as in mj_step)
    const bool last_step = step == p.nsteps - 1;
    if (last_step) {
      if (p.dbg) {
        float* dg = p.dbg + (size_t)fly * DBG_STRIDE;
        if (tid == 0) { dg[DBG_NITER] = (real)niter; dg[DBG_NLS] = (real)nls_total; dg[DBG_NCHG] = (real)nchanged_last; }
        for (int i = tid; i < NV; i += CTA) { dg[DBG_FS + i] = sm[SM_FS + i]; dg[DBG_QACC + i] = qacc[i]; dg[DBG_FC + i] = -sm[SM_GRAD + i] - sm[SM_FS + i]; dg[DBG_QACCE + i] = sm[SM_X + i]; }
        if (tid < 21) dg[DBG_HROWS + NLEG * 177 + tid] = sm[SM_HBB + HB_S + tid];
        for (int s = 0; s < NSLOT; s++) {
          float* c = dg + DBG_CON + (tid * DBG_NSLOT + s) * 6; real fn = real(0.), Wt[6] = {0, 0, 0, 0, 0, 0};
          if (NOSLIP && explicit_f) { if (con[s].D > real(0.)) basis_wrench(con[s], G[s], p.mu, Wt, nullptr); fn = con[s].D > real(0.) ? G[s][0] : real(0.); }
          else contact_forces<false>(con[s], p.mu, Wt, nullptr, &fn);
          const real on = con_on(con[s]);
          c[0] = on; c[1] = on * con_dist(con[s], com);
          c[2] = on * (con[s].r[0] + com[0]); c[3] = on * (con[s].r[1] + com[1]);
          c[4] = on * (con[s].r[2] + com[2]); c[5] = fn;
        }
        for (int i = 0; i < 3; i++) dg[DBG_XPOS + tid * 3 + i] = xpos[i];
        for (int i = tid; i < NV * 6; i += CTA) dg[DBG_CDOF + i] = s_cdof[CDS * (i / 6) + i % 6];
        real nc[1] = {real(0.)};
#pragma unroll
        for (int s = 0; s < NSLOT; s++) nc[0] += con_on(con[s]);
        cta_reduce<1>(nc, s_red, parity, tid, bar);
        if (tid == 0) dg[DBG_NCON] = nc[0];
      }
      if (p.out_actf) {
        float* o = p.out_actf + (size_t)fly * (p.nu_pos + p.nu_adh);
#pragma unroll
        for (int j = 0; j < 3; j++) { int ci = role_int(role[(RF_CIDX + j) * CTA + tid]); if (j < ndof && ci >= 0) o[ci] = actf[j]; }
        int ai = role_int(role[RF_ADH_CIDX * CTA + tid]); if (ai >= 0) o[ai] = adhf;
      }
      if (p.out_sensor) {
        // per-leg contact sensor (world.py:311-331), reduce="netforce": found, force, torque, pos, normal, tangent
        const real sens = (is_leg && role_int(role[RF_LEGSENSOR * CTA + tid]) != 0) ? real(1.) : real(0.);
        real acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // F(3), fn-weighted pos(3), fn sum, count
        real Fc[NSLOT][3], plain[3] = {0, 0, 0};
#pragma unroll
        for (int s = 0; s < NSLOT; s++) {
          real Wt[6] = {0, 0, 0, 0, 0, 0}, fn = real(0.);
          if (NOSLIP && explicit_f) { if (con[s].D > real(0.)) basis_wrench(con[s], G[s], p.mu, Wt, nullptr); fn = con[s].D > real(0.) ? G[s][0] : real(0.); }
          else contact_forces<false>(con[s], p.mu, Wt, nullptr, &fn);
          const real on = sens * con_on(con[s]);
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
        real P3[3] = {0, 0, 0};
        if (acc[7] > real(0.)) for (int i = 0; i < 3; i++) P3[i] = acc[6] > real(1e-15) ? acc[3 + i] / acc[6] : plain[i] / acc[7];
        real T[3] = {0, 0, 0};
#pragma unroll
        for (int s = 0; s < NSLOT; s++) {
          const real on = sens * con_on(con[s]);
          real rr[3] = {on * (con[s].r[0] + com[0] - P3[0]), on * (con[s].r[1] + com[1] - P3[1]), on * (con[s].r[2] + com[2] - P3[2])}, tt[3];
          cross3(rr, Fc[s], tt); T[0] += tt[0]; T[1] += tt[1]; T[2] += tt[2];
        }
#pragma unroll
        for (int off = 1; off < 8; off <<= 1)
#pragma unroll
          for (int i = 0; i < 3; i++) T[i] += __shfl_xor_sync(NMF_FULL, T[i], off, 8);
        if (is_leg && k == 0) {
          float* o = p.out_sensor + ((size_t)fly * NLEG + grp) * 16;
          o[0] = acc[7];
          for (int i = 0; i < 3; i++) { o[1 + i] = acc[7] > real(0.) ? -acc[i] : real(0.); o[4 + i] = acc[7] > real(0.) ? -T[i] : real(0.); o[7 + i] = P3[i]; }
          o[10] = acc[7] > real(0.) ? real(1.) : real(0.); o[11] = real(0.); o[12] = real(0.); o[13] = real(0.); o[14] = acc[7] > real(0.) ? real(1.) : real(0.); o[15] = real(0.);
        }
      }
      if (p.out_xpos || p.out_xquat) {
        block_sync(bar);   // the u staging is free again: reuse it as the pose exchange buffer
        real* ps = sm + SM_U + tid * 8;
        ps[0] = xpos[0]; ps[1] = xpos[1]; ps[2] = xpos[2]; ps[3] = xq[0]; ps[4] = xq[1]; ps[5] = xq[2]; ps[6] = xq[3];
        block_sync(bar);
        for (int sgi = tid; sgi < p.nseg; sgi += CTA) {
          const float* tb = p.seg_tab + sgi * 8; const real* bp = sm + SM_U + __float_as_int(tb[0]) * 8;
          real lp[3] = {tb[1], tb[2], tb[3]}, lq[4] = {tb[4], tb[5], tb[6], tb[7]}, w[3], wq[4];
          qrot(bp + 3, lp, w); qmul(bp + 3, lq, wq);
          if (p.out_xpos) { float* o = p.out_xpos + ((size_t)fly * p.nseg + sgi) * 3; o[0] = bp[0] + w[0]; o[1] = bp[1] + w[1]; o[2] = bp[2] + w[2]; }
          if (p.out_xquat) { float* o = p.out_xquat + ((size_t)fly * p.nseg + sgi) * 4; o[0] = wq[0]; o[1] = wq[1]; o[2] = wq[2]; o[3] = wq[3]; }
        }
        block_sync(bar);
        for (int i = tid; i < NGROUP * U_STRIDE; i += CTA) sm[SM_U + i] = real(0.);   // restore the staging invariant (hub chains: u = 0)
      }
    }

    // ---- advance: qvel += dt a' ; positions integrate with the NEW velocity ; qacc stays as next warm start
    block_sync(bar);
    int* s_fault = reinterpret_cast<int*>(sm + SM_WORK);       // raised by any thread that sees a non-finite velocity (benign race: all write 1)
    if (tid == 0) *s_fault = 0;
    block_sync(bar);
    for (int i = tid; i < NV; i += CTA) {
      const real v = st[S_QVEL + i] + p.dt * sm[SM_X + i];
      st[S_QVEL + i] = v;
      if (!(m_abs(v) < real(3.0e38))) *s_fault = 1;            // NaN or infinity
    }
    block_sync(bar);
    for (int i = tid + 6; i < NV; i += CTA) st[S_QPOS + 1 + i] += p.dt * st[S_QVEL + i];
    if (tid == 0) {
      if (*s_fault) fault |= ST_NONFINITE;
      if (fault) st[S_TIME + 1] = (real)((int)st[S_TIME + 1] | fault);    // sticky until the fly is reset
      for (int i = 0; i < 3; i++) st[S_QPOS + i] += p.dt * st[S_QVEL + i];
      real w[3] = {st[S_QVEL + 3], st[S_QVEL + 4], st[S_QVEL + 5]};
      real n = m_sqrt(dot3(w, w));
      real q[4] = {st[S_QPOS + 3], st[S_QPOS + 4], st[S_QPOS + 5], st[S_QPOS + 6]};
      if (n > real(1e-15)) {
        real sn, cs; sincos_small(real(0.5) * p.dt * n, &sn, &cs);
        real dq[4] = {cs, w[0] / n * sn, w[1] / n * sn, w[2] / n * sn}, nq[4];
        qmul(q, dq, nq); q[0] = nq[0]; q[1] = nq[1]; q[2] = nq[2]; q[3] = nq[3];
      }
      qnormalize(q);
      st[S_QPOS + 3] = q[0]; st[S_QPOS + 4] = q[1]; st[S_QPOS + 5] = q[2]; st[S_QPOS + 6] = q[3];
      st[S_TIME + 2] += real(1.);                 // step count since reset (exact up to 2^24 in float32)
      st[S_TIME] = st[S_TIME + 2] * p.dt;         // time = n dt, not an accumulated sum (5e4 float32 additions drift by 1e-4 relative)
    }
    block_sync(bar);
  }

  // ---- write the record back (TMA bulk store)
  if (p.out_qpos && step0 + nsub == p.nsteps) {
    float* o = p.out_qpos + (size_t)fly * NQ;
    for (int i = tid; i < NQ; i += CTA) o[i] = (float)st[S_QPOS + i];
  }
  if (!p.forward_only) store_record(p, st, sm, fly, tid, published, bar);
}


#ifndef NMF_SIMT_EMU
// Two schedules, one call site of the (large) step body.  A work unit = FPB consecutive flies = what one block steps:
//  * p.queue == nullptr: block b advances unit b by all p.nsteps steps (grid = number of units);
//  * work queue: the launch is cut into items (unit, sub-chunk of p.sub_steps steps) served to a grid that just fills the
//    GPU.  The number of units is rarely a multiple of the resident blocks, and this latency-bound kernel slows down in
//    proportion to the empty slots of a partial last wave; with items a launch is many waves long instead of one or two.
//    The queue is a FIFO of READY units: entries 0..n-1 are implicit (every unit's first sub-chunk), and a block that has
//    written the records back appends the unit again (release) unless that was its last sub-chunk.
//    queue[0] = pop counter, queue[1] = push counter, queue[2 + u] = sub-chunks of unit u done, queue[2 + n + j] = ring entry j.
//    A block may have to wait for a ring entry, but only for items that are running on OTHER blocks (its own previous item
//    has been pushed before it pops), so the wait always ends.
template <int WORLD, int FPB = 1, bool NOSLIP = false>
__device__ __forceinline__ void step_entry(const SP& p) {
  constexpr int SM_FLY = NOSLIP ? SM_TOTAL + NS_COUNT : (WORLD == W_TETHER ? SM_TOTAL : SM_WELD);   // weld rows: tethered world only; noslip region: noslip kernels only
  constexpr bool DYN = (size_t)FPB * SM_FLY * sizeof(real) > 48 * 1024;   // beyond the static limit: dynamic shared memory (opt-in on the host side)
  __shared__ __align__(16) real sm_static[DYN ? 1 : FPB * SM_FLY];
  extern __shared__ __align__(16) unsigned char sm_dynamic[];
  __shared__ int s_fly, s_chunk;
  const int tid = threadIdx.x, slot = fly_slot<FPB>();
  real* sm = (DYN ? reinterpret_cast<real*>(sm_dynamic) : sm_static) + slot * SM_FLY;
  const int n_units = (p.n_flies + FPB - 1) / FPB;
  for (;;) {
    int unit = blockIdx.x, step0 = 0, nsub = p.nsteps;
    if (p.queue) {
      if (tid == 0) {
        const int i = NMF_ATOMIC_ADD(p.queue, 1);
        int f = -1;
        if (i < n_units) f = i;
        else if (i < p.n_items) {
          const int* slot_p = p.queue + 2 + n_units + (i - n_units);
          int v = 0;
          for (unsigned spins = 0; spins < (1u << 24); spins++) {      // bounded: a scheduling bug must not hang the GPU
            v = NMF_LD_ACQUIRE(slot_p);
            if (v) break;
            __nanosleep(64);
          }
          f = v - 1;
          NMF_FENCE_PROXY_ASYNC();            // the records are read through the async proxy (TMA) next
        }
        s_fly = f; s_chunk = f >= 0 ? p.queue[2 + f] : 0;
      }
      block_sync(0);
      unit = warp_uniform(s_fly);
      if (unit < 0) return;
      const int c = warp_uniform(s_chunk);
      if (p.n_chunks > 0) { step0 = p.chunk_start[c]; nsub = p.chunk_start[c + 1] - step0; }
      else { step0 = c * p.sub_steps; nsub = min(p.sub_steps, p.nsteps - step0); }
    }
    const int fly = unit * FPB + slot;
    step_block<WORLD, FPB, NOSLIP>(p, sm, fly < p.n_flies ? fly : -1, step0, nsub, p.queue != nullptr);
    if (!p.queue) return;
    block_sync(0);     // every slot's record store has completed (store_record waited for it); s_fly / s_chunk are free again
    if (tid == 0) {    // hand the unit on
      const int done = s_chunk + 1;            // (s_chunk is rewritten only after the next block_sync)
      p.queue[2 + unit] = done;
      if (step0 + nsub < p.nsteps) {
        NMF_THREADFENCE();
        const int j = NMF_ATOMIC_ADD(p.queue + 1, 1);
        NMF_ST_RELEASE(p.queue + 2 + n_units + j, unit + 1);
      }
    }
  }
}

#endif  // NMF_SIMT_EMU

}  // namespace NMF_NS
}  // namespace nmf
