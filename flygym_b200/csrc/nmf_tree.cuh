// nmf_tree.cuh — general-topology step kernel body, written in terms of `real` (included by nmf_step_all.cuh after
// nmf_step.cuh, once per precision).
//
// Same physics step as nmf_step.cuh (reference GPUSimulation.step -> mujoco_warp.step, src/flygym/warp/simulation.py:260-263;
// Simulation.step -> mj_step, src/flygym/simulation.py:74-76) for ANY free root body carrying a tree of hinge-jointed bodies:
// JointPreset.ALL_BIOLOGICAL / ALL_POSSIBLE skeletons (reference src/flygym/anatomy.py:388-460), ContactBodiesPreset.ALL
// (anatomy.py:519-526), several contact geoms per body.  The star kernels get their speed from the hub + 6 x 8 layout; this
// one reads the topology from tables (nmf_tree_layout.h) and keeps the whole fly in shared memory:
//   * a block of 128 threads owns one fly; each of its 4 warps owns a set of root-child subtrees (legs, abdomen, head with
//     antennae / proboscis / eyes, wings, halteres ...), balanced on the host.  Tree recursions (kinematics, velocities,
//     composite inertias, wrenches) run level by level inside a warp's subtrees with warp-level syncs only; the root body
//     is the only place the warps meet;
//   * the Newton Hessian H = M + J'DJ is formed as a CRBA over contact-augmented spatial inertias (every contact adds X'WX
//     to its body), so it has M's tree sparsity: row i holds entries for the ancestors of DoF i only.  It is factorised
//     L'DL in place in that ancestor-sparse storage, subtree by subtree (one DoF after the other inside a warp, the update
//     of its ancestor rows spread over the lanes); contributions to the 6 x 6 root block are summed per warp and the root
//     block is finished by one thread (the arrowhead scheme of nmf_step.cuh with general subtrees as the "chains");
//   * contacts live in shared memory (general-frame slots, 2 per capsule / 4 per hull with multiccd), so flat and
//     box-column terrain worlds share this one body.
// Solver, line search, integration, outputs and the status word follow nmf_step.cuh line by line.
namespace nmf {
namespace NMF_NS {

typedef TreeParamsT<real> TP;

struct TCon { ContactG c; real sv[3]; real adh; };   // one contact slot in shared memory (TCON_STRIDE reals)
static_assert(sizeof(TCon) == TCON_STRIDE * sizeof(real), "contact slot layout");

__device__ __forceinline__ void tree_sync() { block_sync(0); }

// block-wide sum of N <= 8 values (all TREE_CTA threads)
template <int N>
__device__ __forceinline__ void tree_reduce(real* v, real* s_red, int& parity, int tid) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int n = 0; n < N; n++) v[n] += __shfl_xor_sync(NMF_FULL, v[n], off);
  }
  real* buf = s_red + parity * (TREE_NW * 8);
  if ((tid & 31) == 0) {
#pragma unroll
    for (int n = 0; n < N; n++) buf[(tid >> 5) * 8 + n] = v[n];
  }
  tree_sync();
#pragma unroll
  for (int n = 0; n < N; n++) { real s = real(0.); for (int w = 0; w < TREE_NW; w++) s += buf[w * 8 + n]; v[n] = s; }
  parity ^= 1;
}

// solver parameters and the B*velocity part of the rows of a candidate contact (general frame; see finish_contact)
__device__ __forceinline__ void tree_finish_contact(const TP& p, ContactG& c, real active, real dist, const real* pos, const real* nrm,
                                                    const real* hint, const real* com, const real* cvel, real invw) {
  c.r[0] = pos[0] - com[0]; c.r[1] = pos[1] - com[1]; c.r[2] = pos[2] - com[2];
  c.n[0] = nrm[0]; c.n[1] = nrm[1]; c.n[2] = nrm[2];
  {
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
  const real imp = impedance_of(p.solimp, m_abs(dist - p.margin));
  const real R0 = m_max(real(1e-15), (real(1.) - imp) * invw * (real(1.) + p.mu * p.mu) / imp);
  c.D = active / (real(2.) * (p.mu * p.mu / p.impratio) * R0);
  c.c0 = active * p.cK * imp * (dist - p.margin);
  real vp[3] = {cvel[3] + cvel[1] * c.r[2] - cvel[2] * c.r[1], cvel[4] + cvel[2] * c.r[0] - cvel[0] * c.r[2],
                cvel[5] + cvel[0] * c.r[1] - cvel[1] * c.r[0]};
  real t2v[3]; cross3(c.n, c.t, t2v);
  c.w[0] = active * p.cB * dot3(c.n, vp);
  c.w[1] = active * p.cB * p.mu * dot3(c.t, vp);
  c.w[2] = active * p.cB * p.mu * dot3(t2v, vp);
}

// sphere against the terrain solid (floor plane + grid of box columns): see sphere_terrain in nmf_step.cuh
__device__ __forceinline__ void tree_sphere_terrain(const real* terr, const real* c, real rad, real* nrm, real& dist) {
  const real Px = terr[0], Py = terr[1], hx = terr[2], hy = terr[3];
  nrm[0] = real(0.); nrm[1] = real(0.); nrm[2] = real(1.); dist = c[2] - terr[6] - rad;
  const real fi = m_rint(c[0] / Px), fj = m_rint(c[1] / Py);
  const int i0 = (int)fi, j0 = (int)fj;
  const int sx = c[0] >= fi * Px ? 1 : -1, sy = c[1] >= fj * Py ? 1 : -1;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int i = i0 + ((q & 1) ? sx : 0), j = j0 + ((q & 2) ? sy : 0);
    const real cx = (real)i * Px, cy = (real)j * Py, top = ((i + j) & 1) ? terr[5] : terr[4];
    const real qx = m_min(m_max(c[0], cx - hx), cx + hx), qy = m_min(m_max(c[1], cy - hy), cy + hy), qz = m_min(c[2], top);
    const real dx = c[0] - qx, dy = c[1] - qy, dz = c[2] - qz, d2 = dx * dx + dy * dy + dz * dz;
    real dd, n0, n1, n2;
    if (d2 > real(0.)) { const real inv = m_rsqrt(d2); dd = d2 * inv - rad; n0 = dx * inv; n1 = dy * inv; n2 = dz * inv; }
    else { dd = c[2] - top - rad; n0 = real(0.); n1 = real(0.); n2 = real(1.); }
    if (dd < dist) { dist = dd; nrm[0] = n0; nrm[1] = n1; nrm[2] = n2; }
  }
}

// narrow phase of contact geom g against the ground (plane z = 0, reference world.py:251-260, or the box-column terrain)
__device__ __forceinline__ void tree_collide(const TP& p, int g, const real* xp, const real* q, const real* com, const real* cvel, real invw,
                                             TCon* cs, int& hullv) {
  const TreeDims& d = p.d;
  const int nslot = d.nslot, gtype = p.it[d.i_gtype + g];
  const real* gr = p.rt + d.r_geom + TR_GEOM * g;
  real R[9]; q2mat(q, R);
  const real zn[3] = {real(0.), real(0.), real(1.)}, yh[3] = {real(0.), real(1.), real(0.)}, zero[3] = {real(0.), real(0.), real(0.)};
  int filled = 0;
  if (gtype == 0) {
    real c[3], a[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      c[i] = xp[i] + R[3 * i] * gr[0] + R[3 * i + 1] * gr[1] + R[3 * i + 2] * gr[2];
      a[i] = R[3 * i] * gr[3] + R[3 * i + 1] * gr[4] + R[3 * i + 2] * gr[5];
    }
    const real rad = gr[6], half = gr[7];
    for (int s = 0; s < 2; s++) {
      const real sg = s == 0 ? half : -half;
      real e[3] = {c[0] + sg * a[0], c[1] + sg * a[1], c[2] + sg * a[2]};
      if (p.terrain) {
        real nrm[3], dist; tree_sphere_terrain(p.terr, e, rad, nrm, dist);
        const real back = rad + real(0.5) * dist;
        real pos[3] = {e[0] - back * nrm[0], e[1] - back * nrm[1], e[2] - back * nrm[2]};
        tree_finish_contact(p, cs[s].c, dist <= p.margin ? real(1.) : real(0.), dist, pos, nrm, a, com, cvel, invw);
      } else {
        const real dist = e[2] - rad;
        real pos[3] = {e[0], e[1], real(0.5) * dist};
        tree_finish_contact(p, cs[s].c, e[2] <= p.margin + rad ? real(1.) : real(0.), dist, pos, zn, a, com, cvel, invw);
      }
    }
    filled = 2;
  } else {
    // convex hull: support vertex along -z by a warm-started walk on the hull's vertex graph; with multiccd the neighbours of the
    // support vertex that are within the margin are contacts too, in graph order, up to 4 per geom ([PRIOR] mjc_PlaneConvex)
    const int adr = p.it[d.i_gvadr + g], num = p.it[d.i_gvnum + g];
    int bi = hullv < num ? hullv : 0;
    real best = real(3.0e38);
    if (num > 0) { const real* hv = p.hull + 3 * (adr + bi); best = R[6] * hv[0] + R[7] * hv[1] + R[8] * hv[2]; }
    int extra[3] = {0, 0, 0}, nextra = 0;
    const real zlim = p.margin - xp[2];
    for (int moved = num > 0; moved;) {
      moved = 0; nextra = 0;
      const int n0 = p.hull_nbr_adr[adr + bi], n1 = p.hull_nbr_adr[adr + bi + 1];
      int cand = bi;
      for (int e = n0; e < n1; e++) {
        const int v = p.hull_nbr[e];
        const real* hv = p.hull + 3 * (adr + v);
        const real z = R[6] * hv[0] + R[7] * hv[1] + R[8] * hv[2];
        if (z < best) { best = z; cand = v; moved = 1; }
        if (z <= zlim && nextra < 3) { extra[nextra] = v; nextra++; }
      }
      bi = cand;
    }
    hullv = bi;
    if (!(nslot > 2 && p.multiccd)) nextra = 0;
    const int ns = nslot > 2 ? 4 : 1;
    for (int s = 0; s < ns; s++) {
      const bool on = num > 0 && (s == 0 || s - 1 < nextra);
      const real* hv = p.hull + 3 * (adr + (on ? (s == 0 ? bi : extra[s - 1]) : 0));
      const real h0 = hv[0], h1 = hv[1], h2 = hv[2];
      const real dist = on ? xp[2] + R[6] * h0 + R[7] * h1 + R[8] * h2 : real(1.);
      real pos[3] = {xp[0] + R[0] * h0 + R[1] * h1 + R[2] * h2, xp[1] + R[3] * h0 + R[4] * h1 + R[5] * h2, real(0.5) * dist};
      tree_finish_contact(p, cs[s].c, (on && dist <= p.margin) ? real(1.) : real(0.), dist, pos, zn, yh, com, cvel, invw);
    }
    filled = ns;
  }
  for (int s = filled; s < nslot; s++) tree_finish_contact(p, cs[s].c, real(0.), real(1.), zero, zn, yh, com, cvel, invw);
  for (int s = 0; s < nslot; s++) { cs[s].adh = real(0.); cs[s].sv[0] = cs[s].sv[1] = cs[s].sv[2] = real(0.); }
}

// dense 6 x 6 solve S xb = rhs, S packed lower-triangular (i (i + 1) / 2 + j), L'DL, one thread
__device__ __forceinline__ void tree_root_solve(real* S, real* xb) {
  real dinv[6];
#pragma unroll
  for (int kk = 5; kk >= 0; kk--) {
    dinv[kk] = real(1.0) / S[kk * (kk + 1) / 2 + kk];
#pragma unroll
    for (int j = 0; j < kk; j++) {
      const real l = S[kk * (kk + 1) / 2 + j] * dinv[kk];
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

// Solves H x = x0 in place (x holds the right-hand side on entry), H in ancestor-sparse rows (destroyed: the eliminated rows and 1/D
// remain; L[k][a] = H[k][a] / H[k][k] is applied on the fly, never stored).
// Every warp eliminates the DoFs of its subtrees from the leaves towards the root; what they contribute to the root block /
// the root right-hand side is summed per warp, then thread 0 finishes the 6 x 6 root block and the warps substitute back.
// The tables live in global memory and the 5 resident flies of an SM leave ~28 KB of L1, so every table access on the serial path
// of these loops costs an L2 round trip: one 8-int descriptor per DoF {k, row start, m, pair-list start, descendant-list start,
// number of descendants}, fetched one DoF ahead; the pair and descendant lists are streamed past L1 (__ldcs).
__device__ __forceinline__ void tree_factor_solve(const TP& p, real* sm, int tid) {
  const TreeDims& d = p.d; const int* it = p.it;
  const int lane = tid & 31, w = tid >> 5;
  real* H = sm + d.m_H; real* x = sm + d.m_x; real* dinv = sm + d.m_dinv;
  real* accS = sm + d.m_accS + w * 24; real* rb = sm + d.m_rb + w * 8;
  const int* rowadr = it + d.i_rowadr; const int* col = it + d.i_col;
  const int* pairs = it + d.i_pair; const int* desc = it + d.i_desc;
  const int nHa = d.nHa, wofs = w * 24;
  if (lane < 24) accS[lane] = real(0.);
  if (lane < 8) rb[lane] = real(0.);
  __syncwarp(NMF_FULL);
  const int k0 = it[d.i_wk_adr + w], k1 = it[d.i_wk_adr + w + 1];
  const int* kd = it + d.i_wk;
  int nk = 0, nr0 = 0, nm = 0, ne0 = 0, na = 0;
  if (k0 < k1) { nk = kd[8 * k0]; nr0 = kd[8 * k0 + 1]; nm = kd[8 * k0 + 2]; ne0 = kd[8 * k0 + 3]; na = lane < nm ? col[nr0 + lane + 1] : 0; }
  for (int idx = k0; idx < k1; idx++) {
    const int k = nk, r0 = nr0, m = nm, e0 = ne0, e1 = e0 + m * (m + 1) / 2, a0 = na;     // m proper ancestors a_1 .. a_m
    if (idx + 1 < k1) { nk = kd[8 * idx + 8]; nr0 = kd[8 * idx + 9]; nm = kd[8 * idx + 10]; ne0 = kd[8 * idx + 11]; na = lane < nm ? col[nr0 + lane + 1] : 0; }
    const real ik = real(1.) / H[r0], xk = x[k];
    // right-hand side (L^-T rides along): x[a_p] -= L[k][a_p] x[k]
    if (lane < m) { const real lx = (H[r0 + lane + 1] * ik) * xk; if (a0 >= TREE_NROOT) x[a0] -= lx; else rb[a0] += lx; }
    for (int pp = lane + 33; pp <= m; pp += 32) { const int a = col[r0 + pp]; const real lx = (H[r0 + pp] * ik) * xk; if (a >= TREE_NROOT) x[a] -= lx; else rb[a] += lx; }
    // H[a_p][a_q] -= (H[k][a_p] / H[k][k]) H[k][a_q] over all pairs p <= q, one pair per lane (distinct targets, none in row k; the
    // root block's pairs land in this warp's accumulator behind H); operands of four pairs are in flight before the first store
    for (int e = e0 + lane; e < e1; e += 128) {
      int pk[4], ti[4];
      real hp[4], hq[4], ht[4];
#pragma unroll
      for (int j = 0; j < 4; j++) pk[j] = e + 32 * j < e1 ? __ldcs(pairs + e + 32 * j) : 0;      // 0 = (target 0, p 0, q 0): a harmless dummy
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int tgt = pk[j] & 0xffff;
        ti[j] = tgt + (tgt >= nHa ? wofs : 0);
        hp[j] = H[r0 + ((pk[j] >> 16) & 0xff)]; hq[j] = H[r0 + (pk[j] >> 24)]; ht[j] = H[ti[j]];
      }
#pragma unroll
      for (int j = 0; j < 4; j++) if (pk[j]) H[ti[j]] = ht[j] - (hp[j] * ik) * hq[j];
    }
    if (lane == 0) dinv[k] = ik;
    __syncwarp(NMF_FULL);
  }
  tree_sync();
  if (tid == 0) {
    real S[21], xb[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
#pragma unroll
      for (int j = 0; j <= i; j++) {
        real v = H[rowadr[i] + (i - j)];
        for (int ww = 0; ww < TREE_NW; ww++) v += sm[d.m_accS + ww * 24 + i * (i + 1) / 2 + j];
        S[i * (i + 1) / 2 + j] = v;
      }
      real r = x[i];
      for (int ww = 0; ww < TREE_NW; ww++) r -= sm[d.m_rb + ww * 8 + i];
      xb[i] = r;
    }
    tree_root_solve(S, xb);
#pragma unroll
    for (int i = 0; i < 6; i++) x[i] = xb[i];
  }
  tree_sync();
  // back-substitution  x[k] = (y[k] - sum_a H[k][a] x[a]) / H[k][k]: the root solution first (the last six entries of every row),
  // then the warp's DoFs in ascending order, each pushing its final value to its descendants (one descendant per lane)
  {
    real xb[6];
#pragma unroll
    for (int r = 0; r < 6; r++) xb[r] = x[r];
    for (int idx = k0 + lane; idx < k1; idx += 32) {
      const int k = kd[8 * idx], r0 = kd[8 * idx + 1], m = kd[8 * idx + 2];
      real t = x[k];
#pragma unroll
      for (int r = 0; r < 6; r++) t -= H[r0 + m - r] * xb[r];
      x[k] = t;
    }
  }
  __syncwarp(NMF_FULL);
  int nf0 = 0, nnf = 0;
  if (k0 < k1) { nk = kd[8 * k1 - 8]; nf0 = kd[8 * k1 - 4]; nnf = kd[8 * k1 - 3]; na = lane < nnf ? __ldcs(desc + nf0 + lane) : 0; }
  for (int idx = k1 - 1; idx >= k0; idx--) {
    const int k = nk, f0 = nf0, nf = nnf, pk0 = na;
    if (idx > k0) { nk = kd[8 * idx - 8]; nf0 = kd[8 * idx - 4]; nnf = kd[8 * idx - 3]; na = lane < nnf ? __ldcs(desc + nf0 + lane) : 0; }
    const real xk = x[k] * dinv[k];
    if (lane < nf) x[pk0 >> 16] -= H[pk0 & 0xffff] * xk;
    for (int e = lane + 32; e < nf; e += 32) { const int pk = __ldcs(desc + f0 + e); x[pk >> 16] -= H[pk & 0xffff] * xk; }
    __syncwarp(NMF_FULL);
  }
  for (int idx = k0 + lane; idx < k1; idx += 32) { const int k = kd[8 * idx]; x[k] *= dinv[k]; }
  __syncwarp(NMF_FULL);
}

// body accelerations generated by the DoF vector v:  out_b = out_parent + sum_j cdof_j v_j   (6 reals per body)
__device__ __forceinline__ void tree_dof_to_body(const TP& p, real* sm, int tid, const real* v, real* out) {
  const TreeDims& d = p.d; const int* it = p.it;
  const int lane = tid & 31, w = tid >> 5;
  const real* cdof = sm + d.m_cdof;
  if (tid < 6) { real s = real(0.); for (int k = 0; k < TREE_NROOT; k++) s += cdof[6 * k + tid] * v[k]; out[tid] = s; }
  tree_sync();
  for (int dp = 1; dp <= d.maxd; dp++) {
    const int a0 = it[d.i_wb_adr + w * (d.maxd + 1) + dp], a1 = it[d.i_wb_adr + w * (d.maxd + 1) + dp + 1];
    for (int idx = a0 + lane; idx < a1; idx += 32) {
      const int b = it[d.i_wb + idx], par = it[d.i_parent + b], adr = it[d.i_dofadr + b], nd = it[d.i_ndof + b];
      real s[6];
#pragma unroll
      for (int i = 0; i < 6; i++) s[i] = out[6 * par + i];
      for (int j = 0; j < nd; j++) {
        const real vj = v[adr + j];
#pragma unroll
        for (int i = 0; i < 6; i++) s[i] += cdof[6 * (adr + j) + i] * vj;
      }
#pragma unroll
      for (int i = 0; i < 6; i++) out[6 * b + i] = s[i];
    }
    __syncwarp(NMF_FULL);
  }
  tree_sync();
}

// subtree sums, leaves -> root, of two per-body arrays (n1 / n2 reals per body; n2 may be 0)
template <int n1, int n2>
__device__ __forceinline__ void tree_backward(const TP& p, int tid, real* a1, real* a2) {
  const TreeDims& d = p.d; const int* it = p.it;
  const int lane = tid & 31, w = tid >> 5;
  for (int dp = d.maxd - 1; dp >= 1; dp--) {
    const int b0 = it[d.i_wb_adr + w * (d.maxd + 1) + dp], b1 = it[d.i_wb_adr + w * (d.maxd + 1) + dp + 1];
    for (int idx = b0 + lane; idx < b1; idx += 32) {
      const int b = it[d.i_wb + idx], c0 = it[d.i_child_adr + b], c1 = it[d.i_child_adr + b + 1];
      for (int ci = c0; ci < c1; ci++) {
        const int c = it[d.i_child + ci];
#pragma unroll
        for (int n = 0; n < n1; n++) a1[n1 * b + n] += a1[n1 * c + n];
#pragma unroll
        for (int n = 0; n < n2; n++) a2[n2 * b + n] += a2[n2 * c + n];
      }
    }
    __syncwarp(NMF_FULL);
  }
  tree_sync();
  if (tid < n1 + n2) {
    real* a = tid < n1 ? a1 : a2; const int n = tid < n1 ? tid : tid - n1, st = tid < n1 ? n1 : n2;
    real s = a[n];
    const int c0 = it[d.i_child_adr], c1 = it[d.i_child_adr + 1];
    for (int ci = c0; ci < c1; ci++) s += a[st * it[d.i_child + ci] + n];
    a[n] = s;
  }
  tree_sync();
}

// ------------------------------------------------------------------ the step
__device__ __forceinline__ void tree_step_block(const TP& p, real* sm, const int fly) {
  const TreeDims& d = p.d;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int* it = p.it; const real* rt = p.rt;
  const int nb = d.nb, nv = d.nv, ng = d.ng, nslot = d.nslot, nu = d.nu_pos + d.nu_adh;
  real *st = sm + d.m_state, *xpos = sm + d.m_xpos, *xquat = sm + d.m_xquat, *cinert = sm + d.m_cinert, *crb = sm + d.m_crb;
  real *cdof = sm + d.m_cdof, *cvel = sm + d.m_cvel, *acc = sm + d.m_acc, *yb = sm + d.m_y, *Pb = sm + d.m_P;
  real *fs = sm + d.m_fs, *grad = sm + d.m_grad, *x = sm + d.m_x, *u = sm + d.m_u, *H = sm + d.m_H;
  real *s_red = sm + d.m_red, *misc = sm + d.m_misc;
  real* com = misc;                              // [0..2]
  int* s_fault = reinterpret_cast<int*>(misc + 4);
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(misc + 8);
  real* actf = misc + 16;                        // [nu] actuator forces of this step
  TCon* con = reinterpret_cast<TCon*>(sm + d.m_con);
  int* hullv = reinterpret_cast<int*>(sm + d.m_hullv);
  real* sw = sm + d.m_weld;                      // weld rows of the tethered world (thread 0 owns them; they only touch the root body)
  int parity = 0;
  real* qpos = st + d.s_qpos; real* qvel = st + d.s_qvel; real* qacc = st + d.s_warm; real* ctrl = st + d.s_ctrl;

  for (int g = tid; g < ng; g += TREE_CTA) hullv[g] = 0;
  // ---- state record: one TMA bulk copy; the f64 build widens it (and prefers its own full-precision copy where the float
  //      record still equals what the last f64 launch wrote, see load_record in nmf_step.cuh)
  {
    const float* src = p.state + (size_t)fly * d.s_stride;
    if (std::is_same<real, float>::value) tma_load_n(reinterpret_cast<float*>(st), src, d.s_stride, mbar, tid);
    else {
      float* stage = reinterpret_cast<float*>(sm + d.m_stage);
      tma_load_n(stage, src, d.s_stride, mbar, tid);
      tree_sync();
      if (p.state64) {
        const double* s64 = p.state64 + (size_t)fly * d.s_stride; const float* sh = p.shadow + (size_t)fly * d.s_stride;
        for (int i = tid; i < d.s_stride; i += TREE_CTA) st[i] = (stage[i] == sh[i]) ? (real)s64[i] : (real)stage[i];
      } else {
        for (int i = tid; i < d.s_stride; i += TREE_CTA) st[i] = (real)stage[i];
      }
    }
  }
  tree_sync();

  for (int step = 0; step < p.nsteps; step++) {
    if (p.act_table) {
      const float* row = p.act_table + ((size_t)fly * p.table_T + (size_t)((p.table_t0 + step) % p.table_T)) * p.table_cols;
      for (int i = tid; i < p.table_cols; i += TREE_CTA) ctrl[i] = row[i];
    }
    // ================================================================= A. kinematics
    if (tid == 0) {
      real qh[4] = {qpos[3], qpos[4], qpos[5], qpos[6]}; qnormalize(qh);
      xpos[0] = qpos[0]; xpos[1] = qpos[1]; xpos[2] = qpos[2];
      xquat[0] = qh[0]; xquat[1] = qh[1]; xquat[2] = qh[2]; xquat[3] = qh[3];
    }
    tree_sync();
    for (int dp = 1; dp <= d.maxd; dp++) {
      const int a0 = it[d.i_wb_adr + w * (d.maxd + 1) + dp], a1 = it[d.i_wb_adr + w * (d.maxd + 1) + dp + 1];
      for (int idx = a0 + lane; idx < a1; idx += 32) {
        const int b = it[d.i_wb + idx], par = it[d.i_parent + b], adr = it[d.i_dofadr + b], nd = it[d.i_ndof + b];
        const real* br = rt + d.r_body + TR_BODY * b;
        real qp[4] = {xquat[4 * par], xquat[4 * par + 1], xquat[4 * par + 2], xquat[4 * par + 3]};
        real bp[3] = {br[0], br[1], br[2]}, bq[4] = {br[3], br[4], br[5], br[6]}, t3[3], q[4];
        qrot(qp, bp, t3); qmul(qp, bq, q);
        for (int j = 0; j < nd; j++) {
          const real* dr = rt + d.r_dof + TR_DOF * (adr + j);
          real ax[3] = {dr[0], dr[1], dr[2]}, wa[3]; qrot(q, ax, wa);
          cdof[6 * (adr + j)] = wa[0]; cdof[6 * (adr + j) + 1] = wa[1]; cdof[6 * (adr + j) + 2] = wa[2];
          real sn, cs; sincos_small(real(0.5) * qpos[1 + adr + j], &sn, &cs);
          real ql[4] = {cs, ax[0] * sn, ax[1] * sn, ax[2] * sn}, nq[4];
          qmul(q, ql, nq); q[0] = nq[0]; q[1] = nq[1]; q[2] = nq[2]; q[3] = nq[3];
        }
        qnormalize(q);
#pragma unroll
        for (int i = 0; i < 3; i++) xpos[3 * b + i] = xpos[3 * par + i] + t3[i];
#pragma unroll
        for (int i = 0; i < 4; i++) xquat[4 * b + i] = q[i];
      }
      __syncwarp(NMF_FULL);
    }
    tree_sync();
    // ---- subtree COM of the whole fly, inertias and DoF axes about it
    {
      real v[3] = {real(0.), real(0.), real(0.)};
      for (int b = tid; b < nb; b += TREE_CTA) {
        const real* br = rt + d.r_body + TR_BODY * b;
        real ip[3] = {br[7], br[8], br[9]}, t3[3]; qrot(xquat + 4 * b, ip, t3);
#pragma unroll
        for (int i = 0; i < 3; i++) v[i] += br[16] * (xpos[3 * b + i] + t3[i]);
      }
      tree_reduce<3>(v, s_red, parity, tid);
      if (tid == 0) { com[0] = v[0] * p.inv_total_mass; com[1] = v[1] * p.inv_total_mass; com[2] = v[2] * p.inv_total_mass; }
      tree_sync();
    }
    for (int b = tid; b < nb; b += TREE_CTA) {
      const real* br = rt + d.r_body + TR_BODY * b;
      real R[9]; q2mat(xquat + 4 * b, R);
      const real mass = br[16];
      real Ib[9] = {br[10], br[13], br[14], br[13], br[11], br[15], br[14], br[15], br[12]}, T[9], G[9], off[3];
#pragma unroll
      for (int i = 0; i < 3; i++) off[i] = xpos[3 * b + i] + R[3 * i] * br[7] + R[3 * i + 1] * br[8] + R[3 * i + 2] * br[9] - com[i];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) T[3 * i + j] = R[3 * i] * Ib[j] + R[3 * i + 1] * Ib[3 + j] + R[3 * i + 2] * Ib[6 + j];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = i; j < 3; j++) G[3 * i + j] = T[3 * i] * R[3 * j] + T[3 * i + 1] * R[3 * j + 1] + T[3 * i + 2] * R[3 * j + 2];
      const real o2 = dot3(off, off);
      real* ci = cinert + 10 * b;
      ci[0] = G[0] + mass * (o2 - off[0] * off[0]); ci[1] = G[4] + mass * (o2 - off[1] * off[1]); ci[2] = G[8] + mass * (o2 - off[2] * off[2]);
      ci[3] = G[1] - mass * off[0] * off[1]; ci[4] = G[2] - mass * off[0] * off[2]; ci[5] = G[5] - mass * off[1] * off[2];
      ci[6] = mass * off[0]; ci[7] = mass * off[1]; ci[8] = mass * off[2]; ci[9] = mass;
    }
    for (int k = tid; k < nv; k += TREE_CTA) {
      real* cd = cdof + 6 * k;
      if (k < 3) { cd[0] = cd[1] = cd[2] = real(0.); cd[3] = k == 0 ? real(1.) : real(0.); cd[4] = k == 1 ? real(1.) : real(0.); cd[5] = k == 2 ? real(1.) : real(0.); continue; }
      const int b = it[d.i_dof_body + k];
      real off[3] = {com[0] - xpos[3 * b], com[1] - xpos[3 * b + 1], com[2] - xpos[3 * b + 2]};
      if (k < TREE_NROOT) { real R[9]; q2mat(xquat, R); cd[0] = R[k - 3]; cd[1] = R[3 + k - 3]; cd[2] = R[6 + k - 3]; }
      real l[3]; cross3(cd, off, l);
      cd[3] = l[0]; cd[4] = l[1]; cd[5] = l[2];
    }
    tree_sync();
    // ================================================================= B. velocities, bias accelerations
    if (tid == 0) {
      real cv[6] = {real(0.), real(0.), real(0.), qvel[0], qvel[1], qvel[2]};     // translations first (their cdof_dot vanish)
      real ca[6] = {real(0.), real(0.), real(0.), -p.gx, -p.gy, -p.gz}, cv2[6];
#pragma unroll
      for (int i = 0; i < 6; i++) cv2[i] = cv[i];
      for (int k = 3; k < 6; k++) {
        real cdd[6]; cross_motion(cv, cdof + 6 * k, cdd);
#pragma unroll
        for (int i = 0; i < 6; i++) { ca[i] += cdd[i] * qvel[k]; cv2[i] += cdof[6 * k + i] * qvel[k]; }
      }
#pragma unroll
      for (int i = 0; i < 6; i++) { cvel[i] = cv2[i]; acc[i] = ca[i]; }
    }
    tree_sync();
    for (int dp = 1; dp <= d.maxd; dp++) {
      const int a0 = it[d.i_wb_adr + w * (d.maxd + 1) + dp], a1 = it[d.i_wb_adr + w * (d.maxd + 1) + dp + 1];
      for (int idx = a0 + lane; idx < a1; idx += 32) {
        const int b = it[d.i_wb + idx], par = it[d.i_parent + b], adr = it[d.i_dofadr + b], nd = it[d.i_ndof + b];
        real cv[6], ca[6];
#pragma unroll
        for (int i = 0; i < 6; i++) { cv[i] = cvel[6 * par + i]; ca[i] = acc[6 * par + i]; }
        for (int j = 0; j < nd; j++) {
          real cdd[6]; cross_motion(cv, cdof + 6 * (adr + j), cdd);
          const real qv = qvel[adr + j];
#pragma unroll
          for (int i = 0; i < 6; i++) { ca[i] += cdd[i] * qv; cv[i] += cdof[6 * (adr + j) + i] * qv; }
        }
#pragma unroll
        for (int i = 0; i < 6; i++) { cvel[6 * b + i] = cv[i]; acc[6 * b + i] = ca[i]; }
      }
      __syncwarp(NMF_FULL);
    }
    tree_sync();
    // ================================================================= C. collision, adhesion, smooth forces
    for (int g = tid; g < ng; g += TREE_CTA) {
      const int b = it[d.i_gbody + g];
      tree_collide(p, g, xpos + 3 * b, xquat + 4 * b, com, cvel + 6 * b, rt[d.r_body + TR_BODY * b + 17], con + g * nslot, hullv[g]);
    }
    tree_sync();
    for (int a = tid; a < d.nu_adh; a += TREE_CTA) {
      const int b = it[d.i_adh_body + a], g0 = it[d.i_bg_adr + b], g1 = it[d.i_bg_adr + b + 1];
      const real* ar = rt + d.r_adh + TR_ADH * a;
      real cnt = real(0.);
      for (int gi = g0; gi < g1; gi++) for (int s = 0; s < nslot; s++) cnt += con_on(con[it[d.i_bg + gi] * nslot + s].c);
      const real f = ar[0] * m_min(ar[2], m_max(ar[1], ctrl[d.nu_pos + a]));
      actf[d.nu_pos + a] = f;
      const real pull = cnt > real(0.) ? f / cnt : real(0.);
      for (int gi = g0; gi < g1; gi++) for (int s = 0; s < nslot; s++) con[it[d.i_bg + gi] * nslot + s].adh = pull;
    }
    tree_sync();
    for (int b = tid; b < nb; b += TREE_CTA) {
      real t1[6], t2[6], t3[6], W[6];
      mul_inert(cinert + 10 * b, acc + 6 * b, t1); mul_inert(cinert + 10 * b, cvel + 6 * b, t2); cross_force(cvel + 6 * b, t2, t3);
#pragma unroll
      for (int i = 0; i < 6; i++) W[i] = -(t1[i] + t3[i]);
      const int g0 = it[d.i_bg_adr + b], g1 = it[d.i_bg_adr + b + 1];
      for (int gi = g0; gi < g1; gi++) for (int s = 0; s < nslot; s++) { const TCon& c = con[it[d.i_bg + gi] * nslot + s]; adhesion_wrench(c.c, c.adh, W); }
#pragma unroll
      for (int i = 0; i < 6; i++) { yb[12 * b + i] = W[i]; yb[12 * b + 6 + i] = real(0.); }
#pragma unroll
      for (int i = 0; i < 10; i++) crb[10 * b + i] = cinert[10 * b + i];
    }
    if (p.out_energy && step == p.nsteps - 1) {
      // `energy` flag of the reference model (mujoco_globals.yaml:19): potential = -sum m g.x + joint springs, kinetic = 1/2 v'Mv
      real e[2] = {real(0.), real(0.)};
      for (int b = tid; b < nb; b += TREE_CTA) {
        const real* br = rt + d.r_body + TR_BODY * b;
        real ip[3] = {br[7], br[8], br[9]}, t3[3], t6[6]; qrot(xquat + 4 * b, ip, t3);
        e[0] -= br[16] * (p.gx * (xpos[3 * b] + t3[0]) + p.gy * (xpos[3 * b + 1] + t3[1]) + p.gz * (xpos[3 * b + 2] + t3[2]));
        mul_inert(cinert + 10 * b, cvel + 6 * b, t6); e[1] += real(0.5) * dot6(cvel + 6 * b, t6);
      }
      for (int k = tid; k < nv; k += TREE_CTA) {
        const real* dr = rt + d.r_dof + TR_DOF * k;
        if (k >= TREE_NROOT) { const real dq = qpos[1 + k] - dr[6]; e[0] += real(0.5) * dr[3] * dq * dq; }
        e[1] += real(0.5) * dr[5] * qvel[k] * qvel[k];
      }
      tree_reduce<2>(e, s_red, parity, tid);
      if (tid == 0) { p.out_energy[2 * (size_t)fly] = (float)e[0]; p.out_energy[2 * (size_t)fly + 1] = (float)e[1]; }
    }
    tree_sync();
    tree_backward<12, 10>(p, tid, yb, crb);       // (the upper six of y are not used yet; summing them costs less than a second pass shape)
    for (int k = tid; k < nv; k += TREE_CTA) {
      const int b = it[d.i_dof_body + k];
      const real* dr = rt + d.r_dof + TR_DOF * k;
      real f = dot6(cdof + 6 * k, yb + 12 * b);
      if (k >= TREE_NROOT) {
        const real q = qpos[1 + k], qv = qvel[k];
        f += -dr[3] * (q - dr[6]) - dr[4] * qv;
        const int ci = it[d.i_cidx + k];
        if (ci >= 0) {
          real af = dr[7] * ctrl[ci] - dr[7] * q - dr[8] * qv;
          af = m_min(dr[10], m_max(dr[9], af));
          actf[ci] = af; f += af;
        }
      }
      fs[k] = f;
    }
    // spatial acceleration of every body at the warm-start qacc, contact rows there
    tree_dof_to_body(p, sm, tid, qacc, acc);
    for (int g = tid; g < ng; g += TREE_CTA) {
      const int b = it[d.i_gbody + g];
      for (int s = 0; s < nslot; s++) {
        TCon& c = con[g * nslot + s];
        real ap[3]; project_point(c.c, acc + 6 * b, p.mu, ap);
        c.c.w[0] += ap[0]; c.c.w[1] += ap[1]; c.c.w[2] += ap[2];
      }
    }
    if (p.weld && tid == 0) weld_setup(p, sw, xquat, xpos, com, cvel, acc);     // six always-active rows on the root body
    tree_sync();

    // ================================================================= D. soft-contact solve (primal Newton, exact line search)
    int niter = 0, nls_total = 0, nchanged_last = 0, fault = 0;
    bool explicit_f = false;         // after noslip the contact forces are explicit (no longer a function of the rows)
    real* ns = sm + d.m_ns; int* nsrank = reinterpret_cast<int*>(sm + d.m_nsrank); int* nsidx = reinterpret_cast<int*>(ns + TNS_IDX);
    for (int iter = 0;; iter++) {
      const bool euler = iter > 0 && (nchanged_last == 0 || iter >= p.max_newton);
      if (euler && nchanged_last != 0) fault |= ST_NEWTON_CAP;
      if (euler && p.noslip_iterations > 0 && (ng > 0 || p.weld)) {
        // =================================================================
        // noslip post-solver of the reference's CPU path (mujoco_globals.yaml:15; [PRIOR] mj_solNoSlip), as in nmf_step.cuh: projected
        // Gauss-Seidel on the friction dimensions of the dual problem without the regulariser.  In the basis (n, mu t1, mu t2) of a
        // contact a pair of opposing pyramid edges is one tangential component g with |g| <= the normal force the pair carries;
        //     g <- clamp(g - h / B_gg),   h = jar_t(qacc) + B_tt (g - g_newton),   B_tt = E_t M^-1 E_t'  (2C x 2C, C <= 48).
        // B_tt is built column by column -- a unit wrench at the contact, subtree sums, a solve with the plain inertia matrix in the
        // ancestor-sparse storage, body accelerations, projection on every contact's tangents -- the sweeps run on it in shared memory
        // in the oracle's contact order (geom, then slot), and qacc moves by M^-1 E_t' (g - g_newton).  (Tethered world: see below.)
        // =================================================================
        real* Ss = cvel;
        // x = M^-1 J'(body wrenches in yb[12 b + 6 ..]); afterwards Ss holds the body accelerations of x
        auto m_solve = [&]() {
          tree_backward<12, 0>(p, tid, yb, Pb);
          for (int k = tid; k < nv; k += TREE_CTA) {
            const int b = it[d.i_dof_body + k];
            x[k] = dot6(cdof + 6 * k, yb + 12 * b + 6);
            real P[21]; expand_inert(crb + 10 * b, P);
            real uk[6]; sym6_mul(P, cdof + 6 * k, uk);
#pragma unroll
            for (int i = 0; i < 6; i++) u[6 * k + i] = uk[i];
          }
          tree_sync();
          for (int e = tid; e < d.nH; e += TREE_CTA) {
            const int rc = __ldcs(it + d.i_erow + e), row = rc & 0xffff, cl = rc >> 16;
            real v = dot6(cdof + 6 * cl, u + 6 * row);
            if (row == cl) v += rt[d.r_dof + TR_DOF * row + 5];
            H[e] = v;
          }
          tree_sync();
          tree_factor_solve(p, sm, tid);
          tree_sync();
          tree_dof_to_body(p, sm, tid, x, Ss);
        };
        if (p.weld) {
          // TetheredWorld: the six weld rows are equality rows, which noslip sweeps unclamped ( f_i -= residual_i / A_ii ), as in
          // nmf_step.cuh.  A = J_w M^-1 J_w' (6 x 6) column by column; the rows live on the root body and belong to thread 0.
          for (int j = 0; j < 6; j++) {
            for (int i = tid; i < 12 * nb; i += TREE_CTA) yb[i] = real(0.);
            tree_sync();
            if (tid == 0) {
              const real* r = sw + WL_R; const real* Gm = sw + WL_G; real* W = yb + 6;
              if (j < 3) { real e3[3] = {j == 0 ? real(1.) : real(0.), j == 1 ? real(1.) : real(0.), j == 2 ? real(1.) : real(0.)}, T[3]; cross3(r, e3, T);
                           W[0] = T[0]; W[1] = T[1]; W[2] = T[2]; W[3] = e3[0]; W[4] = e3[1]; W[5] = e3[2]; }
              else { W[0] = Gm[3 * (j - 3)]; W[1] = Gm[3 * (j - 3) + 1]; W[2] = Gm[3 * (j - 3) + 2]; }
            }
            tree_sync();
            m_solve();
            if (tid == 0) { real o6[6]; point_and_rot(sw, Ss, o6); for (int i = 0; i < 6; i++) ns[TNS_B + i * TNS_LD + j] = o6[i]; }
            tree_sync();
          }
          for (int i = tid; i < 12 * nb; i += TREE_CTA) yb[i] = real(0.);
          tree_sync();
          if (tid == 0) {
            real f[6], f0[6], jar[6], improvement0 = real(0.);
            for (int i = 0; i < 6; i++) {
              jar[i] = sw[WL_W + i] + sw[WL_C0 + i]; f0[i] = f[i] = -sw[WL_D + i] * jar[i];
              improvement0 += real(0.5) * f0[i] * f0[i] / sw[WL_D + i];
            }
            for (int sweep = 0; sweep < p.noslip_iterations; sweep++) {
              real improvement = sweep == 0 ? improvement0 : real(0.);
              for (int i = 0; i < 6; i++) {
                real res = jar[i];
                for (int q = 0; q < 6; q++) res += ns[TNS_B + i * TNS_LD + q] * (f[q] - f0[q]);
                const real Aii = ns[TNS_B + i * TNS_LD + i], old = f[i];
                f[i] = old - res / m_max(real(1e-15), Aii);
                const real dd = f[i] - old;
                real change = real(0.5) * dd * dd * Aii + dd * res;
                if (change > real(1e-10)) { f[i] = old; change = real(0.); }
                improvement -= change;
              }
              if (improvement * p.noslip_scale < p.noslip_tol) break;
            }
            real df[6];
            for (int i = 0; i < 6; i++) { sw[WL_F + i] = f[i]; df[i] = f[i] - f0[i]; }
            const real* r = sw + WL_R; const real* Gm = sw + WL_G; real* W = yb + 6;
            real T[3]; cross3(r, df, T);
            for (int i = 0; i < 3; i++) { W[i] = T[i] + Gm[i] * df[3] + Gm[3 + i] * df[4] + Gm[6 + i] * df[5]; W[3 + i] = df[i]; }
          }
          tree_sync();
          m_solve();
          for (int k = tid; k < nv; k += TREE_CTA) qacc[k] += x[k];
          for (int i = tid; i < 6 * nb; i += TREE_CTA) acc[i] += Ss[i];
          if (tid == 0) { real o6[6]; point_and_rot(sw, Ss, o6); for (int i = 0; i < 6; i++) sw[WL_W + i] += o6[i]; }
          tree_sync();
          explicit_f = true;
        }
        if (ng > 0 && tid == 0) {
          int C = 0;
          for (int i = 0; i < ng * nslot; i++) {
            nsrank[i] = -1;
            if (con[i].c.D > real(0.)) { if (C < TNS_MAXC) { nsidx[C] = i; nsrank[i] = C; } C++; }
          }
          reinterpret_cast<int*>(ns + TNS_MISC)[0] = C;
        }
        tree_sync();
        const int C = ng > 0 ? reinterpret_cast<int*>(ns + TNS_MISC)[0] : 0;
        if (C > TNS_MAXC) fault |= ST_NOSLIP_SKIP;
        else if (C > 0) {
          real c0r[1] = {real(0.)};
          if (tid < C) {
            const TCon& c = con[nsidx[tid]];
            real G[3], lim[2]; basis_forces(c.c, G, lim, c0r[0]);
            for (int q = 0; q < 3; q++) ns[TNS_GX + 3 * tid + q] = G[q];
            for (int q = 0; q < 2; q++) { const int i = 2 * tid + q; ns[TNS_G + i] = G[1 + q]; ns[TNS_G0 + i] = G[1 + q]; ns[TNS_LIM + i] = lim[q]; ns[TNS_JT + i] = c.c.w[1 + q]; }
          }
          tree_reduce<1>(c0r, s_red, parity, tid);
          for (int j = 0; j < 2 * C; j++) {
            for (int i = tid; i < 12 * nb; i += TREE_CTA) yb[i] = real(0.);
            tree_sync();
            if (tid == 0) {
              const int ci = nsidx[j >> 1], b = it[d.i_gbody + ci / nslot];
              const TCon& c = con[ci];
              real dd[3], T[3]; contact_tangent(c.c, j & 1, p.mu, dd); cross3(c.c.r, dd, T);
              real* W = yb + 12 * b + 6;
              W[0] = T[0]; W[1] = T[1]; W[2] = T[2]; W[3] = dd[0]; W[4] = dd[1]; W[5] = dd[2];
            }
            tree_sync();
            m_solve();
            if (tid < C) {
              const int ci = nsidx[tid], b = it[d.i_gbody + ci / nslot];
              real o3[3]; project_point(con[ci].c, Ss + 6 * b, p.mu, o3);
              ns[TNS_B + (2 * tid) * TNS_LD + j] = o3[1]; ns[TNS_B + (2 * tid + 1) * TNS_LD + j] = o3[2];
            }
            tree_sync();
          }
          // ---- the sweeps (serial Gauss-Seidel, one thread; 2C <= 48 unknowns)
          if (tid == 0) {
            const int n2 = 2 * C;
            for (int sweep = 0; sweep < p.noslip_iterations; sweep++) {
              real improvement = sweep == 0 ? c0r[0] : real(0.);
              for (int i = 0; i < n2; i++) {
                real h = ns[TNS_JT + i];
                for (int q = 0; q < n2; q++) h += ns[TNS_B + i * TNS_LD + q] * (ns[TNS_G + q] - ns[TNS_G0 + q]);
                const real Bii = ns[TNS_B + i * TNS_LD + i], gold = ns[TNS_G + i], l = ns[TNS_LIM + i];
                real gnew = real(0.);                           // K1 = 4 B_ii below MuJoCo's mjMINVAL: both edges get the mean
                if (real(4.) * Bii >= real(1e-15)) gnew = m_min(l, m_max(-l, gold - h / Bii));
                const real dy = real(0.5) * (gnew - gold);      // y = (f_j - f_j+1) / 2
                real change = real(2.) * Bii * dy * dy + real(2.) * h * dy;
                if (change > real(1e-10)) { gnew = gold; change = real(0.); }
                ns[TNS_G + i] = gnew; improvement -= change;
              }
              if (improvement * p.noslip_scale < p.noslip_tol) break;
            }
          }
          // ---- move: qacc += M^-1 E_t' (g - g_newton); rows and explicit forces follow
          for (int i = tid; i < 12 * nb; i += TREE_CTA) yb[i] = real(0.);
          tree_sync();
          if (tid == 0) {
            for (int r = 0; r < C; r++) {
              const int ci = nsidx[r], b = it[d.i_gbody + ci / nslot];
              real dG[3] = {real(0.), ns[TNS_G + 2 * r] - ns[TNS_GX + 3 * r + 1], ns[TNS_G + 2 * r + 1] - ns[TNS_GX + 3 * r + 2]};
              basis_wrench(con[ci].c, dG, p.mu, yb + 12 * b + 6, nullptr);
              ns[TNS_GX + 3 * r + 1] += dG[1]; ns[TNS_GX + 3 * r + 2] += dG[2];
            }
          }
          tree_sync();
          m_solve();
          for (int k = tid; k < nv; k += TREE_CTA) qacc[k] += x[k];
          for (int i = tid; i < 6 * nb; i += TREE_CTA) acc[i] += Ss[i];
          for (int g = tid; g < ng; g += TREE_CTA) {
            const int b = it[d.i_gbody + g];
            for (int s2 = 0; s2 < nslot; s2++) {
              TCon& c = con[g * nslot + s2];
              real o3[3]; project_point(c.c, Ss + 6 * b, p.mu, o3);
              c.c.w[0] += o3[0]; c.c.w[1] += o3[1]; c.c.w[2] += o3[2];
            }
          }
          tree_sync();
          explicit_f = true;
        }
      }
      // ---- forces of the contacts of every body, contact augmentation of its inertia
      for (int b = tid; b < nb; b += TREE_CTA) {
        real Wc[6] = {0, 0, 0, 0, 0, 0}, A[21];
#pragma unroll
        for (int i = 0; i < 21; i++) A[i] = real(0.);
        const int g0 = it[d.i_bg_adr + b], g1 = it[d.i_bg_adr + b + 1];
        for (int gi = g0; gi < g1; gi++) for (int s = 0; s < nslot; s++) {
          const int ci = it[d.i_bg + gi] * nslot + s;
          const TCon& c = con[ci];
          if (c.c.D > real(0.)) {
            if (explicit_f) basis_wrench(c.c, ns + TNS_GX + 3 * nsrank[ci], p.mu, Wc, nullptr);
            else if (euler) contact_forces<false>(c.c, p.mu, Wc, nullptr, nullptr);
            else contact_forces<true>(c.c, p.mu, Wc, A, nullptr);
          }
        }
        if (p.weld && b == 0) weld_forces(sw, Wc, A, explicit_f);
        real t6[6]; mul_inert(cinert + 10 * b, acc + 6 * b, t6);
#pragma unroll
        for (int i = 0; i < 6; i++) { yb[12 * b + i] = t6[i] - Wc[i]; yb[12 * b + 6 + i] = Wc[i]; }
#pragma unroll
        for (int i = 0; i < 21; i++) Pb[21 * b + i] = A[i];
      }
      tree_sync();
      if (euler) tree_backward<12, 0>(p, tid, yb, Pb); else tree_backward<12, 21>(p, tid, yb, Pb);
      // ---- gradient / right-hand side, u = (crb + A-hat) cdof of every DoF
      for (int k = tid; k < nv; k += TREE_CTA) {
        const int b = it[d.i_dof_body + k];
        const real* dr = rt + d.r_dof + TR_DOF * k;
        const real fc = dot6(cdof + 6 * k, yb + 12 * b + 6);
        const real g = dot6(cdof + 6 * k, yb + 12 * b) + dr[5] * qacc[k] - fs[k];
        const real gk = euler ? -(fs[k] + fc) : g;
        grad[k] = gk; x[k] = -gk;
        real P[21]; expand_inert(crb + 10 * b, P);
        if (!euler) {
#pragma unroll
          for (int i = 0; i < 21; i++) P[i] += Pb[21 * b + i];
        }
        real uk[6]; sym6_mul(P, cdof + 6 * k, uk);
#pragma unroll
        for (int i = 0; i < 6; i++) u[6 * k + i] = uk[i];
      }
      tree_sync();
      for (int e = tid; e < d.nH; e += TREE_CTA) {
        const int rc = __ldcs(it + d.i_erow + e), row = rc & 0xffff, cl = rc >> 16;
        real v = dot6(cdof + 6 * cl, u + 6 * row);
        if (row == cl) { const real* dr = rt + d.r_dof + TR_DOF * row; v += dr[5] + (euler ? p.dt * dr[4] : real(0.)); }
        H[e] = v;
      }
      tree_sync();
      tree_factor_solve(p, sm, tid);
      tree_sync();
      if (euler) { niter = iter; break; }
      // ---- body accelerations of the search direction (into the velocity array, which is free after the contact set-up)
      real* Ss = cvel;
      tree_dof_to_body(p, sm, tid, x, Ss);
      real red[5] = {real(0.), real(0.), real(0.), real(0.), real(0.)};   // s.g , s'Ms , d0 rows(0), d1 rows(0), |s|^2
      for (int b = tid; b < nb; b += TREE_CTA) { real t6[6]; mul_inert(cinert + 10 * b, Ss + 6 * b, t6); red[1] += dot6(Ss + 6 * b, t6); }
      for (int k = tid; k < nv; k += TREE_CTA) { const real xk = x[k]; red[0] += xk * grad[k]; red[1] += rt[d.r_dof + TR_DOF * k + 5] * xk * xk; red[4] += xk * xk; }
      for (int g = tid; g < ng; g += TREE_CTA) {
        const int b = it[d.i_gbody + g];
        for (int s = 0; s < nslot; s++) {
          TCon& c = con[g * nslot + s];
          real dummy = real(0.);
          project_point(c.c, Ss + 6 * b, p.mu, c.sv);
          if (c.c.D > real(0.)) ls_eval(c.c, c.sv, real(0.), red[2], red[3], dummy);
        }
      }
      if (p.weld && tid == 0) { point_and_rot(sw, Ss, sw + WL_SV); weld_ls(sw, real(0.), red[2], red[3]); }
      tree_reduce<5>(red, s_red, parity, tid);
      real alpha = real(0.);
      nchanged_last = 0;
      {
        const real q1 = red[0] - red[2], q2 = red[1];
        real d0 = red[0], d1 = q2 + red[3], lo = real(0.), hi = real(3.0e38);
        const int nls = (red[4] > real(1e-30) && d1 > real(0.)) ? p.max_ls : 0;
        for (int li = 0; li < nls; li++) {
          if (li > 0 && (m_abs(d0) <= Prec<real>::ls_rel * d1 * m_max(m_abs(alpha), Prec<real>::ls_amin) || (hi < real(1.0e38) && hi - lo <= Prec<real>::ls_bracket * hi))) break;
          if (li == nls - 1) fault |= ST_LS_CAP;
          if (d0 < real(0.)) lo = alpha; else hi = alpha;
          real nx = alpha - d0 / d1;
          if (nx <= lo || nx >= hi) nx = (hi > real(1.0e38)) ? real(2.) * m_max(alpha, real(1.)) : real(0.5) * (lo + hi);
          alpha = nx;
          real e[3] = {real(0.), real(0.), real(0.)};
          for (int g = tid; g < ng; g += TREE_CTA)
            for (int s = 0; s < nslot; s++) { const TCon& c = con[g * nslot + s]; if (c.c.D > real(0.)) ls_eval(c.c, c.sv, alpha, e[0], e[1], e[2]); }
          if (p.weld && tid == 0) weld_ls(sw, alpha, e[0], e[1]);
          tree_reduce<3>(e, s_red, parity, tid);
          d0 = q1 + alpha * q2 + e[0]; d1 = q2 + e[1];
          nchanged_last = (int)e[2];
          nls_total++;
        }
      }
      // ---- move
      for (int k = tid; k < nv; k += TREE_CTA) qacc[k] += alpha * x[k];
      for (int i = tid; i < 6 * nb; i += TREE_CTA) acc[i] += alpha * Ss[i];
      for (int g = tid; g < ng; g += TREE_CTA)
        for (int s = 0; s < nslot; s++) { TCon& c = con[g * nslot + s]; c.c.w[0] += alpha * c.sv[0]; c.c.w[1] += alpha * c.sv[1]; c.c.w[2] += alpha * c.sv[2]; }
      if (p.weld && tid == 0) for (int i = 0; i < 6; i++) sw[WL_W + i] += alpha * sw[WL_SV + i];
      tree_sync();
    }
    // x now holds the implicit-damping (Euler) acceleration

    // ---- optional outputs of this step (derived quantities belong to the pre-integration state, as in mj_step)
    if (step == p.nsteps - 1) {
      if (p.dbg && tid == 0) { float* dg = p.dbg + (size_t)fly * TDBG_STRIDE; dg[TDBG_NITER] = (float)niter; dg[TDBG_NLS] = (float)nls_total; real nc = real(0.); for (int i = 0; i < ng * nslot; i++) nc += con_on(con[i].c); dg[TDBG_NCON] = (float)nc; }
      if (p.out_actf) for (int i = tid; i < nu; i += TREE_CTA) p.out_actf[(size_t)fly * nu + i] = (float)actf[i];
      if (p.out_sensor && tid < d.nleg) {
        // per-leg contact sensor (world.py:311-331), reduce="netforce": found, force, torque, pos, normal, tangent
        const int l = tid;
        real F[3] = {0, 0, 0}, Pw[3] = {0, 0, 0}, Pp[3] = {0, 0, 0}, wsum = real(0.), cnt = real(0.);
        for (int g = 0; g < ng; g++) {
          if (it[d.i_leg + it[d.i_gbody + g]] != l) continue;
          for (int s = 0; s < nslot; s++) {
            const TCon& c = con[g * nslot + s];
            if (!(c.c.D > real(0.))) continue;
            real Wt[6] = {0, 0, 0, 0, 0, 0}, fn = real(0.);
            if (explicit_f) { const real* gx = ns + TNS_GX + 3 * nsrank[g * nslot + s]; basis_wrench(c.c, gx, p.mu, Wt, nullptr); fn = gx[0]; }
            else contact_forces<false>(c.c, p.mu, Wt, nullptr, &fn);
#pragma unroll
            for (int i = 0; i < 3; i++) { F[i] += Wt[3 + i]; Pw[i] += fn * (c.c.r[i] + com[i]); Pp[i] += c.c.r[i] + com[i]; }
            wsum += fn; cnt += real(1.);
          }
        }
        real P3[3] = {0, 0, 0}, T[3] = {0, 0, 0};
        if (cnt > real(0.)) for (int i = 0; i < 3; i++) P3[i] = wsum > real(1e-15) ? Pw[i] / wsum : Pp[i] / cnt;
        for (int g = 0; g < ng; g++) {
          if (it[d.i_leg + it[d.i_gbody + g]] != l) continue;
          for (int s = 0; s < nslot; s++) {
            const TCon& c = con[g * nslot + s];
            if (!(c.c.D > real(0.))) continue;
            real Wt[6] = {0, 0, 0, 0, 0, 0}, fn = real(0.);
            if (explicit_f) basis_wrench(c.c, ns + TNS_GX + 3 * nsrank[g * nslot + s], p.mu, Wt, nullptr);
            else contact_forces<false>(c.c, p.mu, Wt, nullptr, &fn);
            real rr[3] = {c.c.r[0] + com[0] - P3[0], c.c.r[1] + com[1] - P3[1], c.c.r[2] + com[2] - P3[2]}, tt[3];
            cross3(rr, Wt + 3, tt); T[0] += tt[0]; T[1] += tt[1]; T[2] += tt[2];
          }
        }
        float* o = p.out_sensor + ((size_t)fly * d.nleg + l) * 16;
        const bool found = cnt > real(0.);
        o[0] = (float)cnt;
        for (int i = 0; i < 3; i++) { o[1 + i] = found ? (float)-F[i] : 0.f; o[4 + i] = found ? (float)-T[i] : 0.f; o[7 + i] = (float)P3[i]; }
        o[10] = found ? 1.f : 0.f; o[11] = 0.f; o[12] = 0.f; o[13] = 0.f; o[14] = found ? 1.f : 0.f; o[15] = 0.f;
      }
      if (p.out_xpos || p.out_xquat) {
        for (int sgi = tid; sgi < d.nseg; sgi += TREE_CTA) {
          const float* tb = p.seg_tab + sgi * 8; const int b = __float_as_int(tb[0]);
          real lp[3] = {tb[1], tb[2], tb[3]}, lq[4] = {tb[4], tb[5], tb[6], tb[7]}, wv[3], wq[4];
          qrot(xquat + 4 * b, lp, wv); qmul(xquat + 4 * b, lq, wq);
          if (p.out_xpos) { float* o = p.out_xpos + ((size_t)fly * d.nseg + sgi) * 3; o[0] = (float)(xpos[3 * b] + wv[0]); o[1] = (float)(xpos[3 * b + 1] + wv[1]); o[2] = (float)(xpos[3 * b + 2] + wv[2]); }
          if (p.out_xquat) { float* o = p.out_xquat + ((size_t)fly * d.nseg + sgi) * 4; o[0] = (float)wq[0]; o[1] = (float)wq[1]; o[2] = (float)wq[2]; o[3] = (float)wq[3]; }
        }
      }
    }
    if (p.forward_only) break;

    // ---- advance: qvel += dt a' ; positions integrate with the NEW velocity ; qacc stays as next warm start
    if (tid == 0) *s_fault = 0;
    tree_sync();
    for (int k = tid; k < nv; k += TREE_CTA) {
      const real v = qvel[k] + p.dt * x[k];
      qvel[k] = v;
      if (!(m_abs(v) < real(3.0e38))) *s_fault = 1;
    }
    tree_sync();
    for (int k = tid + TREE_NROOT; k < nv; k += TREE_CTA) qpos[1 + k] += p.dt * qvel[k];
    if (tid == 0) {
      if (*s_fault) fault |= ST_NONFINITE;
      if (fault) st[d.s_time + 1] = (real)((int)st[d.s_time + 1] | fault);
      for (int i = 0; i < 3; i++) qpos[i] += p.dt * qvel[i];
      real wv[3] = {qvel[3], qvel[4], qvel[5]};
      const real n = m_sqrt(dot3(wv, wv));
      real q[4] = {qpos[3], qpos[4], qpos[5], qpos[6]};
      if (n > real(1e-15)) {
        real sn, cs; sincos_small(real(0.5) * p.dt * n, &sn, &cs);
        real dq[4] = {cs, wv[0] / n * sn, wv[1] / n * sn, wv[2] / n * sn}, nq[4];
        qmul(q, dq, nq); q[0] = nq[0]; q[1] = nq[1]; q[2] = nq[2]; q[3] = nq[3];
      }
      qnormalize(q);
      qpos[3] = q[0]; qpos[4] = q[1]; qpos[5] = q[2]; qpos[6] = q[3];
      st[d.s_time + 2] += real(1.);
      st[d.s_time] = st[d.s_time + 2] * p.dt;
    }
    tree_sync();
  }

  // ---- write the record back
  if (p.out_qpos) { float* o = p.out_qpos + (size_t)fly * d.nq; for (int i = tid; i < d.nq; i += TREE_CTA) o[i] = (float)qpos[i]; }
  if (!p.forward_only) {
    float* dst = p.state + (size_t)fly * d.s_stride;
    if (std::is_same<real, float>::value) tma_store_n(dst, reinterpret_cast<const float*>(st), d.s_stride, tid);
    else {
      float* stage = reinterpret_cast<float*>(sm + d.m_stage);
      tree_sync();
      for (int i = tid; i < d.s_stride; i += TREE_CTA) {
        stage[i] = (float)st[i];
        if (p.state64) { p.state64[(size_t)fly * d.s_stride + i] = (double)st[i]; p.shadow[(size_t)fly * d.s_stride + i] = stage[i]; }
      }
      tma_store_n(dst, stage, d.s_stride, tid);
    }
  }
}

}  // namespace NMF_NS
}  // namespace nmf
