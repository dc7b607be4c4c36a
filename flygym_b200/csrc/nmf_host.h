// nmf_host.h — host-side model ingestion for the sm_100a step path (plain C++).
//
// Parses the baked-model blob (flygym_b200/model.py::NMFModel.to_blob), checks that
// the model has the topology the kernels are written for and flattens it into the
// per-thread "role table" (nmf_layout.h) plus the scalar step parameters.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "nmf_layout.h"

namespace nmf {

struct BlobView {
  const char* p = nullptr; size_t n = 0;
  int32_t nsec() const { int32_t v; memcpy(&v, p + 12, 4); return v; }
  // header + section table inside the buffer, every section's payload inside it too (a truncated or corrupt blob is rejected
  // before anything is dereferenced)
  bool ok() const {
    if (!p || n < 16 || memcmp(p, "NMFB200", 8) != 0) return false;
    const int32_t ns = nsec();
    if (ns < 0 || ns > 4096 || 16 + (size_t)40 * ns > n) return false;
    for (int i = 0; i < ns; i++) {
      const char* e = p + 16 + 40 * i;
      int32_t dt, cnt; int64_t off; memcpy(&dt, e + 24, 4); memcpy(&cnt, e + 28, 4); memcpy(&off, e + 32, 8);
      if ((dt != 0 && dt != 1) || cnt < 0 || off < 0 || (off & 3)) return false;
      const size_t esz = dt == 0 ? 8 : 4;
      if ((size_t)off > n || (size_t)cnt > (n - (size_t)off) / esz) return false;
    }
    return true;
  }
  // section `name` with element size sizeof(T); nullptr when absent, of the wrong type, or shorter than `need` elements
  template <typename T> const T* get(const char* name, int* count = nullptr, int need = 0) const {
    const int32_t ns = nsec();
    for (int i = 0; i < ns; i++) {
      const char* e = p + 16 + 40 * i; char nm[25] = {0}; memcpy(nm, e, 24);
      int32_t dt, cnt; int64_t off; memcpy(&dt, e + 24, 4); memcpy(&cnt, e + 28, 4); memcpy(&off, e + 32, 8);
      if (!strcmp(nm, name)) {
        if ((dt == 0 ? 8 : 4) != (int)sizeof(T) || cnt < need) return nullptr;
        if (count) *count = cnt;
        return reinterpret_cast<const T*>(p + off);
      }
    }
    return nullptr;
  }
};

inline void qmul_d(const double* a, const double* b, double* r) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
inline void q2mat_d(const double* q, double* m) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = 1 - 2 * (x * x + z * z); m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = 1 - 2 * (x * x + y * y);
}
inline float i2f(int v) { float f; memcpy(&f, &v, 4); return f; }
// integer fields of the role table: bit pattern in the float table, plain number in the double one (nmf_step_common.cuh role_int)
template <class real> inline real int_field(int v);
template <> inline float int_field<float>(int v) { return i2f(v); }
template <> inline double int_field<double>(int v) { return (double)v; }

// scalar parameters of the f32 kernels = the f64 ones narrowed (pointers stay null)
inline StepParams narrow(const StepParamsT<double>& d) {
  StepParams f{};
  f.n_flies = d.n_flies; f.nsteps = d.nsteps; f.table_T = d.table_T; f.table_t0 = d.table_t0; f.table_cols = d.table_cols;
  f.forward_only = d.forward_only; f.nu_pos = d.nu_pos; f.nu_adh = d.nu_adh; f.nseg = d.nseg; f.nhubgeom = d.nhubgeom;
  f.dt = (float)d.dt; f.gx = (float)d.gx; f.gy = (float)d.gy; f.gz = (float)d.gz; f.inv_total_mass = (float)d.inv_total_mass;
  f.mu = (float)d.mu; f.cK = (float)d.cK; f.cB = (float)d.cB; f.margin = (float)d.margin; f.impratio = (float)d.impratio;
  for (int i = 0; i < 5; i++) { f.solimp[i] = (float)d.solimp[i]; f.weld_imp[i] = (float)d.weld_imp[i]; }
  f.max_newton = d.max_newton; f.max_ls = d.max_ls; f.terrain = d.terrain; f.weld = d.weld; f.multiccd = d.multiccd;
  f.noslip_iterations = d.noslip_iterations; f.noslip_tol = (float)d.noslip_tol; f.noslip_scale = (float)d.noslip_scale;
  for (int i = 0; i < 8; i++) f.terr[i] = (float)d.terr[i];
  for (int i = 0; i < 3; i++) f.weld_a[i] = (float)d.weld_a[i];
  for (int i = 0; i < 4; i++) f.weld_q[i] = (float)d.weld_q[i];
  f.weld_K = (float)d.weld_K; f.weld_B = (float)d.weld_B; f.weld_ts = (float)d.weld_ts;
  f.weld_invw[0] = (float)d.weld_invw[0]; f.weld_invw[1] = (float)d.weld_invw[1];
  f.sub_steps = d.sub_steps; f.n_items = d.n_items; f.n_chunks = d.n_chunks;
  for (int i = 0; i <= QUEUE_MAX_CHUNKS; i++) f.chunk_start[i] = d.chunk_start[i];
  return f;
}

struct HostModel {
  std::vector<float> role;     // RF_COUNT * CTA (f32 kernels); integer fields as bit patterns
  std::vector<float> hull;     // 3 * nhullvert
  std::vector<double> role64, hull64;   // the same tables for the f64 validation kernels (integer fields as numbers)
  std::vector<int32_t> hull_nbr_adr, hull_nbr;   // CSR adjacency of the hull vertices
  std::vector<float> seg_tab;  // nseg * 8
  std::vector<float> key_state;  // S_STRIDE, the neutral keyframe as a state record
  StepParams par{};            // pointer members left null
  StepParamsT<double> par64{};
  int nu = 0, nseg = 0;
  std::string err;

  bool build(const void* blob, size_t nbytes) {
    BlobView b{(const char*)blob, nbytes};
    if (!b.ok()) { err = "bad model blob"; return false; }
    const int32_t* dims = b.get<int32_t>("dims", nullptr, 10);
    if (!dims) { err = "blob has no dims"; return false; }
    const int nbody = dims[0], nq = dims[1], nv = dims[2], nu_pos = dims[3], nu_adh = dims[4], ngeom = dims[5];
    nseg = dims[7]; const int nleg = dims[8], nhv = dims[9];
    nu = nu_pos + nu_adh;
    if (nbody != 1 + NLEG * NLINK || nv != NV || nq != NQ || nleg != NLEG) { err = "unsupported topology: need hub + 6 legs x 8 links, nv=72"; return false; }
    if (nu_pos < 0 || nu_adh < 0 || nu > MAXU) { err = "too many actuators"; return false; }
    if (ngeom < 0 || nseg < 0 || nhv < 0 || nseg > 4096 || ngeom > CTA) { err = "bad dims"; return false; }
    bool missing = false;
    auto D = [&](const char* n, int need) { const double* q = b.get<double>(n, nullptr, need); if (!q) { missing = true; err = std::string("blob section missing or too short: ") + n; } return q; };
    auto I = [&](const char* n, int need) { const int32_t* q = b.get<int32_t>(n, nullptr, need); if (!q) { missing = true; err = std::string("blob section missing or too short: ") + n; } return q; };
    const double *body_pos = D("body_pos", 3 * nbody), *body_quat = D("body_quat", 4 * nbody), *body_mass = D("body_mass", nbody), *body_ipos = D("body_ipos", 3 * nbody),
                 *body_iquat = D("body_iquat", 4 * nbody), *body_inertia = D("body_inertia", 3 * nbody), *invw = D("body_invweight0", 2 * nbody), *dof_axis = D("dof_axis", 3 * nv),
                 *stiff = D("dof_stiffness", nv), *damp = D("dof_damping", nv), *arm = D("dof_armature", nv), *sref = D("dof_springref", nv),
                 *kp = D("act_kp", nu_pos), *kv = D("act_kv", nu_pos), *frc = D("act_frcrange", 2 * nu_pos), *again = D("adh_gain", nu_adh), *actrl = D("adh_ctrlrange", 2 * nu_adh),
                 *gpos = D("geom_pos", 3 * ngeom), *gquat = D("geom_quat", 4 * ngeom), *gsize = D("geom_size", 2 * ngeom), *hv = D("hull_vert", 3 * nhv), *segpos = D("seg_pos", 3 * nseg),
                 *segquat = D("seg_quat", 4 * nseg), *key_qpos = D("key_qpos", nq), *key_ctrl = D("key_ctrl", nu), *opt = D("opt", 11), *contact = D("contact", 10);
    const int32_t *body_parent = I("body_parent", nbody), *dofadr = I("body_dofadr", nbody), *dofnum = I("body_dofnum", nbody), *body_leg = I("body_leg", nbody),
                  *act_dof = I("act_dof", nu_pos), *adh_body = I("adh_body", nu_adh), *geom_body = I("geom_body", ngeom), *geom_type = I("geom_type", ngeom),
                  *gvadr = I("geom_vertadr", ngeom), *gvnum = I("geom_vertnum", ngeom), *seg_body = I("seg_body", nseg), *leg_root = I("leg_rootbody", nleg);
    if (missing) return false;
    for (int a = 0; a < nu_pos; a++) if (act_dof[a] < 6 || act_dof[a] >= nv) { err = "actuator on a free-joint dof is not supported"; return false; }
    for (int a = 0; a < nu_adh; a++) if (adh_body[a] <= 0 || adh_body[a] >= nbody) { err = "adhesion on the hub is not supported"; return false; }
    for (int g = 0; g < ngeom; g++) {
      if (geom_body[g] < 0 || geom_body[g] >= nbody) { err = "geom_body out of range"; return false; }
      if (gvadr[g] < 0 || gvnum[g] < 0 || gvadr[g] + gvnum[g] > nhv) { err = "hull vertex range out of bounds"; return false; }
    }
    for (int sg = 0; sg < nseg; sg++) if (seg_body[sg] < 0 || seg_body[sg] >= nbody) { err = "seg_body out of range"; return false; }
    static const int want_dofs[NLINK] = {3, 2, 1, 1, 1, 1, 1, 1};
    for (int l = 0; l < NLEG; l++) for (int k = 0; k < NLINK; k++) {
      int bb = 1 + l * NLINK + k;
      if (dofnum[bb] != want_dofs[k] || body_parent[bb] != (k == 0 ? 0 : bb - 1) || body_leg[bb] != l) { err = "unsupported leg chain layout"; return false; }
    }
    role64.assign((size_t)RF_COUNT * CTA, 0.0);
    std::vector<char> int_field_of(RF_COUNT, 0);
    auto set = [&](int field, int tid, double v) { role64[(size_t)field * CTA + tid] = v; };
    auto seti = [&](int field, int tid, int v) { role64[(size_t)field * CTA + tid] = (double)v; int_field_of[field] = 1; };
    auto lane_of_body = [&](int bb) { return bb == 0 ? NLEG * NLINK : bb - 1; };
    double mtot = 0;
    for (int bb = 0; bb < nbody; bb++) mtot += body_mass[bb];
    // body-level fields (hub replicated on all 16 hub lanes)
    for (int tid = 0; tid < CTA; tid++) {
      int bb = tid < NLEG * NLINK ? tid + 1 : 0;
      const bool hublane = tid >= NLEG * NLINK;   // hub chains: identity offset from the hub frame (seeded with the hub pose)
      for (int i = 0; i < 3; i++) set(RF_BPOS + i, tid, hublane ? 0.f : body_pos[3 * bb + i]);
      for (int i = 0; i < 4; i++) set(RF_BQUAT + i, tid, hublane ? (i == 0 ? 1.f : 0.f) : body_quat[4 * bb + i]);
      for (int i = 0; i < 3; i++) set(RF_IPOS + i, tid, body_ipos[3 * bb + i]);
      double Rm[9]; q2mat_d(body_iquat + 4 * bb, Rm);
      double Ib[9];
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += Rm[3 * i + k] * body_inertia[3 * bb + k] * Rm[3 * j + k]; Ib[3 * i + j] = s; }
      set(RF_IB + 0, tid, Ib[0]); set(RF_IB + 1, tid, Ib[4]); set(RF_IB + 2, tid, Ib[8]);
      set(RF_IB + 3, tid, Ib[1]); set(RF_IB + 4, tid, Ib[2]); set(RF_IB + 5, tid, Ib[5]);
      set(RF_MASS, tid, body_mass[bb]); set(RF_INVW, tid, invw[2 * bb]);
      if (hublane && tid != NLEG * NLINK) {       // only lane 48 carries the hub's mass / inertia
        set(RF_MASS, tid, 0.f);
        for (int i = 0; i < 6; i++) set(RF_IB + i, tid, 0.f);
      }
      seti(RF_GTYPE, tid, -1); seti(RF_ADH_CIDX, tid, -1); seti(RF_GIDX, tid, 1 << 20);
      for (int j = 0; j < 3; j++) seti(RF_CIDX + j, tid, -1);
      if (bb > 0) {
        seti(RF_NDOF, tid, dofnum[bb]); seti(RF_DOF0, tid, dofadr[bb]);
        for (int j = 0; j < dofnum[bb]; j++) {
          int d = dofadr[bb] + j;
          for (int i = 0; i < 3; i++) set(RF_AXIS + 3 * j + i, tid, dof_axis[3 * d + i]);
          set(RF_STIFF + j, tid, stiff[d]); set(RF_DAMP + j, tid, damp[d]); set(RF_ARM + j, tid, arm[d]);
          set(RF_SREF + j, tid, sref[d]);
        }
        bool sens = false;   // the leg sensor covers the subtree rooted at the most proximal contact segment
        int l = body_leg[bb]; if (l >= 0 && leg_root[l] >= 0 && bb >= leg_root[l]) sens = true;
        seti(RF_LEGSENSOR, tid, sens ? 1 : 0);
      } else { seti(RF_NDOF, tid, 0); seti(RF_DOF0, tid, 0); }
    }
    // armature / damping of the matrix columns each lane holds in the chain factorisation
    for (int tid = 0; tid < CTA; tid++) {
      const int g = tid / NLINK, t = tid % NLINK;
      if (g >= NLEG) { for (int s = 0; s < 3; s++) { set(RF_CARM + s, tid, 1.f); set(RF_CDMP + s, tid, 0.f); } continue; }
      const int lb = 6 + NLEGDOF * g; const int cols[3] = {t >= 6 ? lb + t - 6 : -1, lb + t + 2, lb + 10};
      for (int s = 0; s < 3; s++) { set(RF_CARM + s, tid, cols[s] >= 0 ? arm[cols[s]] : 0.f); set(RF_CDMP + s, tid, cols[s] >= 0 ? damp[cols[s]] : 0.f); }
    }
    for (int a = 0; a < nu_pos; a++) {
      int d = act_dof[a], bb = -1, j = 0;
      for (int c = 1; c < nbody; c++) if (d >= dofadr[c] && d < dofadr[c] + dofnum[c]) { bb = c; j = d - dofadr[c]; }
      if (bb < 0) { err = "actuator on a free-joint dof is not supported"; return false; }
      int tid = lane_of_body(bb);
      set(RF_KP + j, tid, kp[a]); set(RF_KV + j, tid, kv[a]); set(RF_FLO + j, tid, frc[2 * a]);
      set(RF_FHI + j, tid, frc[2 * a + 1]); seti(RF_CIDX + j, tid, a);
    }
    for (int a = 0; a < nu_adh; a++) {
      int bb = adh_body[a]; if (bb <= 0) { err = "adhesion on the hub is not supported"; return false; }
      int tid = lane_of_body(bb);
      set(RF_ADH_GAIN, tid, again[a]); set(RF_ADH_LO, tid, actrl[2 * a]); set(RF_ADH_HI, tid, actrl[2 * a + 1]);
      seti(RF_ADH_CIDX, tid, nu_pos + a);
    }
    int nhub = 0;
    std::vector<char> used(CTA, 0);
    for (int g = 0; g < ngeom; g++) {
      int bb = geom_body[g], tid;
      if (bb == 0) { if (nhub >= NHUBLANE) { err = "more than 16 contact geoms on the hub"; return false; } tid = NLEG * NLINK + nhub++; }
      else { tid = lane_of_body(bb); if (used[tid]) { err = "more than one contact geom on a leg body"; return false; } }
      used[tid] = 1;
      seti(RF_GTYPE, tid, geom_type[g]);
      double Rm[9]; q2mat_d(gquat + 4 * g, Rm);
      for (int i = 0; i < 3; i++) { set(RF_GPOS + i, tid, gpos[3 * g + i]); set(RF_GAXIS + i, tid, Rm[3 * i + 2]); }
      set(RF_GRAD, tid, gsize[2 * g]); set(RF_GHALF, tid, gsize[2 * g + 1]);
      seti(RF_GVADR, tid, gvadr[g]); seti(RF_GVNUM, tid, gvnum[g]); seti(RF_GIDX, tid, g);
    }
    hull64.assign((size_t)3 * (nhv > 0 ? nhv : 1), 0.0);
    for (int i = 0; i < 3 * nhv; i++) hull64[i] = hv[i];
    hull.assign(hull64.begin(), hull64.end());
    role.resize(role64.size());
    for (int f = 0; f < RF_COUNT; f++) for (int t = 0; t < CTA; t++) {
      const double v = role64[(size_t)f * CTA + t];
      role[(size_t)f * CTA + t] = int_field_of[f] ? i2f((int)v) : (float)v;
    }
    {
      int na = 0, nn = 0; const int32_t* adr = b.get<int32_t>("hull_nbr_adr", &na); const int32_t* nb = b.get<int32_t>("hull_nbr", &nn);
      if (nhv > 0 && (!adr || !nb || na != nhv + 1)) { err = "blob has no hull adjacency (re-bake the model)"; return false; }
      if (nhv > 0) {
        for (int v = 0; v < nhv; v++) if (adr[v] < 0 || adr[v] > adr[v + 1] || adr[v + 1] > nn) { err = "hull adjacency offsets out of bounds"; return false; }
        for (int g = 0; g < ngeom; g++) for (int v = gvadr[g]; v < gvadr[g] + gvnum[g]; v++)
          for (int e = adr[v]; e < adr[v + 1]; e++) if (nb[e] < 0 || nb[e] >= gvnum[g]) { err = "hull adjacency entry out of bounds"; return false; }
      }
      hull_nbr_adr.assign(adr ? adr : nullptr, adr ? adr + na : nullptr); hull_nbr.assign(nb ? nb : nullptr, nb ? nb + nn : nullptr);
      if (hull_nbr_adr.empty()) hull_nbr_adr.push_back(0);
      if (hull_nbr.empty()) hull_nbr.push_back(0);
    }
    seg_tab.assign((size_t)nseg * 8, 0.f);
    for (int s = 0; s < nseg; s++) {
      seg_tab[8 * s] = i2f(lane_of_body(seg_body[s]));
      for (int i = 0; i < 3; i++) seg_tab[8 * s + 1 + i] = (float)segpos[3 * s + i];
      for (int i = 0; i < 4; i++) seg_tab[8 * s + 4 + i] = (float)segquat[4 * s + i];
    }
    key_state.assign(S_STRIDE, 0.f);
    for (int i = 0; i < nq; i++) key_state[S_QPOS + i] = (float)key_qpos[i];
    for (int i = 0; i < nu; i++) key_state[S_CTRL + i] = (float)key_ctrl[i];
    StepParamsT<double>& P = par64;
    P = StepParamsT<double>{};
    P.nu_pos = nu_pos; P.nu_adh = nu_adh; P.nseg = nseg; P.nhubgeom = nhub;
    P.dt = opt[0]; P.gx = opt[1]; P.gy = opt[2]; P.gz = opt[3];
    P.inv_total_mass = (1.0 / mtot);
    P.impratio = opt[10];
    P.mu = contact[0];
    double tc = std::fmax(contact[1], 2 * opt[0]), dr = contact[2];
    auto clampimp = [](double v) { return std::fmin(0.9999, std::fmax(0.0001, v)); };
    double dmax = clampimp(contact[4]);
    P.cK = (1.0 / (dmax * dmax * tc * tc * dr * dr)); P.cB = (2.0 / (dmax * tc));
    P.solimp[0] = clampimp(contact[3]); P.solimp[1] = dmax; P.solimp[2] = std::fmax(0.0, contact[5]);
    P.solimp[3] = clampimp(contact[6]); P.solimp[4] = std::fmax(1.0, contact[7]);
    P.margin = (contact[8] - contact[9]);
    {   // `multiccd` flag of the reference model (mujoco_globals.yaml:18): opt[11] when the blob carries it; it only matters for hull geoms
      int nopt = 0; b.get<double>("opt", &nopt);
      bool hulls = false;
      for (int g = 0; g < ngeom; g++) hulls = hulls || geom_type[g] == 1;
      P.multiccd = (hulls && nopt > 11 && opt[11] != 0.0) ? 1 : 0;
    }
    P.noslip_iterations = (int)opt[8] > 0 ? (int)opt[8] : 0;
    P.noslip_tol = 1e-6;                                  // MuJoCo default; the reference does not set it
    P.noslip_scale = 1.0 / ((opt[9] > 0 ? opt[9] : 1.0) * NV);
    P.max_newton = (int)opt[4]; P.max_ls = (int)opt[6]; P.nsteps = 1;   // reference: iterations=100 (mujoco_globals.yaml:14), ls_iterations=50 (MuJoCo default)
    if (P.max_newton < 1) P.max_newton = 100;
    if (P.max_ls < 1) P.max_ls = 50;
    {  // optional weld section (TetheredWorld): see flygym_b200/model.py WELD_FIELDS
      int nw = 0; const double* wd = b.get<double>("weld", &nw);
      if (wd && nw >= 18 && wd[0] != 0.0) {
        if (ngeom != 0) { err = "a tethered (welded) world cannot have ground-contact geoms"; return false; }
        P.weld = 1;
        for (int i = 0; i < 3; i++) P.weld_a[i] = wd[1 + i];
        for (int i = 0; i < 4; i++) P.weld_q[i] = wd[4 + i];
        const double wtc = std::fmax(wd[8], 2 * opt[0]), wdmax = clampimp(wd[11]);
        P.weld_K = (1.0 / (wdmax * wdmax * wtc * wtc * wd[9] * wd[9])); P.weld_B = (2.0 / (wdmax * wtc));
        P.weld_imp[0] = clampimp(wd[10]); P.weld_imp[1] = wdmax; P.weld_imp[2] = std::fmax(0.0, wd[12]);
        P.weld_imp[3] = clampimp(wd[13]); P.weld_imp[4] = std::fmax(1.0, wd[14]);
        P.weld_ts = wd[15]; P.weld_invw[0] = wd[16]; P.weld_invw[1] = wd[17];
      }
    }
    {  // optional terrain section: {type, Px, Py, hx, hy, top_even, top_odd, z_floor}
      int nt = 0; const double* terr = b.get<double>("terrain", &nt);
      if (terr && nt >= 8 && terr[0] != 0.0) {
        if (terr[0] != 1.0 || !(terr[1] > 0) || !(terr[2] > 0) || !(terr[3] > 0) || !(terr[4] > 0) || terr[3] > 0.5 * terr[1] * (1 + 1e-9) || terr[4] > 0.5 * terr[2] * (1 + 1e-9)) {
          err = "unsupported terrain description"; return false;
        }
        for (int g = 0; g < ngeom; g++) if (geom_type[g] != 0) { err = "terrain worlds need capsule collision geoms (simplify_geom=True)"; return false; }
        P.terrain = 1;
        for (int i = 0; i < 7; i++) P.terr[i] = terr[1 + i];
      }
    }
    par = narrow(par64);
    (void)body_parent;
    return true;
  }
};

}  // namespace nmf
