// nmf_tree_host.h — host-side ingestion of a general-topology model for the tree kernels (plain C++; see nmf_tree_layout.h).
//
// Reads the same blob as nmf_host.h (flygym_b200/model.py::NMFModel.to_blob) but accepts any free root body carrying a tree
// of hinge-jointed bodies (0..3 hinges per body, anchored at the body origin, as the reference composes them:
// src/flygym/compose/fly.py:221-299), any number of capsule / convex-hull contact geoms per body, position actuators on any
// hinge DoF and adhesion actuators on any non-root body.
#pragma once
#include <algorithm>
#include <numeric>

#include "nmf_host.h"
#include "nmf_tree_layout.h"

namespace nmf {

struct TreeModel {
  std::vector<int32_t> itab;
  std::vector<double> rtab64, hull64;
  std::vector<float> rtab, hull;
  std::vector<int32_t> hull_nbr_adr, hull_nbr;
  std::vector<float> seg_tab, key_state;
  TreeParamsT<double> par64{};
  TreeParamsT<float> par{};
  int nu = 0;
  std::string err;

  static int align4(int v) { return (v + 3) & ~3; }
  static constexpr int WELD_REALS = 48;      // >= WL_COUNT of nmf_step.cuh (42: r, G, D, c0, w, sv and the explicit row forces after noslip)

  template <class real> static void plan_smem(TreeDims& d) {
    int o = 0;
    auto take = [&](int n) { const int at = o; o += align4(n); return at; };
    d.m_state = take(d.s_stride);
    d.m_stage = take(sizeof(real) == 8 ? (d.s_stride + 1) / 2 : 0);
    d.m_xpos = take(3 * d.nb); d.m_xquat = take(4 * d.nb); d.m_cinert = take(10 * d.nb); d.m_crb = take(10 * d.nb);
    d.m_cdof = take(6 * d.nv); d.m_cvel = take(6 * d.nb); d.m_acc = take(6 * d.nb); d.m_y = take(12 * d.nb); d.m_P = take(21 * d.nb);
    d.m_fs = take(d.nv); d.m_grad = take(d.nv); d.m_x = take(d.nv); d.m_u = take(6 * d.nv);
    d.m_H = take(d.nH); d.m_accS = take(TREE_NW * 24);      // the per-warp root-block accumulators sit right behind H (one index space)
    d.m_dinv = take(d.nv);
    d.m_con = take(d.ng * d.nslot * TCON_STRIDE);
    d.m_rb = take(TREE_NW * 8); d.m_red = take(2 * TREE_NW * 8);
    d.m_hullv = take(d.ng); d.m_misc = take(16 + d.nu_pos + d.nu_adh); d.m_weld = take(d.ng == 0 ? WELD_REALS : 0);   // (weld rows: only the tethered world, which has no contact geoms)
    d.m_ns = take(d.noslip ? TNS_TOTAL : 0); d.m_nsrank = take(d.noslip ? d.ng * d.nslot : 0);
    d.m_total = o;
  }

  bool build(const void* blob, size_t nbytes) {
    BlobView b{(const char*)blob, nbytes};
    if (!b.ok()) { err = "bad model blob"; return false; }
    const int32_t* dims = b.get<int32_t>("dims", nullptr, 10);
    if (!dims) { err = "blob has no dims"; return false; }
    const int nb = dims[0], nq = dims[1], nv = dims[2], nu_pos = dims[3], nu_adh = dims[4], ng = dims[5], nseg = dims[7], nleg = dims[8], nhv = dims[9];
    nu = nu_pos + nu_adh;
    if (nb < 1 || nb > TREE_MAXBODY || nv < TREE_NROOT || nv > TREE_MAXNV || nq != nv + 1 || ng < 0 || ng > TREE_MAXGEOM || nu_pos < 0 || nu_adh < 0 ||
        nu > TREE_MAXNU || nseg < 0 || nseg > 4096 || nleg < 0 || nleg > 16 || nhv < 0) { err = "model sizes outside what the tree kernels handle"; return false; }
    bool missing = false;
    auto D = [&](const char* n, int need) { const double* q = b.get<double>(n, nullptr, need); if (!q) { missing = true; err = std::string("blob section missing or too short: ") + n; } return q; };
    auto I = [&](const char* n, int need) { const int32_t* q = b.get<int32_t>(n, nullptr, need); if (!q) { missing = true; err = std::string("blob section missing or too short: ") + n; } return q; };
    const double *body_pos = D("body_pos", 3 * nb), *body_quat = D("body_quat", 4 * nb), *body_mass = D("body_mass", nb), *body_ipos = D("body_ipos", 3 * nb),
                 *body_iquat = D("body_iquat", 4 * nb), *body_inertia = D("body_inertia", 3 * nb), *invw = D("body_invweight0", 2 * nb), *dof_axis = D("dof_axis", 3 * nv),
                 *stiff = D("dof_stiffness", nv), *damp = D("dof_damping", nv), *arm = D("dof_armature", nv), *sref = D("dof_springref", nv),
                 *kp = D("act_kp", nu_pos), *kv = D("act_kv", nu_pos), *frc = D("act_frcrange", 2 * nu_pos), *again = D("adh_gain", nu_adh), *actrl = D("adh_ctrlrange", 2 * nu_adh),
                 *gpos = D("geom_pos", 3 * ng), *gquat = D("geom_quat", 4 * ng), *gsize = D("geom_size", 2 * ng), *hv = D("hull_vert", 3 * nhv), *segpos = D("seg_pos", 3 * nseg),
                 *segquat = D("seg_quat", 4 * nseg), *key_qpos = D("key_qpos", nq), *key_ctrl = D("key_ctrl", nu), *opt = D("opt", 11), *contact = D("contact", 10);
    const int32_t *body_parent = I("body_parent", nb), *dofadr = I("body_dofadr", nb), *dofnum = I("body_dofnum", nb), *body_leg = I("body_leg", nb),
                  *dof_body = I("dof_body", nv), *dof_parent = I("dof_parent", nv), *act_dof = I("act_dof", nu_pos), *adh_body = I("adh_body", nu_adh),
                  *geom_body = I("geom_body", ng), *geom_type = I("geom_type", ng), *gvadr = I("geom_vertadr", ng), *gvnum = I("geom_vertnum", ng), *seg_body = I("seg_body", nseg);
    if (missing) return false;
    // ---- topology checks: body 0 = free root (6 DoFs), parents before children, hinge DoFs contiguous and ordered like the bodies
    if (body_parent[0] >= 0 || dofnum[0] != TREE_NROOT || dofadr[0] != 0) { err = "tree kernels need a free root body (6 DoFs) as body 0"; return false; }
    std::vector<int> depth(nb, 0);
    int maxd = 0, next = TREE_NROOT;
    for (int bb = 1; bb < nb; bb++) {
      if (body_parent[bb] < 0 || body_parent[bb] >= bb) { err = "bodies must be ordered parents first"; return false; }
      if (dofnum[bb] < 0 || dofnum[bb] > 3 || dofadr[bb] != next) { err = "a body carries 0..3 hinge DoFs, numbered in body order"; return false; }
      next += dofnum[bb];
      depth[bb] = depth[body_parent[bb]] + 1; maxd = std::max(maxd, depth[bb]);
    }
    if (next != nv || maxd > TREE_MAXDEPTH) { err = "DoF count does not match the bodies"; return false; }
    for (int d = 0; d < nv; d++) {
      const int bb = dof_body[d];
      if (bb < 0 || bb >= nb || d < dofadr[bb] || d >= dofadr[bb] + dofnum[bb]) { err = "dof_body inconsistent"; return false; }
      int want;      // parent DoF: the previous DoF of the same body, else the last DoF of the nearest ancestor that has one
      if (d > dofadr[bb]) want = d - 1;
      else { int a = body_parent[bb]; while (a >= 0 && dofnum[a] == 0) a = body_parent[a]; want = a < 0 ? -1 : dofadr[a] + dofnum[a] - 1; }
      if (dof_parent[d] != want) { err = "dof_parent inconsistent with the body tree"; return false; }
    }
    for (int a = 0; a < nu_pos; a++) if (act_dof[a] < TREE_NROOT || act_dof[a] >= nv) { err = "actuator on a free-joint dof is not supported"; return false; }
    for (int a = 0; a < nu_adh; a++) if (adh_body[a] <= 0 || adh_body[a] >= nb) { err = "adhesion on the root body is not supported"; return false; }
    bool hulls = false;
    for (int g = 0; g < ng; g++) {
      if (geom_body[g] < 0 || geom_body[g] >= nb) { err = "geom_body out of range"; return false; }
      if (geom_type[g] != 0 && geom_type[g] != 1) { err = "unknown geom type"; return false; }
      if (gvadr[g] < 0 || gvnum[g] < 0 || gvadr[g] + gvnum[g] > nhv) { err = "hull vertex range out of bounds"; return false; }
      hulls = hulls || geom_type[g] == 1;
    }
    for (int sg = 0; sg < nseg; sg++) if (seg_body[sg] < 0 || seg_body[sg] >= nb) { err = "seg_body out of range"; return false; }
    TreeDims d{};
    d.noslip = (int)opt[8] > 0 ? 1 : 0;
    d.nb = nb; d.nq = nq; d.nv = nv; d.nu_pos = nu_pos; d.nu_adh = nu_adh; d.ng = ng; d.nseg = nseg; d.nleg = nleg; d.maxd = maxd;
    int nopt = 0; b.get<double>("opt", &nopt);
    const int multiccd = (hulls && nopt > 11 && opt[11] != 0.0) ? 1 : 0;
    d.nslot = multiccd ? 4 : 2;
    d.s_qpos = 0; d.s_qvel = align4(nq); d.s_warm = d.s_qvel + align4(nv); d.s_ctrl = d.s_warm + align4(nv); d.s_time = d.s_ctrl + align4(nu); d.s_stride = d.s_time + 4;

    // ---- subtrees of the root body -> warps (longest processing time first on the DoF count)
    std::vector<int> top(nb, -1);                 // root child a body hangs under
    for (int bb = 1; bb < nb; bb++) top[bb] = body_parent[bb] == 0 ? bb : top[body_parent[bb]];
    std::vector<int> cost(nb, 0), warp_of(nb, 0);
    for (int bb = 1; bb < nb; bb++) cost[top[bb]] += 1 + 2 * dofnum[bb];
    {
      std::vector<int> tops; for (int bb = 1; bb < nb; bb++) if (top[bb] == bb) tops.push_back(bb);
      std::stable_sort(tops.begin(), tops.end(), [&](int a, int c) { return cost[a] > cost[c]; });
      int load[TREE_NW] = {0};
      for (int tb : tops) { int w = 0; for (int q = 1; q < TREE_NW; q++) if (load[q] < load[w]) w = q; load[w] += cost[tb]; warp_of[tb] = w; }
      for (int bb = 1; bb < nb; bb++) warp_of[bb] = warp_of[top[bb]];
    }

    // ---- int table
    itab.clear();
    auto put = [&](const std::vector<int>& v) { const int at = (int)itab.size(); itab.insert(itab.end(), v.begin(), v.end()); while (itab.size() & 3) itab.push_back(0); return at; };
    auto putp = [&](const int32_t* p, int n) { return put(std::vector<int>(p, p + n)); };
    d.i_parent = putp(body_parent, nb); d.i_dofadr = putp(dofadr, nb); d.i_ndof = putp(dofnum, nb); d.i_leg = putp(body_leg, nb);
    {
      std::vector<int> adr(nb + 1, 0), lst;
      for (int bb = 0; bb < nb; bb++) { adr[bb] = (int)lst.size(); for (int c = bb + 1; c < nb; c++) if (body_parent[c] == bb) lst.push_back(c); }
      adr[nb] = (int)lst.size(); d.i_child_adr = put(adr); d.i_child = put(lst);
      std::vector<int> gadr(nb + 1, 0), gl;
      for (int bb = 0; bb < nb; bb++) { gadr[bb] = (int)gl.size(); for (int g = 0; g < ng; g++) if (geom_body[g] == bb) gl.push_back(g); }
      gadr[nb] = (int)gl.size(); d.i_bg_adr = put(gadr); d.i_bg = put(gl);
    }
    d.i_dof_body = putp(dof_body, nv);
    {
      std::vector<int> cidx(nv, -1);
      for (int a = 0; a < nu_pos; a++) { if (cidx[act_dof[a]] >= 0) { err = "two position actuators on one DoF"; return false; } cidx[act_dof[a]] = a; }
      d.i_cidx = put(cidx);
      std::vector<int> rowadr(nv + 1, 0), col, erow;
      for (int k = 0; k < nv; k++) { rowadr[k] = (int)col.size(); for (int a = k; a >= 0; a = dof_parent[a]) { col.push_back(a); erow.push_back(k); } }
      rowadr[nv] = (int)col.size(); d.nH = (int)col.size();
      d.i_rowadr = put(rowadr); d.i_col = put(col);
      for (size_t e = 0; e < col.size(); e++) erow[e] |= col[e] << 16;      // (row | column << 16) of every entry, streamed when H is formed
      d.i_erow = put(erow);
    }
    d.i_gbody = putp(geom_body, ng); d.i_gtype = putp(geom_type, ng); d.i_gvadr = putp(gvadr, ng); d.i_gvnum = putp(gvnum, ng);
    d.i_adh_body = putp(adh_body, nu_adh);
    {
      std::vector<int> adr, lst;
      for (int w = 0; w < TREE_NW; w++) for (int dp = 0; dp <= maxd; dp++) {
        adr.push_back((int)lst.size());
        if (dp >= 1) for (int bb = 1; bb < nb; bb++) if (warp_of[bb] == w && depth[bb] == dp) lst.push_back(bb);
      }
      adr.push_back((int)lst.size()); d.i_wb_adr = put(adr); d.i_wb = put(lst);
    }
    {
      // eliminating DoF k updates H[a_p][a_q] -= (H[k][a_p] / H[k][k]) H[k][a_q] for every pair p <= q of its proper ancestors a_1..a_m
      // (row position = distance up the ancestor chain); pairs inside the root block go to the warp's accumulator instead
      d.nHa = align4(d.nH);
      const int* rowadr = itab.data() + d.i_rowadr; const int* col = itab.data() + d.i_col;
      if (d.nHa + 24 > 0xffff) { err = "model too large for the packed elimination table"; return false; }
      std::vector<int> padr(nv + 1, 0), pl;
      for (int k = 0; k < nv; k++) {
        padr[k] = (int)pl.size();
        if (k < TREE_NROOT) continue;
        const int r0 = rowadr[k], m = rowadr[k + 1] - r0 - 1;
        if (m > 127) { err = "kinematic chain too deep"; return false; }
        for (int pp = 1; pp <= m; pp++) for (int q = pp; q <= m; q++) {
          const int a = col[r0 + pp];
          const int tgt = a >= TREE_NROOT ? rowadr[a] + (q - pp) : d.nHa + a * (a + 1) / 2 + (a - (q - pp));
          pl.push_back(tgt | (pp << 16) | (q << 24));
        }
      }
      // descendants of every non-root DoF k (the DoFs whose rows hold an entry in column k), for the push-style back-substitution
      std::vector<int> dadr(nv + 1, 0), dl;
      for (int k = 0; k < nv; k++) {
        dadr[k] = (int)dl.size();
        if (k < TREE_NROOT) continue;
        for (int j = k + 1; j < nv; j++) for (int e = rowadr[j] + 1; e < rowadr[j + 1]; e++) if (col[e] == k) dl.push_back(e | (j << 16));
      }
      dadr[nv] = (int)dl.size();
      std::vector<int> kadr, kl;
      for (int w = 0; w < TREE_NW; w++) {
        kadr.push_back((int)kl.size() / 8);
        for (int k = nv - 1; k >= TREE_NROOT; k--) if (warp_of[dof_body[k]] == w) {
          const int row[8] = {k, rowadr[k], rowadr[k + 1] - rowadr[k] - 1, padr[k], dadr[k], dadr[k + 1] - dadr[k], 0, 0};
          kl.insert(kl.end(), row, row + 8);
        }
      }
      kadr.push_back((int)kl.size() / 8); d.i_wk_adr = put(kadr); d.i_wk = put(kl);
      d.i_pair = put(pl); d.i_desc = put(dl);
    }
    d.i_total = (int)itab.size();

    // ---- real table
    rtab64.clear();
    double mtot = 0;
    d.r_body = (int)rtab64.size();
    for (int bb = 0; bb < nb; bb++) {
      mtot += body_mass[bb];
      double Rm[9], Ib[9]; q2mat_d(body_iquat + 4 * bb, Rm);
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += Rm[3 * i + k] * body_inertia[3 * bb + k] * Rm[3 * j + k]; Ib[3 * i + j] = s; }
      const double row[TR_BODY] = {body_pos[3 * bb], body_pos[3 * bb + 1], body_pos[3 * bb + 2], body_quat[4 * bb], body_quat[4 * bb + 1], body_quat[4 * bb + 2], body_quat[4 * bb + 3],
                                   body_ipos[3 * bb], body_ipos[3 * bb + 1], body_ipos[3 * bb + 2], Ib[0], Ib[4], Ib[8], Ib[1], Ib[2], Ib[5], body_mass[bb], invw[2 * bb]};
      rtab64.insert(rtab64.end(), row, row + TR_BODY);
    }
    d.r_dof = (int)rtab64.size();
    {
      std::vector<int> act_of(nv, -1);
      for (int a = 0; a < nu_pos; a++) act_of[act_dof[a]] = a;
      for (int k = 0; k < nv; k++) {
        const int a = act_of[k];
        const double row[TR_DOF] = {dof_axis[3 * k], dof_axis[3 * k + 1], dof_axis[3 * k + 2], stiff[k], damp[k], arm[k], sref[k],
                                    a >= 0 ? kp[a] : 0.0, a >= 0 ? kv[a] : 0.0, a >= 0 ? frc[2 * a] : 0.0, a >= 0 ? frc[2 * a + 1] : 0.0};
        rtab64.insert(rtab64.end(), row, row + TR_DOF);
      }
    }
    d.r_geom = (int)rtab64.size();
    for (int g = 0; g < ng; g++) {
      double Rm[9]; q2mat_d(gquat + 4 * g, Rm);
      const double row[TR_GEOM] = {gpos[3 * g], gpos[3 * g + 1], gpos[3 * g + 2], Rm[2], Rm[5], Rm[8], gsize[2 * g], gsize[2 * g + 1]};
      rtab64.insert(rtab64.end(), row, row + TR_GEOM);
    }
    d.r_adh = (int)rtab64.size();
    for (int a = 0; a < nu_adh; a++) { const double row[TR_ADH] = {again[a], actrl[2 * a], actrl[2 * a + 1]}; rtab64.insert(rtab64.end(), row, row + TR_ADH); }
    d.r_total = (int)rtab64.size();
    rtab.assign(rtab64.begin(), rtab64.end());
    hull64.assign((size_t)3 * (nhv > 0 ? nhv : 1), 0.0);
    for (int i = 0; i < 3 * nhv; i++) hull64[i] = hv[i];
    hull.assign(hull64.begin(), hull64.end());
    {
      int na = 0, nn = 0; const int32_t* adr = b.get<int32_t>("hull_nbr_adr", &na); const int32_t* nbp = b.get<int32_t>("hull_nbr", &nn);
      if (nhv > 0 && (!adr || !nbp || na != nhv + 1)) { err = "blob has no hull adjacency (re-bake the model)"; return false; }
      if (nhv > 0) {
        for (int v = 0; v < nhv; v++) if (adr[v] < 0 || adr[v] > adr[v + 1] || adr[v + 1] > nn) { err = "hull adjacency offsets out of bounds"; return false; }
        for (int g = 0; g < ng; g++) for (int v = gvadr[g]; v < gvadr[g] + gvnum[g]; v++)
          for (int e = adr[v]; e < adr[v + 1]; e++) if (nbp[e] < 0 || nbp[e] >= gvnum[g]) { err = "hull adjacency entry out of bounds"; return false; }
      }
      hull_nbr_adr.assign(adr ? adr : nullptr, adr ? adr + na : nullptr); hull_nbr.assign(nbp ? nbp : nullptr, nbp ? nbp + nn : nullptr);
      if (hull_nbr_adr.empty()) hull_nbr_adr.push_back(0);
      if (hull_nbr.empty()) hull_nbr.push_back(0);
    }
    seg_tab.assign((size_t)nseg * 8, 0.f);
    for (int s = 0; s < nseg; s++) {
      seg_tab[8 * s] = i2f(seg_body[s]);
      for (int i = 0; i < 3; i++) seg_tab[8 * s + 1 + i] = (float)segpos[3 * s + i];
      for (int i = 0; i < 4; i++) seg_tab[8 * s + 4 + i] = (float)segquat[4 * s + i];
    }
    key_state.assign(d.s_stride, 0.f);
    for (int i = 0; i < nq; i++) key_state[d.s_qpos + i] = (float)key_qpos[i];
    for (int i = 0; i < nu; i++) key_state[d.s_ctrl + i] = (float)key_ctrl[i];

    TreeParamsT<double>& P = par64;
    P = TreeParamsT<double>{};
    P.nsteps = 1;
    P.dt = opt[0]; P.gx = opt[1]; P.gy = opt[2]; P.gz = opt[3]; P.inv_total_mass = 1.0 / mtot; P.impratio = opt[10];
    P.mu = contact[0];
    const double tc = std::fmax(contact[1], 2 * opt[0]), dr = contact[2];
    auto clampimp = [](double v) { return std::fmin(0.9999, std::fmax(0.0001, v)); };
    const double dmax = clampimp(contact[4]);
    P.cK = 1.0 / (dmax * dmax * tc * tc * dr * dr); P.cB = 2.0 / (dmax * tc);
    P.solimp[0] = clampimp(contact[3]); P.solimp[1] = dmax; P.solimp[2] = std::fmax(0.0, contact[5]);
    P.solimp[3] = clampimp(contact[6]); P.solimp[4] = std::fmax(1.0, contact[7]);
    P.margin = contact[8] - contact[9];
    P.multiccd = multiccd;
    P.max_newton = (int)opt[4]; P.max_ls = (int)opt[6];
    P.noslip_iterations = d.noslip ? (int)opt[8] : 0;
    P.noslip_tol = 1e-6;                                  // MuJoCo default; the reference does not set it
    P.noslip_scale = 1.0 / ((opt[9] > 0 ? opt[9] : 1.0) * nv);
    if (P.max_newton < 1) P.max_newton = 100;
    if (P.max_ls < 1) P.max_ls = 50;
    {
      int nt = 0; const double* terr = b.get<double>("terrain", &nt);
      if (terr && nt >= 8 && terr[0] != 0.0) {
        if (terr[0] != 1.0 || !(terr[1] > 0) || !(terr[2] > 0) || !(terr[3] > 0) || !(terr[4] > 0) || terr[3] > 0.5 * terr[1] * (1 + 1e-9) || terr[4] > 0.5 * terr[2] * (1 + 1e-9)) {
          err = "unsupported terrain description"; return false;
        }
        if (hulls) { err = "terrain worlds need capsule collision geoms (simplify_geom=True)"; return false; }
        P.terrain = 1;
        for (int i = 0; i < 7; i++) P.terr[i] = terr[1 + i];
      }
    }
    {  // optional weld section (TetheredWorld): see flygym_b200/model.py WELD_FIELDS
      int nw = 0; const double* wd = b.get<double>("weld", &nw);
      if (wd && nw >= 18 && wd[0] != 0.0) {
        if (ng != 0) { err = "a tethered (welded) world cannot have ground-contact geoms"; return false; }
        P.weld = 1;
        for (int i = 0; i < 3; i++) P.weld_a[i] = wd[1 + i];
        for (int i = 0; i < 4; i++) P.weld_q[i] = wd[4 + i];
        const double wtc = std::fmax(wd[8], 2 * opt[0]), wdmax = clampimp(wd[11]);
        P.weld_K = 1.0 / (wdmax * wdmax * wtc * wtc * wd[9] * wd[9]); P.weld_B = 2.0 / (wdmax * wtc);
        P.weld_imp[0] = clampimp(wd[10]); P.weld_imp[1] = wdmax; P.weld_imp[2] = std::fmax(0.0, wd[12]);
        P.weld_imp[3] = clampimp(wd[13]); P.weld_imp[4] = std::fmax(1.0, wd[14]);
        P.weld_ts = wd[15]; P.weld_invw[0] = wd[16]; P.weld_invw[1] = wd[17];
      }
    }
    P.d = d; plan_smem<double>(P.d);
    TreeParamsT<float>& F = par;
    F = TreeParamsT<float>{};
    F.nsteps = 1; F.d = d; plan_smem<float>(F.d);
    F.dt = (float)P.dt; F.gx = (float)P.gx; F.gy = (float)P.gy; F.gz = (float)P.gz; F.inv_total_mass = (float)P.inv_total_mass;
    F.mu = (float)P.mu; F.cK = (float)P.cK; F.cB = (float)P.cB; F.margin = (float)P.margin; F.impratio = (float)P.impratio;
    for (int i = 0; i < 5; i++) F.solimp[i] = (float)P.solimp[i];
    for (int i = 0; i < 8; i++) F.terr[i] = (float)P.terr[i];
    F.max_newton = P.max_newton; F.max_ls = P.max_ls; F.multiccd = P.multiccd; F.terrain = P.terrain;
    F.noslip_iterations = P.noslip_iterations; F.noslip_tol = (float)P.noslip_tol; F.noslip_scale = (float)P.noslip_scale;
    F.weld = P.weld; F.weld_K = (float)P.weld_K; F.weld_B = (float)P.weld_B; F.weld_ts = (float)P.weld_ts;
    for (int i = 0; i < 3; i++) F.weld_a[i] = (float)P.weld_a[i];
    for (int i = 0; i < 4; i++) F.weld_q[i] = (float)P.weld_q[i];
    for (int i = 0; i < 5; i++) F.weld_imp[i] = (float)P.weld_imp[i];
    F.weld_invw[0] = (float)P.weld_invw[0]; F.weld_invw[1] = (float)P.weld_invw[1];
    return true;
  }
};

}  // namespace nmf
