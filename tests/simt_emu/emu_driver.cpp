// emu_driver.cpp — runs the *real* kernel source (flygym_b200/csrc/nmf_step.cuh) on the CPU
// through the SIMT emulator.  TEST INFRASTRUCTURE ONLY (see simt_emu.h).
#include "simt_emu.h"
#include "../../flygym_b200/csrc/nmf_host.h"
#include "../../flygym_b200/csrc/nmf_tree_host.h"
#include "../../flygym_b200/csrc/nmf_step_all.cuh"

static float g_sm[8 * (nmf::f32::SM_TOTAL + nmf::f32::NS_COUNT)];
static double g_sm64[nmf::f64::SM_TOTAL + nmf::f64::NS_COUNT];

extern "C" int emu_key_state(const void* blob, size_t nbytes, float* out) {
  nmf::HostModel hm;
  if (!hm.build(blob, nbytes)) { fprintf(stderr, "emu: %s\n", hm.err.c_str()); return -1; }
  memcpy(out, hm.key_state.data(), sizeof(float) * nmf::S_STRIDE);
  return 0;
}

extern "C" int emu_step(const void* blob, size_t nbytes, float* state, int n_flies, int nsteps, float* dbg, float* out_xpos,
                        float* out_xquat, float* out_actf, float* out_sensor, const float* act_table, int table_T, int table_t0, int table_cols,
                        int max_newton, int max_ls, int precision, int fpb, int sub_steps, float* out_energy) {
  nmf::HostModel hm;
  if (!hm.build(blob, nbytes)) { fprintf(stderr, "emu: %s\n", hm.err.c_str()); return -1; }
  if (fpb != 1 && fpb != 2 && fpb != 4 && fpb != 8) return -2;
  // sub_steps > 0: the launch is cut into (fly, sub_steps-step) items served from the work queue to ONE block of fpb slots
  std::vector<int> queue;
  int n_items = n_flies;
  if (sub_steps > 0) {
    const int nchunk = (nsteps + sub_steps - 1) / sub_steps;
    n_items = nchunk * n_flies;
    queue.assign(2 + (size_t)n_flies * (nchunk + 1), 0);
  }
  auto fill = [&](auto& q, const auto* role, const auto* hull) {
    q.state = state; q.role = role; q.hull = hull; q.seg_tab = hm.seg_tab.data(); q.hull_nbr_adr = hm.hull_nbr_adr.data(); q.hull_nbr = hm.hull_nbr.data();
    q.act_table = act_table; q.table_T = table_T; q.table_t0 = table_t0; q.table_cols = table_cols;
    q.out_xpos = out_xpos; q.out_xquat = out_xquat; q.out_actf = out_actf; q.out_sensor = out_sensor; q.out_energy = out_energy; q.dbg = dbg;
    q.n_flies = n_flies; q.nsteps = nsteps;
    if (max_newton > 0) q.max_newton = max_newton;
    if (max_ls > 0) q.max_ls = max_ls;
    q.queue = sub_steps > 0 ? queue.data() : nullptr; q.sub_steps = sub_steps > 0 ? sub_steps : nsteps; q.n_items = n_items;
  };
  if (sub_steps > 0) return -3;    // the work queue needs concurrently running blocks: exercised on the GPU only
  const int n_blocks = (n_flies + fpb - 1) / fpb;
  if (precision == 64) {   // the f64 instantiation of the same source (one fly per block)
    if (fpb != 1) return -2;
    nmf::StepParamsT<double> q = hm.par64;
    fill(q, hm.role64.data(), hm.hull64.data());
    for (int b = 0; b < n_blocks; b++) simt::run_block(nmf::CTA, b, n_blocks, [&]() {
      if (q.noslip_iterations > 0 && q.weld) nmf::f64::step_block<nmf::f64::W_TETHER, 1, true>(q, g_sm64, b, 0, q.nsteps, false);
      else if (q.noslip_iterations > 0) {    // the reference's CPU semantics: noslip post-solver
        if (q.multiccd) nmf::f64::step_block<nmf::f64::W_MESH, 1, true>(q, g_sm64, b, 0, q.nsteps, false);
        else if (q.terrain) nmf::f64::step_block<nmf::f64::W_TERRAIN, 1, true>(q, g_sm64, b, 0, q.nsteps, false);
        else nmf::f64::step_block<nmf::f64::W_FLAT, 1, true>(q, g_sm64, b, 0, q.nsteps, false);
      }
      else if (q.weld) nmf::f64::step_block<nmf::f64::W_TETHER>(q, g_sm64, b, 0, q.nsteps, false);
      else if (q.multiccd) nmf::f64::step_block<nmf::f64::W_MESH>(q, g_sm64, b, 0, q.nsteps, false);
      else if (q.terrain) nmf::f64::step_block<nmf::f64::W_TERRAIN>(q, g_sm64, b, 0, q.nsteps, false);
      else nmf::f64::step_block<nmf::f64::W_FLAT>(q, g_sm64, b, 0, q.nsteps, false); });
    return 0;
  }
  nmf::StepParams p = hm.par;
  fill(p, hm.role.data(), hm.hull.data());
  for (int b = 0; b < n_blocks; b++) simt::run_block(nmf::CTA * fpb, b, n_blocks, [&]() {
    const int slot = threadIdx.x / nmf::CTA; int f = b * fpb + slot; if (f >= n_flies) f = -1;
    float* sm = g_sm + slot * nmf::f32::SM_TOTAL;
#define EMU_RUN(W) switch (fpb) { case 1: nmf::f32::step_block<W, 1>(p, sm, f, 0, p.nsteps, false); break; case 2: nmf::f32::step_block<W, 2>(p, sm, f, 0, p.nsteps, false); break; \
                                  case 4: nmf::f32::step_block<W, 4>(p, sm, f, 0, p.nsteps, false); break; default: nmf::f32::step_block<W, 8>(p, sm, f, 0, p.nsteps, false); }
    if (p.noslip_iterations > 0 && !p.weld && !p.multiccd && !p.terrain) nmf::f32::step_block<nmf::f32::W_FLAT, 1, true>(p, sm, f, 0, p.nsteps, false);
    else if (p.weld) nmf::f32::step_block<nmf::f32::W_TETHER, 1>(p, sm, f, 0, p.nsteps, false);
    else if (p.multiccd) { EMU_RUN(nmf::f32::W_MESH) }
    else if (p.terrain) { EMU_RUN(nmf::f32::W_TERRAIN) }
    else { EMU_RUN(nmf::f32::W_FLAT) }
#undef EMU_RUN
  });
  return 0;
}

// ------------------------------------------------------------------ general-topology (tree) kernels, nmf_tree.cuh
extern "C" int emu_tree_info(const void* blob, size_t nbytes, int* out /* s_stride, s_qpos, s_qvel, s_warm, s_ctrl, s_time, nq, nv, nu, nseg, nleg, smem_f32_bytes, smem_f64_bytes, nH */) {
  nmf::TreeModel tm;
  if (!tm.build(blob, nbytes)) { fprintf(stderr, "emu: %s\n", tm.err.c_str()); return -1; }
  const nmf::TreeDims& d = tm.par.d;
  const int v[14] = {d.s_stride, d.s_qpos, d.s_qvel, d.s_warm, d.s_ctrl, d.s_time, d.nq, d.nv, d.nu_pos + d.nu_adh, d.nseg, d.nleg,
                     d.m_total * 4, tm.par64.d.m_total * 8, d.nH};
  memcpy(out, v, sizeof v);
  return 0;
}
extern "C" int emu_tree_key_state(const void* blob, size_t nbytes, float* out) {
  nmf::TreeModel tm;
  if (!tm.build(blob, nbytes)) { fprintf(stderr, "emu: %s\n", tm.err.c_str()); return -1; }
  memcpy(out, tm.key_state.data(), sizeof(float) * tm.key_state.size());
  return 0;
}
extern "C" int emu_tree_step(const void* blob, size_t nbytes, float* state, int n_flies, int nsteps, float* dbg, float* out_xpos, float* out_xquat,
                             float* out_actf, float* out_sensor, const float* act_table, int table_T, int table_t0, int table_cols,
                             int max_newton, int max_ls, int precision, float* out_energy, int forward_only) {
  nmf::TreeModel tm;
  if (!tm.build(blob, nbytes)) { fprintf(stderr, "emu: %s\n", tm.err.c_str()); return -1; }
  auto fill = [&](auto& q, const auto* rt, const auto* hull) {
    q.state = state; q.it = tm.itab.data(); q.rt = rt; q.hull = hull; q.hull_nbr_adr = tm.hull_nbr_adr.data(); q.hull_nbr = tm.hull_nbr.data();
    q.seg_tab = tm.seg_tab.data(); q.act_table = act_table; q.table_T = table_T; q.table_t0 = table_t0; q.table_cols = table_cols;
    q.out_xpos = out_xpos; q.out_xquat = out_xquat; q.out_actf = out_actf; q.out_sensor = out_sensor; q.out_energy = out_energy; q.dbg = dbg;
    q.n_flies = n_flies; q.nsteps = nsteps; q.forward_only = forward_only;
    if (max_newton > 0) q.max_newton = max_newton;
    if (max_ls > 0) q.max_ls = max_ls;
  };
  if (precision == 64) {
    nmf::TreeParamsT<double> q = tm.par64;
    fill(q, tm.rtab64.data(), tm.hull64.data());
    std::vector<double> sm(q.d.m_total, 0.0);
    for (int b = 0; b < n_flies; b++) simt::run_block(nmf::TREE_CTA, b, n_flies, [&]() { nmf::f64::tree_step_block(q, sm.data(), b); });
    return 0;
  }
  nmf::TreeParamsT<float> q = tm.par;
  fill(q, tm.rtab.data(), tm.hull.data());
  std::vector<float> sm(q.d.m_total, 0.f);
  for (int b = 0; b < n_flies; b++) simt::run_block(nmf::TREE_CTA, b, n_flies, [&]() { nmf::f32::tree_step_block(q, sm.data(), b); });
  return 0;
}
