// emu_driver.cpp — runs the *real* kernel source (flygym_b200/csrc/nmf_step.cuh) on the CPU
// through the SIMT emulator.  TEST INFRASTRUCTURE ONLY (see simt_emu.h).
#include "simt_emu.h"
#include "../../flygym_b200/csrc/nmf_host.h"
#include "../../flygym_b200/csrc/nmf_step_all.cuh"

static float g_sm[nmf::f32::SM_TOTAL];
static double g_sm64[nmf::f64::SM_TOTAL];

extern "C" int emu_key_state(const void* blob, size_t nbytes, float* out) {
  nmf::HostModel hm;
  if (!hm.build(blob, nbytes)) { fprintf(stderr, "emu: %s\n", hm.err.c_str()); return -1; }
  memcpy(out, hm.key_state.data(), sizeof(float) * nmf::S_STRIDE);
  return 0;
}

extern "C" int emu_step(const void* blob, size_t nbytes, float* state, int n_flies, int nsteps, float* dbg, float* out_xpos,
                        float* out_xquat, float* out_actf, float* out_sensor, const float* act_table, int table_T, int table_t0, int table_cols,
                        int max_newton, int max_ls, int precision) {
  nmf::HostModel hm;
  if (!hm.build(blob, nbytes)) { fprintf(stderr, "emu: %s\n", hm.err.c_str()); return -1; }
  nmf::StepParams p = hm.par;
  p.state = state; p.role = hm.role.data(); p.hull = hm.hull.data(); p.seg_tab = hm.seg_tab.data(); p.hull_nbr_adr = hm.hull_nbr_adr.data(); p.hull_nbr = hm.hull_nbr.data();
  p.act_table = act_table; p.table_T = table_T; p.table_t0 = table_t0; p.table_cols = table_cols;
  p.out_xpos = out_xpos; p.out_xquat = out_xquat; p.out_actf = out_actf; p.out_sensor = out_sensor; p.dbg = dbg;
  p.n_flies = n_flies; p.nsteps = nsteps;
  if (max_newton > 0) p.max_newton = max_newton;
  if (max_ls > 0) p.max_ls = max_ls;
  if (precision == 64) {   // the f64 instantiation of the same source
    nmf::StepParamsT<double> q = hm.par64;
    q.state = state; q.role = hm.role64.data(); q.hull = hm.hull64.data(); q.seg_tab = hm.seg_tab.data(); q.hull_nbr_adr = hm.hull_nbr_adr.data(); q.hull_nbr = hm.hull_nbr.data();
    q.act_table = act_table; q.table_T = table_T; q.table_t0 = table_t0; q.table_cols = table_cols;
    q.out_xpos = out_xpos; q.out_xquat = out_xquat; q.out_actf = out_actf; q.out_sensor = out_sensor; q.dbg = dbg;
    q.n_flies = n_flies; q.nsteps = nsteps;
    if (max_newton > 0) q.max_newton = max_newton;
    if (max_ls > 0) q.max_ls = max_ls;
    for (int f = 0; f < n_flies; f++) simt::run_block(nmf::CTA, f, n_flies, [&]() {
      if (q.weld) nmf::f64::step_block<nmf::f64::W_TETHER>(q, g_sm64, f, 0, q.nsteps, false);
      else if (q.terrain) nmf::f64::step_block<nmf::f64::W_TERRAIN>(q, g_sm64, f, 0, q.nsteps, false);
      else nmf::f64::step_block<nmf::f64::W_FLAT>(q, g_sm64, f, 0, q.nsteps, false); });
    return 0;
  }
  for (int f = 0; f < n_flies; f++) simt::run_block(nmf::CTA, f, n_flies, [&]() {
    if (p.weld) nmf::f32::step_block<nmf::f32::W_TETHER>(p, g_sm, f, 0, p.nsteps, false);
    else if (p.terrain) nmf::f32::step_block<nmf::f32::W_TERRAIN>(p, g_sm, f, 0, p.nsteps, false);
    else nmf::f32::step_block<nmf::f32::W_FLAT>(p, g_sm, f, 0, p.nsteps, false); });
  return 0;
}
