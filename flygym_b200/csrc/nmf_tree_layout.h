// nmf_tree_layout.h — tables and parameters of the general-topology step kernels (nmf_tree.cuh; host + device).
//
// The star-topology kernels of nmf_step.cuh are specialised to the reference benchmark skeleton (hub + 6 x 8 links).  Every
// other model the reference's compose layer can emit -- JointPreset.ALL_BIOLOGICAL (126 hinge DoFs: head, proboscis, antennae,
// eyes, abdomen, wings, halteres as well as the legs) and ALL_POSSIBLE (reference src/flygym/anatomy.py:388-460),
// ContactBodiesPreset.ALL (anatomy.py:519-526), ActuatedDOFPreset.ALL -- is a free root body carrying an arbitrary tree of
// hinge-jointed bodies.  The tree kernels step such a model from flat tables: one int table, one real table, offsets below.
#pragma once

namespace nmf {

constexpr int TREE_CTA = 128;                  // threads per fly
constexpr int TREE_NW = TREE_CTA / 32;         // warps per fly: each owns a set of root-child subtrees
constexpr int TREE_NROOT = 6;                  // DoFs of the free root joint
constexpr int TREE_MAXBODY = 96, TREE_MAXNV = 256, TREE_MAXGEOM = 96, TREE_MAXNU = 256, TREE_MAXDEPTH = 40;
constexpr int TCON_STRIDE = 18;                // reals per contact slot in shared memory: ContactG (14) + sv (3) + adhesion pull (1)
// noslip post-solver (models baked with noslip_iterations > 0): at most TNS_MAXC simultaneous contacts, two friction dimensions each.
// Region of TNS_TOTAL reals: B_tt (TNS_LD x TNS_LD) | g | g at Newton | pair limits | tangential rows | explicit basis forces (n, t1, t2)
// per ranked contact | slot of every ranked contact (ints) | count
constexpr int TNS_MAXC = 48, TNS_LD = 2 * TNS_MAXC;
constexpr int TNS_B = 0, TNS_G = TNS_LD * TNS_LD, TNS_G0 = TNS_G + TNS_LD, TNS_LIM = TNS_G0 + TNS_LD, TNS_JT = TNS_LIM + TNS_LD,
              TNS_GX = TNS_JT + TNS_LD, TNS_IDX = TNS_GX + 3 * TNS_MAXC, TNS_MISC = TNS_IDX + TNS_MAXC, TNS_TOTAL = TNS_MISC + 4;

struct TreeDims {
  int nb, nq, nv, nu_pos, nu_adh, ng, nseg, nleg, nslot, nH, maxd;
  // state record (floats): qpos | qvel | qacc_warmstart | ctrl | time, status, step count, pad ; every section 16-byte aligned
  int s_qpos, s_qvel, s_warm, s_ctrl, s_time, s_stride;
  // int table
  int i_parent, i_dofadr, i_ndof, i_leg;       // [nb]
  int i_child_adr, i_child;                    // CSR children of a body: [nb + 1], [nb - 1]
  int i_bg_adr, i_bg;                          // CSR contact geoms of a body: [nb + 1], [ng]
  int i_dof_body, i_cidx;                      // [nv]: body of a DoF, ctrl index of its position actuator (-1 = none)
  int i_rowadr, i_col, i_erow;                 // ancestor-sparse matrix rows: [nv + 1]; [nH] column (ancestor DoF) of every entry; [nH] row | column << 16
  int i_gbody, i_gtype, i_gvadr, i_gvnum;      // [ng]
  int i_adh_body;                              // [nu_adh]
  int i_wb_adr, i_wb;                          // bodies of warp w at tree depth d (d >= 1): adr[w * (maxd + 1) + d .. + 1], list
  int i_wk_adr, i_wk;                          // non-root DoFs of warp w, descending (descendants before ancestors): adr[w .. w + 1]; list of
                                               // 8-int descriptors {DoF k, start of its row, number of proper ancestors m, start of its pair list,
                                               // start of its descendant list, number of descendants, 0, 0}
  int i_pair;                                  // elimination of DoF k: m (m + 1) / 2 packed updates (target entry | p << 16 | q << 24)
  int i_desc;                                  // descendants j of DoF k: packed (entry of H[j][k] | j << 16)
  int nHa;                                     // nH rounded up to 4: targets >= nHa address the per-warp root-block accumulators behind H
  int i_total;
  // real table
  int r_body;     // [nb][18]: pos 3, quat 4, ipos 3, inertia about the COM in the body frame (xx yy zz xy xz yz) 6, mass, invweight
  int r_dof;      // [nv][11]: axis 3 (body frame), stiffness, damping, armature, springref, kp, kv, force lo, force hi
  int r_geom;     // [ng][8]: capsule centre 3, axis 3 (body frame), radius, half length
  int r_adh;      // [nu_adh][3]: gain, ctrl lo, ctrl hi
  int r_total;
  // shared-memory plan (reals)
  int m_state, m_stage, m_xpos, m_xquat, m_cinert, m_crb, m_cdof, m_cvel, m_acc, m_y, m_P, m_fs, m_grad, m_x, m_u, m_H, m_dinv,
      m_con, m_accS, m_rb, m_red, m_hullv, m_misc, m_weld, m_total;
  int noslip;     // the model asks for the noslip post-solver: the two regions below exist
  int m_ns, m_nsrank;   // noslip region (TNS_TOTAL reals); rank of every contact slot among the active ones, -1 = inactive ([ng * nslot] ints)
};
constexpr int TR_BODY = 18, TR_DOF = 11, TR_GEOM = 8, TR_ADH = 3;

template <class real>
struct TreeParamsT {
  float* state;              // [n_flies][s_stride]
  double* state64;           // f64 build: full-precision records between launches (optional)
  float* shadow;             //   and the float records as the last f64 launch left them
  const int* it;             // int table
  const real* rt;            // real table
  const real* hull;          // hull vertices (xyz) in body frames
  const int* hull_nbr_adr;   // CSR adjacency of the hull vertices (indices local to the geom)
  const int* hull_nbr;
  const float* act_table;    // optional [n_flies][table_T][table_cols]
  const float* seg_tab;      // [nseg][8]: body (bit pattern), pos xyz, quat wxyz
  float *out_xpos, *out_xquat, *out_actf, *out_sensor, *out_qpos, *out_energy, *dbg;
  int n_flies, nsteps, table_T, table_t0, table_cols, forward_only;
  TreeDims d;
  real dt, gx, gy, gz, inv_total_mass;
  real mu, cK, cB, margin, impratio;
  real solimp[5];
  int max_newton, max_ls, multiccd, terrain;
  int noslip_iterations; real noslip_tol, noslip_scale;      // as in StepParamsT
  real terr[8];
  // TetheredWorld weld on the root body (reference world.py:350-366), same fields and meaning as in StepParamsT
  int weld;
  real weld_a[3], weld_q[4], weld_K, weld_B, weld_imp[5], weld_ts, weld_invw[2];
};

constexpr int TDBG_NITER = 0, TDBG_NCON = 1, TDBG_NLS = 2, TDBG_STRIDE = 4;

}  // namespace nmf
