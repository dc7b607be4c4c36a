// nmf_layout.h — shared constants of the sm_100a step path (host + device).
//
// Topology handled by the kernels: the reference benchmark model
// (/root/reference/src/flygym_demo/benchmark/time_gpu_simulation.py:21-64):
// one free "hub" body (thorax + all jointless segments fused,
// mujoco_globals.yaml:5 fusestatic) carrying six identical 8-link leg chains with
// (3,2,1,1,1,1,1,1) hinge DoFs -> nv = 6 + 6*11 = 72.
#pragma once

namespace nmf {

constexpr int NLEG = 6;
constexpr int NLINK = 8;        // bodies per leg  (coxa .. tarsus5)
constexpr int NLEGDOF = 11;     // hinge DoFs per leg
constexpr int NV = 6 + NLEG * NLEGDOF;   // 72
constexpr int NQ = NV + 1;               // 73
constexpr int CTA = 64;         // threads per fly: 48 leg-body lanes + 16 hub lanes
constexpr int NHUBLANE = 16;
constexpr int MAXU = 80;        // max controls kept in the state record

// per-fly state record (floats), every section 16-byte aligned so one TMA bulk
// copy (cp.async.bulk) moves the whole record between HBM and shared memory
constexpr int S_QPOS = 0;       // 73 (+3 pad)
constexpr int S_QVEL = 76;      // 72
constexpr int S_WARM = 148;     // 72  qacc_warmstart
constexpr int S_CTRL = 220;     // nu <= MAXU
constexpr int S_TIME = 300;     // time, status word, steps since reset, pad
// status word (S_TIME + 1, a small integer kept as a float; sticky until the fly is reset): device-side faults are reported
// per fly, never by trapping (SURVEY.md 8b error convention)
constexpr int ST_NONFINITE = 1;   // a velocity became NaN / infinite
constexpr int ST_NEWTON_CAP = 2;  // the Newton solver hit its iteration cap with the active set still changing
constexpr int ST_LS_CAP = 4;      // a line search used up its evaluation cap
constexpr int ST_NOSLIP_SKIP = 8; // more simultaneous contacts than the noslip pass handles: the step ran without it
constexpr int S_STRIDE = 304;

// thread-role constant table: role[field * CTA + tid]
enum RoleField {
  RF_BPOS = 0,            // 3  body_pos (parent frame)
  RF_BQUAT = 3,           // 4
  RF_IPOS = 7,            // 3  inertial-frame origin in the body frame
  RF_IB = 10,             // 6  body-frame inertia tensor about the COM: xx yy zz xy xz yz
  RF_MASS = 16,
  RF_INVW = 17,           // body_invweight0[.,0]
  RF_NDOF = 18,           // int
  RF_AXIS = 19,           // 9  local hinge axes (3 x xyz)
  RF_STIFF = 28,          // 3
  RF_DAMP = 31,           // 3
  RF_ARM = 34,            // 3
  RF_SREF = 37,           // 3
  RF_KP = 40,             // 3
  RF_KV = 43,             // 3
  RF_FLO = 46,            // 3
  RF_FHI = 49,            // 3
  RF_CIDX = 52,           // 3 int: ctrl index of the position actuator on this dof, -1 = none
  RF_GTYPE = 55,          // int: -1 none, 0 capsule, 1 convex hull
  RF_GPOS = 56,           // 3  capsule centre (body frame)
  RF_GAXIS = 59,          // 3  capsule axis (body frame)
  RF_GRAD = 62,
  RF_GHALF = 63,
  RF_GVADR = 64,          // int
  RF_GVNUM = 65,          // int
  RF_ADH_GAIN = 66,
  RF_ADH_LO = 67,
  RF_ADH_HI = 68,
  RF_ADH_CIDX = 69,       // int, -1 = no adhesion actuator on this body
  RF_DOF0 = 70,           // int: global dof index of the lane's first dof
  RF_LEGSENSOR = 71,      // int: 1 if this body's contacts count for the leg contact sensor
  RF_CARM = 72,           // 3: armature of the lane's matrix-column DoFs (leg dof t-6 | t+2 | 10); 1 on hub chains
  RF_CDMP = 75,           // 3: damping of the same DoFs
  RF_GIDX = 78,           // int: index of the lane's contact geom in the model's geom order (contact order of the noslip sweeps)
  RF_COUNT = 79
};

// debug dump (floats per fly), only written when a dump buffer is passed
constexpr int DBG_NITER = 0, DBG_NCON = 1, DBG_NLS = 2, DBG_NCHG = 3;
constexpr int DBG_FS = 4;                    // qfrc_smooth[72]
constexpr int DBG_QACC = DBG_FS + NV;        // qacc[72]
constexpr int DBG_FC = DBG_QACC + NV;        // qfrc_constraint[72]
constexpr int DBG_QACCE = DBG_FC + NV;       // Euler (implicit-damping) acceleration[72]
constexpr int DBG_NSLOT = 4;                 // contact slots per thread in the dump (the capsule kernels fill the first two)
constexpr int DBG_CON = DBG_QACCE + NV;      // per thread, DBG_NSLOT slots x (active, dist, x, y, z, fn)
constexpr int DBG_XPOS = DBG_CON + CTA * DBG_NSLOT * 6; // per thread xpos (3)
constexpr int DBG_CDOF = DBG_XPOS + CTA * 3; // cdof[72][6]
constexpr int DBG_HROWS = DBG_CDOF + NV * 6; // Euler matrix rows: 6 legs x 177, then 21 base
constexpr int DBG_STRIDE = DBG_HROWS + NLEG * 177 + 21 + 3;

// `real` = the arithmetic of the kernel instantiation (float for the product path, double for the validation build); buffers
// that live in HBM on behalf of the API (state records, observations, action table) are float32 in both.
constexpr int QUEUE_MAX_CHUNKS = 64;   // sub-chunks per fly and launch the queue buffer is sized for

template <class real>
struct StepParamsT {
  float* state;              // [n_flies][S_STRIDE]
  double* state64;           // f64 build only, optional [n_flies][S_STRIDE]: the records at full precision between launches
  float* shadow;             //   and the float records as the last f64 launch left them (to detect edits made through the API)
  const real* role;          // [RF_COUNT][CTA]
  const real* hull;          // hull vertices (xyz) in body frames
  const float* act_table;    // optional [n_flies][table_T][table_cols] -> ctrl[0:table_cols]; nullptr = use ctrl in state
  const float* seg_tab;      // [nseg][8]: body lane (as float), pos xyz, quat wxyz  (static segments on the hub)
  float* out_xpos;           // optional [n_flies][nseg][3]
  float* out_xquat;          // optional [n_flies][nseg][4]
  float* out_actf;           // optional [n_flies][nu]
  float* out_sensor;         // optional [n_flies][NLEG*16]
  float* out_qpos;           // optional [n_flies][NQ]: qpos after the launch's last step, packed (nmf_step_host reads it back)
  float* out_energy;         // optional [n_flies][2]: potential, kinetic energy of the state the last step started from
  float* dbg;                // optional [n_flies][DBG_STRIDE]
  const int* hull_nbr_adr;   // CSR adjacency of the hull vertices: neighbours of vertex v are hull_nbr[hull_nbr_adr[v] .. hull_nbr_adr[v+1])
  const int* hull_nbr;       //   (indices local to the geom)
  int n_flies, nsteps, table_T, table_t0, table_cols;
  int forward_only;          // 1: evaluate the current state (outputs) without advancing it (mj_forward)
  int nu_pos, nu_adh, nseg, nhubgeom;
  real dt, gx, gy, gz, inv_total_mass;
  real mu, cK, cB, margin, impratio;
  real solimp[5];           // sanitised: d0, dmax, width, midpoint, power
  int max_newton, max_ls;
  int noslip_iterations;     // sweeps of the noslip post-solver (NOSLIP kernel instantiations; mujoco_globals.yaml:15), 0 = off
  real noslip_tol, noslip_scale;   // MuJoCo's noslip_tolerance (1e-6) and the cost scale 1 / (meaninertia * nv)
  int multiccd;              // 1: plane-hull contacts also at the support vertex's neighbours within the margin (W_MESH kernels)
  // terrain: 0 = ground plane z = 0 (FlatGroundWorld); 1 = floor plane + grid of box columns, terr = {Px, Py, hx, hy, top_even,
  // top_odd, z_floor, 0}: column (i, j) covers |x - i Px| <= hx, |y - j Py| <= hy, z <= top_{(i+j)&1}
  int terrain;
  real terr[8];
  // TetheredWorld weld (reference world.py:350-366): the hub-frame point weld_a is pulled onto the world origin and
  // q_hub * weld_q onto the identity by six always-active (equality) rows with their own solref / solimp
  int weld;
  real weld_a[3], weld_q[4], weld_K, weld_B, weld_imp[5], weld_ts, weld_invw[2];
  // work-queue scheduling (nullptr = one block per fly for the whole launch): queue[0] = next work item,
  // queue[1 + fly] = number of sub-chunks of that fly already written back
  int* queue;
  int sub_steps, n_items;
  // item boundaries of a tapered schedule (n_chunks > 0): sub-chunk c of every unit covers steps chunk_start[c] .. chunk_start[c + 1];
  // n_chunks == 0: uniform items of sub_steps steps
  int n_chunks;
  short chunk_start[QUEUE_MAX_CHUNKS + 1];
};
using StepParams = StepParamsT<float>;

}  // namespace nmf
