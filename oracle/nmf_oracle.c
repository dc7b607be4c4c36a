/*
 * nmf_oracle.c — CPU fp64 restatement of the reference's physics step.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product path
 * (flygym_b200/csrc) never links or calls it.
 *
 * PARITY UNPINNED: the reference's arithmetic for this path is the third-party
 * C library MuJoCo 3.6.0 (pinned in /root/reference/uv.lock:1335-1336), reached
 * through exactly one call, `mj.mj_step(self.mj_model, self.mj_data)`
 * (reference src/flygym/simulation.py:76; batched twin mjw.step at
 * src/flygym/warp/simulation.py:263).  MuJoCo is not vendored in the reference,
 * not installed here and not installable (no network), and the reference's tests
 * hold no golden qpos/qvel vectors.  This file restates
 * MuJoCo's published pipeline ("Computation" chapter: kinematics -> comPos ->
 * CRB -> collision -> constraint construction -> transmission -> comVel ->
 * passive -> RNE -> actuation -> smooth acceleration -> Newton solver ->
 * semi-implicit Euler with implicit joint damping) for the model the reference
 * composes (src/flygym/compose/fly.py, world.py, physics.py), written as plain
 * dense linear algebra on a general kinematic tree (free root + hinges).
 *
 * Deliberately NOT structured like the CUDA kernels (those exploit the star
 * topology, fp32, fused scans); agreement between the two is therefore a
 * meaningful cross-check.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MINVAL 1e-15
#define MINIMP 0.0001
#define MAXIMP 0.9999
#define GEOM_CAPSULE 0
#define GEOM_HULL 1

typedef struct {
  /* dims */
  int nbody, nq, nv, nu_pos, nu_adh, ngeom, nsite, nseg, nleg, nhullvert, nu;
  /* model (pointers into owned blob copy) */
  const double *body_pos, *body_quat, *body_mass, *body_ipos, *body_iquat, *body_inertia, *body_invweight0;
  const double *dof_axis, *dof_stiffness, *dof_damping, *dof_armature, *dof_springref;
  const double *act_kp, *act_kv, *act_frcrange, *adh_gain, *adh_ctrlrange;
  const double *geom_pos, *geom_quat, *geom_size, *hull_vert, *site_pos, *seg_pos, *seg_quat;
  const double *key_qpos, *key_ctrl, *opt, *contact, *terrain;   /* terrain: optional section, NULL or type 0 = flat */
  const double* weld;   /* optional section (TetheredWorld): enabled, anchor[3], quat[4], solref[2], solimp[5], torquescale, invweight tran/rot */
  int neq;              /* equality rows at the head of the efc arrays (6 with a weld) */
  const int32_t *body_parent, *body_dofadr, *body_dofnum, *body_leg, *dof_body, *dof_parent;
  const int32_t *act_dof, *adh_body, *geom_body, *geom_type, *geom_vertadr, *geom_vertnum;
  const int32_t *site_body, *seg_body, *leg_rootbody;
  const int32_t *hull_nbr_adr, *hull_nbr;   /* CSR adjacency of the hull vertices (optional; indices local to the geom) */
  int multiccd;                             /* plane-hull: extra contacts at the support vertex's neighbours (mujoco_globals.yaml:18) */
  /* options */
  double dt, grav[3], tolerance, ls_tolerance, meaninertia, impratio;
  int iterations, ls_iterations, noslip_iterations;
  double mu, solref[2], solimp[5], margin, gap;
  void* blob;

  /* state */
  double *qpos, *qvel, *ctrl, *qacc_warmstart, time;
  /* derived */
  double *xpos, *xquat, *xmat, *xipos, *ximat, *xanchor, *xaxis, com[3];
  double *cinert, *crb, *cdof, *cdof_dot, *cvel, *cacc, *cfrc;
  double *M, *qfrc_bias, *qfrc_passive, *qfrc_actuator, *actuator_force, *qfrc_smooth, *qacc_smooth;
  double *qacc, *qfrc_constraint, *adh_moment;
  /* contacts */
  int ncon, maxcon;
  double *con_dist, *con_pos, *con_frame; /* frame: normal, t1, t2 rows */
  int *con_geom;
  /* constraint rows */
  int nefc;
  double *efc_J, *efc_pos, *efc_D, *efc_R, *efc_aref, *efc_vel, *efc_force, *efc_jar;
  int *efc_active;
  /* outputs */
  double *geom_xpos, *geom_xmat, *site_xpos, *seg_xpos, *seg_xquat, *sensordata;
  /* solver stats */
  int solver_niter, noslip_niter; double solver_cost, solver_gradnorm, noslip_tolerance;
  double energy[2];   /* potential, kinetic (mjData.energy with the model's `energy` flag) */
  char err[256];
} nmfo;

/* ------------------------------------------------------------------ small math */
static void cross3(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double dotn(const double* a, const double* b, int n) { double s = 0; for (int i = 0; i < n; i++) s += a[i] * b[i]; return s; }
static void quat_mul(double* r, const double* a, const double* b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void quat_norm(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; } else { for (int i = 0; i < 4; i++) q[i] /= n; }
}
static void quat2mat(double* m, const double* q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = 1 - 2 * (x * x + z * z); m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = 1 - 2 * (x * x + y * y);
}
static void mat_vec(double* r, const double* m, const double* v) {
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
         z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void axis_angle_quat(double* q, const double* axis, double angle) {
  double s = sin(angle * 0.5);
  q[0] = cos(angle * 0.5); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
/* spatial inertia (10-vector: Ixx Iyy Izz Ixy Ixz Iyz, m*off xyz, m) times motion vector (ang, lin) */
static void mul_inert_vec(double* r, const double* i, const double* v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
static void cross_motion(double* r, const double* vel, const double* v) {
  double a[3], b[3], c[3];
  cross3(a, vel, v); cross3(b, vel, v + 3); cross3(c, vel + 3, v);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
static void cross_force(double* r, const double* vel, const double* f) {
  double a[3], b[3], c[3];
  cross3(a, vel, f); cross3(b, vel + 3, f + 3); cross3(c, vel, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}

/* dense Cholesky (lower, in place); returns 0 on success */
static int chol_factor(double* A, int n) {
  for (int j = 0; j < n; j++) {
    double s = A[j * n + j];
    for (int k = 0; k < j; k++) s -= A[j * n + k] * A[j * n + k];
    if (s <= 0) return -1;
    double d = sqrt(s);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double t = A[i * n + j];
      for (int k = 0; k < j; k++) t -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = t / d;
    }
  }
  return 0;
}
static void chol_solve(const double* L, int n, double* x) {
  for (int i = 0; i < n; i++) { double s = x[i]; for (int k = 0; k < i; k++) s -= L[i * n + k] * x[k]; x[i] = s / L[i * n + i]; }
  for (int i = n - 1; i >= 0; i--) { double s = x[i]; for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k]; x[i] = s / L[i * n + i]; }
}

/* ------------------------------------------------------------------ blob */
typedef struct { char name[24]; int32_t dtype; int32_t count; int64_t offset; } section_t;
static const void* find_section(const void* blob, const char* name, int* count) {
  const char* p = (const char*)blob;
  int32_t nsec; memcpy(&nsec, p + 12, 4);
  const char* tab = p + 16;
  for (int i = 0; i < nsec; i++) {
    section_t s; memcpy(&s, tab + i * 40, 40);
    if (strncmp(s.name, name, 24) == 0) { if (count) *count = s.count; return p + s.offset; }
  }
  return NULL;
}
#define SEC(field, nm) do { o->field = find_section(o->blob, nm, NULL); if (!o->field) { free(o->blob); free(o); return NULL; } } while (0)

static double* dalloc(size_t n) { return (double*)calloc(n ? n : 1, sizeof(double)); }

nmfo* nmfo_create(const void* blob, size_t nbytes) {
  if (nbytes < 16 || memcmp(blob, "NMFB200", 8) != 0) return NULL;
  nmfo* o = (nmfo*)calloc(1, sizeof(nmfo));
  o->blob = malloc(nbytes); memcpy(o->blob, blob, nbytes);
  const int32_t* dims = find_section(o->blob, "dims", NULL);
  if (!dims) { free(o->blob); free(o); return NULL; }
  o->nbody = dims[0]; o->nq = dims[1]; o->nv = dims[2]; o->nu_pos = dims[3]; o->nu_adh = dims[4];
  o->ngeom = dims[5]; o->nsite = dims[6]; o->nseg = dims[7]; o->nleg = dims[8]; o->nhullvert = dims[9];
  o->nu = o->nu_pos + o->nu_adh;
  SEC(body_pos, "body_pos"); SEC(body_quat, "body_quat"); SEC(body_mass, "body_mass"); SEC(body_ipos, "body_ipos");
  SEC(body_iquat, "body_iquat"); SEC(body_inertia, "body_inertia"); SEC(body_invweight0, "body_invweight0");
  SEC(dof_axis, "dof_axis"); SEC(dof_stiffness, "dof_stiffness"); SEC(dof_damping, "dof_damping");
  SEC(dof_armature, "dof_armature"); SEC(dof_springref, "dof_springref");
  SEC(act_kp, "act_kp"); SEC(act_kv, "act_kv"); SEC(act_frcrange, "act_frcrange"); SEC(adh_gain, "adh_gain");
  SEC(adh_ctrlrange, "adh_ctrlrange"); SEC(geom_pos, "geom_pos"); SEC(geom_quat, "geom_quat");
  SEC(geom_size, "geom_size"); SEC(hull_vert, "hull_vert"); SEC(site_pos, "site_pos"); SEC(seg_pos, "seg_pos");
  SEC(seg_quat, "seg_quat"); SEC(key_qpos, "key_qpos"); SEC(key_ctrl, "key_ctrl"); SEC(opt, "opt"); SEC(contact, "contact");
  o->terrain = find_section(o->blob, "terrain", NULL);
  o->weld = find_section(o->blob, "weld", NULL);
  if (o->weld && o->weld[0] == 0) o->weld = NULL;
  SEC(body_parent, "body_parent"); SEC(body_dofadr, "body_dofadr"); SEC(body_dofnum, "body_dofnum");
  SEC(body_leg, "body_leg"); SEC(dof_body, "dof_body"); SEC(dof_parent, "dof_parent"); SEC(act_dof, "act_dof");
  SEC(adh_body, "adh_body"); SEC(geom_body, "geom_body"); SEC(geom_type, "geom_type");
  SEC(geom_vertadr, "geom_vertadr"); SEC(geom_vertnum, "geom_vertnum"); SEC(site_body, "site_body");
  SEC(seg_body, "seg_body"); SEC(leg_rootbody, "leg_rootbody");
  o->hull_nbr_adr = find_section(o->blob, "hull_nbr_adr", NULL); o->hull_nbr = find_section(o->blob, "hull_nbr", NULL);
  o->dt = o->opt[0]; o->grav[0] = o->opt[1]; o->grav[1] = o->opt[2]; o->grav[2] = o->opt[3];
  o->iterations = (int)o->opt[4]; o->tolerance = o->opt[5]; o->ls_iterations = (int)o->opt[6];
  o->ls_tolerance = o->opt[7]; o->noslip_iterations = (int)o->opt[8]; o->meaninertia = o->opt[9];
  o->impratio = o->opt[10];
  o->noslip_tolerance = 1e-6;   /* MuJoCo default (the reference does not set it) */
  { int nopt = 0; find_section(o->blob, "opt", &nopt); o->multiccd = nopt > 11 ? (int)o->opt[11] : 0; }
  o->mu = o->contact[0]; o->solref[0] = o->contact[1]; o->solref[1] = o->contact[2];
  for (int i = 0; i < 5; i++) o->solimp[i] = o->contact[3 + i];
  o->margin = o->contact[8]; o->gap = o->contact[9];

  int nb = o->nbody, nv = o->nv;
  o->qpos = dalloc(o->nq); o->qvel = dalloc(nv); o->ctrl = dalloc(o->nu); o->qacc_warmstart = dalloc(nv);
  o->xpos = dalloc(3 * nb); o->xquat = dalloc(4 * nb); o->xmat = dalloc(9 * nb); o->xipos = dalloc(3 * nb);
  o->ximat = dalloc(9 * nb); o->xanchor = dalloc(3 * nv); o->xaxis = dalloc(3 * nv);
  o->cinert = dalloc(10 * nb); o->crb = dalloc(10 * nb); o->cdof = dalloc(6 * nv); o->cdof_dot = dalloc(6 * nv);
  o->cvel = dalloc(6 * nb); o->cacc = dalloc(6 * nb); o->cfrc = dalloc(6 * nb);
  o->M = dalloc((size_t)nv * nv); o->qfrc_bias = dalloc(nv); o->qfrc_passive = dalloc(nv);
  o->qfrc_actuator = dalloc(nv); o->actuator_force = dalloc(o->nu); o->qfrc_smooth = dalloc(nv);
  o->qacc_smooth = dalloc(nv); o->qacc = dalloc(nv); o->qfrc_constraint = dalloc(nv);
  o->adh_moment = dalloc((size_t)o->nu_adh * nv);
  o->maxcon = 4 * o->ngeom;
  o->con_dist = dalloc(o->maxcon); o->con_pos = dalloc(3 * o->maxcon); o->con_frame = dalloc(9 * o->maxcon);
  o->con_geom = (int*)calloc(o->maxcon, sizeof(int));
  int maxefc = 4 * o->maxcon + 6;
  o->efc_J = dalloc((size_t)maxefc * nv); o->efc_pos = dalloc(maxefc); o->efc_D = dalloc(maxefc);
  o->efc_R = dalloc(maxefc); o->efc_aref = dalloc(maxefc); o->efc_vel = dalloc(maxefc);
  o->efc_force = dalloc(maxefc); o->efc_jar = dalloc(maxefc); o->efc_active = (int*)calloc(maxefc, sizeof(int));
  o->geom_xpos = dalloc(3 * o->ngeom); o->geom_xmat = dalloc(9 * o->ngeom); o->site_xpos = dalloc(3 * o->nsite);
  o->seg_xpos = dalloc(3 * o->nseg); o->seg_xquat = dalloc(4 * o->nseg); o->sensordata = dalloc(16 * o->nleg);
  return o;
}

void nmfo_destroy(nmfo* o) {
  if (!o) return;
  double* d[] = {o->qpos, o->qvel, o->ctrl, o->qacc_warmstart, o->xpos, o->xquat, o->xmat, o->xipos, o->ximat, o->xanchor,
                 o->xaxis, o->cinert, o->crb, o->cdof, o->cdof_dot, o->cvel, o->cacc, o->cfrc, o->M, o->qfrc_bias,
                 o->qfrc_passive, o->qfrc_actuator, o->actuator_force, o->qfrc_smooth, o->qacc_smooth, o->qacc,
                 o->qfrc_constraint, o->adh_moment, o->con_dist, o->con_pos, o->con_frame, o->efc_J, o->efc_pos, o->efc_D,
                 o->efc_R, o->efc_aref, o->efc_vel, o->efc_force, o->efc_jar, o->geom_xpos, o->geom_xmat, o->site_xpos,
                 o->seg_xpos, o->seg_xquat, o->sensordata};
  for (size_t i = 0; i < sizeof(d) / sizeof(d[0]); i++) free(d[i]);
  free(o->con_geom); free(o->efc_active); free(o->blob); free(o);
}

/* reference: Simulation.reset -> mj_resetDataKeyframe(neutral)  (simulation.py:59-62) */
void nmfo_reset(nmfo* o) {
  memcpy(o->qpos, o->key_qpos, sizeof(double) * o->nq);
  memset(o->qvel, 0, sizeof(double) * o->nv);
  memcpy(o->ctrl, o->key_ctrl, sizeof(double) * o->nu);
  memset(o->qacc_warmstart, 0, sizeof(double) * o->nv);
  o->time = 0;
}

/* ------------------------------------------------------------------ position stage */
/* mj_kinematics: free root (qpos[0:7]) + hinge chain; hinge anchors at jnt_pos = 0 */
static void kinematics(nmfo* o) {
  int nb = o->nbody;
  quat_norm(o->qpos + 3);
  for (int b = 0; b < nb; b++) {
    double* xp = o->xpos + 3 * b; double* xq = o->xquat + 4 * b;
    int p = o->body_parent[b];
    if (p < 0) {
      memcpy(xp, o->qpos, 24); memcpy(xq, o->qpos + 3, 32);
      for (int k = 0; k < 3; k++) {  /* translational dofs: world axes, anchor at xpos */
        memcpy(o->xanchor + 3 * k, xp, 24); o->xaxis[3 * k + 0] = k == 0; o->xaxis[3 * k + 1] = k == 1; o->xaxis[3 * k + 2] = k == 2;
      }
    } else {
      double t[3]; mat_vec(t, o->xmat + 9 * p, o->body_pos + 3 * b);
      for (int k = 0; k < 3; k++) xp[k] = o->xpos[3 * p + k] + t[k];
      quat_mul(xq, o->xquat + 4 * p, o->body_quat + 4 * b);
      int adr = o->body_dofadr[b];
      for (int j = 0; j < o->body_dofnum[b]; j++) {
        int d = adr + j; double m[9], ql[4];
        quat2mat(m, xq);
        memcpy(o->xanchor + 3 * d, xp, 24);            /* jnt_pos = 0 */
        mat_vec(o->xaxis + 3 * d, m, o->dof_axis + 3 * d);
        axis_angle_quat(ql, o->dof_axis + 3 * d, o->qpos[d + 1]);  /* qpos0 = 0 */
        double nq[4]; quat_mul(nq, xq, ql); memcpy(xq, nq, 32);
        /* off-centre correction vanishes because jnt_pos = 0 */
      }
      quat_norm(xq);
    }
    quat2mat(o->xmat + 9 * b, xq);
    if (p < 0) for (int k = 0; k < 3; k++) { /* rotational free dofs: body-local axes */
      int d = 3 + k; memcpy(o->xanchor + 3 * d, xp, 24);
      o->xaxis[3 * d + 0] = o->xmat[0 + k]; o->xaxis[3 * d + 1] = o->xmat[3 + k]; o->xaxis[3 * d + 2] = o->xmat[6 + k];
    }
    double t[3]; mat_vec(t, o->xmat + 9 * b, o->body_ipos + 3 * b);
    for (int k = 0; k < 3; k++) o->xipos[3 * b + k] = xp[k] + t[k];
    double qi[4]; quat_mul(qi, xq, o->body_iquat + 4 * b); quat2mat(o->ximat + 9 * b, qi);
  }
  for (int g = 0; g < o->ngeom; g++) {
    int b = o->geom_body[g]; double t[3], q[4];
    mat_vec(t, o->xmat + 9 * b, o->geom_pos + 3 * g);
    for (int k = 0; k < 3; k++) o->geom_xpos[3 * g + k] = o->xpos[3 * b + k] + t[k];
    quat_mul(q, o->xquat + 4 * b, o->geom_quat + 4 * g); quat2mat(o->geom_xmat + 9 * g, q);
  }
  for (int s = 0; s < o->nsite; s++) {
    int b = o->site_body[s]; double t[3]; mat_vec(t, o->xmat + 9 * b, o->site_pos + 3 * s);
    for (int k = 0; k < 3; k++) o->site_xpos[3 * s + k] = o->xpos[3 * b + k] + t[k];
  }
  for (int s = 0; s < o->nseg; s++) {
    int b = o->seg_body[s]; double t[3]; mat_vec(t, o->xmat + 9 * b, o->seg_pos + 3 * s);
    for (int k = 0; k < 3; k++) o->seg_xpos[3 * s + k] = o->xpos[3 * b + k] + t[k];
    quat_mul(o->seg_xquat + 4 * s, o->xquat + 4 * b, o->seg_quat + 4 * s);
  }
}

/* mj_comPos: subtree COM of the tree root, cinert and cdof about it */
static void com_pos(nmfo* o) {
  int nb = o->nbody, nv = o->nv; double mt = 0, c[3] = {0, 0, 0};
  for (int b = 0; b < nb; b++) { mt += o->body_mass[b]; for (int k = 0; k < 3; k++) c[k] += o->body_mass[b] * o->xipos[3 * b + k]; }
  for (int k = 0; k < 3; k++) o->com[k] = c[k] / mt;
  for (int b = 0; b < nb; b++) {
    const double* R = o->ximat + 9 * b; const double* I = o->body_inertia + 3 * b; double m = o->body_mass[b];
    double off[3]; for (int k = 0; k < 3; k++) off[k] = o->xipos[3 * b + k] - o->com[k];
    double G[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
      double s = 0; for (int k = 0; k < 3; k++) s += R[3 * i + k] * I[k] * R[3 * j + k];
      G[3 * i + j] = s + m * ((i == j ? dot3(off, off) : 0) - off[i] * off[j]);
    }
    double* ci = o->cinert + 10 * b;
    ci[0] = G[0]; ci[1] = G[4]; ci[2] = G[8]; ci[3] = G[1]; ci[4] = G[2]; ci[5] = G[5];
    ci[6] = m * off[0]; ci[7] = m * off[1]; ci[8] = m * off[2]; ci[9] = m;
  }
  for (int d = 0; d < nv; d++) {
    double* cd = o->cdof + 6 * d;
    if (d < 3) { cd[0] = cd[1] = cd[2] = 0; cd[3] = d == 0; cd[4] = d == 1; cd[5] = d == 2; continue; }
    double off[3]; for (int k = 0; k < 3; k++) off[k] = o->com[k] - o->xanchor[3 * d + k];
    memcpy(cd, o->xaxis + 3 * d, 24); cross3(cd + 3, o->xaxis + 3 * d, off);
  }
}

/* mj_crb: composite rigid body -> dense M (lower+upper filled) */
static void crb(nmfo* o) {
  int nb = o->nbody, nv = o->nv;
  memcpy(o->crb, o->cinert, sizeof(double) * 10 * nb);
  for (int b = nb - 1; b > 0; b--) { int p = o->body_parent[b]; for (int k = 0; k < 10; k++) o->crb[10 * p + k] += o->crb[10 * b + k]; }
  memset(o->M, 0, sizeof(double) * nv * nv);
  for (int i = 0; i < nv; i++) {
    double buf[6]; mul_inert_vec(buf, o->crb + 10 * o->dof_body[i], o->cdof + 6 * i);
    o->M[i * nv + i] = o->dof_armature[i] + dotn(o->cdof + 6 * i, buf, 6);
    for (int j = o->dof_parent[i]; j >= 0; j = o->dof_parent[j]) {
      double v = dotn(o->cdof + 6 * j, buf, 6); o->M[i * nv + j] = v; o->M[j * nv + i] = v;
    }
  }
}

/* mj_jac: translational Jacobian (3 x nv) of a point attached to body b */
static void jac_point(const nmfo* o, double* jacp, int b, const double* point) {
  int nv = o->nv; memset(jacp, 0, sizeof(double) * 3 * nv);
  double off[3]; for (int k = 0; k < 3; k++) off[k] = point[k] - o->com[k];
  int d = o->body_dofadr[b] + o->body_dofnum[b] - 1;
  for (; d >= 0; d = o->dof_parent[d]) {
    const double* cd = o->cdof + 6 * d; double t[3]; cross3(t, cd, off);
    for (int k = 0; k < 3; k++) jacp[k * nv + d] = cd[3 + k] + t[k];
  }
}

/* collision: explicit geom-plane pairs only (world.py:292-309); plane z=0, normal +z */
static void add_contact_n(nmfo* o, int g, double dist, const double* pos, const double* nrm, const double* hint) {
  /* general normal (terrain worlds): tangent = hint orthogonalised against the normal, world x (or y) if they are parallel */
  if (o->ncon >= o->maxcon) return;
  int c = o->ncon++; double* f = o->con_frame + 9 * c;
  o->con_dist[c] = dist; memcpy(o->con_pos + 3 * c, pos, 24); o->con_geom[c] = g;
  memcpy(f, nrm, 24);
  double t = dot3(hint, nrm), y[3] = {hint[0] - t * nrm[0], hint[1] - t * nrm[1], hint[2] - t * nrm[2]};
  if (dot3(y, y) < 1e-12) {
    double e[3] = {fabs(nrm[0]) < 0.9 ? 1.0 : 0.0, fabs(nrm[0]) < 0.9 ? 0.0 : 1.0, 0.0};
    t = dot3(e, nrm); for (int k = 0; k < 3; k++) y[k] = e[k] - t * nrm[k];
  }
  double n = sqrt(dot3(y, y)); for (int k = 0; k < 3; k++) y[k] /= n;
  memcpy(f + 3, y, 24); cross3(f + 6, f, y);
}
/* sphere vs terrain solid (floor plane + grid of box columns): closest point of the solid -> one contact candidate.
 * terrain = {type, Px, Py, hx, hy, top_even, top_odd, z_floor} (flygym_b200/model.py TERRAIN_FIELDS) */
static void sphere_terrain(const nmfo* o, const double* c, double rad, double* nrm, double* dist) {
  const double* T = o->terrain; double Px = T[1], Py = T[2], hx = T[3], hy = T[4];
  nrm[0] = 0; nrm[1] = 0; nrm[2] = 1; *dist = c[2] - T[7] - rad;
  double fi = nearbyint(c[0] / Px), fj = nearbyint(c[1] / Py);
  int i0 = (int)fi, j0 = (int)fj, sx = c[0] >= fi * Px ? 1 : -1, sy = c[1] >= fj * Py ? 1 : -1;
  for (int q = 0; q < 4; q++) {
    int i = i0 + ((q & 1) ? sx : 0), j = j0 + ((q & 2) ? sy : 0);
    double cx = i * Px, cy = j * Py, top = ((i + j) & 1) ? T[6] : T[5];
    double qx = fmin(fmax(c[0], cx - hx), cx + hx), qy = fmin(fmax(c[1], cy - hy), cy + hy), qz = fmin(c[2], top);
    double d[3] = {c[0] - qx, c[1] - qy, c[2] - qz}, d2 = dot3(d, d), dd, n[3] = {0, 0, 1};
    if (d2 > 0) { double l = sqrt(d2); dd = l - rad; for (int k = 0; k < 3; k++) n[k] = d[k] / l; }
    else dd = c[2] - top - rad;   /* centre inside the column: out through the top face */
    if (dd < *dist) { *dist = dd; memcpy(nrm, n, 24); }
  }
}
static void add_contact(nmfo* o, int g, double dist, const double* pos, const double* hint) {
  if (o->ncon >= o->maxcon) return;
  int c = o->ncon++; double* f = o->con_frame + 9 * c;
  o->con_dist[c] = dist; memcpy(o->con_pos + 3 * c, pos, 24); o->con_geom[c] = g;
  f[0] = 0; f[1] = 0; f[2] = 1;
  /* mju_makeFrame: tangent hint orthogonalised against the normal */
  double y[3] = {hint[0], hint[1], hint[2]};
  if (sqrt(dot3(y, y)) < 0.5) { y[0] = 0; y[1] = 1; y[2] = 0; }
  double t = dot3(f, y); for (int k = 0; k < 3; k++) y[k] -= t * f[k];
  double n = sqrt(dot3(y, y));
  if (n < MINVAL) { y[0] = 1; y[1] = 0; y[2] = 0; } else for (int k = 0; k < 3; k++) y[k] /= n;
  memcpy(f + 3, y, 24); cross3(f + 6, f, y);
}
static void collision(nmfo* o) {
  o->ncon = 0;
  for (int g = 0; g < o->ngeom; g++) {
    const double* gp = o->geom_xpos + 3 * g; const double* gm = o->geom_xmat + 9 * g;
    if (o->geom_type[g] == GEOM_CAPSULE) {  /* mjc_PlaneCapsule: two sphere tests, frame aligned with capsule axis */
      double r = o->geom_size[2 * g], h = o->geom_size[2 * g + 1];
      double axis[3] = {gm[2], gm[5], gm[8]};
      for (int s = 0; s < 2; s++) {
        double sg = s == 0 ? 1.0 : -1.0, c[3];
        for (int k = 0; k < 3; k++) c[k] = gp[k] + sg * h * axis[k];
        if (o->terrain && o->terrain[0] != 0) {   /* terrain world: each end sphere against the terrain solid */
          double nrm[3], dist; sphere_terrain(o, c, r, nrm, &dist);
          if (dist > o->margin) continue;
          double pos[3]; for (int k = 0; k < 3; k++) pos[k] = c[k] - (r + dist / 2) * nrm[k];
          add_contact_n(o, g, dist, pos, nrm, axis);
          continue;
        }
        double cdist = c[2];
        if (cdist > o->margin + r) continue;
        double dist = cdist - r, pos[3] = {c[0], c[1], c[2] - (r + dist / 2)};
        add_contact(o, g, dist, pos, axis);
      }
    } else {  /* plane - convex hull: deepest hull vertex (support point along -normal) */
      int b = o->geom_body[g]; const double* R = o->xmat + 9 * b; const double* xp = o->xpos + 3 * b;
      double best = 1e300, bp[3] = {0, 0, 0}; int bi = 0;
      for (int v = 0; v < o->geom_vertnum[g]; v++) {
        const double* lv = o->hull_vert + 3 * (o->geom_vertadr[g] + v); double w[3]; mat_vec(w, R, lv);
        for (int k = 0; k < 3; k++) w[k] += xp[k];
        if (w[2] < best) { best = w[2]; memcpy(bp, w, 24); bi = v; }
      }
      if (best > o->margin) continue;
      double pos[3] = {bp[0], bp[1], bp[2] - best / 2}, zero[3] = {0, 0, 0};
      add_contact(o, g, best, pos, zero);
      /* [PRIOR] mjc_PlaneConvex on a mesh with the `multiccd` flag (mujoco_globals.yaml:18): the neighbours of the support vertex
       * in the hull's vertex graph that are also within the margin become contacts too, up to 4 per geom (graph order). */
      if (o->multiccd && o->hull_nbr_adr && o->hull_nbr) {
        int adr = o->geom_vertadr[g], cnt = 1;
        for (int e = o->hull_nbr_adr[adr + bi]; e < o->hull_nbr_adr[adr + bi + 1] && cnt < 4; e++) {
          const double* lv = o->hull_vert + 3 * (adr + o->hull_nbr[e]); double w[3]; mat_vec(w, R, lv);
          for (int k = 0; k < 3; k++) w[k] += xp[k];
          if (w[2] > o->margin) continue;
          double p2[3] = {w[0], w[1], w[2] - w[2] / 2};
          add_contact(o, g, w[2], p2, zero); cnt++;
        }
      }
    }
  }
}

/* impedance d(r) — getimpedance with sanitised solimp */
static double impedance_of(const double* solimp, double pos_minus_margin) {
  double d0 = fmin(MAXIMP, fmax(MINIMP, solimp[0])), d1 = fmin(MAXIMP, fmax(MINIMP, solimp[1]));
  double width = fmax(0, solimp[2]), mid = fmin(MAXIMP, fmax(MINIMP, solimp[3])), power = fmax(1, solimp[4]);
  if (d0 == d1 || width <= MINVAL) return 0.5 * (d0 + d1);
  double x = fabs(pos_minus_margin) / width;
  if (x >= 1) return d1;
  if (x <= 0) return d0;
  double y;
  if (power == 1) y = x;
  else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
  else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
  return d0 + y * (d1 - d0);
}

static double impedance(const nmfo* o, double pos_minus_margin) { return impedance_of(o->solimp, pos_minus_margin); }

/* weld equality (TetheredWorld, reference world.py:350-366): [PRIOR] mj_instantiateEquality, mjEQ_WELD with body2 = world.
 * Rows 0-2: position of the body-1 point `anchor` (the weld's relpose position, in the hub frame here) minus the world
 * origin; rows 3-5: torquescale * vector part of  q_body1 * relquat;  Jacobian of the rotation rows = 0.5 * vec((0, w) * q).
 * Equality rows are active on both sides (quadratic cost).  diagApprox = body_invweight0 (translational | rotational). */
static void make_weld(nmfo* o) {
  int nv = o->nv; const double* W = o->weld;
  const double *anchor = W + 1, *qfix = W + 4, *solref = W + 8, *solimp = W + 10; double ts = W[15], invw[2] = {W[16], W[17]};
  double p[3], t[3]; mat_vec(t, o->xmat, anchor); for (int k = 0; k < 3; k++) p[k] = o->xpos[k] + t[k];
  double q[4]; quat_mul(q, o->xquat, qfix);
  double cpos[6] = {p[0], p[1], p[2], ts * q[1], ts * q[2], ts * q[3]};
  double* jacp = dalloc(3 * nv); jac_point(o, jacp, 0, p);
  double tc = fmax(solref[0], 2 * o->dt), dmax = fmin(MAXIMP, fmax(MINIMP, solimp[1]));
  double K = 1.0 / (dmax * dmax * tc * tc * solref[1] * solref[1]), B = 2.0 / (dmax * tc);
  for (int r = 0; r < 6; r++) {
    int e = o->nefc++; double* J = o->efc_J + (size_t)e * nv; memset(J, 0, sizeof(double) * nv);
    if (r < 3) memcpy(J, jacp + r * nv, sizeof(double) * nv);
    else for (int d = 0; d < o->body_dofnum[0]; d++) {       /* only the free joint's dofs rotate body 0 */
      const double* w = o->cdof + 6 * d; double c[3]; cross3(c, w, q + 1);
      J[d] = 0.5 * ts * (q[0] * w[r - 3] + c[r - 3]);
    }
    double imp = impedance_of(solimp, cpos[r]);
    o->efc_pos[e] = cpos[r]; o->efc_R[e] = fmax(MINVAL, (1 - imp) * invw[r >= 3] / imp); o->efc_D[e] = 1.0 / o->efc_R[e];
    o->efc_vel[e] = dotn(J, o->qvel, nv);
    o->efc_aref[e] = -B * o->efc_vel[e] - K * imp * cpos[r];
  }
  free(jacp);
}

/* mj_makeConstraint (equality first, then contacts, pyramidal condim 3) + mj_makeImpedance + reference acceleration */
static void make_constraint(nmfo* o) {
  int nv = o->nv; o->nefc = 0; o->neq = 0;
  if (o->weld) { make_weld(o); o->neq = o->nefc; }
  double* jacp = dalloc(3 * nv);
  double tc = fmax(o->solref[0], 2 * o->dt), dampratio = o->solref[1];
  double dmax = fmin(MAXIMP, fmax(MINIMP, o->solimp[1]));
  double K = 1.0 / (dmax * dmax * tc * tc * dampratio * dampratio), B = 2.0 / (dmax * tc);
  double mu = o->mu;
  for (int c = 0; c < o->ncon; c++) {
    int b = o->geom_body[o->con_geom[c]]; const double* f = o->con_frame + 9 * c;
    jac_point(o, jacp, b, o->con_pos + 3 * c);
    double pos = o->con_dist[c], marg = o->margin - o->gap;
    double imp = impedance(o, pos - marg);
    /* diagApprox: translational inverse weight of the two bodies (world = 0); pyramidal rows (1+mu^2) */
    double tran = o->body_invweight0[2 * b];
    double diag0 = tran + mu * mu * tran;
    double R0 = fmax(MINVAL, (1 - imp) * diag0 / imp);
    double mucon = mu / sqrt(o->impratio);      /* pyramidal: con->mu */
    double Rpy = 2 * mucon * mucon * R0;
    for (int r = 0; r < 4; r++) {
      int e = o->nefc++; double* J = o->efc_J + (size_t)e * nv;
      const double* t = f + 3 * (1 + r / 2); double sg = (r & 1) ? -1.0 : 1.0;
      for (int d = 0; d < nv; d++) {
        double jn = f[0] * jacp[d] + f[1] * jacp[nv + d] + f[2] * jacp[2 * nv + d];
        double jt = t[0] * jacp[d] + t[1] * jacp[nv + d] + t[2] * jacp[2 * nv + d];
        J[d] = jn + sg * mu * jt;
      }
      o->efc_pos[e] = pos; o->efc_R[e] = Rpy; o->efc_D[e] = 1.0 / Rpy;
      o->efc_vel[e] = dotn(J, o->qvel, nv);
      o->efc_aref[e] = -B * o->efc_vel[e] - K * imp * (pos - marg);
    }
  }
  free(jacp);
}

/* mj_transmission, adhesion (body) actuators: moment = -(1/n) sum of contact-normal Jacobians */
static void transmission_adhesion(nmfo* o) {
  int nv = o->nv; double* jacp = dalloc(3 * nv);
  memset(o->adh_moment, 0, sizeof(double) * o->nu_adh * nv);
  for (int a = 0; a < o->nu_adh; a++) {
    int cnt = 0; double* mom = o->adh_moment + (size_t)a * nv;
    for (int c = 0; c < o->ncon; c++) {
      int b = o->geom_body[o->con_geom[c]]; if (b != o->adh_body[a]) continue;
      const double* f = o->con_frame + 9 * c; jac_point(o, jacp, b, o->con_pos + 3 * c);
      for (int d = 0; d < nv; d++) mom[d] -= f[0] * jacp[d] + f[1] * jacp[nv + d] + f[2] * jacp[2 * nv + d];
      cnt++;
    }
    if (cnt) for (int d = 0; d < nv; d++) mom[d] /= cnt;
  }
  free(jacp);
}

/* ------------------------------------------------------------------ velocity stage */
static void com_vel(nmfo* o) {
  int nb = o->nbody;
  for (int b = 0; b < nb; b++) {
    double* cv = o->cvel + 6 * b; int p = o->body_parent[b];
    if (p < 0) memset(cv, 0, 48); else memcpy(cv, o->cvel + 6 * p, 48);
    int adr = o->body_dofadr[b], n = o->body_dofnum[b], j = 0;
    if (p < 0) {  /* free joint: translations first (cdof_dot = 0), then rotations against the updated velocity */
      for (int k = 0; k < 3; k++) { memset(o->cdof_dot + 6 * k, 0, 48); for (int i = 0; i < 6; i++) cv[i] += o->cdof[6 * k + i] * o->qvel[k]; }
      for (int k = 3; k < 6; k++) cross_motion(o->cdof_dot + 6 * k, cv, o->cdof + 6 * k);
      for (int k = 3; k < 6; k++) for (int i = 0; i < 6; i++) cv[i] += o->cdof[6 * k + i] * o->qvel[k];
      j = 6;
    }
    for (; j < n; j++) {
      int d = adr + j; cross_motion(o->cdof_dot + 6 * d, cv, o->cdof + 6 * d);
      for (int i = 0; i < 6; i++) cv[i] += o->cdof[6 * d + i] * o->qvel[d];
    }
  }
}
static void passive(nmfo* o) {
  for (int d = 0; d < o->nv; d++)
    o->qfrc_passive[d] = d < 6 ? 0.0 : -o->dof_stiffness[d] * (o->qpos[d + 1] - o->dof_springref[d]) - o->dof_damping[d] * o->qvel[d];
}
/* mj_rne(flg_acc=0): bias forces incl. gravity */
static void rne(nmfo* o) {
  int nb = o->nbody, nv = o->nv;
  for (int b = 0; b < nb; b++) {
    double* ca = o->cacc + 6 * b; int p = o->body_parent[b];
    if (p < 0) { ca[0] = ca[1] = ca[2] = 0; ca[3] = -o->grav[0]; ca[4] = -o->grav[1]; ca[5] = -o->grav[2]; }
    else memcpy(ca, o->cacc + 6 * p, 48);
    int adr = o->body_dofadr[b];
    for (int j = 0; j < o->body_dofnum[b]; j++) for (int i = 0; i < 6; i++) ca[i] += o->cdof_dot[6 * (adr + j) + i] * o->qvel[adr + j];
    double t1[6], t2[6], t3[6];
    mul_inert_vec(t1, o->cinert + 10 * b, ca);
    mul_inert_vec(t2, o->cinert + 10 * b, o->cvel + 6 * b);
    cross_force(t3, o->cvel + 6 * b, t2);
    for (int i = 0; i < 6; i++) o->cfrc[6 * b + i] = t1[i] + t3[i];
  }
  for (int b = nb - 1; b > 0; b--) { int p = o->body_parent[b]; for (int i = 0; i < 6; i++) o->cfrc[6 * p + i] += o->cfrc[6 * b + i]; }
  for (int d = 0; d < nv; d++) o->qfrc_bias[d] = dotn(o->cdof + 6 * d, o->cfrc + 6 * o->dof_body[d], 6);
}

/* mj_fwdActuation: position actuators (joint transmission, gear 1) + adhesion */
static void actuation(nmfo* o) {
  int nv = o->nv; memset(o->qfrc_actuator, 0, sizeof(double) * nv);
  for (int a = 0; a < o->nu_pos; a++) {
    int d = o->act_dof[a];
    double f = o->act_kp[a] * o->ctrl[a] - o->act_kp[a] * o->qpos[d + 1] - o->act_kv[a] * o->qvel[d];
    f = fmin(o->act_frcrange[2 * a + 1], fmax(o->act_frcrange[2 * a], f));
    o->actuator_force[a] = f; o->qfrc_actuator[d] += f;
  }
  for (int a = 0; a < o->nu_adh; a++) {
    double c = fmin(o->adh_ctrlrange[2 * a + 1], fmax(o->adh_ctrlrange[2 * a], o->ctrl[o->nu_pos + a]));
    double f = o->adh_gain[a] * c; o->actuator_force[o->nu_pos + a] = f;
    for (int d = 0; d < nv; d++) o->qfrc_actuator[d] += o->adh_moment[(size_t)a * nv + d] * f;
  }
}

/* ------------------------------------------------------------------ constraint solver (Newton, primal) */
static double constraint_update(nmfo* o, const double* jar, double* force, int* active) {
  double cost = 0;
  for (int e = 0; e < o->nefc; e++) {
    if (jar[e] < 0 || e < o->neq) { force[e] = -o->efc_D[e] * jar[e]; cost += 0.5 * o->efc_D[e] * jar[e] * jar[e]; if (active) active[e] = 1; }
    else { force[e] = 0; if (active) active[e] = 0; }
  }
  return cost;
}
static void mul_M(const nmfo* o, double* r, const double* v) { int nv = o->nv; for (int i = 0; i < nv; i++) r[i] = dotn(o->M + (size_t)i * nv, v, nv); }

static void solve_constraints(nmfo* o) {
  int nv = o->nv, ne = o->nefc; o->solver_niter = 0;
  if (!ne) { memcpy(o->qacc, o->qacc_smooth, sizeof(double) * nv); memset(o->qfrc_constraint, 0, sizeof(double) * nv); return; }
  double *Ma = dalloc(nv), *jar = dalloc(ne), *grad = dalloc(nv), *search = dalloc(nv), *Mv = dalloc(nv), *jv = dalloc(ne);
  double *H = dalloc((size_t)nv * nv), *tmpf = dalloc(ne);
  /* warm start (mj_fwdConstraint) */
  double cost_smooth, cost_warm;
  for (int e = 0; e < ne; e++) jar[e] = dotn(o->efc_J + (size_t)e * nv, o->qacc_smooth, nv) - o->efc_aref[e];
  cost_smooth = constraint_update(o, jar, tmpf, NULL);
  memcpy(o->qacc, o->qacc_warmstart, sizeof(double) * nv);
  mul_M(o, Ma, o->qacc);
  for (int e = 0; e < ne; e++) jar[e] = dotn(o->efc_J + (size_t)e * nv, o->qacc, nv) - o->efc_aref[e];
  cost_warm = constraint_update(o, jar, tmpf, NULL);
  for (int i = 0; i < nv; i++) cost_warm += 0.5 * (Ma[i] - o->qfrc_smooth[i]) * (o->qacc[i] - o->qacc_smooth[i]);
  if (cost_warm > cost_smooth) memcpy(o->qacc, o->qacc_smooth, sizeof(double) * nv);

  double scale = 1.0 / (o->meaninertia * (nv > 1 ? nv : 1));
  mul_M(o, Ma, o->qacc);
  for (int e = 0; e < ne; e++) jar[e] = dotn(o->efc_J + (size_t)e * nv, o->qacc, nv) - o->efc_aref[e];
  double cost = 0, gradnorm = 0;
  for (int iter = 0;; iter++) {
    /* update constraint state, forces, cost, gradient */
    double oldcost = cost;
    cost = constraint_update(o, jar, o->efc_force, o->efc_active);
    for (int i = 0; i < nv; i++) cost += 0.5 * (Ma[i] - o->qfrc_smooth[i]) * (o->qacc[i] - o->qacc_smooth[i]);
    for (int i = 0; i < nv; i++) { double s = 0; for (int e = 0; e < ne; e++) s += o->efc_J[(size_t)e * nv + i] * o->efc_force[e]; o->qfrc_constraint[i] = s; }
    gradnorm = 0;
    for (int i = 0; i < nv; i++) { grad[i] = Ma[i] - o->qfrc_smooth[i] - o->qfrc_constraint[i]; gradnorm += grad[i] * grad[i]; }
    gradnorm = sqrt(gradnorm);
    o->solver_niter = iter; o->solver_cost = cost; o->solver_gradnorm = scale * gradnorm;
    if (iter > 0 && (scale * (oldcost - cost) < o->tolerance || scale * gradnorm < o->tolerance)) break;
    if (iter >= o->iterations) break;
    /* Newton direction: H = M + J' D_active J */
    memcpy(H, o->M, sizeof(double) * nv * nv);
    for (int e = 0; e < ne; e++) if (o->efc_active[e]) {
      const double* J = o->efc_J + (size_t)e * nv; double D = o->efc_D[e];
      for (int i = 0; i < nv; i++) if (J[i] != 0) for (int j = 0; j <= i; j++) H[i * nv + j] += D * J[i] * J[j];
    }
    if (chol_factor(H, nv)) { snprintf(o->err, sizeof o->err, "Newton Hessian not PD"); break; }
    for (int i = 0; i < nv; i++) search[i] = -grad[i];
    chol_solve(H, nv, search);
    /* exact line search on the piecewise-quadratic cost (safeguarded Newton on its derivative) */
    double snorm = sqrt(dotn(search, search, nv));
    if (snorm < MINVAL) break;
    double gtol = o->tolerance * o->ls_tolerance * snorm / scale;
    mul_M(o, Mv, search);
    for (int e = 0; e < ne; e++) jv[e] = dotn(o->efc_J + (size_t)e * nv, search, nv);
    double q1 = 0, q2 = 0;
    for (int i = 0; i < nv; i++) { q1 += search[i] * (Ma[i] - o->qfrc_smooth[i]); q2 += search[i] * Mv[i]; }
    double lo = 0, hi = INFINITY, alpha = 0;
    for (int it = 0; it < o->ls_iterations; it++) {
      double d0 = q1 + alpha * q2, d1 = q2;
      for (int e = 0; e < ne; e++) { double x = jar[e] + alpha * jv[e]; if (x < 0 || e < o->neq) { d0 += o->efc_D[e] * x * jv[e]; d1 += o->efc_D[e] * jv[e] * jv[e]; } }
      if (fabs(d0) < gtol) break;
      if (d0 < 0) lo = alpha; else hi = alpha;
      double nx = alpha - d0 / d1;
      if (nx <= lo || nx >= hi) nx = isinf(hi) ? 2 * (alpha > 0 ? alpha : 1.0) : 0.5 * (lo + hi);
      alpha = nx;
    }
    if (alpha == 0) break;
    for (int i = 0; i < nv; i++) { o->qacc[i] += alpha * search[i]; Ma[i] += alpha * Mv[i]; }
    for (int e = 0; e < ne; e++) jar[e] += alpha * jv[e];
  }
  memcpy(o->efc_jar, jar, sizeof(double) * ne);
  free(Ma); free(jar); free(grad); free(search); free(Mv); free(jv); free(H); free(tmpf);
}

/* ------------------------------------------------------------------ noslip post-solver (CPU reference only)
 * [PRIOR] mj_solNoSlip, run by mj_fwdConstraint after the main solver when opt.noslip_iterations > 0 (the reference sets 5 in
 * mujoco_globals.yaml:15 and strips it on the GPU path, warp/simulation.py:427-448).  A projected Gauss-Seidel on the DUAL
 * problem WITHOUT the regulariser R:  A = J M^-1 J',  b = J qacc_smooth - aref,  residual_i = A_i. f + b_i.
 *   - equality rows: f_i -= residual_i / A_ii                      (unclamped)
 *   - pyramidal contacts, per pair of opposing edges (j, j+1) = (n + mu t_k, n - mu t_k): the sum f_j + f_j+1 = 2 mid (the
 *     normal force carried by the pair) is kept; y = (f_j - f_j+1)/2 minimises the unregularised cost over [-mid, mid]:
 *     K1 = A00 + A11 - 2 A01,  K0 = mid (A00 - A11) + bc0 - bc1  with  bc = residual - Ac f_old,  y = -K0 / K1, clamped.
 *   - an update that would raise the cost (> 1e-10) is undone; sweeps stop when the scaled improvement < noslip_tolerance (1e-6).
 * Afterwards  qfrc_constraint = J' f,  qacc = qacc_smooth + M^-1 qfrc_constraint. */
static double cost_change2(const double* Ac, double* f, const double* old, const double* res) {
  double d0 = f[0] - old[0], d1 = f[1] - old[1];
  double change = 0.5 * (d0 * (Ac[0] * d0 + Ac[1] * d1) + d1 * (Ac[2] * d0 + Ac[3] * d1)) + d0 * res[0] + d1 * res[1];
  if (change > 1e-10) { f[0] = old[0]; f[1] = old[1]; change = 0; }
  return change;
}
static void noslip(nmfo* o) {
  int nv = o->nv, ne = o->nefc, neq = o->neq;
  o->noslip_niter = 0;
  if (!ne || o->noslip_iterations <= 0) return;
  double *L = dalloc((size_t)nv * nv), *Y = dalloc((size_t)ne * nv), *A = dalloc((size_t)ne * ne), *b = dalloc(ne), *f = o->efc_force;
  memcpy(L, o->M, sizeof(double) * nv * nv);
  if (chol_factor(L, nv)) { snprintf(o->err, sizeof o->err, "noslip: M not PD"); free(L); free(Y); free(A); free(b); return; }
  for (int e = 0; e < ne; e++) { memcpy(Y + (size_t)e * nv, o->efc_J + (size_t)e * nv, sizeof(double) * nv); chol_solve(L, nv, Y + (size_t)e * nv); }
  for (int i = 0; i < ne; i++) {
    for (int j = 0; j < ne; j++) A[(size_t)i * ne + j] = dotn(o->efc_J + (size_t)i * nv, Y + (size_t)j * nv, nv);
    b[i] = dotn(o->efc_J + (size_t)i * nv, o->qacc_smooth, nv) - o->efc_aref[i];
  }
  double scale = 1.0 / (o->meaninertia * (nv > 1 ? nv : 1));
  for (int iter = 0; iter < o->noslip_iterations; iter++) {
    double improvement = 0;
    if (iter == 0) for (int i = 0; i < ne; i++) improvement += 0.5 * f[i] * f[i] * o->efc_R[i];
    for (int i = 0; i < neq; i++) {              /* equality rows */
      double res = b[i] + dotn(A + (size_t)i * ne, f, ne), old = f[i], Aii = A[(size_t)i * ne + i];
      f[i] -= res / fmax(MINVAL, Aii);
      double d = f[i] - old, change = 0.5 * d * d * Aii + d * res;
      if (change > 1e-10) { f[i] = old; change = 0; }
      improvement -= change;
    }
    for (int j = neq; j + 1 < ne; j += 2) {      /* pairs of opposing pyramid edges */
      double res[2] = {b[j] + dotn(A + (size_t)j * ne, f, ne), b[j + 1] + dotn(A + (size_t)(j + 1) * ne, f, ne)};
      double old[2] = {f[j], f[j + 1]};
      double Ac[4] = {A[(size_t)j * ne + j], A[(size_t)j * ne + j + 1], A[(size_t)(j + 1) * ne + j], A[(size_t)(j + 1) * ne + j + 1]};
      double bc[2] = {res[0] - (Ac[0] * old[0] + Ac[1] * old[1]), res[1] - (Ac[2] * old[0] + Ac[3] * old[1])};
      double mid = 0.5 * (f[j] + f[j + 1]);
      double K1 = Ac[0] + Ac[3] - Ac[1] - Ac[2], K0 = mid * (Ac[0] - Ac[3]) + bc[0] - bc[1];
      if (K1 < MINVAL) { f[j] = f[j + 1] = mid; }
      else {
        double y = -K0 / K1;
        if (y < -mid) { f[j] = 0; f[j + 1] = 2 * mid; }
        else if (y > mid) { f[j] = 2 * mid; f[j + 1] = 0; }
        else { f[j] = mid + y; f[j + 1] = mid - y; }
      }
      improvement -= cost_change2(Ac, f + j, old, res);
    }
    o->noslip_niter = iter + 1;
    if (improvement * scale < o->noslip_tolerance) break;
  }
  for (int i = 0; i < nv; i++) { double sum = 0; for (int e = 0; e < ne; e++) sum += o->efc_J[(size_t)e * nv + i] * f[e]; o->qfrc_constraint[i] = sum; }
  memcpy(o->qacc, o->qfrc_constraint, sizeof(double) * nv); chol_solve(L, nv, o->qacc);
  for (int i = 0; i < nv; i++) o->qacc[i] += o->qacc_smooth[i];
  for (int e = 0; e < ne; e++) o->efc_jar[e] = dotn(o->efc_J + (size_t)e * nv, o->qacc, nv) - o->efc_aref[e];
  free(L); free(Y); free(A); free(b);
}

/* contact sensors (world.py:311-331): per leg, reduce=netforce over contacts of the leg subtree vs ground.
 * [PRIOR, unverified] layout: found, force(3), torque(3), pos(3), normal(3), tangent(3); net wrench expressed in
 * the world frame (contact frame of the synthetic net contact = identity), force = leg-on-ground. */
static void sensors(nmfo* o) {
  memset(o->sensordata, 0, sizeof(double) * 16 * o->nleg);
  for (int l = 0; l < o->nleg; l++) {
    double* s = o->sensordata + 16 * l; double F[3] = {0, 0, 0}, P[3] = {0, 0, 0}, wsum = 0; int found = 0;
    for (int c = 0; c < o->ncon; c++) {
      if (o->body_leg[o->geom_body[o->con_geom[c]]] != l) continue;
      const double* f = o->con_frame + 9 * c; const double* ef = o->efc_force + o->neq + 4 * c; double fc[3];
      double fn = ef[0] + ef[1] + ef[2] + ef[3], f1 = o->mu * (ef[0] - ef[1]), f2 = o->mu * (ef[2] - ef[3]);
      for (int k = 0; k < 3; k++) fc[k] = fn * f[k] + f1 * f[3 + k] + f2 * f[6 + k];
      for (int k = 0; k < 3; k++) { F[k] += fc[k]; P[k] += fn * o->con_pos[3 * c + k]; }
      wsum += fn; found++;
    }
    s[0] = found;
    if (!found) continue;
    if (wsum > MINVAL) for (int k = 0; k < 3; k++) P[k] /= wsum;
    else { int n = 0; P[0] = P[1] = P[2] = 0; for (int c = 0; c < o->ncon; c++) if (o->body_leg[o->geom_body[o->con_geom[c]]] == l) { for (int k = 0; k < 3; k++) P[k] += o->con_pos[3 * c + k]; n++; } for (int k = 0; k < 3; k++) P[k] /= n; }
    double T[3] = {0, 0, 0};
    for (int c = 0; c < o->ncon; c++) {
      if (o->body_leg[o->geom_body[o->con_geom[c]]] != l) continue;
      const double* f = o->con_frame + 9 * c; const double* ef = o->efc_force + o->neq + 4 * c; double fc[3], r[3], t[3];
      double fn = ef[0] + ef[1] + ef[2] + ef[3], f1 = o->mu * (ef[0] - ef[1]), f2 = o->mu * (ef[2] - ef[3]);
      for (int k = 0; k < 3; k++) { fc[k] = fn * f[k] + f1 * f[3 + k] + f2 * f[6 + k]; r[k] = o->con_pos[3 * c + k] - P[k]; }
      cross3(t, r, fc); for (int k = 0; k < 3; k++) T[k] += t[k];
    }
    for (int k = 0; k < 3; k++) { s[1 + k] = -F[k]; s[4 + k] = -T[k]; s[7 + k] = P[k]; }
    s[10] = 1; s[14] = 1;  /* normal = (1,0,0), tangent = (0,1,0): identity frame */
  }
}

/* mj_energyPos / mj_energyVel (the reference model enables the `energy` flag, mujoco_globals.yaml:19):
 * potential = -sum_b m_b g.xipos_b + sum_hinges 1/2 k (q - springref)^2 ;  kinetic = 1/2 qvel' M qvel */
static void energy(nmfo* o) {
  int nv = o->nv; double ep = 0, ek = 0;
  for (int b = 0; b < o->nbody; b++) ep -= o->body_mass[b] * dot3(o->grav, o->xipos + 3 * b);
  for (int d = 6; d < nv; d++) { double dq = o->qpos[d + 1] - o->dof_springref[d]; ep += 0.5 * o->dof_stiffness[d] * dq * dq; }
  for (int i = 0; i < nv; i++) ek += 0.5 * o->qvel[i] * dotn(o->M + (size_t)i * nv, o->qvel, nv);
  o->energy[0] = ep; o->energy[1] = ek;
}

/* ------------------------------------------------------------------ mj_forward / mj_step */
void nmfo_forward(nmfo* o) {
  int nv = o->nv;
  kinematics(o); com_pos(o); crb(o); collision(o); make_constraint(o); transmission_adhesion(o);
  com_vel(o); passive(o); rne(o); actuation(o);
  for (int i = 0; i < nv; i++) o->qfrc_smooth[i] = o->qfrc_passive[i] - o->qfrc_bias[i] + o->qfrc_actuator[i];
  double* L = dalloc((size_t)nv * nv); memcpy(L, o->M, sizeof(double) * nv * nv);
  if (chol_factor(L, nv)) snprintf(o->err, sizeof o->err, "M not PD");
  memcpy(o->qacc_smooth, o->qfrc_smooth, sizeof(double) * nv); chol_solve(L, nv, o->qacc_smooth);
  free(L);
  solve_constraints(o);
  noslip(o);
  sensors(o);
  energy(o);
}

/* mj_Euler with implicit joint damping (eulerdamp) + mj_advance */
void nmfo_step(nmfo* o) {
  int nv = o->nv;
  nmfo_forward(o);
  double* L = dalloc((size_t)nv * nv); double* a = dalloc(nv);
  memcpy(L, o->M, sizeof(double) * nv * nv);
  for (int i = 0; i < nv; i++) { L[i * nv + i] += o->dt * o->dof_damping[i]; a[i] = o->qfrc_smooth[i] + o->qfrc_constraint[i]; }
  chol_factor(L, nv); chol_solve(L, nv, a);
  memcpy(o->qacc_warmstart, o->qacc, sizeof(double) * nv);
  for (int i = 0; i < nv; i++) o->qvel[i] += o->dt * a[i];
  for (int k = 0; k < 3; k++) o->qpos[k] += o->dt * o->qvel[k];
  double w[3] = {o->qvel[3], o->qvel[4], o->qvel[5]}, n = sqrt(dot3(w, w));
  if (n > MINVAL) {  /* mju_quatIntegrate: q <- q * exp(dt*w/2), w in the body frame */
    double ax[3] = {w[0] / n, w[1] / n, w[2] / n}, dq[4], nq[4];
    axis_angle_quat(dq, ax, o->dt * n); quat_mul(nq, o->qpos + 3, dq); memcpy(o->qpos + 3, nq, 32);
  }
  quat_norm(o->qpos + 3);
  for (int d = 6; d < nv; d++) o->qpos[d + 1] += o->dt * o->qvel[d];
  o->time += o->dt;
  free(L); free(a);
}

void nmfo_step_n(nmfo* o, int n) { for (int i = 0; i < n; i++) nmfo_step(o); }
/* n steps with a per-step table [n][cols] written into ctrl[0:cols] (cols = nu_pos, or nu for position + adhesion inputs) */
void nmfo_step_table_cols(nmfo* o, const double* table, int n, int cols) {
  for (int i = 0; i < n; i++) { memcpy(o->ctrl, table + (size_t)i * cols, sizeof(double) * cols); nmfo_step(o); }
}
/* n steps with a per-step action table [n][nu_pos] written into ctrl[0:nu_pos] */
void nmfo_step_table(nmfo* o, const double* table, int n) {
  for (int i = 0; i < n; i++) { memcpy(o->ctrl, table + (size_t)i * o->nu_pos, sizeof(double) * o->nu_pos); nmfo_step(o); }
}

/* ------------------------------------------------------------------ accessors for the test harness */
int nmfo_dim(const nmfo* o, const char* name) {
  if (!strcmp(name, "nq")) return o->nq; if (!strcmp(name, "nv")) return o->nv; if (!strcmp(name, "nu")) return o->nu;
  if (!strcmp(name, "nbody")) return o->nbody; if (!strcmp(name, "ncon")) return o->ncon; if (!strcmp(name, "nefc")) return o->nefc;
  if (!strcmp(name, "ngeom")) return o->ngeom; if (!strcmp(name, "nsite")) return o->nsite; if (!strcmp(name, "nseg")) return o->nseg;
  if (!strcmp(name, "nleg")) return o->nleg; if (!strcmp(name, "solver_niter")) return o->solver_niter;
  if (!strcmp(name, "noslip_niter")) return o->noslip_niter;
  return -1;
}
/* returns pointer + element count of a named double array (NULL if unknown) */
double* nmfo_array(nmfo* o, const char* name, int* count) {
  int nv = o->nv, nb = o->nbody;
#define A(nm, ptr, cnt) if (!strcmp(name, nm)) { if (count) *count = (cnt); return (ptr); }
  A("qpos", o->qpos, o->nq) A("qvel", o->qvel, nv) A("ctrl", o->ctrl, o->nu) A("qacc_warmstart", o->qacc_warmstart, nv)
  A("time", &o->time, 1) A("xpos", o->xpos, 3 * nb) A("xquat", o->xquat, 4 * nb) A("xipos", o->xipos, 3 * nb)
  A("com", o->com, 3) A("cinert", o->cinert, 10 * nb) A("cdof", o->cdof, 6 * nv) A("M", o->M, nv * nv)
  A("qfrc_bias", o->qfrc_bias, nv) A("qfrc_passive", o->qfrc_passive, nv) A("qfrc_actuator", o->qfrc_actuator, nv)
  A("actuator_force", o->actuator_force, o->nu) A("qfrc_smooth", o->qfrc_smooth, nv) A("qacc_smooth", o->qacc_smooth, nv)
  A("qacc", o->qacc, nv) A("qfrc_constraint", o->qfrc_constraint, nv) A("con_dist", o->con_dist, o->ncon)
  A("con_pos", o->con_pos, 3 * o->ncon) A("con_frame", o->con_frame, 9 * o->ncon) A("efc_J", o->efc_J, o->nefc * nv)
  A("efc_force", o->efc_force, o->nefc) A("efc_aref", o->efc_aref, o->nefc) A("efc_D", o->efc_D, o->nefc)
  A("efc_jar", o->efc_jar, o->nefc) A("efc_pos", o->efc_pos, o->nefc)
  A("site_xpos", o->site_xpos, 3 * o->nsite) A("seg_xpos", o->seg_xpos, 3 * o->nseg) A("seg_xquat", o->seg_xquat, 4 * o->nseg)
  A("sensordata", o->sensordata, 16 * o->nleg) A("geom_xpos", o->geom_xpos, 3 * o->ngeom) A("cvel", o->cvel, 6 * nb)
  A("solver_gradnorm", &o->solver_gradnorm, 1) A("solver_cost", &o->solver_cost, 1) A("energy", o->energy, 2)
#undef A
  return NULL;
}
int nmfo_con_geom(const nmfo* o, int* out, int cap) { int n = o->ncon < cap ? o->ncon : cap; memcpy(out, o->con_geom, sizeof(int) * n); return o->ncon; }
const char* nmfo_last_error(const nmfo* o) { return o->err; }
