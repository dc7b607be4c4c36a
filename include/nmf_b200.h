/*
 * nmf_b200.h — C ABI of the B200-native NeuroMechFly step path (libnmf_b200.so).
 *
 * The reference (flygym 2.0.1) has no FFI of its own for this path: the backend is
 * swapped by subclassing `flygym.Simulation` the way `flygym.warp.GPUSimulation`
 * does (reference src/flygym/warp/simulation.py:28-71,213-263).  The entry points
 * below are what such a subclass binds through ctypes; each one names the reference
 * method(s) it stands behind.  Plain pointers and sizes only; all `float*` buffers
 * marked DEVICE are borrowed device pointers (e.g. torch.Tensor.data_ptr()), HOST
 * buffers are ordinary (preferably pinned) host memory.  Every call returns 0 on
 * success or a negative nmf_status; nmf_last_error() gives the message.
 */
#ifndef NMF_B200_H
#define NMF_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nmf_handle nmf_handle;

enum nmf_status {
  NMF_OK = 0,
  NMF_EINVAL = -1,     /* bad argument / unsupported model topology */
  NMF_ECUDA = -2,      /* CUDA runtime error */
  NMF_ENOTBOUND = -3   /* nmf_bind() has not been called */
};

/* Offsets (in floats) of the sections of one fly's state record, and model sizes. */
typedef struct nmf_info {
  int32_t n_flies, nq, nv, nu_pos, nu_adh, nseg, nleg;
  int32_t state_stride, off_qpos, off_qvel, off_qacc_warmstart, off_ctrl, off_time;
  int32_t dbg_stride;
  float timestep;
  int32_t off_status;   /* per-fly status word (a small integer stored as a float), OR of NMF_ST_* bits; sticky until the fly is reset */
} nmf_info;

/* Device-side faults are reported per fly through the status word of the state record, never by trapping. */
enum nmf_fly_status {
  NMF_ST_NONFINITE = 1,    /* a generalised velocity became NaN / infinite */
  NMF_ST_NEWTON_CAP = 2,   /* the Newton solver hit `iterations` (mujoco_globals.yaml:14) with the active set still changing */
  NMF_ST_LS_CAP = 4,       /* a line search used up its evaluation cap (MuJoCo's ls_iterations) */
  NMF_ST_NOSLIP_SKIP = 8   /* more than 48 simultaneous contacts: the step ran without the noslip pass (noslip models only) */
};

/* Device buffers owned by the caller (PyTorch tensors in the Python host). `state`
 * is required; the observation buffers are optional (NULL = not produced). */
typedef struct nmf_buffers {
  float* state;        /* DEVICE [n_flies][state_stride]  qpos|qvel|qacc_warmstart|ctrl|time */
  float* seg_xpos;     /* DEVICE [n_flies][nseg][3]   -> Simulation.get_body_positions / get_site_positions */
  float* seg_xquat;    /* DEVICE [n_flies][nseg][4]   -> Simulation.get_body_rotations (w,x,y,z) */
  float* act_force;    /* DEVICE [n_flies][nu]        -> Simulation.get_actuator_forces */
  float* sensordata;   /* DEVICE [n_flies][nleg*16]   -> Simulation.get_ground_contact_info */
  float* debug;        /* DEVICE [n_flies][dbg_stride] solver internals for the parity tests (optional) */
  float* energy;       /* DEVICE [n_flies][2]         potential, kinetic energy (the model's `energy` flag, mujoco_globals.yaml:19) */
} nmf_buffers;

/* Simulation.__init__ / GPUSimulation.__init__ (simulation.py:32-57, warp/simulation.py:49-62):
 * ingest the compiled model (blob from flygym_b200.model.NMFModel.to_blob()). */
int nmf_create(const void* model_blob, size_t nbytes, int n_flies, int device, nmf_handle** out);
int nmf_destroy(nmf_handle* h);
int nmf_model_info(const nmf_handle* h, nmf_info* info);
const char* nmf_last_error(const nmf_handle* h);

int nmf_bind(nmf_handle* h, const nmf_buffers* buffers);

/* Simulation.reset / GPUSimulation.reset (simulation.py:59-72, warp/simulation.py:64-71): every fly (or the
 * flies whose byte in the DEVICE mask is non-zero) <- keyframe "neutral". */
int nmf_reset(nmf_handle* h, const uint8_t* mask_or_null, void* cuda_stream);

/* GPUSimulation.step (warp/simulation.py:260-263), `nsteps` times inside one launch.  With an action table
 * (DEVICE [n_flies][table_T][table_cols], as the reference benchmark keeps it: time_gpu_simulation.py:89-98,133-146)
 * step s copies row (table_t0 + s) % table_T into ctrl[0:table_cols]; table_cols = nu_pos (position targets only, the
 * reference's table) or nu_pos + nu_adh (position targets followed by the six adhesion inputs); NULL = use ctrl in the state. */
int nmf_step(nmf_handle* h, int nsteps, const float* action_table_or_null, int table_T, int table_t0, int table_cols, void* cuda_stream);

/* mj_forward for every fly: evaluates the current state (segment poses, actuator forces, contact sensors into the bound
 * observation buffers) without advancing it.  The reference reaches this through mj_data after mj_forward / the first step
 * following Simulation.reset (simulation.py:59-72 resets without forward). */
int nmf_forward(nmf_handle* h, void* cuda_stream);

/* Scheduling of multi-step launches: with more flies than the GPU holds resident fly slots, a launch of nsteps >= 2*sub_steps
 * is cut into (fly, sub_steps-step) work items served from a device-side queue (results are identical; only the order in
 * which flies advance changes).  sub_steps = 0 disables the queue (one block per fly for the whole launch); -1 (default)
 * lets the library pick the sub-chunk length that fills whole waves of resident blocks best. */
int nmf_set_schedule(nmf_handle* h, int sub_steps);

/* Fly slots per thread block of the float32 kernels: 1, 2, 4, 8, or 0 (default) = chosen per launch from the batch size.  The
 * step is bound by instruction fetch, so the slots of a block advance in stages separated by a block-wide barrier (a stage =
 * the kinematics .. smooth-force part of a step, or one solver pass): the block's warps then stream the same code at the same
 * time and share its fetches, while every slot still takes its own work items and nobody waits for a neighbour's Newton
 * iterations.  Results are bit-identical for every setting. */
int nmf_set_flies_per_block(nmf_handle* h, int flies_per_block);

/* set_actuator_inputs / set_leg_adhesion_states (warp/simulation.py:213-258; kernel warp/utils.py:84-104):
 * state.ctrl[:, cols[k]] = src[:, k]   (src DEVICE [n_flies][ncols], cols DEVICE int32[ncols]) */
int nmf_scatter_ctrl(nmf_handle* h, const float* src, const int32_t* cols, int ncols, void* cuda_stream);
/* get_joint_angles / get_joint_velocities (warp/simulation.py:73-115; kernel warp/utils.py:107-127):
 * dst[:, k] = state[:, section_offset + cols[k]] */
int nmf_gather_state(nmf_handle* h, int section_offset, const int32_t* cols, int ncols, float* dst, void* cuda_stream);

/* Host-buffer convenience used for end-to-end timing: copies `actions` (HOST [n_flies][action_cols], action_cols = nu_pos or
 * nu_pos + nu_adh) to the device, steps `nsteps`, copies qpos (HOST [n_flies][nq]) back.  Synchronises the stream.  Large batches
 * are cut into up to four slices pipelined over streams owned by the handle (copies of one slice overlap kernels of the others).
 * With page-locked (pinned / registered) host buffers the pipeline is captured once as a CUDA graph and replayed, the host addresses
 * of the call patched into its copy nodes; pageable buffers, the f64 build and a failed capture take the call-by-call path. */
int nmf_step_host(nmf_handle* h, const float* actions_host, int action_cols, int nsteps, float* qpos_host, void* cuda_stream);

int nmf_set_solver(nmf_handle* h, int max_newton_iterations, int max_linesearch_iterations);

/* Arithmetic of the step kernels: 32 (default; the product path, what GPUSimulation / MuJoCo-Warp compute in) or 64 = the same
 * kernel source instantiated in double precision (what Simulation / MuJoCo's mjtNum computes in).  The 64-bit build is a
 * validation path: it shadows the fp64 oracle over long horizons and so separates algorithmic differences from float32
 * round-off.  The records the API sees stay float32; the library keeps full-precision copies between launches and takes an
 * entry from the float record only where it was edited through the API since (reset, setters, direct writes). */
int nmf_set_precision(nmf_handle* h, int bits);

/* Number of kernels this library has launched on behalf of the handle (bench.py's gpu_launches). */
int64_t nmf_launch_count(const nmf_handle* h);

/* The reference benchmark's action table built on the device: MotionSnippet.get_joint_angles' cubic resampling of the recorded
 * clip (src/flygym_demo/spotlight_data/preprocessing.py:80-142) tiled per world as ReplayTargetData.make_target_angles_all_worlds
 * does (src/flygym_demo/benchmark/time_gpu_simulation.py:73-86; world k replays partition (k + fly_offset) % n_part).
 * coef HOST double[4][n_int][A]: piecewise-cubic coefficients of the spline on the uniform source grid i / fps, highest power
 * first (scipy PPoly layout); last HOST double[A]: the final source sample (interp1d's fill value beyond the grid);
 * out DEVICE float[n_worlds][T][A].  Synchronises the stream. */
int nmf_replay_table(const double* coef, const double* last, int n_int, int A, double fps, double dt, int n_part, int T,
                     int n_worlds, int fly_offset, float* out, void* cuda_stream);

/* ------------------------------------------------------------------------------------------------------------
 * Retina transform and odor-intensity sensor.  FlyGym 2.0.1 ships no implementation (only the v1 parameter block at
 * src/flygym/assets/model/legacy/flygym1_config.yaml:141-200); these entry points are what a re-added
 * `flygym.vision.Retina.raw_image_to_hex_pxls` / odor observation would bind.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct nmf_retina nmf_retina;

/* pixcode HOST int16 [2 eyes][H*W]: 0 = pixel outside every ommatidium, else 2*bin + (channel==blue), bin in 1..n_omm.
 * inv_norm HOST float [2 eyes][n_omm+1][2]: 1/(255*pixel_count) in the channel slot the ommatidium reads, 0 elsewhere. */
int nmf_retina_create(const int16_t* pixcode, const float* inv_norm, int H, int W, int n_omm, int device, nmf_retina** out);
int nmf_retina_destroy(nmf_retina* r);
const char* nmf_retina_last_error(const nmf_retina* r);
int64_t nmf_retina_launch_count(const nmf_retina* r);
/* images DEVICE uint8 [n_flies][2][H][W][3] (16-byte aligned) -> out DEVICE float [n_flies][2][n_omm][2] */
int nmf_retina_forward(nmf_retina* r, const uint8_t* images, int n_flies, float* out, void* cuda_stream);
/* same with HOST buffers (H2D, kernel, D2H, synchronises) */
int nmf_retina_forward_host(nmf_retina* r, const uint8_t* images_host, int n_flies, float* out_host, void* cuda_stream);

/* Eye-camera image formation (v1 camera block flygym1_config.yaml:141-174).  Plain-data
 * parameters, passed by pointer from HOST memory. */
typedef struct nmf_eye_params {
  int32_t eye_seg[2];      /* segment index (in the model's segment order) of l_eye, r_eye */
  float rel_pos[6];        /* camera position in the eye segment frame, per eye */
  float R_local[18];       /* camera-to-segment rotation (row-major 3x3), per eye; the camera looks along its -z, +y is up */
  float cx, cy, inv_f;     /* principal point (pixels) and 1/focal length (1/pixels): f = (H/2)/tan(fovy/2) */
  float inv_check;         /* 1 / checker square size (1/mm) */
  uint32_t ground_lo, ground_hi, sky_g, sky_b;   /* 8-bit colours: the two checker greys, sky green / blue */
  uint32_t body_g, body_b;                       /* 8-bit colour of the fly's own body (nmf_eye_set_body) */
} nmf_eye_params;
/* The fly's own body as the eye cameras see it: ncap <= 64 capsules, capsule k rigidly attached to segment seg[k] with end points
 * cap_a[k], cap_b[k] (segment frame) and radius[k]; HOST arrays, copied.  The visible segments are all but the v1 hidden list
 * (flygym1_config.yaml:148-162).  ncap = 0 (default): ground and sky only. */
int nmf_eye_set_body(nmf_retina* r, const int32_t* seg, const float* cap_a, const float* cap_b, const float* radius, int ncap);
/* raw eye images DEVICE uint8 [n_flies][2][H][W][3] from the segment poses of the last step */
int nmf_eye_render(nmf_retina* r, const nmf_eye_params* prm, const float* seg_xpos, const float* seg_xquat, int n_flies, int nseg,
                   uint8_t* images, void* cuda_stream);
/* fused image formation + Retina (the images are never materialised): out DEVICE float [n_flies][2][n_omm][2] */
int nmf_eye_retina(nmf_retina* r, const nmf_eye_params* prm, const float* seg_xpos, const float* seg_xquat, int n_flies, int nseg,
                   float* out, void* cuda_stream);

/* odor intensity at the 4 sensor sites: out DEVICE float [n_flies][D][4];  all pointers DEVICE.
 * seg_xpos / seg_xquat are the buffers bound with nmf_bind (segment poses of the last step). */
int nmf_odor_intensity(const float* seg_xpos, const float* seg_xquat, int n_flies, int nseg, const int32_t* sensor_seg /*[4]*/,
                       const float* sensor_relpos /*[4][3]*/, const float* src_pos /*[nsrc][3]*/, const float* src_peak /*[nsrc][D]*/,
                       int nsrc, int D, float* out, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
