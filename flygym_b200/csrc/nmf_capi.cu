// nmf_capi.cu — C ABI (include/nmf_b200.h) + kernel launches for the sm_100a step path.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "../../include/nmf_b200.h"
#include "nmf_host.h"
#include "nmf_tree_host.h"
#include "nmf_step_all.cuh"

using namespace nmf;

// ------------------------------------------------------------------ kernels
#ifndef NMF_MINBLOCKS
#define NMF_MINBLOCKS 16   // <= 64 registers/thread: best measured trade-off between occupancy and spills (profiles/)
#endif
extern "C" __global__ void __launch_bounds__(CTA, NMF_MINBLOCKS) nmf_step_kernel(const StepParams p) { f32::step_entry<f32::W_FLAT>(p); }
// FPB fly slots per block, realigned at every stage so that the block's warps share instruction fetches (nmf_step.cuh, step_block)
extern "C" __global__ void __launch_bounds__(2 * CTA, NMF_MINBLOCKS / 2) nmf_step_x2_kernel(const StepParams p) { f32::step_entry<f32::W_FLAT, 2>(p); }
extern "C" __global__ void __launch_bounds__(4 * CTA, NMF_MINBLOCKS / 4) nmf_step_x4_kernel(const StepParams p) { f32::step_entry<f32::W_FLAT, 4>(p); }
extern "C" __global__ void __launch_bounds__(8 * CTA, NMF_MINBLOCKS / 8) nmf_step_x8_kernel(const StepParams p) { f32::step_entry<f32::W_FLAT, 8>(p); }
// terrain worlds (box columns: BASELINE config 3): general contact frames need 8 more registers per lane; measured on B200 at
// 12 / 14 / 16 blocks per SM (80 / 72 / 64 registers): 15.8 / 16.8 / 16.8 M env-steps/s -> occupancy wins over spills here too
#ifndef NMF_MINBLOCKS_TERRAIN
#define NMF_MINBLOCKS_TERRAIN 16
#endif
extern "C" __global__ void __launch_bounds__(CTA, NMF_MINBLOCKS_TERRAIN) nmf_step_terrain_kernel(const StepParams p) { f32::step_entry<f32::W_TERRAIN>(p); }
extern "C" __global__ void __launch_bounds__(2 * CTA, NMF_MINBLOCKS_TERRAIN / 2) nmf_step_terrain_x2_kernel(const StepParams p) { f32::step_entry<f32::W_TERRAIN, 2>(p); }
extern "C" __global__ void __launch_bounds__(4 * CTA, NMF_MINBLOCKS_TERRAIN / 4) nmf_step_terrain_x4_kernel(const StepParams p) { f32::step_entry<f32::W_TERRAIN, 4>(p); }
extern "C" __global__ void __launch_bounds__(8 * CTA, NMF_MINBLOCKS_TERRAIN / 8) nmf_step_terrain_x8_kernel(const StepParams p) { f32::step_entry<f32::W_TERRAIN, 8>(p); }
// flat world with convex-hull geoms and the `multiccd` flag: four contact slots per lane (up to 4 plane-hull contacts per geom).
// 40 of the 64 registers would be contact slots, so these run at 80 registers / 3 blocks of 4 flies per SM
#ifndef NMF_MINBLOCKS_MESH
#define NMF_MINBLOCKS_MESH 12
#endif
extern "C" __global__ void __launch_bounds__(CTA, NMF_MINBLOCKS_MESH) nmf_step_mesh_kernel(const StepParams p) { f32::step_entry<f32::W_MESH>(p); }
extern "C" __global__ void __launch_bounds__(2 * CTA, NMF_MINBLOCKS_MESH / 2) nmf_step_mesh_x2_kernel(const StepParams p) { f32::step_entry<f32::W_MESH, 2>(p); }
extern "C" __global__ void __launch_bounds__(4 * CTA, NMF_MINBLOCKS_MESH / 4) nmf_step_mesh_x4_kernel(const StepParams p) { f32::step_entry<f32::W_MESH, 4>(p); }
extern "C" __global__ void __launch_bounds__(8 * CTA, (NMF_MINBLOCKS_MESH + 7) / 8) nmf_step_mesh_x8_kernel(const StepParams p) { f32::step_entry<f32::W_MESH, 8>(p); }
// TetheredWorld (reference world.py:334-366): no ground contacts, six weld rows on the free body
extern "C" __global__ void __launch_bounds__(CTA, NMF_MINBLOCKS) nmf_step_tether_kernel(const StepParams p) { f32::step_entry<f32::W_TETHER>(p); }
// fp64 instantiations of the same source: a validation build (nmf_set_precision(h, 64)), not a product path -- every
// quantity takes two registers, so occupancy is whatever 255 registers leave
#ifndef NMF_MINBLOCKS_F64
#define NMF_MINBLOCKS_F64 4
#endif
extern "C" __global__ void __launch_bounds__(CTA, NMF_MINBLOCKS_F64) nmf_step_f64_kernel(const StepParamsT<double> p) { f64::step_entry<f64::W_FLAT>(p); }
// the f64 flat kernel with 2 / 4 flies per block in lockstep passes.  Four flies per block run at 128 registers, two blocks per SM
// (8 flies per SM, 952 B of stack): 10.4 M env-steps/s against 9.5 M for one block at 255 registers -- with fetches shared inside a
// block, occupancy wins over spills here as it does in float32 (B200, 4096 flies; one fly per block gains nothing from it: 6.9 M both ways)
#ifndef NMF_MINBLOCKS_F64_X4
#define NMF_MINBLOCKS_F64_X4 2
#endif
extern "C" __global__ void __launch_bounds__(2 * CTA, NMF_MINBLOCKS_F64 / 2) nmf_step_f64_x2_kernel(const StepParamsT<double> p) { f64::step_entry<f64::W_FLAT, 2>(p); }
extern "C" __global__ void __launch_bounds__(4 * CTA, NMF_MINBLOCKS_F64_X4) nmf_step_f64_x4_kernel(const StepParamsT<double> p) { f64::step_entry<f64::W_FLAT, 4>(p); }
extern "C" __global__ void __launch_bounds__(CTA, NMF_MINBLOCKS_F64) nmf_step_terrain_f64_kernel(const StepParamsT<double> p) { f64::step_entry<f64::W_TERRAIN>(p); }
extern "C" __global__ void __launch_bounds__(CTA, NMF_MINBLOCKS_F64) nmf_step_tether_f64_kernel(const StepParamsT<double> p) { f64::step_entry<f64::W_TETHER>(p); }
extern "C" __global__ void __launch_bounds__(CTA, NMF_MINBLOCKS_F64) nmf_step_mesh_f64_kernel(const StepParamsT<double> p) { f64::step_entry<f64::W_MESH>(p); }
// with the noslip post-solver of the reference's CPU path (mujoco_globals.yaml:15; selected by the blob's noslip_iterations > 0):
// the `Simulation` (MuJoCo, float64) semantics, where `GPUSimulation` strips noslip (warp/simulation.py:427-448)
extern "C" __global__ void __launch_bounds__(CTA, NMF_MINBLOCKS_F64) nmf_step_noslip_f64_kernel(const StepParamsT<double> p) { f64::step_entry<f64::W_FLAT, 1, true>(p); }
extern "C" __global__ void __launch_bounds__(CTA, NMF_MINBLOCKS_F64) nmf_step_mesh_noslip_f64_kernel(const StepParamsT<double> p) { f64::step_entry<f64::W_MESH, 1, true>(p); }
extern "C" __global__ void __launch_bounds__(CTA, NMF_MINBLOCKS_F64) nmf_step_terrain_noslip_f64_kernel(const StepParamsT<double> p) { f64::step_entry<f64::W_TERRAIN, 1, true>(p); }
extern "C" __global__ void __launch_bounds__(CTA, NMF_MINBLOCKS_F64) nmf_step_tether_noslip_f64_kernel(const StepParamsT<double> p) { f64::step_entry<f64::W_TETHER, 1, true>(p); }
extern "C" __global__ void __launch_bounds__(CTA, 8) nmf_step_noslip_kernel(const StepParams p) { f32::step_entry<f32::W_FLAT, 1, true>(p); }
extern "C" __global__ void __launch_bounds__(CTA, 8) nmf_step_mesh_noslip_kernel(const StepParams p) { f32::step_entry<f32::W_MESH, 1, true>(p); }
extern "C" __global__ void __launch_bounds__(CTA, 8) nmf_step_terrain_noslip_kernel(const StepParams p) { f32::step_entry<f32::W_TERRAIN, 1, true>(p); }
extern "C" __global__ void __launch_bounds__(CTA, 8) nmf_step_tether_noslip_kernel(const StepParams p) { f32::step_entry<f32::W_TETHER, 1, true>(p); }

// general-topology models (JointPreset.ALL_BIOLOGICAL / ALL_POSSIBLE, ContactBodiesPreset.ALL, ...): nmf_tree.cuh, one block of
// 128 threads per fly, the whole fly in (dynamic) shared memory
#ifndef NMF_TREE_MINBLOCKS
#define NMF_TREE_MINBLOCKS 5     // <= 96 registers: the 45 KB of shared memory of the ALL_BIOLOGICAL fly allow 5 blocks per SM, the registers must too
#endif
extern "C" __global__ void __launch_bounds__(TREE_CTA, NMF_TREE_MINBLOCKS) nmf_tree_step_kernel(const TreeParamsT<float> p) {
  extern __shared__ __align__(16) unsigned char tree_smem[];
  f32::tree_step_block(p, reinterpret_cast<float*>(tree_smem), (int)blockIdx.x);
}
extern "C" __global__ void __launch_bounds__(TREE_CTA) nmf_tree_step_f64_kernel(const TreeParamsT<double> p) {
  extern __shared__ __align__(16) unsigned char tree_smem[];
  f64::tree_step_block(p, reinterpret_cast<double*>(tree_smem), (int)blockIdx.x);
}

__global__ void nmf_reset_kernel(float* state, const float* key, const uint8_t* mask, int n, int stride) {
  int fly = blockIdx.x;
  if (fly >= n || (mask && !mask[fly])) return;
  for (int i = threadIdx.x; i < stride; i += blockDim.x) state[(size_t)fly * stride + i] = key[i];
}

__global__ void nmf_scatter_cols_kernel(float* state, int stride, int off, const float* src, const int32_t* cols, int ncols, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * ncols) return;
  int fly = i / ncols, k = i - fly * ncols;
  state[(size_t)fly * stride + off + cols[k]] = src[i];
}

__global__ void nmf_gather_cols_kernel(const float* state, int stride, int off, const int32_t* cols, int ncols, float* dst, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * ncols) return;
  int fly = i / ncols, k = i - fly * ncols;
  dst[i] = state[(size_t)fly * stride + off + (cols ? cols[k] : k)];
}

// ------------------------------------------------------------------ handle
struct nmf_handle {
  HostModel hm;
  // general-topology models: tables of the tree kernels instead of the role table (tree == true)
  bool tree = false;
  TreeModel tm;
  int *d_it = nullptr; float* d_rt = nullptr; double* d_rt64 = nullptr;
  // record layout of the model (star models: the constants of nmf_layout.h)
  int stride = S_STRIDE, off_qpos = S_QPOS, off_qvel = S_QVEL, off_warm = S_WARM, off_ctrl = S_CTRL, off_time = S_TIME;
  int nq = NQ, nv = NV, nu_pos = 0, nu_adh = 0, nseg = 0, nleg = NLEG;
  int n_flies = 0, device = 0;
  float *d_role = nullptr, *d_hull = nullptr, *d_seg = nullptr, *d_key = nullptr;
  int *d_nbr_adr = nullptr, *d_nbr = nullptr;
  double *d_role64 = nullptr, *d_hull64 = nullptr;   // tables of the f64 validation kernels (uploaded by nmf_set_precision)
  double* d_state64 = nullptr; float* d_shadow = nullptr;   // f64: full-precision records between launches + float image of the last launch
  int precision = 32;
  float *d_act = nullptr, *d_qpos = nullptr;   // staging for nmf_step_host
  int* d_queue = nullptr;                      // work queue: counters, per-fly progress words, ring of ready flies
  int sub_steps = -1;                          // steps per work item: -1 = chosen per launch, 0 = never use the queue
  bool taper = true;                           // heuristic schedule: shrinking items at the end of a launch (env NMF_QUEUE_TAPER=0: uniform)
  int fpb64 = 4;                               // f64 flat kernel: flies per block (1, 2, 4; env NMF_FPB64).  B200, 4096 flies: 6.8 / 8.3 / 9.2 M env-steps/s
  int fpb = 0;                                 // fly slots per block of the f32 flat / terrain kernels: 1, 2, 4, 8 or 0 = chosen per launch (see step_block)
  int resident[9] = {};                        // resident blocks of the model's f32 kernel per fpb (index = fpb)
  int sms = 0;
  static constexpr int MAX_PARTS = 16;         // nmf_step_host: slices of the batch pipelined over private streams
  int host_parts = 4;                          // issued call by call; measured on B200, 4096 flies: 13.4 / 14.4 / 14.6 M env-steps/s end to end with 1 / 2 / 4 slices
  int graph_parts = 4;                         // replayed as a CUDA graph (pinned host buffers).  B200, 4096 flies, us per call (tools/e2e_sweep.py): call by call 278 / 253 with 1 / 4 slices; graph 262 / 248 / 253 / 256 / 256 with 2 / 4 / 6 / 8 / 16
  bool host_graph = true;                      // env NMF_HOST_GRAPH=0 switches the graph path off
  cudaStream_t part_stream[MAX_PARTS] = {}, cap_stream = nullptr;
  cudaEvent_t part_done[MAX_PARTS] = {}, fork = nullptr;
  // the captured pipeline of nmf_step_host (see there); rebuilt when anything that enters the kernels' parameters changes
  struct HostGraph {
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    cudaGraphNode_t h2d[MAX_PARTS] = {}, d2h[MAX_PARTS] = {};
    int parts = 0, cols = 0, nsteps = 0, launches = 0; uint64_t epoch = 0;
    const float* act = nullptr; float* qpos = nullptr;
  } hg;
  uint64_t epoch = 1;                          // bumped by nmf_bind and every setter
  int graph_failures = 0;
  nmf_buffers buf{};
  bool bound = false;
  int64_t launches = 0;
  std::string err;
};


// every entry point that touches the device runs on the handle's device and leaves the caller's current device as it found it
struct DeviceGuard {
  int prev = -1; bool switched = false;
  explicit DeviceGuard(int dev) { if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess; }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return NMF_ECUDA; } } while (0)

static int upload(nmf_handle* h, float** dst, const std::vector<float>& v) {
  CK(cudaMalloc(dst, sizeof(float) * (v.size() ? v.size() : 1)));
  CK(cudaMemcpy(*dst, v.data(), sizeof(float) * v.size(), cudaMemcpyHostToDevice));
  return NMF_OK;
}

static int set_resident_blocks(nmf_handle* h);

extern "C" int nmf_create(const void* blob, size_t nbytes, int n_flies, int device, nmf_handle** out) {
  if (!out) return NMF_EINVAL;
  *out = nullptr;
  nmf_handle* h = new (std::nothrow) nmf_handle;
  if (!h) return NMF_EINVAL;
  *out = h;   // returned even on failure so that nmf_last_error() can be read
  if (n_flies <= 0) { h->err = "n_flies must be positive"; return NMF_EINVAL; }
  const char* force_tree = getenv("NMF_FORCE_TREE");      // A/B runs: step a star-topology model with the general kernels
  if ((force_tree && atoi(force_tree) != 0) || !h->hm.build(blob, nbytes)) {
    // not the hub + 6 x 8 star the fast kernels are specialised to: any free root + hinge tree goes to the tree kernels
    if (!h->tm.build(blob, nbytes)) { h->err = "star kernels: " + (h->hm.err.empty() ? std::string("not tried") : h->hm.err) + "; tree kernels: " + h->tm.err; return NMF_EINVAL; }
    h->tree = true;
    const TreeDims& d = h->tm.par.d;
    h->stride = d.s_stride; h->off_qpos = d.s_qpos; h->off_qvel = d.s_qvel; h->off_warm = d.s_warm; h->off_ctrl = d.s_ctrl; h->off_time = d.s_time;
    h->nq = d.nq; h->nv = d.nv; h->nu_pos = d.nu_pos; h->nu_adh = d.nu_adh; h->nseg = d.nseg; h->nleg = d.nleg;
  } else {
    h->nu_pos = h->hm.par.nu_pos; h->nu_adh = h->hm.par.nu_adh; h->nseg = h->hm.nseg;
  }
  h->n_flies = n_flies; h->device = device;
  {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { h->err = "no such CUDA device"; return NMF_EINVAL; }
  }
  DeviceGuard guard(device);
  int rc;
  const std::vector<int32_t>& nbr_adr = h->tree ? h->tm.hull_nbr_adr : h->hm.hull_nbr_adr;
  const std::vector<int32_t>& nbr = h->tree ? h->tm.hull_nbr : h->hm.hull_nbr;
  if (h->tree) {
    if ((rc = upload(h, &h->d_rt, h->tm.rtab))) return rc;
    CK(cudaMalloc(&h->d_it, sizeof(int) * h->tm.itab.size()));
    CK(cudaMemcpy(h->d_it, h->tm.itab.data(), sizeof(int) * h->tm.itab.size(), cudaMemcpyHostToDevice));
    const size_t smem = (size_t)h->tm.par.d.m_total * sizeof(float);
    int smem_max = 0;
    CK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    if (smem > (size_t)smem_max) { h->err = "model too large for the tree kernels' shared-memory plan"; return NMF_EINVAL; }
    CK(cudaFuncSetAttribute(nmf_tree_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  } else {
    if ((rc = upload(h, &h->d_role, h->hm.role))) return rc;
  }
  if ((rc = upload(h, &h->d_hull, h->tree ? h->tm.hull : h->hm.hull))) return rc;
  if ((rc = upload(h, &h->d_seg, h->tree ? h->tm.seg_tab : h->hm.seg_tab))) return rc;
  if ((rc = upload(h, &h->d_key, h->tree ? h->tm.key_state : h->hm.key_state))) return rc;
  CK(cudaMalloc(&h->d_nbr_adr, sizeof(int) * nbr_adr.size()));
  CK(cudaMemcpy(h->d_nbr_adr, nbr_adr.data(), sizeof(int) * nbr_adr.size(), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&h->d_nbr, sizeof(int) * nbr.size()));
  CK(cudaMemcpy(h->d_nbr, nbr.data(), sizeof(int) * nbr.size(), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&h->d_queue, sizeof(int) * ((size_t)n_flies * QUEUE_MAX_CHUNKS + 2)));
  // staging of nmf_step_host (actions in, packed qpos out) and its pipeline streams: created here, so that no call on the
  // stepping path allocates
  CK(cudaMalloc(&h->d_act, sizeof(float) * (size_t)n_flies * (h->nu_pos + h->nu_adh > MAXU ? h->nu_pos + h->nu_adh : MAXU)));
  CK(cudaMalloc(&h->d_qpos, sizeof(float) * (size_t)n_flies * h->nq));
  for (int k = 0; k < nmf_handle::MAX_PARTS; k++) {
    CK(cudaStreamCreateWithFlags(&h->part_stream[k], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->part_done[k], cudaEventDisableTiming));
  }
  CK(cudaEventCreateWithFlags(&h->fork, cudaEventDisableTiming));
  CK(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
  if (const char* e = getenv("NMF_FPB64")) { int v = atoi(e); if (v == 1 || v == 2 || v == 4) h->fpb64 = v; }
  if (const char* e = getenv("NMF_FPB")) { int v = atoi(e); if (v == 0 || v == 1 || v == 2 || v == 4 || v == 8) h->fpb = v; }
  if (h->hm.par.weld) h->fpb = 1;
  if (!h->tree) {
    int rc2 = set_resident_blocks(h);
    if (rc2) return rc2;
  }
  if (const char* e = getenv("NMF_QUEUE_SUBSTEPS")) h->sub_steps = atoi(e);
  if (const char* e = getenv("NMF_QUEUE_TAPER")) h->taper = atoi(e) != 0;
  if (const char* e = getenv("NMF_HOST_PARTS")) { int v = atoi(e); if (v >= 1 && v <= nmf_handle::MAX_PARTS) h->host_parts = h->graph_parts = v; }
  if (const char* e = getenv("NMF_HOST_GRAPH")) h->host_graph = atoi(e) != 0;
  return NMF_OK;
}

extern "C" int nmf_destroy(nmf_handle* h) {
  if (!h) return NMF_OK;
  DeviceGuard guard(h->device);
  cudaFree(h->d_it); cudaFree(h->d_rt); cudaFree(h->d_rt64); cudaFree(h->d_role); cudaFree(h->d_hull); cudaFree(h->d_seg); cudaFree(h->d_key); cudaFree(h->d_nbr_adr); cudaFree(h->d_nbr); cudaFree(h->d_act); cudaFree(h->d_qpos); cudaFree(h->d_queue); cudaFree(h->d_role64); cudaFree(h->d_hull64); cudaFree(h->d_state64); cudaFree(h->d_shadow);
  for (int k = 0; k < nmf_handle::MAX_PARTS; k++) { if (h->part_stream[k]) cudaStreamDestroy(h->part_stream[k]); if (h->part_done[k]) cudaEventDestroy(h->part_done[k]); }
  if (h->fork) cudaEventDestroy(h->fork);
  if (h->hg.exec) cudaGraphExecDestroy(h->hg.exec);
  if (h->hg.graph) cudaGraphDestroy(h->hg.graph);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  delete h;
  return NMF_OK;
}

extern "C" const char* nmf_last_error(const nmf_handle* h) { return h ? h->err.c_str() : "null handle"; }
extern "C" int64_t nmf_launch_count(const nmf_handle* h) { return h ? h->launches : 0; }

extern "C" int nmf_model_info(const nmf_handle* h, nmf_info* info) {
  if (!h || !info) return NMF_EINVAL;
  info->n_flies = h->n_flies; info->nq = h->nq; info->nv = h->nv; info->nu_pos = h->nu_pos; info->nu_adh = h->nu_adh;
  info->nseg = h->nseg; info->nleg = h->nleg; info->state_stride = h->stride; info->off_qpos = h->off_qpos; info->off_qvel = h->off_qvel;
  info->off_qacc_warmstart = h->off_warm; info->off_ctrl = h->off_ctrl; info->off_time = h->off_time; info->dbg_stride = h->tree ? TDBG_STRIDE : DBG_STRIDE;
  info->off_status = h->off_time + 1;
  info->timestep = h->tree ? h->tm.par.dt : h->hm.par.dt;
  return NMF_OK;
}

extern "C" int nmf_bind(nmf_handle* h, const nmf_buffers* b) {
  if (!h || !b || !b->state) { if (h) h->err = "nmf_bind: state buffer is required"; return NMF_EINVAL; }
  h->buf = *b; h->bound = true; h->epoch++;
  return NMF_OK;
}

extern "C" int nmf_set_solver(nmf_handle* h, int max_newton, int max_ls) {
  if (!h || max_newton < 1 || max_ls < 1) return NMF_EINVAL;
  h->hm.par.max_newton = max_newton; h->hm.par.max_ls = max_ls; h->epoch++;
  h->tm.par.max_newton = max_newton; h->tm.par.max_ls = max_ls;
  return NMF_OK;
}

extern "C" int nmf_reset(nmf_handle* h, const uint8_t* mask, void* stream) {
  if (!h) return NMF_EINVAL;
  if (!h->bound) { h->err = "nmf_reset: not bound"; return NMF_ENOTBOUND; }
  DeviceGuard guard(h->device);
  nmf_reset_kernel<<<h->n_flies, 64, 0, (cudaStream_t)stream>>>(h->buf.state, h->d_key, mask, h->n_flies, h->stride);
  h->launches++;
  CK(cudaGetLastError());
  return NMF_OK;
}

extern "C" int nmf_set_schedule(nmf_handle* h, int sub_steps) {
  if (!h || sub_steps < -1) return NMF_EINVAL;
  h->sub_steps = sub_steps; h->epoch++;
  return NMF_OK;
}

extern "C" int nmf_set_flies_per_block(nmf_handle* h, int fpb) {
  if (!h || (fpb != 0 && fpb != 1 && fpb != 2 && fpb != 4 && fpb != 8)) return NMF_EINVAL;
  h->epoch++;
  if (h->tree) return NMF_OK;                       // one fly per block in the tree kernels
  h->fpb = (h->hm.par.weld || h->hm.par.noslip_iterations > 0) ? 1 : fpb;
  return NMF_OK;
}

static int launch_steps(nmf_handle* h, int nsteps, const float* table, int table_T, int table_t0, int table_cols, bool forward_only, void* stream,
                        int fly0 = 0, int count = -1, float* out_qpos = nullptr);

extern "C" int nmf_step(nmf_handle* h, int nsteps, const float* table, int table_T, int table_t0, int table_cols, void* stream) {
  return launch_steps(h, nsteps, table, table_T, table_t0, table_cols, false, stream);
}

extern "C" int nmf_forward(nmf_handle* h, void* stream) { return launch_steps(h, 1, nullptr, 0, 0, 0, true, stream); }

template <class real> struct KernelSet;
typedef void (*step_kernel_f32)(const StepParams);
static step_kernel_f32 kernel_f32(const StepParams& q, int fpb) {
  const bool weld = q.weld, terrain = q.terrain;
  if (weld) return q.noslip_iterations > 0 ? nmf_step_tether_noslip_kernel : nmf_step_tether_kernel;
  if (q.noslip_iterations > 0) return q.multiccd ? nmf_step_mesh_noslip_kernel : terrain ? nmf_step_terrain_noslip_kernel : nmf_step_noslip_kernel;
  if (q.multiccd) return fpb == 8 ? nmf_step_mesh_x8_kernel : fpb == 4 ? nmf_step_mesh_x4_kernel : fpb == 2 ? nmf_step_mesh_x2_kernel : nmf_step_mesh_kernel;
  if (terrain) return fpb == 8 ? nmf_step_terrain_x8_kernel : fpb == 4 ? nmf_step_terrain_x4_kernel : fpb == 2 ? nmf_step_terrain_x2_kernel : nmf_step_terrain_kernel;
  return fpb == 8 ? nmf_step_x8_kernel : fpb == 4 ? nmf_step_x4_kernel : fpb == 2 ? nmf_step_x2_kernel : nmf_step_kernel;
}
// blocks of 8 fly slots exceed the 48 KB of static shared memory: their per-fly regions live in (opt-in) dynamic shared memory
// the noslip instantiations carry B_tt (96 x 96) behind the fly's regular region: beyond 48 KB as well (step_entry takes the same decision)
template <class NSreal> static size_t noslip_smem(int sm_total, int ns_count) { const size_t b = (size_t)(sm_total + ns_count) * sizeof(NSreal); return b > 48 * 1024 ? b : 0; }
static size_t dyn_smem_f32(const StepParams& q, int fpb) {
  if (q.noslip_iterations > 0) return noslip_smem<float>(f32::SM_TOTAL, f32::NS_COUNT);
  return fpb >= 8 ? (size_t)fpb * f32::SM_WELD * sizeof(float) : 0;
}
static int set_resident_blocks(nmf_handle* h) {
  DeviceGuard guard(h->device);
  CK(cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, h->device));
  for (int fpb = 1; fpb <= 8; fpb *= 2) {
    if ((h->hm.par.weld || h->hm.par.noslip_iterations > 0) && fpb > 1) break;
    int per_sm = 0;
    const size_t dyn = dyn_smem_f32(h->hm.par, fpb);
    if (dyn) CK(cudaFuncSetAttribute(kernel_f32(h->hm.par, fpb), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel_f32(h->hm.par, fpb), CTA * fpb, dyn));
    h->resident[fpb] = per_sm * h->sms;
  }
  return NMF_OK;
}
// fly slots per block for a launch over n flies: the handle's setting, or the largest block that still gives every SM one
// (sharing instruction fetches among the slots of a block is worth more than spreading a small batch thinly)
static int pick_fpb(const nmf_handle* h, int n) {
  if (h->hm.par.weld || h->hm.par.noslip_iterations > 0) return 1;
  if (h->fpb) return h->fpb;
  if (n >= 8 * h->sms) return 8;
  if (n >= 4 * h->sms) return 4;       // (2 flies per block share nothing in the L0 instruction caches and still wait: slower than 1)
  return 1;
}
template <> struct KernelSet<float> {
  static const StepParamsT<float>& base(const nmf_handle* h) { return h->hm.par; }
  static const float* role(const nmf_handle* h) { return h->d_role; }
  static const float* hull(const nmf_handle* h) { return h->d_hull; }
  static int fpb(const nmf_handle* h, int n) { return pick_fpb(h, n); }
  static void launch(const StepParamsT<float>& p, int fpb, int grid, cudaStream_t s) {
    kernel_f32(p, fpb)<<<grid, CTA * fpb, dyn_smem_f32(p, fpb), s>>>(p);
  }
};
template <> struct KernelSet<double> {
  static const StepParamsT<double>& base(const nmf_handle* h) { return h->hm.par64; }
  static const double* role(const nmf_handle* h) { return h->d_role64; }
  static const double* hull(const nmf_handle* h) { return h->d_hull64; }
  static int fpb(const nmf_handle* h, int n) {
    const int want = h->fpb64;
    const bool plain_flat = !h->hm.par.weld && !h->hm.par.terrain && !h->hm.par.multiccd && h->hm.par.noslip_iterations == 0;
    return (plain_flat && (want == 2 || want == 4) && n >= want) ? want : 1;
  }
  static void launch(const StepParamsT<double>& p, int fpb, int grid, cudaStream_t s) {
    if (fpb == 4) {
      const size_t dyn = (size_t)4 * f64::SM_WELD * sizeof(double);
      cudaFuncSetAttribute(nmf_step_f64_x4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);      // (a per-device attribute: a process-wide "done" flag would miss the second GPU)
      nmf_step_f64_x4_kernel<<<grid, 4 * CTA, dyn, s>>>(p); return;
    }
    if (fpb == 2) { nmf_step_f64_x2_kernel<<<grid, 2 * CTA, 0, s>>>(p); return; }
    if (p.noslip_iterations > 0) {
      typedef void (*k64)(const StepParamsT<double>);
      const k64 kern = p.weld ? nmf_step_tether_noslip_f64_kernel : p.multiccd ? nmf_step_mesh_noslip_f64_kernel
                     : p.terrain ? nmf_step_terrain_noslip_f64_kernel : nmf_step_noslip_f64_kernel;
      const size_t dyn = noslip_smem<double>(f64::SM_TOTAL, f64::NS_COUNT);
      if (dyn) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);     // (per device: set on every launch of this validation path)
      kern<<<grid, CTA, dyn, s>>>(p);
    }
    else if (p.weld) nmf_step_tether_f64_kernel<<<grid, CTA, 0, s>>>(p);
    else if (p.multiccd) nmf_step_mesh_f64_kernel<<<grid, CTA, 0, s>>>(p);
    else if (p.terrain) nmf_step_terrain_f64_kernel<<<grid, CTA, 0, s>>>(p);
    else nmf_step_f64_kernel<<<grid, CTA, 0, s>>>(p);
  }
};

// flies [fly0, fly0 + count) only (count < 0: all): the buffers are addressed per fly, so a range is the same launch on offset pointers
template <class real>
static int launch_steps_t(nmf_handle* h, int nsteps, const float* table, int table_T, int table_t0, int table_cols, bool forward_only, void* stream,
                          int fly0, int count, float* out_qpos) {
  StepParamsT<real> p = KernelSet<real>::base(h);
  p.max_newton = h->hm.par.max_newton; p.max_ls = h->hm.par.max_ls;       // nmf_set_solver edits the f32 copy
  p.state = h->buf.state; p.state64 = nullptr; p.shadow = nullptr;
  if (std::is_same<real, double>::value) { p.state64 = h->d_state64; p.shadow = h->d_shadow; }
  p.role = KernelSet<real>::role(h); p.hull = KernelSet<real>::hull(h); p.seg_tab = h->d_seg; p.hull_nbr_adr = h->d_nbr_adr; p.hull_nbr = h->d_nbr;
  p.act_table = table; p.table_T = table_T; p.table_t0 = table_t0; p.table_cols = table ? table_cols : 0;
  p.out_xpos = h->buf.seg_xpos; p.out_xquat = h->buf.seg_xquat; p.out_actf = h->buf.act_force; p.out_sensor = h->buf.sensordata;
  p.out_energy = h->buf.energy; p.out_qpos = out_qpos;      // (out_qpos already points at the first fly of a ranged launch)
  p.dbg = h->buf.debug; p.n_flies = h->n_flies; p.nsteps = nsteps; p.forward_only = forward_only ? 1 : 0;
  const bool ranged = count >= 0 && (fly0 != 0 || count != h->n_flies);
  if (ranged) {
    const size_t f = (size_t)fly0, nu = (size_t)(p.nu_pos + p.nu_adh);
    p.state += f * S_STRIDE; p.n_flies = count;
    if (p.state64) { p.state64 += f * S_STRIDE; p.shadow += f * S_STRIDE; }
    if (p.act_table) p.act_table += f * (size_t)table_T * table_cols;
    if (p.out_xpos) p.out_xpos += f * p.nseg * 3;
    if (p.out_xquat) p.out_xquat += f * p.nseg * 4;
    if (p.out_actf) p.out_actf += f * nu;
    if (p.out_sensor) p.out_sensor += f * NLEG * 16;
    if (p.out_energy) p.out_energy += f * 2;
    if (p.dbg) p.dbg += f * DBG_STRIDE;
  }
  const int fpb = KernelSet<real>::fpb(h, p.n_flies);
  const int n_units = (p.n_flies + fpb - 1) / fpb;      // a block steps a unit of fpb consecutive flies
  int grid = n_units;
  p.queue = nullptr; p.sub_steps = nsteps; p.n_items = n_units; p.n_chunks = 0;
  int sub = h->sub_steps;
  if (sub < 0) {
    // ~8 steps per item.  Measured on B200 with 8 flies per block (profiles/queue_sweep_r02.txt, 4096 flies): a 20-step launch
    // runs at 20.9 / 24.0 / 23.9 / 22.2 M env-steps/s with no queue / items of 3 / 7 / 10 steps, a 100-step launch at
    // 21.7 / 25.1 / 24.6 / 22.7 M with no queue / 10 / 25 / 50: a block's fixed cost per item (record load, set-up code) is
    // small next to the empty slots that long items leave in the last wave.
    const int k = (nsteps + 7) / 8;
    sub = k >= 2 ? (nsteps + k - 1) / k : 0;
  }
  // more units than resident blocks: work queue (one per handle; sized for the f32 kernels' occupancy, so f32 only)
  if (std::is_same<real, float>::value && !ranged && sub > 0 && n_units > h->resident[fpb] && nsteps >= 2 * sub) {
    if (sub * QUEUE_MAX_CHUNKS < nsteps) sub = (nsteps + QUEUE_MAX_CHUNKS - 1) / QUEUE_MAX_CHUNKS;
    int nchunk = (nsteps + sub - 1) / sub;
    p.n_chunks = 0;
    if (h->sub_steps < 0 && h->taper && nsteps <= 30000) {      // (boundaries are stored as shorts)
      // Tapered schedule (the heuristic only; an explicit nmf_set_schedule stays uniform): the launch ends when the last block has
      // finished its last item, and while that item runs the blocks that found the queue empty idle -- up to one item of ~8 steps,
      // 0.7 ms of a 3.6 ms launch of 20 steps.  So the last items of every unit shrink (..., 6, 4, 3, 2 steps): the bulk keeps its
      // low per-item cost and the tail of the launch is a 2-step item.
      int taper[8], nt = 0, tsum = 0;
      for (int t = 2; t < sub && nt < 8 && tsum + t <= nsteps - sub; t = t + 1 > t * 3 / 2 ? t + 1 : t * 3 / 2) { taper[nt++] = t; tsum += t; }
      const int bulk = nsteps - tsum;
      int kb = (bulk + sub - 1) / sub;
      if (kb + nt <= QUEUE_MAX_CHUNKS && nt > 0) {
        int at = 0, c = 0;
        for (int i = 0; i < kb; i++) { p.chunk_start[c++] = (short)at; at += bulk / kb + (i < bulk % kb ? 1 : 0); }
        for (int i = nt - 1; i >= 0; i--) { p.chunk_start[c++] = (short)at; at += taper[i]; }
        p.chunk_start[c] = (short)at;                 // == nsteps
        p.n_chunks = nchunk = c;
      }
    }
    p.queue = h->d_queue; p.sub_steps = sub; p.n_items = nchunk * n_units;
    grid = h->resident[fpb];
    CK(cudaMemsetAsync(h->d_queue, 0, sizeof(int) * ((size_t)n_units * nchunk + 2), (cudaStream_t)stream));
  }
  KernelSet<real>::launch(p, fpb, grid, (cudaStream_t)stream);
  h->launches++;
  CK(cudaGetLastError());
  return NMF_OK;
}

// general-topology models: one block per fly, grid = flies of the (ranged) launch
template <class real>
static int launch_tree_t(nmf_handle* h, int nsteps, const float* table, int table_T, int table_t0, int table_cols, bool forward_only, void* stream,
                         int fly0, int count, float* out_qpos) {
  TreeParamsT<real> p;
  if constexpr (std::is_same<real, double>::value) { p = h->tm.par64; p.rt = h->d_rt64; p.hull = h->d_hull64; p.state64 = h->d_state64; p.shadow = h->d_shadow; }
  else { p = h->tm.par; p.rt = h->d_rt; p.hull = h->d_hull; p.state64 = nullptr; p.shadow = nullptr; }
  p.max_newton = h->tm.par.max_newton; p.max_ls = h->tm.par.max_ls;
  p.state = h->buf.state; p.it = h->d_it; p.hull_nbr_adr = h->d_nbr_adr; p.hull_nbr = h->d_nbr; p.seg_tab = h->d_seg;
  p.act_table = table; p.table_T = table_T; p.table_t0 = table_t0; p.table_cols = table ? table_cols : 0;
  p.out_xpos = h->buf.seg_xpos; p.out_xquat = h->buf.seg_xquat; p.out_actf = h->buf.act_force; p.out_sensor = h->buf.sensordata;
  p.out_energy = h->buf.energy; p.out_qpos = out_qpos; p.dbg = h->buf.debug;
  p.n_flies = h->n_flies; p.nsteps = nsteps; p.forward_only = forward_only ? 1 : 0;
  if (count >= 0 && (fly0 != 0 || count != h->n_flies)) {
    const size_t f = (size_t)fly0, nu = (size_t)(h->nu_pos + h->nu_adh);
    p.state += f * h->stride; p.n_flies = count;
    if (p.state64) { p.state64 += f * h->stride; p.shadow += f * h->stride; }
    if (p.act_table) p.act_table += f * (size_t)table_T * table_cols;
    if (p.out_xpos) p.out_xpos += f * h->nseg * 3;
    if (p.out_xquat) p.out_xquat += f * h->nseg * 4;
    if (p.out_actf) p.out_actf += f * nu;
    if (p.out_sensor) p.out_sensor += f * h->nleg * 16;
    if (p.out_energy) p.out_energy += f * 2;
    if (p.dbg) p.dbg += f * TDBG_STRIDE;
  }
  const size_t smem = (size_t)p.d.m_total * sizeof(real);
  if constexpr (std::is_same<real, double>::value) nmf_tree_step_f64_kernel<<<p.n_flies, TREE_CTA, smem, (cudaStream_t)stream>>>(p);
  else nmf_tree_step_kernel<<<p.n_flies, TREE_CTA, smem, (cudaStream_t)stream>>>(p);
  h->launches++;
  CK(cudaGetLastError());
  return NMF_OK;
}

static int launch_steps(nmf_handle* h, int nsteps, const float* table, int table_T, int table_t0, int table_cols, bool forward_only, void* stream,
                        int fly0, int count, float* out_qpos) {
  if (!h) return NMF_EINVAL;
  if (!h->bound) { h->err = "nmf_step: not bound"; return NMF_ENOTBOUND; }
  if (nsteps <= 0) return NMF_OK;
  if (table && table_T <= 0) { h->err = "nmf_step: action table needs table_T > 0"; return NMF_EINVAL; }
  if (table) table_t0 = ((table_t0 % table_T) + table_T) % table_T;    // any integer start row, negative ones included
  DeviceGuard guard(h->device);
  if (table && table_cols != h->nu_pos && table_cols != h->nu_pos + h->nu_adh) {
    h->err = "nmf_step: action table rows must hold nu_pos (position targets) or nu_pos + nu_adh (+ adhesion) controls"; return NMF_EINVAL;
  }
  if (h->tree)
    return h->precision == 64 ? launch_tree_t<double>(h, nsteps, table, table_T, table_t0, table_cols, forward_only, stream, fly0, count, out_qpos)
                              : launch_tree_t<float>(h, nsteps, table, table_T, table_t0, table_cols, forward_only, stream, fly0, count, out_qpos);
  return h->precision == 64 ? launch_steps_t<double>(h, nsteps, table, table_T, table_t0, table_cols, forward_only, stream, fly0, count, out_qpos)
                            : launch_steps_t<float>(h, nsteps, table, table_T, table_t0, table_cols, forward_only, stream, fly0, count, out_qpos);
}

// Arithmetic of the step kernels: 32 (default, the product path) or 64 = the SAME kernel source instantiated in double precision
// (validation: shadows the fp64 oracle over long horizons; the state records in HBM stay float32, so carry precision across
// steps by fusing them into one launch).
extern "C" int nmf_set_precision(nmf_handle* h, int bits) {
  if (!h || (bits != 32 && bits != 64)) return NMF_EINVAL;
  DeviceGuard guard(h->device);
  if (bits == 64 && !h->d_state64) {
    if (h->tree) {
      const size_t smem = (size_t)h->tm.par64.d.m_total * sizeof(double);
      int smem_max = 0;
      CK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
      if (smem > (size_t)smem_max) { h->err = "model too large for the f64 tree kernel's shared-memory plan"; return NMF_EINVAL; }
      CK(cudaFuncSetAttribute(nmf_tree_step_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CK(cudaMalloc(&h->d_rt64, sizeof(double) * h->tm.rtab64.size()));
      CK(cudaMemcpy(h->d_rt64, h->tm.rtab64.data(), sizeof(double) * h->tm.rtab64.size(), cudaMemcpyHostToDevice));
    } else {
      CK(cudaMalloc(&h->d_role64, sizeof(double) * h->hm.role64.size()));
      CK(cudaMemcpy(h->d_role64, h->hm.role64.data(), sizeof(double) * h->hm.role64.size(), cudaMemcpyHostToDevice));
    }
    const std::vector<double>& hull64 = h->tree ? h->tm.hull64 : h->hm.hull64;
    CK(cudaMalloc(&h->d_hull64, sizeof(double) * hull64.size()));
    CK(cudaMemcpy(h->d_hull64, hull64.data(), sizeof(double) * hull64.size(), cudaMemcpyHostToDevice));
    // full-precision records: start empty; a shadow of NaNs never equals a float record, so the first launch reads the float state
    const size_t nrec = (size_t)h->n_flies * h->stride;
    CK(cudaMalloc(&h->d_state64, sizeof(double) * nrec));
    CK(cudaMalloc(&h->d_shadow, sizeof(float) * nrec));
    CK(cudaMemset(h->d_state64, 0, sizeof(double) * nrec));
    CK(cudaMemset(h->d_shadow, 0xff, sizeof(float) * nrec));
  }
  h->precision = bits; h->epoch++;
  return NMF_OK;
}

extern "C" int nmf_scatter_ctrl(nmf_handle* h, const float* src, const int32_t* cols, int ncols, void* stream) {
  if (!h || !src || !cols || ncols <= 0) return NMF_EINVAL;
  if (!h->bound) return NMF_ENOTBOUND;
  DeviceGuard guard(h->device);
  int total = h->n_flies * ncols;
  nmf_scatter_cols_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->buf.state, h->stride, h->off_ctrl, src, cols, ncols, h->n_flies);
  h->launches++;
  CK(cudaGetLastError());
  return NMF_OK;
}

extern "C" int nmf_gather_state(nmf_handle* h, int off, const int32_t* cols, int ncols, float* dst, void* stream) {
  if (!h || !dst || ncols <= 0 || off < 0 || off >= h->stride) return NMF_EINVAL;
  if (!h->bound) return NMF_ENOTBOUND;
  DeviceGuard guard(h->device);
  int total = h->n_flies * ncols;
  nmf_gather_cols_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->buf.state, h->stride, off, cols, ncols, dst, h->n_flies);
  h->launches++;
  CK(cudaGetLastError());
  return NMF_OK;
}

// Host-buffer step.  With a few thousand flies the batch is cut into slices that run on the handle's own streams, forked from and
// joined to one stream with events: slice k's H2D copy and D2H read-back overlap the other slices' kernels (the kernels of all
// slices are co-resident, so the device sees the same blocks as one launch would give it).
static int issue_host_slices(nmf_handle* h, const float* actions_host, int action_cols, int nsteps, float* qpos_host, cudaStream_t root, int parts) {
  const int n = h->n_flies, NQ = h->nq;
  if (parts > 1) CK(cudaEventRecord(h->fork, root));
  for (int k = 0; k < parts; k++) {
    const int f0 = (int)((long long)n * k / parts), cnt = (int)((long long)n * (k + 1) / parts) - f0;
    cudaStream_t s = parts > 1 ? h->part_stream[k] : root;
    if (parts > 1) CK(cudaStreamWaitEvent(s, h->fork, 0));
    float* d_act = h->d_act + (size_t)f0 * action_cols;
    CK(cudaMemcpyAsync(d_act, actions_host + (size_t)f0 * action_cols, sizeof(float) * (size_t)cnt * action_cols, cudaMemcpyHostToDevice, s));
    // the action block doubles as a 1-row action table: ctrl[0:action_cols] <- actions (position actuators first, then adhesion)
    // the step kernel packs qpos of its flies into d_qpos when it writes the records back (no separate gather launch)
    int rc = launch_steps(h, nsteps, h->d_act, 1, 0, action_cols, false, s, f0, cnt, h->d_qpos + (size_t)f0 * NQ);
    if (rc) return rc;
    CK(cudaMemcpyAsync(qpos_host + (size_t)f0 * NQ, h->d_qpos + (size_t)f0 * NQ, sizeof(float) * (size_t)cnt * NQ, cudaMemcpyDeviceToHost, s));
    if (parts > 1) { CK(cudaEventRecord(h->part_done[k], s)); CK(cudaStreamWaitEvent(root, h->part_done[k], 0)); }
  }
  return NMF_OK;
}

static bool pinned_host(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// The same pipeline as a CUDA graph: issuing 2 copies + 1 launch + 3 event calls per slice costs the host ~15 us per slice; captured
// once and replayed it costs one launch.  Measured on B200 this buys little end to end (248 us against 253 us per call at 4096
// flies, finer slices do not help: the call is bounded by the two waves of the 1-step kernels, 180-200 us, plus the exposed ends of
// the copies), but the host thread is free for ~45 us more per step.  The graph is keyed on everything
// that enters the kernels' parameters (handle epoch, nsteps, columns); only the host addresses may change from call to call -- they
// are patched into the memcpy nodes of the instantiated graph.  Needs pinned (or registered) host buffers; anything else, the f64
// build and a failed capture take the call-by-call path.
static int build_host_graph(nmf_handle* h, const float* actions_host, int action_cols, int nsteps, float* qpos_host, int parts) {
  nmf_handle::HostGraph& g = h->hg;
  if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
  if (g.graph) { cudaGraphDestroy(g.graph); g.graph = nullptr; }
  if (cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); return NMF_ECUDA; }
  const int64_t before = h->launches;
  int rc = issue_host_slices(h, actions_host, action_cols, nsteps, qpos_host, h->cap_stream, parts);
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
  const int launches = (int)(h->launches - before);
  h->launches = before;                                  // (nothing ran yet)
  if (rc != NMF_OK || e != cudaSuccess || !graph) { cudaGetLastError(); if (graph) cudaGraphDestroy(graph); return rc != NMF_OK ? rc : NMF_ECUDA; }
  g.graph = graph;
  if (cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) { cudaGetLastError(); g.exec = nullptr; return NMF_ECUDA; }
  // find the copy nodes of every slice by their device addresses
  size_t nn = 0;
  if (cudaGraphGetNodes(graph, nullptr, &nn) != cudaSuccess) { cudaGetLastError(); return NMF_ECUDA; }
  std::vector<cudaGraphNode_t> nodes(nn);
  if (cudaGraphGetNodes(graph, nodes.data(), &nn) != cudaSuccess) { cudaGetLastError(); return NMF_ECUDA; }
  const int n = h->n_flies, NQ = h->nq;
  int found = 0;
  for (int k = 0; k < parts; k++) { g.h2d[k] = nullptr; g.d2h[k] = nullptr; }
  for (size_t i = 0; i < nn; i++) {
    cudaGraphNodeType t;
    if (cudaGraphNodeGetType(nodes[i], &t) != cudaSuccess || t != cudaGraphNodeTypeMemcpy) continue;
    cudaMemcpy3DParms mp;
    if (cudaGraphMemcpyNodeGetParams(nodes[i], &mp) != cudaSuccess) continue;
    for (int k = 0; k < parts; k++) {
      const int f0 = (int)((long long)n * k / parts);
      if (mp.dstPtr.ptr == (void*)(h->d_act + (size_t)f0 * action_cols) && !g.h2d[k]) { g.h2d[k] = nodes[i]; found++; }
      if (mp.srcPtr.ptr == (void*)(h->d_qpos + (size_t)f0 * NQ) && !g.d2h[k]) { g.d2h[k] = nodes[i]; found++; }
    }
  }
  if (found != 2 * parts) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; return NMF_ECUDA; }
  g.parts = parts; g.cols = action_cols; g.nsteps = nsteps; g.launches = launches; g.epoch = h->epoch; g.act = actions_host; g.qpos = qpos_host;
  return NMF_OK;
}

extern "C" int nmf_step_host(nmf_handle* h, const float* actions_host, int action_cols, int nsteps, float* qpos_host, void* stream_) {
  if (!h || !actions_host || !qpos_host) return NMF_EINVAL;
  if (!h->bound) return NMF_ENOTBOUND;
  DeviceGuard guard(h->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int nu = h->nu_pos + h->nu_adh, n = h->n_flies, NQ = h->nq;
  if (action_cols != h->nu_pos && action_cols != nu) { h->err = "nmf_step_host: actions must have nu_pos or nu_pos + nu_adh columns"; return NMF_EINVAL; }
  if (!h->d_act) { h->err = "nmf_step_host: staging buffers missing"; return NMF_EINVAL; }     // allocated by nmf_create
  if (nsteps <= 0) return NMF_OK;
  // ---- graph path
  int gparts = h->graph_parts;
  while (gparts > 1 && n < 512 * gparts) gparts--;       // at least 512 flies per slice
  if (h->host_graph && h->precision == 32 && gparts > 1 && h->graph_failures < 3) {
    nmf_handle::HostGraph& g = h->hg;
    const bool same = g.exec && g.epoch == h->epoch && g.cols == action_cols && g.nsteps == nsteps && g.parts == gparts;
    bool ok = same;
    if (!same || actions_host != g.act || qpos_host != g.qpos) ok = pinned_host(actions_host) && pinned_host(qpos_host);
    if (ok && !same) {
      ok = build_host_graph(h, actions_host, action_cols, nsteps, qpos_host, gparts) == NMF_OK;
      if (!ok) h->graph_failures++;
    }
    if (ok) {
      for (int k = 0; k < g.parts && ok; k++) {
        const int f0 = (int)((long long)n * k / g.parts), cnt = (int)((long long)n * (k + 1) / g.parts) - f0;
        if (actions_host != g.act)
          ok = cudaGraphExecMemcpyNodeSetParams1D(g.exec, g.h2d[k], h->d_act + (size_t)f0 * action_cols, actions_host + (size_t)f0 * action_cols,
                                                  sizeof(float) * (size_t)cnt * action_cols, cudaMemcpyHostToDevice) == cudaSuccess;
        if (ok && qpos_host != g.qpos)
          ok = cudaGraphExecMemcpyNodeSetParams1D(g.exec, g.d2h[k], qpos_host + (size_t)f0 * NQ, h->d_qpos + (size_t)f0 * NQ,
                                                  sizeof(float) * (size_t)cnt * NQ, cudaMemcpyDeviceToHost) == cudaSuccess;
      }
      if (ok) {
        g.act = actions_host; g.qpos = qpos_host;
        CK(cudaGraphLaunch(g.exec, stream));
        h->launches += g.launches;
        CK(cudaStreamSynchronize(stream));
        return NMF_OK;
      }
      cudaGetLastError(); h->graph_failures++; g.epoch = 0;      // a node update was refused: rebuild next time, this call goes the plain way
    }
  }
  // ---- call by call
  int parts = h->host_parts;
  while (parts > 1 && n < 1024 * parts) parts--;   // at least 1024 flies per slice
  int rc = issue_host_slices(h, actions_host, action_cols, nsteps, qpos_host, stream, parts);
  if (rc) return rc;
  CK(cudaStreamSynchronize(stream));
  return NMF_OK;
}

// ------------------------------------------------------------------ action table of the reference benchmark, built on the device
// MotionSnippet.get_joint_angles (reference src/flygym_demo/spotlight_data/preprocessing.py:80-142) resamples the recorded clip
// onto the simulation time grid with a cubic spline, and ReplayTargetData.make_target_angles_all_worlds
// (src/flygym_demo/benchmark/time_gpu_simulation.py:73-86) tiles it: world k replays partition k % n_part.  The reference
// builds the (n_worlds, T, A) float32 table on the host and uploads it (688 MB at 4096 worlds); here only the spline's
// piecewise-cubic coefficients travel (n_int * 4 * A doubles) and every table entry is one Horner evaluation.
__global__ void nmf_replay_table_kernel(const double* __restrict__ coef, const double* __restrict__ last, int n_int, int A, double fps, double dt,
                                        int n_part, int T, long long total, int fly_offset, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int a = (int)(i % A);
  const long long ws = i / A;
  const int s = (int)(ws % T), world = (int)(ws / T);
  const int part = (world + fly_offset) % n_part;
  const double t = (double)((long long)part * T + s) * dt;
  const double x_last = (double)n_int / fps;                  // last source sample; beyond it interp1d returns fill_value = y[-1]
  double v;
  if (t > x_last) v = last[a];
  else {
    int k = (int)floor(t * fps); k = k < 0 ? 0 : (k >= n_int ? n_int - 1 : k);
    const double u = t - (double)k / fps;
    const double* c = coef + (size_t)k * A + a;               // coef[d][k][a], highest power first (scipy PPoly layout)
    const size_t stride = (size_t)n_int * A;
    v = ((c[0] * u + c[stride]) * u + c[2 * stride]) * u + c[3 * stride];
  }
  out[i] = (float)v;
}

extern "C" int nmf_replay_table(const double* coef_host, const double* last_host, int n_int, int A, double fps, double dt, int n_part, int T,
                                int n_worlds, int fly_offset, float* out_dev, void* stream_) {
  if (!coef_host || !last_host || !out_dev || n_int <= 0 || A <= 0 || n_part <= 0 || T <= 0 || n_worlds <= 0 || !(fps > 0) || !(dt > 0)) return NMF_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  double* d = nullptr;
  const size_t nc = (size_t)4 * n_int * A;
  if (cudaMalloc(&d, sizeof(double) * (nc + A)) != cudaSuccess) return NMF_ECUDA;
  cudaMemcpyAsync(d, coef_host, sizeof(double) * nc, cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(d + nc, last_host, sizeof(double) * A, cudaMemcpyHostToDevice, stream);
  const long long total = (long long)n_worlds * T * A;
  nmf_replay_table_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d, d + nc, n_int, A, fps, dt, n_part, T, total, fly_offset, out_dev);
  const cudaError_t e = cudaGetLastError();
  cudaStreamSynchronize(stream);      // the pageable host arrays and the scratch buffer must outlive the copies / the kernel
  cudaFree(d);
  return e == cudaSuccess ? NMF_OK : NMF_ECUDA;
}
