// nmf_capi.cu — C ABI (include/nmf_b200.h) + kernel launches for the sm_100a step path.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <new>
#include <string>

#include "../../include/nmf_b200.h"
#include "nmf_host.h"
#include "nmf_step.cuh"

using namespace nmf;

// ------------------------------------------------------------------ kernels
#ifndef NMF_MINBLOCKS
#define NMF_MINBLOCKS 16   // <= 64 registers/thread: best measured trade-off between occupancy and spills (profiles/)
#endif
extern "C" __global__ void __launch_bounds__(CTA, NMF_MINBLOCKS) nmf_step_kernel(const StepParams p) {
  __shared__ __align__(16) float sm[SM_TOTAL];
  step_block(p, sm);
}

__global__ void nmf_reset_kernel(float* state, const float* key, const uint8_t* mask, int n) {
  int fly = blockIdx.x;
  if (fly >= n || (mask && !mask[fly])) return;
  for (int i = threadIdx.x; i < S_STRIDE; i += blockDim.x) state[(size_t)fly * S_STRIDE + i] = key[i];
}

__global__ void nmf_scatter_cols_kernel(float* state, int off, const float* src, const int32_t* cols, int ncols, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * ncols) return;
  int fly = i / ncols, k = i - fly * ncols;
  state[(size_t)fly * S_STRIDE + off + cols[k]] = src[i];
}

__global__ void nmf_gather_cols_kernel(const float* state, int off, const int32_t* cols, int ncols, float* dst, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * ncols) return;
  int fly = i / ncols, k = i - fly * ncols;
  dst[i] = state[(size_t)fly * S_STRIDE + off + (cols ? cols[k] : k)];
}

// ------------------------------------------------------------------ handle
struct nmf_handle {
  HostModel hm;
  int n_flies = 0, device = 0;
  float *d_role = nullptr, *d_hull = nullptr, *d_seg = nullptr, *d_key = nullptr;
  int *d_nbr_adr = nullptr, *d_nbr = nullptr;
  float *d_act = nullptr, *d_qpos = nullptr;   // staging for nmf_step_host
  nmf_buffers buf{};
  bool bound = false;
  int64_t launches = 0;
  std::string err;
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return NMF_ECUDA; } } while (0)

static int upload(nmf_handle* h, float** dst, const std::vector<float>& v) {
  CK(cudaMalloc(dst, sizeof(float) * (v.size() ? v.size() : 1)));
  CK(cudaMemcpy(*dst, v.data(), sizeof(float) * v.size(), cudaMemcpyHostToDevice));
  return NMF_OK;
}

extern "C" int nmf_create(const void* blob, size_t nbytes, int n_flies, int device, nmf_handle** out) {
  if (!out) return NMF_EINVAL;
  *out = nullptr;
  nmf_handle* h = new (std::nothrow) nmf_handle;
  if (!h) return NMF_EINVAL;
  *out = h;   // returned even on failure so that nmf_last_error() can be read
  if (n_flies <= 0) { h->err = "n_flies must be positive"; return NMF_EINVAL; }
  if (!h->hm.build(blob, nbytes)) { h->err = h->hm.err; return NMF_EINVAL; }
  h->n_flies = n_flies; h->device = device;
  CK(cudaSetDevice(device));
  int rc;
  if ((rc = upload(h, &h->d_role, h->hm.role))) return rc;
  if ((rc = upload(h, &h->d_hull, h->hm.hull))) return rc;
  if ((rc = upload(h, &h->d_seg, h->hm.seg_tab))) return rc;
  if ((rc = upload(h, &h->d_key, h->hm.key_state))) return rc;
  CK(cudaMalloc(&h->d_nbr_adr, sizeof(int) * h->hm.hull_nbr_adr.size()));
  CK(cudaMemcpy(h->d_nbr_adr, h->hm.hull_nbr_adr.data(), sizeof(int) * h->hm.hull_nbr_adr.size(), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&h->d_nbr, sizeof(int) * h->hm.hull_nbr.size()));
  CK(cudaMemcpy(h->d_nbr, h->hm.hull_nbr.data(), sizeof(int) * h->hm.hull_nbr.size(), cudaMemcpyHostToDevice));
  return NMF_OK;
}

extern "C" int nmf_destroy(nmf_handle* h) {
  if (!h) return NMF_OK;
  cudaFree(h->d_role); cudaFree(h->d_hull); cudaFree(h->d_seg); cudaFree(h->d_key); cudaFree(h->d_nbr_adr); cudaFree(h->d_nbr); cudaFree(h->d_act); cudaFree(h->d_qpos);
  delete h;
  return NMF_OK;
}

extern "C" const char* nmf_last_error(const nmf_handle* h) { return h ? h->err.c_str() : "null handle"; }
extern "C" int64_t nmf_launch_count(const nmf_handle* h) { return h ? h->launches : 0; }

extern "C" int nmf_model_info(const nmf_handle* h, nmf_info* info) {
  if (!h || !info) return NMF_EINVAL;
  info->n_flies = h->n_flies; info->nq = NQ; info->nv = NV; info->nu_pos = h->hm.par.nu_pos; info->nu_adh = h->hm.par.nu_adh;
  info->nseg = h->hm.nseg; info->nleg = NLEG; info->state_stride = S_STRIDE; info->off_qpos = S_QPOS; info->off_qvel = S_QVEL;
  info->off_qacc_warmstart = S_WARM; info->off_ctrl = S_CTRL; info->off_time = S_TIME; info->dbg_stride = DBG_STRIDE;
  info->timestep = h->hm.par.dt;
  return NMF_OK;
}

extern "C" int nmf_bind(nmf_handle* h, const nmf_buffers* b) {
  if (!h || !b || !b->state) { if (h) h->err = "nmf_bind: state buffer is required"; return NMF_EINVAL; }
  h->buf = *b; h->bound = true;
  return NMF_OK;
}

extern "C" int nmf_set_solver(nmf_handle* h, int max_newton, int max_ls) {
  if (!h || max_newton < 1 || max_ls < 1) return NMF_EINVAL;
  h->hm.par.max_newton = max_newton; h->hm.par.max_ls = max_ls;
  return NMF_OK;
}

extern "C" int nmf_reset(nmf_handle* h, const uint8_t* mask, void* stream) {
  if (!h) return NMF_EINVAL;
  if (!h->bound) { h->err = "nmf_reset: not bound"; return NMF_ENOTBOUND; }
  nmf_reset_kernel<<<h->n_flies, 64, 0, (cudaStream_t)stream>>>(h->buf.state, h->d_key, mask, h->n_flies);
  h->launches++;
  CK(cudaGetLastError());
  return NMF_OK;
}

extern "C" int nmf_step(nmf_handle* h, int nsteps, const float* table, int table_T, int table_t0, void* stream) {
  if (!h) return NMF_EINVAL;
  if (!h->bound) { h->err = "nmf_step: not bound"; return NMF_ENOTBOUND; }
  if (nsteps <= 0) return NMF_OK;
  if (table && table_T <= 0) { h->err = "nmf_step: action table needs table_T > 0"; return NMF_EINVAL; }
  StepParams p = h->hm.par;
  p.state = h->buf.state; p.role = h->d_role; p.hull = h->d_hull; p.seg_tab = h->d_seg; p.hull_nbr_adr = h->d_nbr_adr; p.hull_nbr = h->d_nbr;
  p.act_table = table; p.table_T = table_T; p.table_t0 = table_t0;
  p.out_xpos = h->buf.seg_xpos; p.out_xquat = h->buf.seg_xquat; p.out_actf = h->buf.act_force; p.out_sensor = h->buf.sensordata;
  p.dbg = h->buf.debug; p.n_flies = h->n_flies; p.nsteps = nsteps;
  nmf_step_kernel<<<h->n_flies, CTA, 0, (cudaStream_t)stream>>>(p);
  h->launches++;
  CK(cudaGetLastError());
  return NMF_OK;
}

extern "C" int nmf_scatter_ctrl(nmf_handle* h, const float* src, const int32_t* cols, int ncols, void* stream) {
  if (!h || !src || !cols || ncols <= 0) return NMF_EINVAL;
  if (!h->bound) return NMF_ENOTBOUND;
  int total = h->n_flies * ncols;
  nmf_scatter_cols_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->buf.state, S_CTRL, src, cols, ncols, h->n_flies);
  h->launches++;
  CK(cudaGetLastError());
  return NMF_OK;
}

extern "C" int nmf_gather_state(nmf_handle* h, int off, const int32_t* cols, int ncols, float* dst, void* stream) {
  if (!h || !dst || ncols <= 0 || off < 0 || off >= S_STRIDE) return NMF_EINVAL;
  if (!h->bound) return NMF_ENOTBOUND;
  int total = h->n_flies * ncols;
  nmf_gather_cols_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->buf.state, off, cols, ncols, dst, h->n_flies);
  h->launches++;
  CK(cudaGetLastError());
  return NMF_OK;
}

extern "C" int nmf_step_host(nmf_handle* h, const float* actions_host, int nsteps, float* qpos_host, void* stream_) {
  if (!h || !actions_host || !qpos_host) return NMF_EINVAL;
  if (!h->bound) return NMF_ENOTBOUND;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int nu_pos = h->hm.par.nu_pos, n = h->n_flies;
  if (!h->d_act) { CK(cudaMalloc(&h->d_act, sizeof(float) * (size_t)n * nu_pos)); CK(cudaMalloc(&h->d_qpos, sizeof(float) * (size_t)n * NQ)); }
  CK(cudaMemcpyAsync(h->d_act, actions_host, sizeof(float) * (size_t)n * nu_pos, cudaMemcpyHostToDevice, stream));
  // the action block doubles as a 1-row action table: ctrl[0:nu_pos] <- actions (position actuators are ctrl 0..nu_pos-1)
  int rc = nmf_step(h, nsteps, h->d_act, 1, 0, stream);
  if (rc) return rc;
  rc = nmf_gather_state(h, S_QPOS, nullptr, NQ, h->d_qpos, stream);
  if (rc) return rc;
  CK(cudaMemcpyAsync(qpos_host, h->d_qpos, sizeof(float) * (size_t)n * NQ, cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  return NMF_OK;
}
