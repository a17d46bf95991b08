// brb_cabi.cu — the extern "C" boundary declared in include/brb.h: plain pointers and sizes, no torch
// types.  Owns device allocations for the SoA env state, launches the kernels in brb_kernels.cu.
// There is no CPU path: without a CUDA device every entry point returns BRB_ECUDA.

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/brb.h"
#include "brb_internal.h"

extern "C" {
void brb_launch_step(int kind, const BrbModelConsts *c, const BrbState *S, const BrbPerm *perm, const float *actions, float *obs, float *reward,
                     uint8_t *done, uint8_t *truncated, float *terminal_obs, float *ep_return, int32_t *ep_len,
                     const double *replay_u, int max_ctas, cudaStream_t stream);
int brb_step_resident_ctas(int kind, int device);
void brb_launch_group(long long n, const uint8_t *key, const unsigned *hist, unsigned *cursor, int *order, unsigned *hist_zero,
                      unsigned *cursor_zero, unsigned *queue_cursor, cudaStream_t stream);
void brb_launch_reset(int kind, const BrbState *S, float *obs, const double *replay_u, unsigned epoch, cudaStream_t stream);
void brb_launch_done_rows(long long n, const uint8_t *done, const uint8_t *truncated, const float *terminal_obs, const float *ep_return,
                          const int32_t *ep_len, unsigned *block_count, unsigned *block_base, unsigned *ticket, int *n_done,
                          uint32_t *rows, long long max_rows, cudaStream_t stream);
void brb_launch_get_state(const BrbState *S, double *qpos, double *qvel, double *xquat, cudaStream_t stream);
void brb_launch_set_state(const BrbState *S, const double *qpos, const double *qvel, cudaStream_t stream);
void brb_launch_get_elapsed(const BrbState *S, int32_t *out, cudaStream_t stream);
void brb_launch_ffma_probe(float *out, int blocks, int threads, int iters, cudaStream_t stream);
}

struct BrbModel {
  BrbModelConsts consts;
  double *time_table;   // device
  int n_time;
  int device;
};

struct BrbEnv {
  const BrbModel *model;
  BrbState S;
  void *arena;          // one allocation backing every SoA column
  int64_t launches;
  int *order;           // [N] visit order of the next step (see BrbPerm); valid when have_order
  uint8_t *keys;        // [N] group keys published by the last step
  unsigned *hist;       // [2][32] double-buffered histogram, [2][32] cursors behind it, then the step kernel's work-queue cursor
  int have_order, parity, sort_envs;
  int max_ctas;         // resident CTAs of the step kernel on this device (0 = no work queue: one robot per thread)
  // device staging for brb_env_step_host
  float *d_actions, *d_obs, *d_reward, *d_tobs, *d_epret;
  uint8_t *d_done, *d_trunc;
  int32_t *d_eplen;
  // finished-episode compaction for brb_env_step_host_compact
  unsigned *d_blk_count, *d_blk_base, *d_ticket;   // [ceil(N/256)] x 2, [1] ticket followed by [1] n_done
  uint32_t *d_rows;                                // [N][BRB_DONE_ROW_WORDS]
  int32_t *h_ndone;                                // pinned (mapped: the count kernel writes it directly)
  int32_t *h_ndone_dev;                            // device-side address of h_ndone
  cudaStream_t host_stream;
  // reset_all / set_state run on the caller's stream, the host-buffer step on host_stream: the next host step waits for this event
  cudaEvent_t user_ev;
  int user_ev_pending;
  uint32_t reset_epoch;                            // number of reset_all calls so far: folded into the Philox block index
};

// Every entry point runs on the env's / model's device and leaves the caller's current device as it found it.
struct DeviceGuard {
  int prev = -1, target;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) : target(dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
  }
  ~DeviceGuard() { if (prev >= 0 && prev != target) cudaSetDevice(prev); }
};
#define ON_DEVICE(dev) DeviceGuard guard_(dev); CK(guard_.err)

// BRB_DEBUG=1 in the environment prints the CUDA error string behind a BRB_ECUDA status
#define CK(x) do { cudaError_t ck_ = (x); if (ck_ != cudaSuccess) { if (getenv("BRB_DEBUG")) fprintf(stderr, "[brb] %s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(ck_)); cudaGetLastError(); return BRB_ECUDA; } } while (0)

extern "C" int brb_version(void) { return 100; }

extern "C" const char *brb_strerror(int code) {
  switch (code) {
    case BRB_OK: return "ok";
    case BRB_EINVAL: return "invalid argument";
    case BRB_ENOMEM: return "out of memory";
    case BRB_ECUDA: return "CUDA error (no device, launch failure or bad device pointer)";
    default: return "unknown error";
  }
}

extern "C" int brb_model_create(const BrbModelConsts *consts, const double *time_table_host, int n_time, int device, BrbModel **out) {
  if (!consts || !time_table_host || !out || n_time < consts->max_episode_steps + 2) return BRB_EINVAL;
  if (consts->env_kind < BRB_ENV01_V1 || consts->env_kind > BRB_ENV03_V2 || consts->frame_skip < 1) return BRB_EINVAL;
  if (consts->nq < 9 || consts->nq > 16 || consts->nv < 8 || consts->nv > 14) return BRB_EINVAL;
  ON_DEVICE(device);
  BrbModel *m = (BrbModel *)calloc(1, sizeof(BrbModel));
  if (!m) return BRB_ENOMEM;
  m->consts = *consts;
  m->n_time = n_time;
  m->device = device;
  if (cudaMalloc(&m->time_table, sizeof(double) * n_time) != cudaSuccess) { free(m); return BRB_ENOMEM; }
  if (cudaMemcpy(m->time_table, time_table_host, sizeof(double) * n_time, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(m->time_table); free(m); return BRB_ECUDA;
  }
  *out = m;
  return BRB_OK;
}

extern "C" void brb_model_destroy(BrbModel *m) {
  if (!m) return;
  cudaFree(m->time_table);
  free(m);
}

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int brb_env_create(const BrbModel *m, int64_t n, uint64_t seed, int64_t env_id_offset, BrbEnv **out) {
  if (!m || !out || n <= 0 || n > 0x7FFFFFFFLL) return BRB_EINVAL;      // the visit order holds 32-bit env indices
  ON_DEVICE(m->device);
  BrbEnv *e = (BrbEnv *)calloc(1, sizeof(BrbEnv));
  if (!e) return BRB_ENOMEM;
  e->model = m;
  const size_t N = (size_t)n;
  const size_t NQ = (size_t)m->consts.nq, NV = (size_t)m->consts.nv;
  const size_t sz[] = {
      align_up(NQ * N * 8), align_up(NV * N * 8), align_up(4 * N * 8), align_up(8 * N * 4), align_up(N * 8), align_up(N * 8),
      align_up(3 * N * 8), align_up(N * 4), align_up(N * 4), align_up(N * 4), align_up(BRB_NSTATS * 8), align_up(N * 4), align_up(N), align_up(4 * 32 * 4 + 64),
      // staging
      align_up(2 * N * 4), align_up(6 * N * 4), align_up(N * 4), align_up(N), align_up(6 * N * 4), align_up(N * 4), align_up(N), align_up(N * 4),
      // finished-episode compaction
      align_up(((N + 255) / 256) * 4), align_up(((N + 255) / 256) * 4), align_up(2 * 4), align_up(N * BRB_DONE_ROW_WORDS * 4)};
  size_t total = 0;
  for (size_t k = 0; k < sizeof(sz) / sizeof(sz[0]); k++) total += sz[k];
  if (cudaMalloc(&e->arena, total) != cudaSuccess) { cudaGetLastError(); free(e); return BRB_ENOMEM; }
  if (cudaMemset(e->arena, 0, total) != cudaSuccess) { cudaFree(e->arena); free(e); return BRB_ECUDA; }
  char *p = (char *)e->arena;
  int k = 0;
#define TAKE(T) (T *)p; p += sz[k++]
  e->S.qpos = TAKE(double);
  e->S.qvel = TAKE(double);
  e->S.xquat = TAKE(double);
  e->S.aset = TAKE(uint32_t);
  e->S.last_pitch = TAKE(double);
  e->S.ep_return = TAKE(double);
  e->S.v3 = TAKE(double);
  e->S.elapsed = TAKE(int);
  e->S.ep_len = TAKE(int);
  e->S.event = TAKE(uint32_t);
  e->S.stats = TAKE(unsigned long long);
  e->order = TAKE(int);
  e->keys = TAKE(uint8_t);
  e->hist = TAKE(unsigned);
  e->d_actions = TAKE(float);
  e->d_obs = TAKE(float);       // obs | reward | done are adjacent: one device-to-host copy on the host path
  e->d_reward = TAKE(float);
  e->d_done = TAKE(uint8_t);
  e->d_tobs = TAKE(float);
  e->d_epret = TAKE(float);
  e->d_trunc = TAKE(uint8_t);
  e->d_eplen = TAKE(int32_t);
  e->d_blk_count = TAKE(unsigned);
  e->d_blk_base = TAKE(unsigned);
  e->d_ticket = TAKE(unsigned);
  e->d_rows = TAKE(uint32_t);
#undef TAKE
  e->S.n = n;
  e->S.nq = m->consts.nq;
  e->S.nv = m->consts.nv;
  e->S.env0 = env_id_offset;
  e->S.seed = seed;
  e->S.time_table = m->time_table;
  e->parity = 0;
  e->have_order = 0;
  e->sort_envs = getenv("BRB_NO_SORT") ? 0 : 1;
  e->max_ctas = getenv("BRB_NO_QUEUE") ? 0 : brb_step_resident_ctas(m->consts.env_kind, m->device);
  if (cudaStreamCreateWithFlags(&e->host_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaFree(e->arena); free(e); return BRB_ECUDA; }
  if (cudaMallocHost(&e->h_ndone, sizeof(int32_t)) != cudaSuccess) { cudaStreamDestroy(e->host_stream); cudaFree(e->arena); free(e); return BRB_ENOMEM; }
  if (cudaHostGetDevicePointer((void **)&e->h_ndone_dev, e->h_ndone, 0) != cudaSuccess) { cudaGetLastError(); e->h_ndone_dev = nullptr; }
  if (cudaEventCreateWithFlags(&e->user_ev, cudaEventDisableTiming) != cudaSuccess) { cudaFreeHost(e->h_ndone); cudaStreamDestroy(e->host_stream); cudaFree(e->arena); free(e); return BRB_ECUDA; }
  *out = e;
  return BRB_OK;
}

extern "C" void brb_env_destroy(BrbEnv *e) {
  if (!e) return;
  DeviceGuard guard_(e->model->device);
  cudaStreamDestroy(e->host_stream);
  cudaEventDestroy(e->user_ev);
  cudaFreeHost(e->h_ndone);
  cudaFree(e->arena);
  free(e);
}

extern "C" int64_t brb_env_num_envs(const BrbEnv *e) { return e ? e->S.n : 0; }
extern "C" int64_t brb_env_num_launches(const BrbEnv *e) { return e ? e->launches : 0; }

// One VecEnv.step = the fused step kernel + the (tiny) grouping kernel that sorts the envs for the next step.
// All launches of one env object must be stream-ordered with respect to each other, as for any stateful env.
static void launch_step(BrbEnv *e, const float *actions, float *obs, float *reward, uint8_t *done, uint8_t *truncated,
                        float *terminal_obs, float *ep_return, int32_t *ep_len, const double *replay_u, cudaStream_t stream) {
  BrbPerm perm = {nullptr, nullptr, nullptr, nullptr};
  unsigned *hist = e->hist + 32 * e->parity, *cursor = e->hist + 64 + 32 * e->parity, *queue = e->hist + 128;
  if (e->max_ctas > 0) {
    perm.cursor = queue;
    if (!e->sort_envs) cudaMemsetAsync(queue, 0, sizeof(unsigned), stream);     // otherwise the grouping kernel re-arms it
  }
  if (e->sort_envs) {
    perm.in = e->have_order ? e->order : nullptr;
    perm.key_out = e->keys;
    perm.hist = hist;
  }
  brb_launch_step(e->model->consts.env_kind, &e->model->consts, &e->S, &perm, actions, obs, reward, done, truncated, terminal_obs,
                  ep_return, ep_len, replay_u, e->max_ctas, stream);
  e->launches++;
  if (e->sort_envs) {
    brb_launch_group(e->S.n, e->keys, hist, cursor, e->order, e->hist + 32 * (e->parity ^ 1), e->hist + 64 + 32 * (e->parity ^ 1), perm.cursor, stream);
    e->launches++;
    e->have_order = 1;
    e->parity ^= 1;
  }
}

extern "C" int brb_env_reset_all(BrbEnv *e, float *obs, const double *replay_u_reset, void *stream) {
  if (!e || !obs) return BRB_EINVAL;
  ON_DEVICE(e->model->device);
  brb_launch_reset(e->model->consts.env_kind, &e->S, obs, replay_u_reset, e->reset_epoch, (cudaStream_t)stream);
  e->reset_epoch++;
  e->launches++;
  e->have_order = 0;
  CK(cudaGetLastError());
  CK(cudaEventRecord(e->user_ev, (cudaStream_t)stream));
  e->user_ev_pending = 1;
  return BRB_OK;
}

extern "C" int brb_env_step(BrbEnv *e, const float *actions, float *obs, float *reward, uint8_t *done, uint8_t *truncated,
                            float *terminal_obs, float *ep_return, int32_t *ep_len, const double *replay_u, void *stream) {
  if (!e || !actions || !obs) return BRB_EINVAL;
  ON_DEVICE(e->model->device);
  launch_step(e, actions, obs, reward, done, truncated, terminal_obs, ep_return, ep_len, replay_u, (cudaStream_t)stream);
  CK(cudaGetLastError());
  return BRB_OK;
}

extern "C" int brb_env_step_host(BrbEnv *e, const float *actions, float *obs, float *reward, uint8_t *done, uint8_t *truncated,
                                 float *terminal_obs, float *ep_return, int32_t *ep_len) {
  if (!e || !actions || !obs) return BRB_EINVAL;
  ON_DEVICE(e->model->device);
  const size_t N = (size_t)e->S.n;
  cudaStream_t s = e->host_stream;
  if (e->user_ev_pending) { CK(cudaStreamWaitEvent(s, e->user_ev, 0)); e->user_ev_pending = 0; }
  CK(cudaMemcpyAsync(e->d_actions, actions, 2 * N * sizeof(float), cudaMemcpyHostToDevice, s));
  launch_step(e, e->d_actions, e->d_obs, e->d_reward, e->d_done, e->d_trunc, e->d_tobs, e->d_epret, e->d_eplen, nullptr, s);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(obs, e->d_obs, 6 * N * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (reward) CK(cudaMemcpyAsync(reward, e->d_reward, N * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (done) CK(cudaMemcpyAsync(done, e->d_done, N, cudaMemcpyDeviceToHost, s));
  if (truncated) CK(cudaMemcpyAsync(truncated, e->d_trunc, N, cudaMemcpyDeviceToHost, s));
  if (terminal_obs) CK(cudaMemcpyAsync(terminal_obs, e->d_tobs, 6 * N * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (ep_return) CK(cudaMemcpyAsync(ep_return, e->d_epret, N * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (ep_len) CK(cudaMemcpyAsync(ep_len, e->d_eplen, N * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return BRB_OK;
}

// BRB_PROFILE_HOST=1 (kernel-tuning experiments): host-side time line of the host-buffer step, printed every 256 calls
#include <chrono>
static double hostprof_acc[4]; static long hostprof_n; static int hostprof_on = -1;
#define HOSTPROF_T(k) std::chrono::steady_clock::time_point hp_t##k; if (hostprof_on < 0) hostprof_on = getenv("BRB_PROFILE_HOST") ? 1 : 0; if (hostprof_on) hp_t##k = std::chrono::steady_clock::now()
#define HOSTPROF_END() do { if (hostprof_on) { hostprof_acc[0] += std::chrono::duration<double, std::micro>(hp_t1 - hp_t0).count(); \
  hostprof_acc[1] += std::chrono::duration<double, std::micro>(hp_t2 - hp_t1).count(); hostprof_acc[2] += std::chrono::duration<double, std::micro>(hp_t3 - hp_t2).count(); \
  if (++hostprof_n % 256 == 0) { fprintf(stderr, "[brb host path] H2D enqueue %.1f us, launches + D2H enqueue %.1f us, wait %.1f us\n", hostprof_acc[0] / 256, hostprof_acc[1] / 256, hostprof_acc[2] / 256); hostprof_acc[0] = hostprof_acc[1] = hostprof_acc[2] = 0; } } } while (0)

// device-side address of a pinned (page-locked, mapped) host buffer, or NULL for pageable memory
static void *mapped_device_ptr(const void *host) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

extern "C" int brb_env_host_layout(const BrbEnv *e, int64_t offsets[3], int64_t *total_bytes) {
  if (!e || !offsets || !total_bytes) return BRB_EINVAL;
  offsets[0] = 0;
  offsets[1] = (char *)e->d_reward - (char *)e->d_obs;
  offsets[2] = (char *)e->d_done - (char *)e->d_obs;
  *total_bytes = offsets[2] + e->S.n;
  return BRB_OK;
}

extern "C" int brb_env_step_host_compact(BrbEnv *e, const float *actions, float *obs, float *reward, uint8_t *done, int32_t *n_done,
                                         uint32_t *done_rows, int64_t max_rows) {
  if (!e || !actions || !obs || !n_done || (max_rows > 0 && !done_rows) || max_rows < 0) return BRB_EINVAL;
  ON_DEVICE(e->model->device);
  const size_t N = (size_t)e->S.n;
  cudaStream_t s = e->host_stream;
  HOSTPROF_T(0);
  if (e->user_ev_pending) { CK(cudaStreamWaitEvent(s, e->user_ev, 0)); e->user_ev_pending = 0; }
  CK(cudaMemcpyAsync(e->d_actions, actions, 2 * N * sizeof(float), cudaMemcpyHostToDevice, s));
  HOSTPROF_T(1);
  launch_step(e, e->d_actions, e->d_obs, e->d_reward, e->d_done, e->d_trunc, e->d_tobs, e->d_epret, e->d_eplen, nullptr, s);
  // finished-episode rows: when the caller's row buffer is pinned the compaction kernel writes the rows (and the count) straight
  // into host memory over PCIe -- a few KB of posted writes -- so the call needs ONE stream synchronisation; a pageable buffer
  // takes the count first and the rows in a second copy
  uint32_t *rows_dev = (max_rows > 0 && e->h_ndone_dev) ? (uint32_t *)mapped_device_ptr(done_rows) : nullptr;
  const bool direct = rows_dev != nullptr || (max_rows == 0 && e->h_ndone_dev);
  brb_launch_done_rows(e->S.n, e->d_done, e->d_trunc, e->d_tobs, e->d_epret, e->d_eplen, e->d_blk_count, e->d_blk_base, e->d_ticket,
                       direct ? (int *)e->h_ndone_dev : (int *)(e->d_ticket + 1), direct ? rows_dev : e->d_rows, direct ? (long long)max_rows : (long long)N, s);
  e->launches += 2;
  CK(cudaGetLastError());
  if (!direct) CK(cudaMemcpyAsync(e->h_ndone, e->d_ticket + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  // obs | reward | done leave in one copy when the host buffers are laid out like the device staging block (brb_env_host_layout)
  const ptrdiff_t off_r = (char *)e->d_reward - (char *)e->d_obs, off_d = (char *)e->d_done - (char *)e->d_obs;
  // (having the step kernel write obs / reward into the pinned buffers itself was measured slower: 981 vs 907 us per step --
  // scattered 24-byte PCIe writes bunch up when the expensive warps finish together)
  if (reward && done && (char *)reward - (char *)obs == off_r && (char *)done - (char *)obs == off_d) {
    CK(cudaMemcpyAsync(obs, e->d_obs, (size_t)off_d + N, cudaMemcpyDeviceToHost, s));
  } else {
    CK(cudaMemcpyAsync(obs, e->d_obs, 6 * N * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (reward) CK(cudaMemcpyAsync(reward, e->d_reward, N * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (done) CK(cudaMemcpyAsync(done, e->d_done, N, cudaMemcpyDeviceToHost, s));
  }
  HOSTPROF_T(2);
  CK(cudaStreamSynchronize(s));
  HOSTPROF_T(3);
  HOSTPROF_END();
  const int32_t nd = *e->h_ndone;
  *n_done = nd;
  if (nd > max_rows) return BRB_EINVAL;
  if (nd > 0 && !direct) {
    CK(cudaMemcpyAsync(done_rows, e->d_rows, (size_t)nd * BRB_DONE_ROW_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  return BRB_OK;
}

extern "C" int brb_env_get_state(BrbEnv *e, double *qpos, double *qvel, double *xquat, void *stream) {
  if (!e) return BRB_EINVAL;
  ON_DEVICE(e->model->device);
  brb_launch_get_state(&e->S, qpos, qvel, xquat, (cudaStream_t)stream);
  e->launches++;
  CK(cudaGetLastError());
  return BRB_OK;
}

extern "C" int brb_env_set_state(BrbEnv *e, const double *qpos, const double *qvel, void *stream) {
  if (!e || !qpos || !qvel) return BRB_EINVAL;
  ON_DEVICE(e->model->device);
  brb_launch_set_state(&e->S, qpos, qvel, (cudaStream_t)stream);
  e->launches++;
  CK(cudaGetLastError());
  CK(cudaEventRecord(e->user_ev, (cudaStream_t)stream));
  e->user_ev_pending = 1;
  return BRB_OK;
}

extern "C" int brb_env_get_elapsed(BrbEnv *e, int32_t *elapsed, void *stream) {
  if (!e || !elapsed) return BRB_EINVAL;
  ON_DEVICE(e->model->device);
  brb_launch_get_elapsed(&e->S, elapsed, (cudaStream_t)stream);
  e->launches++;
  CK(cudaGetLastError());
  return BRB_OK;
}

extern "C" int brb_env_get_stats(BrbEnv *e, uint64_t out[BRB_NSTATS]) {
  if (!e || !out) return BRB_EINVAL;
  ON_DEVICE(e->model->device);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out, e->S.stats, sizeof(uint64_t) * BRB_NSTATS, cudaMemcpyDeviceToHost));
  return BRB_OK;
}

extern "C" int brb_fp32_peak_flops(int device, double *flops_out, double *ms_out) {
  if (!flops_out) return BRB_EINVAL;
  ON_DEVICE(device);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 4096;
  float *buf;
  CK(cudaMalloc(&buf, sizeof(float) * (size_t)threads * blocks));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 1e30;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0, 0);
    brb_launch_ffma_probe(buf, blocks, threads, iters, 0);
    cudaEventRecord(e1, 0);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(buf); return BRB_ECUDA; }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf);
  const double flop = 2.0 * 8 * 16 * (double)iters * threads * blocks;
  *flops_out = flop / (best * 1e-3);
  if (ms_out) *ms_out = best;
  return BRB_OK;
}
