/*
 * brb.h — C-ABI of the B200-native batched balance-robot environment step.
 *
 * This is the drop-in boundary for the reference's hot path.  The reference has no FFI of its own:
 * the seam is the Gymnasium Env API consumed through SB3's VecEnv, i.e. the Python
 * calls listed next to each entry point below.  All `*_dev` style pointers are DEVICE pointers
 * (e.g. torch tensor .data_ptr()); `stream` is a cudaStream_t passed as void* (NULL = legacy default
 * stream).  Every function returns 0 on success or a negative BRB_E* code; nothing throws; no CPU
 * fallback exists — without a CUDA device every compute entry point returns BRB_ECUDA.
 *
 * Layouts: obs [N,6] f32 row-major; actions [N,2] f32; reward [N] f32; done/truncated [N] u8;
 * terminal_obs [N,6] f32; ep_return [N] f32; ep_len [N] i32; qpos [N,nq] f64; qvel [N,nv] f64 (nq,nv = 9,8; Env03-v2: 16,14).
 */
#ifndef BRB_H
#define BRB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BRB_OK 0
#define BRB_EINVAL (-22)
#define BRB_ENOMEM (-12)
#define BRB_ECUDA (-5)

#define BRB_ENV01_V1 0 /* reference balance_robot/__init__.py:5-10  -> envs/env01_v1.py:10 */
#define BRB_ENV01_V2 1 /* reference balance_robot/__init__.py:12-17 -> envs/env01_v2.py:14 */
#define BRB_ENV01_V3 2 /* reference balance_robot/__init__.py:19-24 -> envs/env01_v3.py:13 */
#define BRB_ENV03_V2 3 /* reference balance_robot/__init__.py:47-52 -> envs/env03_v2.py:14 (robot + fired block) */

#define BRB_FLAG_ACTDERIV_SKIP_CLAMPED 1
/* Opt-in (off by default: the reference has no such rule): an env whose step ends in a pose with contacts the kernels do not
 * generate (BRB_STAT_UNSUPPORTED) is ended as TRUNCATED (TimeLimit.truncated = True, so a trainer bootstraps the value) and
 * auto-reset, instead of stepping on with those contacts missing. */
#define BRB_FLAG_TRUNCATE_UNSUPPORTED 2
/* Env03-v2, opt-in: generate wheel-block contacts (own analytic cylinder-box collider, brb_env03.cuh).  Off (default) = no contact for
 * that pair, poses within reach are counted as unsupported.  On costs ~2.3x step time at 65,536 robots with random actions: the
 * narrow phase and the solve with wheel rows are rare divergent paths inside CTAs that walk the substep loop in lockstep. */
#define BRB_FLAG_WHEEL_BLOCK 4

#define BRB_NSTATS 12
#define BRB_STAT_SUBSTEPS 0          /* env-substeps executed */
#define BRB_STAT_CONTACT_SUBSTEPS 1  /* of which had >= 1 wheel-floor contact */
#define BRB_STAT_SOLVES 2            /* 8x8 factorisations performed */
#define BRB_STAT_NONCONVERGED 3      /* substeps that hit the active-set iteration cap */
#define BRB_STAT_UNSUPPORTED 4       /* env-steps that ended in a pose whose contacts the kernel does not model */
#define BRB_STAT_EPISODES 5          /* episodes finished */
#define BRB_STAT_ENV_STEPS 6
#define BRB_STAT_CONTACT_SLOTS 7     /* sum over contact substeps of the number of wheel-floor contacts (1..4) */
#define BRB_STAT_COUPLED_SUBSTEPS 8  /* Env03-v2: substeps solved through the coupled 14-dof path (block touching the chassis) */
#define BRB_STAT_BLOCK_CONTACT_SUBSTEPS 9 /* Env03-v2: substeps with the block on the floor */
#define BRB_STAT_COUPLED_FALLBACKS 10 /* Env03-v2: coupled substeps finished by the generic line-search solver */
#define BRB_STAT_COUPLED_SOLVES 11    /* Env03-v2: fast coupled Newton steps */

/* Per-model constant block, produced on the host by balance_robot_b200/model.py from the MJCF
 * (stands in for MjModel.from_xml_path, reference envs/RobotBaseEnv.py:56-65).  Passed to the
 * kernels by value (constant bank). */
typedef struct BrbModelConsts {
  float h, grav, mass, mcz;
  float Ixx, Iyy, Izz, Ia;
  float minv_xy[3], minv_uz, minv_wz, minv_blk[10];
  float ox, oz, rad, hl, zfloor, zfloor_lo; /* floor height = zfloor + zfloor_lo (hi/lo split of the fp64 value) */
  float damping, kv, ctrl_lo, ctrl_hi, frc_lo, frc_hi;
  float mu, D, Kimp, Bdamp;
  float impl_W[8], impl_G[3], impl_cinv_full, impl_cinv_damp;
  float impl_Kinv[4][3]; /* (Cinv + G)^-1 = (i00, i01, i11) per servo clamp state, index = clampL + 2 clampR */
  float chassis_half[3], chassis_pos[3];
  int frame_skip, max_episode_steps, env_kind, flags;
  /* Env03-v2 only.  pp[k] = {mu, K, B, D1, d0, d1, width, margin} of the dynamic pairs k = 0 wheel-floor, 1 block-floor,
   * 2 chassis-block (impedance imp(dist) from d0, d1, width; row D = D1 * imp / (1 - imp); aref = -B vel - K imp (dist - margin)) */
  float pp[3][8];
  float blk_half, blk_mass, blk_inertia, blk_radius, chassis_radius;
  float geo_lo[4]; /* fp64 - fp32 residuals of ox, oz, rad, hl: the contact on/off predicate is re-evaluated in fp64 near dist = 0 */
  int nq, nv;
  float wb_D1;     /* Env03-v2: row-weight scale D1 of the wheel-block pair (its other parameters equal the chassis-block pair's pp[2]) */
} BrbModelConsts;

typedef struct BrbModel BrbModel;
typedef struct BrbEnv BrbEnv;

int brb_version(void);
const char *brb_strerror(int code);

/* MjModel.from_xml_path: uploads the constant block and the fp64 step-time table
 * (time_table[k] = data.time after k env steps, k = 0..n_time-1). */
int brb_model_create(const BrbModelConsts *consts, const double *time_table_host, int n_time, int device, BrbModel **out);
void brb_model_destroy(BrbModel *m);

/* gym.make(id) x n_envs (reference sb_rl.py:500).  env_id_offset = global id of env 0 of this shard, so
 * the Philox streams are invariant to how envs are sharded across GPUs. */
int brb_env_create(const BrbModel *m, int64_t n_envs, uint64_t seed, int64_t env_id_offset, BrbEnv **out);
void brb_env_destroy(BrbEnv *e);

/* VecEnv.reset() -> MujocoEnv.reset -> reset_model (reference envs/env01_v1.py:39-58, env01_v2.py:52-71,
 * env01_v3.py:39-54, env03_v1.py:60-83).  replay_u_reset: optional [N,16] ([N,32] for Env03-v2) f64 uniforms replacing Philox. */
int brb_env_reset_all(BrbEnv *e, float *obs, const double *replay_u_reset, void *stream);

/* VecEnv.step(actions) (DummyVecEnv auto-reset + TimeLimit + Monitor around reference
 * envs/env01_v1.py:15-37 / env01_v2.py:28-50 / env01_v3.py:27-37, which call mujoco.mj_step x250).
 * replay_u: optional [N,20] f64 uniforms (4 step-noise slots then 16 reset slots) replacing Philox
 * (Env03-v2: [N,40] = 8 re-fire slots then 32 reset slots).
 * terminal_obs rows are written only where done; ep_return / ep_len hold the running episode
 * statistics (the finished episode's totals where done). Any output pointer except obs may be NULL. */
int brb_env_step(BrbEnv *e, const float *actions, float *obs, float *reward, uint8_t *done, uint8_t *truncated,
                 float *terminal_obs, float *ep_return, int32_t *ep_len, const double *replay_u, void *stream);

/* Same call with HOST buffers (pinned or pageable): copies actions in, steps, copies results out and
 * synchronises the stream.  This is the end-to-end path a host-side trainer (SB3 on CPU tensors) uses. */
int brb_env_step_host(BrbEnv *e, const float *actions, float *obs, float *reward, uint8_t *done, uint8_t *truncated,
                      float *terminal_obs, float *ep_return, int32_t *ep_len);

/* Same step, but the per-episode outputs cross PCIe only for the envs that finished: the device compacts them (ascending
 * env index) into rows of BRB_DONE_ROW_WORDS 32-bit words
 *   [env index (i32) | terminal_observation[6] (f32) | ep_return (f32) | ep_len (i32) | TimeLimit.truncated (i32)]
 * = exactly what SB3's DummyVecEnv / Monitor put into infos[i] of a finished env (sb_rl.py:500-501).  obs / reward / done
 * are full [N] arrays as above.  *n_done receives the number of rows written to done_rows (capacity max_rows rows;
 * BRB_EINVAL if more envs finished than fit). */
#define BRB_DONE_ROW_WORDS 10
int brb_env_step_host_compact(BrbEnv *e, const float *actions, float *obs, float *reward, uint8_t *done, int32_t *n_done,
                              uint32_t *done_rows, int64_t max_rows);

/* Byte offsets of obs / reward / done inside the device staging block of the host path (offsets[0] = 0) and its total size.
 * A caller that carves its three host buffers out of ONE pinned allocation at these offsets gets them back in a single
 * device-to-host copy from brb_env_step_host_compact; any other layout is served by three copies.  (A pinned done_rows
 * buffer is written by the compaction kernel directly, a pageable one through a second copy.) */
int brb_env_host_layout(const BrbEnv *e, int64_t offsets[3], int64_t *total_bytes);

/* MujocoEnv.set_state / data.qpos, data.qvel access (trajectory checks).  xquat = the (stale, Q1) chassis
 * quaternion the observation functions read; may be NULL. */
int brb_env_get_state(BrbEnv *e, double *qpos, double *qvel, double *xquat, void *stream);
int brb_env_set_state(BrbEnv *e, const double *qpos, const double *qvel, void *stream);
/* per-env episode bookkeeping (elapsed steps); out [N] i32 */
int brb_env_get_elapsed(BrbEnv *e, int32_t *elapsed, void *stream);

/* Synchronises and copies the cumulative counters (BRB_STAT_*) to the host. */
int brb_env_get_stats(BrbEnv *e, uint64_t out[BRB_NSTATS]);
int64_t brb_env_num_envs(const BrbEnv *e);
/* kernel launches issued so far by this env object (for bench.py's gpu_launches) */
int64_t brb_env_num_launches(const BrbEnv *e);

/* FP32-pipe peak probe: runs an FFMA-bound kernel and returns achieved FLOP/s (for the roofline denominator). */
int brb_fp32_peak_flops(int device, double *flops_out, double *ms_out);

/* PPO actor-critic forward for the rollout: policy.forward(obs) of SB3's ActorCriticPolicy("MlpPolicy") as the reference
 * configures it (src/sb_rl.py:63-71: pi = vf = [64, 64], tanh, state-independent log_std), one launch for n robots.
 * params: BRB_POLICY_NPARAM fp32 on the device in SB3 state-dict order
 *   mlp_extractor.policy_net.{0,2}.{weight,bias}, mlp_extractor.value_net.{0,2}.{weight,bias},
 *   action_net.{weight,bias}, value_net.{weight,bias}, log_std.
 * obs [n,6]; noise [n,2] standard normal draws (NULL = deterministic: actions = mean).  Outputs: actions [n,2] (the
 * unclipped sample the rollout buffer keeps), actions_clipped [n,2] (what the env receives; may be NULL), values [n],
 * log_prob [n] of the unclipped action.  actions == NULL: critic only (values of the given observations, e.g. V(terminal
 * observation) for the TimeLimit bootstrap).  All pointers are device pointers. */
#define BRB_POLICY_NPARAM 9413
int brb_policy_act(const float *params, const float *obs, const float *noise, int64_t n, float *actions, float *actions_clipped,
                   float *values, float *logp, void *stream);

/* Critic only, for the rows whose mask byte is non-zero (values of the other rows = 0): gamma * V(terminal_observation) of SB3's
 * TimeLimit bootstrap (on_policy_algorithm.collect_rollouts, third party) without a host-side "any truncated?" test. */
int brb_policy_value_masked(const float *params, const float *obs, const uint8_t *mask, int64_t n, float *values, void *stream);

/* One PPO minibatch, forward + loss + backward fused (SB3 PPO.train() inner loop, third party; reference src/sb_rl.py:63-71
 * runs it with SB3's defaults): for the mb samples idx[0..mb) of the rollout buffer (obs [S,6], actions [S,2] unclipped,
 * old_logp / adv / returns [S])
 *   loss = -mean(min(A r, A clamp(r, 1 - clip, 1 + clip))) + vf_coef mean((returns - V)^2) - ent_coef mean(entropy),
 *   r = exp(log_prob - old_logp), A = (adv - adv_stats[0]) * adv_stats[1]   (per-minibatch normalisation: mean, 1 / (std + eps))
 * grad [BRB_POLICY_NPARAM] (parameter-block layout) and stats[4] = {policy_loss, value_loss, approx_kl, clip_fraction}
 * are ACCUMULATED: zero them before the call.  All pointers are device pointers. */
/* Runs on the tensor cores (tcgen05.mma, bf16 hi/lo split = three passes per product, fp32 accumulation in TMEM; csrc/brb_policy_tc.cu).
 * brb_ppo_tc_fault(device) synchronises and returns 1 if one of that kernel's bounded pipeline waits ever timed out (0 otherwise). */
int brb_ppo_tc_fault(int device);
int brb_ppo_grad(const float *params, const float *obs, const float *actions, const float *old_logp, const float *adv, const float *returns,
                 const int64_t *idx, int64_t mb, const float *adv_stats, float clip_range, float vf_coef, float ent_coef, float *grad,
                 float *stats, void *stream);

/* out[0..n) <- a pseudo-random permutation of 0..n-1 keyed by `seed` (device pointer, int64): the minibatch order of one epoch of
 * SB3 PPO.train() (RolloutBuffer.get -> np.random.permutation; third party, reference src/sb_rl.py:63-71) without a sort. */
int brb_random_permutation(int64_t *out, int64_t n, uint64_t seed, void *stream);

/* One optimiser step on the flat parameter block, fused into one launch: g = grad * grad_scale (1 / world after an all-reduce
 * sum), th.nn.utils.clip_grad_norm_(max_grad_norm) (<= 0: no clipping), torch.optim.Adam (SB3 PPO.train(): policy.optimizer.step(),
 * third party; reference src/sb_rl.py:63-71).  m / v = Adam moments [n], step = 1-based step count, norm_out (nullable) receives
 * the pre-clip gradient norm.  grad is zeroed on return (brb_ppo_grad accumulates).  All pointers are device pointers. */
int brb_adam_clip_step(float *params, float *grad, float *m, float *v, int64_t n, float lr, float beta1, float beta2, float eps,
                       int64_t step, float max_grad_norm, float grad_scale, float *norm_out, void *stream);

/* Data-parallel PPO across the GPUs of one node, one process per GPU: the gradient all-reduce (mean over ranks), clipping and Adam
 * as ONE kernel over NVLink peer memory (csrc/brb_policy.cu).  Every rank creates a comm object (a small symmetric block in its own
 * HBM), exports its 64-byte cudaIpc handle, the host side all-gathers the handles (torch.distributed), every rank opens its peers'.
 * Per optimiser step `step` (1-based, the same on every rank): brb_ppo_grad accumulates into brb_comm_grad(c, step) (zero on entry),
 * then brb_comm_allreduce_adam on the same stream.  Sums are formed in rank order on every rank: the replicas stay bit-identical.
 * brb_comm_fault: 1 if a (bounded) wait for a peer ever timed out.  world <= 8, n <= 10,240. */
typedef struct BrbComm BrbComm;
int brb_comm_create(int rank, int world, int device, int64_t n, BrbComm **out);
int brb_comm_export(BrbComm *c, void *handle64);
int brb_comm_open(BrbComm *c, const void *handles /* world x 64 bytes, rank order */);
void brb_comm_destroy(BrbComm *c);
float *brb_comm_grad(BrbComm *c, int64_t step);
int brb_comm_fault(BrbComm *c);
int brb_comm_allreduce_adam(BrbComm *c, float *params, float *m, float *v, float lr, float beta1, float beta2, float eps, int64_t step,
                            float max_grad_norm, float *norm_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif
