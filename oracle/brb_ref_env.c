/*
 * brb_ref_env.c — fp64 CPU ORACLE for the env task logic (test infrastructure; see brb_ref.h).
 *
 * Restates the reference's Python env classes on top of brb_ref_step, with every random draw
 * INJECTED by the caller (the reference mixes a seeded Generator with the global
 * unseeded numpy RNG, so reference runs are not reproducible; parity = same draws in, same values out).
 *
 * Draw slots.  Each step() call takes u_step[4] and u_reset[16] (uniforms in [0,1)):
 *   u_step[0]  noise in _get_reward's get_pitch()          RobotBaseEnv.py:210 via env01_v2.py:19
 *   u_step[1]  noise in the termination test               env01_v2.py:44
 *   u_step[2]  noise in _get_obs's get_pitch()             RobotBaseEnv.py:224
 *   u_step[3]  noise inside get_pitch_dot_alt()            RobotBaseEnv.py:145
 *   u_reset[0..8]   self.np_random.uniform(-0.01, 0.01, 9) env01_v1.py:40-42 / env01_v2.py:53-55
 *   u_reset[9..11]  x_rot, y_rot, z_rot draws              env01_v1.py:46-49 / env01_v2.py:59-62
 *   u_reset[12]     v2: obs pitch noise (RobotBaseEnv.py:224); v3: delay_target_speed draw (env01_v3.py:44)
 *   u_reset[13]     v2: pitch-dot noise (RobotBaseEnv.py:145);  v3: pitch_offset draw (env01_v3.py:52)
 * (v1 and v3 ignore the noise slots; v1/v2 ignore the v3 meaning of slots 12/13.)
 *
 * Env03-v2 (kind 3) uses wider rows: u_reset[32]: [0..15] np_random.uniform(-0.01, 0.01, 16) (env03_v1.py:61-63),
 * [16..18] x/y/z_rot (env03_v1.py:67-70), [19..23] the five draws of set_block_pos_vel (env03_v2.py:42-53: target x,
 * target z, block x/y/z_rot); u_step[8]: [0..4] the five draws of a re-fire inside step() (env03_v1.py:47-49), rest unused.
 * attack_side_front (env03_v2.py:22, drawn once per env instance) is set with brb_ref_env_set_attack_side.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "brb_ref.h"
#include "brb_ref_env.h"

#include <pthread.h>

#define PI 3.14159265358979323846

static const double PITCH_MAX = 0.25, PITCH_DOT_MAX = 1, WHEEL_SPEED_MAX = 170.0, WHEEL_SPEED_DELTA_MAX = 4.0,
                    YAW_MAX = 45.0; /* RobotBaseEnv.py:19-23 */

int brb_ref_sizeof_env(void) { return (int)sizeof(BrbRefEnv); }
int brb_ref_reset_stride(int kind) { return kind == BRB_ENV03_V2 ? 32 : 16; }
int brb_ref_step_stride(int kind) { return kind == BRB_ENV03_V2 ? 8 : 4; }
void brb_ref_env_set_attack_side(BrbRefEnv *e, int front) { e->attack_side_front = front; }

void brb_ref_env_init(BrbRefEnv *e, int kind, int max_episode_steps) {
  memset(e, 0, sizeof *e);
  e->kind = kind;
  e->max_episode_steps = max_episode_steps; /* balance_robot/__init__.py:12-24, :47-52 */
  e->block_delay = 0.5;                     /* env03_v2.py:23 */
  e->attack_side_front = 1;
}

/* scipy.spatial.transform.Rotation.from_quat([x, y, z, w]).as_euler('xyz') restated (third party: scipy==1.14.1 in the
 * reference's conda-environment.yaml:10, 1.18.1 in this image; same algorithm — Bernardes & Viollet, "Quaternion to Euler
 * angles conversion: a direct, general and computationally efficient method", PLoS ONE 2022, as implemented in scipy's
 * _rotation.pyx: _compute_euler_from_quat / _get_angles).  The textbook closed form atan2(2(wx+yz), 1-2(x^2+y^2)) equals
 * it only to the last ulp; the Philox uniforms are 24-bit, which makes reset observations sit on float32 rounding ties
 * often enough that the last ulp shows, so the algorithm is restated operation by operation (checked bit-for-bit against
 * scipy on 4e5 quaternions, and through the reference-class fixtures in tests/test_reference_classes.py).
 * Extrinsic 'xyz': i, j, k = 0, 1, 2, sign = +1, not symmetric.  which = 0 -> x angle (pitch), 2 -> z angle (yaw). */
static double scipy_euler_xyz(const double *q, int which) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  double n = sqrt(x * x + y * y + z * z + w * w); /* from_quat normalises, summing in its own (x, y, z, w) order */
  w /= n; x /= n; y /= n; z /= n;
  const double pi = 3.14159265358979323846;
  double a = w - y, b = x + z, c = y + w, d = z - x;
  double second = 2 * atan2(hypot(c, d), hypot(a, b));
  double half_sum = atan2(b, a), half_diff = atan2(d, c), first, third;
  if (fabs(second) <= 1e-7) { first = 2 * half_sum; third = 0; }            /* gimbal lock: third angle set to zero */
  else if (fabs(second - pi) <= 1e-7) { first = -2 * half_diff; third = 0; } /* (extrinsic order) */
  else { first = half_sum - half_diff; third = half_sum + half_diff; }
  double r = which == 0 ? first : third;                                     /* third *= sign (= +1) */
  if (r < -pi) r += 2 * pi;
  else if (r > pi) r -= 2 * pi;
  return r;
}

/* RobotBaseEnv.get_pitch (RobotBaseEnv.py:127-135): x angle of scipy as_euler('xyz') of the chassis
 * xquat, 0 when w == 0 exactly; xquat is the STALE one (Q1). */
static double true_pitch(const BrbRefEnv *e) {
  const double *q = e->d.xquat[1];
  if (q[0] == 0) return 0;
  return scipy_euler_xyz(q, 0);
}

double brb_ref_env_yaw(const BrbRefEnv *e) { /* RobotBaseEnv.py:177-184 */
  const double *q = e->d.xquat[1];
  if (q[0] == 0) return 0;
  return scipy_euler_xyz(q, 2);
}

/* get_pitch with the per-class override: v2 adds (U - 0.5) * 0.05 (env01_v2.py:16-20),
 * v3 adds pitch_offset (env01_v3.py:23-25) */
static double get_pitch(const BrbRefEnv *e, double u) {
  double p = true_pitch(e);
  if (e->kind == BRB_ENV01_V2) p += (u - 0.5) * 0.05;
  else if (e->kind == BRB_ENV01_V3) p = p + e->pitch_offset;
  return p;
}

/* RobotBaseEnv.get_pitch_dot_alt (RobotBaseEnv.py:142-157); last_time/last_pitch survive resets (Q6) */
static double get_pitch_dot_alt(BrbRefEnv *e, double u) {
  double pitch = get_pitch(e, u), ts = e->d.time, pitch_dot = 0;
  if (e->has_last) {
    double dt = ts - e->last_time;
    if (dt > 0.0) pitch_dot = (pitch - e->last_pitch) / dt;
  }
  e->last_time = ts;
  e->last_pitch = pitch;
  e->has_last = 1;
  return pitch_dot;
}

/* RobotBaseEnv._get_obs (RobotBaseEnv.py:221-246) */
static void get_obs(BrbRefEnv *e, double u_pitch, double u_dot, float *obs) {
  double pitch = get_pitch(e, u_pitch);
  double pitch_dot = get_pitch_dot_alt(e, u_dot);
  double vl = e->d.qvel[6], vr = e->d.qvel[7];
  double wheel_speed = (vl + (-1 * vr)) / 2;
  double wheel_yaw = vl - (-1 * vr);
  obs[0] = (float)(pitch / PITCH_MAX);
  obs[1] = (float)(pitch_dot / PITCH_DOT_MAX);
  obs[2] = (float)(vl / WHEEL_SPEED_MAX * 4);
  obs[3] = (float)(vr / WHEEL_SPEED_MAX * 4);
  obs[4] = (float)((e->target_wheel_speed - wheel_speed) / WHEEL_SPEED_MAX * 4);
  obs[5] = (float)((e->target_yaw - wheel_yaw) / YAW_MAX * 3);
}

/* RobotBaseEnv._get_reward (RobotBaseEnv.py:190-219) and Env01_v3._get_reward (env01_v3.py:56-96) */
static double get_reward(const BrbRefEnv *e, double u) {
  double vl = e->d.qvel[6], vr = e->d.qvel[7];
  if (e->kind == BRB_ENV01_V3) {
    double reward = 0.6;
    double pitch = get_pitch(e, u);
    double wheel_speed = (vl + (-1 * vr)) / 2;
    double tws = e->target_wheel_speed;
    double dv = tws - wheel_speed;
    reward -= fabs(pitch) * 0.05;
    double max_dv = dv < -40.0 ? -40.0 : (dv > 40.0 ? 40.0 : dv);
    double dv_s = fabs(max_dv / 40.0);
    reward -= 0.15 * dv_s;
    if (tws > 0 && tws > wheel_speed) reward += (-1.0 * pitch) * 10.0 * dv_s;
    else if (tws < 0 && tws < wheel_speed) reward += (1.0 * pitch) * 10.0 * dv_s;
    else if (tws > 0 && tws < wheel_speed) reward += (1.0 * pitch) * 10.0 * dv_s;
    else if (tws < 0 && tws > wheel_speed) reward += (-1.0 * pitch) * 10.0 * dv_s;
    double dyd = e->target_yaw - (vl - (-1 * vr));
    reward -= 0.007 * fabs(dyd);
    return reward;
  }
  double reward = 1.0;
  double average_wheel_speed = (vl * -1 + vr) / 2.0;
  double dv = 0 - average_wheel_speed;
  double dyd = 0 - e->d.qvel[5];
  reward -= 0.025 * fabs(dyd);
  double pitch = get_pitch(e, u);
  reward -= fabs(pitch);
  reward += pitch * dv * 0.5;
  return reward;
}

/* scipy Rotation.from_euler('xyz', [a, b, c]).as_quat() -> [x, y, z, w] (extrinsic: Rz(c) Ry(b) Rx(a)) */
void brb_ref_euler_xyz_to_quat_xyzw(double a, double b, double c, double *out) {
  double ca = cos(a / 2), sa = sin(a / 2), cb = cos(b / 2), sb = sin(b / 2), cc = cos(c / 2), sc = sin(c / 2);
  out[0] = sa * cb * cc - ca * sb * sc;
  out[1] = ca * sb * cc + sa * cb * sc;
  out[2] = ca * cb * sc - sa * sb * cc;
  out[3] = ca * cb * cc + sa * sb * sc;
}

/* Env03_v2.set_block_pos_vel (env03_v2.py:25-59): fire the block at the robot from 0.3 m in front / behind.
 * Reads the STALE robot xpos / xquat (Q1); u[0..4] = target x, target z, block x/y/z_rot. */
static void set_block_pos_vel(BrbRefEnv *e, const double *u) {
  const double *robot_pos = e->d.xpos[1];
  double block_attack_angle = -brb_ref_env_yaw(e);
  if (!e->attack_side_front) block_attack_angle += PI;
  double block_pos[3] = {0.3 * sin(block_attack_angle) + robot_pos[0], 0.3 * cos(block_attack_angle) + robot_pos[1], 0.15};
  double target[3] = {(u[0] - 0.5) * 0.02 + robot_pos[0], 0 + robot_pos[1], u[1] * 0.025 + 0.13};
  double v[3] = {target[0] - block_pos[0], target[1] - block_pos[1], target[2] - block_pos[2]};
  double nrm = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  for (int k = 0; k < 3; k++) v[k] = 7.5 * (v[k] / nrm);
  double x_rot = u[2] * 2 * PI, y_rot = u[3] * 2 * PI, z_rot = u[4] * 2 * PI;
  for (int k = 0; k < 3; k++) e->d.qpos[9 + k] = block_pos[k];
  brb_ref_euler_xyz_to_quat_xyzw(x_rot, y_rot, z_rot, e->d.qpos + 12); /* same scalar-last-into-scalar-first quirk */
  for (int k = 0; k < 3; k++) e->d.qvel[8 + k] = v[k];
}

/* MujocoEnv.reset -> mj_resetData -> reset_model (env01_v1.py:39-58, env01_v2.py:52-71, env01_v3.py:39-54) */
void brb_ref_env_reset(const BrbRefModel *m, BrbRefEnv *e, const double *u_reset, float *obs) {
  brb_ref_reset_data(m, &e->d);
  e->elapsed_steps = 0;
  if (e->kind == BRB_ENV01_V3) {
    e->target_wheel_speed = 0;
    e->target_yaw = 0;
    double s = -10.0 + (10.0 - -10.0) * u_reset[12];
    if (s > 0) s += 10; else s -= 10;
    e->delay_target_speed = s;
    e->pitch_offset = -0.0349066 + (0.0349066 - -0.0349066) * u_reset[13];
  }
  double qpos[BRB_MAXNQ];
  for (int i = 0; i < m->nq; i++) qpos[i] = m->qpos0[i] + (-0.01 + (0.01 - -0.01) * u_reset[i]);
  qpos[2] = 0;
  double x_rot, y_rot, z_rot;
  if (e->kind == BRB_ENV03_V2) { /* env03_v1.py:60-83 with Env03_v2.set_block_pos_vel */
    x_rot = (u_reset[16] - 0.5) * 2 * PI;
    y_rot = (u_reset[17] - 0.5) * 0.4;
    z_rot = (u_reset[18] - 0.5) * 0.4;
    brb_ref_euler_xyz_to_quat_xyzw(x_rot, y_rot, z_rot, qpos + 3);
    memcpy(e->d.qpos, qpos, sizeof(double) * m->nq);
    memset(e->d.qvel, 0, sizeof e->d.qvel);
    brb_ref_forward(m, &e->d);
    set_block_pos_vel(e, u_reset + 19);
    e->has_block_timer = 0;
    get_obs(e, 0, 0, obs);
    return;
  }
  x_rot = (u_reset[9] - 0.5) * 2 * PI;
  if (e->kind == BRB_ENV01_V2) {
    y_rot = (u_reset[10] - 0.5) * 0.2;
    z_rot = (u_reset[11] - 0.5) * 2.0;
  } else {
    y_rot = (u_reset[10] - 0.5) * 0.4;
    z_rot = (u_reset[11] - 0.5) * 0.4;
  }
  /* Q3: scalar-LAST quaternion written verbatim into MuJoCo's scalar-FIRST slots */
  brb_ref_euler_xyz_to_quat_xyzw(x_rot, y_rot, z_rot, qpos + 3);
  memcpy(e->d.qpos, qpos, sizeof(double) * m->nq);
  memset(e->d.qvel, 0, sizeof e->d.qvel);
  brb_ref_forward(m, &e->d); /* set_state -> mj_forward: fresh xquat, warm start */
  get_obs(e, u_reset[12], u_reset[13], obs);
}

/* Env01.step / Env01_v2.step / Env01_v3.step (env01_v1.py:15-37, env01_v2.py:28-50, env01_v3.py:27-37)
 * plus gymnasium TimeLimit (truncated when elapsed >= max_episode_steps). */
void brb_ref_env_step(const BrbRefModel *m, BrbRefEnv *e, const float *action, const double *u_step, float *obs,
                      double *reward, int *terminated, int *truncated) {
  if (e->kind == BRB_ENV01_V3) {
    double t = e->d.time;
    if (t > 5.5) e->target_wheel_speed = 3.0 * e->delay_target_speed;
    else if (t > 4.5) e->target_wheel_speed = 2.0 * e->delay_target_speed;
    else if (t > 3.0) e->target_wheel_speed = -1.0 * e->delay_target_speed;
    else if (t > 1.0) e->target_wheel_speed = e->delay_target_speed;
  }
  *reward = get_reward(e, u_step[0]);
  e->d.ctrl[0] = e->d.qvel[6] + (double)action[0] * WHEEL_SPEED_DELTA_MAX;
  e->d.ctrl[1] = e->d.qvel[7] + (double)action[1] * WHEEL_SPEED_DELTA_MAX;
  brb_ref_step(m, &e->d, 250);
  if (e->kind == BRB_ENV03_V2) { /* env03_v1.py:39-49 */
    const double *bv = e->d.qvel + 8;
    if (sqrt(bv[0] * bv[0] + bv[1] * bv[1] + bv[2] * bv[2]) < 0.1 && !e->has_block_timer) {
      e->d.qpos[9] = 10; e->d.qpos[10] = 10; e->d.qpos[11] = 0; /* remove_block, env03_v1.py:85-86 */
      e->has_block_timer = 1;
      e->block_delay_time_start = e->d.time;
    }
    if (e->has_block_timer && (e->d.time - e->block_delay_time_start) > e->block_delay) {
      set_block_pos_vel(e, u_step);
      e->has_block_timer = 0;
    }
  }
  *terminated = fabs(get_pitch(e, u_step[1])) > (50 * PI / 180);
  get_obs(e, u_step[2], u_step[3], obs);
  e->elapsed_steps++;
  *truncated = e->elapsed_steps >= e->max_episode_steps;
}

/* ------------------------------------------------------------------ vectorised front end (DummyVecEnv + Monitor
 * semantics: auto-reset, terminal observation, episode return / length), threads over envs */
struct BrbRefVec {
  BrbRefModel model;
  int n, kind;
  BrbRefEnv *envs;
  double *ep_return;
  int *ep_len;
};

int brb_ref_vec_create(const BrbRefModel *m, int kind, int max_episode_steps, int n, BrbRefVec **out) {
  BrbRefVec *v = (BrbRefVec *)calloc(1, sizeof *v);
  if (!v) return -12;
  v->model = *m;
  v->n = n;
  v->kind = kind;
  v->envs = (BrbRefEnv *)calloc((size_t)n, sizeof(BrbRefEnv));
  v->ep_return = (double *)calloc((size_t)n, sizeof(double));
  v->ep_len = (int *)calloc((size_t)n, sizeof(int));
  if (!v->envs || !v->ep_return || !v->ep_len) return -12;
  for (int i = 0; i < n; i++) brb_ref_env_init(&v->envs[i], kind, max_episode_steps);
  *out = v;
  return 0;
}

void brb_ref_vec_destroy(BrbRefVec *v) {
  if (!v) return;
  free(v->envs); free(v->ep_return); free(v->ep_len); free(v);
}

BrbRefEnv *brb_ref_vec_env(BrbRefVec *v, int i) { return &v->envs[i]; }

/* one contiguous block of envs per thread (plain pthreads; no OpenMP runtime needed) */
typedef struct {
  BrbRefVec *v;
  int lo, hi, is_reset;
  const float *actions;
  const double *u_step, *u_reset;
  float *obs, *reward, *terminal_obs, *ep_return;
  uint8_t *done, *truncated;
  int32_t *ep_len;
} VecJob;

static void *vec_worker(void *arg) {
  VecJob *j = (VecJob *)arg;
  BrbRefVec *v = j->v;
  for (int i = j->lo; i < j->hi; i++) {
    if (j->is_reset) {
      brb_ref_env_reset(&v->model, &v->envs[i], j->u_reset + brb_ref_reset_stride(v->kind) * (size_t)i, j->obs + 6 * (size_t)i);
      v->ep_return[i] = 0;
      v->ep_len[i] = 0;
      continue;
    }
    double r;
    int term, trunc;
    float o[6];
    brb_ref_env_step(&v->model, &v->envs[i], j->actions + 2 * (size_t)i, j->u_step + brb_ref_step_stride(v->kind) * (size_t)i, o, &r, &term, &trunc);
    v->ep_return[i] += r;
    v->ep_len[i] += 1;
    j->reward[i] = (float)r;
    j->done[i] = (uint8_t)(term || trunc);
    j->truncated[i] = (uint8_t)(trunc && !term);
    if (j->ep_return) j->ep_return[i] = (float)v->ep_return[i];
    if (j->ep_len) j->ep_len[i] = v->ep_len[i];
    if (j->done[i]) {
      if (j->terminal_obs) memcpy(j->terminal_obs + 6 * (size_t)i, o, sizeof o);
      brb_ref_env_reset(&v->model, &v->envs[i], j->u_reset + brb_ref_reset_stride(v->kind) * (size_t)i, o);
      v->ep_return[i] = 0;
      v->ep_len[i] = 0;
    }
    memcpy(j->obs + 6 * (size_t)i, o, sizeof o);
  }
  return 0;
}

static void vec_run(VecJob *proto, int nthreads) {
  int n = proto->v->n;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n) nthreads = n > 0 ? n : 1;
  if (nthreads > 256) nthreads = 256;
  VecJob jobs[256];
  pthread_t th[256];
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = *proto;
    jobs[t].lo = (int)((long long)n * t / nthreads);
    jobs[t].hi = (int)((long long)n * (t + 1) / nthreads);
  }
  for (int t = 1; t < nthreads; t++) pthread_create(&th[t], 0, vec_worker, &jobs[t]);
  vec_worker(&jobs[0]);
  for (int t = 1; t < nthreads; t++) pthread_join(th[t], 0);
}

void brb_ref_vec_reset(BrbRefVec *v, const double *u_reset /*[n,16]*/, float *obs /*[n,6]*/, int nthreads) {
  VecJob j;
  memset(&j, 0, sizeof j);
  j.v = v; j.is_reset = 1; j.u_reset = u_reset; j.obs = obs;
  vec_run(&j, nthreads);
}

void brb_ref_vec_step(BrbRefVec *v, const float *actions /*[n,2]*/, const double *u_step /*[n,4]*/,
                      const double *u_reset /*[n,16]*/, float *obs, float *reward, uint8_t *done, uint8_t *truncated,
                      float *terminal_obs, float *ep_return, int32_t *ep_len, int nthreads) {
  VecJob j;
  memset(&j, 0, sizeof j);
  j.v = v; j.actions = actions; j.u_step = u_step; j.u_reset = u_reset; j.obs = obs; j.reward = reward;
  j.done = done; j.truncated = truncated; j.terminal_obs = terminal_obs; j.ep_return = ep_return; j.ep_len = ep_len;
  vec_run(&j, nthreads);
}

/* ------------------------------------------------------------------ Philox4x32-10 (Salmon et al., SC'11) — the
 * counter-based stream the CUDA path draws from; restated here so the oracle can replay the same draws.
 * counter = (env_lo, env_hi, event, block), key = (seed_lo, seed_hi); uniform = (word >> 8) * 2^-24. */
static inline void philox_round(uint32_t *c, const uint32_t *k) {
  uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
  uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1;
  uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
void brb_ref_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k[2] = {key[0], key[1]};
  for (int r = 0; r < 10; r++) {
    philox_round(c, k);
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  memcpy(out, c, sizeof c);
}
/* raw block access: out[n, 4*nblocks] = uniforms of blocks first_block .. first_block+nblocks-1 of `event` */
void brb_ref_philox_blocks(uint64_t seed, uint64_t env0, int n, uint32_t event, uint32_t first_block, int nblocks, double *out) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (int i = 0; i < n; i++) {
    uint64_t env = env0 + (uint64_t)i;
    for (int b = 0; b < nblocks; b++) {
      uint32_t ctr[4] = {(uint32_t)env, (uint32_t)(env >> 32), event, first_block + (uint32_t)b}, w[4];
      brb_ref_philox4x32_10(ctr, key, w);
      for (int k = 0; k < 4; k++) out[((size_t)i * nblocks + b) * 4 + k] = (double)(w[k] >> 8) * (1.0 / 16777216.0);
    }
  }
}
/* fills u_step[n,4] (block 0) and u_reset[n,16] (blocks 1..4) for event index `event` */
void brb_ref_philox_draws(uint64_t seed, uint64_t env0, int n, uint32_t event, double *u_step, double *u_reset) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (int i = 0; i < n; i++) {
    uint64_t env = env0 + (uint64_t)i;
    for (uint32_t b = 0; b < 5; b++) {
      uint32_t ctr[4] = {(uint32_t)env, (uint32_t)(env >> 32), event, b}, w[4];
      brb_ref_philox4x32_10(ctr, key, w);
      double *dst = (b == 0) ? (u_step ? u_step + 4 * (size_t)i : 0) : (u_reset ? u_reset + 16 * (size_t)i + 4 * (b - 1) : 0);
      if (dst)
        for (int k = 0; k < 4; k++) dst[k] = (double)(w[k] >> 8) * (1.0 / 16777216.0);
    }
  }
}
