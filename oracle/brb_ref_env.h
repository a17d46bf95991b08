/* brb_ref_env.h — env-level fp64 CPU ORACLE (test infrastructure only; see brb_ref.h). */
#ifndef BRB_REF_ENV_H
#define BRB_REF_ENV_H
#include <stdint.h>
#include "brb_ref.h"
#ifdef __cplusplus
extern "C" {
#endif

#define BRB_ENV01_V1 0
#define BRB_ENV01_V2 1
#define BRB_ENV01_V3 2
#define BRB_ENV03_V2 3

typedef struct BrbRefEnv {
  int kind, max_episode_steps, elapsed_steps, has_last;
  double last_time, last_pitch;                 /* RobotBaseEnv.py:68-69 (never cleared on reset, Q6) */
  double target_wheel_speed, target_yaw;        /* RobotBaseEnv.py:71-72 */
  double delay_target_speed, pitch_offset;      /* env01_v3.py:18-21 */
  int has_block_timer, attack_side_front;       /* env03_v1.py:22, env03_v2.py:22 */
  double block_delay_time_start, block_delay;   /* env03_v1.py:22-24, env03_v2.py:23 */
  BrbRefData d;
} BrbRefEnv;

typedef struct BrbRefVec BrbRefVec;

int brb_ref_sizeof_env(void);
void brb_ref_env_init(BrbRefEnv *e, int kind, int max_episode_steps);
void brb_ref_env_set_attack_side(BrbRefEnv *e, int front);
int brb_ref_reset_stride(int kind);
int brb_ref_step_stride(int kind);
void brb_ref_env_reset(const BrbRefModel *m, BrbRefEnv *e, const double *u_reset, float *obs);
void brb_ref_env_step(const BrbRefModel *m, BrbRefEnv *e, const float *action, const double *u_step, float *obs,
                      double *reward, int *terminated, int *truncated);
double brb_ref_env_yaw(const BrbRefEnv *e);
void brb_ref_euler_xyz_to_quat_xyzw(double a, double b, double c, double *out);

int brb_ref_vec_create(const BrbRefModel *m, int kind, int max_episode_steps, int n, BrbRefVec **out);
void brb_ref_vec_destroy(BrbRefVec *v);
BrbRefEnv *brb_ref_vec_env(BrbRefVec *v, int i);
void brb_ref_vec_reset(BrbRefVec *v, const double *u_reset, float *obs, int nthreads);
void brb_ref_vec_step(BrbRefVec *v, const float *actions, const double *u_step, const double *u_reset, float *obs,
                      float *reward, uint8_t *done, uint8_t *truncated, float *terminal_obs, float *ep_return,
                      int32_t *ep_len, int nthreads);

void brb_ref_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void brb_ref_philox_blocks(uint64_t seed, uint64_t env0, int n, uint32_t event, uint32_t first_block, int nblocks, double *out);
void brb_ref_philox_draws(uint64_t seed, uint64_t env0, int n, uint32_t event, double *u_step, double *u_reset);
#ifdef __cplusplus
}
#endif
#endif
