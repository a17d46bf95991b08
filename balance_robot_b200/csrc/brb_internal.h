// brb_internal.h — device-side state layout shared by the kernels and the C-ABI translation unit.
#ifndef BRB_INTERNAL_H
#define BRB_INTERNAL_H
#include <stdint.h>

#ifndef BRB_BLOCK
#define BRB_BLOCK 128   // threads per CTA of the step kernel (one env per thread): one warp per SM sub-partition
#endif
#ifndef BRB_BLOCK_ENV03
#define BRB_BLOCK_ENV03 128   // Env03-v2: the 4 warps of a CTA walk the substep loop in lockstep (shared instruction cache); 256: 15.6 vs 14.1 ms
#endif
#ifndef BRB_MINBLOCKS_ENV03
#define BRB_MINBLOCKS_ENV03 2   // Env03-v2 carries the block and the coupled system: 255 registers, 2 CTAs per SM
#endif
#ifndef BRB_MINBLOCKS
#define BRB_MINBLOCKS 3 // __launch_bounds__ min resident CTAs per SM (register cap = 65536 / (BRB_BLOCK * BRB_MINBLOCKS))
#endif
#define BRB_ENV03_V2_WB 4   // internal kernel instance: Env03-v2 with the opt-in wheel-block path compiled in (BRB_FLAG_WHEEL_BLOCK selects it at
                            // launch, so the default instance carries none of it: 10.8 vs 11.4 ms per step with the path compiled in but off)
#define BRB_IS_ENV03(KIND) ((KIND) == BRB_ENV03_V2 || (KIND) == BRB_ENV03_V2_WB)
#define BRB_MAXIT 8     // cap on active-set (Newton) iterations per substep

// Struct-of-arrays env state in HBM: column k of a [K][N] array lives at base + k*N, so a warp's 32
// envs read 32 consecutive values of every column (fully coalesced 256 B / 128 B segments).
struct BrbState {
  long long n;            // envs in this shard
  int nq, nv;             // 9, 8 (Env01-*) or 16, 14 (Env03-v2: + free block)
  long long env0;         // global id of env 0 (Philox counter)
  unsigned long long seed;
  double *qpos;           // [nq][N] x y z qw qx qy qz thL thR [block x y z qw qx qy qz]   (MuJoCo data.qpos)
  double *qvel;           // [nv][N] v_world(3) w_body(3) sL sR [block v_world(3) w_body(3)] (MuJoCo data.qvel)
  double *xquat;          // [4][N]  chassis quaternion as the task logic sees it (one substep stale, Q1)
  uint32_t *aset;         // [N]     converged contact-row active set of the last substep (the solver's warm start)
  double *last_pitch;     // [N]     RobotBaseEnv.last_pitch
  double *ep_return;      // [N]     Monitor running return
  double *v3;             // [3][N]  Env01-v3: target_wheel_speed, delay_target_speed, pitch_offset;
                          //         Env03-v2: block timer start (<0 = none), attack_side_front, block active-set bits
  int *elapsed;           // [N]     TimeLimit._elapsed_steps (also indexes the time table)
  int *ep_len;            // [N]
  uint32_t *event;        // [N]     Philox event counter (0 = reset_all, k = k-th step call)
  const double *time_table;  // [max_episode_steps + 2] fp64 data.time after k env steps
  unsigned long long *stats; // [BRB_NSTATS]
};

// Visit order of the step kernel: `in` = order for this launch (NULL = identity).  The launch publishes a group key per
// env (0 = far airborne, then by landing time, then by wheel-rim contact pattern) plus a histogram; brb_group_kernel counting-sorts them into the
// next launch's order, so that the lanes of a warp mostly run the same contact path.
#define BRB_NLAND 8        // landing-time bins of airborne robots that will touch down during the next step
#define BRB_NCOUPLED 5     // Env03-v2: robots whose block lies on the floor, then those whose block touches the chassis by
                           // number of chassis-block contacts (1-2, 3-4, 5-8), then (highest key, visited first) those with a wheel within reach of the block
#define BRB_NGROUPS (1 + BRB_NLAND + 16 + BRB_NCOUPLED)   // 0 = stays airborne, 1..NLAND = landing, then NLAND + rank of the
                                                          // contact-slot pattern (1..15), then the coupled buckets (<= 31 keys in all)
struct BrbPerm {
  const int *in;
  uint8_t *key_out;    // [N]
  unsigned *hist;      // [32] accumulated by the step kernel, zero on entry
  unsigned *cursor;    // work-queue cursor (zero on entry; NULL = one robot per thread, fixed grid): warps pull 32 robots at a time
};
#endif
