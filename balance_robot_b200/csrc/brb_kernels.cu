// brb_kernels.cu — fused balance-robot env step for sm_100a (B200).
//
// One launch = one VecEnv.step for N robots: reward (pre-step state), 250 substeps of forward dynamics
// + wheel/floor contact solve + implicitfast integration, termination, observation (+ v2 Philox noise),
// TimeLimit truncation, Monitor statistics and auto-reset.  Replaces, for the model of
// envs/env01_v1.xml + envs/robot-02.xml, what the reference gets from
//   Env01.step / Env01_v2.step / Env01_v3.step   (envs/env01_v1.py:15-37, env01_v2.py:28-50, env01_v3.py:27-37)
//   -> mujoco.mj_step(model, data, nstep=250)     (envs/env01_v1.py:24)  [third party, restated: SURVEY.md App. A]
//   + gymnasium TimeLimit / SB3 DummyVecEnv auto-reset / Monitor around it (sb_rl.py:500-501).
//
// Formulation: dynamics are written in the CHASSIS frame, where the joint-space inertia
// M_b is a constant sparse 8x8 (both wheels are axisymmetric about their hinges).  Per substep:
//   f   = -bias(w, s, n_b) + servo/damping torques                         (closed form, gyrostat)
//   H a = f - sum_c P_c' S_c yhat_c,  H = M_b + sum_c P_c' S_c P_c          (primal Newton on the active set)
//   a+  = a - Y(clamp state) [a_sL; a_sR]                                   (implicitfast via Woodbury)
//   qvel += h a+ ; qpos += h qvel (quaternion: q += q (x) (dq - 1))         (semi-implicit, compensated sums)
// Work is FP32; the 17 state accumulators are Kahan-compensated float pairs because h = 2e-5 makes every
// increment ~1e-5 of the value.  Task logic (reward/obs/termination/reset) runs in fp64
// once per env step so it matches the oracle bit-for-bit at the f32 outputs.
//
// No CPU fallback, no Triton, no multi-backend dispatch: this file is the product path.

// BRB_HOST_EMU (tests/host_emu only): the per-env device functions are compiled as plain C++ so the CPU test
// suite can check the exact kernel arithmetic against the oracle without a GPU.  The product library is
// never built with it and contains no host compute path.
#ifdef BRB_HOST_EMU
#include "emu_shim.h"
#define BRB_D static inline
#define BRB_NOINLINE static __attribute__((noinline))
#define __popc(x) __builtin_popcount(x)
#else
#include <cuda_runtime.h>
#define BRB_D __device__ __forceinline__
#define BRB_NOINLINE __device__ __noinline__
#define BRB_CTA_OR(p) (__syncthreads_or(p) != 0)
#define BRB_CTA_SYNC() __syncthreads()
#endif
#include <math.h>
#include <stdint.h>

#include "../../include/brb.h"
#include "brb_internal.h"

#define BRB_PI 3.14159265358979323846

// ---------------------------------------------------------------------------------------------------
// compensated accumulator: value = s - c
struct KF { float s, c; };
BRB_D void kadd(KF &x, float d) {
  float y = d - x.c;
  float t = x.s + y;
  x.c = (t - x.s) - y;
  x.s = t;
}
BRB_D KF ksplit(double v) {
  KF r;
  r.s = (float)v;
  r.c = -(float)(v - (double)r.s);
  return r;
}
BRB_D double kjoin(const KF &x) { return (double)x.s - (double)x.c; }

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11).  counter = (env_lo, env_hi, event, block), key = (seed_lo, seed_hi)
BRB_D void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint64_t m0 = (uint64_t)0xD2511F53u * c0, m1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t hi0 = (uint32_t)(m0 >> 32), lo0 = (uint32_t)m0, hi1 = (uint32_t)(m1 >> 32), lo1 = (uint32_t)m1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
BRB_D void draw4(uint64_t seed, uint64_t env, uint32_t event, uint32_t block, double u[4]) {
  uint32_t w[4];
  philox4x32_10((uint32_t)env, (uint32_t)(env >> 32), event, block, (uint32_t)seed, (uint32_t)(seed >> 32), w);
#pragma unroll
  for (int k = 0; k < 4; k++) u[k] = (double)(w[k] >> 8) * (1.0 / 16777216.0);
}

// ---------------------------------------------------------------------------------------------------
// task-logic helpers (fp64; restate the reference Python, the oracle restates the same logic independently)
// scipy Rotation.from_quat([x,y,z,w]).as_euler('xyz') [third party], angle `WHICH` (0 = x = pitch, 2 = z = yaw): the
// half-angle algorithm scipy implements (Bernardes & Viollet 2022; _rotation.pyx _get_angles), operation by operation,
// because the textbook atan2(2(wx+yz), 1-2(x^2+y^2)) equals it only to the last ulp and the 24-bit Philox uniforms put reset
// observations on float32 rounding ties often enough for that ulp to show.  Gimbal lock (roll = +-90 deg within 1e-7) is
// tested on the tangents instead of scipy's 2 atan2(hypot, hypot): same set up to its boundary.
template <int WHICH>
BRB_D double euler_xyz_of(const double q[4]) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  const double n = sqrt(x * x + y * y + z * z + w * w);
  w /= n; x /= n; y /= n; z /= n;
  const double a = w - y, b = x + z, c = y + w, d = z - x;
  const double half_sum = atan2(b, a), half_diff = atan2(d, c);
  const double sab = a * a + b * b, scd = c * c + d * d, t2 = 2.5e-15;   // tan(0.5e-7)^2
  double r;
  if (scd <= t2 * sab) r = WHICH == 0 ? 2 * half_sum : 0.0;
  else if (sab <= t2 * scd) r = WHICH == 0 ? -2 * half_diff : 0.0;
  else r = WHICH == 0 ? half_sum - half_diff : half_sum + half_diff;
  if (r < -BRB_PI) r += 2 * BRB_PI;
  else if (r > BRB_PI) r -= 2 * BRB_PI;
  return r;
}
BRB_D double pitch_of(const double q[4]) {  // RobotBaseEnv.py:127-135
  if (q[0] == 0.0) return 0.0;
#ifdef BRB_EXP_OLD_EULER
  { double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]); double w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
    return atan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y)); }
#endif
  return euler_xyz_of<0>(q);
}

template <int KIND>
BRB_D double noisy_pitch(double p, double u, double pitch_offset) {
  if (KIND == BRB_ENV01_V2) p += (u - 0.5) * 0.05;          // env01_v2.py:16-20
  else if (KIND == BRB_ENV01_V3) p = p + pitch_offset;      // env01_v3.py:23-25
  return p;
}

template <int KIND>
BRB_D double reward_of(double pitch, double vl, double vr, double yaw_dot, double tws, double tyaw) {
  if (KIND == BRB_ENV01_V3) {  // env01_v3.py:56-96
    double reward = 0.6;
    double wheel_speed = (vl + (-1 * vr)) / 2;
    double dv = tws - wheel_speed;
    reward -= fabs(pitch) * 0.05;
    double max_dv = dv < -40.0 ? -40.0 : (dv > 40.0 ? 40.0 : dv);
    double dv_s = fabs(max_dv / 40.0);
    reward -= 0.15 * dv_s;
    if (tws > 0 && tws > wheel_speed) reward += (-1.0 * pitch) * 10.0 * dv_s;
    else if (tws < 0 && tws < wheel_speed) reward += (1.0 * pitch) * 10.0 * dv_s;
    else if (tws > 0 && tws < wheel_speed) reward += (1.0 * pitch) * 10.0 * dv_s;
    else if (tws < 0 && tws > wheel_speed) reward += (-1.0 * pitch) * 10.0 * dv_s;
    double dyd = tyaw - (vl - (-1 * vr));
    reward -= 0.007 * fabs(dyd);
    return reward;
  }
  double reward = 1.0;  // RobotBaseEnv.py:190-219
  double average_wheel_speed = (vl * -1 + vr) / 2.0;
  double dv = 0 - average_wheel_speed;
  double dyd = 0 - yaw_dot;
  reward -= 0.025 * fabs(dyd);
  reward -= fabs(pitch);
  reward += pitch * dv * 0.5;
  return reward;
}

BRB_D void obs_of(double pitch, double pitch_dot, double vl, double vr, double tws, double tyaw, float o[6]) {
  double wheel_speed = (vl + (-1 * vr)) / 2;   // RobotBaseEnv.py:221-246
  double wheel_yaw = vl - (-1 * vr);
  o[0] = (float)(pitch / 0.25);
  o[1] = (float)(pitch_dot / 1);
  o[2] = (float)(vl / 170.0 * 4);
  o[3] = (float)(vr / 170.0 * 4);
  o[4] = (float)((tws - wheel_speed) / 170.0 * 4);
  o[5] = (float)((tyaw - wheel_yaw) / 45.0 * 3);
}

// reset_model (env01_v1.py:39-58, env01_v2.py:52-71, env01_v3.py:39-54) + first observation.
// u[0..15]: draw-slot layout: DESIGN.md §4.
template <int KIND>
BRB_D void reset_env(const BrbState &S, long long i, const double u[16], float o[6]) {
  const long long N = S.n;
  double tws = 0, dts = 0, poff = 0;
  if (KIND == BRB_ENV01_V3) {
    double s = -10.0 + (10.0 - -10.0) * u[12];
    if (s > 0) s += 10; else s -= 10;
    dts = s;
    poff = -0.0349066 + (0.0349066 - -0.0349066) * u[13];
    S.v3[0 * N + i] = 0.0;
    S.v3[1 * N + i] = dts;
    S.v3[2 * N + i] = poff;
  }
  double qpos[9];
  const double qpos0[9] = {0, 0, 0, 1, 0, 0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 9; k++) qpos[k] = qpos0[k] + (-0.01 + (0.01 - -0.01) * u[k]);
  qpos[2] = 0;
  double x_rot = (u[9] - 0.5) * 2 * BRB_PI, y_rot, z_rot;
  if (KIND == BRB_ENV01_V2) { y_rot = (u[10] - 0.5) * 0.2; z_rot = (u[11] - 0.5) * 2.0; }
  else { y_rot = (u[10] - 0.5) * 0.4; z_rot = (u[11] - 0.5) * 0.4; }
  double sa, ca, sb, cb, sc, cc;
  sincos(x_rot / 2, &sa, &ca); sincos(y_rot / 2, &sb, &cb); sincos(z_rot / 2, &sc, &cc);
  // Q3: scipy's scalar-last [x,y,z,w] lands in MuJoCo's scalar-first slots
  qpos[3] = sa * cb * cc - ca * sb * sc;
  qpos[4] = ca * sb * cc + sa * cb * sc;
  qpos[5] = ca * cb * sc - sa * sb * cc;
  qpos[6] = ca * cb * cc + sa * sb * sc;
#pragma unroll
  for (int k = 0; k < 9; k++) S.qpos[k * N + i] = qpos[k];
#pragma unroll
  for (int k = 0; k < 8; k++) S.qvel[k * N + i] = 0.0;
  S.aset[i] = 0xFFFFu;
  double xq[4];
  double n = sqrt(qpos[3] * qpos[3] + qpos[4] * qpos[4] + qpos[5] * qpos[5] + qpos[6] * qpos[6]);
#pragma unroll
  for (int k = 0; k < 4; k++) { xq[k] = qpos[3 + k] / n; S.xquat[k * N + i] = xq[k]; }
  S.elapsed[i] = 0;
  S.ep_return[i] = 0.0;
  S.ep_len[i] = 0;
  double p = pitch_of(xq);
  double pitch = noisy_pitch<KIND>(p, u[12], poff);
  double pitch2 = noisy_pitch<KIND>(p, u[13], poff);
  S.last_pitch[i] = pitch2;     // Q6: dt <= 0 on the reset observation -> pitch_dot = 0
  obs_of(pitch, 0.0, 0.0, 0.0, tws, 0.0, o);
}

// ---------------------------------------------------------------------------------------------------
// Physics.  Everything below works on registers of one thread (= one robot).  Contact slots are addressed
// with compile-time indices only (template<int CI>), so the 8x8 system and the contact records stay in
// registers (a run-time contact index sends H[] to local memory — measured in profiles/r1a).
#define LT(i, j) ((i) * ((i) + 1) / 2 + (j))   // packed lower-triangular index

#ifdef BRB_TIMELINE
__device__ unsigned long long g_timeline[1 + 4 * 65536];   // debug: per warp-task records of the step kernel
__device__ unsigned g_tl_trips[65536 * 2];                 // debug: loop trips of the warp task that starts at visit-order position 32 k; sum of its lanes' solves
#endif
#ifdef BRB_TRIPSTATS
__device__ unsigned long long g_trip[48 + 32 * 4 + 16 + 2];  // debug: warp-trips, lane-trips, warp-trips with a solve, lanes solving; [8+k]: trips with k lanes in contact; [48+4*key+cls]: robots by incoming group key and contact class of the step
#endif

// smooth force in the solver's coordinates: world linear (gravity = -m g e_z exactly), world angular, wheels
BRB_D void phys_world_force(const BrbModelConsts &c, const struct Phys &P, float (&r)[8]);

// Coordinates of the 8x8 contact system: world-frame linear acceleration (MuJoCo's own dofs 0-2), WORLD-frame angular
// acceleration alpha_w = R alpha_b, wheel accelerations.  In these coordinates the contact frame of the z-up floor
// (n = z, t1 = y, t2 = -x) is a signed permutation, so S_c = Pi' W_c Pi costs nothing and the point map
// P_c = [I | -[r_w]x | w_w] is sparse; the price is M' = blockdiag(R,R,I) M_b blockdiag(R,R,I)' (24 flops per substep).
struct Phys {
  // ---- state carried across substeps
  KF p[3], q[4], th[2], v[3], w[3], s[2];   // world pos, quat (w,x,y,z), wheel angles, world lin vel, body ang vel, wheel speeds
  float uhi[2], ulo[2];                     // clamped ctrl targets, split hi/lo
  unsigned bits;                            // converged pyramid-row active set of the last substep (4 bits per contact slot)
  // ---- working set of the current substep (phys_setup() .. phys_finalize())
  float ex[3], ey[3], ez[3];                // columns of R = chassis x/y/z axes in the world frame
  float fb[8];                              // smooth force in the chassis frame, WITHOUT gravity in the linear part (added per use)
  float cr[4][3], cw[4][3], cy[4][3];       // per contact: r_w (from the chassis origin), wheel column w_w, yhat (n, t1, t2)
  float cD[4];                              // per-contact row weight D (only with position-dependent impedance: Env03-v2)
  unsigned valid, valid_prev;               // bit ci: contact slot ci (2*wheel + rim end) is in contact (valid_prev: at step entry)
  bool clampL, clampR;                      // servo sits on its forcerange (A.9)
  unsigned n_contact, n_solve, n_nonconv, n_slots;
#ifdef BRB_TIMELINE
  unsigned n_trips;
#endif
};

BRB_D void phys_world_force(const BrbModelConsts &c, const Phys &P, float (&r)[8]) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    r[k] = P.ex[k] * P.fb[0] + P.ey[k] * P.fb[1] + P.ez[k] * P.fb[2];
    r[3 + k] = P.ex[k] * P.fb[3] + P.ey[k] * P.fb[4] + P.ez[k] * P.fb[5];
  }
  r[2] -= c.mass * c.grav;
  r[6] = P.fb[6]; r[7] = P.fb[7];
}

// ---- A.3 steps 2-7: kinematics, smooth forces, collision, reference accelerations
// impedance imp(dist) of a dynamic pair (solimp midpoint 0.5, power 2).  pp = {mu, K, B, D1, d0, d1, width, margin}
BRB_D float imp_of(const float *pp, float dist) {
  const float x = __fdividef(fabsf(dist - pp[7]), pp[6]);
  if (x >= 1.f) return pp[5];
  const float y = (x <= 0.5f) ? 2.f * x * x : 1.f - 2.f * (1.f - x) * (1.f - x);
  return pp[4] + y * (pp[5] - pp[4]);
}

// ---- the contact on/off predicate decided like fp64 (A.6: a rim point is in contact iff dist < 0 exactly).
// A sliding contact carries O(10 N) of damping force from its first substep, so a touch-down / lift-off that fires one
// substep early or late moves a velocity by ~1e-3.  The fp32 distances are good to ~1e-8 m at worst (R is built from the
// high halves of the quaternion: measured rms 5e-10, max 4e-9 for the pitch term, 7e-10 / 5e-9 for the roll term);
// whenever one of them lies within BRB_DIST_BAND of zero, the four rim distances are re-evaluated here in fp64 from the full
// compensated state and fp64 geometry constants (hi + lo floats).  0.04 % of robot-substeps take this branch (a rim end
// hovering near the floor does it for many substeps in a row, not only at touch-down; a warp pays for any of its 32 lanes:
// measured +1.4 % step time at this band, +4.5 % at 2e-7), so it is kept to ~40 DFMA: no
// fp64 sqrt or division — rho = sqrt(n1^2 + n2^2) is the fp32 value plus one Newton correction, and |q| = 1 to 1e-14
// (normalised in fp64 at the start of the step, integrated with compensated increments).
#ifndef BRB_DIST_BAND
#define BRB_DIST_BAND 2e-8f
#endif
struct RimDist { float d[4]; };
BRB_D RimDist rim_dist_fp64(const BrbModelConsts &c, const KF (&q)[4], const KF &pz, float rho32, float irho32) {
  const double w = kjoin(q[0]), x = kjoin(q[1]), y = kjoin(q[2]), z = kjoin(q[3]);
  const double n0 = 2 * (x * z - w * y), n1 = 2 * (y * z + w * x), n2 = 1 - 2 * (x * x + y * y);
  const double r0 = (double)rho32, rho = r0 + 0.5 * (double)irho32 * ((n1 * n1 + n2 * n2) - r0 * r0);
  const double ox = (double)c.ox + (double)c.geo_lo[0], oz = (double)c.oz + (double)c.geo_lo[1], rad = (double)c.rad + (double)c.geo_lo[2],
               hl = (double)c.hl + (double)c.geo_lo[3];
  const double hz = kjoin(pz) - ((double)c.zfloor + (double)c.zfloor_lo);
  const double common = hz + oz * n2 - rad * rho, an = fabs(n0);
  const double dd[4] = {common - ox * n0 - hl * an, common - ox * n0 + hl * an, common + ox * n0 - hl * an, common + ox * n0 + hl * an};
  RimDist r;
#pragma unroll
  for (int k = 0; k < 4; k++) r.d[k] = dd[k] < 0.0 ? fminf((float)dd[k], -1e-30f) : fmaxf((float)dd[k], 0.f);   // the sign survives the narrowing
  return r;
}

// One rim end (slot CI = 2*wheel + end).  Branch-free on purpose: the four slots are independent dependency chains that the
// scheduler interleaves; with one divergent region per slot the kernel ran 13 % slower although it executed fewer
// instructions (a warp is latency-bound here: profiles/README.md, round 2).  A slot that is not in contact gets finite
// values that nothing reads: its row pattern is forced to 0 (all-zero weights) in phys_assemble / phys_active_set.
// (Skipping a wheel's two slots behind a warp-uniform vote when no lane has that wheel on the floor was also slower, 0.826 vs
// 0.760 ms: on this path any branch costs more than the work it saves.)
template <int CI, bool VI>
BRB_D void contact_setup(const BrbModelConsts &c, Phys &P, float dist, float sa, const float (&G)[3], const float (&A)[3],
                         const float (&B2)[3], const float (&ww)[3]) {
  constexpr int k = CI >> 1, e = CI & 1;
  const float sg = k ? 1.f : -1.f;
  P.valid |= (dist < 0.f) ? (1u << CI) : 0u;
  const float hd = 0.5f * dist;
  const float cx = sg * c.ox + (e ? -sa : sa) * c.hl;
  // r_w = R (off_k +- a*hl + v) - e_z dist/2 ; w_w = R [axis_k x (r_b - off_k)] = sg (A + hd B2)
  const float rx = P.ex[0] * cx + G[0], ry = P.ex[1] * cx + G[1], rz = P.ex[2] * cx + G[2] - hd;
  const float wx = sg * (A[0] + hd * B2[0]), wy = sg * (A[1] + hd * B2[1]), wz = sg * (A[2] + hd * B2[2]);
  const float sk = P.s[k].s;
  // material-point velocity in the world frame: v + w_w x r_w + s_k w_w
  const float px = P.v[0].s + ww[1] * rz - ww[2] * ry + sk * wx;
  const float py = P.v[1].s + ww[2] * rx - ww[0] * rz + sk * wy;
  const float pz = P.v[2].s + ww[0] * ry - ww[1] * rx + sk * wz;
  P.cr[CI][0] = rx; P.cr[CI][1] = ry; P.cr[CI][2] = rz;
  P.cw[CI][0] = wx; P.cw[CI][1] = wy; P.cw[CI][2] = wz;
  if (VI) {
    const float imp = imp_of(c.pp[0], dist);
    P.cD[CI] = __fdividef(c.pp[0][3] * imp, 1.f - imp);
    P.cy[CI][0] = c.pp[0][2] * pz + c.pp[0][1] * imp * dist;
    P.cy[CI][1] = c.pp[0][2] * py;
    P.cy[CI][2] = -c.pp[0][2] * px;
  } else {
    P.cy[CI][0] = c.Bdamp * pz + c.Kimp * dist;
    P.cy[CI][1] = c.Bdamp * py;
    P.cy[CI][2] = -c.Bdamp * px;
  }
}

template <bool VI = false>
BRB_D void phys_setup(const BrbModelConsts &c, Phys &P) {
  const float qw = P.q[0].s, qx = P.q[1].s, qy = P.q[2].s, qz = P.q[3].s;
  const float xx = qx * qx, yy = qy * qy, zz = qz * qz, xy = qx * qy, xz = qx * qz, yz = qy * qz, wx = qw * qx, wy = qw * qy, wz = qw * qz;
  P.ex[0] = 1.f - 2.f * (yy + zz); P.ey[0] = 2.f * (xy - wz); P.ez[0] = 2.f * (xz + wy);
  P.ex[1] = 2.f * (xy + wz); P.ey[1] = 1.f - 2.f * (xx + zz); P.ez[1] = 2.f * (yz - wx);
  P.ex[2] = 2.f * (xz - wy); P.ey[2] = 2.f * (yz + wx); P.ez[2] = 1.f - 2.f * (xx + yy);
  // world z axis in the chassis frame = third row of R
  const float n0 = P.ex[2], n1 = P.ey[2], n2 = P.ez[2];
  const float w0 = P.w[0].s, w1 = P.w[1].s, w2 = P.w[2].s, sL = P.s[0].s, sR = P.s[1].s;
  {
    const float gm = c.grav * c.mcz;
    P.fb[0] = -c.mcz * (w0 * w2);
    P.fb[1] = -c.mcz * (w1 * w2);
    P.fb[2] = c.mcz * (w0 * w0 + w1 * w1);
    const float Lx = c.Ixx * w0 + c.Ia * (sR - sL), Ly = c.Iyy * w1, Lz = c.Izz * w2;
    P.fb[3] = -(w1 * Lz - w2 * Ly) + gm * n1;
    P.fb[4] = -(w2 * Lx - w0 * Lz) - gm * n0;
    P.fb[5] = -(w0 * Ly - w1 * Lx);
    // servo: force = clip(kv (u - s), forcerange); u - s evaluated with the hi/lo halves (Q7)
    const float dL = (P.uhi[0] - sL) + (P.ulo[0] + P.s[0].c), dR = (P.uhi[1] - sR) + (P.ulo[1] + P.s[1].c);
    float tL = c.kv * dL, tR = c.kv * dR;
    P.clampL = (tL <= c.frc_lo) || (tL >= c.frc_hi);
    P.clampR = (tR <= c.frc_lo) || (tR >= c.frc_hi);
    tL = fminf(c.frc_hi, fmaxf(c.frc_lo, tL));
    tR = fminf(c.frc_hi, fmaxf(c.frc_lo, tR));
    P.fb[6] = tL - c.damping * sL;
    P.fb[7] = tR - c.damping * sR;
  }
  // plane-cylinder collision (A.6), distances evaluated in the chassis frame: both rim ends of both wheels
  P.valid = 0;
  const float rho2 = n1 * n1 + n2 * n2;
  const float irho = rsqrtf(fmaxf(rho2, 1e-30f));
  const float rho = rho2 * irho;
  // oz*nz - rad*rho without cancellation when upright: (oz-rad) nz + rad (nz - rho), nz - rho = -ny^2/(nz+rho)
  const float diff = (n2 > 0.f) ? -__fdividef(n1 * n1, n2 + rho) : (n2 - rho);   // 2 ulp of a term that is <= 1e-2 of the height
  const float hgt = ((P.p[2].s - c.zfloor) - P.p[2].c) - c.zfloor_lo;
  const float common = hgt + (c.oz - c.rad) * n2 + c.rad * diff;
  const float anx = fabsf(n0);
  const float dmin = common - c.ox * anx - c.hl * anx;          // lowest rim point of either wheel
  if (dmin < BRB_DIST_BAND) {
    const float vy = -c.rad * n1 * irho, vz = -c.rad * n2 * irho;   // rim direction toward the floor, chassis frame: (0, vy, vz)
    const float sa = (n0 > 0.f) ? -1.f : 1.f;
    float G[3], A[3], B2[3], ww[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      G[k] = P.ey[k] * vy + P.ez[k] * (c.oz + vz);               // R (0, vy, oz + vz)
      A[k] = P.ez[k] * vy - P.ey[k] * vz;                        // R (0, -vz, vy)
      B2[k] = P.ey[k] * n2 - P.ez[k] * n1;                       // R (0, nz, -ny)
      ww[k] = P.ex[k] * w0 + P.ey[k] * w1 + P.ez[k] * w2;        // world angular velocity
    }
    const float dL0 = common - c.ox * n0, dR0 = common + c.ox * n0, ha = c.hl * anx;
    RimDist rd;
    rd.d[0] = dL0 - ha; rd.d[1] = dL0 + ha; rd.d[2] = dR0 - ha; rd.d[3] = dR0 + ha;
    if (fminf(fminf(fabsf(rd.d[0]), fabsf(rd.d[1])), fminf(fabsf(rd.d[2]), fabsf(rd.d[3]))) < BRB_DIST_BAND)
      rd = rim_dist_fp64(c, P.q, P.p[2], rho, irho);
    contact_setup<0, VI>(c, P, rd.d[0], sa, G, A, B2, ww);
    contact_setup<1, VI>(c, P, rd.d[1], sa, G, A, B2, ww);
    contact_setup<2, VI>(c, P, rd.d[2], sa, G, A, B2, ww);
    contact_setup<3, VI>(c, P, rd.d[3], sa, G, A, B2, ww);
  }
  if (P.valid) { P.n_contact++; P.n_slots += __popc(P.valid); }
}

// ---- active set at acceleration a (world coordinates): z_c = Pi (P_c a) + yhat_c, pyramid rows E_r . z_c < 0 (A.7, A.8).
// Rows within `eps` of the switching surface keep their previous state (`prev`): either choice gives the same
// force to O(D*eps) ~ 2e-5 N, and without the hysteresis fp32 noise can flip such a row back and forth forever.
template <int CI>
BRB_D unsigned contact_bits(const BrbModelConsts &c, const Phys &P, const float (&a)[8], unsigned prev, float mu, float eps) {
  const float ak = a[6 + (CI >> 1)];
  const float rx = P.cr[CI][0], ry = P.cr[CI][1], rz = P.cr[CI][2];
  const float px = a[0] + a[4] * rz - a[5] * ry + ak * P.cw[CI][0];
  const float py = a[1] + a[5] * rx - a[3] * rz + ak * P.cw[CI][1];
  const float pz = a[2] + a[3] * ry - a[4] * rx + ak * P.cw[CI][2];
  const float z0 = pz + P.cy[CI][0];
  const float z1 = mu * (py + P.cy[CI][1]);
  const float z2 = mu * (P.cy[CI][2] - px);
  const unsigned pb = prev >> (4 * CI);
  const float e0 = (pb & 1u) ? eps : -eps, e1 = (pb & 2u) ? eps : -eps, e2 = (pb & 4u) ? eps : -eps, e3 = (pb & 8u) ? eps : -eps;
  const unsigned nb = ((unsigned)(z0 + z1 < e0) | ((unsigned)(z0 - z1 < e1) << 1) | ((unsigned)(z0 + z2 < e2) << 2) | ((unsigned)(z0 - z2 < e3) << 3)) << (4 * CI);
  return (P.valid & (1u << CI)) ? nb : 0u;
}

template <bool VI = false>
BRB_D unsigned phys_active_set(const BrbModelConsts &c, const Phys &P, const float (&a)[8], unsigned prev, float eps = 2e-4f) {
  const float mu = VI ? c.pp[0][0] : c.mu;
  return contact_bits<0>(c, P, a, prev, mu, eps) | contact_bits<1>(c, P, a, prev, mu, eps) | contact_bits<2>(c, P, a, prev, mu, eps) |
         contact_bits<3>(c, P, a, prev, mu, eps);
}

// ---- S_c = Pi' W_c Pi for the 16 pyramid-row patterns b of a contact: {Szz, Syz, Sxz, Syy, Sxx} (D folded in when it is a
//      model constant; unit D for Env03-v2's position-dependent impedance).  One copy per CTA in shared memory.
#ifdef BRB_HOST_EMU
static float g_stab[16 * 5];
#else
__shared__ float g_stab[16 * 5];
#endif
BRB_D void stab_fill(const BrbModelConsts &c, bool vi, int idx) {
  const unsigned b = (unsigned)idx / 5u, j = (unsigned)idx % 5u;
  const float b0 = (float)(b & 1u), b1 = (float)((b >> 1) & 1u), b2 = (float)((b >> 2) & 1u), b3 = (float)((b >> 3) & 1u);
  const float Dc = vi ? 1.f : c.D, muc = vi ? c.pp[0][0] : c.mu;
  const float Dm = Dc * muc, Dmm = Dm * muc;
  g_stab[idx] = j == 0 ? Dc * (b0 + b1 + b2 + b3) : j == 1 ? Dm * (b0 - b1) : j == 2 ? -Dm * (b2 - b3) : j == 3 ? Dmm * (b0 + b1) : Dmm * (b2 + b3);
}

// ---- H += P_c' S P_c, r -= P_c' S yhat-ish for one contact; S = Pi' W_c Pi in world axes:
//      Sxx = W22, Syy = W11, Szz = W00, Syz = W01, Sxz = -W02, Sxy = 0
// S5 = {Szz, Syz, Sxz, Syy, Sxx} of this slot's row pattern (all zero for a slot that is not in contact or has no active row).
// Branch-free like contact_setup: the four slots interleave.
template <int CI>
BRB_D void contact_assemble(const Phys &P, const float (&S5)[5], float (&H)[36], float (&r)[8]) {
  constexpr int kw = 6 + (CI >> 1);
  const float Szz = S5[0], Syz = S5[1], Sxz = S5[2], Syy = S5[3], Sxx = S5[4];
  const float rx = P.cr[CI][0], ry = P.cr[CI][1], rz = P.cr[CI][2];
  const float wx = P.cw[CI][0], wy = P.cw[CI][1], wz = P.cw[CI][2];
  // columns of P: e_x e_y e_z | cx=(0,-rz,ry) cy=(rz,0,-rx) cz=(-ry,rx,0) | w ;  T_j = S p_j
  const float T3x = Sxz * ry, T3y = -Syy * rz + Syz * ry, T3z = -Syz * rz + Szz * ry;
  const float T4x = Sxx * rz - Sxz * rx, T4y = -Syz * rx, T4z = Sxz * rz - Szz * rx;
  const float T5x = -Sxx * ry, T5y = Syy * rx, T5z = -Sxz * ry + Syz * rx;
  const float Twx = Sxx * wx + Sxz * wz, Twy = Syy * wy + Syz * wz, Twz = Sxz * wx + Syz * wy + Szz * wz;
  H[LT(0, 0)] += Sxx; H[LT(2, 0)] += Sxz; H[LT(1, 1)] += Syy; H[LT(2, 1)] += Syz; H[LT(2, 2)] += Szz;
  H[LT(3, 0)] += T3x; H[LT(3, 1)] += T3y; H[LT(3, 2)] += T3z;
  H[LT(4, 0)] += T4x; H[LT(4, 1)] += T4y; H[LT(4, 2)] += T4z;
  H[LT(5, 0)] += T5x; H[LT(5, 1)] += T5y; H[LT(5, 2)] += T5z;
  H[LT(3, 3)] += -rz * T3y + ry * T3z;
  H[LT(4, 3)] += rz * T3x - rx * T3z;
  H[LT(5, 3)] += -ry * T3x + rx * T3y;
  H[LT(4, 4)] += rz * T4x - rx * T4z;
  H[LT(5, 4)] += -ry * T4x + rx * T4y;
  H[LT(5, 5)] += -ry * T5x + rx * T5y;
  H[LT(kw, 0)] += Twx; H[LT(kw, 1)] += Twy; H[LT(kw, 2)] += Twz;
  H[LT(kw, 3)] += -rz * Twy + ry * Twz;
  H[LT(kw, 4)] += rz * Twx - rx * Twz;
  H[LT(kw, 5)] += -ry * Twx + rx * Twy;
  H[LT(kw, kw)] += wx * Twx + wy * Twy + wz * Twz;
  // rhs: g = -S yhat_w with yhat_w = (-y2, y1, y0) the contact-frame vector in world axes
  const float y0 = P.cy[CI][0], y1 = P.cy[CI][1], y2 = P.cy[CI][2];
  const float gx = Sxx * y2 - Sxz * y0, gy = -(Syy * y1 + Syz * y0), gz = Sxz * y2 - Syz * y1 - Szz * y0;
  r[0] += gx; r[1] += gy; r[2] += gz;
  r[3] += -rz * gy + ry * gz;
  r[4] += rz * gx - rx * gz;
  r[5] += -ry * gx + rx * gy;
  r[kw] += wx * gx + wy * gy + wz * gz;
}

// ---- H = M' + sum P'SP (packed lower 8x8), r = f - sum P'S yhat for the active set `bits`
template <bool VI = false>
BRB_D void phys_assemble(const BrbModelConsts &c, const Phys &P, unsigned bits, float (&H)[36], float (&r)[8]) {
#pragma unroll
  for (int k = 0; k < 36; k++) H[k] = 0.f;
  // M' = blockdiag(R, R, I) M_b blockdiag(R, R, I)'
  H[LT(0, 0)] = c.mass; H[LT(1, 1)] = c.mass; H[LT(2, 2)] = c.mass;
  {
    const float kx = c.mcz * P.ez[0], ky = c.mcz * P.ez[1], kz = c.mcz * P.ez[2];   // M_la = -m [c_w]x, c_w = cz ez
    H[LT(4, 0)] = kz;  H[LT(5, 0)] = -ky;
    H[LT(3, 1)] = -kz; H[LT(5, 1)] = kx;
    H[LT(3, 2)] = ky;  H[LT(4, 2)] = -kx;
    const float dx = c.Ixx - c.Iyy, dz = c.Izz - c.Iyy;                              // R I_O R' = Iyy 1 + dx ex ex' + dz ez ez'
    const float ax = dx * P.ex[0], ay = dx * P.ex[1], az = dx * P.ex[2], bx = dz * P.ez[0], by = dz * P.ez[1], bz = dz * P.ez[2];
    H[LT(3, 3)] = c.Iyy + ax * P.ex[0] + bx * P.ez[0];
    H[LT(4, 3)] = ay * P.ex[0] + by * P.ez[0];
    H[LT(5, 3)] = az * P.ex[0] + bz * P.ez[0];
    H[LT(4, 4)] = c.Iyy + ay * P.ex[1] + by * P.ez[1];
    H[LT(5, 4)] = az * P.ex[1] + bz * P.ez[1];
    H[LT(5, 5)] = c.Iyy + az * P.ex[2] + bz * P.ez[2];
    const float hx = c.Ia * P.ex[0], hy = c.Ia * P.ex[1], hz = c.Ia * P.ex[2];       // wheel axes -+ex
    H[LT(6, 3)] = -hx; H[LT(6, 4)] = -hy; H[LT(6, 5)] = -hz;
    H[LT(7, 3)] = hx;  H[LT(7, 4)] = hy;  H[LT(7, 5)] = hz;
    H[LT(6, 6)] = c.Ia; H[LT(7, 7)] = c.Ia;
  }
  phys_world_force(c, P, r);
#ifdef BRB_TRIPSTATS
  {   // solves whose contacts all have the four rows active (S diagonal and equal for every contact)
    const unsigned want = ((P.valid & 1u) ? 0xFu : 0u) | ((P.valid & 2u) ? 0xF0u : 0u) | ((P.valid & 4u) ? 0xF00u : 0u) | ((P.valid & 8u) ? 0xF000u : 0u);
    atomicAdd(&g_trip[192], 1ull);
    if ((bits & want) == want) atomicAdd(&g_trip[193], 1ull);
  }
#endif
  // row-pattern weights of all four slots up front from the 16-entry table in shared memory (stab_fill; pattern 0 = all-zero
  // weights, also forced for a slot that is not in contact): the LDS latency overlaps with M' above
  const unsigned vm = ((P.valid & 1u) ? 0xFu : 0u) | ((P.valid & 2u) ? 0xF0u : 0u) | ((P.valid & 4u) ? 0xF00u : 0u) | ((P.valid & 8u) ? 0xF000u : 0u);
  const unsigned bm = bits & vm;
  float S5[4][5];
#pragma unroll
  for (int ci = 0; ci < 4; ci++) {
    const float *t = g_stab + 5 * ((bm >> (4 * ci)) & 15u);
#ifdef BRB_TRIPSTATS
    if (P.valid & (1u << ci)) atomicAdd(&g_trip[176 + ((bm >> (4 * ci)) & 15u)], 1ull);   // pyramid-row pattern of every contact at every solve
#endif
#pragma unroll
    for (int k = 0; k < 5; k++) S5[ci][k] = VI ? t[k] * P.cD[ci] : t[k];
  }
  contact_assemble<0>(P, S5[0], H, r);
  contact_assemble<1>(P, S5[1], H, r);
  contact_assemble<2>(P, S5[2], H, r);
  contact_assemble<3>(P, S5[3], H, r);
}

// ---- one Newton step on the active set `bits`: (M' + sum P'SP) a = f - sum P'S yhat   (A.8; exact for a fixed set)
template <bool VI = false>
BRB_D void phys_solve(const BrbModelConsts &c, Phys &P, unsigned bits, float (&a)[8]) {
  float H[36], r[8];
  phys_assemble<VI>(c, P, bits, H, r);
  // Cholesky H = L L' and the two triangular solves, straight-line (generated: gen_chol8.py; H(1,0) is structurally zero here)
#include "brb_chol8z.inc"
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = r[k];
  P.n_solve++;
}

// ---- A.9 + A.10: implicitfast velocity update, then positions with the NEW velocities.
// av = world linear acceleration, ab = chassis-frame angular acceleration, a6/a7 = wheels: the SOLVER's acceleration.
BRB_D void phys_finalize(const BrbModelConsts &c, Phys &P, float avx, float avy, float avz, float ab0, float ab1, float ab2, float a6, float a7) {
  {
    // a+ = (M + h Dv)^-1 M a = a - Wm (Cinv + G)^-1 [a6; a7]   (Woodbury on the two wheel dofs; Wm touches uy, wx, sL, sR)
    // (Cinv + G)^-1 depends only on which servos sit on their forcerange: four host-computed 2x2 inverses (constant bank)
    const bool skip = (c.flags & BRB_FLAG_ACTDERIV_SKIP_CLAMPED) != 0;
    const int ci = skip ? ((P.clampL ? 1 : 0) | (P.clampR ? 2 : 0)) : 0;
    const float i00 = c.impl_Kinv[ci][0], i01 = c.impl_Kinv[ci][1], i11 = c.impl_Kinv[ci][2];
    const float y0 = i00 * a6 + i01 * a7, y1 = i01 * a6 + i11 * a7;
    const float duy = c.impl_W[0] * y0 + c.impl_W[1] * y1;      // chassis-frame y component of the linear correction
    avx -= P.ey[0] * duy; avy -= P.ey[1] * duy; avz -= P.ey[2] * duy;
    ab0 -= c.impl_W[2] * y0 + c.impl_W[3] * y1;
    a6 -= c.impl_W[4] * y0 + c.impl_W[5] * y1;
    a7 -= c.impl_W[6] * y0 + c.impl_W[7] * y1;
  }
  const float h = c.h;
  kadd(P.v[0], h * avx); kadd(P.v[1], h * avy); kadd(P.v[2], h * avz);
  kadd(P.w[0], h * ab0); kadd(P.w[1], h * ab1); kadd(P.w[2], h * ab2);
  kadd(P.s[0], h * a6); kadd(P.s[1], h * a7);
  kadd(P.p[0], h * P.v[0].s - h * P.v[0].c);
  kadd(P.p[1], h * P.v[1].s - h * P.v[1].c);
  kadd(P.p[2], h * P.v[2].s - h * P.v[2].c);
  kadd(P.th[0], h * P.s[0].s); kadd(P.th[1], h * P.s[1].s);
  {
    // q <- q (x) [cos(t/2), sin(t/2) w/|w|], t = h |w|; added as the increment q (x) (dq - 1)
    const float qw = P.q[0].s, qx = P.q[1].s, qy = P.q[2].s, qz = P.q[3].s;
    const float n0 = P.w[0].s, n1 = P.w[1].s, n2 = P.w[2].s;
    const float t2 = (h * h) * (n0 * n0 + n1 * n1 + n2 * n2);
    const float sn = (0.5f * h) * (1.f - t2 * (1.f / 24.f));
    const float cm1 = -(0.125f * t2) * (1.f - t2 * (1.f / 48.f));
    const float ex = sn * n0, ey = sn * n1, ez = sn * n2;
    kadd(P.q[0], qw * cm1 - (qx * ex + qy * ey + qz * ez));
    kadd(P.q[1], qx * cm1 + (qw * ex + qy * ez - qz * ey));
    kadd(P.q[2], qy * cm1 + (qw * ey - qx * ez + qz * ex));
    kadd(P.q[3], qz * cm1 + (qw * ez + qx * ey - qy * ex));
  }
}

// ---- nsub substeps as a per-lane state machine.  Every trip of the loop does at most ONE 8x8 solve per
// lane; a lane whose active set changed simply repeats the solve in the next trip while its neighbours move on to
// their next substep, so a warp pays max-over-lanes(nsub + extra solves) trips instead of nsub * max-over-lanes
// (solves per substep).  qstale receives the quaternion before the last integration (Q1).
// One loop trip.  (A second copy of this body without the `done` tests for the first nsub trips -- no lane can be finished before --
// was measured slower: 0.893 vs 0.776 ms, the two copies no longer fit the instruction cache together.)
template <int MAXIT>
BRB_D void phys_trip(const BrbModelConsts &c, Phys &P, const int nsub, KF (&qstale)[4], const unsigned wmask, int &sidx, int &it,
                     bool &need_setup, bool &done, unsigned &was) {
  const bool live = !done;
  if (live && need_setup) {
    phys_setup(c, P);
    const unsigned fresh = P.valid & ~was;     // slots that were not in contact a substep ago start with all four rows active
    P.bits |= ((fresh & 1u) ? 0xFu : 0u) | ((fresh & 2u) ? 0xF0u : 0u) | ((fresh & 4u) ? 0xF00u : 0u) | ((fresh & 8u) ? 0xF000u : 0u);
    was = P.valid;
    need_setup = false;
    it = 0;
  }
  bool conv = true;
  float avx, avy, avz, ab0, ab1, ab2, a6, a7;
#ifdef BRB_TRIPSTATS
  {
    const unsigned act = __ballot_sync(wmask, live);
    const unsigned nv = __popc(__ballot_sync(wmask, live && P.valid != 0u));
    if ((threadIdx.x & 31u) == (unsigned)(__ffs(wmask) - 1)) {
      atomicAdd(&g_trip[0], 1ull); atomicAdd(&g_trip[1], (unsigned long long)__popc(act));
      if (nv) { atomicAdd(&g_trip[2], 1ull); atomicAdd(&g_trip[3], (unsigned long long)nv); }
      atomicAdd(&g_trip[8 + nv], 1ull);
    }
  }
#endif
  __syncwarp(wmask);
  if (live) {
    if (P.valid) {
      float a[8];
      phys_solve(c, P, P.bits, a);
      const unsigned nb = phys_active_set(c, P, a, P.bits);
      conv = (nb == P.bits);
      P.bits = nb;
      if (!conv && ++it >= MAXIT) { P.n_nonconv++; conv = true; }
      avx = a[0]; avy = a[1]; avz = a[2];
      ab0 = P.ex[0] * a[3] + P.ex[1] * a[4] + P.ex[2] * a[5];      // alpha_b = R' alpha_w
      ab1 = P.ey[0] * a[3] + P.ey[1] * a[4] + P.ey[2] * a[5];
      ab2 = P.ez[0] * a[3] + P.ez[1] * a[4] + P.ez[2] * a[5];
      a6 = a[6]; a7 = a[7];
    } else {
      // free flight: a_b = M_b^-1 f_b in the chassis frame, linear part rotated to the world
      const float mg = c.mass * c.grav;
      const float f[8] = {P.fb[0] - mg * P.ex[2], P.fb[1] - mg * P.ey[2], P.fb[2] - mg * P.ez[2], P.fb[3], P.fb[4], P.fb[5], P.fb[6], P.fb[7]};
      const float u0 = c.minv_xy[0] * f[0] + c.minv_xy[1] * f[4];
      const float u1 = c.minv_blk[0] * f[1] + c.minv_blk[1] * f[3] + c.minv_blk[2] * f[6] + c.minv_blk[3] * f[7];
      const float u2 = c.minv_uz * f[2];
      avx = P.ex[0] * u0 + P.ey[0] * u1 + P.ez[0] * u2;
      avy = P.ex[1] * u0 + P.ey[1] * u1 + P.ez[1] * u2;
      avz = P.ex[2] * u0 + P.ey[2] * u1 + P.ez[2] * u2;
      ab0 = c.minv_blk[1] * f[1] + c.minv_blk[4] * f[3] + c.minv_blk[5] * f[6] + c.minv_blk[6] * f[7];
      ab1 = c.minv_xy[1] * f[0] + c.minv_xy[2] * f[4];
      ab2 = c.minv_wz * f[5];
      a6 = c.minv_blk[2] * f[1] + c.minv_blk[5] * f[3] + c.minv_blk[7] * f[6] + c.minv_blk[8] * f[7];
      a7 = c.minv_blk[3] * f[1] + c.minv_blk[6] * f[3] + c.minv_blk[8] * f[6] + c.minv_blk[9] * f[7];
    }
  }
  __syncwarp(wmask);
  if (live && conv) {
    if (sidx == nsub - 1) {
#pragma unroll
      for (int k = 0; k < 4; k++) qstale[k] = P.q[k];
    }
    phys_finalize(c, P, avx, avy, avz, ab0, ab1, ab2, a6, a7);
    if (++sidx >= nsub) done = true;
    need_setup = true;
  }
}

template <int MAXIT>
BRB_D void phys_run(const BrbModelConsts &c, Phys &P, int nsub, KF (&qstale)[4], const unsigned wmask) {
  int sidx = 0, it = 0;
  // The active set of a new substep is seeded with the previous substep's converged set (rows of a contact that just
  // appeared start "all active"): right ~97 % of the time, and the post-solve check catches the rest.
  // phys_setup has a single call site per loop (flag instead of a second inlined copy) to keep the loop body small.
  // The loops are warp-uniform: lanes that finished keep circulating (idle) until every lane of `wmask` (the lanes of this
  // warp that run a robot) is done, which makes the __syncwarp()s legal.  They are there because the compiler otherwise
  // threads the free-flight branch straight into the integration code and the two halves of a mixed warp run it one
  // after the other (ncu source page: phys_finalize executed 1.5x per trip at 20 lanes).
  bool need_setup = true, done = false;
  unsigned was = P.valid_prev;
#ifdef BRB_TIMELINE
  P.n_trips = 0;
#endif
  for (;;) {
    if (!__any_sync(wmask, !done)) break;
#ifdef BRB_TIMELINE
    P.n_trips++;
#endif
    phys_trip<MAXIT>(c, P, nsub, qstale, wmask, sidx, it, need_setup, done, was);
  }
}

// contact-slot mask (bit = 2*wheel + rim end) -> bucket rank, ordered by number of contacts, then by pattern
BRB_D unsigned group_rank(unsigned mask) {
  //                         mask: 0  1  2  3  4  5   6  7   8  9  10  11  12  13  14  15
  const unsigned long long lut = 0x0ull | (1ull << 4) | (3ull << 8) | (6ull << 12) | (2ull << 16) | (5ull << 20) | (9ull << 24) |
                                 (11ull << 28) | (4ull << 32) | (10ull << 36) | (8ull << 40) | (13ull << 44) | (7ull << 48) |
                                 (12ull << 52) | (14ull << 56) | (15ull << 60);
  return (unsigned)((lut >> (4 * (mask & 15u))) & 15ull);
}

// lowest wheel-rim height above the floor for a pose (fp32), used only to group envs for the next step
BRB_D float pose_clearance(const BrbModelConsts &c, float qw, float qx, float qy, float qz, float z) {
  const float nx = 2.f * (qx * qz - qw * qy), ny = 2.f * (qy * qz + qw * qx), nz = 1.f - 2.f * (qx * qx + qy * qy);
  const float rho = sqrtf(ny * ny + nz * nz);
  return (z - c.zfloor) + c.oz * nz - c.rad * rho - (c.ox + c.hl) * fabsf(nx);
}
BRB_D float phys_clearance(const BrbModelConsts &c, const Phys &P) {
  return pose_clearance(c, P.q[0].s, P.q[1].s, P.q[2].s, P.q[3].s, P.p[2].s);
}
// Group key of an AIRBORNE robot for the next step's visit order.  Ballistic prediction: the height is concave in time, so
// the minimum over the step is at one of its ends; the pose at the end is the free-flight one (z + vz T - g T^2/2,
// q (x) exp(w T / 2) to first order; wheel torques only pitch the chassis about the wheel axis, which does not move the rims).
// 0 = stays clear for the whole step; 1..BRB_NLAND = touches down, binned by the predicted landing time so that the lanes
// of a warp switch from the cheap free-flight path to the contact solve within a few substeps of each other.
BRB_D unsigned airborne_key(const BrbModelConsts &c, float qw, float qx, float qy, float qz, float z, float vz, float w0, float w1, float w2) {
  const float T = c.h * (float)c.frame_skip, margin = 1e-4f;
  const float hx = 0.5f * T * w0, hy = 0.5f * T * w1, hz = 0.5f * T * w2;
  float rw = qw - (qx * hx + qy * hy + qz * hz), rx = qx + (qw * hx + qy * hz - qz * hy);
  float ry = qy + (qw * hy - qx * hz + qz * hx), rz = qz + (qw * hz + qx * hy - qy * hx);
  const float inv = rsqrtf(rw * rw + rx * rx + ry * ry + rz * rz);
  rw *= inv; rx *= inv; ry *= inv; rz *= inv;
  const float gT = 0.5f * c.grav * T;
  const float c0 = pose_clearance(c, qw, qx, qy, qz, z) - margin;
  const float cT = pose_clearance(c, rw, rx, ry, rz, z + T * (vz - gT)) - margin;
  if (c0 > 0.f && cT > 0.f) return 0u;
  if (c0 <= 0.f) return 1u;
  // c(t) = c0 + v t - g t^2 / 2 with v fitted to c(T) = cT; first root, as a fraction of the step
  const float v = (cT - c0) / T + gT;
  const float tl = (v + sqrtf(v * v + 2.f * c.grav * c0)) / c.grav;
  const int bin = (int)(tl / T * (float)BRB_NLAND);
  return 1u + (unsigned)(bin < 0 ? 0 : (bin >= BRB_NLAND ? BRB_NLAND - 1 : bin));
}

// ---------------------------------------------------------------------------------------------------
#ifndef BRB_HOST_EMU
__device__ __forceinline__ unsigned long long warp_sum(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}
#endif

// One VecEnv.step for env i of the shard.  stat[8] receives {substeps, contact substeps, solves, non-converged,
// unsupported-pose flag, done flag, group hint for the next step, contact slots in use}.
template <int KIND>
BRB_D void step_env(const BrbModelConsts &c, const BrbState &S, const long long i, const float *__restrict__ actions,
                    float *__restrict__ obs, float *__restrict__ reward, uint8_t *__restrict__ done,
                    uint8_t *__restrict__ truncated, float *__restrict__ terminal_obs, float *__restrict__ ep_return_out,
                    int32_t *__restrict__ ep_len_out, const double *__restrict__ replay_u, unsigned stat[12], const unsigned wmask) {
  const long long N = S.n;
  // ---------------- prologue (fp64 task logic on the pre-step state) ----------------
  double qvel[8], xq[4];
#pragma unroll
  for (int k = 0; k < 8; k++) qvel[k] = S.qvel[k * N + i];
#pragma unroll
  for (int k = 0; k < 4; k++) xq[k] = S.xquat[k * N + i];
  const uint32_t event = S.event[i] + 1u;
  S.event[i] = event;
  double us[4];
  if (replay_u) {
#pragma unroll
    for (int k = 0; k < 4; k++) us[k] = replay_u[i * 20 + k];
  } else if (KIND == BRB_ENV01_V2) {
    draw4(S.seed, (uint64_t)(S.env0 + i), event, 0u, us);
  } else {
    us[0] = us[1] = us[2] = us[3] = 0.0;
  }
  int elapsed = S.elapsed[i];
  double tws = 0.0, poff = 0.0;
  if (KIND == BRB_ENV01_V3) {
    const double t = S.time_table[elapsed], dts = S.v3[1 * N + i];   // env01_v3.py:27-37 (pre-step data.time)
    tws = S.v3[0 * N + i];
    poff = S.v3[2 * N + i];
    if (t > 5.5) tws = 3.0 * dts;
    else if (t > 4.5) tws = 2.0 * dts;
    else if (t > 3.0) tws = -1.0 * dts;
    else if (t > 1.0) tws = dts;
    S.v3[0 * N + i] = tws;
  }
  const double rew = reward_of<KIND>(noisy_pitch<KIND>(pitch_of(xq), us[0], poff), qvel[6], qvel[7], qvel[5], tws, 0.0);
  const float act_x = actions[2 * i], act_y = actions[2 * i + 1];
  double ctrl[2] = {qvel[6] + (double)act_x * 4.0, qvel[7] + (double)act_y * 4.0};   // env01_v1.py:18-23

  // ---------------- 250 substeps ----------------
  Phys st;
  {
    double qn[4], nn = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { qn[k] = S.qpos[(3 + k) * N + i]; nn += qn[k] * qn[k]; }
    nn = 1.0 / sqrt(nn);
#pragma unroll
    for (int k = 0; k < 4; k++) st.q[k] = ksplit(qn[k] * nn);
#pragma unroll
    for (int k = 0; k < 3; k++) { st.p[k] = ksplit(S.qpos[k * N + i]); st.v[k] = ksplit(qvel[k]); st.w[k] = ksplit(qvel[3 + k]); }
#pragma unroll
    for (int k = 0; k < 2; k++) {
      st.th[k] = ksplit(S.qpos[(7 + k) * N + i]);
      st.s[k] = ksplit(qvel[6 + k]);
      const double u = fmin((double)c.ctrl_hi, fmax((double)c.ctrl_lo, ctrl[k]));
      st.uhi[k] = (float)u;
      st.ulo[k] = (float)(u - (double)st.uhi[k]);
    }
    st.bits = S.aset[i];
    st.valid_prev = 0xFu;     // keep the persisted active set as it is on the first substep of the step
    st.n_contact = st.n_solve = st.n_nonconv = st.n_slots = 0;
    // contact records are read branch-free (with zero weights when the slot is not in contact): they must hold finite values
#pragma unroll
    for (int ci = 0; ci < 4; ci++) {
      st.cD[ci] = 0.f;
#pragma unroll
      for (int k = 0; k < 3; k++) st.cr[ci][k] = st.cw[ci][k] = st.cy[ci][k] = 0.f;
    }
  }
  const int nsub = c.frame_skip;
  KF qprev[4];
  phys_run<BRB_MAXIT>(c, st, nsub, qprev, wmask);
  stat[0] = nsub; stat[1] = st.n_contact; stat[2] = st.n_solve; stat[3] = st.n_nonconv; stat[7] = st.n_slots;
#ifdef BRB_TIMELINE
  stat[8] = st.n_trips;
#endif
  {
    // group key for the next step's visit order (no effect on results): robots are bucketed by which wheel-rim contact
    // slots they ended the step with, airborne ones by whether / when they will reach the floor during the next step
    stat[6] = (st.valid == 0u) ? airborne_key(c, st.q[0].s, st.q[1].s, st.q[2].s, st.q[3].s, st.p[2].s, st.v[2].s, st.w[0].s, st.w[1].s, st.w[2].s)
                               : BRB_NLAND + group_rank(st.valid);
  }

  // ---------------- epilogue ----------------
  double qpos[9];
#pragma unroll
  for (int k = 0; k < 3; k++) { qpos[k] = kjoin(st.p[k]); qvel[k] = kjoin(st.v[k]); qvel[3 + k] = kjoin(st.w[k]); }
#pragma unroll
  for (int k = 0; k < 4; k++) qpos[3 + k] = kjoin(st.q[k]);
#pragma unroll
  for (int k = 0; k < 2; k++) { qpos[7 + k] = kjoin(st.th[k]); qvel[6 + k] = kjoin(st.s[k]); }
  {
    double nn = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { xq[k] = kjoin(qprev[k]); nn += xq[k] * xq[k]; }
    nn = 1.0 / sqrt(nn);
#pragma unroll
    for (int k = 0; k < 4; k++) xq[k] *= nn;
  }
  elapsed += 1;
  const double p_true = pitch_of(xq);
  const bool terminated = fabs(noisy_pitch<KIND>(p_true, us[1], poff)) > (50 * BRB_PI / 180);   // env01_v1.py:31
  const bool trunc = elapsed >= c.max_episode_steps;                                            // gymnasium TimeLimit
  const double pitch = noisy_pitch<KIND>(p_true, us[2], poff);
  const double pitch2 = noisy_pitch<KIND>(p_true, us[3], poff);
  const double dt = S.time_table[elapsed] - S.time_table[elapsed - 1];
  double pitch_dot = 0.0;
  if (dt > 0.0) pitch_dot = (pitch2 - S.last_pitch[i]) / dt;   // RobotBaseEnv.py:142-157
  float o[6];
  obs_of(pitch, pitch_dot, qvel[6], qvel[7], tws, 0.0, o);
  const double epr = S.ep_return[i] + rew;
  const int epl = S.ep_len[i] + 1;
  // poses whose contacts this kernel does not model (chassis-floor, wheel lying flat): count them
  {
    const double w = xq[0], x = xq[1], y = xq[2], z = xq[3];
    const double n0 = 2 * (x * z - w * y), n1 = 2 * (y * z + w * x), n2 = 1 - 2 * (x * x + y * y);
    const double hz = qpos[2] - ((double)c.zfloor + (double)c.zfloor_lo);
    double low = hz + n2 * (double)c.chassis_pos[2] + n0 * (double)c.chassis_pos[0] + n1 * (double)c.chassis_pos[1]
                 - fabs(n0) * (double)c.chassis_half[0] - fabs(n1) * (double)c.chassis_half[1] - fabs(n2) * (double)c.chassis_half[2];
    const double rho = sqrt(n1 * n1 + n2 * n2);
    const double tri = hz + (double)c.oz * n2 - fabs(n0) * ((double)c.ox + (double)c.hl) + 0.5 * (double)c.rad * rho;
    stat[4] = (low <= 0.0 || tri <= 0.0) ? 1u : 0u;
  }
  const bool cut = stat[4] != 0u && (c.flags & BRB_FLAG_TRUNCATE_UNSUPPORTED) != 0;      // opt-in: end the episode as truncated

  const bool dn = terminated || trunc || cut;
  if (reward) reward[i] = (float)rew;
  if (done) done[i] = (uint8_t)dn;
  if (truncated) truncated[i] = (uint8_t)((trunc || cut) && !terminated);
  if (ep_return_out) ep_return_out[i] = (float)epr;
  if (ep_len_out) ep_len_out[i] = epl;

  if (dn) {
    stat[5] = 1;
    if (terminal_obs) {
#pragma unroll
      for (int k = 0; k < 6; k++) terminal_obs[i * 6 + k] = o[k];
    }
    double ur[16];
    if (replay_u) {
#pragma unroll
      for (int k = 0; k < 16; k++) ur[k] = replay_u[i * 20 + 4 + k];
    } else {
#pragma unroll
      for (int b = 0; b < 4; b++) draw4(S.seed, (uint64_t)(S.env0 + i), event, 1u + b, ur + 4 * b);
    }
    reset_env<KIND>(S, i, ur, o);
    // reset pose: wheels ~2 cm above the floor (Q11), but v2's +-1 rad of pitch about the body origin plus the roll (Q3)
    // puts some rims on the floor right away
    stat[6] = airborne_key(c, (float)S.qpos[3 * N + i], (float)S.qpos[4 * N + i], (float)S.qpos[5 * N + i], (float)S.qpos[6 * N + i], 0.f, 0.f, 0.f, 0.f, 0.f);
  } else {
#pragma unroll
    for (int k = 0; k < 9; k++) S.qpos[k * N + i] = qpos[k];
#pragma unroll
    for (int k = 0; k < 8; k++) S.qvel[k * N + i] = qvel[k];
    S.aset[i] = st.bits;
#pragma unroll
    for (int k = 0; k < 4; k++) S.xquat[k * N + i] = xq[k];
    S.elapsed[i] = elapsed;
    S.last_pitch[i] = pitch2;
    S.ep_return[i] = epr;
    S.ep_len[i] = epl;
  }
  #pragma unroll
  for (int k = 0; k < 6; k++) obs[i * 6 + k] = o[k];
}

#include "brb_env03.cuh"

#ifndef BRB_HOST_EMU
#ifdef BRB_MAXNREG   // kernel-tuning experiments: explicit register cap instead of the occupancy hint
#define BRB_STEP_BOUNDS(KIND) __maxnreg__(BRB_MAXNREG)
#else
#define BRB_STEP_BOUNDS(KIND) __launch_bounds__(BRB_IS_ENV03(KIND) ? BRB_BLOCK_ENV03 : BRB_BLOCK, BRB_IS_ENV03(KIND) ? BRB_MINBLOCKS_ENV03 : BRB_MINBLOCKS)
#endif
// One robot of the visit order per lane: step it, publish its group key for the next step's order, add to the statistics.
template <int KIND>
__device__ __forceinline__ void step_batch(const BrbModelConsts &c, const BrbState &S, const BrbPerm &perm, const long long tid, const bool ctasync,
                                           const float *__restrict__ actions, float *__restrict__ obs, float *__restrict__ reward,
                                           uint8_t *__restrict__ done, uint8_t *__restrict__ truncated, float *__restrict__ terminal_obs,
                                           float *__restrict__ ep_return_out, int32_t *__restrict__ ep_len_out, const double *__restrict__ replay_u) {
  const bool live = tid < S.n;
  // envs are visited in the order of the partition built by the previous step (grounded robots by contact pattern
  // first, then the landing ones by landing time, airborne last), so a warp's lanes mostly run the same path and the
  // cheap CTAs fill the tail of the launch (state columns are addressed by env id)
  const long long i = live ? (perm.in ? (long long)perm.in[tid] : tid) : 0;
  unsigned stat[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const unsigned wmask = __ballot_sync(0xFFFFFFFFu, live);   // lanes of this warp that run a robot
  if (live) {
    if (BRB_IS_ENV03(KIND)) step_env03<KIND == BRB_ENV03_V2_WB>(c, S, i, actions, obs, reward, done, truncated, terminal_obs, ep_return_out, ep_len_out, replay_u, stat, wmask, ctasync);
    else step_env<KIND>(c, S, i, actions, obs, reward, done, truncated, terminal_obs, ep_return_out, ep_len_out, replay_u, stat, wmask);
  }
  if (perm.key_out) {
    // publish this robot's group key and add it to the histogram the grouping kernel turns into bucket offsets
    const unsigned key = live ? stat[6] : 31u;
#ifdef BRB_TRIPSTATS
    if (live) {
      const unsigned cls = stat[1] == 0u ? 0u : (stat[1] >= stat[0] ? 3u : (2u * stat[1] < stat[0] ? 1u : 2u));
      atomicAdd(&g_trip[48 + 4 * (perm.key_out[i] & 31u) + cls], 1ull);
    }
#endif
    if (live) perm.key_out[i] = (uint8_t)key;
    const unsigned same = __match_any_sync(0xFFFFFFFFu, key);
    if (live && (threadIdx.x & 31u) == (unsigned)(__ffs(same) - 1)) atomicAdd(&perm.hist[key], (unsigned)__popc(same));
  }
#ifdef BRB_TIMELINE
  {
    const unsigned long long ns = warp_sum(stat[2]), nc = warp_sum(stat[1]);
    if ((threadIdx.x & 31) == 0 && tid < 65536 * 32) { g_tl_trips[2 * (tid >> 5)] = stat[8]; g_tl_trips[2 * (tid >> 5) + 1] = (unsigned)(ns - nc); }
  }
#endif
  // statistics: one atomic per warp per counter
  const unsigned long long a0 = warp_sum(stat[0]), a1 = warp_sum(stat[1]), a2 = warp_sum(stat[2]), a3 = warp_sum(stat[3]),
                           a4 = warp_sum(stat[4]), a5 = warp_sum(stat[5]), a6 = warp_sum(live ? 1u : 0u), a7 = warp_sum(stat[7]),
                           stat8 = warp_sum(stat[8]), stat9 = warp_sum(stat[9]), stat10 = warp_sum(stat[10]), stat11 = warp_sum(stat[11]);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&S.stats[BRB_STAT_SUBSTEPS], a0);
    atomicAdd(&S.stats[BRB_STAT_CONTACT_SUBSTEPS], a1);
    atomicAdd(&S.stats[BRB_STAT_SOLVES], a2);
    if (a3) atomicAdd(&S.stats[BRB_STAT_NONCONVERGED], a3);
    if (a4) atomicAdd(&S.stats[BRB_STAT_UNSUPPORTED], a4);
    if (a5) atomicAdd(&S.stats[BRB_STAT_EPISODES], a5);
    atomicAdd(&S.stats[BRB_STAT_ENV_STEPS], a6);
    atomicAdd(&S.stats[BRB_STAT_CONTACT_SLOTS], a7);
    if (BRB_IS_ENV03(KIND)) {
      const unsigned long long a8 = stat8, a9 = stat9;
      if (a8) atomicAdd(&S.stats[BRB_STAT_COUPLED_SUBSTEPS], a8);
      if (a9) atomicAdd(&S.stats[BRB_STAT_BLOCK_CONTACT_SUBSTEPS], a9);
      if (stat10) atomicAdd(&S.stats[10], (unsigned long long)stat10);
      if (stat11) atomicAdd(&S.stats[11], (unsigned long long)stat11);
    }
  }
}

template <int KIND>
__global__ void BRB_STEP_BOUNDS(KIND) brb_step_kernel(const __grid_constant__ BrbModelConsts c, const BrbState S, const BrbPerm perm,
                                                const float *__restrict__ actions, float *__restrict__ obs,
                                                float *__restrict__ reward, uint8_t *__restrict__ done,
                                                uint8_t *__restrict__ truncated, float *__restrict__ terminal_obs,
                                                float *__restrict__ ep_return_out, int32_t *__restrict__ ep_len_out,
                                                const double *__restrict__ replay_u) {
  for (int k = threadIdx.x; k < 16 * 5; k += blockDim.x) stab_fill(c, BRB_IS_ENV03(KIND), k);
  // Work queue (Env01-*): the launch has at most one resident wave of CTAs and every WARP pulls the next 32 robots of the visit
  // order from an atomic cursor until the order is exhausted.  A fixed one-robot-per-thread grid of 512 CTAs on 444 slots runs
  // 1.15 waves: the tail wave costs a whole step-time for 13 % of the robots.
  const bool queue = !BRB_IS_ENV03(KIND) && perm.cursor != nullptr;
  long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // Env03-v2: when every thread of the CTA runs a robot, its warps walk the substep loop in lockstep (CTA barriers at the
  // phase boundaries) so that they share the instruction cache: the loop body is 80 KB per trip, the L1.5 I-cache 32 KB
  const bool ctasync = __syncthreads_and(tid < S.n) != 0;
  for (;;) {
    if (queue) {
      unsigned base = 0;
      if ((threadIdx.x & 31u) == 0u) base = atomicAdd(perm.cursor, 32u);
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if ((long long)base >= S.n) break;
      tid = (long long)base + (threadIdx.x & 31u);
    }
#ifdef BRB_TIMELINE
    unsigned long long tl0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl0));
#endif
    step_batch<KIND>(c, S, perm, tid, ctasync && !queue, actions, obs, reward, done, truncated, terminal_obs, ep_return_out, ep_len_out, replay_u);
#ifdef BRB_TIMELINE
    if ((threadIdx.x & 31u) == 0u) {   // debug: one record per warp task {sm, warp slot, first robot of the visit order, start ns, end ns}
      unsigned long long tl1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl1));
      unsigned smid, wid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
      const unsigned long long slot = atomicAdd(&g_timeline[0], 1ull);
      if (slot < (1ull << 16)) { unsigned long long *r = g_timeline + 1 + 4 * slot; r[0] = ((unsigned long long)smid << 32) | wid; r[1] = (unsigned long long)tid; r[2] = tl0; r[3] = tl1; }
    }
#endif
    if (!queue) break;
  }
}

// Counting sort of the envs by group key -> visit order of the next step (descending key).  hist[] was accumulated by the step kernel;
// order inside a bucket is arbitrary (atomic cursor) and irrelevant to the results, which are per-env deterministic.
__global__ void brb_group_kernel(long long n, const uint8_t *__restrict__ key, const unsigned *__restrict__ hist, unsigned *cursor,
                                 int *__restrict__ order, unsigned *hist_zero, unsigned *cursor_zero, unsigned *queue_cursor) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid == 0 && queue_cursor) *queue_cursor = 0u;     // the step kernel's work-queue cursor, for the next launch
  const bool live = tid < n;
  const unsigned k = live ? key[tid] : 31u;
  // most expensive groups first (grounded robots, then the landing ones, airborne last): the cheap CTAs fill the tail of
  // the launch.  Measured at 65,536 robots: 1.05 ms cheapest-first, 0.93 ms this way (128-thread CTAs).
  unsigned base = 0;
#ifdef BRB_ORDER_CHEAP_FIRST   // kernel-tuning experiment
  for (unsigned j = 0; j < k && j < BRB_NGROUPS; j++) base += hist[j];
#else
  for (unsigned j = k + 1; j < BRB_NGROUPS; j++) base += hist[j];
#endif
  const unsigned lane = threadIdx.x & 31u;
  const unsigned same = __match_any_sync(0xFFFFFFFFu, k);
  const unsigned leader = (unsigned)(__ffs(same) - 1);
  unsigned off = 0;
  if (live && lane == leader) off = atomicAdd(&cursor[k], (unsigned)__popc(same));
  off = __shfl_sync(0xFFFFFFFFu, off, leader);
  if (live) order[base + off + __popc(same & ((1u << lane) - 1u))] = (int)tid;
  if (tid < BRB_NGROUPS) { hist_zero[tid] = 0u; cursor_zero[tid] = 0u; }
}

// ---- finished-episode records of one step, compacted in ascending env order (host path: only these rows cross PCIe).
// Row = BRB_DONE_ROW_WORDS 32-bit words: [env index (i32) | terminal_obs[6] (f32) | ep_return (f32) | ep_len (i32) | truncated (i32)].
// Pass 1: per-CTA counts; the last CTA to arrive scans them into exclusive bases and the total.  Pass 2: rows.
__global__ void brb_done_count_kernel(long long n, const uint8_t *__restrict__ done, unsigned *__restrict__ block_count,
                                      unsigned *__restrict__ block_base, unsigned *ticket, int *n_done) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cnt = __syncthreads_count(i < n && done[i] != 0);
  __shared__ bool last;
  __shared__ unsigned part[256];
  if (threadIdx.x == 0) {
    block_count[blockIdx.x] = (unsigned)cnt;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  const unsigned nb = gridDim.x, per = (nb + blockDim.x - 1) / blockDim.x;
  const unsigned lo = threadIdx.x * per, hi = min(nb, lo + per);
  unsigned sum = 0;
  for (unsigned b = lo; b < hi; b++) sum += ((volatile unsigned *)block_count)[b];
  part[threadIdx.x] = sum;
  __syncthreads();
  for (unsigned o = 1; o < blockDim.x; o <<= 1) {       // inclusive scan of the 256 partial sums
    const unsigned v = threadIdx.x >= o ? part[threadIdx.x - o] : 0u;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned base = part[threadIdx.x] - sum;
  for (unsigned b = lo; b < hi; b++) { block_base[b] = base; base += ((volatile unsigned *)block_count)[b]; }
  if (threadIdx.x == blockDim.x - 1) { *n_done = (int)part[threadIdx.x]; *ticket = 0u; }     // n_done may be mapped host memory
}

__global__ void brb_done_rows_kernel(long long n, const uint8_t *__restrict__ done, const uint8_t *__restrict__ truncated,
                                     const float *__restrict__ terminal_obs, const float *__restrict__ ep_return,
                                     const int32_t *__restrict__ ep_len, const unsigned *__restrict__ block_base, uint32_t *__restrict__ rows,
                                     long long max_rows) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool d = i < n && done[i] != 0;
  __shared__ unsigned wcount[8];
  const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xFFFFFFFFu, d);
  if (lane == 0) wcount[w] = __popc(bal);
  __syncthreads();
  if (!d) return;
  unsigned r = block_base[blockIdx.x] + __popc(bal & ((1u << lane) - 1u));
  for (unsigned k = 0; k < w; k++) r += wcount[k];
  if ((long long)r >= max_rows) return;     // rows may be the caller's (mapped, pinned) host buffer: never write past its capacity
  uint32_t *row = rows + (size_t)r * BRB_DONE_ROW_WORDS;
  row[0] = (uint32_t)i;
#pragma unroll
  for (int k = 0; k < 6; k++) row[1 + k] = __float_as_uint(terminal_obs[i * 6 + k]);
  row[7] = __float_as_uint(ep_return[i]);
  row[8] = (uint32_t)ep_len[i];
  row[9] = (uint32_t)truncated[i];
}

template <int KIND>
__global__ void brb_reset_kernel(const BrbState S, float *__restrict__ obs, const double *__restrict__ replay_u, const unsigned epoch) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.n) return;
  S.event[i] = 0u;
  float o[6];
  if (KIND == BRB_ENV03_V2) {
    double ur[32];
    if (replay_u) { for (int k = 0; k < 32; k++) ur[k] = replay_u[i * 32 + k]; }
    else { for (int b = 0; b < 8; b++) draw4(S.seed, (uint64_t)(S.env0 + i), 0u, 1u + b + 16u * epoch, ur + 4 * b); }
    reset_env03(S, i, ur, o);
  } else {
    double ur[16];
    if (replay_u) {
#pragma unroll
      for (int k = 0; k < 16; k++) ur[k] = replay_u[i * 16 + k];
    } else {
#pragma unroll
      for (int b = 0; b < 4; b++) draw4(S.seed, (uint64_t)(S.env0 + i), 0u, 1u + b + 16u * epoch, ur + 4 * b);
    }
    reset_env<KIND>(S, i, ur, o);
  }
#pragma unroll
  for (int k = 0; k < 6; k++) obs[i * 6 + k] = o[k];
}

__global__ void brb_get_state_kernel(const BrbState S, double *qpos, double *qvel, double *xquat) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.n) return;
  if (qpos) for (int k = 0; k < S.nq; k++) qpos[i * S.nq + k] = S.qpos[k * S.n + i];
  if (qvel) for (int k = 0; k < S.nv; k++) qvel[i * S.nv + k] = S.qvel[k * S.n + i];
  if (xquat) for (int k = 0; k < 4; k++) xquat[i * 4 + k] = S.xquat[k * S.n + i];
}

// MujocoEnv.set_state + mj_forward: kinematics become fresh, warm start restarts
__global__ void brb_set_state_kernel(const BrbState S, const double *qpos, const double *qvel) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.n) return;
  for (int k = 0; k < S.nq; k++) S.qpos[k * S.n + i] = qpos[i * S.nq + k];
  for (int k = 0; k < S.nv; k++) S.qvel[k * S.n + i] = qvel[i * S.nv + k];
  S.aset[i] = 0xFFFFu;
  double nn = 0;
  for (int k = 0; k < 4; k++) nn += qpos[i * S.nq + 3 + k] * qpos[i * S.nq + 3 + k];
  nn = 1.0 / sqrt(nn);
  for (int k = 0; k < 4; k++) S.xquat[k * S.n + i] = qpos[i * S.nq + 3 + k] * nn;
}

__global__ void brb_get_elapsed_kernel(const BrbState S, int32_t *out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < S.n) out[i] = S.elapsed[i];
}

// FFMA-bound probe for the FP32-pipe roofline denominator: 8 independent chains per thread
__global__ void brb_ffma_probe_kernel(float *out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int k = 0; k < iters; k++) {
#pragma unroll
    for (int j = 0; j < 16; j++) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// ---------------------------------------------------------------------------------------------------
// launch wrappers used by the C-ABI translation unit (brb_cabi.cu)
extern "C" int brb_step_resident_ctas(int kind, int device) {
  int per_sm = 0, sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaError_t e = cudaSuccess;
  switch (kind) {
    case BRB_ENV01_V1: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, brb_step_kernel<BRB_ENV01_V1>, BRB_BLOCK, 0); break;
    case BRB_ENV01_V2: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, brb_step_kernel<BRB_ENV01_V2>, BRB_BLOCK, 0); break;
    case BRB_ENV01_V3: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, brb_step_kernel<BRB_ENV01_V3>, BRB_BLOCK, 0); break;
    default: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, brb_step_kernel<BRB_ENV03_V2>, BRB_BLOCK_ENV03, 0); break;
  }
  if (e != cudaSuccess || per_sm < 1 || sms < 1) { cudaGetLastError(); return 0; }
  return per_sm * sms;
}

extern "C" void brb_launch_step(int kind, const BrbModelConsts *c, const BrbState *S, const BrbPerm *perm, const float *actions, float *obs, float *reward,
                                uint8_t *done, uint8_t *truncated, float *terminal_obs, float *ep_return, int32_t *ep_len,
                                const double *replay_u, int max_ctas, cudaStream_t stream) {
  unsigned grid = (unsigned)((S->n + BRB_BLOCK - 1) / BRB_BLOCK);
  if (perm->cursor != nullptr && kind != BRB_ENV03_V2 && grid > (unsigned)max_ctas) grid = (unsigned)max_ctas;   // one resident wave, warps pull work
  switch (kind) {
    case BRB_ENV01_V1:
      brb_step_kernel<BRB_ENV01_V1><<<grid, BRB_BLOCK, 0, stream>>>(*c, *S, *perm, actions, obs, reward, done, truncated, terminal_obs, ep_return, ep_len, replay_u);
      break;
    case BRB_ENV01_V2:
      brb_step_kernel<BRB_ENV01_V2><<<grid, BRB_BLOCK, 0, stream>>>(*c, *S, *perm, actions, obs, reward, done, truncated, terminal_obs, ep_return, ep_len, replay_u);
      break;
    case BRB_ENV01_V3:
      brb_step_kernel<BRB_ENV01_V3><<<grid, BRB_BLOCK, 0, stream>>>(*c, *S, *perm, actions, obs, reward, done, truncated, terminal_obs, ep_return, ep_len, replay_u);
      break;
    default:
      if (c->flags & BRB_FLAG_WHEEL_BLOCK)
        brb_step_kernel<BRB_ENV03_V2_WB><<<(unsigned)((S->n + BRB_BLOCK_ENV03 - 1) / BRB_BLOCK_ENV03), BRB_BLOCK_ENV03, 0, stream>>>(*c, *S, *perm, actions, obs, reward, done, truncated, terminal_obs, ep_return, ep_len, replay_u);
      else
        brb_step_kernel<BRB_ENV03_V2><<<(unsigned)((S->n + BRB_BLOCK_ENV03 - 1) / BRB_BLOCK_ENV03), BRB_BLOCK_ENV03, 0, stream>>>(*c, *S, *perm, actions, obs, reward, done, truncated, terminal_obs, ep_return, ep_len, replay_u);
      break;
  }
}

#ifdef BRB_TIMELINE
extern "C" void brb_timeline(unsigned long long *out, int reset) {
  cudaDeviceSynchronize();
  if (out) { cudaMemcpyFromSymbol(out, g_timeline, sizeof(g_timeline)); cudaMemcpyFromSymbol(out + 1 + 4 * 65536, g_tl_trips, sizeof(g_tl_trips)); }
  if (reset) { unsigned long long z = 0; cudaMemcpyToSymbol(g_timeline, &z, sizeof(z)); }
}
#endif
#ifdef BRB_TRIPSTATS
extern "C" void brb_tripstats(unsigned long long *out) { cudaMemcpyFromSymbol(out, g_trip, sizeof(unsigned long long) * (48 + 128 + 18)); }
#endif

extern "C" void brb_launch_group(long long n, const uint8_t *key, const unsigned *hist, unsigned *cursor, int *order, unsigned *hist_zero,
                                 unsigned *cursor_zero, unsigned *queue_cursor, cudaStream_t stream) {
  brb_group_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, key, hist, cursor, order, hist_zero, cursor_zero, queue_cursor);
}

extern "C" void brb_launch_done_rows(long long n, const uint8_t *done, const uint8_t *truncated, const float *terminal_obs, const float *ep_return,
                                     const int32_t *ep_len, unsigned *block_count, unsigned *block_base, unsigned *ticket, int *n_done,
                                     uint32_t *rows, long long max_rows, cudaStream_t stream) {
  const unsigned nb = (unsigned)((n + 255) / 256);
  brb_done_count_kernel<<<nb, 256, 0, stream>>>(n, done, block_count, block_base, ticket, n_done);
  brb_done_rows_kernel<<<nb, 256, 0, stream>>>(n, done, truncated, terminal_obs, ep_return, ep_len, block_base, rows, max_rows);
}

// epoch = number of earlier reset_all calls on this env object: VecEnv.reset() called again draws NEW start states (block index
// 1 + b + 16 epoch of event 0), as the reference's RNG streams keep advancing across resets; epoch 0 is what the parity tests replay
extern "C" void brb_launch_reset(int kind, const BrbState *S, float *obs, const double *replay_u, unsigned epoch, cudaStream_t stream) {
  const unsigned grid = (unsigned)((S->n + 127) / 128);
  switch (kind) {
    case BRB_ENV01_V1: brb_reset_kernel<BRB_ENV01_V1><<<grid, 128, 0, stream>>>(*S, obs, replay_u, epoch); break;
    case BRB_ENV01_V2: brb_reset_kernel<BRB_ENV01_V2><<<grid, 128, 0, stream>>>(*S, obs, replay_u, epoch); break;
    case BRB_ENV01_V3: brb_reset_kernel<BRB_ENV01_V3><<<grid, 128, 0, stream>>>(*S, obs, replay_u, epoch); break;
    default: brb_reset_kernel<BRB_ENV03_V2><<<grid, 128, 0, stream>>>(*S, obs, replay_u, epoch); break;
  }
}

extern "C" void brb_launch_get_state(const BrbState *S, double *qpos, double *qvel, double *xquat, cudaStream_t stream) {
  brb_get_state_kernel<<<(unsigned)((S->n + 127) / 128), 128, 0, stream>>>(*S, qpos, qvel, xquat);
}
extern "C" void brb_launch_set_state(const BrbState *S, const double *qpos, const double *qvel, cudaStream_t stream) {
  brb_set_state_kernel<<<(unsigned)((S->n + 127) / 128), 128, 0, stream>>>(*S, qpos, qvel);
}
extern "C" void brb_launch_get_elapsed(const BrbState *S, int32_t *out, cudaStream_t stream) {
  brb_get_elapsed_kernel<<<(unsigned)((S->n + 127) / 128), 128, 0, stream>>>(*S, out);
}
extern "C" void brb_launch_ffma_probe(float *out, int blocks, int threads, int iters, cudaStream_t stream) {
  brb_ffma_probe_kernel<<<blocks, threads, 0, stream>>>(out, iters, 0.999999f, 1e-7f);
}
#endif  // !BRB_HOST_EMU
