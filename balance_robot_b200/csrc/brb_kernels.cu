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
#else
#include <cuda_runtime.h>
#define BRB_D __device__ __forceinline__
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
BRB_D double pitch_of(const double q[4]) {  // RobotBaseEnv.py:127-135
  if (q[0] == 0.0) return 0.0;
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
  return atan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y));
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
  for (int k = 0; k < 8; k++) { S.qvel[k * N + i] = 0.0; S.warm[k * N + i] = 0.f; }
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
// packed lower-triangular index
#define LT(i, j) ((i) * ((i) + 1) / 2 + (j))

struct Contact { float rx, ry, rz, y0, y1, y2; };

// One physics substep.  All arguments are registers of the calling thread.
struct Sub {
  KF p[3], q[4], th[2], v[3], w[3], s[2];   // world pos, quat (w,x,y,z), wheel angles, world lin vel, body ang vel, wheel speeds
  float a[8];                               // body-frame solver acceleration (warm start for the next substep)
  float uhi[2], ulo[2];                     // clamped ctrl targets, split hi/lo
  unsigned n_contact, n_solve, n_nonconv;
};

template <int MAXIT>
BRB_D void substep(const BrbModelConsts &c, Sub &st) {
  const float qw = st.q[0].s, qx = st.q[1].s, qy = st.q[2].s, qz = st.q[3].s;
  // rows of R = world axes expressed in the chassis frame
  const float xx = qx * qx, yy = qy * qy, zz = qz * qz, xy = qx * qy, xz = qx * qz, yz = qy * qz, wx = qw * qx, wy = qw * qy, wz = qw * qz;
  const float Xb0 = 1.f - 2.f * (yy + zz), Xb1 = 2.f * (xy - wz), Xb2 = 2.f * (xz + wy);
  const float Yb0 = 2.f * (xy + wz), Yb1 = 1.f - 2.f * (xx + zz), Yb2 = 2.f * (yz - wx);
  const float Zb0 = 2.f * (xz - wy), Zb1 = 2.f * (yz + wx), Zb2 = 1.f - 2.f * (xx + yy);
  const float w0 = st.w[0].s, w1 = st.w[1].s, w2 = st.w[2].s, sL = st.s[0].s, sR = st.s[1].s;

  // ---- smooth forces in the chassis frame (A.3 steps 4-6) ----
  float f[8];
  {
    const float mg = c.mass * c.grav, gm = c.grav * c.mcz;
    f[0] = -c.mcz * (w0 * w2) - mg * Zb0;
    f[1] = -c.mcz * (w1 * w2) - mg * Zb1;
    f[2] = c.mcz * (w0 * w0 + w1 * w1) - mg * Zb2;
    const float Lx = c.Ixx * w0 + c.Ia * (sR - sL), Ly = c.Iyy * w1, Lz = c.Izz * w2;
    f[3] = -(w1 * Lz - w2 * Ly) + gm * Zb1;
    f[4] = -(w2 * Lx - w0 * Lz) - gm * Zb0;
    f[5] = -(w0 * Ly - w1 * Lx);
  }
  bool clampL, clampR;
  {
    // servo: force = clip(kv (u - s), forcerange); u - s evaluated with the hi/lo halves (Q7)
    float dL = (st.uhi[0] - sL) + (st.ulo[0] + st.s[0].c), dR = (st.uhi[1] - sR) + (st.ulo[1] + st.s[1].c);
    float tL = c.kv * dL, tR = c.kv * dR;
    clampL = (tL <= c.frc_lo) || (tL >= c.frc_hi);
    clampR = (tR <= c.frc_lo) || (tR >= c.frc_hi);
    tL = fminf(c.frc_hi, fmaxf(c.frc_lo, tL));
    tR = fminf(c.frc_hi, fmaxf(c.frc_lo, tR));
    f[6] = tL - c.damping * sL;
    f[7] = tR - c.damping * sR;
  }

  // ---- plane-cylinder collision in the chassis frame (A.6) ----
  Contact ct[4];
  unsigned valid = 0;
  {
    const float rho2 = Zb1 * Zb1 + Zb2 * Zb2;
    const float irho = rsqrtf(fmaxf(rho2, 1e-30f));
    const float rho = rho2 * irho;
    const float vy = -c.rad * Zb1 * irho, vz = -c.rad * Zb2 * irho;
    // oz*nz - rad*rho without cancellation when upright: (oz-rad) nz + rad (nz - rho), nz - rho = -ny^2/(nz+rho)
    const float diff = (Zb2 > 0.f) ? -(Zb1 * Zb1) / (Zb2 + rho) : (Zb2 - rho);
    const float hgt = ((st.p[2].s - c.zfloor) - st.p[2].c) - c.zfloor_lo;
    const float common = hgt + (c.oz - c.rad) * Zb2 + c.rad * diff;
    const float anx = fabsf(Zb0);
    const float sa = (Zb0 > 0.f) ? -1.f : 1.f;
    const float u0 = st.v[0].s * Xb0 + st.v[1].s * Yb0 + st.v[2].s * Zb0;   // chassis-frame linear velocity
    const float u1 = st.v[0].s * Xb1 + st.v[1].s * Yb1 + st.v[2].s * Zb1;
    const float u2 = st.v[0].s * Xb2 + st.v[1].s * Yb2 + st.v[2].s * Zb2;
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const float sg = k ? 1.f : -1.f;
      const float d0 = common + sg * c.ox * Zb0;
      const float sk = k ? sR : sL;
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const float dist = e ? d0 + c.hl * anx : d0 - c.hl * anx;
        if (dist < 0.f) {
          const int ci = 2 * k + e;
          valid |= 1u << ci;
          const float hd = 0.5f * dist;
          const float rx = sg * c.ox + (e ? -sa : sa) * c.hl - Zb0 * hd;
          const float ry = vy - Zb1 * hd;
          const float rz = c.oz + vz - Zb2 * hd;
          // material-point velocity: u + w x r + s_k * sg * (0, -(rz-oz), ry)
          const float wy_ = -sg * (rz - c.oz), wz_ = sg * ry;
          const float px = u0 + w1 * rz - w2 * ry;
          const float py = u1 + w2 * rx - w0 * rz + sk * wy_;
          const float pz = u2 + w0 * ry - w1 * rx + sk * wz_;
          ct[ci].rx = rx; ct[ci].ry = ry; ct[ci].rz = rz;
          ct[ci].y0 = c.Bdamp * (Zb0 * px + Zb1 * py + Zb2 * pz) + c.Kimp * dist;
          ct[ci].y1 = c.Bdamp * (Yb0 * px + Yb1 * py + Yb2 * pz);
          ct[ci].y2 = -c.Bdamp * (Xb0 * px + Xb1 * py + Xb2 * pz);
        }
      }
    }
  }

  float a[8];
  if (valid == 0) {
    // ---- free flight: a = M_b^-1 f ----
    a[0] = c.minv_xy[0] * f[0] + c.minv_xy[1] * f[4];
    a[4] = c.minv_xy[1] * f[0] + c.minv_xy[2] * f[4];
    a[2] = c.minv_uz * f[2];
    a[5] = c.minv_wz * f[5];
    a[1] = c.minv_blk[0] * f[1] + c.minv_blk[1] * f[3] + c.minv_blk[2] * f[6] + c.minv_blk[3] * f[7];
    a[3] = c.minv_blk[1] * f[1] + c.minv_blk[4] * f[3] + c.minv_blk[5] * f[6] + c.minv_blk[6] * f[7];
    a[6] = c.minv_blk[2] * f[1] + c.minv_blk[5] * f[3] + c.minv_blk[7] * f[6] + c.minv_blk[8] * f[7];
    a[7] = c.minv_blk[3] * f[1] + c.minv_blk[6] * f[3] + c.minv_blk[8] * f[6] + c.minv_blk[9] * f[7];
  } else {
    st.n_contact++;
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = st.a[k];
    unsigned used = 0xFFFFFFFFu;
    const float mu = c.mu;
    int it = 0;
    for (;; it++) {
      // ---- active set at the current acceleration: z_c = B (P_c a) + yhat_c, rows E_r . z_c < 0 ----
      unsigned bits = 0;
#pragma unroll
      for (int ci = 0; ci < 4; ci++) {
        if (valid & (1u << ci)) {
          const float sg = (ci >> 1) ? 1.f : -1.f;
          const float ak = a[6 + (ci >> 1)];
          const float rx = ct[ci].rx, ry = ct[ci].ry, rz = ct[ci].rz;
          const float px = a[0] + a[4] * rz - a[5] * ry;
          const float py = a[1] + a[5] * rx - a[3] * rz - ak * sg * (rz - c.oz);
          const float pz = a[2] + a[3] * ry - a[4] * rx + ak * sg * ry;
          const float z0 = Zb0 * px + Zb1 * py + Zb2 * pz + ct[ci].y0;
          const float z1 = mu * (Yb0 * px + Yb1 * py + Yb2 * pz + ct[ci].y1);
          const float z2 = mu * (ct[ci].y2 - (Xb0 * px + Xb1 * py + Xb2 * pz));
          bits |= (unsigned)(z0 + z1 < 0.f) << (4 * ci);
          bits |= (unsigned)(z0 - z1 < 0.f) << (4 * ci + 1);
          bits |= (unsigned)(z0 + z2 < 0.f) << (4 * ci + 2);
          bits |= (unsigned)(z0 - z2 < 0.f) << (4 * ci + 3);
        }
      }
      if (bits == used) break;
      if (it >= MAXIT) { st.n_nonconv++; break; }
      used = bits;
      st.n_solve++;

      // ---- H = M_b + sum_c P_c' S_c P_c (packed lower), rhs = f - sum_c P_c' B' W_c yhat_c ----
      float H[36];
#pragma unroll
      for (int k = 0; k < 36; k++) H[k] = 0.f;
      H[LT(0, 0)] = c.mass; H[LT(1, 1)] = c.mass; H[LT(2, 2)] = c.mass;
      H[LT(4, 0)] = c.mcz; H[LT(3, 1)] = -c.mcz;
      H[LT(3, 3)] = c.Ixx; H[LT(4, 4)] = c.Iyy; H[LT(5, 5)] = c.Izz;
      H[LT(6, 3)] = -c.Ia; H[LT(7, 3)] = c.Ia; H[LT(6, 6)] = c.Ia; H[LT(7, 7)] = c.Ia;
      float r[8];
#pragma unroll
      for (int k = 0; k < 8; k++) r[k] = f[k];
#pragma unroll
      for (int ci = 0; ci < 4; ci++) {
        const unsigned b = (bits >> (4 * ci)) & 15u;
        if ((valid & (1u << ci)) && b) {
          const float b0 = (float)(b & 1u), b1 = (float)((b >> 1) & 1u), b2 = (float)((b >> 2) & 1u), b3 = (float)((b >> 3) & 1u);
          const float Dm = c.D * mu, Dmm = Dm * mu;
          const float W00 = c.D * (b0 + b1 + b2 + b3), W01 = Dm * (b0 - b1), W02 = Dm * (b2 - b3);
          const float W11 = Dmm * (b0 + b1), W22 = Dmm * (b2 + b3);
          // rows of W B with B = [Zb; Yb; -Xb]
          const float g00 = W00 * Zb0 + W01 * Yb0 - W02 * Xb0, g01 = W00 * Zb1 + W01 * Yb1 - W02 * Xb1, g02 = W00 * Zb2 + W01 * Yb2 - W02 * Xb2;
          const float g10 = W01 * Zb0 + W11 * Yb0, g11 = W01 * Zb1 + W11 * Yb1, g12 = W01 * Zb2 + W11 * Yb2;
          const float g20 = W02 * Zb0 - W22 * Xb0, g21 = W02 * Zb1 - W22 * Xb1, g22 = W02 * Zb2 - W22 * Xb2;
          // S = B' (W B), symmetric
          const float S00 = Zb0 * g00 + Yb0 * g10 - Xb0 * g20, S01 = Zb0 * g01 + Yb0 * g11 - Xb0 * g21, S02 = Zb0 * g02 + Yb0 * g12 - Xb0 * g22;
          const float S11 = Zb1 * g01 + Yb1 * g11 - Xb1 * g21, S12 = Zb1 * g02 + Yb1 * g12 - Xb1 * g22;
          const float S22 = Zb2 * g02 + Yb2 * g12 - Xb2 * g22;
          const float rx = ct[ci].rx, ry = ct[ci].ry, rz = ct[ci].rz;
          const float sg = (ci >> 1) ? 1.f : -1.f;
          const float wy_ = -sg * (rz - c.oz), wz_ = sg * ry;
          // columns of P: e_x e_y e_z | cx=(0,-rz,ry) cy=(rz,0,-rx) cz=(-ry,rx,0) | w=(0,wy_,wz_)
          // T_j = S p_j for the angular and wheel columns
          const float T3x = -rz * S01 + ry * S02, T3y = -rz * S11 + ry * S12, T3z = -rz * S12 + ry * S22;
          const float T4x = rz * S00 - rx * S02, T4y = rz * S01 - rx * S12, T4z = rz * S02 - rx * S22;
          const float T5x = -ry * S00 + rx * S01, T5y = -ry * S01 + rx * S11, T5z = -ry * S02 + rx * S12;
          const float Twx = wy_ * S01 + wz_ * S02, Twy = wy_ * S11 + wz_ * S12, Twz = wy_ * S12 + wz_ * S22;
          H[LT(0, 0)] += S00; H[LT(1, 0)] += S01; H[LT(2, 0)] += S02; H[LT(1, 1)] += S11; H[LT(2, 1)] += S12; H[LT(2, 2)] += S22;
          H[LT(3, 0)] += T3x; H[LT(3, 1)] += T3y; H[LT(3, 2)] += T3z;
          H[LT(4, 0)] += T4x; H[LT(4, 1)] += T4y; H[LT(4, 2)] += T4z;
          H[LT(5, 0)] += T5x; H[LT(5, 1)] += T5y; H[LT(5, 2)] += T5z;
          H[LT(3, 3)] += -rz * T3y + ry * T3z;
          H[LT(4, 3)] += rz * T3x - rx * T3z;
          H[LT(5, 3)] += -ry * T3x + rx * T3y;
          H[LT(4, 4)] += rz * T4x - rx * T4z;
          H[LT(5, 4)] += -ry * T4x + rx * T4y;
          H[LT(5, 5)] += -ry * T5x + rx * T5y;
          const int kw = 6 + (ci >> 1);
          H[LT(kw, 0)] += Twx; H[LT(kw, 1)] += Twy; H[LT(kw, 2)] += Twz;
          H[LT(kw, 3)] += -rz * Twy + ry * Twz;
          H[LT(kw, 4)] += rz * Twx - rx * Twz;
          H[LT(kw, 5)] += -ry * Twx + rx * Twy;
          H[LT(kw, kw)] += wy_ * Twy + wz_ * Twz;
          // rhs: g = -B' (W yhat)
          const float t0 = W00 * ct[ci].y0 + W01 * ct[ci].y1 + W02 * ct[ci].y2;
          const float t1 = W01 * ct[ci].y0 + W11 * ct[ci].y1;
          const float t2 = W02 * ct[ci].y0 + W22 * ct[ci].y2;
          const float gx = -(t0 * Zb0 + t1 * Yb0 - t2 * Xb0), gy = -(t0 * Zb1 + t1 * Yb1 - t2 * Xb1), gz = -(t0 * Zb2 + t1 * Yb2 - t2 * Xb2);
          r[0] += gx; r[1] += gy; r[2] += gz;
          r[3] += -rz * gy + ry * gz;
          r[4] += rz * gx - rx * gz;
          r[5] += -ry * gx + rx * gy;
          r[kw] += wy_ * gy + wz_ * gz;
        }
      }
      // ---- Cholesky H = L L' in registers, then solve ----
      float inv[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        float d = H[LT(j, j)];
#pragma unroll
        for (int k = 0; k < j; k++) d -= H[LT(j, k)] * H[LT(j, k)];
        const float id = rsqrtf(d);
        inv[j] = id;
#pragma unroll
        for (int i = j + 1; i < 8; i++) {
          float t = H[LT(i, j)];
#pragma unroll
          for (int k = 0; k < j; k++) t -= H[LT(i, k)] * H[LT(j, k)];
          H[LT(i, j)] = t * id;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; i++) {
        float t = r[i];
#pragma unroll
        for (int k = 0; k < i; k++) t -= H[LT(i, k)] * r[k];
        r[i] = t * inv[i];
      }
#pragma unroll
      for (int i = 7; i >= 0; i--) {
        float t = r[i];
#pragma unroll
        for (int k = i + 1; k < 8; k++) t -= H[LT(k, i)] * r[k];
        r[i] = t * inv[i];
      }
#pragma unroll
      for (int k = 0; k < 8; k++) a[k] = r[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; k++) st.a[k] = a[k];   // qacc_warmstart <- solver qacc (A.3 step 8)

  // ---- implicitfast (A.9): a+ = (M + h Dv)^-1 M a = a - Wm (Cinv + G)^-1 [a6; a7] ----
  {
    const bool skip = (c.flags & BRB_FLAG_ACTDERIV_SKIP_CLAMPED) != 0;
    const float cL = (skip && clampL) ? c.impl_cinv_damp : c.impl_cinv_full;
    const float cR = (skip && clampR) ? c.impl_cinv_damp : c.impl_cinv_full;
    const float k00 = cL + c.impl_G[0], k01 = c.impl_G[1], k11 = cR + c.impl_G[2];
    const float idet = 1.f / (k00 * k11 - k01 * k01);
    const float y0 = (k11 * a[6] - k01 * a[7]) * idet, y1 = (k00 * a[7] - k01 * a[6]) * idet;
    a[1] -= c.impl_W[0] * y0 + c.impl_W[1] * y1;
    a[3] -= c.impl_W[2] * y0 + c.impl_W[3] * y1;
    a[6] -= c.impl_W[4] * y0 + c.impl_W[5] * y1;
    a[7] -= c.impl_W[6] * y0 + c.impl_W[7] * y1;
  }

  // ---- mj_advance (A.10): velocities, then positions with the NEW velocities ----
  const float h = c.h;
  kadd(st.v[0], h * (Xb0 * a[0] + Xb1 * a[1] + Xb2 * a[2]));
  kadd(st.v[1], h * (Yb0 * a[0] + Yb1 * a[1] + Yb2 * a[2]));
  kadd(st.v[2], h * (Zb0 * a[0] + Zb1 * a[1] + Zb2 * a[2]));
  kadd(st.w[0], h * a[3]); kadd(st.w[1], h * a[4]); kadd(st.w[2], h * a[5]);
  kadd(st.s[0], h * a[6]); kadd(st.s[1], h * a[7]);
  kadd(st.p[0], h * st.v[0].s - h * st.v[0].c);
  kadd(st.p[1], h * st.v[1].s - h * st.v[1].c);
  kadd(st.p[2], h * st.v[2].s - h * st.v[2].c);
  kadd(st.th[0], h * st.s[0].s); kadd(st.th[1], h * st.s[1].s);
  {
    // q <- q (x) [cos(t/2), sin(t/2) w/|w|], t = h |w|; added as the increment q (x) (dq - 1)
    const float n0 = st.w[0].s, n1 = st.w[1].s, n2 = st.w[2].s;
    const float t2 = (h * h) * (n0 * n0 + n1 * n1 + n2 * n2);
    const float sn = (0.5f * h) * (1.f - t2 * (1.f / 24.f));
    const float cm1 = -(0.125f * t2) * (1.f - t2 * (1.f / 48.f));
    const float ex = sn * n0, ey = sn * n1, ez = sn * n2;
    const float dw = qw * cm1 - (qx * ex + qy * ey + qz * ez);
    const float dx = qx * cm1 + (qw * ex + qy * ez - qz * ey);
    const float dy = qy * cm1 + (qw * ey - qx * ez + qz * ex);
    const float dz = qz * cm1 + (qw * ez + qx * ey - qy * ex);
    kadd(st.q[0], dw); kadd(st.q[1], dx); kadd(st.q[2], dy); kadd(st.q[3], dz);
  }
}

// ---------------------------------------------------------------------------------------------------
#ifndef BRB_HOST_EMU
__device__ __forceinline__ unsigned long long warp_sum(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}
#endif

// One VecEnv.step for env i of the shard.  stat[6] receives {substeps, contact substeps, solves, non-converged,
// unsupported-pose flag, done flag}.
template <int KIND>
BRB_D void step_env(const BrbModelConsts &c, const BrbState &S, const long long i, const float *__restrict__ actions,
                    float *__restrict__ obs, float *__restrict__ reward, uint8_t *__restrict__ done,
                    uint8_t *__restrict__ truncated, float *__restrict__ terminal_obs, float *__restrict__ ep_return_out,
                    int32_t *__restrict__ ep_len_out, const double *__restrict__ replay_u, unsigned stat[6]) {
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
  Sub st;
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
#pragma unroll
    for (int k = 0; k < 8; k++) st.a[k] = S.warm[k * N + i];
    st.n_contact = st.n_solve = st.n_nonconv = 0;
  }
  const int nsub = c.frame_skip;
  KF qprev[4];
  for (int sidx = 0; sidx < nsub; sidx++) {
    if (sidx == nsub - 1) {
#pragma unroll
      for (int k = 0; k < 4; k++) qprev[k] = st.q[k];   // Q1: kinematics seen by the task logic are one substep stale
    }
    substep<BRB_MAXIT>(c, st);
  }
  stat[0] = nsub; stat[1] = st.n_contact; stat[2] = st.n_solve; stat[3] = st.n_nonconv;

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
  const bool dn = terminated || trunc;
  if (reward) reward[i] = (float)rew;
  if (done) done[i] = (uint8_t)dn;
  if (truncated) truncated[i] = (uint8_t)(trunc && !terminated);
  if (ep_return_out) ep_return_out[i] = (float)epr;
  if (ep_len_out) ep_len_out[i] = epl;

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
  } else {
#pragma unroll
    for (int k = 0; k < 9; k++) S.qpos[k * N + i] = qpos[k];
#pragma unroll
    for (int k = 0; k < 8; k++) { S.qvel[k * N + i] = qvel[k]; S.warm[k * N + i] = st.a[k]; }
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

#ifndef BRB_HOST_EMU
template <int KIND>
__global__ void __launch_bounds__(BRB_BLOCK) brb_step_kernel(const __grid_constant__ BrbModelConsts c, const BrbState S,
                                                             const float *__restrict__ actions, float *__restrict__ obs,
                                                             float *__restrict__ reward, uint8_t *__restrict__ done,
                                                             uint8_t *__restrict__ truncated, float *__restrict__ terminal_obs,
                                                             float *__restrict__ ep_return_out, int32_t *__restrict__ ep_len_out,
                                                             const double *__restrict__ replay_u) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned stat[6] = {0, 0, 0, 0, 0, 0};
  if (i < S.n) step_env<KIND>(c, S, i, actions, obs, reward, done, truncated, terminal_obs, ep_return_out, ep_len_out, replay_u, stat);
  // statistics: one atomic per warp per counter
  const unsigned long long a0 = warp_sum(stat[0]), a1 = warp_sum(stat[1]), a2 = warp_sum(stat[2]), a3 = warp_sum(stat[3]),
                           a4 = warp_sum(stat[4]), a5 = warp_sum(stat[5]), a6 = warp_sum(i < S.n ? 1u : 0u);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&S.stats[BRB_STAT_SUBSTEPS], a0);
    atomicAdd(&S.stats[BRB_STAT_CONTACT_SUBSTEPS], a1);
    atomicAdd(&S.stats[BRB_STAT_SOLVES], a2);
    if (a3) atomicAdd(&S.stats[BRB_STAT_NONCONVERGED], a3);
    if (a4) atomicAdd(&S.stats[BRB_STAT_UNSUPPORTED], a4);
    if (a5) atomicAdd(&S.stats[BRB_STAT_EPISODES], a5);
    atomicAdd(&S.stats[BRB_STAT_ENV_STEPS], a6);
  }
}

template <int KIND>
__global__ void brb_reset_kernel(const BrbState S, float *__restrict__ obs, const double *__restrict__ replay_u) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.n) return;
  double ur[16];
  if (replay_u) {
#pragma unroll
    for (int k = 0; k < 16; k++) ur[k] = replay_u[i * 16 + k];
  } else {
#pragma unroll
    for (int b = 0; b < 4; b++) draw4(S.seed, (uint64_t)(S.env0 + i), 0u, 1u + b, ur + 4 * b);
  }
  S.event[i] = 0u;
  float o[6];
  reset_env<KIND>(S, i, ur, o);
#pragma unroll
  for (int k = 0; k < 6; k++) obs[i * 6 + k] = o[k];
}

__global__ void brb_get_state_kernel(const BrbState S, double *qpos, double *qvel, double *xquat) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.n) return;
  if (qpos) for (int k = 0; k < 9; k++) qpos[i * 9 + k] = S.qpos[k * S.n + i];
  if (qvel) for (int k = 0; k < 8; k++) qvel[i * 8 + k] = S.qvel[k * S.n + i];
  if (xquat) for (int k = 0; k < 4; k++) xquat[i * 4 + k] = S.xquat[k * S.n + i];
}

// MujocoEnv.set_state + mj_forward: kinematics become fresh, warm start restarts
__global__ void brb_set_state_kernel(const BrbState S, const double *qpos, const double *qvel) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.n) return;
  for (int k = 0; k < 9; k++) S.qpos[k * S.n + i] = qpos[i * 9 + k];
  for (int k = 0; k < 8; k++) { S.qvel[k * S.n + i] = qvel[i * 8 + k]; S.warm[k * S.n + i] = 0.f; }
  double nn = 0;
  for (int k = 0; k < 4; k++) nn += qpos[i * 9 + 3 + k] * qpos[i * 9 + 3 + k];
  nn = 1.0 / sqrt(nn);
  for (int k = 0; k < 4; k++) S.xquat[k * S.n + i] = qpos[i * 9 + 3 + k] * nn;
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
extern "C" void brb_launch_step(int kind, const BrbModelConsts *c, const BrbState *S, const float *actions, float *obs, float *reward,
                                uint8_t *done, uint8_t *truncated, float *terminal_obs, float *ep_return, int32_t *ep_len,
                                const double *replay_u, cudaStream_t stream) {
  const unsigned grid = (unsigned)((S->n + BRB_BLOCK - 1) / BRB_BLOCK);
  switch (kind) {
    case BRB_ENV01_V1:
      brb_step_kernel<BRB_ENV01_V1><<<grid, BRB_BLOCK, 0, stream>>>(*c, *S, actions, obs, reward, done, truncated, terminal_obs, ep_return, ep_len, replay_u);
      break;
    case BRB_ENV01_V2:
      brb_step_kernel<BRB_ENV01_V2><<<grid, BRB_BLOCK, 0, stream>>>(*c, *S, actions, obs, reward, done, truncated, terminal_obs, ep_return, ep_len, replay_u);
      break;
    default:
      brb_step_kernel<BRB_ENV01_V3><<<grid, BRB_BLOCK, 0, stream>>>(*c, *S, actions, obs, reward, done, truncated, terminal_obs, ep_return, ep_len, replay_u);
      break;
  }
}

extern "C" void brb_launch_reset(int kind, const BrbState *S, float *obs, const double *replay_u, cudaStream_t stream) {
  const unsigned grid = (unsigned)((S->n + 127) / 128);
  switch (kind) {
    case BRB_ENV01_V1: brb_reset_kernel<BRB_ENV01_V1><<<grid, 128, 0, stream>>>(*S, obs, replay_u); break;
    case BRB_ENV01_V2: brb_reset_kernel<BRB_ENV01_V2><<<grid, 128, 0, stream>>>(*S, obs, replay_u); break;
    default: brb_reset_kernel<BRB_ENV01_V3><<<grid, 128, 0, stream>>>(*S, obs, replay_u); break;
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
