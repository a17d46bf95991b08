// emu.cpp — TEST-ONLY host harness around the CUDA kernels' per-env functions (compiled with -DBRB_HOST_EMU).
// Not part of the product; never shipped in libbrb_cuda.so.
#include <stdlib.h>
#include <string.h>
#include "../../balance_robot_b200/csrc/brb_kernels.cu"

struct Emu {
  BrbModelConsts c;
  BrbState S;
  double *tt;
  unsigned long long stats[BRB_NSTATS];
};

extern "C" Emu *emu_create(const BrbModelConsts *c, const double *tt, int n_time, long long n, unsigned long long seed, long long env0) {
  Emu *e = (Emu *)calloc(1, sizeof(Emu));
  e->c = *c;
  e->tt = (double *)malloc(sizeof(double) * n_time);
  memcpy(e->tt, tt, sizeof(double) * n_time);
  BrbState &S = e->S;
  S.n = n; S.env0 = env0; S.seed = seed; S.nq = c->nq; S.nv = c->nv;
  S.qpos = (double *)calloc((size_t)c->nq * n, 8); S.qvel = (double *)calloc((size_t)c->nv * n, 8); S.xquat = (double *)calloc(4 * n, 8);
  S.aset = (uint32_t *)calloc(n, 4); S.last_pitch = (double *)calloc(n, 8); S.ep_return = (double *)calloc(n, 8);
  S.v3 = (double *)calloc(3 * n, 8); S.elapsed = (int *)calloc(n, 4); S.ep_len = (int *)calloc(n, 4);
  S.event = (uint32_t *)calloc(n, 4); S.time_table = e->tt; S.stats = e->stats;
  return e;
}
extern "C" void emu_destroy(Emu *e) {
  BrbState &S = e->S;
  free(S.qpos); free(S.qvel); free(S.xquat); free(S.aset); free(S.last_pitch); free(S.ep_return); free(S.v3);
  free(S.elapsed); free(S.ep_len); free(S.event); free(e->tt); free(e);
}
template <int KIND> static void reset_all(Emu *e, float *obs, const double *replay) {
  for (long long i = 0; i < e->S.n; i++) {
    e->S.event[i] = 0;
    if (KIND == BRB_ENV03_V2) {
      double ur[32];
      if (replay) memcpy(ur, replay + 32 * i, sizeof ur);
      else for (int b = 0; b < 8; b++) draw4(e->S.seed, (uint64_t)(e->S.env0 + i), 0u, 1u + b, ur + 4 * b);
      reset_env03(e->S, i, ur, obs + 6 * i);
    } else {
      double ur[16];
      if (replay) memcpy(ur, replay + 16 * i, sizeof ur);
      else for (int b = 0; b < 4; b++) draw4(e->S.seed, (uint64_t)(e->S.env0 + i), 0u, 1u + b, ur + 4 * b);
      reset_env<KIND>(e->S, i, ur, obs + 6 * i);
    }
  }
}
extern "C" void emu_reset(Emu *e, float *obs, const double *replay) {
  switch (e->c.env_kind) {
    case BRB_ENV01_V1: reset_all<BRB_ENV01_V1>(e, obs, replay); break;
    case BRB_ENV01_V2: reset_all<BRB_ENV01_V2>(e, obs, replay); break;
    case BRB_ENV01_V3: reset_all<BRB_ENV01_V3>(e, obs, replay); break;
    default: reset_all<BRB_ENV03_V2>(e, obs, replay); break;
  }
}
template <int KIND> static void step_all(Emu *e, const float *actions, float *obs, float *reward, uint8_t *done, uint8_t *trunc,
                                         float *tobs, float *epr, int32_t *epl, const double *replay) {
  for (int k = 0; k < 16 * 5; k++) stab_fill(e->c, KIND == BRB_ENV03_V2, k);     // the kernel's per-CTA shared-memory table
  for (long long i = 0; i < e->S.n; i++) {
    unsigned stat[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (KIND == BRB_ENV03_V2) {
      if (e->c.flags & BRB_FLAG_WHEEL_BLOCK) step_env03<true>(e->c, e->S, i, actions, obs, reward, done, trunc, tobs, epr, epl, replay, stat, 1u, false);
      else step_env03<false>(e->c, e->S, i, actions, obs, reward, done, trunc, tobs, epr, epl, replay, stat, 1u, false);
    }
    else step_env<KIND>(e->c, e->S, i, actions, obs, reward, done, trunc, tobs, epr, epl, replay, stat, 1u);
    for (int k = 0; k < 6; k++) e->stats[k] += stat[k];
    e->stats[BRB_STAT_CONTACT_SLOTS] += stat[7];
    e->stats[BRB_STAT_COUPLED_SUBSTEPS] += stat[8];
    e->stats[BRB_STAT_BLOCK_CONTACT_SUBSTEPS] += stat[9];
    e->stats[10] += stat[10];
    e->stats[11] += stat[11];
    e->stats[BRB_STAT_ENV_STEPS] += 1;
  }
}
extern "C" void emu_step(Emu *e, const float *actions, float *obs, float *reward, uint8_t *done, uint8_t *trunc, float *tobs,
                         float *epr, int32_t *epl, const double *replay) {
  switch (e->c.env_kind) {
    case BRB_ENV01_V1: step_all<BRB_ENV01_V1>(e, actions, obs, reward, done, trunc, tobs, epr, epl, replay); break;
    case BRB_ENV01_V2: step_all<BRB_ENV01_V2>(e, actions, obs, reward, done, trunc, tobs, epr, epl, replay); break;
    case BRB_ENV01_V3: step_all<BRB_ENV01_V3>(e, actions, obs, reward, done, trunc, tobs, epr, epl, replay); break;
    default: step_all<BRB_ENV03_V2>(e, actions, obs, reward, done, trunc, tobs, epr, epl, replay); break;
  }
}
extern "C" void emu_get_state(Emu *e, double *qpos, double *qvel, double *xquat) {
  const long long n = e->S.n;
  for (long long i = 0; i < n; i++) {
    const int nq = e->S.nq, nv = e->S.nv;
    for (int k = 0; k < nq; k++) qpos[i * nq + k] = e->S.qpos[k * n + i];
    for (int k = 0; k < nv; k++) qvel[i * nv + k] = e->S.qvel[k * n + i];
    if (xquat) for (int k = 0; k < 4; k++) xquat[i * 4 + k] = e->S.xquat[k * n + i];
  }
}
extern "C" void emu_set_state(Emu *e, const double *qpos, const double *qvel) {
  const long long n = e->S.n;
  for (long long i = 0; i < n; i++) {
    const int nq = e->S.nq, nv = e->S.nv;
    for (int k = 0; k < nq; k++) e->S.qpos[k * n + i] = qpos[i * nq + k];
    for (int k = 0; k < nv; k++) e->S.qvel[k * n + i] = qvel[i * nv + k];
    e->S.aset[i] = 0xFFFFu;
    double nn = 0;
    for (int k = 0; k < 4; k++) nn += qpos[i * nq + 3 + k] * qpos[i * nq + 3 + k];
    nn = 1.0 / sqrt(nn);
    for (int k = 0; k < 4; k++) e->S.xquat[k * n + i] = qpos[i * nq + 3 + k] * nn;
  }
}
extern "C" void emu_get_stats(Emu *e, unsigned long long *out) { memcpy(out, e->stats, sizeof e->stats); }

// test hook: the fp32 cylinder-box collider of brb_env03.cuh on its own (tests/test_env03_parity.py compares it with the oracle's)
extern "C" int emu_cyl_box(const float *c, const float *a, float R, float L, const float *b, const float *E_rows, float h, float margin,
                           float *dist, float *nrm, float *pos) {
  CylBoxIn ci;
  for (int j = 0; j < 3; j++) { ci.cc[j] = c[j]; ci.a[j] = a[j]; ci.b[j] = b[j]; ci.E[0][j] = E_rows[j]; ci.E[1][j] = E_rows[3 + j]; ci.E[2][j] = E_rows[6 + j]; }
  ci.R = R; ci.L = L; ci.h = h; ci.margin = margin;
  const CylBoxOut co = env03_cyl_box(ci);
  *dist = co.dist;
  for (int j = 0; j < 3; j++) { nrm[j] = co.n[j]; pos[j] = co.pos[j]; }
  return co.hit;
}
