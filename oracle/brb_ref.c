/*
 * brb_ref.c — fp64 CPU ORACLE (test infrastructure; see brb_ref.h for the usage rule and the
 * "parity unpinned" statement).
 *
 * Restates, for the model class {free-joint root bodies, hinge children, plane/box/cylinder geoms,
 * velocity servos, pyramidal condim-3 contacts, Newton solver, implicitfast}, what the reference
 * obtains from mujoco.mj_step (reference envs/env01_v1.py:24, envs/env01_v2.py:37) with the model
 * of envs/env01_v1.xml + envs/robot-02.xml.  Section numbers (A.x) refer to SURVEY.md Appendix A.
 *
 * The dynamics are computed with projected Newton-Euler sums over bodies in world coordinates
 * (mass matrix = sum_b m Jv'Jv + Jw' I Jw, bias = sum_b Jv' m (a_vp - g) + Jw' (I alpha_vp + W x I W)),
 * which yields the same M(q) and c(q,v) as MuJoCo's CRB / RNE.  It is deliberately a different
 * algorithm from the body-frame closed form the CUDA kernel uses, so the two check each other.
 */
#include "brb_ref.h"

#include <math.h>
#include <string.h>

#define MINVAL 1e-15
#define MINIMP 0.0001
#define MAXIMP 0.9999

/* ------------------------------------------------------------------ small vector helpers */
static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(double *r, const double *a, const double *b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void matvec3(double *r, const double *R, const double *v) { /* R row-major */
  double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline double normalize3(double *v) {
  double n = sqrt(dot3(v, v));
  if (n < MINVAL) { v[0] = 1; v[1] = 0; v[2] = 0; return 0.0; }
  v[0] /= n; v[1] /= n; v[2] /= n;
  return n;
}
static inline void normalize4(double *q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static inline void mulquat(double *r, const double *a, const double *b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static inline void quat2mat(double *R, const double *q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
}
static inline void axisangle2quat(double *q, const double *axis, double angle) {
  if (angle == 0) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  double s = sin(angle * 0.5);
  q[0] = cos(angle * 0.5); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}

/* dense Cholesky of the leading n x n block of A (row stride ld); lower factor in place. 0 on success */
static int chol_factor(double *A, int n, int ld) {
  for (int j = 0; j < n; j++) {
    double s = A[j * ld + j];
    for (int k = 0; k < j; k++) s -= A[j * ld + k] * A[j * ld + k];
    if (s <= 0) return -1;
    s = sqrt(s);
    A[j * ld + j] = s;
    for (int i = j + 1; i < n; i++) {
      double t = A[i * ld + j];
      for (int k = 0; k < j; k++) t -= A[i * ld + k] * A[j * ld + k];
      A[i * ld + j] = t / s;
    }
  }
  return 0;
}
static void chol_solve(const double *L, int n, int ld, double *x) {
  for (int i = 0; i < n; i++) {
    double t = x[i];
    for (int k = 0; k < i; k++) t -= L[i * ld + k] * x[k];
    x[i] = t / L[i * ld + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double t = x[i];
    for (int k = i + 1; k < n; k++) t -= L[k * ld + i] * x[k];
    x[i] = t / L[i * ld + i];
  }
}

int brb_ref_sizeof_model(void) { return (int)sizeof(BrbRefModel); }
int brb_ref_sizeof_data(void) { return (int)sizeof(BrbRefData); }
int brb_ref_sizeof_contact(void) { return (int)sizeof(BrbRefContact); }

/* ------------------------------------------------------------------ A.3 step 2: kinematics */
void brb_ref_kinematics(const BrbRefModel *m, BrbRefData *d) {
  memset(d->xpos[0], 0, sizeof d->xpos[0]);
  d->xquat[0][0] = 1; d->xquat[0][1] = d->xquat[0][2] = d->xquat[0][3] = 0;
  quat2mat(d->xmat[0], d->xquat[0]);
  memset(d->xipos[0], 0, sizeof d->xipos[0]);
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parent[b], j = m->body_jnt[b];
    double *xp = d->xpos[b], *xq = d->xquat[b];
    if (j >= 0 && m->jnt_type[j] == BRB_JNT_FREE) {
      int a = m->jnt_qposadr[j];
      memcpy(xp, d->qpos + a, 3 * sizeof(double));
      memcpy(xq, d->qpos + a + 3, 4 * sizeof(double));
      normalize4(xq);
      memcpy(d->xanchor[j], xp, 3 * sizeof(double));
      d->xaxis[j][0] = d->xaxis[j][1] = 0; d->xaxis[j][2] = 1;
    } else {
      double t[3];
      matvec3(t, d->xmat[p], m->body_pos[b]);
      for (int k = 0; k < 3; k++) xp[k] = d->xpos[p][k] + t[k];
      mulquat(xq, d->xquat[p], m->body_quat[b]);
      if (j >= 0) { /* hinge */
        double R[9], qr[4], off[3];
        quat2mat(R, xq);
        matvec3(off, R, m->jnt_pos[j]);
        for (int k = 0; k < 3; k++) d->xanchor[j][k] = xp[k] + off[k];
        matvec3(d->xaxis[j], R, m->jnt_axis[j]);
        int a = m->jnt_qposadr[j];
        axisangle2quat(qr, m->jnt_axis[j], d->qpos[a] - m->qpos0[a]);
        mulquat(xq, xq, qr);
        quat2mat(R, xq);
        matvec3(off, R, m->jnt_pos[j]);
        for (int k = 0; k < 3; k++) xp[k] = d->xanchor[j][k] - off[k];
      }
      normalize4(xq);
    }
    quat2mat(d->xmat[b], xq);
    double t[3];
    matvec3(t, d->xmat[b], m->body_ipos[b]);
    for (int k = 0; k < 3; k++) d->xipos[b][k] = xp[k] + t[k];
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_body[g];
    double t[3], Rg[9];
    matvec3(t, d->xmat[b], m->geom_pos[g]);
    for (int k = 0; k < 3; k++) d->geom_xpos[g][k] = d->xpos[b][k] + t[k];
    quat2mat(Rg, m->geom_quat[g]);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++)
        d->geom_xmat[g][3 * r + c] = d->xmat[b][3 * r] * Rg[c] + d->xmat[b][3 * r + 1] * Rg[3 + c] + d->xmat[b][3 * r + 2] * Rg[6 + c];
  }
}

/* A.7: point Jacobians in MuJoCo's dof convention (free joint: world-frame linear velocity of the
 * body origin, BODY-frame angular velocity; hinge: rotation about the world-frame axis through the anchor) */
void brb_ref_jac(const BrbRefModel *m, const BrbRefData *d, int body, const double p[3], double *jacp, double *jacr) {
  int nv = m->nv;
  if (jacp) memset(jacp, 0, 3 * nv * sizeof(double));
  if (jacr) memset(jacr, 0, 3 * nv * sizeof(double));
  for (int b = body; b > 0; b = m->body_parent[b]) {
    int j = m->body_jnt[b];
    if (j < 0) continue;
    int a = m->jnt_dofadr[j];
    if (m->jnt_type[j] == BRB_JNT_FREE) {
      double r[3] = {p[0] - d->xpos[b][0], p[1] - d->xpos[b][1], p[2] - d->xpos[b][2]};
      for (int k = 0; k < 3; k++) {
        if (jacp) jacp[k * nv + a + k] = 1.0;
        double col[3] = {d->xmat[b][k], d->xmat[b][3 + k], d->xmat[b][6 + k]}, c[3];
        cross3(c, col, r);
        for (int i = 0; i < 3; i++) {
          if (jacr) jacr[i * nv + a + 3 + k] = col[i];
          if (jacp) jacp[i * nv + a + 3 + k] = c[i];
        }
      }
    } else {
      double r[3] = {p[0] - d->xanchor[j][0], p[1] - d->xanchor[j][1], p[2] - d->xanchor[j][2]}, c[3];
      cross3(c, d->xaxis[j], r);
      for (int i = 0; i < 3; i++) {
        if (jacr) jacr[i * nv + a] = d->xaxis[j][i];
        if (jacp) jacp[i * nv + a] = c[i];
      }
    }
  }
}

static void body_world_inertia(const BrbRefModel *m, const BrbRefData *d, int b, double *Iw) {
  const double *R = d->xmat[b], *I = m->body_inertia[b];
  double T[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) T[3 * r + c] = R[3 * r] * I[c] + R[3 * r + 1] * I[3 + c] + R[3 * r + 2] * I[6 + c];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) Iw[3 * r + c] = T[3 * r] * R[3 * c] + T[3 * r + 1] * R[3 * c + 1] + T[3 * r + 2] * R[3 * c + 2];
}

/* A.3 step 3: joint-space inertia M(q), dense nv x nv (row stride BRB_MAXNV) */
void brb_ref_mass_matrix(const BrbRefModel *m, BrbRefData *d) {
  int nv = m->nv;
  double jp[3 * BRB_MAXNV], jr[3 * BRB_MAXNV], Iw[9], IJ[3 * BRB_MAXNV];
  memset(d->qM, 0, sizeof d->qM);
  for (int b = 1; b < m->nbody; b++) {
    brb_ref_jac(m, d, b, d->xipos[b], jp, jr);
    body_world_inertia(m, d, b, Iw);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < nv; c++) IJ[r * nv + c] = Iw[3 * r] * jr[c] + Iw[3 * r + 1] * jr[nv + c] + Iw[3 * r + 2] * jr[2 * nv + c];
    for (int i = 0; i < nv; i++)
      for (int k = 0; k < nv; k++) {
        double s = 0;
        for (int r = 0; r < 3; r++) s += m->body_mass[b] * jp[r * nv + i] * jp[r * nv + k] + jr[r * nv + i] * IJ[r * nv + k];
        d->qM[i * BRB_MAXNV + k] += s;
      }
  }
}

/* A.3 step 4: qfrc_bias = C(q,v) v + G(q) via velocity-product accelerations (qacc = 0) */
void brb_ref_bias(const BrbRefModel *m, BrbRefData *d) {
  int nv = m->nv;
  double W[BRB_MAXBODY][3], Al[BRB_MAXBODY][3], aO[BRB_MAXBODY][3];
  double jp[3 * BRB_MAXNV], jr[3 * BRB_MAXNV], Iw[9];
  memset(d->qfrc_bias, 0, sizeof d->qfrc_bias);
  memset(W, 0, sizeof W); memset(Al, 0, sizeof Al); memset(aO, 0, sizeof aO);
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parent[b], j = m->body_jnt[b];
    if (j >= 0 && m->jnt_type[j] == BRB_JNT_FREE) {
      int a = m->jnt_dofadr[j];
      matvec3(W[b], d->xmat[b], d->qvel + a + 3); /* world angular velocity; alpha_vp = 0, a_origin_vp = 0 */
    } else {
      double ra[3], t[3], t2[3], aa[3];
      const double *anchor = (j >= 0) ? d->xanchor[j] : d->xpos[b];
      for (int k = 0; k < 3; k++) ra[k] = anchor[k] - d->xpos[p][k];
      cross3(t, Al[p], ra);
      cross3(t2, W[p], ra); cross3(t2, W[p], t2);
      for (int k = 0; k < 3; k++) aa[k] = aO[p][k] + t[k] + t2[k];
      memcpy(W[b], W[p], sizeof W[b]); memcpy(Al[b], Al[p], sizeof Al[b]);
      if (j >= 0) {
        double qd = d->qvel[m->jnt_dofadr[j]], wxa[3];
        cross3(wxa, W[p], d->xaxis[j]);
        for (int k = 0; k < 3; k++) { W[b][k] += d->xaxis[j][k] * qd; Al[b][k] += wxa[k] * qd; }
      }
      double ro[3];
      for (int k = 0; k < 3; k++) ro[k] = d->xpos[b][k] - anchor[k];
      cross3(t, Al[b], ro);
      cross3(t2, W[b], ro); cross3(t2, W[b], t2);
      for (int k = 0; k < 3; k++) aO[b][k] = aa[k] + t[k] + t2[k];
    }
    double rc[3], t[3], t2[3], F[3], N[3], IW[3], IA[3];
    for (int k = 0; k < 3; k++) rc[k] = d->xipos[b][k] - d->xpos[b][k];
    cross3(t, Al[b], rc);
    cross3(t2, W[b], rc); cross3(t2, W[b], t2);
    for (int k = 0; k < 3; k++) F[k] = m->body_mass[b] * (aO[b][k] + t[k] + t2[k] - m->gravity[k]);
    body_world_inertia(m, d, b, Iw);
    matvec3(IW, Iw, W[b]); matvec3(IA, Iw, Al[b]);
    cross3(N, W[b], IW);
    for (int k = 0; k < 3; k++) N[k] += IA[k];
    brb_ref_jac(m, d, b, d->xipos[b], jp, jr);
    for (int c = 0; c < nv; c++)
      for (int r = 0; r < 3; r++) d->qfrc_bias[c] += jp[r * nv + c] * F[r] + jr[r * nv + c] * N[r];
  }
}

double brb_ref_energy(const BrbRefModel *m, BrbRefData *d, double *kinetic, double *potential) {
  brb_ref_kinematics(m, d);
  brb_ref_mass_matrix(m, d);
  double ke = 0, pe = 0;
  for (int i = 0; i < m->nv; i++)
    for (int k = 0; k < m->nv; k++) ke += 0.5 * d->qvel[i] * d->qM[i * BRB_MAXNV + k] * d->qvel[k];
  for (int b = 1; b < m->nbody; b++) pe -= m->body_mass[b] * dot3(m->gravity, d->xipos[b]);
  if (kinetic) *kinetic = ke;
  if (potential) *potential = pe;
  return ke + pe;
}

/* ------------------------------------------------------------------ mj_setConst restatement */
int brb_ref_model_finalize(BrbRefModel *m) {
  static BrbRefData d; /* large; finalize is called once per model from one thread */
  memset(&d, 0, sizeof d);
  memset(m->qpos0, 0, sizeof m->qpos0);
  for (int j = 0; j < m->njnt; j++)
    if (m->jnt_type[j] == BRB_JNT_FREE) {
      int a = m->jnt_qposadr[j], b = m->jnt_body[j];
      for (int k = 0; k < 3; k++) m->qpos0[a + k] = m->body_pos[b][k];
      for (int k = 0; k < 4; k++) m->qpos0[a + 3 + k] = m->body_quat[b][k];
    }
  if (m->solver_tolerance <= 0) m->solver_tolerance = 1e-13;
  memcpy(d.qpos, m->qpos0, sizeof d.qpos);
  brb_ref_kinematics(m, &d);
  brb_ref_mass_matrix(m, &d);
  int nv = m->nv;
  double tr = 0;
  for (int i = 0; i < nv; i++) tr += d.qM[i * BRB_MAXNV + i];
  m->meaninertia = tr / nv;
  double L[BRB_MAXNV * BRB_MAXNV];
  memcpy(L, d.qM, sizeof L);
  if (chol_factor(L, nv, BRB_MAXNV)) return -1;
  m->body_invweight0[0][0] = m->body_invweight0[0][1] = 0;
  for (int b = 1; b < m->nbody; b++) {
    double jp[3 * BRB_MAXNV], jr[3 * BRB_MAXNV], col[BRB_MAXNV];
    brb_ref_jac(m, &d, b, d.xipos[b], jp, jr);
    double tt = 0, rr = 0;
    for (int r = 0; r < 3; r++) {
      memcpy(col, jp + r * nv, nv * sizeof(double));
      chol_solve(L, nv, BRB_MAXNV, col);
      for (int c = 0; c < nv; c++) tt += jp[r * nv + c] * col[c];
      memcpy(col, jr + r * nv, nv * sizeof(double));
      chol_solve(L, nv, BRB_MAXNV, col);
      for (int c = 0; c < nv; c++) rr += jr[r * nv + c] * col[c];
    }
    m->body_invweight0[b][0] = tt / 3 > MINVAL ? tt / 3 : MINVAL;
    m->body_invweight0[b][1] = rr / 3 > MINVAL ? rr / 3 : MINVAL;
  }
  return 0;
}

void brb_ref_reset_data(const BrbRefModel *m, BrbRefData *d) {
  memset(d, 0, sizeof *d);
  memcpy(d->qpos, m->qpos0, sizeof d->qpos);
}

/* ------------------------------------------------------------------ A.6 colliders */
static BrbRefContact *new_contact(BrbRefData *d, const BrbRefModel *m, int pair, double dist, const double *pos, const double *normal) {
  if (d->ncon >= BRB_MAXCON) return 0;
  BrbRefContact *c = &d->contact[d->ncon++];
  memset(c, 0, sizeof *c);
  c->dist = dist;
  memcpy(c->pos, pos, 3 * sizeof(double));
  memcpy(c->frame, normal, 3 * sizeof(double));
  c->pair = pair;
  c->dim = m->pair_condim[pair];
  c->body1 = m->geom_body[m->pair_geom1[pair]];
  c->body2 = m->geom_body[m->pair_geom2[pair]];
  c->includemargin = m->pair_margin[pair] - m->pair_gap[pair];
  memcpy(c->friction, m->pair_friction[pair], sizeof c->friction);
  memcpy(c->solref, m->pair_solref[pair], sizeof c->solref);
  memcpy(c->solimp, m->pair_solimp[pair], sizeof c->solimp);
  return c;
}

static void collide_plane_cylinder(const BrbRefModel *m, BrbRefData *d, int pair) {
  int g1 = m->pair_geom1[pair], g2 = m->pair_geom2[pair];
  const double *mat1 = d->geom_xmat[g1], *mat2 = d->geom_xmat[g2], *pos1 = d->geom_xpos[g1], *pos2 = d->geom_xpos[g2];
  double margin = m->pair_margin[pair], radius = m->geom_size[g2][0], halflen = m->geom_size[g2][1];
  double n[3] = {mat1[2], mat1[5], mat1[8]}, a[3] = {mat2[2], mat2[5], mat2[8]}, v[3], pos[3];
  double prjaxis = dot3(n, a);
  if (prjaxis > 0) { a[0] = -a[0]; a[1] = -a[1]; a[2] = -a[2]; prjaxis = -prjaxis; }
  double rel[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
  double dist0 = dot3(n, rel);
  for (int k = 0; k < 3; k++) v[k] = a[k] * prjaxis - n[k];
  double len2 = dot3(v, v);
  if (len2 >= MINVAL * MINVAL) {
    double s = radius / sqrt(len2);
    for (int k = 0; k < 3; k++) v[k] *= s;
  } else {
    v[0] = mat2[0] * radius; v[1] = mat2[3] * radius; v[2] = mat2[6] * radius;
  }
  double prjvec = dot3(v, n);
  for (int k = 0; k < 3; k++) a[k] *= halflen;
  prjaxis *= halflen;
  double dist = dist0 + prjaxis + prjvec;
  if (dist > margin) return;
  for (int k = 0; k < 3; k++) pos[k] = pos2[k] + v[k] + a[k] - n[k] * dist * 0.5;
  new_contact(d, m, pair, dist, pos, n);
  dist = dist0 - prjaxis + prjvec;
  if (dist <= margin) {
    for (int k = 0; k < 3; k++) pos[k] = pos2[k] + v[k] - a[k] - n[k] * dist * 0.5;
    new_contact(d, m, pair, dist, pos, n);
  }
  double prjvec1 = -prjvec * 0.5;
  dist = dist0 + prjaxis + prjvec1;
  if (dist <= margin) {
    double s[3];
    cross3(s, v, a);
    normalize3(s);
    double sc = radius * sqrt(3.0) * 0.5;
    for (int sign = 1; sign >= -1; sign -= 2) {
      for (int k = 0; k < 3; k++) pos[k] = pos2[k] + sign * sc * s[k] + a[k] - v[k] * 0.5 - n[k] * dist * 0.5;
      new_contact(d, m, pair, dist, pos, n);
    }
  }
}

static void collide_plane_box(const BrbRefModel *m, BrbRefData *d, int pair) {
  int g1 = m->pair_geom1[pair], g2 = m->pair_geom2[pair];
  const double *mat1 = d->geom_xmat[g1], *mat2 = d->geom_xmat[g2], *pos1 = d->geom_xpos[g1], *pos2 = d->geom_xpos[g2];
  double margin = m->pair_margin[pair];
  const double *size = m->geom_size[g2];
  double n[3] = {mat1[2], mat1[5], mat1[8]};
  double rel[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
  double dist0 = dot3(n, rel);
  int cnt = 0;
  for (int i = 0; i < 8; i++) {
    double vec[3] = {(i & 1 ? size[0] : -size[0]), (i & 2 ? size[1] : -size[1]), (i & 4 ? size[2] : -size[2])}, corner[3], pos[3];
    matvec3(corner, mat2, vec);
    double ldist = dot3(n, corner);
    if (dist0 + ldist > margin || ldist > 0) continue;
    double dist = dist0 + ldist;
    for (int k = 0; k < 3; k++) pos[k] = corner[k] + pos2[k] - n[k] * dist * 0.5;
    new_contact(d, m, pair, dist, pos, n);
    if (++cnt >= 4) return;
  }
}


/* box-box: separating-axis test over the 15 candidate axes, then either face clipping (reference face = axis of least
 * penetration, incident face of the other box clipped against its side planes, Sutherland-Hodgman) or the closest points
 * of the two edges.  NOT a restatement of MuJoCo's mjc_BoxBox (engine_collision_box.c, several hundred lines that are not
 * reproducible from memory): it yields the same contact manifold class (<= 8 points, normal = axis of least penetration,
 * point = midway between the surfaces) but individual points can differ.  Contact normal points from geom1 to geom2. */
static void collide_box_box(const BrbRefModel *m, BrbRefData *d, int pair) {
  int g1 = m->pair_geom1[pair], g2 = m->pair_geom2[pair];
  const double *R1 = d->geom_xmat[g1], *R2 = d->geom_xmat[g2], *p1 = d->geom_xpos[g1], *p2 = d->geom_xpos[g2];
  const double *h1 = m->geom_size[g1], *h2 = m->geom_size[g2];
  double margin = m->pair_margin[pair];
  double dp[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  double A[3][3], B[3][3]; /* box axes (columns of R) as rows here */
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < 3; k++) { A[i][k] = R1[3 * k + i]; B[i][k] = R2[3 * k + i]; }
  double best = -1e30, bestn[3] = {0, 0, 1};
  int bestkind = -1, bi = 0, bj = 0; /* kind 0: face of 1, 1: face of 2, 2: edge-edge */
  for (int kind = 0; kind < 2; kind++)
    for (int i = 0; i < 3; i++) {
      const double *L = kind == 0 ? A[i] : B[i];
      double r1 = 0, r2 = 0;
      for (int k = 0; k < 3; k++) { r1 += h1[k] * fabs(dot3(A[k], L)); r2 += h2[k] * fabs(dot3(B[k], L)); }
      double t = dot3(dp, L), s = fabs(t) - (r1 + r2);
      if (s > margin) return;
      if (s > best) {
        best = s; bestkind = kind; bi = i;
        for (int k = 0; k < 3; k++) bestn[k] = t >= 0 ? L[k] : -L[k];
      }
    }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double L[3];
      cross3(L, A[i], B[j]);
      double n = sqrt(dot3(L, L));
      if (n < 1e-6) continue;
      for (int k = 0; k < 3; k++) L[k] /= n;
      double r1 = 0, r2 = 0;
      for (int k = 0; k < 3; k++) { r1 += h1[k] * fabs(dot3(A[k], L)); r2 += h2[k] * fabs(dot3(B[k], L)); }
      double t = dot3(dp, L), s = fabs(t) - (r1 + r2);
      if (s > margin) return;
      if (s > best + 1e-3 * fabs(best) + 1e-9 && s > 0.95 * best + (best < 0 ? 0.05 * best : 0) ) {
        /* edge axes win only when clearly better than the best face axis (standard bias towards face contacts) */
        if (s > best + 1e-4) { best = s; bestkind = 2; bi = i; bj = j; for (int k = 0; k < 3; k++) bestn[k] = t >= 0 ? L[k] : -L[k]; }
      }
    }
  if (bestkind == 2) {
    /* closest points of edge bi of box 1 and edge bj of box 2: pick the edges supporting along +-n */
    double c1[3], c2[3];
    for (int k = 0; k < 3; k++) { c1[k] = p1[k]; c2[k] = p2[k]; }
    for (int a = 0; a < 3; a++) {
      if (a != bi) { double sg = dot3(A[a], bestn) > 0 ? 1 : -1; for (int k = 0; k < 3; k++) c1[k] += sg * h1[a] * A[a][k]; }
      if (a != bj) { double sg = dot3(B[a], bestn) > 0 ? -1 : 1; for (int k = 0; k < 3; k++) c2[k] += sg * h2[a] * B[a][k]; }
    }
    const double *u = A[bi], *v = B[bj];
    double w[3] = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]};
    double uv = dot3(u, v), uw = dot3(u, w), vw = dot3(v, w), den = 1 - uv * uv;
    double sa = den > 1e-12 ? (uv * vw - uw) / den : 0, sb = den > 1e-12 ? (vw - uv * uw) / den : 0;
    sa = fmax(-h1[bi], fmin(h1[bi], sa)); sb = fmax(-h2[bj], fmin(h2[bj], sb));
    double q1[3], q2[3], pos[3];
    for (int k = 0; k < 3; k++) { q1[k] = c1[k] + sa * u[k]; q2[k] = c2[k] + sb * v[k]; pos[k] = 0.5 * (q1[k] + q2[k]); }
    new_contact(d, m, pair, best, pos, bestn);
    return;
  }
  /* face contact: reference box/face, incident box */
  const double (*RA)[3] = bestkind == 0 ? A : B, (*IB)[3] = bestkind == 0 ? B : A;
  const double *hr = bestkind == 0 ? h1 : h2, *hi = bestkind == 0 ? h2 : h1, *pr = bestkind == 0 ? p1 : p2, *pi = bestkind == 0 ? p2 : p1;
  double nref[3]; /* outward normal of the reference face (points towards the incident box) */
  for (int k = 0; k < 3; k++) nref[k] = bestkind == 0 ? bestn[k] : -bestn[k];
  /* incident face: the face of the incident box most anti-parallel to nref */
  int ia = 0; double mind = 1e30, isg = 1;
  for (int a = 0; a < 3; a++) { double t = dot3(IB[a], nref); if (-fabs(t) < mind) { mind = -fabs(t); ia = a; isg = t > 0 ? -1 : 1; } }
  int a1 = (ia + 1) % 3, a2 = (ia + 2) % 3;
  double poly[16][3], tmp[16][3];
  int np = 4;
  for (int c = 0; c < 4; c++) {
    double s1 = (c == 0 || c == 3) ? 1 : -1, s2 = (c < 2) ? 1 : -1;
    for (int k = 0; k < 3; k++) poly[c][k] = pi[k] + isg * hi[ia] * IB[ia][k] + s1 * hi[a1] * IB[a1][k] + s2 * hi[a2] * IB[a2][k];
  }
  int r1 = (bi + 1) % 3, r2 = (bi + 2) % 3;
  for (int side = 0; side < 4; side++) { /* clip against the four side planes of the reference face */
    const double *ax = RA[side < 2 ? r1 : r2];
    double sg = (side & 1) ? -1 : 1, lim = hr[side < 2 ? r1 : r2];
    int nn = 0;
    for (int c = 0; c < np; c++) {
      const double *P = poly[c], *Q = poly[(c + 1) % np];
      double rel1[3] = {P[0] - pr[0], P[1] - pr[1], P[2] - pr[2]}, rel2[3] = {Q[0] - pr[0], Q[1] - pr[1], Q[2] - pr[2]};
      double dP = sg * dot3(rel1, ax) - lim, dQ = sg * dot3(rel2, ax) - lim;
      if (dP <= 0) { memcpy(tmp[nn++], P, sizeof(double) * 3); }
      if ((dP < 0 && dQ > 0) || (dP > 0 && dQ < 0)) {
        double t = dP / (dP - dQ);
        for (int k = 0; k < 3; k++) tmp[nn][k] = P[k] + t * (Q[k] - P[k]);
        nn++;
      }
      if (nn >= 15) break;
    }
    np = nn;
    memcpy(poly, tmp, sizeof(double) * 3 * np);
    if (np == 0) return;
  }
  int cnt = 0;
  for (int c = 0; c < np && cnt < 8; c++) {
    double rel[3] = {poly[c][0] - pr[0], poly[c][1] - pr[1], poly[c][2] - pr[2]};
    double depth = dot3(rel, nref) - hr[bi]; /* signed distance of the incident point from the reference face */
    if (depth > margin) continue;
    double pos[3];
    for (int k = 0; k < 3; k++) pos[k] = poly[c][k] - nref[k] * depth * 0.5;
    new_contact(d, m, pair, depth, pos, bestn);
    cnt++;
  }
}

/* cylinder-box (Env03-v2: wheel vs block).  MuJoCo 3.2.0 sends this pair to libccd's MPR through mjc_Convex: an iterative portal
 * refinement whose result on a curved feature is only good to its tolerance (1e-6 m of support distance, i.e. about 1 % of normal
 * direction on the 34 mm wheel).  It is NOT restated.  Own analytic collider, same contact class (one point, normal from geom1 to
 * geom2, position midway between the surfaces):
 *   sep(d) = d.(c2 - c1) - h_cyl(d) - h_box(d)  is a lower bound of the signed distance for every unit d (support functions), and
 *   the signed distance is its maximum.  It is evaluated on the directions at which the maximum can sit for these two shapes:
 *   box face normals (3), the cylinder axis, axis x box edge (3: edge against the curved side), the radial direction to each box vertex
 *   (8: vertex against the curved side), the direction from the nearest rim point to each vertex (8), and for each box
 *   edge direction the best direction perpendicular to it (edge against a rim circle: a one-dimensional search, see below).  The contact point comes from two alternating projections between the
 *   support feature of the box and of the cylinder along the chosen direction. */
static double cyl_box_sep(const double *dir, const double *delta, const double *a, double R, double L, const double E[3][3], const double *h) {
  double da = dot3(dir, a), s = dot3(dir, delta) - L * fabs(da) - R * sqrt(fmax(0.0, 1.0 - da * da));
  for (int j = 0; j < 3; j++) s -= h[j] * fabs(dot3(dir, E[j]));
  return s;
}
#define CYLBOX_TAU 0.02
static double soft_sign(double x) { return fmax(-1.0, fmin(1.0, x / CYLBOX_TAU)); }
/* cos(2 pi k / 16); sin(x) = cos(x - pi/2) = entry (k + 12) & 15 */
static const double CYLBOX_COS16[16] = {1.0, 0.92387953251128674, 0.70710678118654752, 0.38268343236508977, 0.0, -0.38268343236508977,
                                        -0.70710678118654752, -0.92387953251128674, -1.0, -0.92387953251128674, -0.70710678118654752,
                                        -0.38268343236508977, 0.0, 0.38268343236508977, 0.70710678118654752, 0.92387953251128674};
static double cyl_box_g(double c, double s, double D1, double D2, double A1, double A2, double R, double L, double h1, double h2) {
  const double da = A1 * c + A2 * s;
  return D1 * c + D2 * s - L * fabs(da) - R * sqrt(fmax(0.0, 1.0 - da * da)) - h1 * fabs(c) - h2 * fabs(s);
}
/* candidate direction v (any length; `orient`: flip it towards the box centre first): keeps the direction of the largest separation */
static void cyl_box_try(const double *v, int orient, const double *delta, const double *a, double R, double L, const double E[3][3],
                        const double *h, double *best, double *bd) {
  double n2 = dot3(v, v);
  if (n2 <= 1e-16) return;
  double in = 1.0 / sqrt(n2);
  if (orient && dot3(v, delta) < 0) in = -in;
  double t[3] = {v[0] * in, v[1] * in, v[2] * in}, s = cyl_box_sep(t, delta, a, R, L, E, h);
  if (s > *best) { *best = s; bd[0] = t[0]; bd[1] = t[1]; bd[2] = t[2]; }
}

/* returns 1 and fills dist / normal / pos when the shapes are within `margin`; exposed for tests */
int brb_ref_cylinder_box(const double c[3], const double a_in[3], double R, double L, const double b[3], const double Emat[9] /* rows = box axes */,
                         const double h[3], double margin, double *dist, double normal[3], double pos[3]) {
  double a[3] = {a_in[0], a_in[1], a_in[2]}, E[3][3], delta[3] = {b[0] - c[0], b[1] - c[1], b[2] - c[2]};
  normalize3(a);
  for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) E[i][k] = Emat[3 * i + k];
  double best = -1e30, bd[3] = {0, 0, 1};
  for (int i = 0; i < 3; i++) cyl_box_try(E[i], 1, delta, a, R, L, E, h, &best, bd);
  cyl_box_try(a, 1, delta, a, R, L, E, h, &best, bd);
  for (int i = 0; i < 3; i++) { double x[3]; cross3(x, a, E[i]); if (dot3(x, x) > 1e-10) cyl_box_try(x, 1, delta, a, R, L, E, h, &best, bd); }
  for (int vi = 0; vi < 8; vi++) {
    double u[3];
    for (int k = 0; k < 3; k++) u[k] = delta[k] + ((vi & 1) ? h[0] : -h[0]) * E[0][k] + ((vi & 2) ? h[1] : -h[1]) * E[1][k] + ((vi & 4) ? h[2] : -h[2]) * E[2][k];
    double ua = dot3(u, a), up[3] = {u[0] - ua * a[0], u[1] - ua * a[1], u[2] - ua * a[2]};
    double rho2 = dot3(up, up);
    if (rho2 <= 1e-16) continue;
    /* vertex against the curved side: radial direction (no re-orientation: it points from the axis to the vertex) */
    cyl_box_try(up, 0, delta, a, R, L, E, h, &best, bd);
    /* vertex against the rim of the cap on its side of the cylinder (the other rim is farther): from the nearest rim point to the vertex */
    {
      double sg = ua < 0 ? -1.0 : 1.0, ir = R / sqrt(rho2), t[3];
      for (int k = 0; k < 3; k++) t[k] = u[k] - sg * L * a[k] - up[k] * ir;
      cyl_box_try(t, 0, delta, a, R, L, E, h, &best, bd);
    }
  }
  /* box EDGE against a rim (or the curved side, or a cap): the separating direction is perpendicular to the edge, d(phi) = cos(phi) E_a1 +
   * sin(phi) E_a2 for an edge along E_ax.  In that plane sep is the one-dimensional function
   *   g(phi) = D1 c + D2 s - L |A1 c + A2 s| - R sqrt(1 - (A1 c + A2 s)^2) - h1 |c| - h2 |s|        (all four parallel edges and both caps at once)
   * which is maximised by a 16-point scan followed by a golden-section search with a fixed number of steps inside the best cell
   * (parametrised by the tangent of the offset, so no trigonometry per step). */
  for (int ax = 0; ax < 3; ax++) {
    const int a1 = (ax + 1) % 3, a2 = (ax + 2) % 3;
    const double D1 = dot3(E[a1], delta), D2 = dot3(E[a2], delta), A1 = dot3(E[a1], a), A2 = dot3(E[a2], a);
    int kb = 0;
    double gb = -1e30;
    for (int k = 0; k < 16; k++) {
      const double g = cyl_box_g(CYLBOX_COS16[k], CYLBOX_COS16[(k + 12) & 15], D1, D2, A1, A2, R, L, h[a1], h[a2]);
      if (g > gb) { gb = g; kb = k; }
    }
    const double ck = CYLBOX_COS16[kb], sk = CYLBOX_COS16[(kb + 12) & 15], tmax = 0.41421356237309503 /* tan(pi/8) */, gr = 0.6180339887498949;
    double lo = -tmax, hi = tmax, x1 = hi - gr * (hi - lo), x2 = lo + gr * (hi - lo), f1 = 0, f2 = 0;
    int fresh = 2;
    for (int it = 0; it < 10; it++) {
      for (int w = 0; w < 2; w++) {
        if (fresh != 2 && w != fresh) continue;
        const double t = w ? x2 : x1, in = 1.0 / sqrt(1.0 + t * t);
        const double g = cyl_box_g((ck - t * sk) * in, (sk + t * ck) * in, D1, D2, A1, A2, R, L, h[a1], h[a2]);
        if (w) f2 = g; else f1 = g;
      }
      if (f1 > f2) { hi = x2; x2 = x1; f2 = f1; x1 = hi - gr * (hi - lo); fresh = 0; }
      else { lo = x1; x1 = x2; f1 = f2; x2 = lo + gr * (hi - lo); fresh = 1; }
    }
    const double t = 0.5 * (lo + hi), in = 1.0 / sqrt(1.0 + t * t), cf = (ck - t * sk) * in, sf = (sk + t * ck) * in;
    double tdir[3];
    for (int k = 0; k < 3; k++) tdir[k] = cf * E[a1][k] + sf * E[a2][k];
    cyl_box_try(tdir, 0, delta, a, R, L, E, h, &best, bd);
  }
  if (best > margin) return 0;
  /* Contact point: support features of the two shapes along bd, refined by two alternating projections.  A feature is "selected" by
   * the sign of bd along an axis; within CYLBOX_TAU (0.02 rad) of perpendicular the selection is blended linearly with the free
   * coordinate (the projection of the other shape's point), so the point moves continuously from the middle of a flat-on-flat patch
   * to its deeper end as the shapes tilt instead of jumping with the sign of a 1e-6 rad tilt. */
  double sa = soft_sign(dot3(bd, a)), sb[3], dp[3], rdir[3] = {0, 0, 0}, wr = 0.0;
  for (int j = 0; j < 3; j++) sb[j] = soft_sign(dot3(bd, E[j]));
  { double da = dot3(bd, a); for (int k = 0; k < 3; k++) dp[k] = bd[k] - da * a[k]; double pm = sqrt(dot3(dp, dp));
    if (pm > 1e-12) { for (int k = 0; k < 3; k++) rdir[k] = dp[k] / pm; wr = fmin(1.0, pm / CYLBOX_TAU); } }
  double pc[3], qb[3];
  for (int k = 0; k < 3; k++) pc[k] = c[k] + sa * L * a[k] + wr * R * rdir[k];
  for (int pass = 0; pass < 2; pass++) {
    for (int k = 0; k < 3; k++) qb[k] = b[k];
    for (int j = 0; j < 3; j++) {
      double rel[3] = {pc[0] - b[0], pc[1] - b[1], pc[2] - b[2]};
      double l = -sb[j] * h[j] + (1.0 - fabs(sb[j])) * fmax(-h[j], fmin(h[j], dot3(rel, E[j])));
      for (int k = 0; k < 3; k++) qb[k] += l * E[j][k];
    }
    double rel[3] = {qb[0] - c[0], qb[1] - c[1], qb[2] - c[2]}, ta = dot3(rel, a);
    double t = sa * L + (1.0 - fabs(sa)) * fmax(-L, fmin(L, ta));
    double rv[3];
    for (int k = 0; k < 3; k++) rv[k] = rel[k] - ta * a[k];                /* radial part of the box point, inside the disc */
    { double n2 = dot3(rv, rv); if (n2 > R * R) { double in = R / sqrt(n2); for (int k = 0; k < 3; k++) rv[k] *= in; } }
    for (int k = 0; k < 3; k++) pc[k] = c[k] + t * a[k] + wr * R * rdir[k] + (1.0 - wr) * rv[k];
  }
  *dist = best;
  for (int k = 0; k < 3; k++) { normal[k] = bd[k]; pos[k] = 0.5 * (pc[k] + qb[k]); }
  return 1;
}

static void collide_cylinder_box(const BrbRefModel *m, BrbRefData *d, int pair) {
  int g1 = m->pair_geom1[pair], g2 = m->pair_geom2[pair];
  const double *R1 = d->geom_xmat[g1], *R2 = d->geom_xmat[g2];
  double a[3] = {R1[2], R1[5], R1[8]}, E[9], dist, n[3], pos[3];
  for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) E[3 * i + k] = R2[3 * k + i];
  if (brb_ref_cylinder_box(d->geom_xpos[g1], a, m->geom_size[g1][0], m->geom_size[g1][1], d->geom_xpos[g2], E, m->geom_size[g2],
                           m->pair_margin[pair], &dist, n, pos))
    new_contact(d, m, pair, dist, pos, n);
}

static void make_frame(double *f) { /* mju_makeFrame with an undefined y axis */
  normalize3(f);
  f[3] = f[4] = f[5] = 0;
  if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1;
  double t = dot3(f, f + 3);
  for (int k = 0; k < 3; k++) f[3 + k] -= t * f[k];
  normalize3(f + 3);
  cross3(f + 6, f, f + 3);
}

static void collision(const BrbRefModel *m, BrbRefData *d) {
  d->ncon = 0;
  for (int p = 0; p < m->npair; p++) {
    int t1 = m->geom_type[m->pair_geom1[p]], t2 = m->geom_type[m->pair_geom2[p]];
    if (t1 == BRB_GEOM_PLANE && t2 == BRB_GEOM_CYLINDER) collide_plane_cylinder(m, d, p);
    else if (t1 == BRB_GEOM_PLANE && t2 == BRB_GEOM_BOX) collide_plane_box(m, d, p);
    else if (t1 == BRB_GEOM_BOX && t2 == BRB_GEOM_BOX) collide_box_box(m, d, p);
    else if (t1 == BRB_GEOM_CYLINDER && t2 == BRB_GEOM_BOX && (m->flags & BRB_FLAG_CYLINDER_BOX)) collide_cylinder_box(m, d, p);
  }
  for (int i = 0; i < d->ncon; i++) make_frame(d->contact[i].frame);
}

/* ------------------------------------------------------------------ A.7 constraint rows */
static double get_impedance(const double *solimp_in, double pos, double margin) {
  double s[5];
  memcpy(s, solimp_in, sizeof s);
  s[0] = fmin(MAXIMP, fmax(MINIMP, s[0])); s[1] = fmin(MAXIMP, fmax(MINIMP, s[1]));
  s[2] = fmax(0, s[2]); s[3] = fmin(MAXIMP, fmax(MINIMP, s[3])); s[4] = fmax(1, s[4]);
  if (s[0] == s[1] || s[2] <= MINVAL) return 0.5 * (s[0] + s[1]);
  double x = (pos - margin) / s[2];
  if (x < 0) x = -x;
  if (x >= 1) return s[1];
  if (x <= 0) return s[0];
  double y;
  if (s[4] == 1) y = x;
  else if (x <= s[3]) y = pow(x, s[4]) / pow(s[3], s[4] - 1);
  else y = 1 - pow(1 - x, s[4]) / pow(1 - s[3], s[4] - 1);
  return s[0] + y * (s[1] - s[0]);
}

static void make_constraint(const BrbRefModel *m, BrbRefData *d) {
  int nv = m->nv;
  d->nefc = 0;
  double jp1[3 * BRB_MAXNV], jp2[3 * BRB_MAXNV], jc[3 * BRB_MAXNV];
  for (int ci = 0; ci < d->ncon; ci++) {
    BrbRefContact *c = &d->contact[ci];
    c->exclude = (c->dist >= c->includemargin);
    c->efc_address = -1;
    if (c->exclude) continue;
    int nrow = (c->dim == 1) ? 1 : 2 * (c->dim - 1);
    if (c->dim != 1 && c->dim != 3) continue; /* only condim 1 and 3 occur in the supported scenes */
    if (d->nefc + nrow > BRB_MAXEFC) break;
    brb_ref_jac(m, d, c->body1, c->pos, jp1, 0);
    brb_ref_jac(m, d, c->body2, c->pos, jp2, 0);
    for (int r = 0; r < 3; r++)
      for (int k = 0; k < nv; k++) {
        double s = 0;
        for (int i = 0; i < 3; i++) s += c->frame[3 * r + i] * (jp2[i * nv + k] - jp1[i * nv + k]);
        jc[r * nv + k] = s;
      }
    c->efc_address = d->nefc;
    double tran = m->body_invweight0[c->body1][0] + m->body_invweight0[c->body2][0];
    double imp = get_impedance(c->solimp, c->dist, c->includemargin);
    /* reference accel parameters (standard solref, refsafe on) */
    double tc = c->solref[0], dr = c->solref[1], dmax = fmin(MAXIMP, fmax(MINIMP, c->solimp[1])), K, B;
    if (tc > 0) {
      tc = fmax(tc, 2 * m->timestep);
      K = 1.0 / fmax(MINVAL, dmax * dmax * tc * tc * dr * dr);
      B = 2.0 / fmax(MINVAL, dmax * tc);
    } else {
      K = -tc / fmax(MINVAL, dmax * dmax);
      B = -dr / fmax(MINVAL, dmax);
    }
    double R0 = 0;
    for (int r = 0; r < nrow; r++) {
      int e = d->nefc + r;
      double *J = d->efc_J + e * BRB_MAXNV;
      double diag;
      if (c->dim == 1) {
        memcpy(J, jc, nv * sizeof(double));
        diag = tran;
      } else {
        double mu = c->friction[r / 2], sg = (r & 1) ? -1.0 : 1.0;
        const double *jt = jc + (1 + r / 2) * nv;
        for (int k = 0; k < nv; k++) J[k] = jc[k] + sg * mu * jt[k];
        diag = tran * (1 + mu * mu);
      }
      double Rr = fmax(MINVAL, (1 - imp) / imp * diag);
      if (r == 0) R0 = (m->flags & BRB_FLAG_RPY_FROM_FIRST_ROW) ? Rr : fmax(MINVAL, (1 - imp) / imp * tran);
      if (c->dim > 1) Rr = 2 * c->friction[0] * c->friction[0] * R0;
      d->efc_R[e] = Rr;
      d->efc_D[e] = 1.0 / Rr;
      d->efc_pos[e] = c->dist;
      d->efc_margin[e] = c->includemargin;
      double vel = 0;
      for (int k = 0; k < nv; k++) vel += J[k] * d->qvel[k];
      d->efc_vel[e] = vel;
      d->efc_aref[e] = -B * vel - K * imp * (c->dist - c->includemargin);
    }
    d->nefc += nrow;
  }
}

/* ------------------------------------------------------------------ A.8 Newton solver (primal) */
static double solver_cost(const BrbRefModel *m, const BrbRefData *d, const double *a) {
  int nv = m->nv;
  double cost = 0, da[BRB_MAXNV];
  for (int i = 0; i < nv; i++) da[i] = a[i] - d->qacc_smooth[i];
  for (int i = 0; i < nv; i++)
    for (int k = 0; k < nv; k++) cost += 0.5 * da[i] * d->qM[i * BRB_MAXNV + k] * da[k];
  for (int e = 0; e < d->nefc; e++) {
    double jar = -d->efc_aref[e];
    for (int k = 0; k < nv; k++) jar += d->efc_J[e * BRB_MAXNV + k] * a[k];
    if (jar < 0) cost += 0.5 * d->efc_D[e] * jar * jar;
  }
  return cost;
}

static void solve_constraints(const BrbRefModel *m, BrbRefData *d) {
  int nv = m->nv, ne = d->nefc;
  double a[BRB_MAXNV];
  d->solver_niter = 0;
  memset(d->qfrc_constraint, 0, sizeof d->qfrc_constraint);
  if (ne == 0) {
    memcpy(d->qacc, d->qacc_smooth, sizeof d->qacc);
    return;
  }
  /* start from the better of warm start and unconstrained acceleration */
  if (solver_cost(m, d, d->qacc_warmstart) < solver_cost(m, d, d->qacc_smooth)) memcpy(a, d->qacc_warmstart, sizeof a);
  else memcpy(a, d->qacc_smooth, sizeof a);
  double scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  double jar[BRB_MAXEFC], jp[BRB_MAXEFC], g[BRB_MAXNV], p[BRB_MAXNV], H[BRB_MAXNV * BRB_MAXNV];
  unsigned char act[BRB_MAXEFC];
  for (int it = 0; it < 100; it++) {
    for (int e = 0; e < ne; e++) {
      double s = -d->efc_aref[e];
      for (int k = 0; k < nv; k++) s += d->efc_J[e * BRB_MAXNV + k] * a[k];
      jar[e] = s;
      act[e] = s < 0;
    }
    /* gradient g = M (a - a_smooth) + J' D jar_active */
    double gn = 0;
    for (int i = 0; i < nv; i++) {
      double s = 0;
      for (int k = 0; k < nv; k++) s += d->qM[i * BRB_MAXNV + k] * (a[k] - d->qacc_smooth[k]);
      for (int e = 0; e < ne; e++)
        if (act[e]) s += d->efc_J[e * BRB_MAXNV + i] * d->efc_D[e] * jar[e];
      g[i] = s;
      gn += s * s;
    }
    if (scale * sqrt(gn) < m->solver_tolerance) break;
    d->solver_niter++;
    /* Hessian H = M + J' D_active J; Newton direction p = -H^-1 g */
    memcpy(H, d->qM, sizeof H);
    for (int e = 0; e < ne; e++)
      if (act[e]) {
        const double *J = d->efc_J + e * BRB_MAXNV;
        for (int i = 0; i < nv; i++) {
          double s = d->efc_D[e] * J[i];
          if (s == 0) continue;
          for (int k = 0; k <= i; k++) H[i * BRB_MAXNV + k] += s * J[k];
        }
      }
    if (chol_factor(H, nv, BRB_MAXNV)) break;
    for (int i = 0; i < nv; i++) p[i] = -g[i];
    chol_solve(H, nv, BRB_MAXNV, p);
    for (int e = 0; e < ne; e++) {
      double s = 0;
      for (int k = 0; k < nv; k++) s += d->efc_J[e * BRB_MAXNV + k] * p[k];
      jp[e] = s;
    }
    /* full step keeps the active set -> exact minimiser of the (locally quadratic) cost: done */
    int same = 1;
    for (int e = 0; e < ne; e++)
      if ((jar[e] + jp[e] < 0) != act[e]) { same = 0; break; }
    if (same) {
      for (int i = 0; i < nv; i++) a[i] += p[i];
      break; /* consistent active set + full Newton step = exact minimiser */
    }
    /* exact line search on the piecewise-quadratic cost: root of the increasing piecewise-linear phi' */
    double pMp = 0, pMd = 0;
    for (int i = 0; i < nv; i++) {
      double s = 0;
      for (int k = 0; k < nv; k++) s += d->qM[i * BRB_MAXNV + k] * p[k];
      pMp += p[i] * s;
      pMd += s * (a[i] - d->qacc_smooth[i]);
    }
    double bp[BRB_MAXEFC + 1];
    int nb = 0;
    for (int e = 0; e < ne; e++)
      if (jp[e] != 0) {
        double t = -jar[e] / jp[e];
        if (t > 0) bp[nb++] = t;
      }
    for (int i = 1; i < nb; i++) { /* insertion sort */
      double t = bp[i];
      int k = i - 1;
      while (k >= 0 && bp[k] > t) { bp[k + 1] = bp[k]; k--; }
      bp[k + 1] = t;
    }
    double lo = 0, alpha = 1;
    for (int seg = 0; seg <= nb; seg++) {
      double hi = (seg < nb) ? bp[seg] : -1;
      double mid = (seg < nb) ? 0.5 * (lo + hi) : lo + 1.0;
      double c0 = pMd, c1 = pMp;
      for (int e = 0; e < ne; e++)
        if (jar[e] + mid * jp[e] < 0) { c0 += d->efc_D[e] * jar[e] * jp[e]; c1 += d->efc_D[e] * jp[e] * jp[e]; }
      double root = -c0 / c1;
      if (seg == nb || root <= hi) { alpha = root < lo ? lo : root; break; }
      lo = hi;
    }
    for (int i = 0; i < nv; i++) a[i] += alpha * p[i];
  }
  memcpy(d->qacc, a, sizeof a);
  for (int e = 0; e < ne; e++) {
    double s = -d->efc_aref[e];
    for (int k = 0; k < nv; k++) s += d->efc_J[e * BRB_MAXNV + k] * a[k];
    d->efc_force[e] = s < 0 ? -d->efc_D[e] * s : 0;
    if (d->efc_force[e] != 0)
      for (int k = 0; k < nv; k++) d->qfrc_constraint[k] += d->efc_J[e * BRB_MAXNV + k] * d->efc_force[e];
  }
}

/* ------------------------------------------------------------------ mj_forward */
static void forward_impl(const BrbRefModel *m, BrbRefData *d, double *Lm) {
  int nv = m->nv;
  brb_ref_kinematics(m, d);
  brb_ref_mass_matrix(m, d);
  memcpy(Lm, d->qM, sizeof(double) * BRB_MAXNV * BRB_MAXNV);
  chol_factor(Lm, nv, BRB_MAXNV);
  collision(m, d);
  /* velocity stage */
  memset(d->qfrc_passive, 0, sizeof d->qfrc_passive);
  for (int j = 0; j < m->njnt; j++)
    if (m->jnt_type[j] == BRB_JNT_HINGE) d->qfrc_passive[m->jnt_dofadr[j]] = -m->jnt_damping[j] * d->qvel[m->jnt_dofadr[j]];
  brb_ref_bias(m, d);
  /* actuation (A.3 step 5, Q7) */
  memset(d->qfrc_actuator, 0, sizeof d->qfrc_actuator);
  for (int u = 0; u < m->nu; u++) {
    int dof = m->jnt_dofadr[m->act_jnt[u]];
    double c = d->ctrl[u];
    if (m->act_ctrllimited[u]) c = fmin(m->act_ctrlrange[u][1], fmax(m->act_ctrlrange[u][0], c));
    double f = m->act_kv[u] * c - m->act_kv[u] * (m->act_gear[u] * d->qvel[dof]);
    if (m->act_forcelimited[u]) f = fmin(m->act_forcerange[u][1], fmax(m->act_forcerange[u][0], f));
    d->actuator_force[u] = f;
    d->qfrc_actuator[dof] += m->act_gear[u] * f;
  }
  for (int i = 0; i < nv; i++) {
    d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
    d->qacc_smooth[i] = d->qfrc_smooth[i];
  }
  chol_solve(Lm, nv, BRB_MAXNV, d->qacc_smooth);
  make_constraint(m, d);
  solve_constraints(m, d);
  memcpy(d->qacc_warmstart, d->qacc, sizeof d->qacc);
}

void brb_ref_forward(const BrbRefModel *m, BrbRefData *d) {
  double Lm[BRB_MAXNV * BRB_MAXNV];
  forward_impl(m, d, Lm);
}

/* ------------------------------------------------------------------ A.9 / A.10: implicitfast + advance */
void brb_ref_step(const BrbRefModel *m, BrbRefData *d, int nstep) {
  int nv = m->nv;
  double h = m->timestep;
  double Lm[BRB_MAXNV * BRB_MAXNV], A[BRB_MAXNV * BRB_MAXNV], qacc[BRB_MAXNV];
  for (int s = 0; s < nstep; s++) {
    forward_impl(m, d, Lm);
    d->stat_substeps++;
    d->stat_contact_substeps += d->nefc > 0;
    d->stat_newton_iters += d->solver_niter;
    d->stat_efc_rows += d->nefc;
    /* (M - h dF/dv) a+ = qfrc_smooth + qfrc_constraint; dF/dv = passive damping + actuator bias */
    memcpy(A, d->qM, sizeof A);
    for (int j = 0; j < m->njnt; j++)
      if (m->jnt_type[j] == BRB_JNT_HINGE) A[m->jnt_dofadr[j] * (BRB_MAXNV + 1)] += h * m->jnt_damping[j];
    for (int u = 0; u < m->nu; u++) {
      int dof = m->jnt_dofadr[m->act_jnt[u]];
      if ((m->flags & BRB_FLAG_ACTDERIV_SKIP_CLAMPED) && m->act_forcelimited[u] &&
          (d->actuator_force[u] <= m->act_forcerange[u][0] || d->actuator_force[u] >= m->act_forcerange[u][1]))
        continue;
      A[dof * (BRB_MAXNV + 1)] += h * m->act_kv[u] * m->act_gear[u] * m->act_gear[u];
    }
    chol_factor(A, nv, BRB_MAXNV);
    for (int i = 0; i < nv; i++) qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
    chol_solve(A, nv, BRB_MAXNV, qacc);
    /* mj_advance: velocity first, then positions with the NEW velocity */
    for (int i = 0; i < nv; i++) d->qvel[i] += h * qacc[i];
    for (int j = 0; j < m->njnt; j++) {
      int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
      if (m->jnt_type[j] == BRB_JNT_FREE) {
        for (int k = 0; k < 3; k++) d->qpos[qa + k] += h * d->qvel[da + k];
        double w[3] = {d->qvel[da + 3], d->qvel[da + 4], d->qvel[da + 5]}, qr[4];
        double ang = h * normalize3(w);
        axisangle2quat(qr, w, ang);
        normalize4(d->qpos + qa + 3);
        mulquat(d->qpos + qa + 3, d->qpos + qa + 3, qr);
      } else {
        d->qpos[qa] += h * d->qvel[da];
      }
    }
    d->time += h;
  }
}
