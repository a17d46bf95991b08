// brb_env03.cuh — Env03-v2: the balance robot plus a free 4 cm block fired at it (reference envs/env03_v1.py:17-113,
// envs/env03_v2.py:14-59, scene envs/env03_v1.xml:31-37).  Included by brb_kernels.cu (same translation unit, so it
// shares the robot physics, Philox and task-logic helpers).
//
// What changes against Env01: there are no explicit contact pairs, so every pair is a *dynamic* pair
// with mixed default parameters — mu = 1, position-dependent impedance (solimp 0.9/0.95/0.001), the block geom's margin
// 0.002 and solref (0.005, 0.9) averaged in.  Contacts modelled: wheel-floor (plane-cylinder), block-floor (plane-box),
// chassis-block (box-box: SAT + face clipping, the oracle's algorithm in fp32).  Not modelled (counted as unsupported):
// wheel-block (MuJoCo uses libccd MPR there) and chassis-floor.
//
// Solver structure per substep: while the block does not touch the chassis the two systems are independent — the
// robot's 8x8 (same code as Env01 with per-contact row weights) and the block's 6x6 (isotropic cube, M = diag(m, I)).
// During an impact (a few ms every 0.5 s+) all contacts go through one generic dense 14-dof Newton solve
// (env03_coupled_solve, local-memory arrays, not inlined: rare path).
#ifndef BRB_ENV03_CUH
#define BRB_ENV03_CUH

#define LT6(i, j) ((i) * ((i) + 1) / 2 + (j))

struct Blk {
  // state (plain fp32: the block only has to agree with the oracle to the 1e-5 bar, and it does without compensation)
  float p[3], q[4], v[3], w[3];   // world position, quaternion (w,x,y,z), world linear velocity, BODY angular velocity
  // working set of the current substep
  float ex[3], ey[3], ez[3];      // columns of the block's rotation matrix
  float cr[4][3], cy[4][3], cD[4];   // floor contacts: point relative to the block centre, yhat, row weight (dynamic index: local memory)
  int nc;                         // number of floor contacts (<= 4)
  unsigned bits;                  // active pyramid rows, 4 bits per contact slot
};

BRB_D void blk_frame(Blk &B) {
  const float qw = B.q[0], qx = B.q[1], qy = B.q[2], qz = B.q[3];
  const float xx = qx * qx, yy = qy * qy, zz = qz * qz, xy = qx * qy, xz = qx * qz, yz = qy * qz, wx = qw * qx, wy = qw * qy, wz = qw * qz;
  B.ex[0] = 1.f - 2.f * (yy + zz); B.ey[0] = 2.f * (xy - wz); B.ez[0] = 2.f * (xz + wy);
  B.ex[1] = 2.f * (xy + wz); B.ey[1] = 1.f - 2.f * (xx + zz); B.ez[1] = 2.f * (yz - wx);
  B.ex[2] = 2.f * (xz - wy); B.ey[2] = 2.f * (yz + wx); B.ez[2] = 1.f - 2.f * (xx + yy);
}

// plane-box (mjc_PlaneBox, A.6): corners in index order, at most 4, skipping corners above the centre
BRB_D void blk_setup(const BrbModelConsts &c, Blk &B) {
  blk_frame(B);
  B.nc = 0;
  const float h = c.blk_half, *pp = c.pp[1];
  const float dist0 = (B.p[2] - c.zfloor) - c.zfloor_lo;
  if (dist0 - c.blk_radius > pp[7]) return;
  const float wwx = B.ex[0] * B.w[0] + B.ey[0] * B.w[1] + B.ez[0] * B.w[2];   // world angular velocity
  const float wwy = B.ex[1] * B.w[0] + B.ey[1] * B.w[1] + B.ez[1] * B.w[2];
  const float wwz = B.ex[2] * B.w[0] + B.ey[2] * B.w[1] + B.ez[2] * B.w[2];
  for (int i = 0; i < 8 && B.nc < 4; i++) {
    const float sx = (i & 1) ? h : -h, sy = (i & 2) ? h : -h, sz = (i & 4) ? h : -h;
    const float cx = B.ex[0] * sx + B.ey[0] * sy + B.ez[0] * sz;
    const float cy = B.ex[1] * sx + B.ey[1] * sy + B.ez[1] * sz;
    const float cz = B.ex[2] * sx + B.ey[2] * sy + B.ez[2] * sz;
    const float dist = dist0 + cz;
    if (dist > pp[7] || cz > 0.f) continue;
    if (dist >= pp[7]) { continue; }            // inside the margin band edge: no constraint row (dist >= includemargin)
    const int k = B.nc++;
    const float rz = cz - 0.5f * dist;
    const float px = B.v[0] + wwy * rz - wwz * cy, py = B.v[1] + wwz * cx - wwx * rz, pz = B.v[2] + wwx * cy - wwy * cx;
    const float imp = imp_of(pp, dist);
    B.cr[k][0] = cx; B.cr[k][1] = cy; B.cr[k][2] = rz;
    B.cD[k] = __fdividef(pp[3] * imp, 1.f - imp);
    B.cy[k][0] = pp[2] * pz + pp[1] * imp * (dist - pp[7]);
    B.cy[k][1] = pp[2] * py;
    B.cy[k][2] = -pp[2] * px;
  }
}

BRB_D unsigned blk_active_set(const BrbModelConsts &c, const Blk &B, const float (&a)[6], unsigned prev, float eps = 2e-4f) {
  unsigned bits = 0;
  const float mu = c.pp[1][0];
  for (int k = 0; k < B.nc; k++) {
    const float rx = B.cr[k][0], ry = B.cr[k][1], rz = B.cr[k][2];
    const float px = a[0] + a[4] * rz - a[5] * ry, py = a[1] + a[5] * rx - a[3] * rz, pz = a[2] + a[3] * ry - a[4] * rx;
    const float z0 = pz + B.cy[k][0], z1 = mu * (py + B.cy[k][1]), z2 = mu * (B.cy[k][2] - px);
    const unsigned pb = prev >> (4 * k);
    const float e0 = (pb & 1u) ? eps : -eps, e1 = (pb & 2u) ? eps : -eps, e2 = (pb & 4u) ? eps : -eps, e3 = (pb & 8u) ? eps : -eps;
    bits |= ((unsigned)(z0 + z1 < e0) | ((unsigned)(z0 - z1 < e1) << 1) | ((unsigned)(z0 + z2 < e2) << 2) | ((unsigned)(z0 - z2 < e3) << 3)) << (4 * k);
  }
  return bits;
}

// block: H = diag(m, I) + sum P' S P (packed lower 6x6), r = f - sum P' S yhat, P = [1 | -[r]x], world coordinates
BRB_D void blk_assemble(const BrbModelConsts &c, const Blk &B, unsigned bits, float (&H)[21], float (&r)[6]) {
#pragma unroll
  for (int k = 0; k < 21; k++) H[k] = 0.f;
  H[LT6(0, 0)] = c.blk_mass; H[LT6(1, 1)] = c.blk_mass; H[LT6(2, 2)] = c.blk_mass;
  H[LT6(3, 3)] = c.blk_inertia; H[LT6(4, 4)] = c.blk_inertia; H[LT6(5, 5)] = c.blk_inertia;
  r[0] = 0.f; r[1] = 0.f; r[2] = -c.blk_mass * c.grav; r[3] = 0.f; r[4] = 0.f; r[5] = 0.f;
  const float mu = c.pp[1][0];
  for (int k = 0; k < B.nc; k++) {
    const unsigned b = (bits >> (4 * k)) & 15u;
    if (!b) continue;
    const float b0 = (float)(b & 1u), b1 = (float)((b >> 1) & 1u), b2 = (float)((b >> 2) & 1u), b3 = (float)((b >> 3) & 1u);
    const float Dc = B.cD[k], Dm = Dc * mu, Dmm = Dm * mu;
    const float Szz = Dc * (b0 + b1 + b2 + b3), Syz = Dm * (b0 - b1), Sxz = -Dm * (b2 - b3), Syy = Dmm * (b0 + b1), Sxx = Dmm * (b2 + b3);
    const float rx = B.cr[k][0], ry = B.cr[k][1], rz = B.cr[k][2];
    const float T3x = Sxz * ry, T3y = -Syy * rz + Syz * ry, T3z = -Syz * rz + Szz * ry;
    const float T4x = Sxx * rz - Sxz * rx, T4y = -Syz * rx, T4z = Sxz * rz - Szz * rx;
    const float T5x = -Sxx * ry, T5y = Syy * rx, T5z = -Sxz * ry + Syz * rx;
    H[LT6(0, 0)] += Sxx; H[LT6(2, 0)] += Sxz; H[LT6(1, 1)] += Syy; H[LT6(2, 1)] += Syz; H[LT6(2, 2)] += Szz;
    H[LT6(3, 0)] += T3x; H[LT6(3, 1)] += T3y; H[LT6(3, 2)] += T3z;
    H[LT6(4, 0)] += T4x; H[LT6(4, 1)] += T4y; H[LT6(4, 2)] += T4z;
    H[LT6(5, 0)] += T5x; H[LT6(5, 1)] += T5y; H[LT6(5, 2)] += T5z;
    H[LT6(3, 3)] += -rz * T3y + ry * T3z;
    H[LT6(4, 3)] += rz * T3x - rx * T3z;
    H[LT6(5, 3)] += -ry * T3x + rx * T3y;
    H[LT6(4, 4)] += rz * T4x - rx * T4z;
    H[LT6(5, 4)] += -ry * T4x + rx * T4y;
    H[LT6(5, 5)] += -ry * T5x + rx * T5y;
    const float y0 = B.cy[k][0], y1 = B.cy[k][1], y2 = B.cy[k][2];
    const float gx = Sxx * y2 - Sxz * y0, gy = -(Syy * y1 + Syz * y0), gz = Sxz * y2 - Syz * y1 - Szz * y0;
    r[0] += gx; r[1] += gy; r[2] += gz;
    r[3] += -rz * gy + ry * gz;
    r[4] += rz * gx - rx * gz;
    r[5] += -ry * gx + rx * gy;
  }
}

// block alone: one Newton step on its floor-contact active set
BRB_D void blk_solve(const BrbModelConsts &c, const Blk &B, unsigned bits, float (&a)[6]) {
  float H[21], r[6];
  blk_assemble(c, B, bits, H, r);
#include "brb_chol6.inc"
#pragma unroll
  for (int k = 0; k < 6; k++) a[k] = r[k];
}

// semi-implicit Euler for the free cube (no joint damping, no actuator: implicitfast reduces to Euler)
BRB_D void blk_finalize(const BrbModelConsts &c, Blk &B, const float (&a)[6]) {
  const float h = c.h;
  B.v[0] += h * a[0]; B.v[1] += h * a[1]; B.v[2] += h * a[2];
  B.w[0] += h * (B.ex[0] * a[3] + B.ex[1] * a[4] + B.ex[2] * a[5]);      // body-frame angular acceleration = R' alpha_w
  B.w[1] += h * (B.ey[0] * a[3] + B.ey[1] * a[4] + B.ey[2] * a[5]);
  B.w[2] += h * (B.ez[0] * a[3] + B.ez[1] * a[4] + B.ez[2] * a[5]);
  B.p[0] += h * B.v[0]; B.p[1] += h * B.v[1]; B.p[2] += h * B.v[2];
  const float qw = B.q[0], qx = B.q[1], qy = B.q[2], qz = B.q[3];
  const float t2 = (h * h) * (B.w[0] * B.w[0] + B.w[1] * B.w[1] + B.w[2] * B.w[2]);
  const float sn = (0.5f * h) * (1.f - t2 * (1.f / 24.f)), cm1 = -(0.125f * t2) * (1.f - t2 * (1.f / 48.f));
  const float ex = sn * B.w[0], ey = sn * B.w[1], ez = sn * B.w[2];
  float nw = qw + (qw * cm1 - (qx * ex + qy * ey + qz * ez));
  float nx = qx + (qx * cm1 + (qw * ex + qy * ez - qz * ey));
  float ny = qy + (qy * cm1 + (qw * ey - qx * ez + qz * ex));
  float nz = qz + (qz * cm1 + (qw * ez + qx * ey - qy * ex));
  const float inv = rsqrtf(nw * nw + nx * nx + ny * ny + nz * nz);
  B.q[0] = nw * inv; B.q[1] = nx * inv; B.q[2] = ny * inv; B.q[3] = nz * inv;
}

// ------------------------------------------------------------------------------------------------ chassis-block (box-box)
// Generic contact record of the coupled path.  Bodies: A (normal points away from it) and B; either may be the robot
// (ra, wheel column) or the block (rb) or the world (absent).
struct GContact {
  float n[3], t1[3], t2[3];     // contact frame (world)
  float ra[3], wa[3];           // robot: point relative to the chassis origin, wheel column (zero for chassis contacts)
  float rb[3];                  // block: point relative to its centre
  float y[3], D, mu;
  int robot_sign, block_sign, wheel;   // +1 body B, -1 body A, 0 absent; wheel = 0/1 or -1
};
#define BRB_MAXGC 16

BRB_D void make_frame3(float *n, float *t1, float *t2) {   // mju_makeFrame with an undefined y axis (A.6)
  t1[0] = 0.f; t1[1] = 0.f; t1[2] = 0.f;
  if (n[1] < 0.5f && n[1] > -0.5f) t1[1] = 1.f; else t1[2] = 1.f;
  const float d = n[0] * t1[0] + n[1] * t1[1] + n[2] * t1[2];
  t1[0] -= d * n[0]; t1[1] -= d * n[1]; t1[2] -= d * n[2];
  const float inv = rsqrtf(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
  t1[0] *= inv; t1[1] *= inv; t1[2] *= inv;
  t2[0] = n[1] * t1[2] - n[2] * t1[1]; t2[1] = n[2] * t1[0] - n[0] * t1[2]; t2[2] = n[0] * t1[1] - n[1] * t1[0];
}

BRB_D float dot3f(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// Box-box in two stages (the oracle's collide_box_box in fp32).
// Stage 1, env03_sat: separating-axis test over the 15 axes on the relative rotation C = A'B (all scalar registers,
//   fully unrolled): r1 + r2 and the centre distance along an axis follow from C, |C| and t = A' dp.  Returns 0 when an
//   axis separates the boxes by more than `margin`, else the axis of least penetration (kind 0/1: face of box 1/2,
//   kind 2: edge i of box 1 x edge j of box 2), its signed direction (1 -> 2) and the separation along it.
// Stage 2, env03_manifold: face clipping (Sutherland-Hodgman against the reference face's side planes) or the closest
//   points of the two edges; <= 8 points.  Run-time indexed arrays -> local memory, not inlined.
BRB_D int env03_sat(const float (&A)[3][3], const float (&h1)[3], const float (&Bx)[3][3], float h2, const float (&dp)[3], float margin,
                    int &kind, int &bi, int &bj, float &best, float (&bestn)[3]) {
  float Cm[3][3], Ca[3][3], t[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    t[i] = dot3f(dp, A[i]);
#pragma unroll
    for (int j = 0; j < 3; j++) { Cm[i][j] = dot3f(A[i], Bx[j]); Ca[i][j] = fabsf(Cm[i][j]); }
  }
  best = -1e30f; kind = -1; bi = 0; bj = 0;
  float sgn = 1.f;
#pragma unroll
  for (int i = 0; i < 3; i++) {          // faces of box 1
    const float s = fabsf(t[i]) - (h1[i] + h2 * (Ca[i][0] + Ca[i][1] + Ca[i][2]));
    if (s > margin) return 0;
    if (s > best) { best = s; kind = 0; bi = i; sgn = t[i] >= 0.f ? 1.f : -1.f; }
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {          // faces of box 2
    const float tl = t[0] * Cm[0][j] + t[1] * Cm[1][j] + t[2] * Cm[2][j];
    const float s = fabsf(tl) - (h1[0] * Ca[0][j] + h1[1] * Ca[1][j] + h1[2] * Ca[2][j] + h2);
    if (s > margin) return 0;
    if (s > best) { best = s; kind = 1; bi = j; sgn = tl >= 0.f ? 1.f : -1.f; }
  }
#pragma unroll
  for (int i = 0; i < 3; i++) {
    constexpr int nxt[3] = {1, 2, 0}, prv[3] = {2, 0, 1};
    const int i1 = nxt[i], i2 = prv[i];
#pragma unroll
    for (int j = 0; j < 3; j++) {        // edge i of box 1 x edge j of box 2
      const int j1 = nxt[j], j2 = prv[j];
      const float n2 = 1.f - Cm[i][j] * Cm[i][j];
      if (n2 < 1e-12f) continue;
      const float inv = rsqrtf(n2);
      const float tl = (t[i2] * Cm[i1][j] - t[i1] * Cm[i2][j]) * inv;
      const float s = fabsf(tl) - (h1[i1] * Ca[i2][j] + h1[i2] * Ca[i1][j] + h2 * (Ca[i][j2] + Ca[i][j1])) * inv;
      if (s > margin) return 0;
      if (s > best + 1e-4f) { best = s; kind = 2; bi = i; bj = j; sgn = tl >= 0.f ? 1.f : -1.f; }
    }
  }
  if (kind == 2) {
    float L[3] = {A[bi][1] * Bx[bj][2] - A[bi][2] * Bx[bj][1], A[bi][2] * Bx[bj][0] - A[bi][0] * Bx[bj][2], A[bi][0] * Bx[bj][1] - A[bi][1] * Bx[bj][0]};
    const float inv = rsqrtf(dot3f(L, L));
#pragma unroll
    for (int k = 0; k < 3; k++) bestn[k] = sgn * inv * L[k];
  } else {
#pragma unroll
    for (int k = 0; k < 3; k++) bestn[k] = sgn * (kind == 0 ? A[bi][k] : Bx[bi][k]);
  }
  return 1;
}

#ifdef BRB_HOST_EMU
static
#else
__device__ __noinline__
#endif
int env03_manifold(const float *p1, const float (*A)[3], const float *h1, const float *p2, const float (*Bx)[3], const float *h2,
                   float margin, int bestkind, int bi, int bj, float best, const float *bestn, float (*pos)[3], float *dist) {
  if (bestkind == 2) {
    float c1[3] = {p1[0], p1[1], p1[2]}, c2[3] = {p2[0], p2[1], p2[2]};
    for (int a = 0; a < 3; a++) {
      if (a != bi) { const float sg = dot3f(A[a], bestn) > 0.f ? 1.f : -1.f; for (int k = 0; k < 3; k++) c1[k] += sg * h1[a] * A[a][k]; }
      if (a != bj) { const float sg = dot3f(Bx[a], bestn) > 0.f ? -1.f : 1.f; for (int k = 0; k < 3; k++) c2[k] += sg * h2[a] * Bx[a][k]; }
    }
    const float *u = A[bi], *v = Bx[bj];
    const float w[3] = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]};
    const float uv = dot3f(u, v), uw = dot3f(u, w), vw = dot3f(v, w), den = 1.f - uv * uv;
    float sa = den > 1e-12f ? (uv * vw - uw) / den : 0.f, sb = den > 1e-12f ? (vw - uv * uw) / den : 0.f;
    sa = fmaxf(-h1[bi], fminf(h1[bi], sa)); sb = fmaxf(-h2[bj], fminf(h2[bj], sb));
    for (int k = 0; k < 3; k++) pos[0][k] = 0.5f * ((c1[k] + sa * u[k]) + (c2[k] + sb * v[k]));
    dist[0] = best;
    return 1;
  }
  const float (*RA)[3] = bestkind == 0 ? A : Bx, (*IB)[3] = bestkind == 0 ? Bx : A;
  const float *hr = bestkind == 0 ? h1 : h2, *hi = bestkind == 0 ? h2 : h1, *pr = bestkind == 0 ? p1 : p2, *pi = bestkind == 0 ? p2 : p1;
  float nref[3];
  for (int k = 0; k < 3; k++) nref[k] = bestkind == 0 ? bestn[k] : -bestn[k];
  int ia = 0; float mind = 1e30f, isg = 1.f;
  for (int a = 0; a < 3; a++) { const float t = dot3f(IB[a], nref); if (-fabsf(t) < mind) { mind = -fabsf(t); ia = a; isg = t > 0.f ? -1.f : 1.f; } }
  const int a1 = (ia + 1) % 3, a2 = (ia + 2) % 3;
  float poly[16][3], tmp[16][3];
  int np = 4;
  for (int cc = 0; cc < 4; cc++) {
    const float s1 = (cc == 0 || cc == 3) ? 1.f : -1.f, s2 = (cc < 2) ? 1.f : -1.f;
    for (int k = 0; k < 3; k++) poly[cc][k] = pi[k] + isg * hi[ia] * IB[ia][k] + s1 * hi[a1] * IB[a1][k] + s2 * hi[a2] * IB[a2][k];
  }
  const int r1 = (bi + 1) % 3, r2 = (bi + 2) % 3;
  for (int side = 0; side < 4; side++) {
    const float *ax = RA[side < 2 ? r1 : r2];
    const float sg = (side & 1) ? -1.f : 1.f, lim = hr[side < 2 ? r1 : r2];
    int nn = 0;
    for (int cc = 0; cc < np; cc++) {
      const float *P = poly[cc], *Q = poly[(cc + 1) % np];
      const float e1[3] = {P[0] - pr[0], P[1] - pr[1], P[2] - pr[2]}, e2[3] = {Q[0] - pr[0], Q[1] - pr[1], Q[2] - pr[2]};
      const float dP = sg * dot3f(e1, ax) - lim, dQ = sg * dot3f(e2, ax) - lim;
      if (dP <= 0.f) { for (int k = 0; k < 3; k++) tmp[nn][k] = P[k]; nn++; }
      if ((dP < 0.f && dQ > 0.f) || (dP > 0.f && dQ < 0.f)) {
        const float t = dP / (dP - dQ);
        for (int k = 0; k < 3; k++) tmp[nn][k] = P[k] + t * (Q[k] - P[k]);
        nn++;
      }
      if (nn >= 15) break;
    }
    np = nn;
    for (int cc = 0; cc < np; cc++) for (int k = 0; k < 3; k++) poly[cc][k] = tmp[cc][k];
    if (np == 0) return 0;
  }
  int cnt = 0;
  for (int cc = 0; cc < np && cnt < 8; cc++) {
    const float rel[3] = {poly[cc][0] - pr[0], poly[cc][1] - pr[1], poly[cc][2] - pr[2]};
    const float depth = dot3f(rel, nref) - hr[bi];
    if (depth > margin) continue;
    for (int k = 0; k < 3; k++) pos[cnt][k] = poly[cc][k] - nref[k] * depth * 0.5f;
    dist[cnt] = depth;
    cnt++;
  }
  return cnt;
}

// Generic dense Newton solve over all 14 dofs (robot: world lin, world ang, wheels; block: world lin, world ang) for the
// substeps in which the block touches the chassis.  Active-set iteration with hysteresis, exact for a fixed set.
#ifdef BRB_HOST_EMU
static
#else
__device__ __noinline__
#endif
int env03_coupled_solve(const BrbModelConsts &c, const Phys &P, const GContact *gc, int ngc, float *acc /*[14]*/) {
  float J[BRB_MAXGC * 3][14];
  for (int k = 0; k < ngc; k++) {
    const GContact &g = gc[k];
    const float *dirs[3] = {g.n, g.t1, g.t2};
    for (int r = 0; r < 3; r++) {
      const float *d = dirs[r];
      float *row = J[3 * k + r];
      for (int j = 0; j < 14; j++) row[j] = 0.f;
      if (g.robot_sign) {
        const float s = (float)g.robot_sign;
        row[0] = s * d[0]; row[1] = s * d[1]; row[2] = s * d[2];
        row[3] = s * (g.ra[1] * d[2] - g.ra[2] * d[1]); row[4] = s * (g.ra[2] * d[0] - g.ra[0] * d[2]); row[5] = s * (g.ra[0] * d[1] - g.ra[1] * d[0]);
        if (g.wheel >= 0) row[6 + g.wheel] = s * dot3f(g.wa, d);
      }
      if (g.block_sign) {
        const float s = (float)g.block_sign;
        row[8] = s * d[0]; row[9] = s * d[1]; row[10] = s * d[2];
        row[11] = s * (g.rb[1] * d[2] - g.rb[2] * d[1]); row[12] = s * (g.rb[2] * d[0] - g.rb[0] * d[2]); row[13] = s * (g.rb[0] * d[1] - g.rb[1] * d[0]);
      }
    }
  }
  // mass matrix in these coordinates (robot block as in phys_solve, block diagonal) and smooth force
  float M[14][14], f[14];
  for (int i = 0; i < 14; i++) for (int j = 0; j < 14; j++) M[i][j] = 0.f;
  {
    M[0][0] = M[1][1] = M[2][2] = c.mass;
    const float kx = c.mcz * P.ez[0], ky = c.mcz * P.ez[1], kz = c.mcz * P.ez[2];
    M[4][0] = kz; M[5][0] = -ky; M[3][1] = -kz; M[5][1] = kx; M[3][2] = ky; M[4][2] = -kx;
    const float dx = c.Ixx - c.Iyy, dz = c.Izz - c.Iyy;
    for (int i = 0; i < 3; i++) for (int j = 0; j <= i; j++) M[3 + i][3 + j] = (i == j ? c.Iyy : 0.f) + dx * P.ex[i] * P.ex[j] + dz * P.ez[i] * P.ez[j];
    for (int j = 0; j < 3; j++) { M[6][3 + j] = -c.Ia * P.ex[j]; M[7][3 + j] = c.Ia * P.ex[j]; }
    M[6][6] = M[7][7] = c.Ia;
    M[8][8] = M[9][9] = M[10][10] = c.blk_mass;
    M[11][11] = M[12][12] = M[13][13] = c.blk_inertia;
    for (int i = 0; i < 14; i++) for (int j = i + 1; j < 14; j++) M[i][j] = M[j][i];
    { float fr[8]; phys_world_force(c, P, fr); for (int k = 0; k < 8; k++) f[k] = fr[k]; }
    f[8] = 0.f; f[9] = 0.f; f[10] = -c.blk_mass * c.grav; f[11] = f[12] = f[13] = 0.f;
  }
  // Newton with an exact line search on the piecewise-quadratic cost (the oracle's algorithm, A.8): start from the
  // unconstrained acceleration, direction p = -H^-1 g on the current active set, full step if it keeps the set (then the
  // point is the exact minimiser), otherwise the root of the increasing piecewise-linear phi'(alpha).
  const int nrow = 4 * ngc;
  float a[14], jar[4 * BRB_MAXGC], jp[4 * BRB_MAXGC], Dr[4 * BRB_MAXGC];
  int nonconv = 0;
  {
    float L[14][14], r[14];
    for (int i = 0; i < 14; i++) { r[i] = f[i]; for (int j = 0; j <= i; j++) L[i][j] = M[i][j]; }
    for (int j = 0; j < 14; j++) {
      float d = L[j][j];
      for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k];
      const float id = rsqrtf(d);
      L[j][j] = id;
      for (int i = j + 1; i < 14; i++) { float t = L[i][j]; for (int k = 0; k < j; k++) t -= L[i][k] * L[j][k]; L[i][j] = t * id; }
    }
    for (int i = 0; i < 14; i++) { float t = r[i]; for (int k = 0; k < i; k++) t -= L[i][k] * r[k]; r[i] = t * L[i][i]; }
    for (int i = 13; i >= 0; i--) { float t = r[i]; for (int k = i + 1; k < 14; k++) t -= L[k][i] * r[k]; r[i] = t * L[i][i]; }
    for (int i = 0; i < 14; i++) a[i] = r[i];
  }
  for (int it = 0;; it++) {
    float H[14][14], g[14], p[14];
    for (int i = 0; i < 14; i++) {
      float t = -f[i];
      for (int j = 0; j < 14; j++) t += M[i][j] * a[j];
      g[i] = t;
      for (int j = 0; j <= i; j++) H[i][j] = M[i][j];
    }
    for (int k = 0; k < ngc; k++) {
      const GContact &gk = gc[k];
      for (int rr = 0; rr < 4; rr++) {
        const float *jt = J[3 * k + 1 + (rr >> 1)], sg = (rr & 1) ? -gk.mu : gk.mu;
        float row[14], z = gk.y[0] + sg * gk.y[1 + (rr >> 1)];
        for (int j = 0; j < 14; j++) { row[j] = J[3 * k][j] + sg * jt[j]; z += row[j] * a[j]; }
        jar[4 * k + rr] = z;
        Dr[4 * k + rr] = gk.D;
        if (z < 0.f) {
          for (int i = 0; i < 14; i++) {
            const float di = gk.D * row[i];
            if (di == 0.f) continue;
            g[i] += di * z;
            for (int j = 0; j <= i; j++) H[i][j] += di * row[j];
          }
        }
      }
    }
    for (int j = 0; j < 14; j++) {   // Cholesky (lower), in place; diagonal keeps 1/L_jj
      float d = H[j][j];
      for (int k = 0; k < j; k++) d -= H[j][k] * H[j][k];
      const float id = rsqrtf(d);
      H[j][j] = id;
      for (int i = j + 1; i < 14; i++) { float t = H[i][j]; for (int k = 0; k < j; k++) t -= H[i][k] * H[j][k]; H[i][j] = t * id; }
    }
    for (int i = 0; i < 14; i++) { float t = -g[i]; for (int k = 0; k < i; k++) t -= H[i][k] * p[k]; p[i] = t * H[i][i]; }
    for (int i = 13; i >= 0; i--) { float t = p[i]; for (int k = i + 1; k < 14; k++) t -= H[k][i] * p[k]; p[i] = t * H[i][i]; }
    bool same = true;
    for (int k = 0; k < ngc; k++) {
      const GContact &gk = gc[k];
      for (int rr = 0; rr < 4; rr++) {
        const float *jt = J[3 * k + 1 + (rr >> 1)], sg = (rr & 1) ? -gk.mu : gk.mu;
        float z = 0.f;
        for (int j = 0; j < 14; j++) z += (J[3 * k][j] + sg * jt[j]) * p[j];
        jp[4 * k + rr] = z;
        const float now = jar[4 * k + rr], nxt = now + z;
        // rows that stay within the hysteresis band of the switching surface do not count as a change
        if ((now < 0.f) != (nxt < 0.f) && fabsf(nxt) > 2e-4f) same = false;
      }
    }
    if (same) { for (int i = 0; i < 14; i++) a[i] += p[i]; break; }
    if (it >= 40) { nonconv = 1; for (int i = 0; i < 14; i++) a[i] += p[i]; break; }
    float pMp = 0.f, pg = 0.f;
    for (int i = 0; i < 14; i++) {
      float t = 0.f, u = -f[i];
      for (int j = 0; j < 14; j++) { t += M[i][j] * p[j]; u += M[i][j] * a[j]; }
      pMp += p[i] * t;
      pg += p[i] * u;
    }
    float lo = 0.f, alpha = 1.f;
    for (int seg = 0; seg <= nrow; seg++) {
      float hi = 3.0e38f;                           // next breakpoint above lo
      for (int r = 0; r < nrow; r++)
        if (jp[r] != 0.f) { const float t = -jar[r] / jp[r]; if (t > lo && t < hi) hi = t; }
      const bool last = hi > 1.0e38f;
      const float mid = last ? lo + 1.f : 0.5f * (lo + hi);
      float c0 = pg, c1 = pMp;
      for (int r = 0; r < nrow; r++)
        if (jar[r] + mid * jp[r] < 0.f) { c0 += Dr[r] * jar[r] * jp[r]; c1 += Dr[r] * jp[r] * jp[r]; }
      const float root = -c0 / c1;
      if (last || root <= hi) { alpha = root < lo ? lo : root; break; }
      lo = hi;
    }
    for (int i = 0; i < 14; i++) a[i] += alpha * p[i];
  }
  for (int i = 0; i < 14; i++) acc[i] = a[i];
  return nonconv;
}

// ------------------------------------------------------------------------------------------------ wheel-block (cylinder-box)
// MuJoCo 3.2.0 sends this pair to libccd's MPR (mjc_Convex), an iterative portal refinement whose result on the curved wheel surface
// is only good to its tolerance; it is not restated.  Own analytic collider, the oracle's algorithm (oracle/brb_ref.c:
// brb_ref_cylinder_box) in fp32: sep(d) = d.(c_box - c_cyl) - h_cyl(d) - h_box(d) is a lower bound of the signed distance for every
// unit d and the distance is its maximum; it is evaluated on the directions at which the maximum can sit -- box face normals, the
// wheel axis, axis x box edge, the radial direction to each box vertex, from the nearest rim point of either cap to each vertex, and
// the best direction perpendicular to each box edge direction (one-dimensional scan + golden-section search, fixed step count).  One contact: normal from the wheel to
// the block, position from two alternating projections between the two support features.  Rare path: not inlined, loops not unrolled.
#define CYLBOX_TAU 0.02f
BRB_D float soft_signf(float x) { return fmaxf(-1.f, fminf(1.f, x / CYLBOX_TAU)); }
// sqrt(max(x, 0)) as x * rsqrt(x): the IEEE sqrtf has an out-of-line slow path (32 CALLs in this function), and the collider runs on one
// lane of a warp in cold code, where every instruction counts
BRB_D float cb_sqrtf(float x) { const float y = fmaxf(x, 1e-30f); return y * rsqrtf(y); }
BRB_D float cyl_box_sep(const float *d, const float *delta, const float *a, float R, float L, const float (*E)[3], float h) {
  const float da = dot3f(d, a);
  return dot3f(d, delta) - L * fabsf(da) - R * cb_sqrtf(1.f - da * da) - h * (fabsf(dot3f(d, E[0])) + fabsf(dot3f(d, E[1])) + fabsf(dot3f(d, E[2])));
}
BRB_D void cyl_box_try(const float *v, bool orient, const float *delta, const float *a, float R, float L, const float (*E)[3], float h, float &best, float *bd) {
  const float n2 = dot3f(v, v);
  if (n2 <= 1e-16f) return;
  float in = rsqrtf(n2);
  if (orient && dot3f(v, delta) < 0.f) in = -in;
  const float t[3] = {v[0] * in, v[1] * in, v[2] * in};
  const float sp = cyl_box_sep(t, delta, a, R, L, E, h);
  if (sp > best) { best = sp; bd[0] = t[0]; bd[1] = t[1]; bd[2] = t[2]; }
}
BRB_D float cyl_box_cos16(int k) {      // cos(2 pi k / 16); sin = entry (k + 12) & 15
  const int q = k & 7;                  // cos(pi - x) = -cos(x): fold to the first half, then to the first quadrant
  const float v = q == 0 ? 1.f : (q == 1 ? 0.92387953251128674f : (q == 2 ? 0.70710678118654752f : (q == 3 ? 0.38268343236508977f :
                  (q == 4 ? 0.f : (q == 5 ? -0.38268343236508977f : (q == 6 ? -0.70710678118654752f : -0.92387953251128674f))))));
  return (k & 8) ? -v : v;
}
BRB_D float cyl_box_g(float c, float s, float D1, float D2, float A1, float A2, float R, float L, float h) {
  const float da = A1 * c + A2 * s;
  return D1 * c + D2 * s - L * fabsf(da) - R * cb_sqrtf(1.f - da * da) - h * (fabsf(c) + fabsf(s));
}
// arguments and result travel BY VALUE: through pointers the box axes were local-memory loads in every one of the ~100 separation
// evaluations (145 us per call, one lane active; the first version of the opt-in path ran at 87 ms per step)
struct CylBoxIn { float cc[3], a[3], b[3], E[3][3], R, L, h, margin; };
struct CylBoxOut { int hit; float dist, n[3], pos[3]; };
#ifdef BRB_HOST_EMU
static
#else
__device__ __noinline__
#endif
CylBoxOut env03_cyl_box(const CylBoxIn in_) {
  const float cc[3] = {in_.cc[0], in_.cc[1], in_.cc[2]}, a[3] = {in_.a[0], in_.a[1], in_.a[2]}, b[3] = {in_.b[0], in_.b[1], in_.b[2]};
  const float E[3][3] = {{in_.E[0][0], in_.E[0][1], in_.E[0][2]}, {in_.E[1][0], in_.E[1][1], in_.E[1][2]}, {in_.E[2][0], in_.E[2][1], in_.E[2][2]}};
  const float R = in_.R, L = in_.L, h = in_.h, margin = in_.margin;
  CylBoxOut out;
  out.hit = 0; out.dist = 0.f;
  for (int k = 0; k < 3; k++) { out.n[k] = 0.f; out.pos[k] = 0.f; }
  const float delta[3] = {b[0] - cc[0], b[1] - cc[1], b[2] - cc[2]};
  float best = -1e30f, bd[3] = {0.f, 0.f, 1.f};
#pragma unroll
  for (int i = 0; i < 3; i++) cyl_box_try(E[i], true, delta, a, R, L, E, h, best, bd);
  cyl_box_try(a, true, delta, a, R, L, E, h, best, bd);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const float x[3] = {a[1] * E[i][2] - a[2] * E[i][1], a[2] * E[i][0] - a[0] * E[i][2], a[0] * E[i][1] - a[1] * E[i][0]};
    if (dot3f(x, x) > 1e-10f) cyl_box_try(x, true, delta, a, R, L, E, h, best, bd);
  }
#pragma unroll 1
  for (int vi = 0; vi < 8; vi++) {
    float u[3];
    for (int k = 0; k < 3; k++) u[k] = delta[k] + ((vi & 1) ? h : -h) * E[0][k] + ((vi & 2) ? h : -h) * E[1][k] + ((vi & 4) ? h : -h) * E[2][k];
    const float ua = dot3f(u, a), up[3] = {u[0] - ua * a[0], u[1] - ua * a[1], u[2] - ua * a[2]};
    const float rho2 = dot3f(up, up);
    if (rho2 <= 1e-16f) continue;
    cyl_box_try(up, false, delta, a, R, L, E, h, best, bd);
    {   // ... and against the rim of the cap on the vertex's side of the wheel (the other rim is farther from it)
      const float sg = ua < 0.f ? -1.f : 1.f, ir = R * rsqrtf(rho2);
      float t[3];
      for (int k = 0; k < 3; k++) t[k] = u[k] - sg * L * a[k] - up[k] * ir;
      cyl_box_try(t, false, delta, a, R, L, E, h, best, bd);
    }
  }
  if (best > margin) return out;     // already separated by more than the margin along one of the cheap directions: the maximum only grows
  // box edge against a rim / the curved side / a cap: the separating direction is perpendicular to the edge, d(phi) = cos(phi) E_a1 +
  // sin(phi) E_a2; sep restricted to that plane is a one-dimensional function (all four parallel edges and both caps at once), maximised
  // by a 16-point scan and a golden-section search with a fixed number of steps inside the best cell (no trigonometry per step)
#pragma unroll
  for (int ax = 0; ax < 3; ax++) {      // unrolled: E[a1], E[a2] stay register-indexed
    const int a1 = (ax + 1) % 3, a2 = (ax + 2) % 3;
    const float D1 = dot3f(E[a1], delta), D2 = dot3f(E[a2], delta), A1 = dot3f(E[a1], a), A2 = dot3f(E[a2], a);
    int kb = 0;
    float gb = -1e30f;
#pragma unroll 1
    for (int k = 0; k < 16; k++) {
      const float g = cyl_box_g(cyl_box_cos16(k), cyl_box_cos16((k + 12) & 15), D1, D2, A1, A2, R, L, h);
      if (g > gb) { gb = g; kb = k; }
    }
    const float ck = cyl_box_cos16(kb), sk = cyl_box_cos16((kb + 12) & 15), tmax = 0.41421356237309503f, gr = 0.6180339887498949f;
    float lo = -tmax, hi = tmax, x1 = hi - gr * (hi - lo), x2 = lo + gr * (hi - lo), f1 = 0.f, f2 = 0.f;
    int fresh = 2;
#pragma unroll 1
    for (int it = 0; it < 10; it++) {
#pragma unroll 1
      for (int w = 0; w < 2; w++) {
        if (fresh != 2 && w != fresh) continue;
        const float t = w ? x2 : x1, in = rsqrtf(1.f + t * t);
        const float g = cyl_box_g((ck - t * sk) * in, (sk + t * ck) * in, D1, D2, A1, A2, R, L, h);
        if (w) f2 = g; else f1 = g;
      }
      if (f1 > f2) { hi = x2; x2 = x1; f2 = f1; x1 = hi - gr * (hi - lo); fresh = 0; }
      else { lo = x1; x1 = x2; f1 = f2; x2 = lo + gr * (hi - lo); fresh = 1; }
    }
    const float t = 0.5f * (lo + hi), in = rsqrtf(1.f + t * t), cf = (ck - t * sk) * in, sf = (sk + t * ck) * in;
    float td[3];
    for (int k = 0; k < 3; k++) td[k] = cf * E[a1][k] + sf * E[a2][k];
    cyl_box_try(td, false, delta, a, R, L, E, h, best, bd);
  }
  if (best > margin) return out;
  // contact point: support features along bd refined by two alternating projections; feature selection blended over 0.02 rad around
  // perpendicular (see the oracle: continuous from the middle of a flat-on-flat patch to its deeper end)
  const float sa = soft_signf(dot3f(bd, a));
  float sb[3], rdir[3] = {0.f, 0.f, 0.f}, wr = 0.f;
  for (int j = 0; j < 3; j++) sb[j] = soft_signf(dot3f(bd, E[j]));
  {
    const float da = dot3f(bd, a), dp[3] = {bd[0] - da * a[0], bd[1] - da * a[1], bd[2] - da * a[2]}, pm = cb_sqrtf(dot3f(dp, dp));
    if (pm > 1e-12f) { for (int k = 0; k < 3; k++) rdir[k] = dp[k] / pm; wr = fminf(1.f, pm / CYLBOX_TAU); }
  }
  float pc[3], qb[3];
  for (int k = 0; k < 3; k++) pc[k] = cc[k] + sa * L * a[k] + wr * R * rdir[k];
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
    for (int k = 0; k < 3; k++) qb[k] = b[k];
    const float rel[3] = {pc[0] - b[0], pc[1] - b[1], pc[2] - b[2]};
    for (int j = 0; j < 3; j++) {
      const float l = -sb[j] * h + (1.f - fabsf(sb[j])) * fmaxf(-h, fminf(h, dot3f(rel, E[j])));
      for (int k = 0; k < 3; k++) qb[k] += l * E[j][k];
    }
    const float rl[3] = {qb[0] - cc[0], qb[1] - cc[1], qb[2] - cc[2]}, ta = dot3f(rl, a);
    const float t = sa * L + (1.f - fabsf(sa)) * fmaxf(-L, fminf(L, ta));
    float rv[3];
    for (int k = 0; k < 3; k++) rv[k] = rl[k] - ta * a[k];
    { const float n2 = dot3f(rv, rv); if (n2 > R * R) { const float in = R * rsqrtf(n2); for (int k = 0; k < 3; k++) rv[k] *= in; } }
    for (int k = 0; k < 3; k++) pc[k] = cc[k] + t * a[k] + wr * R * rdir[k] + (1.f - wr) * rv[k];
  }
  out.hit = 1;
  out.dist = best;
  for (int k = 0; k < 3; k++) { out.n[k] = bd[k]; out.pos[k] = 0.5f * (pc[k] + qb[k]); }
  return out;
}

// Wheel-block contacts of the current substep: at most one per wheel, each with its own frame and the wheel's column of the point map
// (the contact point rides on the spinning wheel).  Run-time indexed -> local memory; only touched while a wheel is within reach of the block.
struct WBSet {
  float n[2][3], t1[2][3], t2[2][3];
  float ra[2][3], rb[2][3], w[2][3], y[2][3], D[2];
  int wheel[2];
  int nw;
  unsigned bits;      // 4 pyramid rows per contact
  bool near;          // the narrow phase ran for a wheel this substep
};

BRB_D void wb_setup(const BrbModelConsts &c, const Phys &P, const Blk &B, WBSet &W) {
  W.nw = 0;
  W.near = false;
  const float *pp = c.pp[2];
  const float reach = c.blk_radius + sqrtf(c.rad * c.rad + c.hl * c.hl) + pp[7];
#pragma unroll 1
  for (int k = 0; k < 2; k++) {
    const float sg = k ? 1.f : -1.f;
    float cw[3];
    for (int j = 0; j < 3; j++) cw[j] = P.p[j].s + P.ex[j] * (sg * c.ox) + P.ez[j] * c.oz;
    const float dx = B.p[0] - cw[0], dy = B.p[1] - cw[1], dz = B.p[2] - cw[2];
    if (dx * dx + dy * dy + dz * dz > reach * reach) continue;
    const float ax[3] = {P.ex[0], P.ex[1], P.ex[2]};
    const float Bx[3][3] = {{B.ex[0], B.ex[1], B.ex[2]}, {B.ey[0], B.ey[1], B.ey[2]}, {B.ez[0], B.ez[1], B.ez[2]}};
    {
      // mid phase: separated by more than the margin along a box face normal or the wheel axis?  (four of the collider's own candidate
      // directions: if one of them already exceeds the margin so does the maximum, i.e. the collider would return "no contact")
      const float dl[3] = {dx, dy, dz};
      bool sepd = false;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const float *Lq = q < 3 ? Bx[q] : ax;
        const float la = dot3f(Lq, ax);
        const float sc = c.hl * fabsf(la) + c.rad * sqrtf(fmaxf(0.f, 1.f - la * la));
        const float sbx = c.blk_half * (fabsf(dot3f(Lq, Bx[0])) + fabsf(dot3f(Lq, Bx[1])) + fabsf(dot3f(Lq, Bx[2])));
        sepd = sepd || (fabsf(dot3f(Lq, dl)) - sc - sbx > pp[7]);
      }
      if (sepd) continue;
    }
    W.near = true;
    CylBoxIn ci;
    for (int j = 0; j < 3; j++) { ci.cc[j] = cw[j]; ci.a[j] = ax[j]; ci.b[j] = B.p[j]; ci.E[0][j] = Bx[0][j]; ci.E[1][j] = Bx[1][j]; ci.E[2][j] = Bx[2][j]; }
    ci.R = c.rad; ci.L = c.hl; ci.h = c.blk_half; ci.margin = pp[7];
    const CylBoxOut co = env03_cyl_box(ci);
    if (!co.hit || co.dist >= pp[7]) continue;
    const float dist = co.dist;
    float nn[3] = {co.n[0], co.n[1], co.n[2]};
    const float pos[3] = {co.pos[0], co.pos[1], co.pos[2]};
    const int m = W.nw++;
    float t1[3], t2[3];
    make_frame3(nn, t1, t2);
    const float wr[3] = {P.ex[0] * P.w[0].s + P.ey[0] * P.w[1].s + P.ez[0] * P.w[2].s, P.ex[1] * P.w[0].s + P.ey[1] * P.w[1].s + P.ez[1] * P.w[2].s,
                         P.ex[2] * P.w[0].s + P.ey[2] * P.w[1].s + P.ez[2] * P.w[2].s};
    const float wb[3] = {B.ex[0] * B.w[0] + B.ey[0] * B.w[1] + B.ez[0] * B.w[2], B.ex[1] * B.w[0] + B.ey[1] * B.w[1] + B.ez[1] * B.w[2],
                         B.ex[2] * B.w[0] + B.ey[2] * B.w[1] + B.ez[2] * B.w[2]};
    float ra[3], rb[3], rc[3];
    for (int j = 0; j < 3; j++) { ra[j] = pos[j] - P.p[j].s; rb[j] = pos[j] - B.p[j]; rc[j] = pos[j] - cw[j]; }
    // wheel column: velocity of the material point per unit wheel speed = axis_k x (r - wheel centre), axis_k = sg ex
    const float wv[3] = {sg * (ax[1] * rc[2] - ax[2] * rc[1]), sg * (ax[2] * rc[0] - ax[0] * rc[2]), sg * (ax[0] * rc[1] - ax[1] * rc[0])};
    const float sk = P.s[k].s;
    const float dv[3] = {(B.v[0] + wb[1] * rb[2] - wb[2] * rb[1]) - (P.v[0].s + wr[1] * ra[2] - wr[2] * ra[1] + sk * wv[0]),
                         (B.v[1] + wb[2] * rb[0] - wb[0] * rb[2]) - (P.v[1].s + wr[2] * ra[0] - wr[0] * ra[2] + sk * wv[1]),
                         (B.v[2] + wb[0] * rb[1] - wb[1] * rb[0]) - (P.v[2].s + wr[0] * ra[1] - wr[1] * ra[0] + sk * wv[2])};
    const float imp = imp_of(pp, dist);
    W.D[m] = __fdividef(c.wb_D1 * imp, 1.f - imp);
    W.wheel[m] = k;
    for (int j = 0; j < 3; j++) { W.n[m][j] = nn[j]; W.t1[m][j] = t1[j]; W.t2[m][j] = t2[j]; W.ra[m][j] = ra[j]; W.rb[m][j] = rb[j]; W.w[m][j] = wv[j]; }
    W.y[m][0] = pp[2] * dot3f(nn, dv) + pp[1] * imp * (dist - pp[7]);
    W.y[m][1] = pp[2] * dot3f(t1, dv);
    W.y[m][2] = pp[2] * dot3f(t2, dv);
  }
}

BRB_D unsigned wb_active_set(const BrbModelConsts &c, const WBSet &W, const float (&ar)[8], const float (&ab)[6], unsigned prev, float eps = 2e-4f) {
  unsigned bits = 0;
  const float mu = c.pp[2][0];
  for (int k = 0; k < W.nw; k++) {
    const float *ra = W.ra[k], *rb = W.rb[k], *wv = W.w[k];
    const float aw = W.wheel[k] ? ar[7] : ar[6];
    const float dx = (ab[0] + ab[4] * rb[2] - ab[5] * rb[1]) - (ar[0] + ar[4] * ra[2] - ar[5] * ra[1] + aw * wv[0]);
    const float dy = (ab[1] + ab[5] * rb[0] - ab[3] * rb[2]) - (ar[1] + ar[5] * ra[0] - ar[3] * ra[2] + aw * wv[1]);
    const float dz = (ab[2] + ab[3] * rb[1] - ab[4] * rb[0]) - (ar[2] + ar[3] * ra[1] - ar[4] * ra[0] + aw * wv[2]);
    const float z0 = W.n[k][0] * dx + W.n[k][1] * dy + W.n[k][2] * dz + W.y[k][0];
    const float z1 = mu * (W.t1[k][0] * dx + W.t1[k][1] * dy + W.t1[k][2] * dz + W.y[k][1]);
    const float z2 = mu * (W.t2[k][0] * dx + W.t2[k][1] * dy + W.t2[k][2] * dz + W.y[k][2]);
    const unsigned pb = prev >> (4 * k);
    const float e0 = (pb & 1u) ? eps : -eps, e1 = (pb & 2u) ? eps : -eps, e2 = (pb & 4u) ? eps : -eps, e3 = (pb & 8u) ? eps : -eps;
    bits |= ((unsigned)(z0 + z1 < e0) | ((unsigned)(z0 - z1 < e1) << 1) | ((unsigned)(z0 + z2 < e2) << 2) | ((unsigned)(z0 - z2 < e3) << 3)) << (4 * k);
  }
  return bits;
}

// ------------------------------------------------------------------------------------------------ substep driver
struct Env03Stats { unsigned coupled, blk_contact, unsupported, fallback, csolves, coupled_last, blk_last, wb_last; };

// gathers every contact of the substep into generic records and runs the coupled solve
BRB_D int env03_coupled_substep(const BrbModelConsts &c, const Phys &P, const Blk &B, const float (*bpos)[3], const float *bdist,
                                const float *bn, int nbb, const WBSet &W, float *acc) {
  GContact gc[BRB_MAXGC];
  int n = 0;
  // wheel-floor contacts of the robot (already set up in P): frame = world axes (n = z, t1 = y, t2 = -x)
  for (int ci = 0; ci < 4; ci++) {
    if (!(P.valid & (1u << ci))) continue;
    GContact &g = gc[n++];
    g.n[0] = 0.f; g.n[1] = 0.f; g.n[2] = 1.f; g.t1[0] = 0.f; g.t1[1] = 1.f; g.t1[2] = 0.f; g.t2[0] = -1.f; g.t2[1] = 0.f; g.t2[2] = 0.f;
    for (int k = 0; k < 3; k++) { g.ra[k] = P.cr[ci][k]; g.wa[k] = P.cw[ci][k]; g.rb[k] = 0.f; g.y[k] = P.cy[ci][k]; }
    g.D = P.cD[ci]; g.mu = c.pp[0][0]; g.robot_sign = 1; g.block_sign = 0; g.wheel = ci >> 1;
  }
  for (int k2 = 0; k2 < B.nc; k2++) {
    GContact &g = gc[n++];
    g.n[0] = 0.f; g.n[1] = 0.f; g.n[2] = 1.f; g.t1[0] = 0.f; g.t1[1] = 1.f; g.t1[2] = 0.f; g.t2[0] = -1.f; g.t2[1] = 0.f; g.t2[2] = 0.f;
    for (int k = 0; k < 3; k++) { g.rb[k] = B.cr[k2][k]; g.ra[k] = 0.f; g.wa[k] = 0.f; g.y[k] = B.cy[k2][k]; }
    g.D = B.cD[k2]; g.mu = c.pp[1][0]; g.robot_sign = 0; g.block_sign = 1; g.wheel = -1;
  }
  // chassis-block contacts: body A = robot (chassis), body B = block; normal from chassis to block
  const float *pp = c.pp[2];
  const float wr[3] = {P.ex[0] * P.w[0].s + P.ey[0] * P.w[1].s + P.ez[0] * P.w[2].s, P.ex[1] * P.w[0].s + P.ey[1] * P.w[1].s + P.ez[1] * P.w[2].s,
                       P.ex[2] * P.w[0].s + P.ey[2] * P.w[1].s + P.ez[2] * P.w[2].s};
  const float wb[3] = {B.ex[0] * B.w[0] + B.ey[0] * B.w[1] + B.ez[0] * B.w[2], B.ex[1] * B.w[0] + B.ey[1] * B.w[1] + B.ez[1] * B.w[2],
                       B.ex[2] * B.w[0] + B.ey[2] * B.w[1] + B.ez[2] * B.w[2]};
  const float rp[3] = {P.p[0].s, P.p[1].s, P.p[2].s};
  for (int k2 = 0; k2 < nbb && n < BRB_MAXGC; k2++) {
    if (bdist[k2] >= pp[7]) continue;
    GContact &g = gc[n++];
    for (int k = 0; k < 3; k++) g.n[k] = bn[k];
    make_frame3(g.n, g.t1, g.t2);
    for (int k = 0; k < 3; k++) { g.ra[k] = bpos[k2][k] - rp[k]; g.rb[k] = bpos[k2][k] - B.p[k]; g.wa[k] = 0.f; }
    const float va[3] = {P.v[0].s + wr[1] * g.ra[2] - wr[2] * g.ra[1], P.v[1].s + wr[2] * g.ra[0] - wr[0] * g.ra[2], P.v[2].s + wr[0] * g.ra[1] - wr[1] * g.ra[0]};
    const float vb[3] = {B.v[0] + wb[1] * g.rb[2] - wb[2] * g.rb[1], B.v[1] + wb[2] * g.rb[0] - wb[0] * g.rb[2], B.v[2] + wb[0] * g.rb[1] - wb[1] * g.rb[0]};
    const float dv[3] = {vb[0] - va[0], vb[1] - va[1], vb[2] - va[2]};
    const float imp = imp_of(pp, bdist[k2]);
    g.D = pp[3] * imp / (1.f - imp); g.mu = pp[0];
    g.y[0] = pp[2] * dot3f(g.n, dv) + pp[1] * imp * (bdist[k2] - pp[7]);
    g.y[1] = pp[2] * dot3f(g.t1, dv);
    g.y[2] = pp[2] * dot3f(g.t2, dv);
    g.robot_sign = -1; g.block_sign = 1; g.wheel = -1;
  }
  // wheel-block contacts: body A = robot (wheel: its column rides along), body B = block
  for (int k2 = 0; k2 < W.nw && n < BRB_MAXGC; k2++) {
    GContact &g = gc[n++];
    for (int k = 0; k < 3; k++) { g.n[k] = W.n[k2][k]; g.t1[k] = W.t1[k2][k]; g.t2[k] = W.t2[k2][k]; g.ra[k] = W.ra[k2][k]; g.rb[k] = W.rb[k2][k];
                                  g.wa[k] = W.w[k2][k]; g.y[k] = W.y[k2][k]; }
    g.D = W.D[k2]; g.mu = c.pp[2][0]; g.robot_sign = -1; g.block_sign = 1; g.wheel = W.wheel[k2];
  }
  return env03_coupled_solve(c, P, gc, n, acc);
}

// ------------------------------------------------------------------------------------------------ fast coupled path
// Chassis-block contacts of the current substep (general contact frame).  Indexed at run time -> local memory; the
// accumulators they feed (H, Hb, Cc) are indexed at compile time and stay in registers.
struct CBSet {
  float n[3], t1[3], t2[3];          // one contact frame for the whole box-box manifold (registers)
  float ra[8][3], rb[8][3], y[8][3], D[8];   // per contact (run-time index: local memory)
  int nc;
  unsigned bits;   // 4 pyramid rows per contact
};

BRB_D void cb_setup(const BrbModelConsts &c, const Phys &P, const Blk &B, const float (*bpos)[3], const float *bdist, const float *bn,
                    int nbb, CBSet &Q) {
  const float *pp = c.pp[2];
  const float wr[3] = {P.ex[0] * P.w[0].s + P.ey[0] * P.w[1].s + P.ez[0] * P.w[2].s, P.ex[1] * P.w[0].s + P.ey[1] * P.w[1].s + P.ez[1] * P.w[2].s,
                       P.ex[2] * P.w[0].s + P.ey[2] * P.w[1].s + P.ez[2] * P.w[2].s};
  const float wb[3] = {B.ex[0] * B.w[0] + B.ey[0] * B.w[1] + B.ez[0] * B.w[2], B.ex[1] * B.w[0] + B.ey[1] * B.w[1] + B.ez[1] * B.w[2],
                       B.ex[2] * B.w[0] + B.ey[2] * B.w[1] + B.ez[2] * B.w[2]};
  float nn[3] = {bn[0], bn[1], bn[2]}, t1[3], t2[3];
  make_frame3(nn, t1, t2);
  for (int k = 0; k < 3; k++) { Q.n[k] = nn[k]; Q.t1[k] = t1[k]; Q.t2[k] = t2[k]; }
  int n = 0;
  for (int k2 = 0; k2 < nbb && n < 8; k2++) {
    if (bdist[k2] >= pp[7]) continue;
    float ra[3], rb[3];
    for (int k = 0; k < 3; k++) { ra[k] = bpos[k2][k] - P.p[k].s; rb[k] = bpos[k2][k] - B.p[k]; }
    const float dv[3] = {(B.v[0] + wb[1] * rb[2] - wb[2] * rb[1]) - (P.v[0].s + wr[1] * ra[2] - wr[2] * ra[1]),
                         (B.v[1] + wb[2] * rb[0] - wb[0] * rb[2]) - (P.v[1].s + wr[2] * ra[0] - wr[0] * ra[2]),
                         (B.v[2] + wb[0] * rb[1] - wb[1] * rb[0]) - (P.v[2].s + wr[0] * ra[1] - wr[1] * ra[0])};
    const float imp = imp_of(pp, bdist[k2]);
    Q.D[n] = __fdividef(pp[3] * imp, 1.f - imp);
    for (int k = 0; k < 3; k++) { Q.ra[n][k] = ra[k]; Q.rb[n][k] = rb[k]; }
    Q.y[n][0] = pp[2] * dot3f(nn, dv) + pp[1] * imp * (bdist[k2] - pp[7]);
    Q.y[n][1] = pp[2] * dot3f(t1, dv);
    Q.y[n][2] = pp[2] * dot3f(t2, dv);
    n++;
  }
  Q.nc = n;
}

BRB_D unsigned cb_active_set(const BrbModelConsts &c, const CBSet &Q, const float (&ar)[8], const float (&ab)[6], unsigned prev, float eps = 2e-4f) {
  unsigned bits = 0;
  const float mu = c.pp[2][0];
  for (int k = 0; k < Q.nc; k++) {
    const float *ra = Q.ra[k], *rb = Q.rb[k];
    const float dx = (ab[0] + ab[4] * rb[2] - ab[5] * rb[1]) - (ar[0] + ar[4] * ra[2] - ar[5] * ra[1]);
    const float dy = (ab[1] + ab[5] * rb[0] - ab[3] * rb[2]) - (ar[1] + ar[5] * ra[0] - ar[3] * ra[2]);
    const float dz = (ab[2] + ab[3] * rb[1] - ab[4] * rb[0]) - (ar[2] + ar[3] * ra[1] - ar[4] * ra[0]);
    const float z0 = Q.n[0] * dx + Q.n[1] * dy + Q.n[2] * dz + Q.y[k][0];
    const float z1 = mu * (Q.t1[0] * dx + Q.t1[1] * dy + Q.t1[2] * dz + Q.y[k][1]);
    const float z2 = mu * (Q.t2[0] * dx + Q.t2[1] * dy + Q.t2[2] * dz + Q.y[k][2]);
    const unsigned pb = prev >> (4 * k);
    const float e0 = (pb & 1u) ? eps : -eps, e1 = (pb & 2u) ? eps : -eps, e2 = (pb & 4u) ? eps : -eps, e3 = (pb & 8u) ? eps : -eps;
    bits |= ((unsigned)(z0 + z1 < e0) | ((unsigned)(z0 - z1 < e1) << 1) | ((unsigned)(z0 + z2 < e2) << 2) | ((unsigned)(z0 - z2 < e3) << 3)) << (4 * k);
  }
  return bits;
}

// One Newton step of the coupled robot+block system on the given active sets: assemble the robot's 8x8 (H, r), the
// block's 6x6 (Hb, rb) and the 6x6 coupling Cc from the chassis-block contacts, eliminate the block (brb_schur6.inc),
// solve the reduced 8x8 (brb_chol8.inc), back-substitute the block.  Exact for fixed active sets (A.8).
// WB = true additionally assembles the wheel-block contacts of W (own frame per contact, wheel column in the robot's point map, so the
// coupling block has rows for the wheel dofs: Cc[8][6], brb_schur6w.inc); that instance is not inlined (coupled_solve_wb): rare path.
template <bool WB>
BRB_D void coupled_solve_impl(const BrbModelConsts &c, const Phys &P, const Blk &B, const CBSet &Q, const WBSet &W, float (&ar)[8], float (&ab)[6]) {
  float H[36], r[8], Hb[21], rb[6], Cc[8][6];     // rows 6, 7 of Cc exist only for WB (never touched otherwise)
  phys_assemble<true>(c, P, P.valid ? P.bits : 0u, H, r);
  blk_assemble(c, B, B.bits, Hb, rb);
#pragma unroll
  for (int i = 0; i < (WB ? 8 : 6); i++)
#pragma unroll
    for (int j = 0; j < 6; j++) Cc[i][j] = 0.f;
  const float mu = c.pp[2][0];
  for (int k = 0; k < Q.nc; k++) {
    const unsigned b = (Q.bits >> (4 * k)) & 15u;
    if (!b) continue;
    const float b0 = (float)(b & 1u), b1 = (float)((b >> 1) & 1u), b2 = (float)((b >> 2) & 1u), b3 = (float)((b >> 3) & 1u);
    const float Dc = Q.D[k], Dm = Dc * mu, Dmm = Dm * mu;
    const float W00 = Dc * (b0 + b1 + b2 + b3), W01 = Dm * (b0 - b1), W02 = Dm * (b2 - b3), W11 = Dmm * (b0 + b1), W22 = Dmm * (b2 + b3);
    const float *n = Q.n, *t1 = Q.t1, *t2 = Q.t2;
    float g0[3], g1[3], g2[3], S[3][3];
#pragma unroll
    for (int j = 0; j < 3; j++) { g0[j] = W00 * n[j] + W01 * t1[j] + W02 * t2[j]; g1[j] = W01 * n[j] + W11 * t1[j]; g2[j] = W02 * n[j] + W22 * t2[j]; }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) S[i][j] = n[i] * g0[j] + t1[i] * g1[j] + t2[i] * g2[j];
    const float y0 = Q.y[k][0], y1 = Q.y[k][1], y2 = Q.y[k][2];
    const float u0 = W00 * y0 + W01 * y1 + W02 * y2, u1 = W01 * y0 + W11 * y1, u2 = W02 * y0 + W22 * y2;
    float gv[3];
#pragma unroll
    for (int j = 0; j < 3; j++) gv[j] = u0 * n[j] + u1 * t1[j] + u2 * t2[j];
    const float ax = Q.ra[k][0], ay = Q.ra[k][1], az = Q.ra[k][2], bx = Q.rb[k][0], by = Q.rb[k][1], bz = Q.rb[k][2];
    // T_j = S c_j with c_x = (0,-rz,ry), c_y = (rz,0,-rx), c_z = (-ry,rx,0), for the robot point (Ta) and the block point (Tb)
    float Ta[3][3], Tb[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      Ta[0][i] = -az * S[i][1] + ay * S[i][2]; Ta[1][i] = az * S[i][0] - ax * S[i][2]; Ta[2][i] = -ay * S[i][0] + ax * S[i][1];
      Tb[0][i] = -bz * S[i][1] + by * S[i][2]; Tb[1][i] = bz * S[i][0] - bx * S[i][2]; Tb[2][i] = -by * S[i][0] + bx * S[i][1];
    }
    // c_i . v helpers
#define CDOT_A(i, v) ((i) == 0 ? (-az * (v)[1] + ay * (v)[2]) : (i) == 1 ? (az * (v)[0] - ax * (v)[2]) : (-ay * (v)[0] + ax * (v)[1]))
#define CDOT_B(i, v) ((i) == 0 ? (-bz * (v)[1] + by * (v)[2]) : (i) == 1 ? (bz * (v)[0] - bx * (v)[2]) : (-by * (v)[0] + bx * (v)[1]))
    // robot block (J = -P_ra): + P_ra' S P_ra ; rhs + P_ra' gv
    H[LT(0, 0)] += S[0][0]; H[LT(1, 0)] += S[1][0]; H[LT(2, 0)] += S[2][0]; H[LT(1, 1)] += S[1][1]; H[LT(2, 1)] += S[2][1]; H[LT(2, 2)] += S[2][2];
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
      for (int j = 0; j < 3; j++) H[LT(3 + i, j)] += Ta[i][j];
#pragma unroll
      for (int j = 0; j <= i; j++) H[LT(3 + i, 3 + j)] += CDOT_A(i, Ta[j]);
      r[i] += gv[i];
      r[3 + i] += CDOT_A(i, gv);
    }
    // block (J = +P_rb): + P_rb' S P_rb ; rhs - P_rb' gv
    Hb[LT6(0, 0)] += S[0][0]; Hb[LT6(1, 0)] += S[1][0]; Hb[LT6(2, 0)] += S[2][0]; Hb[LT6(1, 1)] += S[1][1]; Hb[LT6(2, 1)] += S[2][1]; Hb[LT6(2, 2)] += S[2][2];
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
      for (int j = 0; j < 3; j++) Hb[LT6(3 + i, j)] += Tb[i][j];
#pragma unroll
      for (int j = 0; j <= i; j++) Hb[LT6(3 + i, 3 + j)] += CDOT_B(i, Tb[j]);
      rb[i] -= gv[i];
      rb[3 + i] -= CDOT_B(i, gv);
    }
    // coupling: Cc[robot dof][block dof] = -(P_ra' S P_rb)
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        Cc[i][j] -= S[i][j];
        Cc[i][3 + j] -= Tb[j][i];
        Cc[3 + i][j] -= Ta[i][j];
        Cc[3 + i][3 + j] -= CDOT_A(i, Tb[j]);
      }
#undef CDOT_A
#undef CDOT_B
  }
  if (WB) {
    for (int k = 0; k < W.nw; k++) {
      const unsigned b = (W.bits >> (4 * k)) & 15u;
      if (!b) continue;
      const float b0 = (float)(b & 1u), b1 = (float)((b >> 1) & 1u), b2 = (float)((b >> 2) & 1u), b3 = (float)((b >> 3) & 1u);
      const float Dc = W.D[k], Dm = Dc * mu, Dmm = Dm * mu;
      const float W00 = Dc * (b0 + b1 + b2 + b3), W01 = Dm * (b0 - b1), W02 = Dm * (b2 - b3), W11 = Dmm * (b0 + b1), W22 = Dmm * (b2 + b3);
      const float *n = W.n[k], *t1 = W.t1[k], *t2 = W.t2[k];
      float g0[3], g1[3], g2[3], S[3][3];
      for (int j = 0; j < 3; j++) { g0[j] = W00 * n[j] + W01 * t1[j] + W02 * t2[j]; g1[j] = W01 * n[j] + W11 * t1[j]; g2[j] = W02 * n[j] + W22 * t2[j]; }
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) S[i][j] = n[i] * g0[j] + t1[i] * g1[j] + t2[i] * g2[j];
      const float y0 = W.y[k][0], y1 = W.y[k][1], y2 = W.y[k][2];
      const float u0 = W00 * y0 + W01 * y1 + W02 * y2, u1 = W01 * y0 + W11 * y1, u2 = W02 * y0 + W22 * y2;
      float gv[3];
      for (int j = 0; j < 3; j++) gv[j] = u0 * n[j] + u1 * t1[j] + u2 * t2[j];
      // point maps: robot P_a = [1 | -[ra]x | wv e_kw'], block P_b = [1 | -[rb]x]; columns c_0 = (0,-z,y), c_1 = (z,0,-x), c_2 = (-y,x,0)
      float ca[4][3], cb[3][3];          // robot: the three angular columns and the wheel column; block: the three angular columns
      {
        const float ax = W.ra[k][0], ay = W.ra[k][1], az = W.ra[k][2], bx = W.rb[k][0], by = W.rb[k][1], bz = W.rb[k][2];
        ca[0][0] = 0.f; ca[0][1] = -az; ca[0][2] = ay;  ca[1][0] = az; ca[1][1] = 0.f; ca[1][2] = -ax;  ca[2][0] = -ay; ca[2][1] = ax; ca[2][2] = 0.f;
        cb[0][0] = 0.f; cb[0][1] = -bz; cb[0][2] = by;  cb[1][0] = bz; cb[1][1] = 0.f; cb[1][2] = -bx;  cb[2][0] = -by; cb[2][1] = bx; cb[2][2] = 0.f;
        for (int j = 0; j < 3; j++) ca[3][j] = W.w[k][j];
      }
      float Sa[4][3], Sb[3][3];          // S c for every non-trivial column
      for (int q = 0; q < 4; q++)
        for (int i = 0; i < 3; i++) Sa[q][i] = S[i][0] * ca[q][0] + S[i][1] * ca[q][1] + S[i][2] * ca[q][2];
      for (int q = 0; q < 3; q++)
        for (int i = 0; i < 3; i++) Sb[q][i] = S[i][0] * cb[q][0] + S[i][1] * cb[q][1] + S[i][2] * cb[q][2];
      const bool right = W.wheel[k] != 0;
      // robot block: + P_a' S P_a, rhs + P_a' gv   (dof order: lin 0-2, ang 3-5, wheel kw)
      H[LT(0, 0)] += S[0][0]; H[LT(1, 0)] += S[1][0]; H[LT(2, 0)] += S[2][0]; H[LT(1, 1)] += S[1][1]; H[LT(2, 1)] += S[2][1]; H[LT(2, 2)] += S[2][2];
      for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) H[LT(3 + i, j)] += Sa[i][j];
        for (int j = 0; j <= i; j++) H[LT(3 + i, 3 + j)] += dot3f(ca[i], Sa[j]);
        r[i] += gv[i];
        r[3 + i] += dot3f(ca[i], gv);
      }
      {
        float hw[7];
        for (int j = 0; j < 3; j++) { hw[j] = Sa[3][j]; hw[3 + j] = dot3f(ca[j], Sa[3]); }
        hw[6] = dot3f(ca[3], Sa[3]);
        const float rw = dot3f(ca[3], gv);
        if (right) { for (int j = 0; j < 6; j++) H[LT(7, j)] += hw[j]; H[LT(7, 7)] += hw[6]; r[7] += rw; }
        else { for (int j = 0; j < 6; j++) H[LT(6, j)] += hw[j]; H[LT(6, 6)] += hw[6]; r[6] += rw; }
      }
      // block: + P_b' S P_b, rhs - P_b' gv
      Hb[LT6(0, 0)] += S[0][0]; Hb[LT6(1, 0)] += S[1][0]; Hb[LT6(2, 0)] += S[2][0]; Hb[LT6(1, 1)] += S[1][1]; Hb[LT6(2, 1)] += S[2][1]; Hb[LT6(2, 2)] += S[2][2];
      for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) Hb[LT6(3 + i, j)] += Sb[i][j];
        for (int j = 0; j <= i; j++) Hb[LT6(3 + i, 3 + j)] += dot3f(cb[i], Sb[j]);
        rb[i] -= gv[i];
        rb[3 + i] -= dot3f(cb[i], gv);
      }
      // coupling: Cc[robot dof][block dof] = -(P_a' S P_b)
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
          Cc[i][j] -= S[i][j];
          Cc[i][3 + j] -= Sb[j][i];
          Cc[3 + i][j] -= Sa[i][j];
          Cc[3 + i][3 + j] -= dot3f(ca[i], Sb[j]);
        }
      for (int j = 0; j < 3; j++) {
        const float cl = Sa[3][j], cg = dot3f(ca[3], Sb[j]);
        if (right) { Cc[7][j] -= cl; Cc[7][3 + j] -= cg; } else { Cc[6][j] -= cl; Cc[6][3 + j] -= cg; }
      }
    }
#define BRB_SCHUR6_ELIMINATE
#include "brb_schur6w.inc"
#undef BRB_SCHUR6_ELIMINATE
#include "brb_chol8.inc"
#define BRB_SCHUR6_BACKSUB
#include "brb_schur6w.inc"
#undef BRB_SCHUR6_BACKSUB
  } else {
#define BRB_SCHUR6_ELIMINATE
#include "brb_schur6.inc"
#undef BRB_SCHUR6_ELIMINATE
#include "brb_chol8.inc"
#define BRB_SCHUR6_BACKSUB
#include "brb_schur6.inc"
#undef BRB_SCHUR6_BACKSUB
  }
#pragma unroll
  for (int k = 0; k < 8; k++) ar[k] = r[k];
#pragma unroll
  for (int k = 0; k < 6; k++) ab[k] = rb[k];
}
BRB_D void coupled_solve_fast(const BrbModelConsts &c, const Phys &P, const Blk &B, const CBSet &Q, const WBSet &W, float (&ar)[8], float (&ab)[6]) {
  coupled_solve_impl<false>(c, P, B, Q, W, ar, ab);
}
#ifdef BRB_HOST_EMU
static __attribute__((noinline))
#else
__device__ __noinline__
#endif
void coupled_solve_wb(const BrbModelConsts &c, const Phys &P, const Blk &B, const CBSet &Q, const WBSet &W, float (&ar)[8], float (&ab)[6]) {
  coupled_solve_impl<true>(c, P, B, Q, W, ar, ab);
}

// does the block touch the chassis box this substep?  bounding spheres first, then the SAT collider
BRB_D int env03_detect(const BrbModelConsts &c, const Phys &P, const Blk &B, float (*bpos)[3], float *bdist, float *bn) {
  float pc[3];
#pragma unroll
  for (int k = 0; k < 3; k++) pc[k] = P.p[k].s + P.ex[k] * c.chassis_pos[0] + P.ey[k] * c.chassis_pos[1] + P.ez[k] * c.chassis_pos[2];
  const float dx = B.p[0] - pc[0], dy = B.p[1] - pc[1], dz = B.p[2] - pc[2];
  const float reach = c.chassis_radius + c.blk_radius + c.pp[2][7];
  if (dx * dx + dy * dy + dz * dz > reach * reach) return 0;
  const float A[3][3] = {{P.ex[0], P.ex[1], P.ex[2]}, {P.ey[0], P.ey[1], P.ey[2]}, {P.ez[0], P.ez[1], P.ez[2]}};
  const float Bx[3][3] = {{B.ex[0], B.ex[1], B.ex[2]}, {B.ey[0], B.ey[1], B.ey[2]}, {B.ez[0], B.ez[1], B.ez[2]}};
  const float h1[3] = {c.chassis_half[0], c.chassis_half[1], c.chassis_half[2]}, dp[3] = {dx, dy, dz};
  int kind, bi, bj;
  float best;
  float nn[3];
  if (!env03_sat(A, h1, Bx, c.blk_half, dp, c.pp[2][7], kind, bi, bj, best, nn)) return 0;
  bn[0] = nn[0]; bn[1] = nn[1]; bn[2] = nn[2];
  const float h2[3] = {c.blk_half, c.blk_half, c.blk_half};
  return env03_manifold(pc, A, h1, B.p, Bx, h2, c.pp[2][7], kind, bi, bj, best, bn, bpos, bdist);
}

// nsub substeps as a per-lane state machine with ONE solve site: the robot (8 dofs), the block (6 dofs) and their coupling
// always go through coupled_solve_fast (with no chassis-block contact the coupling block is zero and the elimination
// decouples exactly).  A single code path keeps the loop body inside the instruction cache: the first version had
// separate robot / block / coupled solvers and was fetch-bound (ncu: stall_no_instruction 4.9 per issue, profiles/).
template <int MAXIT, bool WBON>
BRB_D void phys03_run(const BrbModelConsts &c, Phys &P, Blk &B, int nsub, KF (&qstale)[4], float (&pstale)[3], Env03Stats &es, const unsigned wmask, const bool ctasync) {
  int sidx = 0, it = 0, nbb = 0, qprev_nc = -1;
  bool need_setup = true, done = false;   // warp-uniform loop + explicit reconvergence: see phys_run
  float bpos[8][3], bdist[8], bn[3];
  CBSet Q;
  Q.nc = 0; Q.bits = 0xFFFFFFFFu;
  WBSet W;
  W.nw = 0; W.bits = 0xFFu;
  int wprev_nw = -1;
  unsigned was = 0u;
  int wasn = -1;
  for (;;) {
    if (!(ctasync ? BRB_CTA_OR(!done) : __any_sync(wmask, !done))) break;
    if (!done && need_setup) {
      phys_setup<true>(c, P);
      const unsigned fresh = P.valid & ~was;
      P.bits |= ((fresh & 1u) ? 0xFu : 0u) | ((fresh & 2u) ? 0xF0u : 0u) | ((fresh & 4u) ? 0xF00u : 0u) | ((fresh & 8u) ? 0xF000u : 0u);
      blk_setup(c, B);
      if (B.nc != wasn) B.bits = 0xFFFFu;
    }
    if (ctasync) BRB_CTA_SYNC();
    if (!done && need_setup) {
      nbb = env03_detect(c, P, B, bpos, bdist, bn);
      Q.nc = 0;
      if (nbb > 0) {
        cb_setup(c, P, B, bpos, bdist, bn, nbb, Q);
        if (Q.nc != qprev_nc) Q.bits = 0xFFFFFFFFu;
        if (Q.nc > 0) es.coupled++;
      }
      if (WBON) {
        wb_setup(c, P, B, W);
        if (W.nw != wprev_nw) W.bits = 0xFFu;
        if (W.nw > 0 && Q.nc == 0) es.coupled++;
        wprev_nw = W.nw;
        es.wb_last = W.near ? 1u : 0u;
      }
      es.coupled_last = (unsigned)(Q.nc + (WBON ? W.nw : 0));
      es.blk_last = B.nc > 0 ? 1u : 0u;
      qprev_nc = Q.nc;
      was = P.valid; wasn = B.nc;
      need_setup = false;
      it = 0;
    }
    float ar[8], ab[6];
    bool conv = true;
#ifdef BRB_TRIPSTATS
    {
      const unsigned act = __ballot_sync(wmask, !done);
      const unsigned nv = __popc(__ballot_sync(wmask, !done && Q.nc > 0));
      if ((threadIdx.x & 31u) == (unsigned)(__ffs(wmask) - 1)) {
        atomicAdd(&g_trip[0], 1ull); atomicAdd(&g_trip[1], (unsigned long long)__popc(act));
        if (nv) { atomicAdd(&g_trip[2], 1ull); atomicAdd(&g_trip[3], (unsigned long long)nv); }
      }
    }
#endif
    if (ctasync) BRB_CTA_SYNC(); else __syncwarp(wmask);
    if (done) {
    } else if (P.valid || B.nc > 0 || Q.nc > 0 || (WBON && W.nw > 0)) {
      if (WBON && W.nw > 0) coupled_solve_wb(c, P, B, Q, W, ar, ab);
      else coupled_solve_fast(c, P, B, Q, W, ar, ab);
      es.csolves++;
      // rows within eps of their switching surface keep their state; when the undamped iteration starts to cycle the band
      // is widened (x4 per extra solve from the third on), which freezes the flapping rows at a force error <= D eps
      const float eps = it < 2 ? 2e-4f : (it == 2 ? 8e-4f : (it == 3 ? 3.2e-3f : (it == 4 ? 1.28e-2f : 5.12e-2f)));
      const unsigned nr = P.valid ? phys_active_set<true>(c, P, ar, P.bits, eps) : P.bits;
      const unsigned nbl = blk_active_set(c, B, ab, B.bits, eps), nq = cb_active_set(c, Q, ar, ab, Q.bits, eps);
      const unsigned nwb = (WBON && W.nw > 0) ? wb_active_set(c, W, ar, ab, W.bits, eps) : W.bits;
      conv = (nr == P.bits) && (nbl == B.bits) && (nq == Q.bits) && (nwb == W.bits);
      P.bits = nr; B.bits = nbl; Q.bits = nq; W.bits = nwb;
      if (!conv && ++it >= 7) {
        // the undamped active-set iteration is cycling: finish with the line-search Newton, and seed the next substep with
        // the active sets of ITS solution (otherwise the same cycle — and the same fallback — repeats substep after substep)
        float acc[14];
        es.fallback++;
        if (env03_coupled_substep(c, P, B, bpos, bdist, bn, Q.nc > 0 ? nbb : 0, W, acc)) P.n_nonconv++;
#pragma unroll
        for (int k = 0; k < 8; k++) ar[k] = acc[k];
#pragma unroll
        for (int k = 0; k < 6; k++) ab[k] = acc[8 + k];
        if (P.valid) P.bits = phys_active_set<true>(c, P, ar, P.bits);
        B.bits = blk_active_set(c, B, ab, B.bits);
        Q.bits = cb_active_set(c, Q, ar, ab, Q.bits);
        if (WBON && W.nw > 0) W.bits = wb_active_set(c, W, ar, ab, W.bits);
        conv = true;
      }
    } else {
      // both bodies in free flight: a_b(robot) = M_b^-1 f_b in the chassis frame, block = gravity
      const float mg = c.mass * c.grav;
      const float f[8] = {P.fb[0] - mg * P.ex[2], P.fb[1] - mg * P.ey[2], P.fb[2] - mg * P.ez[2], P.fb[3], P.fb[4], P.fb[5], P.fb[6], P.fb[7]};
      const float u0 = c.minv_xy[0] * f[0] + c.minv_xy[1] * f[4];
      const float u1 = c.minv_blk[0] * f[1] + c.minv_blk[1] * f[3] + c.minv_blk[2] * f[6] + c.minv_blk[3] * f[7];
      const float u2 = c.minv_uz * f[2];
      const float b0 = c.minv_blk[1] * f[1] + c.minv_blk[4] * f[3] + c.minv_blk[5] * f[6] + c.minv_blk[6] * f[7];
      const float b1 = c.minv_xy[1] * f[0] + c.minv_xy[2] * f[4], b2 = c.minv_wz * f[5];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        ar[k] = P.ex[k] * u0 + P.ey[k] * u1 + P.ez[k] * u2;
        ar[3 + k] = P.ex[k] * b0 + P.ey[k] * b1 + P.ez[k] * b2;          // world-frame angular acceleration
      }
      ar[6] = c.minv_blk[2] * f[1] + c.minv_blk[5] * f[3] + c.minv_blk[7] * f[6] + c.minv_blk[8] * f[7];
      ar[7] = c.minv_blk[3] * f[1] + c.minv_blk[6] * f[3] + c.minv_blk[8] * f[6] + c.minv_blk[9] * f[7];
      ab[0] = 0.f; ab[1] = 0.f; ab[2] = -c.grav; ab[3] = 0.f; ab[4] = 0.f; ab[5] = 0.f;
    }
    if (ctasync) BRB_CTA_SYNC(); else __syncwarp(wmask);
    if (!done && conv) {
      if (sidx == nsub - 1) {
#pragma unroll
        for (int k = 0; k < 4; k++) qstale[k] = P.q[k];
#pragma unroll
        for (int k = 0; k < 3; k++) pstale[k] = P.p[k].s - P.p[k].c;       // Q1: xpos is one substep stale too
      }
      if (B.nc > 0) es.blk_contact++;
      phys_finalize(c, P, ar[0], ar[1], ar[2], P.ex[0] * ar[3] + P.ex[1] * ar[4] + P.ex[2] * ar[5],
                    P.ey[0] * ar[3] + P.ey[1] * ar[4] + P.ey[2] * ar[5], P.ez[0] * ar[3] + P.ez[1] * ar[4] + P.ez[2] * ar[5], ar[6], ar[7]);
      blk_finalize(c, B, ab);
      if (++sidx >= nsub) done = true;
      need_setup = true;
    }
  }
}

BRB_D double yaw_of(const double q[4]) {   // RobotBaseEnv.py:177-184
  if (q[0] == 0.0) return 0.0;
  return euler_xyz_of<2>(q);
}

// Env03_v2.set_block_pos_vel (env03_v2.py:25-59).  u[0..4] = target x, target z, block x/y/z_rot.
BRB_D void env03_fire_block(const double robot_pos[3], const double xquat[4], bool attack_front, const double *u, double *bq /*[7]*/,
                            double *bv /*[3]*/) {
  double ang = -yaw_of(xquat);
  if (!attack_front) ang += BRB_PI;
  double sn, cs;
  sincos(ang, &sn, &cs);
  const double bp[3] = {0.3 * sn + robot_pos[0], 0.3 * cs + robot_pos[1], 0.15};
  const double tg[3] = {(u[0] - 0.5) * 0.02 + robot_pos[0], 0 + robot_pos[1], u[1] * 0.025 + 0.13};
  double v[3] = {tg[0] - bp[0], tg[1] - bp[1], tg[2] - bp[2]};
  const double nrm = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  for (int k = 0; k < 3; k++) { bv[k] = 7.5 * (v[k] / nrm); bq[k] = bp[k]; }
  const double x_rot = u[2] * 2 * BRB_PI, y_rot = u[3] * 2 * BRB_PI, z_rot = u[4] * 2 * BRB_PI;
  double sa, ca, sb, cb, sc, cc;
  sincos(x_rot / 2, &sa, &ca); sincos(y_rot / 2, &sb, &cb); sincos(z_rot / 2, &sc, &cc);
  bq[3] = sa * cb * cc - ca * sb * sc;      // scalar-last quaternion into scalar-first slots, as for the robot (Q3)
  bq[4] = ca * sb * cc + sa * cb * sc;
  bq[5] = ca * cb * sc - sa * sb * cc;
  bq[6] = ca * cb * cc + sa * sb * sc;
}

BRB_D bool env03_attack_front(const BrbState &S, long long i) {   // env03_v2.py:22: np.random.random() > 0.5, once per env
  double u[4];
  draw4(S.seed, (uint64_t)(S.env0 + i), 0xFFFFFFFFu, 0u, u);
  return u[0] > 0.5;
}

// Env03.reset_model (env03_v1.py:60-83) with Env03_v2.set_block_pos_vel; u[32]: 16 jitter, 3 robot rotation, 5 block draws
BRB_D void reset_env03(const BrbState &S, long long i, const double *u, float o[6]) {
  const long long N = S.n;
  double qpos[16];
  const double qpos0[16] = {0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0};
  for (int k = 0; k < 16; k++) qpos[k] = qpos0[k] + (-0.01 + (0.01 - -0.01) * u[k]);
  qpos[2] = 0;
  const double x_rot = (u[16] - 0.5) * 2 * BRB_PI, y_rot = (u[17] - 0.5) * 0.4, z_rot = (u[18] - 0.5) * 0.4;
  double sa, ca, sb, cb, sc, cc;
  sincos(x_rot / 2, &sa, &ca); sincos(y_rot / 2, &sb, &cb); sincos(z_rot / 2, &sc, &cc);
  qpos[3] = sa * cb * cc - ca * sb * sc;
  qpos[4] = ca * sb * cc + sa * cb * sc;
  qpos[5] = ca * cb * sc - sa * sb * cc;
  qpos[6] = ca * cb * cc + sa * sb * sc;
  double xq[4];
  const double n = sqrt(qpos[3] * qpos[3] + qpos[4] * qpos[4] + qpos[5] * qpos[5] + qpos[6] * qpos[6]);
  for (int k = 0; k < 4; k++) { xq[k] = qpos[3 + k] / n; S.xquat[k * N + i] = xq[k]; }
  double bv[3];
  env03_fire_block(qpos, xq, env03_attack_front(S, i), u + 19, qpos + 9, bv);     // set_state ran mj_forward: xpos/xquat fresh
  for (int k = 0; k < 16; k++) S.qpos[k * N + i] = qpos[k];
  for (int k = 0; k < 14; k++) S.qvel[k * N + i] = 0.0;
  for (int k = 0; k < 3; k++) S.qvel[(8 + k) * N + i] = bv[k];
  S.aset[i] = 0xFFFFu;
  S.v3[0 * N + i] = -1.0;          // block_delay_time_start = None
  S.v3[2 * N + i] = 65535.0;       // block active set
  S.elapsed[i] = 0;
  S.ep_return[i] = 0.0;
  S.ep_len[i] = 0;
  const double p = pitch_of(xq);
  S.last_pitch[i] = p;
  obs_of(p, 0.0, 0.0, 0.0, 0.0, 0.0, o);
}

// One VecEnv.step of Env03-v2 for env i (Env03.step, env03_v1.py:26-58).  replay_u row = 8 re-fire slots + 32 reset slots.
template <bool WBON>
BRB_D void step_env03(const BrbModelConsts &c, const BrbState &S, const long long i, const float *__restrict__ actions,
                      float *__restrict__ obs, float *__restrict__ reward, uint8_t *__restrict__ done,
                      uint8_t *__restrict__ truncated, float *__restrict__ terminal_obs, float *__restrict__ ep_return_out,
                      int32_t *__restrict__ ep_len_out, const double *__restrict__ replay_u, unsigned stat[12], const unsigned wmask, const bool ctasync) {
  const long long N = S.n;
  double qvel[14], xq[4];
  for (int k = 0; k < 14; k++) qvel[k] = S.qvel[k * N + i];
  for (int k = 0; k < 4; k++) xq[k] = S.xquat[k * N + i];
  const uint32_t event = S.event[i] + 1u;
  S.event[i] = event;
  int elapsed = S.elapsed[i];
  const double rew = reward_of<BRB_ENV01_V1>(pitch_of(xq), qvel[6], qvel[7], qvel[5], 0.0, 0.0);   // RobotBaseEnv._get_reward
  const double ctrl[2] = {qvel[6] + (double)actions[2 * i] * 4.0, qvel[7] + (double)actions[2 * i + 1] * 4.0};

  Phys st;
  Blk B;
  {
    double qn[4], nn = 0;
    for (int k = 0; k < 4; k++) { qn[k] = S.qpos[(3 + k) * N + i]; nn += qn[k] * qn[k]; }
    nn = 1.0 / sqrt(nn);
    for (int k = 0; k < 4; k++) st.q[k] = ksplit(qn[k] * nn);
    for (int k = 0; k < 3; k++) { st.p[k] = ksplit(S.qpos[k * N + i]); st.v[k] = ksplit(qvel[k]); st.w[k] = ksplit(qvel[3 + k]); }
    for (int k = 0; k < 2; k++) {
      st.th[k] = ksplit(S.qpos[(7 + k) * N + i]);
      st.s[k] = ksplit(qvel[6 + k]);
      const double u = fmin((double)c.ctrl_hi, fmax((double)c.ctrl_lo, ctrl[k]));
      st.uhi[k] = (float)u;
      st.ulo[k] = (float)(u - (double)st.uhi[k]);
    }
    st.bits = S.aset[i];
    st.n_contact = st.n_solve = st.n_nonconv = st.n_slots = 0;
    for (int ci = 0; ci < 4; ci++) {     // wheel-rim contact records are read branch-free: finite values from the start
      st.cD[ci] = 0.f;
      for (int k = 0; k < 3; k++) st.cr[ci][k] = st.cw[ci][k] = st.cy[ci][k] = 0.f;
    }
    nn = 0;
    for (int k = 0; k < 4; k++) { qn[k] = S.qpos[(12 + k) * N + i]; nn += qn[k] * qn[k]; }
    nn = 1.0 / sqrt(nn);
    for (int k = 0; k < 4; k++) B.q[k] = (float)(qn[k] * nn);
    for (int k = 0; k < 3; k++) { B.p[k] = (float)S.qpos[(9 + k) * N + i]; B.v[k] = (float)qvel[8 + k]; B.w[k] = (float)qvel[11 + k]; }
    B.bits = (unsigned)S.v3[2 * N + i];
    B.nc = 0;
  }
  KF qprev[4];
  float pstale[3];
  Env03Stats es = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
  phys03_run<BRB_MAXIT, WBON>(c, st, B, c.frame_skip, qprev, pstale, es, wmask, ctasync);
  stat[0] = c.frame_skip; stat[1] = st.n_contact; stat[2] = st.n_solve; stat[3] = st.n_nonconv; stat[7] = st.n_slots;
  stat[8] = es.coupled; stat[9] = es.blk_contact; stat[10] = es.fallback; stat[11] = es.csolves;
  {
    stat[6] = (st.valid == 0u) ? airborne_key(c, st.q[0].s, st.q[1].s, st.q[2].s, st.q[3].s, st.p[2].s, st.v[2].s, st.w[0].s, st.w[1].s, st.w[2].s)
                               : BRB_NLAND + group_rank(st.valid);
    // robots whose block was touching the chassis in the last substep (impacts last ~6 env steps) get their own bucket, so
    // the warps running the coupled assembly are not diluted by robots on the uncoupled path
    // (one bucket per contact count class: the coupled assembly loops over the contacts, lanes with fewer of them idle)
    if (es.coupled_last) stat[6] = BRB_NGROUPS - 4 + (es.coupled_last <= 2u ? 0u : (es.coupled_last <= 4u ? 1u : 2u));
    else if (es.blk_last) stat[6] = BRB_NGROUPS - 5;      // block resting / sliding on the floor: its floor contacts are a rare path
    // a wheel within reach of the block runs the cylinder-box collider (and, in contact, the not-inlined solve with wheel rows) every
    // substep: in lockstep one such robot holds up its CTA, so they get CTAs of their own, visited first
    if (es.wb_last) stat[6] = BRB_NGROUPS - 1;
  }

  double qpos[16];
  for (int k = 0; k < 3; k++) { qpos[k] = kjoin(st.p[k]); qvel[k] = kjoin(st.v[k]); qvel[3 + k] = kjoin(st.w[k]); }
  for (int k = 0; k < 4; k++) qpos[3 + k] = kjoin(st.q[k]);
  for (int k = 0; k < 2; k++) { qpos[7 + k] = kjoin(st.th[k]); qvel[6 + k] = kjoin(st.s[k]); }
  for (int k = 0; k < 3; k++) { qpos[9 + k] = (double)B.p[k]; qvel[8 + k] = (double)B.v[k]; qvel[11 + k] = (double)B.w[k]; }
  for (int k = 0; k < 4; k++) qpos[12 + k] = (double)B.q[k];
  {
    double nn = 0;
    for (int k = 0; k < 4; k++) { xq[k] = kjoin(qprev[k]); nn += xq[k] * xq[k]; }
    nn = 1.0 / sqrt(nn);
    for (int k = 0; k < 4; k++) xq[k] *= nn;
  }
  elapsed += 1;
  const double tnow = S.time_table[elapsed];
  // block logic (env03_v1.py:39-49): remove a block that came to rest, re-fire it 0.5 s later
  double timer = S.v3[0 * N + i];
  {
    const double bs = sqrt(qvel[8] * qvel[8] + qvel[9] * qvel[9] + qvel[10] * qvel[10]);
    if (bs < 0.1 && timer < 0.0) { qpos[9] = 10; qpos[10] = 10; qpos[11] = 0; timer = tnow; }
    if (timer >= 0.0 && (tnow - timer) > 0.5) {
      double uf[8];
      if (replay_u) { for (int k = 0; k < 8; k++) uf[k] = replay_u[i * 40 + k]; }
      else { draw4(S.seed, (uint64_t)(S.env0 + i), event, 9u, uf); draw4(S.seed, (uint64_t)(S.env0 + i), event, 10u, uf + 4); }
      const double rp[3] = {(double)pstale[0], (double)pstale[1], (double)pstale[2]};
      double bv[3];
      env03_fire_block(rp, xq, env03_attack_front(S, i), uf, qpos + 9, bv);
      for (int k = 0; k < 3; k++) qvel[8 + k] = bv[k];
      timer = -1.0;
    }
  }
  const double p_true = pitch_of(xq);
  const bool terminated = fabs(p_true) > (50 * BRB_PI / 180);
  const bool trunc = elapsed >= c.max_episode_steps;
  const double dt = tnow - S.time_table[elapsed - 1];
  double pitch_dot = 0.0;
  if (dt > 0.0) pitch_dot = (p_true - S.last_pitch[i]) / dt;
  float o[6];
  obs_of(p_true, pitch_dot, qvel[6], qvel[7], 0.0, 0.0, o);
  const double epr = S.ep_return[i] + rew;
  const int epl = S.ep_len[i] + 1;
  // poses whose contacts this kernel does not model, evaluated on the post-step state like Env01's count: chassis on the floor,
  // a wheel lying flat, and a wheel within reach of the block (conservative separating-axis test on the block's face normals, the
  // wheel axis and the centre line: "not separated on these five axes" over-counts, it never misses a touching pair)
  {
    const double w = qpos[3], x = qpos[4], y = qpos[5], z = qpos[6];
    const double R[3][3] = {{1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)}, {2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)},
                            {2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)}};
    const double n0 = R[2][0], n1 = R[2][1], n2 = R[2][2];
    const double hz = qpos[2] - ((double)c.zfloor + (double)c.zfloor_lo);
    const double low = hz + n2 * (double)c.chassis_pos[2] + n0 * (double)c.chassis_pos[0] + n1 * (double)c.chassis_pos[1]
                       - fabs(n0) * (double)c.chassis_half[0] - fabs(n1) * (double)c.chassis_half[1] - fabs(n2) * (double)c.chassis_half[2];
    const double rho = sqrt(n1 * n1 + n2 * n2);
    const double tri = hz + (double)c.oz * n2 - fabs(n0) * ((double)c.ox + (double)c.hl) + 0.5 * (double)c.rad * rho;
    bool near_wheel = false;
    const double bw = qpos[12], bx = qpos[13], by = qpos[14], bz = qpos[15];
    const double E[3][3] = {{1 - 2 * (by * by + bz * bz), 2 * (bx * by + bw * bz), 2 * (bx * bz - bw * by)},      // block axes as rows
                            {2 * (bx * by - bw * bz), 1 - 2 * (bx * bx + bz * bz), 2 * (by * bz + bw * bx)},
                            {2 * (bx * bz + bw * by), 2 * (by * bz - bw * bx), 1 - 2 * (bx * bx + by * by)}};
    const double ax[3] = {R[0][0], R[1][0], R[2][0]};                                                       // wheel axes = chassis x
    for (int k = 0; k < 2; k++) {
      const double sg = k ? 1.0 : -1.0, ox = sg * (double)c.ox, oz = (double)c.oz;
      const double d[3] = {qpos[9] - (qpos[0] + R[0][0] * ox + R[0][2] * oz), qpos[10] - (qpos[1] + R[1][0] * ox + R[1][2] * oz),
                           qpos[11] - (qpos[2] + R[2][0] * ox + R[2][2] * oz)};
      const double dl = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      bool sep = false;
      for (int a = 0; a < 5 && !sep; a++) {
        double L[3];
        if (a < 3) { L[0] = E[a][0]; L[1] = E[a][1]; L[2] = E[a][2]; }
        else if (a == 3) { L[0] = ax[0]; L[1] = ax[1]; L[2] = ax[2]; }
        else { if (dl < 1e-9) break; L[0] = d[0] / dl; L[1] = d[1] / dl; L[2] = d[2] / dl; }
        const double la = L[0] * ax[0] + L[1] * ax[1] + L[2] * ax[2];
        const double sc = (double)c.hl * fabs(la) + (double)c.rad * sqrt(fmax(0.0, 1.0 - la * la));
        double sb = 0.0;
        for (int e = 0; e < 3; e++) sb += (double)c.blk_half * fabs(L[0] * E[e][0] + L[1] * E[e][1] + L[2] * E[e][2]);
        sep = fabs(L[0] * d[0] + L[1] * d[1] + L[2] * d[2]) > sc + sb + (double)c.pp[2][7];
      }
      near_wheel = near_wheel || !sep;
    }
    if (WBON) near_wheel = false;     // the pair is generated (wb_setup): nothing unsupported about it
#ifdef BRB_PROBE_UNSUP   // kernel-tuning experiment: count one cause at a time (1 = chassis-floor, 2 = wheel flat, 4 = wheel near block)
    stat[4] = (((BRB_PROBE_UNSUP & 1) && low <= 0.0) || ((BRB_PROBE_UNSUP & 2) && tri <= 0.0) || ((BRB_PROBE_UNSUP & 4) && near_wheel)) ? 1u : 0u;
#else
    stat[4] = (low <= 0.0 || tri <= 0.0 || near_wheel) ? 1u : 0u;
#endif
    if (low <= 0.0) es.unsupported |= 1u;
    if (near_wheel) es.unsupported |= 2u;
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
    stat[6] = 0;
    if (terminal_obs) for (int k = 0; k < 6; k++) terminal_obs[i * 6 + k] = o[k];
    double ur[32];
    if (replay_u) { for (int k = 0; k < 32; k++) ur[k] = replay_u[i * 40 + 8 + k]; }
    else { for (int b = 0; b < 8; b++) draw4(S.seed, (uint64_t)(S.env0 + i), event, 1u + b, ur + 4 * b); }
    reset_env03(S, i, ur, o);
  } else {
    for (int k = 0; k < 16; k++) S.qpos[k * N + i] = qpos[k];
    for (int k = 0; k < 14; k++) S.qvel[k * N + i] = qvel[k];
    for (int k = 0; k < 4; k++) S.xquat[k * N + i] = xq[k];
    S.aset[i] = st.bits;
    S.v3[0 * N + i] = timer;
    S.v3[2 * N + i] = (double)B.bits;
    S.elapsed[i] = elapsed;
    S.last_pitch[i] = p_true;
    S.ep_return[i] = epr;
    S.ep_len[i] = epl;
  }
  for (int k = 0; k < 6; k++) obs[i * 6 + k] = o[k];
}
#endif
