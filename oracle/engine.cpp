// ORACLE - TEST INFRASTRUCTURE ONLY (see engine.h).  Dense, generic, FP64 restatement of MuJoCo's
// mj_forward / mj_step pipeline for the reference's scene.  Deliberately written in the "textbook"
// dense style (explicit body Jacobians, explicit efc_J, dense Cholesky) so that it shares no
// structure with the matrix-free, structure-specialised CUDA kernels it is used to check.
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "ccd.h"

namespace orc {

// ------------------------------------------------------------------------------------------------
// small math
// ------------------------------------------------------------------------------------------------
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline double norm3(const double* a) { return std::sqrt(dot3(a, a)); }
static inline void mulquat(double* r, const double* a, const double* b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static inline void normquat(double* q) {
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  for (int i = 0; i < 4; i++) q[i] /= n;
}
static inline void quat2mat(double* m, const double* q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
static inline void mulmatvec3(double* r, const double* m, const double* v) {
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  double y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  double z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void mulmatTvec3(double* r, const double* m, const double* v) {
  double x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  double y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  double z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void mulmat3(double* r, const double* a, const double* b) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  std::memcpy(r, t, sizeof t);
}
static inline void axisangle2quat(double* q, const double* axis, double angle) {
  double s = std::sin(angle * 0.5);
  q[0] = std::cos(angle * 0.5); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}

// ------------------------------------------------------------------------------------------------
// position stage: kinematics
// ------------------------------------------------------------------------------------------------
static const double kCamPos[3][3] = {{0, 0, 2.0}, {1.2, -0.3, 0.8}, {-0.07, 0.0, 0.055}};
static const int kCamBody[3] = {0, 0, 9};

void kinematics_only(Data& d) {
  for (int k = 0; k < 3; k++) d.xpos[0][k] = 0;
  d.xquat[0][0] = 1; d.xquat[0][1] = d.xquat[0][2] = d.xquat[0][3] = 0;
  quat2mat(d.xmat[0], d.xquat[0]);
  for (int k = 0; k < 3; k++) d.xipos[0][k] = 0;
  for (int b = 1; b < NBODY; b++) {
    int p = mm_body_parent[b], j = mm_body_jnt[b];
    double* xp = d.xpos[b];
    double* xq = d.xquat[b];
    if (j >= 0 && mm_jnt_type[j] == 0) {  // free
      int qa = mm_jnt_qposadr[j];
      normquat(d.qpos + qa + 3);
      for (int k = 0; k < 3; k++) xp[k] = d.qpos[qa + k];
      for (int k = 0; k < 4; k++) xq[k] = d.qpos[qa + 3 + k];
      for (int k = 0; k < 3; k++) { d.xanchor[j][k] = xp[k]; d.xaxis[j][k] = (k == 2); }
    } else {
      double t[3];
      mulmatvec3(t, d.xmat[p], mm_body_pos[b]);
      for (int k = 0; k < 3; k++) xp[k] = d.xpos[p][k] + t[k];
      mulquat(xq, d.xquat[p], mm_body_quat[b]);
      if (j >= 0) {
        double R[9], ax[3];
        quat2mat(R, xq);
        mulmatvec3(ax, R, mm_jnt_axis[j]);  // joint axis in world (before the joint's own motion)
        double q = d.qpos[mm_jnt_qposadr[j]];  // qpos0 = 0, jnt_pos = 0
        for (int k = 0; k < 3; k++) { d.xaxis[j][k] = ax[k]; }
        if (mm_jnt_type[j] == 3) {  // hinge about the body origin
          double qr[4], qn[4];
          axisangle2quat(qr, mm_jnt_axis[j], q);
          mulquat(qn, xq, qr);
          for (int k = 0; k < 4; k++) xq[k] = qn[k];
          for (int k = 0; k < 3; k++) d.xanchor[j][k] = xp[k];
        } else {  // slide
          for (int k = 0; k < 3; k++) d.xanchor[j][k] = xp[k];
          for (int k = 0; k < 3; k++) xp[k] += ax[k] * q;
        }
      }
      normquat(xq);
    }
    quat2mat(d.xmat[b], xq);
    double c[3];
    mulmatvec3(c, d.xmat[b], mm_body_ipos[b]);
    for (int k = 0; k < 3; k++) d.xipos[b][k] = xp[k] + c[k];
  }
  for (int g = 0; g < NGEOM; g++) {
    int b = mm_geom_body[g];
    double t[3], R[9];
    mulmatvec3(t, d.xmat[b], mm_geom_pos[g]);
    for (int k = 0; k < 3; k++) d.geom_xpos[g][k] = d.xpos[b][k] + t[k];
    quat2mat(R, mm_geom_quat[g]);
    mulmat3(d.geom_xmat[g], d.xmat[b], R);
  }
  // cameras: overhead (xyaxes 1 0 0 0 1 0 -> identity), side (xyaxes 0 1 0 -0.4 0 0.9), wrist (env.py:57-64)
  {
    static const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    std::memcpy(d.cam_xmat[0], I3, sizeof I3);
    for (int k = 0; k < 3; k++) d.cam_xpos[0][k] = kCamPos[0][k];
    double x[3] = {0, 1, 0}, y[3] = {-0.4, 0, 0.9}, z[3];
    double pr = dot3(x, y);
    for (int k = 0; k < 3; k++) y[k] -= pr * x[k];
    double ny = norm3(y);
    for (int k = 0; k < 3; k++) y[k] /= ny;
    cross3(z, x, y);
    for (int r = 0; r < 3; r++) { d.cam_xmat[1][3 * r] = x[r]; d.cam_xmat[1][3 * r + 1] = y[r]; d.cam_xmat[1][3 * r + 2] = z[r]; }
    for (int k = 0; k < 3; k++) d.cam_xpos[1][k] = kCamPos[1][k];
    double wq[4] = {-0.0616, -0.7044, 0.7044, 0.0616}, wr[9], t[3];
    normquat(wq);
    quat2mat(wr, wq);
    int hb = kCamBody[2];
    mulmat3(d.cam_xmat[2], d.xmat[hb], wr);
    mulmatvec3(t, d.xmat[hb], kCamPos[2]);
    for (int k = 0; k < 3; k++) d.cam_xpos[2][k] = d.xpos[hb][k] + t[k];
  }
}

// Jacobian of a world point attached to `body`.  jacp/jacr are 3 x NV row-major.
void jac(const Data& d, double* jacp, double* jacr, const double point[3], int body) {
  if (jacp) std::memset(jacp, 0, sizeof(double) * 3 * NV);
  if (jacr) std::memset(jacr, 0, sizeof(double) * 3 * NV);
  int b = body;
  while (b > 0) {
    int j = mm_body_jnt[b];
    if (j >= 0) {
      int da = mm_jnt_dofadr[j];
      if (mm_jnt_type[j] == 3) {
        double r[3], c[3];
        for (int k = 0; k < 3; k++) r[k] = point[k] - d.xanchor[j][k];
        cross3(c, d.xaxis[j], r);
        for (int k = 0; k < 3; k++) {
          if (jacr) jacr[k * NV + da] = d.xaxis[j][k];
          if (jacp) jacp[k * NV + da] = c[k];
        }
      } else if (mm_jnt_type[j] == 2) {
        for (int k = 0; k < 3; k++)
          if (jacp) jacp[k * NV + da] = d.xaxis[j][k];
      } else {  // free: 3 world-frame translations, 3 body-frame rotations
        for (int k = 0; k < 3; k++)
          if (jacp) jacp[k * NV + da + k] = 1.0;
        double r[3];
        for (int k = 0; k < 3; k++) r[k] = point[k] - d.xpos[b][k];
        for (int a = 0; a < 3; a++) {
          double ax[3] = {d.xmat[b][a], d.xmat[b][3 + a], d.xmat[b][6 + a]}, c[3];
          cross3(c, ax, r);
          for (int k = 0; k < 3; k++) {
            if (jacr) jacr[k * NV + da + 3 + a] = ax[k];
            if (jacp) jacp[k * NV + da + 3 + a] = c[k];
          }
        }
      }
    }
    b = mm_body_parent[b];
  }
}

// ------------------------------------------------------------------------------------------------
// inertia matrix (dense, from body Jacobians) and bias forces (Newton-Euler projected by J^T)
// ------------------------------------------------------------------------------------------------
static void world_inertia(const Data& d, int b, double* Iw) {
  double t[9], Rt[9];
  const double* R = d.xmat[b];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Rt[3 * i + j] = R[3 * j + i];
  mulmat3(t, R, mm_body_inertia[b]);
  mulmat3(Iw, t, Rt);
}

static void make_M(Data& d) {
  std::memset(d.M, 0, sizeof d.M);
  static thread_local double jp[3 * NV], jr[3 * NV];
  for (int b = 1; b < NBODY; b++) {
    if (mm_body_mass[b] <= 0 || mm_body_weld[b] == 0) continue;
    jac(d, jp, jr, d.xipos[b], b);
    double Iw[9];
    world_inertia(d, b, Iw);
    double m = mm_body_mass[b];
    for (int i = 0; i < NV; i++) {
      double ijr[3];
      for (int r = 0; r < 3; r++) ijr[r] = Iw[3 * r] * jr[i] + Iw[3 * r + 1] * jr[NV + i] + Iw[3 * r + 2] * jr[2 * NV + i];
      for (int k = 0; k < NV; k++) {
        double s = m * (jp[i] * jp[k] + jp[NV + i] * jp[NV + k] + jp[2 * NV + i] * jp[2 * NV + k]);
        s += ijr[0] * jr[k] + ijr[1] * jr[NV + k] + ijr[2] * jr[2 * NV + k];
        d.M[k * NV + i] += s;
      }
    }
  }
  for (int j = 0; j < NJNT; j++)
    if (mm_jnt_type[j] != 0) d.M[mm_jnt_dofadr[j] * (NV + 1)] += mm_jnt_armature[j];
}

static void make_bias(Data& d) {
  // forward pass in world coordinates: angular velocity w, origin velocity v, bias accelerations
  // (zero joint acceleration), gravity folded in as a_world = -g
  double w[NBODY][3], al[NBODY][3], a[NBODY][3];
  for (int k = 0; k < 3; k++) { w[0][k] = 0; al[0][k] = 0; a[0][k] = -mm_gravity[k]; }
  std::memset(d.qfrc_bias, 0, sizeof d.qfrc_bias);
  static thread_local double jp[3 * NV], jr[3 * NV];
  for (int b = 1; b < NBODY; b++) {
    int p = mm_body_parent[b], j = mm_body_jnt[b];
    if (j >= 0 && mm_jnt_type[j] == 0) {
      int da = mm_jnt_dofadr[j];
      mulmatvec3(w[b], d.xmat[b], d.qvel + da + 3);
      for (int k = 0; k < 3; k++) { al[b][k] = 0; a[b][k] = -mm_gravity[k]; }
    } else {
      double r[3], t[3], t2[3];
      for (int k = 0; k < 3; k++) r[k] = d.xpos[b][k] - d.xpos[p][k];
      for (int k = 0; k < 3; k++) { w[b][k] = w[p][k]; al[b][k] = al[p][k]; }
      cross3(t, al[p], r);
      cross3(t2, w[p], r);
      double t3[3];
      cross3(t3, w[p], t2);
      for (int k = 0; k < 3; k++) a[b][k] = a[p][k] + t[k] + t3[k];
      if (j >= 0) {
        double qd = d.qvel[mm_jnt_dofadr[j]];
        double aq[3] = {d.xaxis[j][0] * qd, d.xaxis[j][1] * qd, d.xaxis[j][2] * qd};
        if (mm_jnt_type[j] == 3) {
          cross3(t, w[p], aq);
          for (int k = 0; k < 3; k++) { w[b][k] += aq[k]; al[b][k] += t[k]; }
        } else {
          cross3(t, w[p], aq);
          for (int k = 0; k < 3; k++) a[b][k] += 2 * t[k];
        }
      }
    }
    if (mm_body_mass[b] <= 0 || mm_body_weld[b] == 0) continue;
    double c[3], t[3], t2[3], ac[3], f[3], n[3], Iw[9], Iwv[3];
    for (int k = 0; k < 3; k++) c[k] = d.xipos[b][k] - d.xpos[b][k];
    cross3(t, al[b], c);
    cross3(t2, w[b], c);
    double t3[3];
    cross3(t3, w[b], t2);
    for (int k = 0; k < 3; k++) ac[k] = a[b][k] + t[k] + t3[k];
    for (int k = 0; k < 3; k++) f[k] = mm_body_mass[b] * ac[k];
    world_inertia(d, b, Iw);
    mulmatvec3(n, Iw, al[b]);
    mulmatvec3(Iwv, Iw, w[b]);
    cross3(t, w[b], Iwv);
    for (int k = 0; k < 3; k++) n[k] += t[k];
    jac(d, jp, jr, d.xipos[b], b);
    for (int i = 0; i < NV; i++)
      d.qfrc_bias[i] += jp[i] * f[0] + jp[NV + i] * f[1] + jp[2 * NV + i] * f[2] + jr[i] * n[0] + jr[NV + i] * n[1] +
                        jr[2 * NV + i] * n[2];
  }
}

// dense Cholesky helpers (lower triangular in place, row-major n x n); returns min pivot
static double chol_factor(double* A, int n) {
  double minp = 1e300;
  for (int j = 0; j < n; j++) {
    double s = A[j * n + j];
    for (int k = 0; k < j; k++) s -= A[j * n + k] * A[j * n + k];
    minp = std::min(minp, s);
    if (s < MINVAL) s = MINVAL;
    double l = std::sqrt(s);
    A[j * n + j] = l;
    for (int i = j + 1; i < n; i++) {
      double t = A[i * n + j];
      for (int k = 0; k < j; k++) t -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = t / l;
    }
  }
  return minp;
}
static void chol_solve(const double* L, int n, double* x) {
  for (int i = 0; i < n; i++) {
    double s = x[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * x[k];
    x[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
}

// ------------------------------------------------------------------------------------------------
// collision
// ------------------------------------------------------------------------------------------------
static void make_frame(double* f) {  // mju_makeFrame rule: tangent seed (0,1,0) unless |n_y| >= 0.5
  double n = norm3(f);
  for (int k = 0; k < 3; k++) f[k] /= n;
  double t[3] = {0, 0, 0};
  if (f[1] < 0.5 && f[1] > -0.5) t[1] = 1; else t[2] = 1;
  double pr = dot3(f, t);
  for (int k = 0; k < 3; k++) t[k] -= pr * f[k];
  double nt = norm3(t);
  for (int k = 0; k < 3; k++) f[3 + k] = t[k] / nt;
  cross3(f + 6, f, f + 3);
}

static void add_contact(Data& d, int g1, int g2, const double* pos, const double* normal, double dist) {
  Contact c;
  c.dist = dist;
  for (int k = 0; k < 3; k++) { c.pos[k] = pos[k]; c.frame[k] = normal[k]; }
  make_frame(c.frame);
  c.dim = std::max(mm_geom_condim[g1], mm_geom_condim[g2]);
  double f[3];
  for (int k = 0; k < 3; k++) f[k] = std::max(mm_geom_friction[g1][k], mm_geom_friction[g2][k]);
  c.friction[0] = c.friction[1] = f[0]; c.friction[2] = f[1]; c.friction[3] = c.friction[4] = f[2];
  c.solref[0] = 0.02; c.solref[1] = 1.0;
  c.solimp[0] = 0.9; c.solimp[1] = 0.95; c.solimp[2] = 0.001; c.solimp[3] = 0.5; c.solimp[4] = 2.0;
  c.geom1 = g1; c.geom2 = g2; c.efc_address = -1;
  d.contact.push_back(c);
}

// plane (g1) vs box (g2): each penetrating corner, at most 4, contact midway between corner and plane
static void plane_box(Data& d, int g1, int g2) {
  const double* pp = d.geom_xpos[g1];
  const double* pm = d.geom_xmat[g1];
  double n[3] = {pm[2], pm[5], pm[8]};
  const double* bp = d.geom_xpos[g2];
  const double* bm = d.geom_xmat[g2];
  const double* s = mm_geom_size[g2];
  int cnt = 0;
  for (int i = 0; i < 8 && cnt < 4; i++) {
    double loc[3] = {(i & 1 ? s[0] : -s[0]), (i & 2 ? s[1] : -s[1]), (i & 4 ? s[2] : -s[2])}, c[3];
    mulmatvec3(c, bm, loc);
    for (int k = 0; k < 3; k++) c[k] += bp[k];
    double r[3] = {c[0] - pp[0], c[1] - pp[1], c[2] - pp[2]};
    double dist = dot3(r, n);
    if (dist < 0) {
      double pos[3];
      for (int k = 0; k < 3; k++) pos[k] = c[k] - n[k] * dist * 0.5;
      add_contact(d, g1, g2, pos, n, dist);
      cnt++;
    }
  }
}

// box-box: separating-axis test over 15 axes, then reference-face clipping (<= 8 points, each at the
// mid-surface with its own depth) or a single edge-edge point.  Normal from g1 to g2.
static void box_box(Data& d, int g1, int g2) {
  const double *pa = d.geom_xpos[g1], *Ra = d.geom_xmat[g1], *sa = mm_geom_size[g1];
  const double *pb = d.geom_xpos[g2], *Rb = d.geom_xmat[g2], *sb = mm_geom_size[g2];
  double A[3][3], B[3][3];
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < 3; k++) { A[i][k] = Ra[3 * k + i]; B[i][k] = Rb[3 * k + i]; }
  double dp[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
  double C[3][3], Q[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { C[i][j] = dot3(A[i], B[j]); Q[i][j] = std::fabs(C[i][j]); }
  double best = 1e300, bn[3] = {0, 0, 0};
  int code = -1;
  // face axes of A then B
  for (int i = 0; i < 3; i++) {
    double t = dot3(dp, A[i]);
    double pen = sa[i] + sb[0] * Q[i][0] + sb[1] * Q[i][1] + sb[2] * Q[i][2] - std::fabs(t);
    if (pen < 0) return;
    if (pen < best) { best = pen; code = i; double sg = t < 0 ? -1 : 1; for (int k = 0; k < 3; k++) bn[k] = sg * A[i][k]; }
  }
  for (int j = 0; j < 3; j++) {
    double t = dot3(dp, B[j]);
    double pen = sb[j] + sa[0] * Q[0][j] + sa[1] * Q[1][j] + sa[2] * Q[2][j] - std::fabs(t);
    if (pen < 0) return;
    // a face of the second box must be clearly better: exactly parallel faces (left/right pads, cube
    // on table) tie to rounding error and the choice would otherwise flip with the last bit
    if (pen < best * (1 - 1e-6) - 1e-12) { best = pen; code = 3 + j; double sg = t < 0 ? -1 : 1; for (int k = 0; k < 3; k++) bn[k] = sg * B[j][k]; }
  }
  // edge x edge axes (face axes preferred: an edge axis must be clearly better)
  double ebest = 1e300, en[3] = {0, 0, 0};
  int ecode = -1;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double L[3];
      cross3(L, A[i], B[j]);
      double ln = norm3(L);
      if (ln < 1e-6) continue;
      for (int k = 0; k < 3; k++) L[k] /= ln;
      double t = dot3(dp, L);
      double ra = 0, rb = 0;
      for (int k = 0; k < 3; k++) { ra += sa[k] * std::fabs(dot3(A[k], L)); rb += sb[k] * std::fabs(dot3(B[k], L)); }
      double pen = ra + rb - std::fabs(t);
      if (pen < 0) return;
      if (pen < ebest) { ebest = pen; ecode = 6 + 3 * i + j; double sg = t < 0 ? -1 : 1; for (int k = 0; k < 3; k++) en[k] = sg * L[k]; }
    }
  d.stats.narrow_tests++;
  if (ecode >= 0 && ebest * 1.05 < best) {
    // edge-edge: closest points of the two supporting edges
    int i = (ecode - 6) / 3, j = (ecode - 6) % 3;
    double ea[3], eb[3];
    for (int k = 0; k < 3; k++) { ea[k] = pa[k]; eb[k] = pb[k]; }
    for (int a = 0; a < 3; a++) {
      if (a != i) { double sg = dot3(en, A[a]) > 0 ? 1 : -1; for (int k = 0; k < 3; k++) ea[k] += sg * sa[a] * A[a][k]; }
      if (a != j) { double sg = dot3(en, B[a]) > 0 ? -1 : 1; for (int k = 0; k < 3; k++) eb[k] += sg * sb[a] * B[a][k]; }
    }
    double w[3] = {ea[0] - eb[0], ea[1] - eb[1], ea[2] - eb[2]};
    double uu = 1, vv = 1, uv = C[i][j], uw = dot3(A[i], w), vw = dot3(B[j], w);
    double den = uu * vv - uv * uv;
    double s = (uv * vw - vv * uw) / den, t = (uu * vw - uv * uw) / den;
    s = std::max(-sa[i], std::min(sa[i], s));
    t = std::max(-sb[j], std::min(sb[j], t));
    double pos[3];
    for (int k = 0; k < 3; k++) pos[k] = 0.5 * ((ea[k] + s * A[i][k]) + (eb[k] + t * B[j][k]));
    add_contact(d, g1, g2, pos, en, -ebest);
    return;
  }
  // face case: reference box owns the axis, incident box is the other one
  bool refA = code < 3;
  int ax = refA ? code : code - 3;
  const double* pr = refA ? pa : pb;
  const double* sr = refA ? sa : sb;
  double (*Rr)[3] = refA ? A : B;
  const double* pi = refA ? pb : pa;
  const double* si = refA ? sb : sa;
  double (*Ri)[3] = refA ? B : A;
  double nref[3];  // outward normal of the reference face (points toward the incident box)
  for (int k = 0; k < 3; k++) nref[k] = refA ? bn[k] : -bn[k];
  // incident face: most anti-parallel to nref
  int iax = 0;
  double mx = -1;
  for (int a = 0; a < 3; a++) { double v = std::fabs(dot3(nref, Ri[a])); if (v > mx) { mx = v; iax = a; } }
  double isg = dot3(nref, Ri[iax]) > 0 ? -1 : 1;
  int u = (iax + 1) % 3, v = (iax + 2) % 3;
  double poly[16][3];
  int np = 4;
  static const int su[4] = {1, -1, -1, 1}, sv[4] = {1, 1, -1, -1};
  for (int q = 0; q < 4; q++)
    for (int k = 0; k < 3; k++)
      poly[q][k] = pi[k] + isg * si[iax] * Ri[iax][k] + su[q] * si[u] * Ri[u][k] + sv[q] * si[v] * Ri[v][k];
  // clip against the 4 side planes of the reference face
  int t1 = (ax + 1) % 3, t2 = (ax + 2) % 3;
  for (int side = 0; side < 4 && np > 0; side++) {
    int ta = side < 2 ? t1 : t2;
    double sg = (side & 1) ? -1.0 : 1.0;
    double lim = sr[ta];
    double out[16][3];
    int no = 0;
    for (int q = 0; q < np; q++) {
      const double* P = poly[q];
      const double* Qn = poly[(q + 1) % np];
      double rp[3] = {P[0] - pr[0], P[1] - pr[1], P[2] - pr[2]}, rq[3] = {Qn[0] - pr[0], Qn[1] - pr[1], Qn[2] - pr[2]};
      double dP = sg * dot3(rp, Rr[ta]) - lim, dQ = sg * dot3(rq, Rr[ta]) - lim;
      // 1 nm band: vertices lying on a side plane (exactly aligned pads) count as inside and do not
      // spawn sliver crossings whose existence would depend on the last bit
      const double ce = 1e-9;
      if (dP <= ce) { for (int k = 0; k < 3; k++) out[no][k] = P[k]; no++; }
      if ((dP < -ce && dQ > ce) || (dP > ce && dQ < -ce)) {
        double tt = dP / (dP - dQ);
        for (int k = 0; k < 3; k++) out[no][k] = P[k] + tt * (Qn[k] - P[k]);
        no++;
      }
    }
    np = no;
    for (int q = 0; q < np; q++) for (int k = 0; k < 3; k++) poly[q][k] = out[q][k];
  }
  double sgn = dot3(nref, Rr[ax]) > 0 ? 1.0 : -1.0;
  int cnt = 0;
  for (int q = 0; q < np && cnt < 8; q++) {
    double r[3] = {poly[q][0] - pr[0], poly[q][1] - pr[1], poly[q][2] - pr[2]};
    double depth = sr[ax] - sgn * dot3(r, Rr[ax]);
    if (depth <= 0) continue;
    double pos[3];
    for (int k = 0; k < 3; k++) pos[k] = poly[q][k] + nref[k] * depth * 0.5;
    add_contact(d, g1, g2, pos, bn, -depth);
    cnt++;
  }
}

// general convex pair through GJK/EPA (single contact; MuJoCo multiccd is off by default)
static void convex_convex(Data& d, int g1, int g2) {
  ccd::Shape s1, s2;
  auto fill = [&](ccd::Shape& s, int g) {
    s.type = mm_geom_type[g];
    s.pos = d.geom_xpos[g];
    s.mat = d.geom_xmat[g];
    s.size = mm_geom_size[g];
    s.verts = &mm_hull[mm_geom_vadr[g]][0];
    s.nvert = mm_geom_vnum[g];
  };
  fill(s1, g1);
  fill(s2, g2);
  d.stats.ccd_tests++;
  double pos[3], nrm[3], depth;
  if (ccd::penetration(s1, s2, pos, nrm, &depth)) add_contact(d, g1, g2, pos, nrm, -depth);
}

// plane (g1) vs convex mesh (g2): deepest vertex (single contact)
static void plane_convex(Data& d, int g1, int g2) {
  const double* pp = d.geom_xpos[g1];
  const double* pm = d.geom_xmat[g1];
  double n[3] = {pm[2], pm[5], pm[8]};
  double nl[3];
  mulmatTvec3(nl, d.geom_xmat[g2], n);
  const double* V = &mm_hull[mm_geom_vadr[g2]][0];
  int best = 0;
  double bv = 1e300;
  for (int i = 0; i < mm_geom_vnum[g2]; i++) {
    double v = dot3(V + 3 * i, nl);
    if (v < bv) { bv = v; best = i; }
  }
  double c[3];
  mulmatvec3(c, d.geom_xmat[g2], V + 3 * best);
  for (int k = 0; k < 3; k++) c[k] += d.geom_xpos[g2][k];
  double r[3] = {c[0] - pp[0], c[1] - pp[1], c[2] - pp[2]};
  double dist = dot3(r, n);
  if (dist < 0) {
    double pos[3];
    for (int k = 0; k < 3; k++) pos[k] = c[k] - n[k] * dist * 0.5;
    add_contact(d, g1, g2, pos, n, dist);
  }
}

static void collision(Data& d) {
  d.contact.clear();
  for (int p = 0; p < MM_NPAIR; p++) {
    int g1 = mm_pair[p][0], g2 = mm_pair[p][1];
    int t1 = mm_geom_type[g1], t2 = mm_geom_type[g2];
    // bounding-sphere rejection (planes have no bound)
    if (t1 != 0) {
      double c1[3], c2[3], t[3];
      mulmatvec3(t, d.geom_xmat[g1], mm_geom_bcenter[g1]);
      for (int k = 0; k < 3; k++) c1[k] = d.geom_xpos[g1][k] + t[k];
      mulmatvec3(t, d.geom_xmat[g2], mm_geom_bcenter[g2]);
      for (int k = 0; k < 3; k++) c2[k] = d.geom_xpos[g2][k] + t[k];
      double r[3] = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]};
      double rs = mm_geom_rbound[g1] + mm_geom_rbound[g2];
      if (dot3(r, r) > rs * rs) continue;
    } else {
      double t[3], c2[3];
      mulmatvec3(t, d.geom_xmat[g2], mm_geom_bcenter[g2]);
      for (int k = 0; k < 3; k++) c2[k] = d.geom_xpos[g2][k] + t[k] - d.geom_xpos[g1][k];
      double n[3] = {d.geom_xmat[g1][2], d.geom_xmat[g1][5], d.geom_xmat[g1][8]};
      if (dot3(c2, n) > mm_geom_rbound[g2]) continue;
    }
    if (t1 == 0 && t2 == 6) plane_box(d, g1, g2);
    else if (t1 == 6 && t2 == 6) box_box(d, g1, g2);
    else if (d.flags & 1) continue;
    else if (t1 == 0 && t2 == 7) plane_convex(d, g1, g2);
    else if (t1 == 0) continue;  // plane-cylinder: the table legs are static
    else convex_convex(d, g1, g2);
  }
  d.ncon = (int)d.contact.size();
}

// ------------------------------------------------------------------------------------------------
// constraint rows
// ------------------------------------------------------------------------------------------------
static double impedance(const double* solimp, double pos_minus_margin) {
  double dmin = std::min(0.9999, std::max(1e-4, solimp[0]));
  double dmax = std::min(0.9999, std::max(1e-4, solimp[1]));
  double width = solimp[2];
  double mid = std::min(0.9999, std::max(1e-4, solimp[3]));
  double power = std::max(1.0, solimp[4]);
  if (dmin == dmax || width <= MINVAL) return 0.5 * (dmin + dmax);
  double x = std::fabs(pos_minus_margin) / width;
  if (x >= 1) return dmax;
  if (x <= 0) return dmin;
  double y;
  if (x <= mid) y = std::pow(x, power) / std::pow(mid, power - 1);
  else y = 1 - std::pow(1 - x, power) / std::pow(1 - mid, power - 1);
  return dmin + y * (dmax - dmin);
}

static void kbi(const double* solref, const double* solimp, double pos, double* K, double* B, double* imp) {
  double tc = std::max(solref[0], 2 * MM_TIMESTEP);  // refsafe
  double dr = solref[1];
  double dmax = std::min(0.9999, std::max(1e-4, solimp[1]));
  *K = 1.0 / std::max(MINVAL, dmax * dmax * tc * tc * dr * dr);
  *B = 2.0 / std::max(MINVAL, dmax * tc);
  *imp = impedance(solimp, pos);
}

static void add_row(Data& d, const double* J, double pos, double diagApprox, const double* solref,
                    const double* solimp, int type, double Roverride = -1) {
  double K, B, imp;
  kbi(solref, solimp, pos, &K, &B, &imp);
  double R = std::max(MINVAL, (1 - imp) / imp * diagApprox);
  if (Roverride > 0) R = Roverride;
  double vel = 0;
  for (int i = 0; i < NV; i++) vel += J[i] * d.qvel[i];
  d.efc_J.insert(d.efc_J.end(), J, J + NV);
  d.efc_pos.push_back(pos);
  d.efc_R.push_back(R);
  d.efc_D.push_back(1.0 / R);
  d.efc_vel.push_back(vel);
  d.efc_aref.push_back(-B * vel - K * imp * pos);
  d.efc_type.push_back(type);
}

static void make_constraint(Data& d) {
  d.efc_J.clear(); d.efc_pos.clear(); d.efc_D.clear(); d.efc_R.clear(); d.efc_aref.clear();
  d.efc_vel.clear(); d.efc_type.clear();
  static const double dsolref[2] = {0.02, 1.0}, dsolimp[5] = {0.9, 0.95, 0.001, 0.5, 2.0};
  double J[NV];
  // joint equality (finger coupling)
  std::memset(J, 0, sizeof J);
  J[MM_EQ_DOF1] = 1; J[MM_EQ_DOF2] = -1;
  add_row(d, J, d.qpos[7] - d.qpos[8], mm_dof_invweight0[MM_EQ_DOF1] + mm_dof_invweight0[MM_EQ_DOF2], mm_eq_solref,
          mm_eq_solimp, 0);
  // joint limits
  for (int j = 0; j < NJNT; j++) {
    if (!mm_jnt_limited[j] || mm_jnt_type[j] == 0) continue;
    int da = mm_jnt_dofadr[j];
    double q = d.qpos[mm_jnt_qposadr[j]];
    for (int side = 0; side < 2; side++) {
      double dist = side == 0 ? q - mm_jnt_range[j][0] : mm_jnt_range[j][1] - q;
      if (dist < 0) {
        std::memset(J, 0, sizeof J);
        J[da] = side == 0 ? 1 : -1;
        add_row(d, J, dist, mm_dof_invweight0[da], dsolref, dsolimp, 1);
      }
    }
  }
  // pyramidal contacts
  static thread_local double jp1[3 * NV], jr1[3 * NV], jp2[3 * NV], jr2[3 * NV];
  for (auto& c : d.contact) {
    int b1 = mm_geom_body[c.geom1], b2 = mm_geom_body[c.geom2];
    jac(d, jp1, jr1, c.pos, b1);
    jac(d, jp2, jr2, c.pos, b2);
    double Jc[6][NV];  // contact-frame Jacobian: 3 translational rows then 3 rotational rows
    for (int r = 0; r < 3; r++)
      for (int i = 0; i < NV; i++) {
        double tp = 0, tr = 0;
        for (int k = 0; k < 3; k++) {
          tp += c.frame[3 * r + k] * (jp2[k * NV + i] - jp1[k * NV + i]);
          tr += c.frame[3 * r + k] * (jr2[k * NV + i] - jr1[k * NV + i]);
        }
        Jc[r][i] = tp; Jc[3 + r][i] = tr;
      }
    double tran = mm_body_invweight0[b1][0] + mm_body_invweight0[b2][0];
    c.efc_address = (int)d.efc_pos.size();
    double K, B, imp;
    kbi(c.solref, c.solimp, c.dist, &K, &B, &imp);
    double diag0 = tran * (1 + c.friction[0] * c.friction[0]);
    double R0 = std::max(MINVAL, (1 - imp) / imp * diag0);
    double Rpy = 2 * c.friction[0] * c.friction[0] * R0;  // impratio = 1
    for (int k = 1; k < c.dim; k++) {
      // rows J_n +/- mu_k J_k ; k = 1,2 tangents, 3 torsion (rotation about the normal), 4,5 rolling
      const double* Jk = k < 3 ? Jc[k] : (k == 3 ? Jc[3] : Jc[k]);
      double mu = c.friction[k - 1];
      for (int sg = 0; sg < 2; sg++) {
        for (int i = 0; i < NV; i++) J[i] = Jc[0][i] + (sg == 0 ? mu : -mu) * Jk[i];
        add_row(d, J, c.dist, diag0, c.solref, c.solimp, 2, Rpy);
      }
    }
  }
  d.nefc = (int)d.efc_pos.size();
  d.efc_force.assign(d.nefc, 0.0);
}

// ------------------------------------------------------------------------------------------------
// Newton solver on the primal (acceleration) problem
// ------------------------------------------------------------------------------------------------
struct Solver {
  Data& d;
  int n;
  std::vector<double> Jaref, Jv;
  double Ma[NV], grad[NV], Mgrad[NV], search[NV], Mv[NV];
  double cost, gauss;
  explicit Solver(Data& dd) : d(dd), n(dd.nefc), Jaref(dd.nefc), Jv(dd.nefc) {}

  void mulM(double* r, const double* x) {
    for (int i = 0; i < NV; i++) { double s = 0; for (int k = 0; k < NV; k++) s += d.M[i * NV + k] * x[k]; r[i] = s; }
  }
  void mulJ(double* r, const double* x) {
    for (int e = 0; e < n; e++) { double s = 0; const double* J = &d.efc_J[(size_t)e * NV]; for (int k = 0; k < NV; k++) s += J[k] * x[k]; r[e] = s; }
  }
  bool active(int e) const { return d.efc_type[e] == 0 || Jaref[e] < 0; }
  void update_constraint() {
    double c = 0;
    for (int e = 0; e < n; e++) {
      if (active(e)) { d.efc_force[e] = -d.efc_D[e] * Jaref[e]; c += 0.5 * d.efc_D[e] * Jaref[e] * Jaref[e]; }
      else d.efc_force[e] = 0;
    }
    for (int i = 0; i < NV; i++) {
      double s = 0;
      for (int e = 0; e < n; e++) s += d.efc_J[(size_t)e * NV + i] * d.efc_force[e];
      d.qfrc_constraint[i] = s;
    }
    gauss = 0;
    for (int i = 0; i < NV; i++) gauss += 0.5 * (Ma[i] - d.qfrc_smooth[i]) * (d.qacc[i] - d.qacc_smooth[i]);
    cost = c + gauss;
  }
  void update_gradient() {
    for (int i = 0; i < NV; i++) grad[i] = Ma[i] - d.qfrc_smooth[i] - d.qfrc_constraint[i];
    static thread_local double H[NV * NV];
    std::memcpy(H, d.M, sizeof(double) * NV * NV);
    for (int e = 0; e < n; e++) {
      if (!active(e)) continue;
      const double* J = &d.efc_J[(size_t)e * NV];
      double D = d.efc_D[e];
      for (int i = 0; i < NV; i++) {
        if (J[i] == 0) continue;
        double di = D * J[i];
        for (int k = 0; k <= i; k++) H[i * NV + k] += di * J[k];
      }
    }
    chol_factor(H, NV);
    for (int i = 0; i < NV; i++) Mgrad[i] = grad[i];
    chol_solve(H, NV, Mgrad);
    for (int i = 0; i < NV; i++) search[i] = -Mgrad[i];
  }
  // derivative and curvature of the cost along the search direction at step alpha
  void eval(double alpha, double qg1, double qg2, double* d1, double* d2) {
    double a1 = qg1 + 2 * alpha * qg2, a2 = 2 * qg2;
    for (int e = 0; e < n; e++) {
      double x = Jaref[e] + alpha * Jv[e];
      if (d.efc_type[e] == 0 || x < 0) { a1 += d.efc_D[e] * Jv[e] * x; a2 += d.efc_D[e] * Jv[e] * Jv[e]; }
    }
    d.stats.ls_evals++;
    *d1 = a1; *d2 = a2;
  }
  // exact 1-D minimisation: safeguarded Newton on the (monotone, piecewise-linear) derivative
  double linesearch(double scale_inv) {
    double snorm = 0;
    for (int i = 0; i < NV; i++) snorm += search[i] * search[i];
    snorm = std::sqrt(snorm);
    if (snorm < MINVAL) return 0;
    double gtol = 1e-8 * 0.01 * snorm * scale_inv;
    mulM(Mv, search);
    mulJ(Jv.data(), search);
    double qg1 = 0, qg2 = 0;
    for (int i = 0; i < NV; i++) { qg1 += search[i] * (Ma[i] - d.qfrc_smooth[i]); qg2 += 0.5 * search[i] * Mv[i]; }
    double lo = 0, hi = -1, a = 0, d1, d2;
    eval(0, qg1, qg2, &d1, &d2);
    if (d1 >= 0) return 0;  // not a descent direction (converged)
    for (int it = 0; it < 50; it++) {
      double an = a - d1 / d2;
      if (hi > 0 && !(an > lo && an < hi)) an = 0.5 * (lo + hi);
      a = an;
      eval(a, qg1, qg2, &d1, &d2);
      if (std::fabs(d1) < gtol) break;
      if (d1 < 0) lo = a; else hi = a;
    }
    return a;
  }
};

static void fwd_constraint(Data& d) {
  d.solver_niter = 0;
  if (d.nefc == 0) {
    std::memcpy(d.qacc, d.qacc_smooth, sizeof d.qacc);
    std::memset(d.qfrc_constraint, 0, sizeof d.qfrc_constraint);
    std::memcpy(d.qacc_warmstart, d.qacc, sizeof d.qacc);
    return;
  }
  Solver s(d);
  const double scale_inv = MM_MEANINERTIA * NV;  // 1/scale
  const double scale = 1.0 / scale_inv;
  // warmstart: pick the cheaper of qacc_warmstart and qacc_smooth
  double cost_ws, cost_sm;
  {
    std::memcpy(d.qacc, d.qacc_warmstart, sizeof d.qacc);
    s.mulM(s.Ma, d.qacc);
    s.mulJ(s.Jaref.data(), d.qacc);
    for (int e = 0; e < s.n; e++) s.Jaref[e] -= d.efc_aref[e];
    s.update_constraint();
    cost_ws = s.cost;
    std::memcpy(d.qacc, d.qacc_smooth, sizeof d.qacc);
    s.mulM(s.Ma, d.qacc);
    s.mulJ(s.Jaref.data(), d.qacc);
    for (int e = 0; e < s.n; e++) s.Jaref[e] -= d.efc_aref[e];
    s.update_constraint();
    cost_sm = s.cost;
    if (cost_ws < cost_sm) {
      std::memcpy(d.qacc, d.qacc_warmstart, sizeof d.qacc);
      s.mulM(s.Ma, d.qacc);
      s.mulJ(s.Jaref.data(), d.qacc);
      for (int e = 0; e < s.n; e++) s.Jaref[e] -= d.efc_aref[e];
      s.update_constraint();
    }
  }
  s.update_gradient();
  int iter = 0;
  while (iter < 100) {
    double alpha = s.linesearch(scale_inv);
    if (alpha == 0) break;
    for (int i = 0; i < NV; i++) { d.qacc[i] += alpha * s.search[i]; s.Ma[i] += alpha * s.Mv[i]; }
    for (int e = 0; e < s.n; e++) s.Jaref[e] += alpha * s.Jv[e];
    double oldcost = s.cost;
    s.update_constraint();
    s.update_gradient();
    iter++;
    double improvement = scale * (oldcost - s.cost);
    double gn = 0;
    for (int i = 0; i < NV; i++) gn += s.grad[i] * s.grad[i];
    double gradient = scale * std::sqrt(gn);
    if (improvement < 1e-8 || gradient < 1e-8) break;
  }
  d.solver_niter = iter;
  std::memcpy(d.qacc_warmstart, d.qacc, sizeof d.qacc);
}

// ------------------------------------------------------------------------------------------------
// pipeline
// ------------------------------------------------------------------------------------------------
static void fwd_actuation(Data& d) {
  std::memset(d.qfrc_actuator, 0, sizeof d.qfrc_actuator);
  for (int a = 0; a < NU; a++) {
    double c = std::max(mm_act_ctrlrange[a][0], std::min(mm_act_ctrlrange[a][1], d.ctrl[a]));
    double len, vel;
    if (mm_act_trntype[a] == 0) { len = d.qpos[mm_act_trnid[a]]; vel = d.qvel[mm_act_trnid[a]]; }
    else {
      len = vel = 0;
      for (int t = 0; t < 2; t++) { len += mm_tendon_coef[t] * d.qpos[mm_tendon_dof[t]]; vel += mm_tendon_coef[t] * d.qvel[mm_tendon_dof[t]]; }
    }
    double f = mm_act_gain[a] * c + mm_act_bias[a][0] + mm_act_bias[a][1] * len + mm_act_bias[a][2] * vel;
    d.act_saturated[a] = 0;
    if (f <= mm_act_forcerange[a][0]) { f = mm_act_forcerange[a][0]; d.act_saturated[a] = 1; }
    else if (f >= mm_act_forcerange[a][1]) { f = mm_act_forcerange[a][1]; d.act_saturated[a] = 1; }
    d.actuator_force[a] = f;
    if (mm_act_trntype[a] == 0) d.qfrc_actuator[mm_act_trnid[a]] += f;
    else for (int t = 0; t < 2; t++) d.qfrc_actuator[mm_tendon_dof[t]] += mm_tendon_coef[t] * f;
  }
}

void forward(Data& d) {
  kinematics_only(d);
  make_M(d);
  collision(d);
  make_constraint(d);
  // velocity-dependent
  for (int i = 0; i < NV; i++) d.qfrc_passive[i] = 0;
  for (int j = 0; j < NJNT; j++)
    if (mm_jnt_type[j] != 0) d.qfrc_passive[mm_jnt_dofadr[j]] = -mm_jnt_damping[j] * d.qvel[mm_jnt_dofadr[j]];
  make_bias(d);
  fwd_actuation(d);
  for (int i = 0; i < NV; i++) d.qfrc_smooth[i] = d.qfrc_passive[i] - d.qfrc_bias[i] + d.qfrc_actuator[i];
  static thread_local double L[NV * NV];
  std::memcpy(L, d.M, sizeof d.M);
  chol_factor(L, NV);
  std::memcpy(d.qacc_smooth, d.qfrc_smooth, sizeof d.qacc_smooth);
  chol_solve(L, NV, d.qacc_smooth);
  fwd_constraint(d);
  d.stats.ncon += d.ncon; d.stats.nefc += d.nefc; d.stats.newton_iters += d.solver_niter;
  d.stats.forwards++;
  for (int r = 0; r < d.nefc; r++) {
    int nz = 0;
    for (int i = 0; i < NV; i++) nz += d.efc_J[(size_t)r * NV + i] != 0.0;
    d.stats.nnzJ += nz; d.stats.nnzJ2 += (long long)nz * nz;
  }
  d.stats.max_ncon = std::max(d.stats.max_ncon, d.ncon);
  d.stats.max_nefc = std::max(d.stats.max_nefc, d.nefc);
  d.stats.max_newton = std::max(d.stats.max_newton, d.solver_niter);
}

static void quat_integrate(double* q, const double* w, double h) {
  double n = norm3(w);
  if (n * h < MINVAL) return;  // no rotation
  double ax[3] = {w[0] / n, w[1] / n, w[2] / n}, dq[4], r[4];
  axisangle2quat(dq, ax, n * h);
  mulquat(r, q, dq);
  normquat(r);
  for (int k = 0; k < 4; k++) q[k] = r[k];
}

// implicitfast: (M - h*qDeriv) a = qfrc_smooth + qfrc_constraint, semi-implicit position update
static void implicitfast(Data& d) {
  const double h = MM_TIMESTEP;
  static thread_local double MH[NV * NV];
  std::memcpy(MH, d.M, sizeof d.M);
  for (int j = 0; j < NJNT; j++)
    if (mm_jnt_type[j] != 0) MH[mm_jnt_dofadr[j] * (NV + 1)] += h * mm_jnt_damping[j];
  for (int a = 0; a < NU; a++) {
    if (d.act_saturated[a]) continue;  // clamped actuator: no velocity derivative
    double kd = mm_act_bias[a][2];
    if (mm_act_trntype[a] == 0) MH[mm_act_trnid[a] * (NV + 1)] -= h * kd;
    else
      for (int t = 0; t < 2; t++)
        for (int u = 0; u < 2; u++)
          MH[mm_tendon_dof[t] * NV + mm_tendon_dof[u]] -= h * kd * mm_tendon_coef[t] * mm_tendon_coef[u];
  }
  double acc[NV];
  for (int i = 0; i < NV; i++) acc[i] = d.qfrc_smooth[i] + d.qfrc_constraint[i];
  chol_factor(MH, NV);
  chol_solve(MH, NV, acc);
  for (int i = 0; i < NV; i++) d.qvel[i] += h * acc[i];
  for (int j = 0; j < NJNT; j++) {
    int qa = mm_jnt_qposadr[j], da = mm_jnt_dofadr[j];
    if (mm_jnt_type[j] == 0) {
      for (int k = 0; k < 3; k++) d.qpos[qa + k] += h * d.qvel[da + k];
      quat_integrate(d.qpos + qa + 3, d.qvel + da + 3, h);
    } else d.qpos[qa] += h * d.qvel[da];
  }
  d.time += h;
}

static bool bad_state(const Data& d) {
  for (int i = 0; i < NQ; i++) if (!(std::fabs(d.qpos[i]) < 1e10)) return true;
  for (int i = 0; i < NV; i++) if (!(std::fabs(d.qvel[i]) < 1e10)) return true;
  return false;
}

void step(Data& d) {
  if (bad_state(d)) reset_keyframe(d);  // mj_checkPos / mj_checkVel auto-reset (to qpos0 in MuJoCo; keyframe here)
  forward(d);
  bool badacc = false;
  for (int i = 0; i < NV; i++) if (!(std::fabs(d.qacc[i]) < 1e10)) badacc = true;
  if (badacc) { reset_keyframe(d); forward(d); }
  implicitfast(d);
  d.stats.substeps++;
}

void reset_keyframe(Data& d) {
  Stats keep = d.stats;
  int flags = d.flags;
  for (int i = 0; i < NQ; i++) d.qpos[i] = mm_key_qpos[i];
  for (int i = 0; i < NV; i++) { d.qvel[i] = 0; d.qacc_warmstart[i] = 0; d.qacc[i] = 0; }
  for (int i = 0; i < NU; i++) d.ctrl[i] = mm_key_ctrl[i];
  d.time = 0;
  d.stats = keep;
  d.flags = flags;
}

}  // namespace orc
