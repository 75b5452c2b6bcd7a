// ORACLE - TEST INFRASTRUCTURE ONLY (see engine.h / hotpath.h).
#include "hotpath.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace orc {

static const double HOME_QPOS[7] = {1.5708, -0.2, 0.0, -2.1, 0.0, 1.8, 0.785};             // controller.py:8
static const double TARGET_ORI[9] = {0.0, 1.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, -1.0};           // controller.py:12-18

// ---- pose utilities -------------------------------------------------------------------------
// pose_utils.py:48-82 : rotation matrix -> (x, y, z, w), four-branch, no sign canonicalisation
static void rotmat_to_quat_xyzw(const double* R, double* q) {
  double tr = R[0] + R[4] + R[8], x, y, z, w;
  if (tr > 0) {
    double s = 2.0 * std::sqrt(tr + 1.0);
    w = 0.25 * s; x = (R[7] - R[5]) / s; y = (R[2] - R[6]) / s; z = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    double s = 2.0 * std::sqrt(1.0 + R[0] - R[4] - R[8]);
    w = (R[7] - R[5]) / s; x = 0.25 * s; y = (R[1] + R[3]) / s; z = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    double s = 2.0 * std::sqrt(1.0 + R[4] - R[0] - R[8]);
    w = (R[2] - R[6]) / s; x = (R[1] + R[3]) / s; y = 0.25 * s; z = (R[5] + R[7]) / s;
  } else {
    double s = 2.0 * std::sqrt(1.0 + R[8] - R[0] - R[4]);
    w = (R[3] - R[1]) / s; x = (R[2] + R[6]) / s; y = (R[5] + R[7]) / s; z = 0.25 * s;
  }
  q[0] = x; q[1] = y; q[2] = z; q[3] = w;
}

// pose_utils.py:154-181 : 8-DOF and 10-DOF float32 encodings of (pos, R, gripper)
static void encode_pose(const double* p, const double* R, float g, float* out8, float* out10) {
  double q[4];
  rotmat_to_quat_xyzw(R, q);
  for (int k = 0; k < 3; k++) { out8[k] = (float)p[k]; out10[k] = (float)p[k]; }
  for (int k = 0; k < 4; k++) out8[3 + k] = (float)q[k];
  out8[7] = g;
  for (int k = 0; k < 6; k++) out10[3 + k] = (float)R[k];
  out10[9] = g;
}

// gym_env.py:252-281 : only the translation of the decoded SE(3) reaches the controller
void decode_action(const Env& e, const float* a, double target[3], float* gripper) {
  int m = e.action_mode;
  *gripper = m == ABS_POS ? a[3] : ((m == EE_POS_QUAT_G || m == EE_POS_QUAT_G_REL) ? a[7] : a[9]);
  if (m == EE_POS_QUAT_G_REL || m == EE_POS_ROT6D_G_REL) {
    for (int r = 0; r < 3; r++)
      target[r] = e.init_R[3 * r] * (double)a[0] + e.init_R[3 * r + 1] * (double)a[1] + e.init_R[3 * r + 2] * (double)a[2] +
                  e.init_pos[r];
  } else {
    for (int r = 0; r < 3; r++) target[r] = (double)a[r];
  }
}

// controller.py:21-43
void orientation_error(const double* Rc, const double* Rt, double out[3]) {
  double E[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) E[3 * i + j] = Rt[3 * i] * Rc[3 * j] + Rt[3 * i + 1] * Rc[3 * j + 1] + Rt[3 * i + 2] * Rc[3 * j + 2];
  double tv = std::max(-1.0, std::min(1.0, (E[0] + E[4] + E[8] - 1) / 2));
  double ang = std::acos(tv);
  if (ang < 1e-6) { out[0] = out[1] = out[2] = 0; return; }
  double s = 2 * std::sin(ang);
  out[0] = (E[7] - E[5]) / s * ang; out[1] = (E[2] - E[6]) / s * ang; out[2] = (E[3] - E[1]) / s * ang;
}

// general matrix inverse by Gauss-Jordan with partial pivoting (np.linalg.inv stand-in)
static void inv_n(const double* A, double* Ai, int n) {
  double M[6][12];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) { M[i][j] = A[i * n + j]; M[i][n + j] = i == j; }
  for (int c = 0; c < n; c++) {
    int p = c;
    for (int r = c + 1; r < n; r++) if (std::fabs(M[r][c]) > std::fabs(M[p][c])) p = r;
    if (p != c) for (int j = 0; j < 2 * n; j++) std::swap(M[c][j], M[p][j]);
    double iv = 1.0 / M[c][c];
    for (int j = 0; j < 2 * n; j++) M[c][j] *= iv;
    for (int r = 0; r < n; r++) {
      if (r == c) continue;
      double f = M[r][c];
      if (f != 0) for (int j = 0; j < 2 * n; j++) M[r][j] -= f * M[c][j];
    }
  }
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Ai[i * n + j] = M[i][n + j];
}

// controller.py:87-137 (reads xpos/xmat/xaxis/xanchor as left by the last position stage: possibly stale)
void ik_compute(const Data& d, const double target[3], double qt[7]) {
  double jp[3 * NV], jr[3 * NV];
  const double* ee = d.xpos[BODY_HAND];
  jac(d, jp, jr, ee, BODY_HAND);
  double J[6][7], e[6];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 7; c++) { J[r][c] = jp[r * NV + c]; J[3 + r][c] = jr[r * NV + c]; }
  for (int k = 0; k < 3; k++) e[k] = 1.0 * (target[k] - ee[k]);
  orientation_error(d.xmat[BODY_HAND], TARGET_ORI, e + 3);
  double A[36], Ai[36];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      double s = 0;
      for (int c = 0; c < 7; c++) s += J[i][c] * J[j][c];
      A[6 * i + j] = s + (i == j ? 1e-3 : 0.0);
    }
  inv_n(A, Ai, 6);
  double P[7][6];  // J^T inv(JJT)
  for (int c = 0; c < 7; c++)
    for (int j = 0; j < 6; j++) { double s = 0; for (int i = 0; i < 6; i++) s += J[i][c] * Ai[6 * i + j]; P[c][j] = s; }
  double dq[7];
  for (int c = 0; c < 7; c++) { double s = 0; for (int j = 0; j < 6; j++) s += P[c][j] * e[j]; dq[c] = s; }
  double z[7];
  for (int c = 0; c < 7; c++) z[c] = 0.5 * (HOME_QPOS[c] - d.qpos[c]);
  for (int c = 0; c < 7; c++) {
    double s = 0;
    for (int k = 0; k < 7; k++) {
      double n = (c == k ? 1.0 : 0.0);
      for (int j = 0; j < 6; j++) n -= P[c][j] * J[j][k];
      s += n * z[k];
    }
    dq[c] += s;
  }
  double nn = 0;
  for (int c = 0; c < 7; c++) nn += dq[c] * dq[c];
  nn = std::sqrt(nn);
  if (nn > 5.0) for (int c = 0; c < 7; c++) dq[c] *= 5.0 / nn;
  for (int c = 0; c < 7; c++) {
    double q = d.qpos[c] + dq[c];
    double lo = mm_jnt_range[c][0], hi = mm_jnt_range[c][1];
    if (lo < hi) q = std::max(lo, std::min(hi, q));
    qt[c] = q;
  }
}

// controller.py:139-145
bool ik_reached(const Data& d, const double t[3]) {
  const double* ee = d.xpos[BODY_HAND];
  double s = 0;
  for (int k = 0; k < 3; k++) s += (ee[k] - t[k]) * (ee[k] - t[k]);
  return std::sqrt(s) < 0.02;
}

// cameras.py:56-104 : pinhole projection with the signed camera-z quirk
static void project(const Data& d, int cam, double fovy_deg, const double* p, float* out) {
  const double size = 224.0;
  double f = (size / 2.0) / std::tan(fovy_deg * M_PI / 180.0 / 2.0);
  double rel[3] = {p[0] - d.cam_xpos[cam][0], p[1] - d.cam_xpos[cam][1], p[2] - d.cam_xpos[cam][2]};
  const double* m = d.cam_xmat[cam];
  double c[3];
  for (int j = 0; j < 3; j++) c[j] = rel[0] * m[j] + rel[1] * m[3 + j] + rel[2] * m[6 + j];
  double depth = c[2];
  if (std::fabs(depth) < 1e-6) depth = 1e-6;
  double px = f * c[0] / depth + size / 2.0, py = -f * c[1] / depth + size / 2.0;
  out[0] = (float)(px / size); out[1] = (float)(py / size);
}

// gym_env.py:283-339 (state part + keypoints; images are out of scope)
void env_obs(const Env& e, float* o) {
  const Data& d = e.d;
  const double* ee = d.xpos[BODY_HAND];
  const double* R = d.xmat[BODY_HAND];
  float g = (float)(d.ctrl[7] / 255.0);
  for (int k = 0; k < 3; k++) o[k] = (float)ee[k];
  o[3] = g;
  for (int k = 0; k < 7; k++) o[4 + k] = (float)d.qpos[k];
  encode_pose(ee, R, g, o + 11, o + 19);
  // T_rel = inv(T_init) @ T_cur
  double Rr[9], pr[3], dp[3] = {ee[0] - e.init_pos[0], ee[1] - e.init_pos[1], ee[2] - e.init_pos[2]};
  for (int i = 0; i < 3; i++) {
    pr[i] = e.init_R[i] * dp[0] + e.init_R[3 + i] * dp[1] + e.init_R[6 + i] * dp[2];
    for (int j = 0; j < 3; j++) Rr[3 * i + j] = e.init_R[i] * R[j] + e.init_R[3 + i] * R[3 + j] + e.init_R[6 + i] * R[6 + j];
  }
  encode_pose(pr, Rr, g, o + 29, o + 37);
  for (int k = 0; k < 3; k++) { o[47 + k] = (k == e.bin_idx); o[50 + k] = (k == e.obj_idx); }
  // keypoints: KEYPOINT_BODIES = 3 objects, 3 bins, hand (constants.py:29-37)
  static const int kb[7] = {16, 17, 18, 13, 14, 15, 9};
  for (int k = 0; k < 7; k++) {
    project(d, 0, 45.0, d.xpos[kb[k]], o + 53 + 2 * k);
    project(d, 2, 128.0, d.xpos[kb[k]], o + 67 + 2 * k);
  }
  o[81] = e.tgt_obj_kp[0]; o[82] = e.tgt_obj_kp[1]; o[83] = e.tgt_bin_kp[0]; o[84] = e.tgt_bin_kp[1];
}

// gym_env.py:477-534 (placement drawn on the host with numpy's PCG64 and passed in as obj_xy)
void env_reset(Env& e, const double* obj_xy, int obj_idx, int bin_idx, const double* yaw_cs) {
  reset_keyframe(e.d);
  forward(e.d);
  e.step_count = 0;
  if (obj_xy) {  // randomization.py:52-65 + env.py:160-161
    for (int o = 0; o < 3; o++) {
      double* q = e.d.qpos + 9 + 7 * o;
      q[0] = obj_xy[2 * o]; q[1] = obj_xy[2 * o + 1]; q[2] = 0.26; q[3] = 1; q[4] = q[5] = q[6] = 0;
      if (yaw_cs) { q[3] = yaw_cs[2 * o]; q[6] = yaw_cs[2 * o + 1]; }  // randomize_yaw=True, randomization.py:55-62
    }
    forward(e.d);
  }
  for (int k = 0; k < 3; k++) e.init_pos[k] = e.d.xpos[BODY_HAND][k];
  for (int k = 0; k < 9; k++) e.init_R[k] = e.d.xmat[BODY_HAND][k];
  e.has_grasped = e.has_lifted = e.above_target = e.has_placed = false;
  e.hwm_set = false;
  for (int k = 0; k < 5; k++) e.hwm[k] = 0;
  e.obj_idx = obj_idx; e.bin_idx = bin_idx;
  project(e.d, 0, 45.0, e.d.xpos[BODY_OBJ0 + obj_idx], e.tgt_obj_kp);
  project(e.d, 0, 45.0, e.d.xpos[BODY_BIN0 + bin_idx], e.tgt_bin_kp);
  fsm_reset(e);
}

// gym_env.py:341-350
static bool robot_collision(const Data& d) {
  for (const auto& c : d.contact) {
    int b1 = mm_geom_body[c.geom1], b2 = mm_geom_body[c.geom2];
    bool r1 = b1 >= 1 && b1 <= 11, r2 = b2 >= 1 && b2 <= 11;
    bool o1 = b1 >= BODY_TABLE && b1 < BODY_OBJ0, o2 = b2 >= BODY_TABLE && b2 < BODY_OBJ0;
    if ((r1 && o2) || (r2 && o1)) return true;
  }
  return false;
}

static double dist3(const double* a, const double* b) {
  return std::sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
}

// gym_env.py:352-434
static void staged_reward(Env& e, double* reward, bool* done) {
  const Data& d = e.d;
  const double D_MAX = 0.5, GRASP_Z = 0.35, LIFT_Z = 0.42;
  const double* obj = d.xpos[BODY_OBJ0 + e.obj_idx];
  const double* bin = d.xpos[BODY_BIN0 + e.bin_idx];
  const double* ee = d.xpos[BODY_HAND];
  bool closed = d.ctrl[7] == 0.0;
  if (!e.has_grasped && obj[2] > GRASP_Z && closed) e.has_grasped = true;
  if (!e.has_lifted && obj[2] > LIFT_Z && closed) e.has_lifted = true;
  double xy = std::hypot(obj[0] - bin[0], obj[1] - bin[1]);
  if (!e.above_target && e.has_lifted && xy < 0.06) e.above_target = true;
  bool placed = xy < 0.05 && obj[2] < bin[2] + 0.06;
  if (!e.has_placed && placed) e.has_placed = true;
  double r[5];
  r[0] = e.has_grasped ? 1.0 : 1.0 - std::min(dist3(ee, obj) / D_MAX, 1.0);
  r[1] = !e.has_grasped ? 0.0 : (e.has_lifted ? 1.0 : std::max(0.0, std::min((obj[2] - 0.30) / (LIFT_Z - 0.30), 1.0)));
  r[2] = !e.has_lifted ? 0.0 : (e.above_target ? 1.0 : 1.0 - std::min(xy / D_MAX, 1.0));
  r[3] = !e.above_target ? 0.0 : (e.has_placed ? 1.0 : 1.0 - std::max(0.0, std::min((obj[2] - bin[2]) / 0.25, 1.0)));
  r[4] = !e.has_placed ? 0.0 : 1.0 - std::min(dist3(ee, e.init_pos) / D_MAX, 1.0);
  e.hwm_set = true;
  for (int k = 0; k < 5; k++) e.hwm[k] = std::max(e.hwm[k], r[k]);
  if (robot_collision(d)) { *reward = -1.0; *done = true; return; }
  double s = 0;
  bool all = true;
  for (int k = 0; k < 5; k++) { s += e.hwm[k]; all = all && e.hwm[k] >= 0.90; }
  *reward = s / 5.0;
  *done = all;
}

// gym_env.py:436-470
static void compute_reward(Env& e, double* reward, bool* success) {
  const Data& d = e.d;
  const double* obj = d.xpos[BODY_OBJ0 + e.obj_idx];
  const double* bin = d.xpos[BODY_BIN0 + e.bin_idx];
  const double* ee = d.xpos[BODY_HAND];
  double xy = std::hypot(obj[0] - bin[0], obj[1] - bin[1]);
  bool succ = xy < 0.05 && obj[2] < bin[2] + 0.06;
  if (e.reward_type == SPARSE) { *reward = succ ? 1.0 : 0.0; *success = succ; return; }
  if (e.reward_type == STAGED) { staged_reward(e, reward, success); return; }
  double r = 0.0;
  r -= dist3(ee, obj);
  if (obj[2] > 0.30) { r += 2.0; r -= dist3(obj, bin); }
  if (succ) r += 10.0;
  *reward = r; *success = succ;
}

// gym_env.py:536-581
void env_step(Env& e, const float* action, float* obs, double* reward, int* terminated, int* truncated, int* success,
              float* rc) {
  double target[3];
  float g;
  decode_action(e, action, target, &g);
  e.d.ctrl[7] = g > 0.5f ? 255.0 : 0.0;
  for (int s = 0; s < 16; s++) {
    double q[7];
    ik_compute(e.d, target, q);
    for (int k = 0; k < 7; k++) e.d.ctrl[k] = q[k];
    step(e.d);
  }
  forward(e.d);
  e.step_count++;
  double r;
  bool succ;
  compute_reward(e, &r, &succ);
  if (e.reward_type == STAGED) {
    *terminated = (r < 0 || succ);
    *success = succ && r >= 0;
    if (rc && e.hwm_set) {
      double s = 0;
      for (int k = 0; k < 5; k++) { rc[1 + k] = (float)(e.hwm[k] / 5.0); s += e.hwm[k] / 5.0; }
      rc[0] = (float)s;
    }
  } else {
    *terminated = succ;
    *success = succ;
  }
  *truncated = e.step_count >= e.max_episode_steps;
  *reward = r;
  if (obs) env_obs(e, obs);
}

// ---- FSM (pick_and_place.py:167-277) ----------------------------------------------------------
void fsm_reset(Env& e) {
  e.fsm_state = 1; e.task_index = 0; e.settle_counter = 0; e.gripper_open = 1; e.has_target = 0;
  for (int k = 0; k < 3; k++) e.target[k] = e.transit_end[k] = 0;
}

void fsm_plan(Env& e, int n) {
  const Data& d = e.d;
  const double* obj = d.xpos[BODY_OBJ0 + e.obj_idx];
  const double* bin = d.xpos[BODY_BIN0 + e.bin_idx];
  auto set_t = [&](double x, double y, double z) { e.target[0] = x; e.target[1] = y; e.target[2] = z; e.has_target = 1; };
  switch (e.fsm_state) {
    case 1:  // IDLE
      if (e.task_index >= 1) { e.fsm_state = 11; return; }
      e.gripper_open = 1;
      set_t(obj[0], obj[1], 0.44);
      e.fsm_state = 2;
      return;
    case 2:  // PRE_GRASP
      if (ik_reached(d, e.target)) { set_t(obj[0], obj[1], 0.36); e.fsm_state = 3; }
      return;
    case 3:  // GRASP
      if (ik_reached(d, e.target)) { e.gripper_open = 0; e.settle_counter = 150; e.fsm_state = 4; }
      return;
    case 4:  // CLOSE_GRIPPER
      e.settle_counter -= n;
      if (e.settle_counter <= 0) { set_t(obj[0], obj[1], 0.55); e.fsm_state = 5; }
      return;
    case 5:  // LIFT
      if (ik_reached(d, e.target)) {
        e.transit_end[0] = bin[0]; e.transit_end[1] = bin[1]; e.transit_end[2] = 0.55;
        e.fsm_state = 6;
      }
      return;
    case 6: {  // MOVE_TO_BIN
      double diff[3] = {e.transit_end[0] - e.target[0], e.transit_end[1] - e.target[1], e.transit_end[2] - e.target[2]};
      double dist = std::sqrt(diff[0] * diff[0] + diff[1] * diff[1] + diff[2] * diff[2]);
      double st = 0.001 * n;
      if (dist > st) for (int k = 0; k < 3; k++) e.target[k] += diff[k] * (st / dist);
      else for (int k = 0; k < 3; k++) e.target[k] = e.transit_end[k];
      if (dist <= 0.02) { e.settle_counter = 100; e.fsm_state = 7; }
      return;
    }
    case 7:  // SETTLE_AT_BIN
      e.settle_counter -= n;
      if (e.settle_counter <= 0) { set_t(bin[0], bin[1], 0.45); e.fsm_state = 8; }
      return;
    case 8:  // LOWER_TO_BIN
      if (ik_reached(d, e.target)) { e.gripper_open = 1; e.settle_counter = 150; e.fsm_state = 9; }
      return;
    case 9:  // RELEASE
      e.settle_counter -= n;
      if (e.settle_counter <= 0) { set_t(0.0, 0.3, 0.55); e.fsm_state = 10; }
      return;
    case 10:  // RETREAT
      if (ik_reached(d, e.target)) { e.task_index += 1; e.fsm_state = 1; }
      return;
    default:
      return;
  }
}

// scripts/generate_dataset.py:145-148 : abs_pos action from the FSM target
void fsm_action(const Env& e, float a[4]) {
  const double* src = e.has_target ? e.target : e.d.xpos[BODY_HAND];
  for (int k = 0; k < 3; k++) a[k] = (float)src[k];
  a[3] = e.gripper_open ? 1.0f : 0.0f;
}

}  // namespace orc
