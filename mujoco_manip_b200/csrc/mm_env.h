// Environment layer around the physics substep: what PickPlaceGymEnv.step / reset do per env
// (mujoco_manip/gym_env.py:477-581) plus the scripted FSM (mujoco_manip/pick_and_place.py:167-277).
#pragma once
#include "mm_core.h"

namespace mm {

enum { MODE_ABS_POS = 0, MODE_QUAT = 1, MODE_ROT6D = 2, MODE_QUAT_REL = 3, MODE_ROT6D_REL = 4 };
enum { REWARD_DENSE = 0, REWARD_SPARSE = 1, REWARD_STAGED = 2 };
constexpr int OBS_DIM = 85;      // 53 state floats + 32 keypoint floats (layout: oracle/hotpath.cpp env_obs)
constexpr int ACTION_STRIDE = 10;
constexpr int ACTION_REPEAT = 16;  // constants.py:26
constexpr int FSM_TASKS_STRIDE = 20, FSM_MAX_TASKS = 9;

// Env-major state arrays in HBM (FP64 storage regardless of the compute type; a G-lane group reads
// one env's contiguous rows with consecutive lanes -> coalesced).
struct StatePtrs {
  double* qpos;      // [N,30]
  double* qvel;      // [N,27]
  double* ctrl;      // [N,8]
  double* warm;      // [N,27]  qacc_warmstart
  double* tinit;     // [N,12]  initial EE pose: pos(3) + R(9)   (gym_env.py:245-250)
  double* eepose;    // [N,12]  EE pose after the last forward
  double* fsm_f;     // [N,6]   FSM target(3) + transit_end(3)
  double* hwm;       // [N,5]   staged-reward high-water marks
  double* kin;       // [N,18]  qpos of the last position stage: arm + fingers (9), cube positions (9)
  int* step_count;   // [N]
  int* task;         // [N,2]   object index, bin index
  int* fsm_i;        // [N,5]   state(1..11), task_index, settle_counter, gripper_open, has_target
  int* fsm_tasks;    // [N,20]  the FSM's task list: count, then (object, bin) index pairs (pick_and_place.py:91)
  int* flags;        // [N]     bits 0..3 staged stickies, bit 4 hwm initialised
  int* diag;         // [N,4]   ncon, newton iterations (last forward), overflow bits, non-finite resets
};

struct StepOut {
  float* obs;            // [N,85]
  float* reward;         // [N]
  unsigned char* terminated;  // [N]
  unsigned char* truncated;   // [N]
  unsigned char* success;     // [N]
  float* reward_components;   // [N,6] or null
};

template <class T, int G>
MM_HDN void load_state(const Grp<G>& g, Scratch<T>& s, const StatePtrs& st, long e) {
  for (int i = g.lane; i < NQ; i += G) s.qpos[i] = (T)st.qpos[e * NQ + i];
  for (int i = g.lane; i < NV; i += G) s.qvel[i] = (T)st.qvel[e * NV + i];
  if (g.lane == 0) s.warm_g = st.warm + e * NV;
  for (int i = g.lane; i < NU; i += G) s.ctrl[i] = (T)st.ctrl[e * NU + i];
  // (only members of the persistent part: stage A runs on a shared-memory slice that ends before the solver's temporaries)
  if (g.lane == 0) { s.overflow = 0; s.ncon = 0; s.nbox = 0; s.ncvx = 0; s.qbase = 0; s.nsurv = 0; s.niter = 0; s.prof = 0; for (int k = 0; k < 8; k++) s.tph[k] = 0; }
  g.sync();
}

template <class T, int G>
MM_HDN void store_state(const Grp<G>& g, const Scratch<T>& s, const StatePtrs& st, long e) {
  for (int i = g.lane; i < NQ; i += G) st.qpos[e * NQ + i] = (double)s.qpos[i];
  for (int i = g.lane; i < NV; i += G) st.qvel[e * NV + i] = (double)s.qvel[i];  // (qacc_warmstart is written by solve)
  for (int i = g.lane; i < NU; i += G) st.ctrl[e * NU + i] = (double)s.ctrl[i];
  for (int i = g.lane; i < 12; i += G)
    st.eepose[e * 12 + i] = (double)(i < 3 ? s.bpos[DB_HAND][i] : s.bR[DB_HAND][i - 3]);
  // kinematics of the last position stage (what data.xpos / mj_jac describe in the reference)
  for (int i = g.lane; i < 18; i += G) st.kin[e * 18 + i] = (double)(i < 9 ? s.tmp6[KIN_ROW + i / 6][i % 6] : s.bpos[DB_CUBE0 + (i - 9) / 3][(i - 9) % 3]);
  if (g.lane == 0) {
    st.diag[e * 4 + 0] = s.ncon; st.diag[e * 4 + 1] = s.niter; st.diag[e * 4 + 2] |= s.overflow;
  }
}

// pose_utils.py:48-82
template <class T>
MM_HD void rotmat_to_quat_xyzw(const T* R, T* q) {
  T tr = R[0] + R[4] + R[8], x, y, z, w;
  if (tr > 0) {
    T sc = 2 * tsqrt(tr + 1);
    w = (T)0.25 * sc; x = (R[7] - R[5]) / sc; y = (R[2] - R[6]) / sc; z = (R[3] - R[1]) / sc;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    T sc = 2 * tsqrt(1 + R[0] - R[4] - R[8]);
    w = (R[7] - R[5]) / sc; x = (T)0.25 * sc; y = (R[1] + R[3]) / sc; z = (R[2] + R[6]) / sc;
  } else if (R[4] > R[8]) {
    T sc = 2 * tsqrt(1 + R[4] - R[0] - R[8]);
    w = (R[2] - R[6]) / sc; x = (R[1] + R[3]) / sc; y = (T)0.25 * sc; z = (R[5] + R[7]) / sc;
  } else {
    T sc = 2 * tsqrt(1 + R[8] - R[0] - R[4]);
    w = (R[3] - R[1]) / sc; x = (R[2] + R[6]) / sc; y = (R[5] + R[7]) / sc; z = (T)0.25 * sc;
  }
  q[0] = x; q[1] = y; q[2] = z; q[3] = w;
}

MM_HD void encode_pose(const double* p, const double* R, float gr, float* o8, float* o10) {
  double q[4];
  rotmat_to_quat_xyzw(R, q);
  for (int k = 0; k < 3; k++) { o8[k] = (float)p[k]; o10[k] = (float)p[k]; }
  for (int k = 0; k < 4; k++) o8[3 + k] = (float)q[k];
  o8[7] = gr;
  for (int k = 0; k < 6; k++) o10[3 + k] = (float)R[k];
  o10[9] = gr;
}

// cameras.py:85-104 pinhole projection (signed camera-z quirk kept)
MM_HD void project_kp(const double* cpos, const double* cmat, double f, const double* p, float* out) {
  double rel[3] = {p[0] - cpos[0], p[1] - cpos[1], p[2] - cpos[2]}, c[3];
  for (int j = 0; j < 3; j++) c[j] = rel[0] * cmat[j] + rel[1] * cmat[3 + j] + rel[2] * cmat[6 + j];
  double depth = c[2];
  if (fabs(depth) < 1e-6) depth = 1e-6;
  out[0] = (float)((f * c[0] / depth + 112.0) / 224.0);
  out[1] = (float)((-f * c[1] / depth + 112.0) / 224.0);
}

constexpr double F_OVERHEAD = 270.39191851172757;  // 112 / tan(22.5 deg)
constexpr double F_WRIST = 54.62604669662306;      // 112 / tan(64 deg)
MM_HD void bin_pos(int b, double* p) {  // pick_and_place_scene.xml:64,78,92
  p[0] = b == 0 ? -0.3 : (b == 1 ? 0.0 : 0.3); p[1] = b == 1 ? 0.65 : 0.55; p[2] = 0.24;
}

// observation packing (gym_env.py:283-339, state part + keypoints).  Executed by one lane in FP64 on
// the FK results (the amount of work is ~200 flops; it is not on the critical path).
template <class T>
MM_HDN void write_obs(const Scratch<T>& s, const StatePtrs& st, long e, float* o, const float* tgt_kp) {
  double ee[3], R[9];
  for (int k = 0; k < 3; k++) ee[k] = (double)s.bpos[DB_HAND][k];
  for (int k = 0; k < 9; k++) R[k] = (double)s.bR[DB_HAND][k];
  float gr = (float)((double)s.ctrl[7] / 255.0);
  for (int k = 0; k < 3; k++) o[k] = (float)ee[k];
  o[3] = gr;
  for (int k = 0; k < 7; k++) o[4 + k] = (float)(double)s.qpos[k];
  encode_pose(ee, R, gr, o + 11, o + 19);
  const double* ti = st.tinit + e * 12;
  const double* Ri = ti + 3;
  double Rr[9], pr[3], dp[3] = {ee[0] - ti[0], ee[1] - ti[1], ee[2] - ti[2]};
  for (int i = 0; i < 3; i++) {
    pr[i] = Ri[i] * dp[0] + Ri[3 + i] * dp[1] + Ri[6 + i] * dp[2];
    for (int j = 0; j < 3; j++) Rr[3 * i + j] = Ri[i] * R[j] + Ri[3 + i] * R[3 + j] + Ri[6 + i] * R[6 + j];
  }
  encode_pose(pr, Rr, gr, o + 29, o + 37);
  int obj = st.task[e * 2], bin = st.task[e * 2 + 1];
  for (int k = 0; k < 3; k++) { o[47 + k] = (k == bin) ? 1.f : 0.f; o[50 + k] = (k == obj) ? 1.f : 0.f; }
  // keypoints: 3 cubes, 3 bins, hand for the overhead and the wrist camera
  const double ocp[3] = {0, 0, 2.0}, ocm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  // wrist camera (env.py:57-64): pos (-0.07, 0, 0.055), quat (-0.0616, -0.7044, 0.7044, 0.0616) normalised
  const double wq0 = -0.0616, wq1 = -0.7044, wq2 = 0.7044, wq3 = 0.0616;
  double wn = 1.0 / sqrt(wq0 * wq0 + wq1 * wq1 + wq2 * wq2 + wq3 * wq3);
  double wq[4] = {wq0 * wn, wq1 * wn, wq2 * wn, wq3 * wn}, wr[9], wm[9], wp[3];
  quat2mat(wr, wq);
  matmul3(wm, R, wr);
  const double wl[3] = {-0.07, 0.0, 0.055};
  rot(wp, R, wl);
  for (int k = 0; k < 3; k++) wp[k] += ee[k];
  for (int k = 0; k < 7; k++) {
    double p[3];
    if (k < 3) for (int a = 0; a < 3; a++) p[a] = (double)s.bpos[DB_CUBE0 + k][a];
    else if (k < 6) bin_pos(k - 3, p);
    else for (int a = 0; a < 3; a++) p[a] = ee[a];
    project_kp(ocp, ocm, F_OVERHEAD, p, o + 53 + 2 * k);
    project_kp(wp, wm, F_WRIST, p, o + 67 + 2 * k);
  }
  for (int k = 0; k < 4; k++) o[81 + k] = tgt_kp[k];
}

MM_HD double dist3d(const double* a, const double* b) {
  return sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
}

// rewards + termination (gym_env.py:352-470, 562-577); one lane, FP64
template <class T>
MM_HDN void write_reward(const Scratch<T>& s, const StatePtrs& st, long e, int reward_type, int max_steps,
                         bool robot_collision, const StepOut& out) {
  int obj = st.task[e * 2], bin = st.task[e * 2 + 1];
  double op[3], bp[3], ee[3];
  for (int k = 0; k < 3; k++) { op[k] = (double)s.bpos[DB_CUBE0 + obj][k]; ee[k] = (double)s.bpos[DB_HAND][k]; }
  bin_pos(bin, bp);
  double xy = sqrt((op[0] - bp[0]) * (op[0] - bp[0]) + (op[1] - bp[1]) * (op[1] - bp[1]));
  bool succ = xy < 0.05 && op[2] < bp[2] + 0.06;
  double r;
  bool term, info_succ;
  if (reward_type == REWARD_SPARSE) { r = succ ? 1.0 : 0.0; term = succ; info_succ = succ; }
  else if (reward_type == REWARD_STAGED) {
    const double D_MAX = 0.5, GRASP_Z = 0.35, LIFT_Z = 0.42;
    int f = st.flags[e];
    bool closed = (double)s.ctrl[7] == 0.0;
    if (!(f & 1) && op[2] > GRASP_Z && closed) f |= 1;
    if (!(f & 2) && op[2] > LIFT_Z && closed) f |= 2;
    if (!(f & 4) && (f & 2) && xy < 0.06) f |= 4;
    if (!(f & 8) && succ) f |= 8;
    double rr[5];
    const double* ti = st.tinit + e * 12;
    double dm;
    dm = dist3d(ee, op) / D_MAX; rr[0] = (f & 1) ? 1.0 : 1.0 - (dm < 1.0 ? dm : 1.0);
    double lift = (op[2] - 0.30) / (LIFT_Z - 0.30);
    rr[1] = !(f & 1) ? 0.0 : ((f & 2) ? 1.0 : (lift < 0 ? 0.0 : (lift > 1 ? 1.0 : lift)));
    dm = xy / D_MAX; rr[2] = !(f & 2) ? 0.0 : ((f & 4) ? 1.0 : 1.0 - (dm < 1.0 ? dm : 1.0));
    double ha = (op[2] - bp[2]) / 0.25;
    rr[3] = !(f & 4) ? 0.0 : ((f & 8) ? 1.0 : 1.0 - (ha < 0 ? 0.0 : (ha > 1 ? 1.0 : ha)));
    dm = dist3d(ee, ti) / D_MAX; rr[4] = !(f & 8) ? 0.0 : 1.0 - (dm < 1.0 ? dm : 1.0);
    double* hw = st.hwm + e * 5;
    double sum = 0;
    bool all = true;
    for (int k = 0; k < 5; k++) { if (rr[k] > hw[k]) hw[k] = rr[k]; sum += hw[k]; all = all && hw[k] >= 0.90; }
    f |= 16;
    st.flags[e] = f;
    bool done;
    if (robot_collision) { r = -1.0; done = true; }
    else { r = sum / 5.0; done = all; }
    term = r < 0 || done;
    info_succ = done && r >= 0;
    if (out.reward_components) {
      float* rc = out.reward_components + e * 6;
      double tot = 0;
      for (int k = 0; k < 5; k++) { rc[1 + k] = (float)(hw[k] / 5.0); tot += hw[k] / 5.0; }
      rc[0] = (float)tot;
    }
  } else {
    r = -dist3d(ee, op);
    if (op[2] > 0.30) { r += 2.0; r -= dist3d(op, bp); }
    if (succ) r += 10.0;
    term = succ; info_succ = succ;
  }
  int sc = st.step_count[e] + 1;
  st.step_count[e] = sc;
  out.reward[e] = (float)r;
  out.terminated[e] = term;
  out.truncated[e] = sc >= max_steps;
  out.success[e] = info_succ;
}

template <class T, int G>
MM_HDN bool any_robot_collision(const Grp<G>& g, const Scratch<T>& s, const Work<T>& w) {
  int hit = 0;
  for (int c = g.lane; c < s.ncon; c += G) hit |= (w.cmeta[c] >> META_ROBOBS_BIT) & 1;
  return g.any(hit);
}

template <class T, int G>
MM_HDN bool state_bad(const Grp<G>& g, const Scratch<T>& s) {
  int bad = 0;
  for (int i = g.lane; i < NQ; i += G) bad |= !(tabs(s.qpos[i]) < (T)1e10);
  for (int i = g.lane; i < NV; i += G) bad |= !(tabs(s.qvel[i]) < (T)1e10);
  return g.any(bad);
}

// ------------------------------------------------------------------------------------------------
// One PickPlaceGymEnv.step (gym_env.py:536-581) as a sequence of batch-wide STAGES.  The 16 x (IK, mj_step) +
// trailing mj_forward of an env are 17 rounds of
//   stage A  (group per env)   [first round: load state, decode action; later rounds: fused behind stage C of the
//                              round before, same kernel]  IK -> position + velocity stage
//                              (kinematics, CRBA, RNEA, actuation, qacc_smooth) -> broad phase -> box / plane
//                              narrow phase -> the env's convex candidates are pushed on the batch-wide queue
//   convex   (warp per PAIR)   GJK + EPA of every queued (env, geom pair): a pile-up env's 25 hull pairs run on 25
//                              warps instead of one after the other on the env's own warp
//   stage C  (group per env)   contact list assembly (candidate order: results do not depend on the schedule) ->
//                              constraint rows -> Newton solver -> implicitfast integration
//                              [last round: reward, termination, observation, state store]
// The env's image between stages (Scratch<T>[0, SCRATCH_PERSIST)) lives in global memory (L2 resident).
// ------------------------------------------------------------------------------------------------
struct CvxItem { int env, ci; };

template <class T>
struct CvxQueue {
  CvxItem* items;    // [cap]
  CvxRes<T>* res;    // [cap]
  int* count;        // items pushed in this round
  int* head;         // next item to be taken (convex kernel)
  int cap;
};

// contact-rich envs of a round: stage A lists them, and stage C runs them with a whole CTA each (Grp<128>) next to the
// warp-per-env launch of the others
struct HeavyList {
  unsigned char* flag;  // [N] 1 = listed this round (the warp-per-env stage C skips it); null = feature off
  int* items;           // [cap] env ids
  int* count;           // listed this round
  int min_load;         // box contacts + convex candidates from which an env counts as contact-rich
};

#ifdef __CUDA_ARCH__
#define MM_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#else
#define MM_ATOMIC_ADD(p, v) ([&]() { int old_ = *(p); *(p) += (v); return old_; }())
#endif

// bytes of Scratch<T> stage A may touch: the persistent image + the H / tmp6 scratch (kinematics, dynamics, clip polygons)
template <class T> MM_HDN constexpr size_t scratch_a_bytes() { return (offsetof(Scratch<T>, pairK_s) + 15) / 16 * 16; }
template <class T> MM_HDN constexpr int ctx_stride() { return (int)((SCRATCH_PERSIST(T) + 15) / 16 * 16); }

template <class T, int G>
MM_HD void ctx_copy(const Grp<G>& g, void* dst, const void* src) {
  constexpr int NW = ctx_stride<T>() / 16;
  struct alignas(16) W16 { unsigned x, y, z, w; };
  const W16* s_ = reinterpret_cast<const W16*>(src);
  W16* d_ = reinterpret_cast<W16*>(dst);
  for (int i = g.lane; i < NW; i += G) d_[i] = s_[i];
  g.sync();
}

// Image load / store of a stage kernel.  A full warp per env on the device: ONE bulk-async copy by the TMA unit
// (cp.async.bulk, 4.3 KB, completion on an mbarrier in the env's scratch / a bulk group) instead of 9 LDG.128 + 9 STS.128
// per lane through the registers; -DMM_TMA_IMAGE=0, sub-warp groups, the CTA-per-env path and the host build copy by lanes.
#ifndef MM_TMA_IMAGE
#define MM_TMA_IMAGE 1
#endif
template <class T, int G>
MM_HD void ctx_load(const Grp<G>& g, Scratch<T>& s, const void* src) {
#if defined(__CUDA_ARCH__) && MM_TMA_IMAGE
  if constexpr (G == 32) {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&s.mbar), dst = (unsigned)__cvta_generic_to_shared(&s);
    const unsigned bytes = (unsigned)ctx_stride<T>();
    if (g.lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                   "r"(bytes), "r"(bar)
                   : "memory");
    }
    g.sync();  // the barrier is initialised before any lane polls it
    unsigned done;
    do {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                   : "=r"(done)
                   : "r"(bar)
                   : "memory");
    } while (!done);
    return;
  }
#endif
  ctx_copy<T, G>(g, &s, src);
}
template <class T, int G>
MM_HD void ctx_store(const Grp<G>& g, void* dst, const Scratch<T>& s) {
#if defined(__CUDA_ARCH__) && MM_TMA_IMAGE
  if constexpr (G == 32) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this lane's shared-memory writes -> visible to the copy unit
    g.sync();
    if (g.lane == 0) {
      const unsigned src = (unsigned)__cvta_generic_to_shared(&s), bytes = (unsigned)ctx_stride<T>();
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      // (the shared-memory source must stay until it has been read; the global write itself completes by the end of the grid)
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    g.sync();
    return;
  }
#endif
  ctx_copy<T, G>(g, dst, &s);
}

template <class T, int G>
MM_HDN void reset_bad_state(const Grp<G>& g, Scratch<T>& s, const ModelDev<T>& md, const StatePtrs& st, long e) {
  // Non-finite state: mj_checkPos / mj_checkVel would warn and reset the data; here the env is put back on
  // the keyframe, flagged (diag[3]) so the host can count it, and stepping continues (as mj_step does).
  if (state_bad<T, G>(g, s)) {
    for (int i = g.lane; i < NQ; i += G) s.qpos[i] = md.key_qpos[i];
    for (int i = g.lane; i < NV; i += G) { s.qvel[i] = 0; s.warm_g[i] = 0; }
    if (g.lane == 0) st.diag[e * 4 + 3] += 1;
    g.sync();
  }
}

// decode_action (gym_env.py:252-281): only the translation reaches the controller; rotation is decoded and
// dropped by the reference, so it is not computed here.
template <class T>
MM_HD void decode_action(Scratch<T>& s, const StatePtrs& st, long e, const float* a, int mode) {
  float gr = mode == MODE_ABS_POS ? a[3] : ((mode == MODE_QUAT || mode == MODE_QUAT_REL) ? a[7] : a[9]);
  if (mode == MODE_QUAT_REL || mode == MODE_ROT6D_REL) {
    const double* ti = st.tinit + e * 12;
    for (int r = 0; r < 3; r++)
      s.target[r] = (T)(ti[3 + 3 * r] * (double)a[0] + ti[3 + 3 * r + 1] * (double)a[1] + ti[3 + 3 * r + 2] * (double)a[2] + ti[r]);
  } else for (int r = 0; r < 3; r++) s.target[r] = (T)(double)a[r];
  s.ctrl[7] = gr > 0.5f ? (T)255 : (T)0;
}

// stage A on a loaded scratch: IK .. queue push of round `sub`
template <class T, int G>
MM_HDN void stage_a_body(const Grp<G>& g, Scratch<T>& s, const ModelDev<T>& md, Work<T>& w, const StatePtrs& st, long e, int sub,
                         const CvxQueue<T>& q, const HeavyList& hv) {
  if (sub != ACTION_REPEAT) ik<T, G>(g, s, md);
  reset_bad_state<T, G>(g, s, md, st, e);
  fk<T, G>(g, s, md);
  dyn_smooth<T, G>(g, s, md);
  broad_phase<T, G>(g, s, md, w);
  narrow_box<T, G>(g, s, md, w);
  // convex candidates -> batch-wide queue (the slice [qbase, qbase + ncvx) keeps them in candidate order)
  const GeomDev<T>& gm = *md.geom;
  int ncvx = list_convex<T, G>(g, s, gm, w, [](int, int) {});
  int qbase = 0;
  if (g.lane == 0 && ncvx > 0) qbase = MM_ATOMIC_ADD(q.count, ncvx);
  qbase = g.bcast(qbase, 0);
  if (qbase + ncvx > q.cap) {  // queue full (sized for 64 convex pairs per env on average): flagged, never silent
    ncvx = q.cap - qbase < 0 ? 0 : q.cap - qbase;
    if (g.lane == 0) s.overflow |= 8;
  }
  CvxItem* items = q.items + qbase;
  const int lim = ncvx, env = (int)e;
  list_convex<T, G>(g, s, gm, w, [=](int k, int ci) { if (k < lim) { items[k].env = env; items[k].ci = ci; } });
  if (g.lane == 0) {
    s.qbase = qbase; s.ncvx = ncvx;
    if (hv.flag) {
      int heavy = s.ncon + ncvx >= hv.min_load;
      hv.flag[e] = (unsigned char)heavy;
      if (heavy) { int at = MM_ATOMIC_ADD(hv.count, 1); hv.items[at] = (int)e; }
    }
  }
  g.sync();
}

// stage A of round 0 (the later rounds' stage A runs fused behind stage C of the round before, see stage_c)
template <class T, int G>
MM_HDN void stage_a(const Grp<G>& g, Scratch<T>& s, const ModelDev<T>& md, Work<T>& w, const StatePtrs& st, long e, int sub,
                    const float* action, int mode, char* ctx_base, const CvxQueue<T>& q, const HeavyList& hv) {
  void* ctx = ctx_base + (size_t)e * ctx_stride<T>();
  if (sub == 0) {
    load_state<T, G>(g, s, st, e);
    fk<T, G>(g, s, md);  // state after reset / the previous step's trailing mj_forward
    if (g.lane == 0) decode_action(s, st, e, action + e * ACTION_STRIDE, mode);
    g.sync();
  } else {
    ctx_load<T, G>(g, s, ctx);
    if (g.lane == 0) s.warm_g = st.warm + e * NV;
    g.sync();
  }
  stage_a_body<T, G>(g, s, md, w, st, e, sub, q, hv);
  ctx_store<T, G>(g, ctx, s);
}

// Stage C of round `sub`.  FUSE: unless it is the last round, stage A of round sub + 1 follows at once for the same
// env (no batch-wide barrier and no image round trip between them: `qn` / `hvn` are the NEXT round's queue and list,
// which alternate between two buffers so that this round's results stay readable for the envs still in stage C).
// Measured on the B200 the fused form is 15-25 % SLOWER (stage A then runs at stage C's register budget and
// occupancy), so the library launches the two stages separately; the switch stays for experiments (MM_FUSE_CA=1).
template <class T, int G, bool FUSE>
MM_HDN void stage_c(const Grp<G>& g, Scratch<T>& s, const ModelDev<T>& md, Work<T>& w, const StatePtrs& st, long e, int sub,
                    char* ctx_base, const CvxQueue<T>& q, const CvxQueue<T>& qn, const HeavyList& hvn, int reward_type,
                    int max_steps, const StepOut& out, const float* tgt_kp_all) {
  void* ctx = ctx_base + (size_t)e * ctx_stride<T>();
  ctx_load<T, G>(g, s, ctx);
  if (g.lane == 0) s.warm_g = st.warm + e * NV;
  g.sync();
  w.cvx = q.res + s.qbase;
  assemble_contacts<T, G>(g, s, md, w);
  {
    Work<T> wr = w;  // (a copy: the fused stage A below must see the global row arrays again)
    if (ROWS_S > 0 && s.ncon <= ROWS_S) {  // few contacts: the solver rows live in shared memory (mm_core.h, MM_ROWS_S)
      wr.Jaref = s.rows_s;
      if (MM_ROWS_N >= 2) wr.Jv = s.rows_s + 6 * ROWS_S;
      if (MM_ROWS_N == 3) wr.aref = s.rows_s + 12 * ROWS_S;
    }
    make_constraints<T, G>(g, s, md, wr);
    solve<T, G>(g, s, md, wr);
  }
  if (sub != ACTION_REPEAT) {
    integrate<T, G>(g, s, md);
    if (FUSE) stage_a_body<T, G>(g, s, md, w, st, e, sub + 1, qn, hvn);
    ctx_store<T, G>(g, ctx, s);
    return;
  }
  // last round: the trailing mj_forward is done; reward, termination, observation and the state store are the
  // epilogue kernel's (stage_finish), which keeps ~2,300 SASS instructions out of this kernel's instruction footprint
  ctx_store<T, G>(g, ctx, s);
}

// Epilogue of the step for env e (gym_env.py:562-579): reward / termination, packed observation, state store; works on
// the env's image and contact list as the last stage C left them.
template <class T, int G>
MM_HDN void stage_finish(const Grp<G>& g, Scratch<T>& s, Work<T>& w, const StatePtrs& st, long e, char* ctx_base, int reward_type,
                         int max_steps, const StepOut& out, const float* tgt_kp_all) {
  ctx_load<T, G>(g, s, ctx_base + (size_t)e * ctx_stride<T>());
  bool rc = reward_type == REWARD_STAGED ? any_robot_collision<T, G>(g, s, w) : false;
  for (int i = g.lane; i < 9; i += G) s.tmp6[KIN_ROW + i / 6][i % 6] = s.qpos[i];
  g.sync();
  store_state<T, G>(g, s, st, e);
  if (g.lane == 0) {
    write_reward<T>(s, st, e, reward_type, max_steps, rc, out);
    write_obs<T>(s, st, e, out.obs + e * OBS_DIM, tgt_kp_all + e * 4);
  }
}

// convex stage: item i of the queue by one group; body poses from the env's global image
template <class T, int G>
MM_HDN void stage_convex(const Grp<G>& g, const GeomDev<T>& gm, const CvxQueue<T>& q, int i, const char* ctx_base,
                         T (*bpos)[3], T (*bR)[9], const EpaMem<T>& em) {
  CvxItem it = q.items[i];
  const Scratch<T>* cs = reinterpret_cast<const Scratch<T>*>(ctx_base + (size_t)it.env * ctx_stride<T>());
  for (int k = g.lane; k < NDB * 3; k += G) bpos[k / 3][k % 3] = cs->bpos[k / 3][k % 3];
  for (int k = g.lane; k < NDB * 9; k += G) bR[k / 9][k % 9] = cs->bR[k / 9][k % 9];
  g.sync();
  convex_pair<T, G>(g, bpos, bR, gm, it.ci, em, q.res + i);
}

// Reset env e to the keyframe (+ optional object placement), gym_env.py:477-534.
template <class T, int G>
MM_HDN void env_reset(const Grp<G>& g, Scratch<T>& s, const ModelDev<T>& md, Work<T>& w, const StatePtrs& st, long e,
                      const double* obj_xy /*[6] or null*/, const double* yaw_cs /*[6] or null*/, int obj, int bin, float* obs,
                      float* tgt_kp_all) {
  if (g.lane == 0) s.warm_g = st.warm + e * NV;
  g.sync();
  for (int i = g.lane; i < NQ; i += G) s.qpos[i] = md.key_qpos[i];
  for (int i = g.lane; i < NV; i += G) { s.qvel[i] = 0; s.warm_g[i] = 0; }
  for (int i = g.lane; i < NU; i += G) s.ctrl[i] = md.key_ctrl[i];
  if (g.lane == 0) { s.overflow = 0; s.ncon = 0; s.npair = 0; s.nspec = 0; s.niter = 0; s.prof = 0; }
  g.sync();
  for (int pass = 0; pass < (obj_xy ? 2 : 1); pass++) {
    if (pass == 1) {  // randomization.py:52-65 + env.py:160-161
      if (g.lane == 0)
        for (int o = 0; o < 3; o++) {
          T* q = s.qpos + 9 + 7 * o;
          q[0] = (T)obj_xy[2 * o]; q[1] = (T)obj_xy[2 * o + 1]; q[2] = (T)0.26; q[3] = 1; q[4] = q[5] = q[6] = 0;
          if (yaw_cs) { q[3] = (T)yaw_cs[2 * o]; q[6] = (T)yaw_cs[2 * o + 1]; }  // randomize_yaw, randomization.py:55-62
        }
      g.sync();
    }
    forward<T, G>(g, s, md, w);  // env.py:116-117
  }
  for (int i = g.lane; i < 9; i += G) s.tmp6[KIN_ROW + i / 6][i % 6] = s.qpos[i];
  g.sync();
  store_state<T, G>(g, s, st, e);
  if (g.lane == 0) {
    for (int i = 0; i < 12; i++) st.tinit[e * 12 + i] = (double)(i < 3 ? s.bpos[DB_HAND][i] : s.bR[DB_HAND][i - 3]);
    st.step_count[e] = 0;
    st.task[e * 2] = obj; st.task[e * 2 + 1] = bin;
    st.flags[e] = 0;
    for (int k = 0; k < 5; k++) st.hwm[e * 5 + k] = 0;
    int* fi = st.fsm_i + e * 5;
    fi[0] = 1; fi[1] = 0; fi[2] = 0; fi[3] = 1; fi[4] = 0;
    int* ft = st.fsm_tasks + e * FSM_TASKS_STRIDE;  // the expert of this episode works on the env's own task
    ft[0] = 1; ft[1] = obj; ft[2] = bin;
    for (int k = 0; k < 6; k++) st.fsm_f[e * 6 + k] = 0;
    st.diag[e * 4 + 2] = 0;
    // target keypoints frozen at reset (gym_env.py:519-531)
    const double ocp[3] = {0, 0, 2.0}, ocm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double p[3];
    for (int a = 0; a < 3; a++) p[a] = (double)s.bpos[DB_CUBE0 + obj][a];
    project_kp(ocp, ocm, F_OVERHEAD, p, tgt_kp_all + e * 4);
    bin_pos(bin, p);
    project_kp(ocp, ocm, F_OVERHEAD, p, tgt_kp_all + e * 4 + 2);
  }
  g.sync();
  if (g.lane == 0 && obs) write_obs<T>(s, st, e, obs + e * OBS_DIM, tgt_kp_all + e * 4);
}

enum { OP_IK = 1, OP_FORWARD = 2, OP_INTEGRATE = 4 };

// Engine-level calls on env e: IKController.compute + set_arm_ctrl (controller.py:87-137 with the stale
// kinematics of the last position stage), mj_forward, mj_step = forward + integrate (env.py:117-121).
template <class T, int G>
MM_HDN void env_ops(const Grp<G>& g, Scratch<T>& s, const ModelDev<T>& md, Work<T>& w, const StatePtrs& st, long e,
                    int ops, const double* target) {
  load_state<T, G>(g, s, st, e);
  // kinematics as the last position stage left them
  for (int i = g.lane; i < 9; i += G) { s.fs[i] = s.qpos[i]; s.qpos[i] = (T)st.kin[e * 18 + i]; }
  g.sync();
  fk<T, G>(g, s, md);
  for (int i = g.lane; i < 9; i += G) s.qpos[i] = s.fs[i];
  if (g.lane == 0 && target) for (int k = 0; k < 3; k++) s.target[k] = (T)target[3 * e + k];
  g.sync();
  if ((ops & OP_IK) && target) ik<T, G>(g, s, md);
  if (ops & OP_FORWARD) {
    if (state_bad<T, G>(g, s)) {
      for (int i = g.lane; i < NQ; i += G) s.qpos[i] = md.key_qpos[i];
      for (int i = g.lane; i < NV; i += G) { s.qvel[i] = 0; s.warm_g[i] = 0; }
      if (g.lane == 0) st.diag[e * 4 + 3] += 1;
      g.sync();
    }
    forward<T, G>(g, s, md, w);
    for (int i = g.lane; i < 9; i += G) s.tmp6[KIN_ROW + i / 6][i % 6] = s.qpos[i];  // position-stage qpos for store_state
    g.sync();
    if (ops & OP_INTEGRATE) integrate<T, G>(g, s, md);
  } else {
    for (int i = g.lane; i < 9; i += G) s.tmp6[KIN_ROW + i / 6][i % 6] = (T)st.kin[e * 18 + i];
    g.sync();
  }
  store_state<T, G>(g, s, st, e);
}

// Scripted FSM, one plan(n_steps) call (pick_and_place.py:167-277) + the abs_pos action it implies
// (scripts/generate_dataset.py:145-148).  Thread per env; reads the cached EE pose and cube positions.
MM_HDN inline void fsm_plan_one(const StatePtrs& st, long e, int n, float* action_out /*[4] or null*/) {
  int* fi = st.fsm_i + e * 5;
  double* tg = st.fsm_f + e * 6;
  double* te = tg + 3;
  const double* ee = st.eepose + e * 12;
  // the FSM's own task list (pick_and_place.py:91,151-165); task_index >= count only while IDLE -> DONE
  const int* ft = st.fsm_tasks + e * FSM_TASKS_STRIDE;
  int ntask = ft[0] < FSM_MAX_TASKS ? ft[0] : FSM_MAX_TASKS;
  int ti = fi[1] < ntask ? fi[1] : (ntask > 0 ? ntask - 1 : 0);
  int obj = ntask > 0 ? ft[1 + 2 * ti] : 0, bin = ntask > 0 ? ft[2 + 2 * ti] : 0;
  const double* op = st.kin + e * 18 + 9 + 3 * obj;  // data.xpos of the object at the last position stage
  double bp[3];
  bin_pos(bin, bp);
  double dd = sqrt((ee[0] - tg[0]) * (ee[0] - tg[0]) + (ee[1] - tg[1]) * (ee[1] - tg[1]) + (ee[2] - tg[2]) * (ee[2] - tg[2]));
  bool reached = dd < 0.02;
  switch (fi[0]) {
    case 1:
      if (fi[1] >= ntask) { fi[0] = 11; break; }
      fi[3] = 1; tg[0] = op[0]; tg[1] = op[1]; tg[2] = 0.44; fi[4] = 1; fi[0] = 2;
      break;
    case 2:
      if (reached) { tg[0] = op[0]; tg[1] = op[1]; tg[2] = 0.36; fi[0] = 3; }
      break;
    case 3:
      if (reached) { fi[3] = 0; fi[2] = 150; fi[0] = 4; }
      break;
    case 4:
      fi[2] -= n;
      if (fi[2] <= 0) { tg[0] = op[0]; tg[1] = op[1]; tg[2] = 0.55; fi[0] = 5; }
      break;
    case 5:
      if (reached) { te[0] = bp[0]; te[1] = bp[1]; te[2] = 0.55; fi[0] = 6; }
      break;
    case 6: {
      double df[3] = {te[0] - tg[0], te[1] - tg[1], te[2] - tg[2]};
      double dist = sqrt(df[0] * df[0] + df[1] * df[1] + df[2] * df[2]);
      double stp = 0.001 * n;
      if (dist > stp) for (int k = 0; k < 3; k++) tg[k] += df[k] * (stp / dist);
      else for (int k = 0; k < 3; k++) tg[k] = te[k];
      if (dist <= 0.02) { fi[2] = 100; fi[0] = 7; }
      break;
    }
    case 7:
      fi[2] -= n;
      if (fi[2] <= 0) { tg[0] = bp[0]; tg[1] = bp[1]; tg[2] = 0.45; fi[0] = 8; }
      break;
    case 8:
      if (reached) { fi[3] = 1; fi[2] = 150; fi[0] = 9; }
      break;
    case 9:
      fi[2] -= n;
      if (fi[2] <= 0) { tg[0] = 0.0; tg[1] = 0.3; tg[2] = 0.55; fi[0] = 10; }
      break;
    case 10:
      if (reached) { fi[1] += 1; fi[0] = 1; }
      break;
    default:
      break;
  }
  if (action_out) {
    const double* src = fi[4] ? tg : ee;
    for (int k = 0; k < 3; k++) action_out[k] = (float)src[k];
    action_out[3] = fi[3] ? 1.0f : 0.0f;
  }
}

}  // namespace mm
