// ORACLE - TEST INFRASTRUCTURE ONLY (see engine.h).  C ABI over the oracle for ctypes:
// used by oracle/fake_mujoco.py (the stand-in `mujoco` module the reference's own Python runs on),
// by tests/ as the parity checker and by bench.py's cpu_baseline / --impl reference legs.
#define MM_WANT_NAMES
#include <chrono>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "hotpath.h"
#include "ccd.h"

using namespace orc;

extern "C" {

void* orc_new() {
  Env* e = new Env();
  std::memset(&e->d.stats, 0, sizeof(Stats));
  e->d.flags = 0;
  e->d.ncon = e->d.nefc = 0;
  reset_keyframe(e->d);
  forward(e->d);
  return e;
}
void orc_free(void* h) { delete (Env*)h; }

void orc_config(void* h, int action_mode, int reward_type, int max_steps, int flags) {
  Env* e = (Env*)h;
  e->action_mode = action_mode; e->reward_type = reward_type; e->max_episode_steps = max_steps; e->d.flags = flags;
}

double* orc_ptr(void* h, const char* name) {
  Data& d = ((Env*)h)->d;
  if (!strcmp(name, "qpos")) return d.qpos;
  if (!strcmp(name, "qvel")) return d.qvel;
  if (!strcmp(name, "ctrl")) return d.ctrl;
  if (!strcmp(name, "qacc")) return d.qacc;
  if (!strcmp(name, "qacc_warmstart")) return d.qacc_warmstart;
  if (!strcmp(name, "qacc_smooth")) return d.qacc_smooth;
  if (!strcmp(name, "qfrc_bias")) return d.qfrc_bias;
  if (!strcmp(name, "qfrc_smooth")) return d.qfrc_smooth;
  if (!strcmp(name, "qfrc_constraint")) return d.qfrc_constraint;
  if (!strcmp(name, "qfrc_actuator")) return d.qfrc_actuator;
  if (!strcmp(name, "xpos")) return &d.xpos[0][0];
  if (!strcmp(name, "xmat")) return &d.xmat[0][0];
  if (!strcmp(name, "xquat")) return &d.xquat[0][0];
  if (!strcmp(name, "xipos")) return &d.xipos[0][0];
  if (!strcmp(name, "geom_xpos")) return &d.geom_xpos[0][0];
  if (!strcmp(name, "geom_xmat")) return &d.geom_xmat[0][0];
  if (!strcmp(name, "cam_xpos")) return &d.cam_xpos[0][0];
  if (!strcmp(name, "cam_xmat")) return &d.cam_xmat[0][0];
  if (!strcmp(name, "M")) return d.M;
  if (!strcmp(name, "time")) return &d.time;
  if (!strcmp(name, "init_pos")) return ((Env*)h)->init_pos;
  if (!strcmp(name, "init_R")) return ((Env*)h)->init_R;
  return nullptr;
}

// constraint row i: out = {pos, D, R, aref, force}; returns the row type (0 equality, 1 limit, 2 contact) or -1
int orc_efc(void* h, int i, double* out) {
  Data& d = ((Env*)h)->d;
  if (i < 0 || i >= d.nefc) return -1;
  out[0] = d.efc_pos[i]; out[1] = d.efc_D[i]; out[2] = d.efc_R[i]; out[3] = d.efc_aref[i]; out[4] = d.efc_force[i];
  return d.efc_type[i];
}

int orc_ncon(void* h) { return ((Env*)h)->d.ncon; }
int orc_nefc(void* h) { return ((Env*)h)->d.nefc; }
int orc_niter(void* h) { return ((Env*)h)->d.solver_niter; }
void orc_contact(void* h, int i, int* g1, int* g2, double* dist, double* pos, double* frame) {
  const Contact& c = ((Env*)h)->d.contact[i];
  *g1 = c.geom1; *g2 = c.geom2; *dist = c.dist;
  if (pos) std::memcpy(pos, c.pos, sizeof c.pos);
  if (frame) std::memcpy(frame, c.frame, sizeof c.frame);
}

void orc_mj_step(void* h) { step(((Env*)h)->d); }
void orc_mj_forward(void* h) { forward(((Env*)h)->d); }
void orc_mj_kinematics(void* h) { kinematics_only(((Env*)h)->d); }
void orc_mj_reset_keyframe(void* h) { reset_keyframe(((Env*)h)->d); }
void orc_mj_jac(void* h, double* jp, double* jr, const double* point, int body) { jac(((Env*)h)->d, jp, jr, point, body); }

void orc_env_reset(void* h, const double* obj_xy, int obj_idx, int bin_idx) { env_reset(*(Env*)h, obj_xy, obj_idx, bin_idx); }
void orc_env_reset_yaw(void* h, const double* obj_xy, const double* yaw_cs, int obj_idx, int bin_idx) {
  env_reset(*(Env*)h, obj_xy, obj_idx, bin_idx, yaw_cs);
}
void orc_env_step(void* h, const float* action, float* obs, double* reward, int* term, int* trunc, int* succ, float* rc) {
  env_step(*(Env*)h, action, obs, reward, term, trunc, succ, rc);
}
void orc_env_obs(void* h, float* obs) { env_obs(*(Env*)h, obs); }
void orc_ik(void* h, const double* target, double* q) { ik_compute(((Env*)h)->d, target, q); }
void orc_decode(void* h, const float* action, double* target, float* g) { decode_action(*(Env*)h, action, target, g); }
void orc_fsm_reset(void* h) { fsm_reset(*(Env*)h); }
void orc_fsm_plan(void* h, int n) { fsm_plan(*(Env*)h, n); }
void orc_fsm_action(void* h, float* a) { fsm_action(*(Env*)h, a); }
void orc_fsm_get(void* h, int* st, double* target, double* transit_end) {
  Env* e = (Env*)h;
  st[0] = e->fsm_state; st[1] = e->task_index; st[2] = e->settle_counter; st[3] = e->gripper_open; st[4] = e->has_target;
  std::memcpy(target, e->target, sizeof e->target);
  std::memcpy(transit_end, e->transit_end, sizeof e->transit_end);
}
void orc_stats(void* h, long long* out) {
  const Stats& s = ((Env*)h)->d.stats;
  out[0] = s.substeps; out[1] = s.ncon; out[2] = s.nefc; out[3] = s.newton_iters; out[4] = s.ls_evals;
  out[5] = s.narrow_tests; out[6] = s.ccd_tests; out[7] = s.max_ncon; out[8] = s.max_nefc; out[9] = s.max_newton;
  out[10] = s.forwards; out[11] = s.nnzJ; out[12] = s.nnzJ2;
}
void orc_stats_clear(void* h) { std::memset(&((Env*)h)->d.stats, 0, sizeof(Stats)); }

// model tables for the fake module
int orc_model_int(const char* name) {
  if (!strcmp(name, "nq")) return NQ;
  if (!strcmp(name, "nv")) return NV;
  if (!strcmp(name, "nu")) return NU;
  if (!strcmp(name, "nbody")) return NBODY;
  if (!strcmp(name, "ngeom")) return NGEOM;
  if (!strcmp(name, "njnt")) return NJNT;
  return -1;
}
const char* orc_body_name(int i) { return mm_body_name[i]; }
const char* orc_geom_name(int i) { return mm_geom_name[i]; }
const char* orc_jnt_name(int i) { return mm_jnt_name[i]; }
int orc_geom_bodyid(int i) { return mm_geom_body[i]; }
int orc_jnt_qposadr(int i) { return mm_jnt_qposadr[i]; }
void orc_jnt_range(int i, double* r) { r[0] = mm_jnt_range[i][0]; r[1] = mm_jnt_range[i][1]; }

// One scripted-FSM expert episode (scripts/generate_dataset.py:140-196 loop): returns success of the
// final state; *length = number of env steps taken.
int orc_run_fsm_episode(void* h, const double* obj_xy, int obj_idx, int bin_idx, int max_steps, int* length,
                        int* phase_hist /*12 or null*/) {
  Env* e = (Env*)h;
  int keep_mode = e->action_mode;
  e->action_mode = ABS_POS;
  env_reset(*e, obj_xy, obj_idx, bin_idx);
  int n = 0, succ = 0, term, trunc;
  double r;
  while (e->fsm_state != 11 && n < max_steps) {
    fsm_plan(*e, 16);
    float a[4];
    fsm_action(*e, a);
    if (phase_hist) phase_hist[e->fsm_state]++;
    env_step(*e, a, nullptr, &r, &term, &trunc, &succ, nullptr);
    n++;
  }
  *length = n;
  e->action_mode = keep_mode;
  return succ;
}

// CPU baseline: `n_envs` independent envs, `n_steps` env-steps each, random world-frame targets
// (SURVEY 8d config 2 distribution) in the given action mode, spread over `nthreads` host threads.
// Returns env-steps per second.
double orc_bench_random2(int n_envs, int n_steps, int mode, unsigned seed, int nthreads, int flags, long long* stats_out /*13 or null*/) {
  std::vector<Env*> envs(n_envs);
  for (int i = 0; i < n_envs; i++) {
    envs[i] = (Env*)orc_new();
    envs[i]->action_mode = mode;
    envs[i]->d.flags = flags;
    env_reset(*envs[i], nullptr, 0, 0);
  }
  auto work = [&](int t) {
    for (int i = t; i < n_envs; i += nthreads) {
      std::mt19937 rng(seed + 7919u * i);
      std::uniform_real_distribution<float> ux(-0.3f, 0.3f), uy(0.30f, 0.65f), uz(0.30f, 0.60f), u01(0.f, 1.f);
      Env& e = *envs[i];
      float obs[OBS_FULL_DIM];
      for (int s = 0; s < n_steps; s++) {
        double w[3] = {ux(rng), uy(rng), uz(rng)};
        float a[10] = {0};
        int gi = mode == ABS_POS ? 3 : ((mode == EE_POS_QUAT_G || mode == EE_POS_QUAT_G_REL) ? 7 : 9);
        if (mode == EE_POS_QUAT_G_REL || mode == EE_POS_ROT6D_G_REL) {
          for (int k = 0; k < 3; k++) {
            double dp[3] = {w[0] - e.init_pos[0], w[1] - e.init_pos[1], w[2] - e.init_pos[2]};
            a[k] = (float)(e.init_R[k] * dp[0] + e.init_R[3 + k] * dp[1] + e.init_R[6 + k] * dp[2]);
          }
        } else for (int k = 0; k < 3; k++) a[k] = (float)w[k];
        if (gi == 7) a[6] = 1.f; else if (gi == 9) { a[3] = 1.f; a[7] = 1.f; }
        a[gi] = u01(rng) > 0.5f ? 1.f : 0.f;
        double r; int te, tr, su;
        env_step(e, a, obs, &r, &te, &tr, &su, nullptr);
      }
    }
  };
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) th.emplace_back(work, t);
  for (auto& x : th) x.join();
  double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (stats_out) {
    for (int k = 0; k < 13; k++) stats_out[k] = 0;
    for (auto e : envs) {
      long long o[13];
      orc_stats(e, o);
      for (int k = 0; k < 13; k++) {
        if (k >= 7 && k <= 9) stats_out[k] = std::max(stats_out[k], o[k]);
        else stats_out[k] += o[k];
      }
    }
  }
  for (auto e : envs) delete e;
  return (double)n_envs * n_steps / dt;
}

void orc_ccd_marks(long long* out) {
  out[0] = ccd::marks().max_faces; out[1] = ccd::marks().max_iters; out[2] = ccd::marks().calls; out[3] = ccd::marks().iters;
  out[4] = ccd::marks().gjk_calls; out[5] = ccd::marks().gjk_iters; out[6] = ccd::marks().gjk_hits; out[7] = ccd::marks().gjk_maxed;
}

double orc_bench_random(int n_envs, int n_steps, int mode, unsigned seed, int nthreads, int flags) {
  return orc_bench_random2(n_envs, n_steps, mode, seed, nthreads, flags, nullptr);
}

}  // extern "C"
