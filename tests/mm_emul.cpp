// TEST-ONLY host build of the kernel source with a 1-lane group (G = 1).
// Lives in tests/ and is built by tests/hostlib.py (g++) into tests/_build/libmm_emul.so so that the kernel LOGIC can be checked
// against the oracle on machines without a GPU (`-m "not gpu"` tests).  It is never loaded by the
// mujoco_manip_b200 package: the product path is the CUDA library and fails loudly without it.
#define MM_MODEL_HOST_FILL
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <type_traits>

// Operation-counting scalar: the kernel source instantiated with it executes the very same algorithm and counts every
// floating-point add / subtract / multiply / divide / square root / trigonometric call it performs (1 each) - the
// ALGORITHMIC work of a step at G = 1, free of lane redundancy - per stage (bench.py roofline).
struct Cnt {
  double v;
  static long long n;
  Cnt() = default;
  template <class U, class = typename std::enable_if<std::is_arithmetic<U>::value>::type>
  Cnt(U x) : v((double)x) {}
  explicit operator double() const { return v; }
  explicit operator float() const { return (float)v; }
  explicit operator int() const { return (int)v; }
  Cnt& operator+=(Cnt o) { v += o.v; n++; return *this; }
  Cnt& operator-=(Cnt o) { v -= o.v; n++; return *this; }
  Cnt& operator*=(Cnt o) { v *= o.v; n++; return *this; }
  Cnt& operator/=(Cnt o) { v /= o.v; n++; return *this; }
};
long long Cnt::n = 0;
inline Cnt operator+(Cnt a, Cnt b) { Cnt::n++; return Cnt(a.v + b.v); }
inline Cnt operator-(Cnt a, Cnt b) { Cnt::n++; return Cnt(a.v - b.v); }
inline Cnt operator*(Cnt a, Cnt b) { Cnt::n++; return Cnt(a.v * b.v); }
inline Cnt operator/(Cnt a, Cnt b) { Cnt::n++; return Cnt(a.v / b.v); }
inline Cnt operator-(Cnt a) { return Cnt(-a.v); }
inline bool operator<(Cnt a, Cnt b) { return a.v < b.v; }
inline bool operator>(Cnt a, Cnt b) { return a.v > b.v; }
inline bool operator<=(Cnt a, Cnt b) { return a.v <= b.v; }
inline bool operator>=(Cnt a, Cnt b) { return a.v >= b.v; }
inline bool operator==(Cnt a, Cnt b) { return a.v == b.v; }
inline bool operator!=(Cnt a, Cnt b) { return a.v != b.v; }
namespace mm {
inline Cnt tsqrt(Cnt x) { Cnt::n++; return Cnt(std::sqrt(x.v)); }
inline Cnt trsqrt(Cnt x) { Cnt::n += 2; return Cnt(1.0 / std::sqrt(x.v)); }
inline void tsincos(Cnt x, Cnt* s, Cnt* c) { Cnt::n += 2; s->v = std::sin(x.v); c->v = std::cos(x.v); }
inline Cnt tacos(Cnt x) { Cnt::n++; return Cnt(std::acos(x.v)); }
inline Cnt tsin(Cnt x) { Cnt::n++; return Cnt(std::sin(x.v)); }
}  // namespace mm

#include "../mujoco_manip_b200/csrc/mm_env.h"

using namespace mm;

namespace {
template <class T>
struct Ctx {
  ModelDev<T> md;
  GeomDev<T> gm;
  Scratch<T> s;
  std::vector<T> wr;
  std::vector<int> wi;
  Work<T> w;
  Ctx() : wr(WORK_REALS), wi(WORK_INTS) {
    fill_model(md);
    fill_geom(gm);
    md.geom = &gm;
    w = make_work(wr.data(), wi.data());
    std::memset(&s, 0, sizeof s);
  }
};
template <class T>
Ctx<T>& ctx() {
  static thread_local Ctx<T> c;
  return c;
}

StatePtrs state_from(void** p) {
  StatePtrs st;
  st.qpos = (double*)p[0]; st.qvel = (double*)p[1]; st.ctrl = (double*)p[2]; st.warm = (double*)p[3];
  st.tinit = (double*)p[4]; st.eepose = (double*)p[5]; st.fsm_f = (double*)p[6]; st.hwm = (double*)p[7]; st.kin = (double*)p[8];
  st.step_count = (int*)p[9]; st.task = (int*)p[10]; st.fsm_i = (int*)p[11]; st.fsm_tasks = (int*)p[12]; st.flags = (int*)p[13]; st.diag = (int*)p[14];
  return st;
}
StepOut out_from(void** p) {
  StepOut o;
  o.obs = (float*)p[0]; o.reward = (float*)p[1]; o.terminated = (unsigned char*)p[2]; o.truncated = (unsigned char*)p[3];
  o.success = (unsigned char*)p[4]; o.reward_components = (float*)p[5];
  return o;
}

template <class T>
void do_reset(int n, void** sp, const unsigned char* mask, const double* obj_xy, const double* yaw_cs, const int* task,
              float* obs, float* tgt) {
  Ctx<T>& c = ctx<T>();
  StatePtrs st = state_from(sp);
  Grp<1> g{0, 1u};
  for (long e = 0; e < n; e++) {
    if (mask && !mask[e]) continue;
    env_reset<T, 1>(g, c.s, c.md, c.w, st, e, obj_xy ? obj_xy + 6 * e : nullptr, yaw_cs ? yaw_cs + 6 * e : nullptr, task[2 * e],
                    task[2 * e + 1], obs, tgt);
  }
}
// One control step of n envs through the STAGE functions the CUDA kernels run (mm_env.h): 17 rounds of stage A for
// every env, the batch-wide convex queue, stage C for every env - with per-env contact lists and images as on the device.
// convex-queue statistics of everything stepped so far: items, hits, env-rounds; per geom: items, hits
static long long g_qstat[3], g_qpair[NGEOM][2];
template <class T>
void do_step(int n, void** sp, const float* actions, int mode, int reward_type, int max_steps, void** op, const float* tgt,
             long long* flops = nullptr /*[3]: stage A, convex stage, stage C*/) {
  long long mark = Cnt::n;
  auto tick = [&](int k) { if (flops) flops[k] += Cnt::n - mark; mark = Cnt::n; };
  Ctx<T>& c = ctx<T>();
  StatePtrs st = state_from(sp);
  StepOut out = out_from(op);
  Grp<1> g{0, 1u};
  std::vector<T> wer((size_t)n * WORK_REALS);
  std::vector<int> wei((size_t)n * WORK_INTS);
  std::vector<char> image((size_t)n * ctx_stride<T>());
  const int cap = n * 64;
  std::vector<CvxItem> items(2 * cap);   // two buffers: round r's results are read while round r + 1's items are pushed
  std::vector<CvxRes<T>> res(2 * cap);
  std::vector<T> everts(EPA_MAXV * 6), eface(EPA_MAXF * 4);
  std::vector<int> eints(EPA_INTS);
  static T bpos[NDB][3], bR[NDB][9];
  int count[ACTION_REPEAT + 2] = {0}, head[ACTION_REPEAT + 2] = {0};
  const HeavyList nohv{nullptr, nullptr, nullptr, 0};
  auto queue = [&](int sub) { return CvxQueue<T>{items.data() + (sub & 1) * cap, res.data() + (sub & 1) * cap, &count[sub], &head[sub], cap}; };
  auto work = [&](long e) { return make_work(wer.data() + e * WORK_REALS, wei.data() + e * WORK_INTS); };
  for (long e = 0; e < n; e++) {  // stage A of round 0
    Work<T> w = work(e);
    // stage A owns only the first scratch_a_bytes of the scratch on the device: everything behind must stay untouched
    unsigned char* tail = reinterpret_cast<unsigned char*>(&c.s) + scratch_a_bytes<T>();
    size_t ntail = sizeof(Scratch<T>) - scratch_a_bytes<T>();
    std::memset(tail, 0xA5, ntail);
    stage_a<T, 1>(g, c.s, c.md, w, st, e, 0, actions, mode, image.data(), queue(0), nohv);
    for (size_t k = 0; k < ntail; k++)
      if (tail[k] != 0xA5) { std::fprintf(stderr, "stage A wrote outside its scratch slice (byte %zu)\n", scratch_a_bytes<T>() + k); std::abort(); }
  }
  tick(0);
  for (int sub = 0; sub <= ACTION_REPEAT; sub++) {
    CvxQueue<T> q = queue(sub), qn = queue(sub + 1);
    EpaMem<T> em;
    em.vert = everts.data(); em.face = eface.data(); em.fidx = eints.data(); em.edge = em.fidx + EPA_MAXF; em.canon = em.edge + EPA_MAXE; em.ecan = em.canon + EPA_MAXV;
    int cnt = count[sub] < cap ? count[sub] : cap;
    g_qstat[0] += cnt; g_qstat[2] += n;
    for (int i = cnt - 1; i >= 0; i--) {  // any order: results are addressed by queue position
#ifdef MM_DEBUG_EPA
      long before = mm_debug_epa_iters;
      stage_convex<T, 1>(g, c.gm, q, i, image.data(), bpos, bR, em);
      long it = mm_debug_epa_iters - before;
      if (it >= 40) {
        int ci = q.items[i].ci, a = c.gm.pair[ci][0], b = c.gm.pair[ci][1];
        std::fprintf(stderr, "EPA %ld iterations: env %d pair %d (geom %d type %d body %d | geom %d type %d body %d) hit %d depth %g\n", it,
                     q.items[i].env, ci, a, c.gm.type[a], c.gm.body[a], b, c.gm.type[b], c.gm.body[b], q.res[i].hit, (double)q.res[i].depth);
      }
#else
      stage_convex<T, 1>(g, c.gm, q, i, image.data(), bpos, bR, em);
#endif
      {
        int ci = q.items[i].ci;
        g_qstat[1] += q.res[i].hit;
        g_qpair[c.gm.pair[ci][0]][0]++; g_qpair[c.gm.pair[ci][0]][1] += q.res[i].hit;
        g_qpair[c.gm.pair[ci][1]][0]++; g_qpair[c.gm.pair[ci][1]][1] += q.res[i].hit;
      }
    }
    tick(1);
    for (long e = 0; e < n; e++) {  // stage C of round sub
      Work<T> w = work(e);
      stage_c<T, 1, false>(g, c.s, c.md, w, st, e, sub, image.data(), q, qn, nohv, reward_type, max_steps, out, tgt);
    }
    tick(2);
    if (sub == ACTION_REPEAT)
      for (long e = 0; e < n; e++) {  // epilogue kernel
        Work<T> w = work(e);
        stage_finish<T, 1>(g, c.s, w, st, e, image.data(), reward_type, max_steps, out, tgt);
      }
    if (sub < ACTION_REPEAT)
      for (long e = 0; e < n; e++) {  // stage A of round sub + 1 (its own launch on the device)
        Work<T> w = work(e);
        stage_a<T, 1>(g, c.s, c.md, w, st, e, sub + 1, actions, mode, image.data(), qn, nohv);
      }
    tick(0);
  }
}
}  // namespace

extern "C" {
void emul_queue_stats(long long* tot /*[3]*/, long long* per_geom /*[NGEOM][2]*/) {
  for (int k = 0; k < 3; k++) tot[k] = g_qstat[k];
  for (int k = 0; k < NGEOM; k++) { per_geom[2 * k] = g_qpair[k][0]; per_geom[2 * k + 1] = g_qpair[k][1]; }
}

void emul_reset_yaw(int n, void** state, const unsigned char* mask, const double* obj_xy, const double* yaw_cs,
                    const int* task, float* obs, float* tgt_kp, int use_float) {
  if (use_float) do_reset<float>(n, state, mask, obj_xy, yaw_cs, task, obs, tgt_kp);
  else do_reset<double>(n, state, mask, obj_xy, yaw_cs, task, obs, tgt_kp);
}

void emul_reset(int n, void** state, const unsigned char* mask, const double* obj_xy, const int* task, float* obs,
                float* tgt_kp, int use_float) {
  emul_reset_yaw(n, state, mask, obj_xy, nullptr, task, obs, tgt_kp, use_float);
}

void emul_step(int n, void** state, const float* actions, int mode, int reward_type, int max_steps, void** out,
               const float* tgt_kp, int use_float) {
  if (use_float) do_step<float>(n, state, actions, mode, reward_type, max_steps, out, tgt_kp);
  else do_step<double>(n, state, actions, mode, reward_type, max_steps, out, tgt_kp);
}

// One control step executed with the operation-counting scalar: flops[0..2] += floating-point operations of stage A,
// the convex stage and stage C over the n envs (state advances exactly as in emul_step with use_float = 0).
void emul_step_counted(int n, void** state, const float* actions, int mode, int reward_type, int max_steps, void** out,
                       const float* tgt_kp, long long* flops) {
  do_step<Cnt>(n, state, actions, mode, reward_type, max_steps, out, tgt_kp, flops);
}

void emul_ops(int n, void** state, int ops, const double* target, int use_float) {
  StatePtrs st = state_from(state);
  Grp<1> g{0, 1u};
  for (long e = 0; e < n; e++) {
    if (use_float) { Ctx<float>& c = ctx<float>(); env_ops<float, 1>(g, c.s, c.md, c.w, st, e, ops, target); }
    else { Ctx<double>& c = ctx<double>(); env_ops<double, 1>(g, c.s, c.md, c.w, st, e, ops, target); }
  }
}

void emul_fsm_plan(int n, void** state, int nsteps, float* actions /*[n,10]*/) {
  StatePtrs st = state_from(state);
  for (long e = 0; e < n; e++) fsm_plan_one(st, e, nsteps, actions ? actions + e * ACTION_STRIDE : nullptr);
}

// One forward pass on a raw state with every intermediate exposed (FP64), for stage-wise parity tests.
void emul_forward_debug(const double* qpos, const double* qvel, const double* ctrl, const double* warm, double* Mr,
                        double* fs, double* as, double* qacc, double* fc, double* bpos, double* bR, int* ncon,
                        double* cpos, double* cn, double* cdist, int* cmeta, int* niter, double* cD, double* aref) {
  Ctx<double>& c = ctx<double>();
  Scratch<double>& s = c.s;
  Grp<1> g{0, 1u};
  for (int i = 0; i < NQ; i++) s.qpos[i] = qpos[i];
  static thread_local double warm_buf[NV];
  for (int i = 0; i < NV; i++) { s.qvel[i] = qvel[i]; warm_buf[i] = warm[i]; }
  s.warm_g = warm_buf;
  for (int i = 0; i < NU; i++) s.ctrl[i] = ctrl[i];
  s.overflow = 0;
  forward<double, 1>(g, s, c.md, c.w);
  std::memcpy(Mr, s.Mr, sizeof s.Mr);
  std::memcpy(fs, s.fs, sizeof s.fs);
  std::memcpy(as, s.as, sizeof s.as);
  std::memcpy(qacc, s.qacc, sizeof s.qacc);
  std::memcpy(fc, s.fc, sizeof s.fc);
  std::memcpy(bpos, s.bpos, sizeof s.bpos);
  std::memcpy(bR, s.bR, sizeof s.bR);
  *ncon = s.ncon;
  *niter = s.niter;
  for (int k = 0; k < s.ncon; k++) {
    for (int d = 0; d < 3; d++) { cpos[3 * k + d] = c.w.cpos[d * MAXCON + k]; cn[3 * k + d] = c.w.cn[d * MAXCON + k]; }
    cdist[k] = c.w.cdist[k];
    cmeta[k] = c.w.cmeta[k];
    if (cD) cD[k] = c.w.cD[k];
    if (aref) for (int r = 0; r < 6; r++) aref[6 * k + r] = c.w.aref[6 * k + r];
  }
}

int emul_scratch_bytes(int use_float) { return use_float ? (int)sizeof(Scratch<float>) : (int)sizeof(Scratch<double>); }
}
