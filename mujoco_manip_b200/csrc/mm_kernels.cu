// CUDA kernels (sm_100a) + the C ABI of include/mm_manip.h.
// A control step is a sequence of batch-wide stage kernels (mm_launch.cuh); inside a stage one G-lane group works on
// one environment (G = 32 warp per env; 16 / 8 = two / four envs per warp) with its matrices in shared memory, and one
// warp on one convex geom pair.  No CPU fallback: every entry point needs a CUDA device.
#define MM_MODEL_HOST_FILL
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mm_manip.h"
#include "mm_launch.cuh"
#include "mm_rng.h"

using namespace mm;

namespace {

thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return -1; }
#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_)); \
  } while (0)

__global__ void k_fsm(StatePtrs st, long n, int nsteps, float* actions) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) fsm_plan_one(st, e, nsteps, actions ? actions + e * ACTION_STRIDE : nullptr);
}

// The four action encodings of an expert abs_pos action (scripts/generate_dataset.py:56-80): pose = (target
// position, TARGET_ORI) in the world frame and relative to the episode's initial EE pose.  Thread per env.
// out [N,36] = pos_quat_g (8) | pos_rot6d_g (10) | pos_quat_g_rel (8) | pos_rot6d_g_rel (10)
__global__ void k_expert(StatePtrs st, long n, const float* abs_actions, float* out) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const float* a = abs_actions + e * ACTION_STRIDE;
  const double R[9] = {0, 1, 0, 1, 0, 0, 0, 0, -1};  // TARGET_ORI, controller.py:12-18
  double p[3] = {(double)a[0], (double)a[1], (double)a[2]};
  float* o = out + e * 36;
  encode_pose(p, R, a[3], o, o + 8);
  const double* ti = st.tinit + e * 12;
  const double* Ri = ti + 3;
  double Rr[9], pr[3], dp[3] = {p[0] - ti[0], p[1] - ti[1], p[2] - ti[2]};
  for (int i = 0; i < 3; i++) {
    pr[i] = Ri[i] * dp[0] + Ri[3 + i] * dp[1] + Ri[6 + i] * dp[2];
    for (int j = 0; j < 3; j++) Rr[3 * i + j] = Ri[i] * R[j] + Ri[3 + i] * R[3 + j] + Ri[6 + i] * R[6 + j];
  }
  encode_pose(pr, Rr, a[3], o + 18, o + 26);
}

// Philox placement + task draw for every env (mm_rng.h); thread per env
__global__ void k_sample(unsigned long long seed, long long gid0, const long long* episode, long n, double xlo, double xhi,
                         double ylo, double yhi, double min_sep, int npool, double* xy, int* task_draw, int* attempts) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  unsigned long long gid = (unsigned long long)(gid0 + e);
  unsigned ep = (unsigned)episode[e];
  double p[6];
  int na = philox_place(seed, gid, ep, xlo, xhi, ylo, yhi, min_sep, 1000, p);
  for (int k = 0; k < 6; k++) xy[6 * e + k] = p[k];
  if (task_draw) task_draw[e] = philox_task(seed, gid, ep, npool);
  if (attempts) attempts[e] = na;
}

// Philox yaw draw of the three cubes (mm_rng.h): theta [N,3] and (cos, sin)(theta / 2) [N,6]
__global__ void k_sample_yaw(unsigned long long seed, long long gid0, const long long* episode, long n, double* theta,
                             double* cs) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  for (int o = 0; o < 3; o++) {
    double th = philox_yaw(seed, (unsigned long long)(gid0 + e), (unsigned)episode[e], o);
    if (theta) theta[3 * e + o] = th;
    cs[6 * e + 2 * o] = cos(0.5 * th);
    cs[6 * e + 2 * o + 1] = sin(0.5 * th);
  }
}

// Episode draw of the vectorised env for the envs selected by `mask`: placement (rejection sampler), task (fixed /
// cycled by global id / drawn) and yaw from the env's Philox stream (mm_rng.h), then the episode counter advances.
struct SampleArgs {
  unsigned long long seed;
  long long gid0;
  long long* episode;
  const unsigned char* mask;
  long n;
  double xlo, xhi, ylo, yhi, min_sep;
  const int* pool;  // [npool][2] (object index, bin index)
  int npool, task_mode;  // 0 first pool entry | 1 pool[gid % npool] | 2 Philox draw
  int do_xy, do_yaw, advance;
  double* xy; int* task; int* attempts; double* theta; double* yaw_cs; double* stats;
};
__global__ void k_sample_episode(SampleArgs a) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n || (a.mask && !a.mask[e])) return;
  unsigned long long gid = (unsigned long long)(a.gid0 + e);
  unsigned ep = (unsigned)a.episode[e];
  if (a.do_xy) {
    double p[6];
    int na = philox_place(a.seed, gid, ep, a.xlo, a.xhi, a.ylo, a.yhi, a.min_sep, 1000, p);
    for (int k = 0; k < 6; k++) a.xy[6 * e + k] = p[k];
    if (a.attempts) a.attempts[e] = na;
    if (na == 0 && a.stats) atomicAdd(a.stats + 6, 1.0);  // the reference raises RuntimeError here (randomization.py:84-87)
  }
  if (a.task) {
    int k = a.task_mode == 2 ? philox_task(a.seed, gid, ep, a.npool) : (a.task_mode == 1 ? (int)(gid % (unsigned long long)a.npool) : 0);
    a.task[2 * e] = a.pool[2 * k]; a.task[2 * e + 1] = a.pool[2 * k + 1];
  }
  if (a.do_yaw)
    for (int o = 0; o < 3; o++) {
      double th = philox_yaw(a.seed, gid, ep, o);
      if (a.theta) a.theta[3 * e + o] = th;
      a.yaw_cs[6 * e + 2 * o] = cos(0.5 * th);
      a.yaw_cs[6 * e + 2 * o + 1] = sin(0.5 * th);
    }
  if (a.advance) a.episode[e] = (long long)ep + 1;
}

// Load-aware schedule of the next step (library-side, no host round trip): a counting sort of the envs by the busy time
// of their previous step (SM cycles / 256, SCHED_BINS bins of 2^SCHED_SHIFT units), heaviest first, dealt round-robin
// over the chunks so that every chunk starts its most expensive envs first and neighbouring warps run envs of similar
// cost (same code at the same time: shorter stage tails, better instruction-cache hit rate).  One CTA; results of the
// step never depend on the order (every env writes only its own slices; convex results are addressed by queue slot).
constexpr int SCHED_BINS = 2048, SCHED_SHIFT = 6, SCHED_THREADS = 1024;
__global__ void __launch_bounds__(SCHED_THREADS) k_schedule(const int* __restrict__ work, int* __restrict__ order, int n,
                                                            int chunk, int nchunk) {
  __shared__ int hist[SCHED_BINS];   // counts, then running start offsets (descending bins)
  __shared__ int part[SCHED_THREADS / 32];
  const int tid = threadIdx.x;
  for (int b = tid; b < SCHED_BINS; b += SCHED_THREADS) hist[b] = 0;
  __syncthreads();
  auto bin_of = [](int w) { int b = w >> SCHED_SHIFT; return b < 0 ? 0 : (b >= SCHED_BINS ? SCHED_BINS - 1 : b); };
  for (int i = tid; i < n; i += SCHED_THREADS) atomicAdd(&hist[bin_of(work[i])], 1);
  __syncthreads();
  // exclusive prefix over the bins in DESCENDING order: thread t owns bins SCHED_BINS-1-2t and SCHED_BINS-2-2t
  const int b0 = SCHED_BINS - 1 - 2 * tid, b1 = b0 - 1;
  const int c0 = hist[b0], c1 = hist[b1];
  int v = c0 + c1, incl = v;
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += t; }
  if ((tid & 31) == 31) part[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    int pv = part[tid], pi = pv;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, pi, o); if (tid >= o) pi += t; }
    part[tid] = pi - pv;
  }
  __syncthreads();
  const int excl = part[tid >> 5] + incl - v;
  hist[b0] = excl;
  hist[b1] = excl + c0;
  __syncthreads();
  const int last_sz = n - (nchunk - 1) * chunk;  // the last chunk may be shorter
  for (int i = tid; i < n; i += SCHED_THREADS) {
    int r = atomicAdd(&hist[bin_of(work[i])], 1);  // rank in descending order of work
    int c, pos;
    if (nchunk == 1) { c = 0; pos = r; }
    else if (r < last_sz * nchunk) { c = r % nchunk; pos = r / nchunk; }
    else { int r2 = r - last_sz * nchunk; c = r2 % (nchunk - 1); pos = last_sz + r2 / (nchunk - 1); }
    order[c * chunk + pos] = i;
  }
}

// Bookkeeping of the vectorised env after a step: running returns, episode statistics, the mask of finished envs,
// the last observation of finished episodes, diagnostics.  One thread per env; the statistics are block-reduced.
__global__ void k_post_step(StatePtrs st, StepOut out, long n, double* ep_return, unsigned char* reset_mask, float* final_obs,
                            double* stats, int auto_reset) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  double v[6] = {0, 0, 0, 0, 0, 0};  // episodes, successes, sum length, sum return, non-finite resets, overflow episodes
  if (e < n) {
    double ret = ep_return[e] + (double)out.reward[e];
    bool done = out.terminated[e] || out.truncated[e];
    int bad = st.diag[4 * e + 3];
    if (bad) { v[4] = bad; st.diag[4 * e + 3] = 0; }
    if (done) {
      v[0] = 1; v[1] = out.success[e] ? 1 : 0; v[2] = st.step_count[e]; v[3] = ret;
      v[5] = st.diag[4 * e + 2] ? 1 : 0;
    }
    ep_return[e] = (done && auto_reset) ? 0.0 : ret;
    if (reset_mask) reset_mask[e] = done ? 1 : 0;
  }
  if (final_obs) {  // coalesced copy of this block's observations
    long base = (long)blockIdx.x * blockDim.x * OBS_DIM, lim = n * OBS_DIM;
    for (long i = base + threadIdx.x; i < base + (long)blockDim.x * OBS_DIM && i < lim; i += blockDim.x) final_obs[i] = out.obs[i];
  }
  if (stats) {
#pragma unroll
    for (int k = 0; k < 6; k++) {
      double x = v[k];
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) == 0 && x != 0.0) atomicAdd(stats + k, x);
    }
  }
}

// FMA-throughput microbenchmark: the measured denominator of the CUDA-core roofline (bench.py)
template <class T>
__global__ void __launch_bounds__(256) k_peak(T* out, int iters, T a, T b) {
  T x[16];
#pragma unroll
  for (int k = 0; k < 16; k++) x[k] = (T)(threadIdx.x + k);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = x[k] * a + b;
  }
  T acc = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) acc += x[k];
  if (acc == (T)-12345) out[0] = acc;  // never true; keeps the chain alive
}

}  // namespace

struct mm_handle {
  mm_config cfg;
  void* d_model = nullptr;
  void* d_geom = nullptr;
  void* d_work_reals = nullptr;  // per-env workspace: contacts, survivors, solver rows
  int* d_work_ints = nullptr;
  char* d_ctx = nullptr;         // env images between the stage kernels
  int convex_grid = 0, heavy_grid = 0;
  unsigned char* d_hflag = nullptr;  // [N] contact-rich flag of the current round
  int* d_h_items = nullptr;          // [nchunk][chunk]
  int heavy_min = 0;                 // 0 = every env takes the warp-per-env stage C
  // per-stage device timing (mm_stage_times): events recorded around every stage launch of mm_step on its own stream
  bool timing = false;
  struct Timed { int kind; cudaEvent_t a, b; };
  std::vector<Timed> timed;
  std::vector<cudaEvent_t> ev_pool;
  // CUDA-graph path of mm_step: the ~450 launches, forks and joins of one control step are captured once per distinct
  // argument set (state / action / output pointers, action mode) and replayed with one cudaGraphLaunch (MM_GRAPH=0: off)
  struct StepGraph { StepParams key; cudaGraphExec_t exec; long long launches; unsigned long long stamp; };
  std::vector<StepGraph> graphs;
  bool use_graph = true;
  unsigned long long graph_clock = 0;
  cudaStream_t gstream = nullptr;    // capture origin; also carries the replay when the caller's stream is the legacy default stream
  cudaEvent_t ev_gin = nullptr, ev_gout = nullptr;
  int group_a = 0;                   // MM_GROUP_A=16: stage A by 16-lane groups (two envs per warp) next to a 32-lane stage C (experiment)
  bool round_major = true;           // MM_ISSUE=0: launches issued chunk after chunk instead of round after round
  bool fuse_ca = false;              // MM_FUSE_CA=1: stage A fused behind stage C (measured slower; experiment switch)
  cudaStream_t hside[16] = {};        // sibling streams of `side` for the contact-rich stage C
  cudaEvent_t ev_x[16] = {}, ev_h[16] = {};
  // chunks of the batch and their convex-pair queues
  long chunk = 0;
  int nchunk = 0, nstream = 0;
  void* d_q_items = nullptr;   // [nchunk][q_cap]
  void* d_q_res = nullptr;
  int* d_q_ctr = nullptr;      // [nchunk][2][NROUND] counts, heads
  int q_cap = 0;
  cudaStream_t side[16] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[16] = {};
  float* d_tgt = nullptr;
  const double* yaw_cs = nullptr;  // caller-owned, used by mm_reset when placements are given
  // staging for the host-buffer path
  float* d_actions = nullptr;
  float* d_obs = nullptr;
  float* d_reward = nullptr;
  unsigned char* d_flags = nullptr;  // terminated | truncated | success, N each
  long long launches = 0;
  long long* d_cycles = nullptr;
  // library-side load-aware schedule (k_schedule): on for batches of more than 32 envs unless MM_BALANCE=0 or the caller
  // supplies an order of its own (mm_set_schedule)
  bool balance = false;
  int* d_bal_order = nullptr;
  int* d_bal_work = nullptr;
  const int* d_order = nullptr;  // mm_set_schedule
  int* d_work = nullptr;  // optional per-env cycle counts (mm_set_schedule)
};

// every entry point runs on the handle's device whatever the caller's current device is
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define GUARD(h)                                                     \
  DeviceGuard guard_((h)->cfg.device);                               \
  if (!guard_.ok) return fail("cannot select the handle's CUDA device")

namespace {

size_t real_bytes(const mm_config* c) { return c->precision ? 4 : 8; }

typedef cudaError_t (*prepare_fn)();
typedef cudaError_t (*launch_fn)(int, const StepParams&, int, int, cudaStream_t);
typedef cudaError_t (*resident_fn)(int*, int*);
int inst_index(const mm_config& c) { return (c.precision ? 3 : 0) + (c.group == 32 ? 0 : (c.group == 16 ? 1 : 2)); }
const prepare_fn PREPARE[6] = {prepare_f64_32, prepare_f64_16, prepare_f64_8, prepare_f32_32, prepare_f32_16, prepare_f32_8};
const resident_fn RESIDENT[6] = {resident_f64_32, resident_f64_16, resident_f64_8, resident_f32_32, resident_f32_16, resident_f32_8};
const launch_fn LAUNCH[6] = {launch_f64_32, launch_f64_16, launch_f64_8, launch_f32_32, launch_f32_16, launch_f32_8};

template <class T>
int upload_model(mm_handle* h) {
  static GeomDev<T> gm;  // ~40 KB: filled once per process
  fill_geom(gm);
  CK(cudaMalloc(&h->d_geom, sizeof gm));
  CK(cudaMemcpy(h->d_geom, &gm, sizeof gm, cudaMemcpyHostToDevice));
  ModelDev<T> m;
  std::memset(&m, 0, sizeof m);
  fill_model(m);
  m.geom = reinterpret_cast<const GeomDev<T>*>(h->d_geom);
  CK(cudaMalloc(&h->d_model, sizeof m));
  CK(cudaMemcpy(h->d_model, &m, sizeof m, cudaMemcpyHostToDevice));
  return 0;
}

StatePtrs to_ptrs(const mm_state* s) {
  StatePtrs st;
  st.qpos = s->qpos; st.qvel = s->qvel; st.ctrl = s->ctrl; st.warm = s->warm; st.tinit = s->tinit; st.eepose = s->eepose;
  st.fsm_f = s->fsm_f; st.hwm = s->hwm; st.kin = s->kin; st.step_count = s->step_count; st.task = s->task; st.fsm_i = s->fsm_i; st.fsm_tasks = s->fsm_tasks;
  st.flags = s->flags; st.diag = s->diag;
  return st;
}

}  // namespace

extern "C" {

const char* mm_last_error(void) { return g_err.c_str(); }

size_t mm_workspace_bytes(const mm_config* cfg) {
  size_t n = (size_t)cfg->num_envs, rb = real_bytes(cfg);
  // per env: staging buffers, contact list + solver rows, stage image, share of the convex queue
  size_t per_env = 4 * 4 + (ACTION_STRIDE + OBS_DIM + 1) * 4 + 3 + (size_t)WORK_REALS * rb + (size_t)WORK_INTS * 4 +
                   (cfg->precision ? ctx_stride<float>() : ctx_stride<double>()) +
                   64 * (sizeof(CvxItem) + (cfg->precision ? sizeof(CvxRes<float>) : sizeof(CvxRes<double>)));
  return n * per_env;
}

int mm_create(const mm_config* cfg, mm_handle** out) {
  if (!cfg || !out) return fail("mm_create: null argument");
  if (cfg->num_envs <= 0) return fail("mm_create: num_envs must be positive");
  if (cfg->group != 8 && cfg->group != 16 && cfg->group != 32) return fail("mm_create: group must be 8, 16 or 32");
  if (cfg->precision != 0 && cfg->precision != 1) return fail("mm_create: precision must be 0 (f64) or 1 (f32)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("mm_create: no CUDA device (this library has no CPU path)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail("mm_create: no such CUDA device");
  mm_handle* h = new mm_handle();
  h->cfg = *cfg;
  GUARD(h);
  size_t n = (size_t)cfg->num_envs, rb = real_bytes(cfg);
  if (cfg->precision == 0) {
    if (upload_model<double>(h) != 0) return -1;
  } else {
    if (upload_model<float>(h) != 0) return -1;
  }
  if (PREPARE[inst_index(h->cfg)]() != cudaSuccess) return fail("mm_create: kernel attribute set-up failed");
  CK(RESIDENT[inst_index(h->cfg)](&h->convex_grid, &h->heavy_grid));
  CK(cudaMalloc(&h->d_work_reals, n * WORK_REALS * rb));
  CK(cudaMalloc(&h->d_work_ints, n * WORK_INTS * sizeof(int)));
  size_t cstride = cfg->precision ? ctx_stride<float>() : ctx_stride<double>();
  CK(cudaMalloc(&h->d_ctx, n * cstride));
  CK(cudaMemset(h->d_ctx, 0, n * cstride));
  // chunks: MM_CHUNK envs each (default: the batch in MM_STREAMS pieces, at most 16,384 and at least 256 envs per piece)
  h->nstream = (int)env_long("MM_STREAMS", n < 8192 ? 8 : 4);  // measured: 8 for small batches (shorter tails), 4 from 8,192 envs on
  if (h->nstream < 1) h->nstream = 1;
  if (h->nstream > 16) h->nstream = 16;
  long chunk = env_long("MM_CHUNK", 0);
  if (chunk <= 0) {
    chunk = ((long)n + h->nstream - 1) / h->nstream;
    if (chunk > 16384) chunk = 16384;
    if (chunk < 256) chunk = 256;
  }
  if (((long)n + chunk - 1) / chunk > MAX_CHUNKS) chunk = ((long)n + MAX_CHUNKS - 1) / MAX_CHUNKS;
  h->chunk = chunk;
  h->nchunk = (int)(((long)n + chunk - 1) / chunk);
  if (h->nstream > h->nchunk) h->nstream = h->nchunk;
  h->q_cap = (int)(chunk * 64);
  size_t res_bytes = cfg->precision ? sizeof(CvxRes<float>) : sizeof(CvxRes<double>);
  CK(cudaMalloc(&h->d_q_items, (size_t)h->nchunk * 2 * h->q_cap * sizeof(CvxItem)));
  CK(cudaMalloc(&h->d_q_res, (size_t)h->nchunk * 2 * h->q_cap * res_bytes));
  CK(cudaMalloc(&h->d_q_ctr, (size_t)h->nchunk * 4 * NCTR * sizeof(int)));
  // contact-rich envs (box contacts + convex candidates >= MM_HEAVY; 0 = the CTA-per-env path is off, the default: at the
  // benchmarked batch sizes stage C is bound by the throughput of all envs, not by its slowest one - no gain measured)
  h->heavy_min = h->heavy_grid > 0 ? (int)env_long("MM_HEAVY", 0) : 0;
  h->fuse_ca = env_long("MM_FUSE_CA", 0) != 0;
  h->round_major = env_long("MM_ISSUE", 1) != 0;
  CK(cudaMalloc(&h->d_hflag, 2 * n));
  CK(cudaMemset(h->d_hflag, 0, 2 * n));
  CK(cudaMalloc(&h->d_h_items, (size_t)h->nchunk * 2 * chunk * sizeof(int)));
  for (int i = 0; i < h->nstream; i++) {
    CK(cudaStreamCreateWithFlags(&h->side[i], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&h->hside[i], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_x[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_h[i], cudaEventDisableTiming));
  }
  CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  h->use_graph = env_long("MM_GRAPH", 1) != 0;
  h->group_a = (int)env_long("MM_GROUP_A", cfg->group);
  if (h->group_a != cfg->group) {
    if (h->group_a != 16 && h->group_a != 8 && h->group_a != 32) return fail("mm_create: MM_GROUP_A must be 32, 16 or 8");
    mm_config ca = *cfg;
    ca.group = h->group_a;
    if (PREPARE[inst_index(ca)]() != cudaSuccess) return fail("mm_create: kernel attribute set-up failed (stage A group)");
  }
  h->balance = env_long("MM_BALANCE", 1) != 0 && n > 32;
  if (h->balance) {
    CK(cudaMalloc(&h->d_bal_order, n * sizeof(int)));
    CK(cudaMalloc(&h->d_bal_work, n * sizeof(int)));
    CK(cudaMemset(h->d_bal_work, 0, n * sizeof(int)));
  }
  CK(cudaStreamCreateWithFlags(&h->gstream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&h->ev_gin, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&h->ev_gout, cudaEventDisableTiming));
  CK(cudaMalloc(&h->d_tgt, n * 4 * sizeof(float)));
  CK(cudaMemset(h->d_tgt, 0, n * 4 * sizeof(float)));
  CK(cudaMalloc(&h->d_actions, n * ACTION_STRIDE * sizeof(float)));
  CK(cudaMalloc(&h->d_obs, n * OBS_DIM * sizeof(float)));
  CK(cudaMalloc(&h->d_reward, n * sizeof(float)));
  CK(cudaMalloc(&h->d_flags, n * 3));
  *out = h;
  return 0;
}

void mm_destroy(mm_handle* h) {
  if (!h) return;
  DeviceGuard guard_(h->cfg.device);
  for (int i = 0; i < h->nstream; i++) {
    if (h->side[i]) cudaStreamDestroy(h->side[i]);
    if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
    if (h->hside[i]) cudaStreamDestroy(h->hside[i]);
    if (h->ev_x[i]) cudaEventDestroy(h->ev_x[i]);
    if (h->ev_h[i]) cudaEventDestroy(h->ev_h[i]);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  for (auto& e : h->graphs) cudaGraphExecDestroy(e.exec);
  cudaFree(h->d_bal_order); cudaFree(h->d_bal_work);
  if (h->gstream) cudaStreamDestroy(h->gstream);
  if (h->ev_gin) cudaEventDestroy(h->ev_gin);
  if (h->ev_gout) cudaEventDestroy(h->ev_gout);
  for (auto& t : h->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
  for (auto e : h->ev_pool) cudaEventDestroy(e);
  cudaFree(h->d_model); cudaFree(h->d_geom); cudaFree(h->d_work_reals); cudaFree(h->d_work_ints); cudaFree(h->d_ctx);
  cudaFree(h->d_q_items); cudaFree(h->d_q_res);
  cudaFree(h->d_q_ctr); cudaFree(h->d_tgt); cudaFree(h->d_hflag); cudaFree(h->d_h_items);
  cudaFree(h->d_actions); cudaFree(h->d_obs); cudaFree(h->d_reward); cudaFree(h->d_flags);
  delete h;
}

namespace {
void base_params(mm_handle* h, const mm_state* st, StepParams& p) {
  p.st = to_ptrs(st);
  p.model = h->d_model;
  p.work_reals = h->d_work_reals; p.work_ints = h->d_work_ints; p.ctx = h->d_ctx;
  p.tgt_kp = h->d_tgt; p.n = h->cfg.num_envs; p.slot0 = 0; p.nslot = h->cfg.num_envs;
  p.reward_type = h->cfg.reward_type; p.max_steps = h->cfg.max_episode_steps;
  p.q_cap = h->q_cap;
}
}  // namespace

int mm_reset(mm_handle* h, const mm_state* st, const uint8_t* mask, const double* obj_xy, const int32_t* task,
             float* obs, void* stream) {
  if (!h || !st || !task) return fail("mm_reset: null argument");
  GUARD(h);
  StepParams p{};
  base_params(h, st, p);
  p.mask = mask; p.obj_xy = obj_xy; p.yaw_cs = obj_xy ? h->yaw_cs : nullptr; p.task = task; p.obs = obs;
  h->launches++;
  CK(LAUNCH[inst_index(h->cfg)](3, p, 0, 0, (cudaStream_t)stream));
  return 0;
}

namespace {
// The launches of one control step on `main` (and the handle's side streams, forked from and joined back to it).
int enqueue_step(mm_handle* h, const StepParams& p, cudaStream_t main) {
  const launch_fn launch = LAUNCH[inst_index(h->cfg)];
  mm_config cfg_a = h->cfg;
  cfg_a.group = h->group_a;
  const launch_fn launch_a = LAUNCH[inst_index(cfg_a)];
  const size_t rb = real_bytes(&h->cfg);
  const size_t res_bytes = h->cfg.precision ? sizeof(CvxRes<float>) : sizeof(CvxRes<double>);
  CK(cudaMemsetAsync(h->d_q_ctr, 0, (size_t)h->nchunk * 4 * NCTR * sizeof(int), main));
  if (h->balance && p.order == h->d_bal_order) {
    k_schedule<<<1, SCHED_THREADS, 0, main>>>(p.work, h->d_bal_order, (int)p.n, (int)h->chunk, h->nchunk);
    CK(cudaGetLastError());
    h->launches++;
  }
  // (with the contact-rich path on, even a single chunk runs on a side stream: its sibling stream needs one to pair with)
  const bool forked = h->nstream > 1 || h->heavy_min > 0;
  if (forked) {
    CK(cudaEventRecord(h->ev_fork, main));
    for (int i = 0; i < h->nstream; i++) CK(cudaStreamWaitEvent(h->side[i], h->ev_fork, 0));
  }
  // the chunks of a stream run one after the other.  Issue order: round-major (all chunks' round r before any round
  // r + 1, so that every stream has work from the first microseconds on) unless MM_ISSUE=0 (chunk after chunk)
  const bool fuse = h->fuse_ca;  // experiment switch: stage A of round r + 1 inside the stage C kernel of round r
  auto params_of = [&](int c) {
    StepParams pc = p;
    pc.slot0 = (long)c * h->chunk;
    pc.nslot = pc.slot0 + h->chunk <= p.n ? h->chunk : p.n - pc.slot0;
    pc.q_items = (char*)h->d_q_items + (size_t)c * 2 * h->q_cap * sizeof(CvxItem);
    pc.q_res = (char*)h->d_q_res + (size_t)c * 2 * h->q_cap * res_bytes;
    pc.q_count = h->d_q_ctr + (size_t)c * 4 * NCTR;
    pc.q_head = pc.q_count + NCTR;
    pc.h_count = pc.q_count + 2 * NCTR;
    pc.h_head = pc.q_count + 3 * NCTR;
    pc.hflag = h->heavy_min > 0 ? h->d_hflag : nullptr;
    pc.h_items = h->d_h_items + (size_t)c * 2 * h->chunk;
    pc.h_cap = (int)h->chunk;
    pc.heavy_min = h->heavy_min;
    return pc;
  };
  // launch + (optionally) a pair of timing events on the launching stream; kind 0 stage A | 1 convex | 2 stage C | 3 heavy
  auto timed_launch = [&](const StepParams& pc, int which, int kind, int sub, int grid_x, cudaStream_t st_) -> cudaError_t {
    if (!h->timing) return (which == 0 ? launch_a : launch)(which, pc, sub, grid_x, st_);
    cudaEvent_t ev[2];
    for (int k = 0; k < 2; k++) {
      if (h->ev_pool.empty()) { cudaError_t e = cudaEventCreate(&ev[k]); if (e != cudaSuccess) return e; }
      else { ev[k] = h->ev_pool.back(); h->ev_pool.pop_back(); }
    }
    cudaError_t e = cudaEventRecord(ev[0], st_);
    if (e != cudaSuccess) return e;
    e = (which == 0 ? launch_a : launch)(which, pc, sub, grid_x, st_);
    if (e != cudaSuccess) return e;
    e = cudaEventRecord(ev[1], st_);
    h->timed.push_back({kind, ev[0], ev[1]});
    return e;
  };
  // one round of one chunk (after the last round: the epilogue kernel)
  auto issue = [&](int c, int sub) -> int {
    const int si = c % h->nstream;
    cudaStream_t s = forked ? h->side[si] : main;
    const StepParams pc = params_of(c);
    if (sub == 0 || !fuse) CK(timed_launch(pc, 0, 0, sub, 0, s));
    CK(timed_launch(pc, 1, 1, sub, h->convex_grid, s));
    if (h->heavy_min > 0) {  // contact-rich envs: a CTA each, next to the warp-per-env launch of the others
      CK(cudaEventRecord(h->ev_x[si], s));
      CK(cudaStreamWaitEvent(h->hside[si], h->ev_x[si], 0));
      CK(timed_launch(pc, fuse ? 7 : 5, 3, sub, h->heavy_grid, h->hside[si]));
      CK(cudaEventRecord(h->ev_h[si], h->hside[si]));
    }
    CK(timed_launch(pc, fuse ? 6 : 2, 2, sub, 0, s));
    if (h->heavy_min > 0) CK(cudaStreamWaitEvent(s, h->ev_h[si], 0));
    if (sub == NROUND - 1) {
      CK(launch(8, pc, 0, 0, s));  // reward, termination, observation, state store
      h->launches += 1 + (fuse ? 1 : NROUND) + (h->heavy_min > 0 ? 3 : 2) * NROUND;
    }
    return 0;
  };
  if (h->round_major) {
    for (int sub = 0; sub < NROUND; sub++)
      for (int c = 0; c < h->nchunk; c++) { int e = issue(c, sub); if (e) return e; }
  } else {
    for (int c = 0; c < h->nchunk; c++)
      for (int sub = 0; sub < NROUND; sub++) { int e = issue(c, sub); if (e) return e; }
  }
  if (forked)
    for (int i = 0; i < h->nstream; i++) {
      CK(cudaEventRecord(h->ev_join[i], h->side[i]));
      CK(cudaStreamWaitEvent(main, h->ev_join[i], 0));
    }
  return 0;
}
}  // namespace

int mm_step(mm_handle* h, const mm_state* st, const float* actions, int action_mode, const mm_step_out* out,
            void* stream) {
  if (!h || !st || !actions || !out) return fail("mm_step: null argument");
  if (action_mode < 0 || action_mode > 4) return fail("mm_step: bad action_mode");
  GUARD(h);
  cudaStream_t main = (cudaStream_t)stream;
  StepParams p{};
  base_params(h, st, p);
  p.out.obs = out->obs; p.out.reward = out->reward; p.out.terminated = out->terminated; p.out.truncated = out->truncated;
  p.out.success = out->success; p.out.reward_components = out->reward_components;
  p.actions = actions; p.mode = action_mode;
  p.cycles = h->d_cycles;
  p.order = h->d_order;
  p.work = h->d_work;
  if (h->balance && !h->d_order) {  // the library's own schedule (k_schedule at the head of the step)
    p.order = h->d_bal_order;
    if (!p.work) p.work = h->d_bal_work;
  }
  if (h->use_graph && !h->timing) {
    mm_handle::StepGraph* g = nullptr;
    for (auto& e : h->graphs)
      if (std::memcmp(&e.key, &p, sizeof(StepParams)) == 0) { g = &e; break; }
    if (!g) {
      if (h->graphs.size() >= 64) {  // least recently used out
        size_t lru = 0;
        for (size_t k = 1; k < h->graphs.size(); k++) if (h->graphs[k].stamp < h->graphs[lru].stamp) lru = k;
        cudaGraphExecDestroy(h->graphs[lru].exec);
        h->graphs.erase(h->graphs.begin() + lru);
      }
      const long long l0 = h->launches;
      cudaGraph_t graph = nullptr;
      CK(cudaStreamBeginCapture(h->gstream, cudaStreamCaptureModeRelaxed));
      int rc = enqueue_step(h, p, h->gstream);
      cudaError_t ee = cudaStreamEndCapture(h->gstream, &graph);
      const long long per_step = h->launches - l0;
      h->launches = l0;
      if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
      CK(ee);
      mm_handle::StepGraph e;
      std::memcpy(&e.key, &p, sizeof(StepParams));
      e.launches = per_step; e.stamp = 0;
      ee = cudaGraphInstantiate(&e.exec, graph, 0);
      cudaGraphDestroy(graph);
      CK(ee);
      h->graphs.push_back(e);
      g = &h->graphs.back();
    }
    g->stamp = ++h->graph_clock;
    if (main == nullptr || main == cudaStreamLegacy) {  // the legacy stream takes no graph launch: go through gstream
      CK(cudaEventRecord(h->ev_gin, main));
      CK(cudaStreamWaitEvent(h->gstream, h->ev_gin, 0));
      CK(cudaGraphLaunch(g->exec, h->gstream));
      CK(cudaEventRecord(h->ev_gout, h->gstream));
      CK(cudaStreamWaitEvent(main, h->ev_gout, 0));
    } else {
      CK(cudaGraphLaunch(g->exec, main));
    }
    h->launches += g->launches;
    return 0;
  }
  return enqueue_step(h, p, main);
}

int mm_step_host_async(mm_handle* h, const mm_state* st, const float* h_actions, int action_mode, float* h_obs,
                       float* h_reward, uint8_t* h_terminated, uint8_t* h_truncated, uint8_t* h_success, void* stream) {
  if (!h || !st || !h_actions) return fail("mm_step_host: null argument");
  GUARD(h);
  cudaStream_t s = (cudaStream_t)stream;
  size_t n = (size_t)h->cfg.num_envs;
  CK(cudaMemcpyAsync(h->d_actions, h_actions, n * ACTION_STRIDE * sizeof(float), cudaMemcpyHostToDevice, s));
  mm_step_out o;
  o.obs = h->d_obs; o.reward = h->d_reward; o.terminated = h->d_flags; o.truncated = h->d_flags + n;
  o.success = h->d_flags + 2 * n; o.reward_components = nullptr;
  int rc = mm_step(h, st, h->d_actions, action_mode, &o, stream);
  if (rc) return rc;
  if (h_obs) CK(cudaMemcpyAsync(h_obs, h->d_obs, n * OBS_DIM * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (h_reward) CK(cudaMemcpyAsync(h_reward, h->d_reward, n * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (h_terminated) CK(cudaMemcpyAsync(h_terminated, h->d_flags, n, cudaMemcpyDeviceToHost, s));
  if (h_truncated) CK(cudaMemcpyAsync(h_truncated, h->d_flags + n, n, cudaMemcpyDeviceToHost, s));
  if (h_success) CK(cudaMemcpyAsync(h_success, h->d_flags + 2 * n, n, cudaMemcpyDeviceToHost, s));
  return 0;
}

int mm_step_host(mm_handle* h, const mm_state* st, const float* h_actions, int action_mode, float* h_obs,
                 float* h_reward, uint8_t* h_terminated, uint8_t* h_truncated, uint8_t* h_success, void* stream) {
  int rc = mm_step_host_async(h, st, h_actions, action_mode, h_obs, h_reward, h_terminated, h_truncated, h_success, stream);
  if (rc) return rc;
  GUARD(h);
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

int mm_host_staging(mm_handle* h, mm_step_out* out) {
  if (!h || !out) return fail("mm_host_staging: null argument");
  size_t n = (size_t)h->cfg.num_envs;
  out->obs = h->d_obs; out->reward = h->d_reward; out->terminated = h->d_flags; out->truncated = h->d_flags + n;
  out->success = h->d_flags + 2 * n; out->reward_components = nullptr;
  return 0;
}

int mm_fsm_plan(mm_handle* h, const mm_state* st, int n_steps, float* actions_out, void* stream) {
  if (!h || !st) return fail("mm_fsm_plan: null argument");
  GUARD(h);
  long n = h->cfg.num_envs;
  k_fsm<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(to_ptrs(st), n, n_steps, actions_out);
  CK(cudaGetLastError());
  h->launches++;
  return 0;
}

int mm_sample_placements(mm_handle* h, uint64_t seed, int64_t env_id_offset, const int64_t* episode_index, double x_lo,
                         double x_hi, double y_lo, double y_hi, double min_separation, int32_t npool, double* obj_xy,
                         int32_t* task_draw, int32_t* attempts, void* stream) {
  if (!h || !episode_index || !obj_xy) return fail("mm_sample_placements: null argument");
  GUARD(h);
  if (npool <= 0) return fail("mm_sample_placements: npool must be positive");
  long n = h->cfg.num_envs;
  k_sample<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      (unsigned long long)seed, (long long)env_id_offset, (const long long*)episode_index, n, x_lo, x_hi, y_lo, y_hi,
      min_separation, npool, obj_xy, task_draw, attempts);
  CK(cudaGetLastError());
  h->launches++;
  return 0;
}

int mm_set_placement_yaw(mm_handle* h, const double* yaw_cs) {
  if (!h) return fail("mm_set_placement_yaw: null handle");
  h->yaw_cs = yaw_cs;
  return 0;
}

int mm_sample_yaw(mm_handle* h, uint64_t seed, int64_t env_id_offset, const int64_t* episode_index, double* theta,
                  double* yaw_cs, void* stream) {
  if (!h || !episode_index || !yaw_cs) return fail("mm_sample_yaw: null argument");
  GUARD(h);
  long n = h->cfg.num_envs;
  k_sample_yaw<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      (unsigned long long)seed, (long long)env_id_offset, (const long long*)episode_index, n, theta, yaw_cs);
  CK(cudaGetLastError());
  h->launches++;
  return 0;
}

int mm_sample_episode(mm_handle* h, uint64_t seed, int64_t env_id_offset, int64_t* episode_index, const uint8_t* mask,
                      double x_lo, double x_hi, double y_lo, double y_hi, double min_separation, const int32_t* pool,
                      int32_t npool, int32_t task_mode, int32_t randomize_xy, int32_t randomize_yaw, int32_t advance,
                      double* obj_xy, int32_t* task, int32_t* attempts, double* yaw_theta, double* yaw_cs, double* stats,
                      void* stream) {
  if (!h || !episode_index) return fail("mm_sample_episode: null argument");
  if (randomize_xy && !obj_xy) return fail("mm_sample_episode: randomize_xy needs obj_xy");
  if (randomize_yaw && !yaw_cs) return fail("mm_sample_episode: randomize_yaw needs yaw_cs");
  if (task && (!pool || npool <= 0 || task_mode < 0 || task_mode > 2)) return fail("mm_sample_episode: bad task pool / mode");
  GUARD(h);
  SampleArgs a;
  a.seed = (unsigned long long)seed; a.gid0 = (long long)env_id_offset; a.episode = (long long*)episode_index; a.mask = mask;
  a.n = h->cfg.num_envs; a.xlo = x_lo; a.xhi = x_hi; a.ylo = y_lo; a.yhi = y_hi; a.min_sep = min_separation;
  a.pool = pool; a.npool = npool; a.task_mode = task_mode; a.do_xy = randomize_xy; a.do_yaw = randomize_yaw; a.advance = advance;
  a.xy = obj_xy; a.task = task; a.attempts = attempts; a.theta = yaw_theta; a.yaw_cs = yaw_cs; a.stats = stats;
  k_sample_episode<<<(unsigned)((a.n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
  CK(cudaGetLastError());
  h->launches++;
  return 0;
}

int mm_post_step(mm_handle* h, const mm_state* st, const mm_step_out* out, double* ep_return, uint8_t* reset_mask,
                 float* final_obs, double* stats, int32_t auto_reset, void* stream) {
  if (!h || !st || !out || !ep_return) return fail("mm_post_step: null argument");
  GUARD(h);
  StepOut o;
  o.obs = out->obs; o.reward = out->reward; o.terminated = out->terminated; o.truncated = out->truncated;
  o.success = out->success; o.reward_components = out->reward_components;
  long n = h->cfg.num_envs;
  k_post_step<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(to_ptrs(st), o, n, ep_return, reset_mask, final_obs,
                                                                           stats, auto_reset);
  CK(cudaGetLastError());
  h->launches++;
  return 0;
}

int mm_measure_fma_peak(int device, int fp64, double* tflops) {
  if (!tflops) return fail("mm_measure_fma_peak: null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("mm_measure_fma_peak: no CUDA device");
  DeviceGuard guard_(device);
  if (!guard_.ok) return fail("mm_measure_fma_peak: cannot select the device");
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  void* buf;
  CK(cudaMalloc(&buf, 64));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int iters = 4096, blocks = prop.multiProcessorCount * 8, threads = 256;
  double best = 0;
  for (int rep = 0; rep < 6; rep++) {
    CK(cudaEventRecord(e0));
    if (fp64) k_peak<double><<<blocks, threads>>>((double*)buf, iters, 1.0000001, 1e-9);
    else k_peak<float><<<blocks, threads>>>((float*)buf, iters, 1.0000001f, 1e-9f);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * 16 * (double)iters * blocks * threads;
    double tf = fl / (ms * 1e-3) * 1e-12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
  *tflops = best;
  return 0;
}

int mm_ops(mm_handle* h, const mm_state* st, int ops, const double* target, void* stream) {
  if (!h || !st) return fail("mm_ops: null argument");
  if (ops <= 0 || ops > 7) return fail("mm_ops: ops must be a combination of MM_OP_IK, MM_OP_FORWARD, MM_OP_INTEGRATE");
  if ((ops & MM_OP_INTEGRATE) && !(ops & MM_OP_FORWARD)) return fail("mm_ops: MM_OP_INTEGRATE needs MM_OP_FORWARD");
  if ((ops & MM_OP_IK) && !target) return fail("mm_ops: MM_OP_IK needs a target array");
  GUARD(h);
  StepParams p{};
  base_params(h, st, p);
  p.ops = ops; p.target = target;
  h->launches++;
  CK(LAUNCH[inst_index(h->cfg)](4, p, 0, 0, (cudaStream_t)stream));
  return 0;
}

int mm_expert_actions(mm_handle* h, const mm_state* st, const float* abs_actions, float* encodings, void* stream) {
  if (!h || !st || !abs_actions || !encodings) return fail("mm_expert_actions: null argument");
  GUARD(h);
  long n = h->cfg.num_envs;
  k_expert<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(to_ptrs(st), n, abs_actions, encodings);
  CK(cudaGetLastError());
  h->launches++;
  return 0;
}

int mm_set_schedule(mm_handle* h, const int32_t* order, int32_t* work) {
  if (!h) return fail("mm_set_schedule: null handle");
  h->d_order = order;
  h->d_work = work;
  return 0;
}

int mm_set_cycle_buffer(mm_handle* h, long long* cycles) {
  if (!h) return fail("mm_set_cycle_buffer: null handle");
  h->d_cycles = cycles;
  return 0;
}

int mm_stage_timing(mm_handle* h, int32_t on) {
  if (!h) return fail("mm_stage_timing: null handle");
  h->timing = on != 0;
  return 0;
}

int mm_stage_times(mm_handle* h, double* ms, long long* launches) {
  if (!h || !ms) return fail("mm_stage_times: null argument");
  GUARD(h);
  CK(cudaDeviceSynchronize());
  for (int k = 0; k < 4; k++) { ms[k] = 0; if (launches) launches[k] = 0; }
  for (auto& t : h->timed) {
    float f = 0;
    CK(cudaEventElapsedTime(&f, t.a, t.b));
    ms[t.kind] += (double)f;
    if (launches) launches[t.kind]++;
    h->ev_pool.push_back(t.a);
    h->ev_pool.push_back(t.b);
  }
  h->timed.clear();
  return 0;
}

int mm_launch_count(mm_handle* h, long long* out) {
  if (!h || !out) return fail("mm_launch_count: null argument");
  *out = h->launches;
  return 0;
}

}  // extern "C"
