// Kernel templates of the step / reset launches and their per-(precision, group) entry points.
// Each (T, G) pair is instantiated in its own translation unit (mm_inst_*.cu) so the library builds
// in parallel; mm_kernels.cu holds the C ABI and dispatches through the mm_inst_* functions.
//
// One control step (mm_step) = 17 rounds of three batch-wide STAGE kernels (see mm_env.h):
//   k_stage_a  group per env   IK, kinematics, smooth dynamics, broad phase, box / plane narrow phase, queue push
//   k_convex   warp per PAIR   GJK + EPA of every queued (env, geom pair) of the batch
//   k_stage_c  group per env   contact assembly, constraint rows, Newton solver, integration [+ reward / obs / store]
// Every kernel is small enough for the instruction cache and has its own register / shared-memory budget, so many
// more warps are resident per SM than in one fused kernel, and a pile-up env's hull pairs are tested concurrently.
// The batch is cut into chunks that run their 51 launches on alternating streams: the tail of one chunk's kernel
// (its slowest env) overlaps with the other chunk's work, and a chunk's env images stay L2 resident.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

#include "mm_env.h"

namespace mm {

// CTA shapes (warps per CTA, minimum resident CTAs per SM = register budget) of the three stage kernels
// (measured on the B200, 4,096 / 16,384 / 65,536 envs: one-warp CTAs for the per-env stages - a finished env frees its
// slot at once - at 128 / 168 registers, and 255 registers for the convex kernel, whose simplex and shape state
// otherwise spills; profiles/r02_variants.txt)
#ifndef MM_WA
#define MM_WA 1
#endif
#ifndef MM_MINB_A
#define MM_MINB_A 16
#endif
#ifndef MM_WC
#define MM_WC 1
#endif
#ifndef MM_MINB_C
#define MM_MINB_C 12
#endif
#ifndef MM_WX
#define MM_WX 4
#endif
#ifndef MM_MINB_X
#define MM_MINB_X 2
#endif
// stage C / fused kernels: G = 8 packs four envs into a warp, so fewer warps fit the shared memory of a CTA
template <int G> constexpr int warps_c() { return G == 8 ? (MM_WC > 2 ? 2 : MM_WC) : MM_WC; }
constexpr int MAX_CHUNKS = 64;
constexpr int NROUND = ACTION_REPEAT + 1;
constexpr int NCTR = NROUND + 1;  // counters per kind and chunk (the last round's fused stage A does not exist: spare slot)

struct StepParams {
  StatePtrs st;
  StepOut out;
  const float* actions;
  const void* model;
  // per-env workspace (contacts, survivors, solver rows) and env images between the stages
  void* work_reals;   // [N][WORK_REALS]
  int* work_ints;     // [N][WORK_INTS]
  char* ctx;          // [N][ctx_stride]
  // convex-pair queue of this chunk
  void* q_items;
  void* q_res;
  int* q_count;       // [NROUND + 1]
  int* q_head;        // [NROUND + 1]
  int q_cap;
  // contact-rich envs of this chunk (stage C by a whole CTA each)
  unsigned char* hflag;   // [2][N] or null (two buffers alternate between the rounds, like the queue)
  int* h_items;           // [2][h_cap]
  int h_cap;
  int* h_count;           // [NROUND]
  int* h_head;            // [NROUND]
  int heavy_min;
  float* tgt_kp;
  const unsigned char* mask;
  const double* obj_xy;
  const double* yaw_cs;  // [N,6] cos(theta/2), sin(theta/2) of the three cubes, or null
  const int* task;
  float* obs;
  long n;                // envs of the handle
  long slot0, nslot;     // this launch covers schedule slots [slot0, slot0 + nslot)
  int mode, reward_type, max_steps;
  const double* target;  // [N,3] world EE targets for the IK op, or null
  int ops;
  const int* order;      // [N] env processed by each slot (heaviest first) or null = identity
  int* work;             // [N] out: busy cycles / 256 of each env's step, or null
  long long* cycles;     // [N,9] or null: SM clock cycles of each env's step: total, then per stage kernel (profiling aid)
};

template <class T, int G>
__device__ __forceinline__ void setup_group(Grp<G>& g) {
  g.lane = threadIdx.x % G;
  int inwarp = (threadIdx.x % 32) / G;
  g.mask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (inwarp * G));
}

template <class T>
__device__ __forceinline__ CvxQueue<T> queue_of(const StepParams& p, int sub) {
  CvxQueue<T> q;  // two buffers alternate between the rounds (stage C of round r reads, the fused stage A of r + 1 writes)
  q.items = reinterpret_cast<CvxItem*>(p.q_items) + (size_t)(sub & 1) * p.q_cap;
  q.res = reinterpret_cast<CvxRes<T>*>(p.q_res) + (size_t)(sub & 1) * p.q_cap;
  q.count = p.q_count + sub;
  q.head = p.q_head + sub;
  q.cap = p.q_cap;
  return q;
}

__device__ __forceinline__ HeavyList heavy_of(const StepParams& p, int sub) {
  HeavyList hv;
  hv.flag = p.hflag ? p.hflag + (size_t)(sub & 1) * p.n : nullptr;
  hv.items = p.h_items + (size_t)(sub & 1) * p.h_cap;
  hv.count = p.h_count + sub;
  hv.min_load = p.heavy_min;
  return hv;
}

template <class T>
__device__ __forceinline__ Work<T> work_of(const StepParams& p, long e) {
  return make_work(reinterpret_cast<T*>(p.work_reals) + e * WORK_REALS, p.work_ints + e * WORK_INTS);
}

template <class T> MM_HDN constexpr size_t scratch_c_bytes() { return (sizeof(Scratch<T>) + 15) / 16 * 16; }

template <class T, int G, int W>
__global__ void __launch_bounds__(32 * W, MM_MINB_A) k_stage_a(StepParams p, int sub) {
  extern __shared__ __align__(16) unsigned char smem[];
  Grp<G> g;
  setup_group<T, G>(g);
  const ModelDev<T>* md = reinterpret_cast<const ModelDev<T>*>(p.model);
  int gi = threadIdx.x / G;
  long slot = (long)blockIdx.x * (32 * W / G) + gi;
  if (slot >= p.nslot) return;
  slot += p.slot0;
  long e = p.order ? p.order[slot] : slot;
  long long t0 = clock64();
  Scratch<T>& s = *reinterpret_cast<Scratch<T>*>(smem + gi * scratch_a_bytes<T>());
  Work<T> w = work_of<T>(p, e);
  stage_a<T, G>(g, s, *md, w, p.st, e, sub, p.actions, p.mode, p.ctx, queue_of<T>(p, sub), heavy_of(p, sub));
  if (g.lane == 0) {
    long long dt = clock64() - t0;
    if (p.work) p.work[e] = (sub == 0 ? 0 : p.work[e]) + (int)(dt >> 8);
    if (p.cycles) { p.cycles[9 * e + 1] += dt; p.cycles[9 * e] += dt; }
  }
}

template <class T, int G, int W, bool FUSE>
__global__ void __launch_bounds__(32 * W, MM_MINB_C) k_stage_c(StepParams p, int sub) {
  extern __shared__ __align__(16) unsigned char smem[];
  Grp<G> g;
  setup_group<T, G>(g);
  const ModelDev<T>* md = reinterpret_cast<const ModelDev<T>*>(p.model);
  int gi = threadIdx.x / G;
  long slot = (long)blockIdx.x * (32 * W / G) + gi;
  if (slot >= p.nslot) return;
  slot += p.slot0;
  long e = p.order ? p.order[slot] : slot;
  if (p.hflag && p.hflag[(size_t)(sub & 1) * p.n + e]) return;  // contact-rich env: k_stage_c_heavy runs it
  long long t0 = clock64();
  Scratch<T>& s = *reinterpret_cast<Scratch<T>*>(smem + gi * scratch_c_bytes<T>());
  Work<T> w = work_of<T>(p, e);
  stage_c<T, G, FUSE>(g, s, *md, w, p.st, e, sub, p.ctx, queue_of<T>(p, sub), queue_of<T>(p, sub + 1), heavy_of(p, sub + 1),
                      p.reward_type, p.max_steps, p.out, p.tgt_kp);
  if (g.lane == 0) {
    long long dt = clock64() - t0;
    if (p.work) p.work[e] += (int)(dt >> 8);
    if (p.cycles) { p.cycles[9 * e + 3] += dt; p.cycles[9 * e] += dt; }
  }
}

// epilogue of the step (reward, termination, observation, state store): warp per env on the env images
template <class T, int G, int W>
__global__ void __launch_bounds__(32 * W, MM_MINB_A) k_finish(StepParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  Grp<G> g;
  setup_group<T, G>(g);
  int gi = threadIdx.x / G;
  long slot = (long)blockIdx.x * (32 * W / G) + gi;
  if (slot >= p.nslot) return;
  slot += p.slot0;
  long e = p.order ? p.order[slot] : slot;
  Scratch<T>& s = *reinterpret_cast<Scratch<T>*>(smem + gi * scratch_a_bytes<T>());
  Work<T> w = work_of<T>(p, e);
  stage_finish<T, G>(g, s, w, p.st, e, p.ctx, p.reward_type, p.max_steps, p.out, p.tgt_kp);
}

// stage C of the contact-rich envs of a round: persistent 128-thread CTAs, one env at a time per CTA (Grp<128>)
#ifndef MM_MINB_H
#define MM_MINB_H 3
#endif
template <class T, bool FUSE>
__global__ void __launch_bounds__(128, MM_MINB_H) k_stage_c_heavy(StepParams p, int sub) {
  extern __shared__ __align__(16) unsigned char smem[];
  Grp<128> g;
  g.lane = threadIdx.x;
  g.mask = 0xffffffffu;
  g.xs = reinterpret_cast<double*>(smem + scratch_c_bytes<T>());
  g.par = 0;
  const ModelDev<T>* md = reinterpret_cast<const ModelDev<T>*>(p.model);
  Scratch<T>& s = *reinterpret_cast<Scratch<T>*>(smem);
  __shared__ int s_item;
  const int count = p.h_count[sub];
  while (true) {
    if (threadIdx.x == 0) s_item = atomicAdd(p.h_head + sub, 1);
    __syncthreads();
    int i = s_item;
    __syncthreads();
    if (i >= count) break;
    long e = p.h_items[(size_t)(sub & 1) * p.h_cap + i];
    long long t0 = clock64();
    Work<T> w = work_of<T>(p, e);
    stage_c<T, 128, FUSE>(g, s, *md, w, p.st, e, sub, p.ctx, queue_of<T>(p, sub), queue_of<T>(p, sub + 1), heavy_of(p, sub + 1),
                          p.reward_type, p.max_steps, p.out, p.tgt_kp);
    if (g.lane == 0) {
      long long dt = clock64() - t0;
      if (p.work) p.work[e] += (int)(dt >> 8);
      if (p.cycles) { p.cycles[9 * e + 3] += dt; p.cycles[9 * e] += dt; }
    }
    __syncthreads();
  }
}
template <class T> size_t smem_h() { return scratch_c_bytes<T>() + (8 + 128) * sizeof(double); }

// convex stage: persistent warps take (env, geom pair) items off the queue of this round
template <class T>
struct ConvexSmem {
  T vert[EPA_MAXV * 6];  // polytope vertices next to its faces: the canonical-id scan, the visibility test and the face
                         // construction of every EPA iteration read them (from global memory they cost an L2 round trip each)
  T face[EPA_MAXF * 4];
  T bpos[NDB][3], bR[NDB][9];
  int fidx[EPA_MAXF], edge[EPA_MAXE], canon[EPA_MAXV], ecan[EPA_MAXE];
};
template <class T, int W>
__global__ void __launch_bounds__(32 * W, MM_MINB_X) k_convex(StepParams p, int sub) {
  extern __shared__ __align__(16) unsigned char smem[];
  Grp<32> g;
  setup_group<T, 32>(g);
  const ModelDev<T>* md = reinterpret_cast<const ModelDev<T>*>(p.model);
  const GeomDev<T>& gm = *md->geom;
  int wi = threadIdx.x / 32;
  ConvexSmem<T>& cs = reinterpret_cast<ConvexSmem<T>*>(smem)[wi];
  CvxQueue<T> q = queue_of<T>(p, sub);
  int count = *q.count;
  if (count > q.cap) count = q.cap;
  EpaMem<T> em;
  em.vert = cs.vert;
  em.face = cs.face; em.fidx = cs.fidx; em.edge = cs.edge; em.canon = cs.canon; em.ecan = cs.ecan;
  // Software pipeline over the queue: while a pair is in GJK / EPA, the item of the next pair is being loaded and the
  // claim of the one after that (an atomic on the queue head) is in flight; of the env image only the poses of the
  // pair's two bodies are fetched, one value per lane.  One exposed L2 round trip per pair instead of three.
  const char* ctx = p.ctx;
  int i = 0;
  if (g.lane == 0) i = atomicAdd(q.head, 1);
  i = __shfl_sync(0xffffffffu, i, 0);
  if (i >= count) return;
  CvxItem it = q.items[i];
  // (claiming ahead holds a pair back from the warps that are idle at the end of a short queue: only with >= 6 pairs per warp)
  const bool ahead = count > 6 * (int)(gridDim.x * W);
  int raw_next = 0;
  if (ahead && g.lane == 0) raw_next = atomicAdd(q.head, 1);
  while (true) {
    long long t0 = clock64();
    const Scratch<T>* img = reinterpret_cast<const Scratch<T>*>(ctx + (size_t)it.env * ctx_stride<T>());
    const int ga = gm.pair[it.ci][0], gb = gm.pair[it.ci][1];
    const int body = g.lane < 12 ? gm.body[ga] : gm.body[gb], k = g.lane < 12 ? g.lane : g.lane - 12;
    const bool mine = g.lane < 24 && body >= 0;
    T v = 0;
    if (mine) v = k < 3 ? img->bpos[body][k] : img->bR[body][k - 3];
    // next item (its index was claimed during the previous pair) and the claim after it
    int inext = count;
    CvxItem itn = it;
    if (ahead) {
      inext = __shfl_sync(0xffffffffu, raw_next, 0);
      if (inext < count) {
        itn = q.items[inext];
        if (g.lane == 0) raw_next = atomicAdd(q.head, 1);
      }
    }
    if (mine) { if (k < 3) cs.bpos[body][k] = v; else cs.bR[body][k - 3] = v; }
    __syncwarp();
    convex_pair<T, 32>(g, cs.bpos, cs.bR, gm, it.ci, em, q.res + i);
    if (g.lane == 0 && (p.work || p.cycles)) {
      long long dt = clock64() - t0;
      int e = it.env;
      if (p.work) atomicAdd(p.work + e, (int)(dt >> 8));
      if (p.cycles) { atomicAdd((unsigned long long*)p.cycles + 9 * e + 2, (unsigned long long)dt); }
    }
    if (!ahead) {  // short queue: claim the next pair only now
      if (g.lane == 0) inext = atomicAdd(q.head, 1);
      inext = __shfl_sync(0xffffffffu, inext, 0);
      if (inext < count) itn = q.items[inext];
    }
    if (inext >= count) break;
    i = inext;
    it = itn;
  }
}

template <class T, int G>
struct FusedCfg {  // reset / engine-level ops: the fused forward of an env by its own group
  static constexpr int W = warps_c<G>();
  static constexpr int THREADS = 32 * W;
  static constexpr int ENVS = THREADS / G;
};

template <class T, int G>
__global__ void __launch_bounds__(FusedCfg<T, G>::THREADS, MM_MINB_C) k_reset(StepParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  Grp<G> g;
  setup_group<T, G>(g);
  const ModelDev<T>* md = reinterpret_cast<const ModelDev<T>*>(p.model);
  int gi = threadIdx.x / G;
  long e = (long)blockIdx.x * FusedCfg<T, G>::ENVS + gi;
  if (e < p.n && !(p.mask && !p.mask[e])) {
    Scratch<T>& s = *reinterpret_cast<Scratch<T>*>(smem + gi * scratch_c_bytes<T>());
    Work<T> w = work_of<T>(p, e);
    env_reset<T, G>(g, s, *md, w, p.st, e, p.obj_xy ? p.obj_xy + 6 * e : nullptr, p.yaw_cs ? p.yaw_cs + 6 * e : nullptr, p.task[2 * e],
                    p.task[2 * e + 1], p.obs, p.tgt_kp);
  }
}

template <class T, int G>
__global__ void __launch_bounds__(FusedCfg<T, G>::THREADS, MM_MINB_C) k_ops(StepParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  Grp<G> g;
  setup_group<T, G>(g);
  const ModelDev<T>* md = reinterpret_cast<const ModelDev<T>*>(p.model);
  int gi = threadIdx.x / G;
  long e = (long)blockIdx.x * FusedCfg<T, G>::ENVS + gi;
  if (e < p.n) {
    Scratch<T>& s = *reinterpret_cast<Scratch<T>*>(smem + gi * scratch_c_bytes<T>());
    Work<T> w = work_of<T>(p, e);
    env_ops<T, G>(g, s, *md, w, p.st, e, p.ops, p.target);
  }
}

inline long env_long(const char* name, long dflt) {
  const char* e = getenv(name);
  return e ? atol(e) : dflt;
}

template <class T, int G> size_t smem_a() { return (32 * MM_WA / G) * scratch_a_bytes<T>(); }
template <class T, int G> size_t smem_c() { return (32 * warps_c<G>() / G) * scratch_c_bytes<T>(); }
template <class T> size_t smem_x() { return MM_WX * sizeof(ConvexSmem<T>); }

template <class T, int G>
cudaError_t inst_prepare() {
  cudaError_t e = cudaFuncSetAttribute(k_stage_a<T, G, MM_WA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a<T, G>());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_stage_c<T, G, warps_c<G>(), false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c<T, G>());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_stage_c<T, G, warps_c<G>(), true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c<T, G>());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_convex<T, MM_WX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x<T>());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_finish<T, G, MM_WA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a<T, G>());
  if (e != cudaSuccess) return e;
  if (G == 32) {
    e = cudaFuncSetAttribute(k_stage_c_heavy<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h<T>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_stage_c_heavy<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h<T>());
    if (e != cudaSuccess) return e;
  }
  e = cudaFuncSetAttribute(k_ops<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c<T, G>());
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_reset<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c<T, G>());
}

// grids of the persistent kernels of this instantiation (convex stage, contact-rich stage C): every CTA that can be resident
template <class T, int G>
cudaError_t inst_resident(int* convex_grid, int* heavy_grid) {
  int dev = 0, sms = 0, per = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_convex<T, MM_WX>, 32 * MM_WX, smem_x<T>());
  if (e != cudaSuccess) return e;
  *convex_grid = (per > 0 ? per : 1) * sms;
  *heavy_grid = 0;
  if (G == 32) {
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_stage_c_heavy<T, false>, 128, smem_h<T>());
    if (e != cudaSuccess) return e;
    *heavy_grid = (per > 0 ? per : 1) * sms;
  }
  return cudaSuccess;
}

// which: 0 = stage A of round `sub`, 1 = convex stage, 2 = stage C, 3 = reset, 4 = ops, 5 = stage C of the
// contact-rich envs; 6 / 7 = 2 / 5 with stage A of the next round fused behind; 8 = epilogue of the step.  `grid_x`: grid of the persistent kernels.
template <class T, int G>
cudaError_t inst_launch(int which, const StepParams& p, int sub, int grid_x, cudaStream_t s) {
  if (which == 0) {
    constexpr int EPB = 32 * MM_WA / G;
    k_stage_a<T, G, MM_WA><<<(unsigned)((p.nslot + EPB - 1) / EPB), 32 * MM_WA, smem_a<T, G>(), s>>>(p, sub);
  } else if (which == 8) {
    constexpr int EPB = 32 * MM_WA / G;
    k_finish<T, G, MM_WA><<<(unsigned)((p.nslot + EPB - 1) / EPB), 32 * MM_WA, smem_a<T, G>(), s>>>(p);
  } else if (which == 1) {
    k_convex<T, MM_WX><<<(unsigned)grid_x, 32 * MM_WX, smem_x<T>(), s>>>(p, sub);
  } else if (which == 2 || which == 6) {
    constexpr int EPB = 32 * warps_c<G>() / G;
    unsigned grid = (unsigned)((p.nslot + EPB - 1) / EPB);
    if (which == 2) k_stage_c<T, G, warps_c<G>(), false><<<grid, 32 * warps_c<G>(), smem_c<T, G>(), s>>>(p, sub);
    else k_stage_c<T, G, warps_c<G>(), true><<<grid, 32 * warps_c<G>(), smem_c<T, G>(), s>>>(p, sub);
  } else if (which == 5 || which == 7) {
    if (G == 32) {
      if (which == 5) k_stage_c_heavy<T, false><<<(unsigned)grid_x, 128, smem_h<T>(), s>>>(p, sub);
      else k_stage_c_heavy<T, true><<<(unsigned)grid_x, 128, smem_h<T>(), s>>>(p, sub);
    }
  } else {
    constexpr int EPB = FusedCfg<T, G>::ENVS;
    unsigned grid = (unsigned)((p.n + EPB - 1) / EPB);
    if (which == 3) k_reset<T, G><<<grid, FusedCfg<T, G>::THREADS, smem_c<T, G>(), s>>>(p);
    else k_ops<T, G><<<grid, FusedCfg<T, G>::THREADS, smem_c<T, G>(), s>>>(p);
  }
  return cudaGetLastError();
}

// entry points defined by the mm_inst_*.cu units
#define MM_DECL_INST(NAME)                 \
  cudaError_t prepare_##NAME();            \
  cudaError_t resident_##NAME(int* convex_grid, int* heavy_grid); \
  cudaError_t launch_##NAME(int which, const StepParams& p, int sub, int grid_x, cudaStream_t s);
MM_DECL_INST(f64_32) MM_DECL_INST(f64_16) MM_DECL_INST(f64_8)
MM_DECL_INST(f32_32) MM_DECL_INST(f32_16) MM_DECL_INST(f32_8)

#define MM_DEFINE_INST(NAME, T, G)                                                              \
  namespace mm {                                                                                \
  cudaError_t prepare_##NAME() { return inst_prepare<T, G>(); }                                 \
  cudaError_t resident_##NAME(int* convex_grid, int* heavy_grid) { return inst_resident<T, G>(convex_grid, heavy_grid); } \
  cudaError_t launch_##NAME(int which, const StepParams& p, int sub, int grid_x, cudaStream_t s) { return inst_launch<T, G>(which, p, sub, grid_x, s); } \
  }

}  // namespace mm
