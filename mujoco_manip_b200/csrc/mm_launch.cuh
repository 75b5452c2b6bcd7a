// Kernel templates of the step / reset launches and their per-(precision, group) entry points.
// Each (T, G) pair is instantiated in its own translation unit (mm_inst_*.cu) so the library builds
// in parallel; mm_kernels.cu holds the C ABI and dispatches through the mm_inst_* functions.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

#include "mm_env.h"

namespace mm {

// CTA shape.  The number of resident envs per SM is set by the shared-memory scratch (~17.5 KB per env in
// FP64).  G == 32: ONE CTA per SM holding as many warps (= envs) as fit, executed phase-synchronously
// (Grp::phase) so that its warps share the instruction cache.  G < 32: one warp per CTA (32 / G envs).
#ifndef MM_CTAS_PER_SM
#define MM_CTAS_PER_SM 2  // phase-synchronous G == 32 kernel: CTAs sharing an SM, so that the barrier waits of one
                          // overlap with the work of the other (measured: 1 -> 132k, 2 -> 151k, 3 -> 122k env-steps/s at 4096 envs)
#endif
#ifndef MM_WARPS_PER_CTA
#define MM_WARPS_PER_CTA 6
#endif
template <class T, int G>
struct BlockCfg {
  static constexpr int FIT = (int)((227 * 1024 - 2048) / sizeof(Scratch<T>));
  // 6 warps per CTA measured best in FP64 (7 fit after the scratch diet, but a 7-env CTA takes 40 % longer than a
  // 6-env one: more members to wait for at every barrier; 4096 envs: 166k vs 138k env-steps/s); the shared memory
  // left over serves as L1 for the workspace and the local-memory spills
  static constexpr int WFIT = (FIT > 16 ? 16 : FIT) / MM_CTAS_PER_SM;
  static constexpr int WARPS = G == 32 ? ((sizeof(T) == 8 && WFIT > MM_WARPS_PER_CTA) ? MM_WARPS_PER_CTA : WFIT) : 1;
  static constexpr int THREADS = 32 * WARPS;
  static constexpr int ENVS = THREADS / G;
  static constexpr int MINB = G == 32 ? MM_CTAS_PER_SM : ((227 * 1024) / (ENVS * (int)sizeof(Scratch<T>) + 1024) > 16
                                                 ? 16 : (227 * 1024) / (ENVS * (int)sizeof(Scratch<T>) + 1024));
};

struct StepParams {
  StatePtrs st;
  StepOut out;
  const float* actions;
  const void* model;
  void* work_reals;   // [pool_ctas * ENVS] per-env workspaces, handed out per RESIDENT CTA (see acquire_work)
  int* work_ints;
  int* pool_flags;    // [pool_ctas] 0 = free
  int pool_ctas;
  float* tgt_kp;
  const unsigned char* mask;
  const double* obj_xy;
  const double* yaw_cs;  // [N,6] cos(theta/2), sin(theta/2) of the three cubes, or null
  const int* task;
  float* obs;
  long n;
  int mode, reward_type, max_steps;
  const double* target;  // [N,3] world EE targets for the IK op, or null
  int ops;
  const int* order;      // [N] env processed by each slot (envs of similar cost share a CTA) or null = identity
  int* work;             // [N] out: busy cycles / 256 of each env's step, or null
  int phase_level;       // barrier density of the phase-synchronous G == 32 kernel (Grp::ps)
  long long* cycles;  // [N,9] or null: SM clock cycles of each env's step: total, then per stage (profiling aid)
};

template <class T, int G, int GPB = BlockCfg<T, G>::ENVS>
__device__ __forceinline__ bool setup(const StepParams& p, unsigned char* smem, const ModelDev<T>*& md, Scratch<T>*& sc,
                                      Grp<G>& g, long& e) {
  md = reinterpret_cast<const ModelDev<T>*>(p.model);  // read-only, global memory (L1 resident)
  int gi = threadIdx.x / G;
  e = (long)blockIdx.x * GPB + gi;
  g.busy = 0;
  g.mark = 0;
  sc = reinterpret_cast<Scratch<T>*>(smem) + gi;
  g.lane = threadIdx.x % G;
  int inwarp = (threadIdx.x % 32) / G;
  g.mask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (inwarp * G));
  g.ps = 0;
  return e < p.n;
}

// The per-env workspace (contacts, constraint rows, EPA vertices: ~47 KB) is scratch inside one launch, so it is
// pooled per RESIDENT CTA instead of per env: a few hundred CTAs are in flight however many envs there are, and
// their workspaces stay in L2 (N x 47 KB does not).  A CTA claims a free pool entry with a CAS scan that starts at
// its own index (the pool is at least as large as the number of CTAs that can be resident) and frees it on exit.
__device__ __forceinline__ int acquire_work(const StepParams& p) {
  __shared__ int s_slot;
  if (threadIdx.x == 0) {
    int slot;
    if ((int)gridDim.x <= p.pool_ctas) slot = (int)blockIdx.x;
    else {
      slot = (int)(blockIdx.x % (unsigned)p.pool_ctas);
      while (atomicCAS(p.pool_flags + slot, 0, 1) != 0) slot = slot + 1 == p.pool_ctas ? 0 : slot + 1;
    }
    s_slot = slot;
  }
  __syncthreads();
  return s_slot;
}
__device__ __forceinline__ void release_work(const StepParams& p, int slot) {
  __syncthreads();
  if (threadIdx.x == 0 && (int)gridDim.x > p.pool_ctas) { __threadfence(); atomicExch(p.pool_flags + slot, 0); }
}

// W = warps per CTA.  Two variants of the G == 32 kernel are built: W = BlockCfg::WARPS (6 in FP64) for large
// batches and W = MM_WARPS_SMALL (4) for small ones, where shorter CTAs fill the tail of the launch better
// (4096 envs: 186k vs 170k env-steps/s; 16384 envs: 228k vs 238k).
#ifndef MM_MINB_SMALL
#define MM_MINB_SMALL MM_CTAS_PER_SM  // CTAs per SM the short-CTA variant is compiled for (sets its register budget)
#endif
template <class T, int G, int W>
__global__ void __launch_bounds__(32 * W, W == BlockCfg<T, G>::WARPS ? BlockCfg<T, G>::MINB : MM_MINB_SMALL) k_step(StepParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ModelDev<T>* md;
  Scratch<T>* sc;
  Grp<G> g;
  long e;
  int pool = acquire_work(p);
  bool valid = setup<T, G, 32 * W / G>(p, smem, md, sc, g, e);
  bool dummy = false;
  if (G == 32) g.ps = p.phase_level;
  if (!valid) { e = p.n - 1; dummy = true; }  // padding group: replays the last env without storing
  if (p.order) e = p.order[e];
#ifdef __CUDA_ARCH__
  g.mark = clock64();
#endif
  long wslot = (long)pool * BlockCfg<T, G>::ENVS + threadIdx.x / G;
  Work<T> w = make_work(reinterpret_cast<T*>(p.work_reals) + wslot * WORK_REALS, p.work_ints + wslot * WORK_INTS);
  long long t0 = p.cycles ? clock64() : 0;
  env_step<T, G>(g, *sc, *md, w, p.st, e, p.actions, p.mode, p.reward_type, p.max_steps, p.out, p.tgt_kp, dummy,
                 p.cycles ? p.cycles + 9 * e : nullptr);
  if (p.cycles && g.lane == 0 && !dummy) p.cycles[9 * e] = clock64() - t0;
  if (p.work && g.lane == 0 && !dummy) p.work[e] = (int)((g.busy + (clock64() - g.mark)) >> 8);
  release_work(p, pool);
}

template <class T, int G>
__global__ void __launch_bounds__(BlockCfg<T, G>::THREADS, BlockCfg<T, G>::MINB) k_reset(StepParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ModelDev<T>* md;
  Scratch<T>* sc;
  Grp<G> g;
  long e;
  int pool = acquire_work(p);
  bool valid = setup<T, G>(p, smem, md, sc, g, e);
  if (valid && p.mask && !p.mask[e]) valid = false;
  if (valid) {
    long wslot = (long)pool * BlockCfg<T, G>::ENVS + threadIdx.x / G;
    Work<T> w = make_work(reinterpret_cast<T*>(p.work_reals) + wslot * WORK_REALS, p.work_ints + wslot * WORK_INTS);
    env_reset<T, G>(g, *sc, *md, w, p.st, e, p.obj_xy ? p.obj_xy + 6 * e : nullptr, p.yaw_cs ? p.yaw_cs + 6 * e : nullptr, p.task[2 * e],
                    p.task[2 * e + 1], p.obs, p.tgt_kp);
  }
  release_work(p, pool);
}


// MM_EXTRA_SMEM (bytes, environment variable) pads the dynamic shared memory: an occupancy probe for profiling only
inline size_t extra_smem() {
  static long v = -1;
  if (v < 0) { const char* e = getenv("MM_EXTRA_SMEM"); v = e ? atol(e) : 0; }
  return (size_t)v;
}
template <class T, int G>
__global__ void __launch_bounds__(BlockCfg<T, G>::THREADS, BlockCfg<T, G>::MINB) k_ops(StepParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ModelDev<T>* md;
  Scratch<T>* sc;
  Grp<G> g;
  long e;
  int pool = acquire_work(p);
  if (setup<T, G>(p, smem, md, sc, g, e)) {
    long wslot = (long)pool * BlockCfg<T, G>::ENVS + threadIdx.x / G;
    Work<T> w = make_work(reinterpret_cast<T*>(p.work_reals) + wslot * WORK_REALS, p.work_ints + wslot * WORK_INTS);
    env_ops<T, G>(g, *sc, *md, w, p.st, e, p.ops, p.target);
  }
  release_work(p, pool);
}

#ifndef MM_WARPS_SMALL
#define MM_WARPS_SMALL 4
#endif
#ifndef MM_SMALL_BATCH
#define MM_SMALL_BATCH 8192  // envs per GPU below which the short-CTA variant of the step kernel is launched
#endif
inline long small_batch() {  // MM_SMALL_BATCH (environment variable) overrides the compiled threshold: tuning aid
  static long v = -1;
  if (v < 0) { const char* e = getenv("MM_SMALL_BATCH"); v = e ? atol(e) : MM_SMALL_BATCH; }
  return v;
}
template <class T, int G>
constexpr int small_warps() { return G == 32 ? (BlockCfg<T, G>::WARPS < MM_WARPS_SMALL ? BlockCfg<T, G>::WARPS : MM_WARPS_SMALL) : 1; }

template <class T, int G>
size_t smem_bytes() { return BlockCfg<T, G>::ENVS * sizeof(Scratch<T>) + extra_smem(); }

template <class T, int G>
cudaError_t inst_prepare() {
  size_t sm = smem_bytes<T, G>();
  cudaError_t e = cudaFuncSetAttribute(k_step<T, G, BlockCfg<T, G>::WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_step<T, G, small_warps<T, G>()>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_ops<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_reset<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
}

// CTAs of this instantiation that can be resident on the device (sizes the workspace pool)
template <class T, int G>
cudaError_t inst_resident(int* ctas, int* envs_per_cta) {
  int dev = 0, sms = 0, per = 0, best = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  size_t sm = smem_bytes<T, G>();
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_step<T, G, BlockCfg<T, G>::WARPS>, BlockCfg<T, G>::THREADS, sm);
  if (e != cudaSuccess) return e;
  best = per;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_step<T, G, small_warps<T, G>()>, 32 * small_warps<T, G>(),
                                                    small_warps<T, G>() * sizeof(Scratch<T>) + extra_smem());
  if (e != cudaSuccess) return e;
  best = per > best ? per : best;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_reset<T, G>, BlockCfg<T, G>::THREADS, sm);
  if (e != cudaSuccess) return e;
  best = per > best ? per : best;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_ops<T, G>, BlockCfg<T, G>::THREADS, sm);
  if (e != cudaSuccess) return e;
  best = per > best ? per : best;
  *ctas = best * sms;
  *envs_per_cta = BlockCfg<T, G>::ENVS;
  return cudaSuccess;
}

template <class T, int G>
cudaError_t inst_launch(int which, const StepParams& p, cudaStream_t s) {  // which: 0 step, 1 reset, 2 ops
  constexpr int BLOCK = BlockCfg<T, G>::THREADS;
  constexpr int GPB = BLOCK / G;
  unsigned grid = (unsigned)((p.n + GPB - 1) / GPB);
  size_t sm = smem_bytes<T, G>();
  if (which == 1) k_reset<T, G><<<grid, BLOCK, sm, s>>>(p);
  else if (which == 2) k_ops<T, G><<<grid, BLOCK, sm, s>>>(p);
  else if (G == 32 && p.n < small_batch()) {
    constexpr int W = small_warps<T, G>();
    unsigned g2 = (unsigned)((p.n + W - 1) / W);
    k_step<T, G, W><<<g2, 32 * W, W * sizeof(Scratch<T>) + extra_smem(), s>>>(p);
  } else k_step<T, G, BlockCfg<T, G>::WARPS><<<grid, BLOCK, sm, s>>>(p);
  return cudaGetLastError();
}

// entry points defined by the mm_inst_*.cu units
#define MM_DECL_INST(NAME)                 \
  cudaError_t prepare_##NAME();            \
  cudaError_t resident_##NAME(int* ctas, int* envs_per_cta); \
  cudaError_t launch_##NAME(int which, const StepParams& p, cudaStream_t s);
MM_DECL_INST(f64_32) MM_DECL_INST(f64_16) MM_DECL_INST(f64_8)
MM_DECL_INST(f32_32) MM_DECL_INST(f32_16) MM_DECL_INST(f32_8)

#define MM_DEFINE_INST(NAME, T, G)                                                              \
  namespace mm {                                                                                \
  cudaError_t prepare_##NAME() { return inst_prepare<T, G>(); }                                 \
  cudaError_t resident_##NAME(int* ctas, int* envs_per_cta) { return inst_resident<T, G>(ctas, envs_per_cta); } \
  cudaError_t launch_##NAME(int which, const StepParams& p, cudaStream_t s) { return inst_launch<T, G>(which, p, s); } \
  }

}  // namespace mm
