// Lane-group abstraction: G lanes of a warp cooperate on ONE environment.
//   G = 32 : warp per env          G = 16 / 8 : two / four envs per warp
//   G = 1  : host build only (tests/ "1-lane emulation" of the very same kernel source, used to debug
//            the kernel logic against the oracle on machines without a GPU; never a product path).
// All loops over per-env items are written `for (i = g.lane; i < n; i += G)`, all reductions go
// through the group (xor-butterfly shuffles inside the aligned G-lane segment of the warp).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define MM_HD __host__ __device__ __forceinline__
#define MM_HDN __host__ __device__
#define MM_HDX __host__ __device__  // stage functions stay inline (out-of-line versions measured 10-15% slower)
#define MM_HDL __host__ __device__ __noinline__  // big leaf functions: ONE copy in the kernel (instruction-cache footprint)
// solver pieces that stage C calls from several places (update_constraint, mulJ, ...): -DMM_NOINLINE_SOLVER asks for
// one out-of-line copy each (instruction-cache experiment; the compiler clones most of them anyway: 235 -> 221 KB)
#ifdef MM_NOINLINE_SOLVER
#define MM_HDS __host__ __device__ __noinline__
#else
#define MM_HDS __host__ __device__
#endif
#else
#define MM_HD inline
#define MM_HDN
#define MM_HDX
#define MM_HDL
#define MM_HDS
#endif

// address-space hints for the out-of-line functions (a plain reference would compile to generic loads)
#ifdef __CUDA_ARCH__
#define MM_IN_SHARED(p) __builtin_assume(__isShared(p))
#define MM_IN_GLOBAL(p) __builtin_assume(__isGlobal(p))
#else
#define MM_IN_SHARED(p)
#define MM_IN_GLOBAL(p)
#endif

namespace mm {

template <int G>
struct Grp {
  int lane;       // 0..G-1 inside the group
  unsigned mask;  // lanes of this group inside the warp
  // G == 32: literal full mask, so that syncs and shuffles compile to single instructions (a run-time
  // mask makes the compiler emit a MATCH.ANY / vote / divergence-check sequence around every one of them)
  MM_HD unsigned m() const { return G == 32 ? 0xffffffffu : mask; }
  // warp-scope view of the group (a group never spans warps here; Grp<128> below does)
  static constexpr int WL = G;  // lanes of the group inside one warp
  MM_HD int wlane() const { return lane; }
  MM_HD int warp() const { return 0; }
  MM_HD static constexpr int nwarps() { return 1; }
  template <class T> MM_HD T wshfl_up(T v, int o) const { return shfl_up(v, o); }
  template <class T> MM_HD T wshfl_down(T v, int o) const { return shfl_down(v, o); }
  MM_HD int wany(int pred) const { return any(pred); }

  MM_HD void sync() const {
#ifdef __CUDA_ARCH__
    if (G > 1) __syncwarp(m());
#endif
  }
  template <class T>
  MM_HD T sum(T v) const {
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m(), v, o);
#endif
    return v;
  }
  MM_HD int any(int pred) const {
#ifdef __CUDA_ARCH__
    if (G > 1) return (__ballot_sync(m(), pred) & mask) != 0;
#endif
    return pred != 0;
  }
  // bit i set when lane i of the group has pred != 0
  MM_HD unsigned ballot(int pred) const {
#ifdef __CUDA_ARCH__
    if (G > 1) {
      unsigned b = __ballot_sync(m(), pred) & mask;
      return G == 32 ? b : (b >> (__ffs(m()) - 1));
    }
#endif
    return pred ? 1u : 0u;
  }
  // (value, index) reductions with the tie rule of a linear scan: among equal values the LOWEST index wins.
  // Full warp (G == 32): three integer warp-reduce instructions (redux.sync) on an order-preserving integer image of
  // the value - high word, low word among the lanes that hold the best high word, then the lowest index among the
  // exact matches - instead of a five-level shuffle butterfly; same result bit for bit (-0 and +0 compare equal in
  // both forms: the value is normalised with + 0.0 first).
#ifdef __CUDA_ARCH__
  __device__ __forceinline__ static unsigned long long order_key(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v + 0.0);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
  }
  __device__ __forceinline__ static unsigned long long order_key(float v) {
    unsigned u = (unsigned)__float_as_int(v + 0.0f);
    return (unsigned long long)((u >> 31) ? ~u : (u | 0x80000000u)) << 32;
  }
  template <class T>
  __device__ __forceinline__ void arg_redux(T& v, int& i, bool want_max, bool want_value = true) const {
    unsigned long long k = order_key(v);
    if (!want_max) k = ~k;
    unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    bool c = hi == mh;
    unsigned ml = __reduce_max_sync(0xffffffffu, c ? lo : 0u);
    c = c && lo == ml;
    int win = __reduce_min_sync(0xffffffffu, c ? i : 0x7fffffff);
    if (want_value) {  // (warp votes and shuffles are never dead code to the compiler: callers that only need the index say so)
      int src = __ffs(__ballot_sync(0xffffffffu, c && i == win)) - 1;
      v = __shfl_sync(0xffffffffu, v, src);
    }
    i = win;
  }
#endif
  // index of the maximum / minimum only (lowest index among equals); `v` is left unspecified
  template <class T>
  MM_HD void argmax_index(T& v, int& i) const {
#ifdef __CUDA_ARCH__
    if (G == 32) { arg_redux(v, i, true, false); return; }
#endif
    argmax(v, i);
  }
  template <class T>
  MM_HD void argmin_index(T& v, int& i) const {
#ifdef __CUDA_ARCH__
    if (G == 32) { arg_redux(v, i, false, false); return; }
#endif
    argmin(v, i);
  }
  template <class T>
  MM_HD void argmax(T& v, int& i) const {
#ifdef __CUDA_ARCH__
    if (G == 32) { arg_redux(v, i, true); return; }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      T ov = __shfl_xor_sync(m(), v, o);
      int oi = __shfl_xor_sync(m(), i, o);
      if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
#endif
  }
  template <class T>
  MM_HD void argmin(T& v, int& i) const {
#ifdef __CUDA_ARCH__
    if (G == 32) { arg_redux(v, i, false); return; }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      T ov = __shfl_xor_sync(m(), v, o);
      int oi = __shfl_xor_sync(m(), i, o);
      if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
#endif
  }
  MM_HD int imin(int v) const {
#ifdef __CUDA_ARCH__
    if (G == 32) return __reduce_min_sync(0xffffffffu, v);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) { int t = __shfl_xor_sync(m(), v, o); v = t < v ? t : v; }
#endif
    return v;
  }
  MM_HD int isum(int v) const {
#ifdef __CUDA_ARCH__
    if (G == 32) return __reduce_add_sync(0xffffffffu, v);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m(), v, o);
#endif
    return v;
  }
  MM_HD int imax(int v) const {
#ifdef __CUDA_ARCH__
    if (G == 32) return __reduce_max_sync(0xffffffffu, v);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) { int t = __shfl_xor_sync(m(), v, o); v = t > v ? t : v; }
#endif
    return v;
  }
  // exclusive prefix sum over the lanes of the group; *total receives the group sum
  MM_HD int scan_excl(int v, int* total) const {
    int incl = v;
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
      int t = __shfl_up_sync(m(), incl, o, G);
      if (lane >= o) incl += t;
    }
    *total = __shfl_sync(m(), incl, G - 1, G);
#else
    *total = incl;
#endif
    return incl - v;
  }
  // the same for 0 / 1 flags: one ballot and two population counts instead of a five-level shuffle scan
  MM_HD int scan_flag(int flag, int* total) const {
#ifdef __CUDA_ARCH__
    if (G > 1) {
      const unsigned b = ballot(flag);
      *total = __popc(b);
      return __popc(b & ((1u << lane) - 1u));
    }
#endif
    *total = flag ? 1 : 0;
    return 0;
  }
  template <class T>
  MM_HD T shfl_up(T v, int o) const {
#ifdef __CUDA_ARCH__
    return __shfl_up_sync(m(), v, o, G);
#else
    return v;
#endif
  }
  template <class T>
  MM_HD T shfl_down(T v, int o) const {
#ifdef __CUDA_ARCH__
    return __shfl_down_sync(m(), v, o, G);
#else
    return v;
#endif
  }
  template <class T>
  MM_HD T bcast(T v, int src) const {
#ifdef __CUDA_ARCH__
    return __shfl_sync(m(), v, src, G);
#else
    return v;
#endif
  }
};

// A whole 128-thread CTA on ONE environment (stage C of contact-rich envs, mm_launch.cuh k_stage_c_heavy): loops over
// contacts / rows / matrix entries spread over four warps, group syncs are CTA barriers, reductions go through a small
// shared scratch (fixed combination order: deterministic).  Only what stage C uses is provided.
#if defined(__CUDACC__)
template <>
struct Grp<128> {
  int lane;
  unsigned mask;
  double* xs;        // shared scratch of the CTA: [2][4] reduction slots, then [128] exchange slots
  mutable int par;   // alternating reduction slot set (every thread calls the reductions in the same order)
  static constexpr int WL = 32;
  __device__ __forceinline__ unsigned m() const { return 0xffffffffu; }
  __device__ __forceinline__ int wlane() const { return lane & 31; }
  __device__ __forceinline__ int warp() const { return lane >> 5; }
  __device__ __forceinline__ static constexpr int nwarps() { return 4; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  template <class T> __device__ __forceinline__ T wshfl_up(T v, int o) const { return __shfl_up_sync(0xffffffffu, v, o); }
  template <class T> __device__ __forceinline__ T wshfl_down(T v, int o) const { return __shfl_down_sync(0xffffffffu, v, o); }
  __device__ __forceinline__ int wany(int pred) const { return __any_sync(0xffffffffu, pred); }
  __device__ __forceinline__ int any(int pred) const { return __syncthreads_or(pred) != 0; }
  template <class T>
  __device__ __forceinline__ T sum(T v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    T* slot = reinterpret_cast<T*>(xs + 4 * par);
    if ((lane & 31) == 0) slot[lane >> 5] = v;
    __syncthreads();
    T r = (slot[0] + slot[1]) + (slot[2] + slot[3]);
    par ^= 1;
    return r;
  }
  __device__ __forceinline__ int isum(int v) const { return sum<int>(v); }
  __device__ __forceinline__ int bcast(int v, int src) const {
    int* slot = reinterpret_cast<int*>(xs + 4 * par);
    if (lane == src) slot[0] = v;
    __syncthreads();
    int r = slot[0];
    par ^= 1;
    return r;
  }
  __device__ __forceinline__ int scan_flag(int flag, int* total) const { return scan_excl(flag ? 1 : 0, total); }
  // exclusive prefix sum over the 128 lanes; *total receives the group sum
  __device__ __forceinline__ int scan_excl(int v, int* total) const {
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((lane & 31) >= o) incl += t;
    }
    int* slot = reinterpret_cast<int*>(xs + 4 * par);
    if ((lane & 31) == 31) slot[lane >> 5] = incl;
    __syncthreads();
    int w = lane >> 5, before = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { int t = slot[k]; tot += t; if (k < w) before += t; }
    par ^= 1;
    *total = tot;
    return before + incl - v;
  }
};
#endif

// ---- small vector helpers ----------------------------------------------------------------------
template <class T> MM_HD T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class T> MM_HD void cross3(T* r, const T* a, const T* b) {
  T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
template <class T> MM_HD T dot6(const T* a, const T* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
// r = R (row-major 3x3) * v
template <class T> MM_HD void rot(T* r, const T* R, const T* v) {
  T x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  T y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  T z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <class T> MM_HD void rotT(T* r, const T* R, const T* v) {
  T x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  T y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  T z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <class T> MM_HD void matmul3(T* r, const T* a, const T* b) {
  T t[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
#pragma unroll
  for (int i = 0; i < 9; i++) r[i] = t[i];
}
template <class T> MM_HD T tmax(T a, T b) { return a > b ? a : b; }
template <class T> MM_HD T tmin(T a, T b) { return a < b ? a : b; }
template <class T> MM_HD T tabs(T a) { return a < 0 ? -a : a; }
template <class T> MM_HD T tclamp(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }
MM_HD int tpopc(int x) {
#ifdef __CUDA_ARCH__
  return __popc((unsigned)x);
#else
  return __builtin_popcount((unsigned)x);
#endif
}
// index of the lowest set bit (x != 0)
MM_HD int tctz(unsigned x) {
#ifdef __CUDA_ARCH__
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}
MM_HD float tsqrt(float x) { return sqrtf(x); }
MM_HDL static double tsqrt(double x) { return sqrt(x); }  // ~150 SASS instructions per expansion: keep one copy
// Cholesky pivot: 1 / sqrt(d).  On the device one reciprocal square root (1 ulp) replaces the square root and the
// division; d >= MINVAL_D > 0 at every call.
MM_HD double trsqrt(double d) {
#ifdef __CUDA_ARCH__
  return rsqrt(d);
#else
  return 1.0 / sqrt(d);
#endif
}
MM_HD float trsqrt(float d) { return 1.0f / sqrtf(d); }
MM_HD void tsincos(float x, float* s, float* c) {
#ifdef __CUDA_ARCH__
  sincosf(x, s, c);
#else
  *s = sinf(x); *c = cosf(x);
#endif
}
MM_HD void tsincos(double x, double* s, double* c) {
#ifdef __CUDA_ARCH__
  sincos(x, s, c);
#else
  *s = sin(x); *c = cos(x);
#endif
}
MM_HD float tacos(float x) { return acosf(x); }
MM_HD double tacos(double x) { return acos(x); }
MM_HD float tsin(float x) { return sinf(x); }
MM_HD double tsin(double x) { return sin(x); }

}  // namespace mm
