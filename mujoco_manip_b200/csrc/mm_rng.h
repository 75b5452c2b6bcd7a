// Counter-based object placement (replaces the per-env numpy Generator of the reference,
// mujoco_manip/randomization.py:70-98 + gym_env.py:496-517, for the vectorised env).
//
// Philox4x32-10, key = 64-bit seed, counter = (global env id lo, hi, episode index, block):
//   attempt a of the rejection sampler uses blocks 4a+0..4a+2 (12 words -> six 53-bit uniforms:
//   x of the three cubes, then y of the three cubes - the reference's draw order), block 3 word 0
//   is the task draw; with randomize_yaw, cube o takes its yaw uniform from words 0,1 of block 4(o+1)+3.  The stream depends only on (seed, global env id, episode), never on how envs
//   are sharded over GPUs.  oracle/philox.py is the CPU statement of the same rule (bit-exact).
#pragma once
#include "mm_group.h"

namespace mm {

struct Philox {
  uint32_t k0, k1;
  MM_HD static uint32_t mulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
  MM_HD void block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t* out) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; r++) {
      uint32_t h0 = mulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
      uint32_t h1 = mulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
      uint32_t n0 = h1 ^ c1 ^ a, n1 = l1, n2 = h0 ^ c3 ^ b, n3 = l0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};

// no fused multiply-add here: the CPU statement uses separate IEEE multiply and add
MM_HD double affine_rn(double lo, double span, double u) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(lo, __dmul_rn(span, u));
#else
  volatile double p = span * u;
  return lo + p;
#endif
}
MM_HD double sq_sum_rn(double dx, double dy) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
#else
  volatile double a = dx * dx, b = dy * dy;
  return a + b;
#endif
}

MM_HD double u53(uint32_t hi, uint32_t lo) {
  return (double)(((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

// returns the number of attempts used (1..max_attempts), or 0 when every attempt was rejected
// (the reference raises RuntimeError there, randomization.py:84-87); xy = x0,y0,x1,y1,x2,y2
MM_HD int philox_place(uint64_t seed, uint64_t gid, uint32_t episode, double xlo, double xhi, double ylo, double yhi,
                       double min_sep, int max_attempts, double* xy) {
  Philox ph{(uint32_t)seed, (uint32_t)(seed >> 32)};
  const double sx = xhi - xlo, sy = yhi - ylo, ms2 = min_sep * min_sep;
  for (int a = 0; a < max_attempts; a++) {
    uint32_t w[12];
    for (int k = 0; k < 3; k++) ph.block((uint32_t)gid, (uint32_t)(gid >> 32), episode, (uint32_t)(4 * a + k), w + 4 * k);
    double x[3], y[3];
    for (int j = 0; j < 3; j++) {
      x[j] = affine_rn(xlo, sx, u53(w[2 * j], w[2 * j + 1]));
      y[j] = affine_rn(ylo, sy, u53(w[6 + 2 * j], w[6 + 2 * j + 1]));
    }
    bool ok = true;
    for (int i = 0; i < 3; i++)
      for (int j = i + 1; j < 3; j++)
        if (sq_sum_rn(x[i] - x[j], y[i] - y[j]) < ms2) ok = false;
    if (ok || a == max_attempts - 1) {
      for (int j = 0; j < 3; j++) { xy[2 * j] = x[j]; xy[2 * j + 1] = y[j]; }
      return ok ? a + 1 : 0;
    }
  }
  return 0;
}

MM_HD int philox_task(uint64_t seed, uint64_t gid, uint32_t episode, int npool) {
  Philox ph{(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t w[4];
  ph.block((uint32_t)gid, (uint32_t)(gid >> 32), episode, 3u, w);
  return (int)(((uint64_t)w[0] * (uint64_t)npool) >> 32);
}

// yaw of cube o (randomization.py:55-62: theta = uniform(0, 2 pi), drawn after the accepted placement)
MM_HD double philox_yaw(uint64_t seed, uint64_t gid, uint32_t episode, int o) {
  Philox ph{(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t w[4];
  ph.block((uint32_t)gid, (uint32_t)(gid >> 32), episode, (uint32_t)(4 * (o + 1) + 3), w);
  return affine_rn(0.0, 6.283185307179586, u53(w[0], w[1]));
}

}  // namespace mm
