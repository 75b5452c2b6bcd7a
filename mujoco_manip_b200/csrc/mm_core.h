// Physics substep of the pick-and-place scene for ONE environment handled by a G-lane group.
//
// Replaces, for this scene only, what the reference reaches through libmujoco on its step hot path:
//   mj_step    (mujoco_manip/env.py:119-121, gym_env.py:558)
//   mj_forward (mujoco_manip/env.py:117,161, gym_env.py:560)
//   mj_jac     (mujoco_manip/controller.py:101-108)
// Algorithms (SURVEY.md Appendix A for the semantics that are reproduced):
//   * kinematics of the 7R+gripper chain, world-origin spatial axes S_i for all 27 dofs
//   * CRBA with compact rigid-body inertias (m, h, Ibar) about the world origin -> 9x9 robot block;
//     the cube blocks are constant diag(m, I)
//   * RNEA in the same coordinates for the bias forces
//   * two-level broad phase, box-box / plane-box / plane-hull narrow phase + GJK / EPA for hulls and cylinders
//     (mm_ccd.h), contacts grouped by BODY PAIR
//   * pyramidal soft constraints solved by Newton with exact line search, MATRIX-FREE:
//     J is never stored; every J*x, J^T*f and J^T D J goes through per-body-pair 6-vectors / 6x6 blocks
//   * H = M + J^T D J assembled per touching body pair and factored by coupling structure (lone cubes: 6x6 on one lane)
//   * implicitfast integration (block diagonal: 9x9 robot factor, scalar cube updates)
// Execution: G = 32 lanes per env inside batch-wide stage kernels (mm_env.h, mm_launch.cuh); one warp per geom pair in the convex stage.
#pragma once
#include "mm_ccd.h"

namespace mm {

#ifdef __CUDA_ARCH__
#define MM_TICK_(s, g, slot, t0) do { if ((s).prof) { long long t1_ = clock64(); if ((g).lane == 0) (s).tph[slot] += (unsigned)((t1_ - (t0)) >> 6); (t0) = t1_; } } while (0)
#define MM_T0(s) ((s).prof ? clock64() : 0)
#else
#define MM_TICK_(s, g, slot, t0) do { } while (0)
#define MM_T0(s) 0
#endif
#define MM_TICK(s, g, slot, t0) MM_TICK_(s, g, slot, t0)

constexpr int MAXCON = 256;   // contacts per env (oracle max: 44 in scripted episodes, 76 in the table-collision stress run, 143 in random-action pile-ups)
constexpr int MAXROW = MAXCON * 6;
// MM_ROWS_S > 0: the solver rows Jaref and Jv (MM_ROWS_N = 3: aref too) of an env with at most MM_ROWS_S contacts live
// in shared memory instead of the global workspace: rows are re-read right after they are written, and every such
// read misses L1.  12 contacts (three cubes at rest on the table, 62 % of the benchmarked envs) x 2 arrays is what
// fits beside 12 resident envs per SM; taking the space for a third array from the shared pair tables (MM_PAIR_S)
// slows the contact-rich envs down, which the small batches wait for (profiles/r02_variants.txt).
#ifndef MM_ROWS_S
#define MM_ROWS_S 12
#endif
#ifndef MM_ROWS_N
#define MM_ROWS_N 2
#endif
#ifndef MM_PAIR_S
#define MM_PAIR_S 16
#endif
constexpr int ROWS_S = MM_ROWS_S;
#if MM_ROWS_S
#define MM_ROWS_GLOBAL(p)
#else
#define MM_ROWS_GLOBAL(p) MM_IN_GLOBAL(p)
#endif
#if MM_ROWS_S && MM_ROWS_N == 3
#define MM_AREF_GLOBAL(p)
#else
#define MM_AREF_GLOBAL(p) MM_IN_GLOBAL(p)
#endif
constexpr int MAXPAIR_S = MM_PAIR_S; // simultaneously touching body pairs whose tables live in shared memory (the common case)
#ifndef MM_PAIR_SPILL_AT
#define MM_PAIR_SPILL_AT MAXPAIR_S
#endif
constexpr int MAXPAIR = 96;   // every ordered (class, class) key of the 780 candidate pairs (94, tools/modelc.py): exact, no cap;
                              // envs with more than MAXPAIR_S touching pairs keep their pair tables in the global workspace
constexpr int MAXSPEC = 10;   // equality + at most one limit row per robot joint
constexpr int MAXSURV = 384;  // geom pairs surviving the first level of the broad phase
constexpr double MINVAL_D = 1e-15;

// meta word of a contact
//  bits 0-6 pair slot | 7-10 class A | 11-14 class B | 15 condim-4 (cube) | 16-21 active-row bits | 22 robot-obstacle
constexpr int META_KEY_SHIFT = 7, META_DIM4_BIT = 15, META_ACT_SHIFT = 16, META_ROBOBS_BIT = 22;
MM_HD int meta_slot(int m) { return m & 127; }
MM_HD int meta_dim4(int m) { return (m >> META_DIM4_BIT) & 1; }
MM_HD int meta_key(int m) { return (m >> META_KEY_SHIFT) & 0xFF; }  // class A | class B << 4
// candidate pairs (and therefore contacts) are ordered by (class A, class B)
MM_HD int sort_key(int m) { return (((m >> META_KEY_SHIFT) & 15) << 4) | ((m >> (META_KEY_SHIFT + 4)) & 15); }

// rows of Scratch::S: one spatial axis per dof (the cube translations carry their unit vectors, so that every product
// with S is a plain row access without a case split)
constexpr int NSROW = NV;
MM_HD int srow(int dof) { return dof; }
MM_HD bool is_cube_translation(int dof) { return dof >= NROB && ((dof - NROB) % 6) < 3; }
// rows of the 6-vector scratch: 16 suffice (14 dofs of a body pair + 2 rows that park the position-stage qpos);
// the FP32 build keeps 27 because its EPA workspace (integers are as wide as reals there) needs the room
template <class T> constexpr int TMP6_ROWS() { return sizeof(T) == 8 ? 16 : 48; }
constexpr int KIN_ROW = 14;  // tmp6 rows 14, 15: arm + finger qpos of the last position stage (store_state)

template <class T>
struct alignas(16) Scratch {  // (images travel as 16-byte words: ctx_copy)
  // ---- persistent part: the image of an env that travels between the stage kernels of one control step
  //      (mm_stage.h: ctx_load / ctx_store copy [0, SCRATCH_PERSIST) to / from global memory) ----
  T qpos[NQ], qvel[NV], ctrl[NU];
  T bpos[NDB][3], bR[NDB][9];
  T S[NSROW][6];   // spatial axis of every dof about the world origin (angular 3, linear 3)
  T Mr[NROB * NROB];
  T fs[NV], as[NV];
  T actf[NU];
  T target[3];
  int actsat[NU];
  int ncon, nbox, nsurv, overflow, niter, prof;
  int qbase, ncvx;  // this env's slice of the convex-pair queue: items / results [qbase, qbase + ncvx)
  unsigned tph[8];  // profiling (mm_set_cycle_buffer): busy cycles / 64 in kinematics+dynamics | broad phase | narrow
                    // phase | its convex (GJK / EPA) part | constraint rows | solver | IK | integration
  // ---- stage temporaries ----
  double* warm_g;  // qacc_warmstart of this env in the global state (read at the start of solve, written at its end)
  unsigned long long mbar;  // mbarrier of the bulk-async image load (ctx_load, device only)
  // contiguous block that is dead during collision (re-used there as clip scratch and EPA polytope)
  T H[NV * NV];
  T tmp6[TMP6_ROWS<T>()][6];
  T pairK_s[MAXPAIR_S][21], pairW_s[MAXPAIR_S][6], pairF_s[MAXPAIR_S][6];
  T rows_s[ROWS_S ? MM_ROWS_N * 6 * ROWS_S : 1];
  T qacc[NV], Ma[NV], search[NV], Mv[NV], fc[NV];
  // -----------------------------------------------------------------------------------------------------------
  T specD[MAXSPEC], specJaref[MAXSPEC], specJv[MAXSPEC], specAref[MAXSPEC];
  int pairkey_s[MAXPAIR_S];
  int pairmd_s[MAXPAIR_S], pairmb_s[MAXPAIR_S];  // dof masks of the pair: dofs of exactly one of the two bodies | dofs of body B
  // pair tables of this forward pass: the shared arrays above, or (more than MAXPAIR_S touching pairs) the global workspace
  T (*pairK)[21];
  T (*pairW)[6];
  T (*pairF)[6];
  int *pairkey, *pairmd, *pairmb;
  int specdof[MAXSPEC];  // -1: equality row (e7 - e8); else dof | (negative sign ? 256 : 0)
  int npair, nspec, hvalid;
  int lone;       // bit c: cube c touches neither the robot nor another cube -> its 6x6 block of H is independent
  int n_il;       // dofs of the coupled part: robot (9) + the cubes that are not `lone`
  signed char il[NV], dl[16];
};
#define SCRATCH_PERSIST(T) (offsetof(Scratch<T>, warm_g))

// result of the general convex test of one queued geom pair (mm_ccd.h), consumed in candidate order by assemble_contacts
template <class T>
struct CvxRes { T pos[3], nrm[3], depth; int hit, ci; };

// per-env slice of the global workspace (streamed, coalesced across lanes: index = contact / row)
template <class T>
struct Work {
  // contact list of the env (lives from the narrow phase to the end of the solver)
  T* cpos;  // [3][MAXCON]
  T* cn;    // [3][MAXCON]
  T* ct1;   // [3][MAXCON]
  T* cdist; // [MAXCON]
  T* cD;    // [MAXCON]
  int* cmeta;  // [MAXCON]
  int* surv;   // [MAXSURV] geom pairs surviving the broad phase
  // solver rows (alive inside stage C only) + spill space
  T* aref;  // [MAXROW]
  T* Jaref; // [MAXROW]
  T* Jv;    // [MAXROW]
  T* pairbig;    // [MAXPAIR][33] pair tables of an env with more than MAXPAIR_S touching body pairs
  int* pairbig_i;  // [3][MAXPAIR]
  CvxRes<T>* cvx;  // results of this env's convex pairs: its slice of the batch-wide queue (stage kernels), or - fused
                   // forward - the head of the row arrays, which are idle until the constraint rows are made
  EpaMem<T> epa;   // fused forward only: vertices at the tail of the row arrays (the convex kernel has its own polytope)
};
constexpr int WORK_REALS = MAXCON * 11 + MAXROW * 3 + MAXPAIR * 33;
constexpr int WORK_INTS = MAXCON + MAXSURV + 3 * MAXPAIR;
template <class T>
MM_HD Work<T> make_work(T* reals, int* ints) {
  Work<T> w;
  w.cpos = reals; w.cn = reals + 3 * MAXCON; w.ct1 = reals + 6 * MAXCON; w.cdist = reals + 9 * MAXCON;
  w.cD = reals + 10 * MAXCON;
  w.aref = reals + 11 * MAXCON; w.Jaref = w.aref + MAXROW; w.Jv = w.Jaref + MAXROW;
  w.pairbig = w.Jv + MAXROW;
  w.cmeta = ints; w.surv = ints + MAXCON; w.pairbig_i = w.surv + MAXSURV;
  w.cvx = reinterpret_cast<CvxRes<T>*>(w.aref);
  w.epa.vert = w.aref + 3 * MAXROW - EPA_MAXV * 6;
  w.epa.face = nullptr; w.epa.fidx = nullptr; w.epa.edge = nullptr; w.epa.canon = nullptr; w.epa.ecan = nullptr;  // shared memory (collide)
  static_assert(MAXSURV * sizeof(CvxRes<T>) + EPA_MAXV * 6 * sizeof(T) <= 3 * MAXROW * sizeof(T), "fused convex scratch does not fit the row arrays");
  static_assert((MAXCON * 11) % 2 == 0 && WORK_REALS % 2 == 0, "8-byte alignment of the slices");
  return w;
}

MM_HD int dofmask(int cls) {
  if (cls == 0) return 0;
  if (cls <= 7) return (1 << cls) - 1;
  if (cls == 8) return 0x7F | (1 << 7);
  if (cls == 9) return 0x7F | (1 << 8);
  return 0x3F << (9 + 6 * (cls - CLS_CUBE0));
}

// ------------------------------------------------------------------------------------------------
// group Cholesky (lower, in place, row-major ld = n) and solve.  Every factor routine below leaves the RECIPROCAL
// of the pivot on the diagonal: the triangular solves multiply instead of dividing on their sequential chain.
// ------------------------------------------------------------------------------------------------
template <class T, int G>
MM_HDS void chol_factor(const Grp<G>& g, T* A, int n) {
  for (int j = 0; j < n; j++) {
    T d = A[j * n + j];
    if (d < (T)MINVAL_D) d = (T)MINVAL_D;
    T inv = trsqrt(d);
    g.sync();
    for (int i = j + 1 + g.lane; i < n; i += G) A[i * n + j] *= inv;
    if (g.lane == 0) A[j * n + j] = inv;
    g.sync();
    // trailing update: one lane per row, or 2 / 4 lanes per row (each a contiguous share of the row) while the
    // rows left are few - every entry is still updated once per pivot by exactly one lane, so the result is the same
    int m = n - j - 1;
    int P = G >= 4 * m ? 4 : (G >= 2 * m ? 2 : 1);
    if (G < 32) P = 1;
    if (P == 1) {
      for (int i = j + 1 + g.lane; i < n; i += G) {
        T lij = A[i * n + j];
        for (int k = j + 1; k <= i; k++) A[i * n + k] -= lij * A[k * n + j];
      }
    } else {
      int r = g.lane / P, h = g.lane % P;
      if (r < m) {
        int i = j + 1 + r, len = i - j;
        T lij = A[i * n + j];
        int k0 = j + 1 + (len * h) / P, k1 = j + 1 + (len * (h + 1)) / P;
        for (int k = k0; k < k1; k++) A[i * n + k] -= lij * A[k * n + j];
      }
    }
    g.sync();
  }
}

template <class T, int G>
MM_HDS void chol_solve(const Grp<G>& g, const T* L, int n, T* x) {
  for (int k = 0; k < n; k++) {
    T xk = x[k] * L[k * n + k];
    g.sync();
    if (g.lane == 0) x[k] = xk;
    for (int i = k + 1 + g.lane; i < n; i += G) x[i] -= L[i * n + k] * xk;
    g.sync();
  }
  for (int k = n - 1; k >= 0; k--) {
    T xk = x[k] * L[k * n + k];
    g.sync();
    if (g.lane == 0) x[k] = xk;
    for (int i = g.lane; i < k; i += G) x[i] -= L[k * n + i] * xk;
    g.sync();
  }
}

// ---- structure-aware factorisation of H = M + J^T D J --------------------------------------------
// H couples the robot block (9 dofs) with a cube block (6 dofs) only while a robot geom touches that
// cube, and two cube blocks only while the cubes touch.  A `lone` cube is factored / solved by ONE lane
// in registers (6x6); the coupled part (robot + non-lone cubes, index list s.il) cooperatively.  Entries
// outside the coupling pattern stay exactly zero and are skipped.
template <class T>
MM_HD void chol6_local(T* A) {  // A -> H[b][b], row stride NV; lower triangle in place
  T a[21];
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) a[i * (i + 1) / 2 + j] = A[i * NV + j];
#pragma unroll
  for (int j = 0; j < 6; j++) {
    T d = a[j * (j + 1) / 2 + j];
    if (d < (T)MINVAL_D) d = (T)MINVAL_D;
    T inv = trsqrt(d);
    a[j * (j + 1) / 2 + j] = inv;
#pragma unroll
    for (int i = j + 1; i < 6; i++) a[i * (i + 1) / 2 + j] *= inv;
#pragma unroll
    for (int i = j + 1; i < 6; i++)
#pragma unroll
      for (int k = j + 1; k <= i; k++) a[i * (i + 1) / 2 + k] -= a[i * (i + 1) / 2 + j] * a[k * (k + 1) / 2 + j];
  }
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) A[i * NV + j] = a[i * (i + 1) / 2 + j];
}

template <class T>
MM_HD void solve6_local(const T* L, T* x) {  // L -> H[b][b] (factor), x -> vector + b
  T v[6];
#pragma unroll
  for (int i = 0; i < 6; i++) v[i] = x[i];
#pragma unroll
  for (int i = 0; i < 6; i++) {
#pragma unroll
    for (int k = 0; k < i; k++) v[i] -= L[i * NV + k] * v[k];
    v[i] *= L[i * NV + i];
  }
#pragma unroll
  for (int i = 5; i >= 0; i--) {
#pragma unroll
    for (int k = i + 1; k < 6; k++) v[i] -= L[k * NV + i] * v[k];
    v[i] *= L[i * NV + i];
  }
#pragma unroll
  for (int i = 0; i < 6; i++) x[i] = v[i];
}

// cooperative Cholesky of the rows / columns listed in il[0..n) of the NV x NV matrix A (zero entries skipped)
template <class T, int G>
MM_HDS void chol_factor_list(const Grp<G>& g, T* A, const signed char* il, int n) {
  for (int jj = 0; jj < n; jj++) {
    int j = il[jj];
    T d = A[j * NV + j];
    if (d < (T)MINVAL_D) d = (T)MINVAL_D;
    T inv = trsqrt(d);
    g.sync();
    for (int ii = jj + 1 + g.lane; ii < n; ii += G) A[il[ii] * NV + j] *= inv;
    if (g.lane == 0) A[j * NV + j] = inv;
    g.sync();
    int m = n - jj - 1;  // trailing update, 1 / 2 / 4 lanes per row as in chol_factor
    int P = G >= 4 * m ? 4 : (G >= 2 * m ? 2 : 1);
    if (G < 32) P = 1;
    if (P == 1) {
      for (int ii = jj + 1 + g.lane; ii < n; ii += G) {
        int i = il[ii];
        T lij = A[i * NV + j];
        if (lij == 0) continue;
        for (int kk = jj + 1; kk <= ii; kk++) { int k = il[kk]; A[i * NV + k] -= lij * A[k * NV + j]; }
      }
    } else {
      int r = g.lane / P, h = g.lane % P;
      if (r < m) {
        int ii = jj + 1 + r, len = ii - jj;
        int i = il[ii];
        T lij = A[i * NV + j];
        if (lij != 0) {
          int c0 = jj + 1 + (len * h) / P, c1 = jj + 1 + (len * (h + 1)) / P;
          for (int kk = c0; kk < c1; kk++) { int k = il[kk]; A[i * NV + k] -= lij * A[k * NV + j]; }
        }
      }
    }
    g.sync();
  }
}

// H x = b for the structured factor: lone cubes on lanes 0..2, the coupled part sequentially on lane 3
// when it is just the robot block, cooperatively otherwise
template <class T, int G>
MM_HDS void solve_H(const Grp<G>& g, const Scratch<T>& s, T* x) {
  const T* L = s.H;
  int n = s.n_il;
  if (G >= 4 || G == 1) {
    for (int c = (G == 1 ? 0 : g.lane); c < 3; c += (G == 1 ? 1 : G))
      if ((s.lone >> c) & 1) solve6_local(L + (9 + 6 * c) * (NV + 1), x + 9 + 6 * c);
  }
  if (n == NROB) {
    if (g.lane == (G >= 4 ? 3 : 0)) {
      for (int i = 0; i < NROB; i++) {
        T v = x[i];
        for (int k = 0; k < i; k++) v -= L[i * NV + k] * x[k];
        x[i] = v * L[i * NV + i];
      }
      for (int i = NROB - 1; i >= 0; i--) {
        T v = x[i];
        for (int k = i + 1; k < NROB; k++) v -= L[k * NV + i] * x[k];
        x[i] = v * L[i * NV + i];
      }
    }
    g.sync();
    return;
  }
  const signed char* il = s.il;
  for (int kk = 0; kk < n; kk++) {
    int k = il[kk];
    T xk = x[k] * L[k * NV + k];
    g.sync();
    if (g.lane == 0) x[k] = xk;
    for (int ii = kk + 1 + g.lane; ii < n; ii += G) { int i = il[ii]; x[i] -= L[i * NV + k] * xk; }
    g.sync();
  }
  for (int kk = n - 1; kk >= 0; kk--) {
    int k = il[kk];
    T xk = x[k] * L[k * NV + k];
    g.sync();
    if (g.lane == 0) x[k] = xk;
    for (int ii = g.lane; ii < kk; ii += G) { int i = il[ii]; x[i] -= L[k * NV + i] * xk; }
    g.sync();
  }
}

template <class T>
MM_HD T S_comp(const Scratch<T>& s, int dof, int c) {
  return s.S[dof][c];
}
template <class T>
MM_HD T S_dot(const Scratch<T>& s, int dof, const T* x) {
  return dot6(s.S[dof], x);
}
template <class T>
MM_HD void S_get(const Scratch<T>& s, int dof, T* out) {
  const T* S = s.S[dof];
  for (int a = 0; a < 6; a++) out[a] = S[a];
}

// ------------------------------------------------------------------------------------------------
// kinematics + world-origin spatial axes (replaces mj_kinematics / mj_comPos / mj_jac data)
// ------------------------------------------------------------------------------------------------
template <class T>
MM_HD void quat2mat(T* m, const T* q) {
  T w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}

template <class T, int G>
MM_HDX void fk(const Grp<G>& g, Scratch<T>& s, const ModelDev<T>& md) {
  MM_IN_SHARED(&s);
  MM_IN_GLOBAL(&md);
  for (int i = g.lane; i < NARM; i += G) tsincos(s.qpos[i], &s.tmp6[i][0], &s.tmp6[i][1]);
  g.sync();
  for (int t = g.lane; t < 4; t += G) {
    if (t == 0) {
      T Rp[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pp[3] = {0, 0, 0};
      for (int k = 0; k < NARM; k++) {
        T v[3], Rf[9];
        rot(v, Rp, md.link_pos[k]);
        for (int a = 0; a < 3; a++) { pp[a] += v[a]; s.bpos[k][a] = pp[a]; }
        matmul3(Rf, Rp, md.link_R[k]);
        T sn = s.tmp6[k][0], cs = s.tmp6[k][1];
        for (int r = 0; r < 3; r++) {
          Rp[3 * r] = Rf[3 * r] * cs + Rf[3 * r + 1] * sn;
          Rp[3 * r + 1] = Rf[3 * r + 1] * cs - Rf[3 * r] * sn;
          Rp[3 * r + 2] = Rf[3 * r + 2];
        }
        for (int a = 0; a < 9; a++) s.bR[k][a] = Rp[a];
      }
      T v[3], Rh[9], ph[3];
      rot(v, Rp, md.link_pos[7]);
      for (int a = 0; a < 3; a++) { ph[a] = pp[a] + v[a]; s.bpos[DB_HAND][a] = ph[a]; }
      matmul3(Rh, Rp, md.link_R[7]);
      for (int a = 0; a < 9; a++) s.bR[DB_HAND][a] = Rh[a];
      for (int f = 0; f < 2; f++) {
        T Rf[9];
        matmul3(Rf, Rh, md.link_R[8 + f]);
        rot(v, Rh, md.link_pos[8 + f]);
        T q = s.qpos[7 + f];
        for (int a = 0; a < 3; a++) s.bpos[DB_LF + f][a] = ph[a] + v[a] + Rf[3 * a + 1] * q;
        for (int a = 0; a < 9; a++) s.bR[DB_LF + f][a] = Rf[a];
      }
    } else {
      int j = t - 1;
      T* q = s.qpos + 9 + 7 * j + 3;
      T nn = tsqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
      if (nn < (T)MINVAL_D) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
      else { T inv = (T)1 / nn; for (int a = 0; a < 4; a++) q[a] *= inv; }
      for (int a = 0; a < 3; a++) s.bpos[DB_CUBE0 + j][a] = s.qpos[9 + 7 * j + a];
      quat2mat(s.bR[DB_CUBE0 + j], q);
    }
  }
  g.sync();
  for (int i = g.lane; i < NV; i += G) {
    T* S = s.S[i];
    if (is_cube_translation(i)) {
      int k = (i - NROB) % 6;
#pragma unroll
      for (int a = 0; a < 6; a++) S[a] = a == 3 + k ? (T)1 : (T)0;
      continue;
    }
    if (i < NARM) {
      T a[3] = {s.bR[i][2], s.bR[i][5], s.bR[i][8]};
      S[0] = a[0]; S[1] = a[1]; S[2] = a[2];
      cross3(S + 3, s.bpos[i], a);
    } else if (i < NROB) {
      const T* R = s.bR[DB_LF + (i - 7)];
      S[0] = S[1] = S[2] = 0; S[3] = R[1]; S[4] = R[4]; S[5] = R[7];
    } else {
      int d = i - 9, j = d / 6, k = d % 6;
      const T* R = s.bR[DB_CUBE0 + j];
      T a[3] = {R[k - 3], R[3 + k - 3], R[6 + k - 3]};
      S[0] = a[0]; S[1] = a[1]; S[2] = a[2];
      cross3(S + 3, s.bpos[DB_CUBE0 + j], a);
    }
  }
  g.sync();
}

// ------------------------------------------------------------------------------------------------
// smooth dynamics: robot mass matrix (CRBA), bias (RNEA), passive + actuation, qacc_smooth (A5)
// ------------------------------------------------------------------------------------------------
// compact inertia about the world origin: {m, h[3], Ixx, Iyy, Izz, Ixy, Ixz, Iyz};  (n;f) = I (w;v)
template <class T>
MM_HD void inertia_apply(const T* I, const T* V, T* F) {
  const T* w = V; const T* v = V + 3; const T* h = I + 1;
  T hv[3], hw[3];
  cross3(hv, h, v);
  cross3(hw, h, w);
  F[0] = I[4] * w[0] + I[7] * w[1] + I[8] * w[2] + hv[0];
  F[1] = I[7] * w[0] + I[5] * w[1] + I[9] * w[2] + hv[1];
  F[2] = I[8] * w[0] + I[9] * w[1] + I[6] * w[2] + hv[2];
  F[3] = I[0] * v[0] - hw[0]; F[4] = I[0] * v[1] - hw[1]; F[5] = I[0] * v[2] - hw[2];
}

MM_HD int ib_parent(int k) { return k <= 6 ? k - 1 : 6; }

template <class T, int G>
MM_HDX void dyn_smooth(const Grp<G>& g, Scratch<T>& s, const ModelDev<T>& md) {
  MM_IN_SHARED(&s);
  MM_IN_GLOBAL(&md);
  // temporaries live in the (not yet needed) H region: [2][9][10] individual / composite inertias
  T (*inert)[NROB][10] = reinterpret_cast<T (*)[NROB][10]>(s.H + 300);
  // individual compact inertias
  for (int ib = g.lane; ib < NROB; ib += G) {
    int d = ib < 7 ? ib : ib + 1;
    const T* R = s.bR[d];
    T c[3], v[3];
    rot(v, R, md.ib_com[ib]);
    for (int a = 0; a < 3; a++) c[a] = s.bpos[d][a] + v[a];
    const T* J = md.ib_inertia[ib];
    T Ib[9] = {J[0], J[3], J[4], J[3], J[1], J[5], J[4], J[5], J[2]}, t[9], Iw[9], Rt[9];
    for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) Rt[3 * i + k] = R[3 * k + i];
    matmul3(t, R, Ib);
    matmul3(Iw, t, Rt);
    T m = md.ib_mass[ib], cc = dot3(c, c);
    T* o = inert[0][ib];
    o[0] = m; o[1] = m * c[0]; o[2] = m * c[1]; o[3] = m * c[2];
    o[4] = Iw[0] + m * (cc - c[0] * c[0]); o[5] = Iw[4] + m * (cc - c[1] * c[1]); o[6] = Iw[8] + m * (cc - c[2] * c[2]);
    o[7] = Iw[1] - m * c[0] * c[1]; o[8] = Iw[2] - m * c[0] * c[2]; o[9] = Iw[5] - m * c[1] * c[2];
  }
  // spatial velocities / bias accelerations along the chain (lane 0), world origin coordinates
  T* Vb = s.H;            // [9][6]
  T* Ab = s.H + 54;       // [9][6]
  T* fb = s.H + 108;      // [9][6]
  if (g.lane == 0) {
    for (int k = 0; k < NROB; k++) {
      int p = ib_parent(k);
      T Vp[6], Ap[6];
      if (p < 0) { for (int a = 0; a < 6; a++) { Vp[a] = 0; Ap[a] = 0; } Ap[5] = -md.gravity_z; }
      else for (int a = 0; a < 6; a++) { Vp[a] = Vb[6 * p + a]; Ap[a] = Ab[6 * p + a]; }
      const T* S = s.S[k];
      T qd = s.qvel[k];
      // A_k = A_p + (V_p x S_k) qd ; V_k = V_p + S_k qd
      T c1[3], c2[3], c3[3];
      cross3(c1, Vp, S);          // w x Sw
      cross3(c2, Vp, S + 3);      // w x Sv
      cross3(c3, Vp + 3, S);      // v x Sw
      for (int a = 0; a < 3; a++) {
        Ab[6 * k + a] = Ap[a] + c1[a] * qd;
        Ab[6 * k + 3 + a] = Ap[3 + a] + (c2[a] + c3[a]) * qd;
      }
      for (int a = 0; a < 6; a++) Vb[6 * k + a] = Vp[a] + S[a] * qd;
    }
  }
  g.sync();
  // composite inertias (component-parallel suffix sums)
  for (int c = g.lane; c < 10; c += G) {
    inert[1][8][c] = inert[0][8][c];
    inert[1][7][c] = inert[0][7][c];
    T acc = inert[0][6][c] + inert[0][7][c] + inert[0][8][c];
    inert[1][6][c] = acc;
    for (int k = 5; k >= 0; k--) { acc += inert[0][k][c]; inert[1][k][c] = acc; }
  }
  // body forces f_k = I_k A_k + V_k x* (I_k V_k)
  for (int k = g.lane; k < NROB; k += G) {
    T IA[6], IV[6];
    inertia_apply(inert[0][k], Ab + 6 * k, IA);
    inertia_apply(inert[0][k], Vb + 6 * k, IV);
    const T* w = Vb + 6 * k; const T* v = w + 3;
    T c1[3], c2[3], c3[3];
    cross3(c1, w, IV);       // w x n
    cross3(c2, v, IV + 3);   // v x f
    cross3(c3, w, IV + 3);   // w x f
    for (int a = 0; a < 3; a++) { fb[6 * k + a] = IA[a] + c1[a] + c2[a]; fb[6 * k + 3 + a] = IA[3 + a] + c3[a]; }
  }
  // actuators (A5): position servos on the arm, tendon servo on the gripper
  for (int a = g.lane; a < NU; a += G) {
    T c = tclamp(s.ctrl[a], md.ctrl_lo[a], md.ctrl_hi[a]);
    T len = a < NARM ? s.qpos[a] : (T)0.5 * (s.qpos[7] + s.qpos[8]);
    T vel = a < NARM ? s.qvel[a] : (T)0.5 * (s.qvel[7] + s.qvel[8]);
    T f = md.act_gain[a] * c + md.act_b1[a] * len + md.act_b2[a] * vel;
    int sat = 0;
    if (f <= md.frc_lo[a]) { f = md.frc_lo[a]; sat = 1; }
    else if (f >= md.frc_hi[a]) { f = md.frc_hi[a]; sat = 1; }
    s.actf[a] = f;
    s.actsat[a] = sat;
  }
  g.sync();
  // F_j = Ic_j S_j
  for (int j = g.lane; j < NROB; j += G) inertia_apply(inert[1][j], s.S[j], s.tmp6[j]);
  // subtree force sums (component-parallel)
  for (int c = g.lane; c < 6; c += G) {
    T acc = fb[6 * 6 + c] + fb[6 * 7 + c] + fb[6 * 8 + c];
    fb[6 * 6 + c] = acc;
    for (int k = 5; k >= 0; k--) { acc += fb[6 * k + c]; fb[6 * k + c] = acc; }
  }
  g.sync();
  for (int e = g.lane; e < NROB * NROB; e += G) {
    int i = e / NROB, j = e % NROB;
    int lo = i < j ? i : j, hi = i < j ? j : i;
    T v = 0;
    if (hi <= 6 || lo <= 6 || lo == hi) v = dot6(s.S[lo], s.tmp6[hi]);
    if (lo >= 7 && hi >= 7 && lo != hi) v = 0;
    if (i == j) v += md.armature[i];
    s.Mr[e] = v;
  }
  for (int i = g.lane; i < NV; i += G) {
    if (i < NROB) {
      T bias = dot6(s.S[i], fb + 6 * i);
      T act = i < NARM ? s.actf[i] : (T)0.5 * s.actf[7];
      s.fs[i] = -md.damping[i] * s.qvel[i] - bias + act;
    } else {
      int k = (i - 9) % 6;
      s.fs[i] = k == 2 ? md.cube_mass * md.gravity_z : (T)0;
    }
  }
  g.sync();
  // qacc_smooth: robot block via Cholesky (factor kept in H[200..280]), cubes are diagonal
  T* L = s.H + 200;
  for (int e = g.lane; e < NROB * NROB; e += G) L[e] = s.Mr[e];
  for (int i = g.lane; i < NV; i += G) {
    if (i < NROB) s.as[i] = s.fs[i];
    else s.as[i] = s.fs[i] / (((i - 9) % 6) < 3 ? md.cube_mass : md.cube_inertia);
  }
  g.sync();
  chol_factor<T, G>(g, L, NROB);
  chol_solve<T, G>(g, L, NROB, s.as);
}

// y = M x  (robot block + diagonal cubes)
template <class T, int G>
MM_HDN void mulM(const Grp<G>& g, const Scratch<T>& s, const ModelDev<T>& md, const T* x, T* y) {
  for (int i = g.lane; i < NV; i += G) {
    if (i < NROB) {
      T a = 0;
      for (int k = 0; k < NROB; k++) a += s.Mr[i * NROB + k] * x[k];
      y[i] = a;
    } else y[i] = x[i] * (((i - 9) % 6) < 3 ? md.cube_mass : md.cube_inertia);
  }
}

// ------------------------------------------------------------------------------------------------
// collision: all 47 geoms / 780 candidate pairs - plane-box, box-box, plane-hull here, every pair
// with a mesh hull or a cylinder through GJK + EPA (mm_ccd.h)
// ------------------------------------------------------------------------------------------------
template <class T>
struct BoxRef { const T* c; const T* R; const T* s; };

// world pose of geom `gi`: position into pos[3]; returns the orientation (body frame or identity)
template <class T>
MM_HD const T* geom_pose(const Scratch<T>& s, const GeomDev<T>& gm, int gi, const T* ident, T* pos) {
  int body = gm.body[gi];
  if (body < 0) { pos[0] = gm.pos[gi][0]; pos[1] = gm.pos[gi][1]; pos[2] = gm.pos[gi][2]; return ident; }
  const T* R = s.bR[body];
  T v[3];
  rot(v, R, gm.pos[gi]);
  for (int k = 0; k < 3; k++) pos[k] = s.bpos[body][k] + v[k];
  return R;
}
// centre of the bounding sphere of geom `gi`
template <class T>
MM_HD void geom_bcenter(const Scratch<T>& s, const GeomDev<T>& gm, int gi, T* c) {
  int body = gm.body[gi];
  if (body < 0) { c[0] = gm.bc[gi][0]; c[1] = gm.bc[gi][1]; c[2] = gm.bc[gi][2]; return; }
  T v[3];
  rot(v, s.bR[body], gm.bc[gi]);
  for (int k = 0; k < 3; k++) c[k] = s.bpos[body][k] + v[k];
}

// Per-lane scratch of the box narrow phase (shared memory, NARROW_SCR reals per lane, NARROW_LANES lanes
// of a group work at a time): [0,36) axes A, B and their products C, |C|; [36,60) clip polygon;
// [60,84) clip output, finally the contact points; [27,35) finally the contact distances.
constexpr int NARROW_SCR = 84, NARROW_LANES = 8, SCR_PTS = 60, SCR_DIST = 27;

// returns number of contact points; normal nrm (A -> B); points / distances (negative) in the scratch
template <class T>
MM_HDL int box_box(const BoxRef<T>& A_, const BoxRef<T>& B_, T* nrm, T* scr) {
  const T *pa = A_.c, *Ra = A_.R, *sa = A_.s, *pb = B_.c, *Rb = B_.R, *sb = B_.s;
  T (*A)[3] = reinterpret_cast<T (*)[3]>(scr);
  T (*B)[3] = reinterpret_cast<T (*)[3]>(scr + 9);
  T (*C)[3] = reinterpret_cast<T (*)[3]>(scr + 18);
  T (*Q)[3] = reinterpret_cast<T (*)[3]>(scr + 27);
  T (*poly)[3] = reinterpret_cast<T (*)[3]>(scr + 36);
  T (*outp)[3] = reinterpret_cast<T (*)[3]>(scr + 60);
  T (*pts)[3] = outp;
  T* dist = scr + SCR_DIST;
  for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) { A[i][k] = Ra[3 * k + i]; B[i][k] = Rb[3 * k + i]; }
  T dp[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { C[i][j] = dot3(A[i], B[j]); Q[i][j] = tabs(C[i][j]); }
  T best = (T)1e30, bn[3] = {0, 0, 0};
  int code = -1;
#pragma unroll 1
  for (int i = 0; i < 3; i++) {
    T t = dot3(dp, A[i]);
    T pen = sa[i] + sb[0] * Q[i][0] + sb[1] * Q[i][1] + sb[2] * Q[i][2] - tabs(t);
    if (pen < 0) return 0;
    if (pen < best) { best = pen; code = i; T sg = t < 0 ? (T)-1 : (T)1; for (int k = 0; k < 3; k++) bn[k] = sg * A[i][k]; }
  }
#pragma unroll 1
  for (int j = 0; j < 3; j++) {
    T t = dot3(dp, B[j]);
    T pen = sb[j] + sa[0] * Q[0][j] + sa[1] * Q[1][j] + sa[2] * Q[2][j] - tabs(t);
    if (pen < 0) return 0;
    // a face of B must be clearly better than A's (parallel faces tie to rounding error)
    if (pen < best * (T)(1 - 1e-6) - (T)1e-12) { best = pen; code = 3 + j; T sg = t < 0 ? (T)-1 : (T)1; for (int k = 0; k < 3; k++) bn[k] = sg * B[j][k]; }
  }
  T ebest = (T)1e30, en[3] = {0, 0, 0};
  int ecode = -1;
#pragma unroll 1
  for (int i = 0; i < 3; i++)
#pragma unroll 1
    for (int j = 0; j < 3; j++) {
      T L[3];
      cross3(L, A[i], B[j]);
      T ln = tsqrt(dot3(L, L));
      if (ln < (T)1e-6) continue;
      T inv = (T)1 / ln;
      for (int k = 0; k < 3; k++) L[k] *= inv;
      T t = dot3(dp, L), ra = 0, rb = 0;
      for (int k = 0; k < 3; k++) { ra += sa[k] * tabs(dot3(A[k], L)); rb += sb[k] * tabs(dot3(B[k], L)); }
      T pen = ra + rb - tabs(t);
      if (pen < 0) return 0;
      if (pen < ebest) { ebest = pen; ecode = 3 * i + j; T sg = t < 0 ? (T)-1 : (T)1; for (int k = 0; k < 3; k++) en[k] = sg * L[k]; }
    }
  if (ecode >= 0 && ebest * (T)1.05 < best) {
    int i = ecode / 3, j = ecode % 3;
    T ea[3] = {pa[0], pa[1], pa[2]}, eb[3] = {pb[0], pb[1], pb[2]};
#pragma unroll 1
    for (int a = 0; a < 3; a++) {
      if (a != i) { T sg = dot3(en, A[a]) > 0 ? (T)1 : (T)-1; for (int k = 0; k < 3; k++) ea[k] += sg * sa[a] * A[a][k]; }
      if (a != j) { T sg = dot3(en, B[a]) > 0 ? (T)-1 : (T)1; for (int k = 0; k < 3; k++) eb[k] += sg * sb[a] * B[a][k]; }
    }
    T w[3] = {ea[0] - eb[0], ea[1] - eb[1], ea[2] - eb[2]};
    T uv = C[i][j], uw = dot3(A[i], w), vw = dot3(B[j], w);
    T den = (T)1 - uv * uv;
    T sp = (uv * vw - uw) / den, tp = (vw - uv * uw) / den;
    sp = tclamp(sp, -sa[i], sa[i]);
    tp = tclamp(tp, -sb[j], sb[j]);
    for (int k = 0; k < 3; k++) { pts[0][k] = (T)0.5 * ((ea[k] + sp * A[i][k]) + (eb[k] + tp * B[j][k])); nrm[k] = en[k]; }
    dist[0] = -ebest;
    return 1;
  }
  bool refA = code < 3;
  int ax = refA ? code : code - 3;
  const T* pr = refA ? pa : pb; const T* sr = refA ? sa : sb;
  const T* pi = refA ? pb : pa; const T* si = refA ? sb : sa;
  T (*Rr)[3] = refA ? A : B;
  T (*Ri)[3] = refA ? B : A;
  T nref[3];
  for (int k = 0; k < 3; k++) { nref[k] = refA ? bn[k] : -bn[k]; nrm[k] = bn[k]; }
  int iax = 0;
  T mx = -1;
  for (int a = 0; a < 3; a++) { T v = tabs(dot3(nref, Ri[a])); if (v > mx) { mx = v; iax = a; } }
  T isg = dot3(nref, Ri[iax]) > 0 ? (T)-1 : (T)1;
  int u = (iax + 1) % 3, v = (iax + 2) % 3;
  int np = 4;
#pragma unroll 1
  for (int q = 0; q < 4; q++) {
    T su = (q == 0 || q == 3) ? (T)1 : (T)-1, sv = q < 2 ? (T)1 : (T)-1;
    for (int k = 0; k < 3; k++)
      poly[q][k] = pi[k] + isg * si[iax] * Ri[iax][k] + su * si[u] * Ri[u][k] + sv * si[v] * Ri[v][k];
  }
  int t1 = (ax + 1) % 3, t2 = (ax + 2) % 3;
#pragma unroll 1
  for (int side = 0; side < 4 && np > 0; side++) {
    int ta = side < 2 ? t1 : t2;
    T sg = (side & 1) ? (T)-1 : (T)1, lim = sr[ta];
    int no = 0;
#pragma unroll 1
    for (int q = 0; q < np; q++) {
      const T* P = poly[q];
      const T* Qn = poly[(q + 1) % np];
      T rp[3] = {P[0] - pr[0], P[1] - pr[1], P[2] - pr[2]}, rq[3] = {Qn[0] - pr[0], Qn[1] - pr[1], Qn[2] - pr[2]};
      T dP = sg * dot3(rp, Rr[ta]) - lim, dQ = sg * dot3(rq, Rr[ta]) - lim;
      // 1 nm band: vertices on a side plane (exactly aligned pads) are inside, no sliver crossings
      const T ce = (T)1e-9;
      if (dP <= ce && no < 8) { for (int k = 0; k < 3; k++) outp[no][k] = P[k]; no++; }
      if (((dP < -ce && dQ > ce) || (dP > ce && dQ < -ce)) && no < 8) {
        T tt = dP / (dP - dQ);
        for (int k = 0; k < 3; k++) outp[no][k] = P[k] + tt * (Qn[k] - P[k]);
        no++;
      }
    }
    np = no;
    for (int q = 0; q < np; q++) for (int k = 0; k < 3; k++) poly[q][k] = outp[q][k];
  }
  T sgn = dot3(nref, Rr[ax]) > 0 ? (T)1 : (T)-1;
  int cnt = 0;
#pragma unroll 1
  for (int q = 0; q < np && cnt < 8; q++) {
    T r[3] = {poly[q][0] - pr[0], poly[q][1] - pr[1], poly[q][2] - pr[2]};
    T depth = sr[ax] - sgn * dot3(r, Rr[ax]);
    if (depth <= 0) continue;
    for (int k = 0; k < 3; k++) pts[cnt][k] = poly[q][k] + nref[k] * depth * (T)0.5;
    dist[cnt] = -depth;
    cnt++;
  }
  return cnt;
}

// floor plane z = 0 (normal +z) vs box: penetrating corners, at most 4
template <class T>
MM_HDL int plane_box(const BoxRef<T>& B_, T* nrm, T* scr) {
  nrm[0] = 0; nrm[1] = 0; nrm[2] = 1;
  int cnt = 0;
#pragma unroll 1
  for (int i = 0; i < 8 && cnt < 4; i++) {
    T loc[3] = {(i & 1) ? B_.s[0] : -B_.s[0], (i & 2) ? B_.s[1] : -B_.s[1], (i & 4) ? B_.s[2] : -B_.s[2]}, c[3];
    rot(c, B_.R, loc);
    for (int k = 0; k < 3; k++) c[k] += B_.c[k];
    T d = c[2];
    if (d < 0) {
      scr[SCR_PTS + 3 * cnt] = c[0]; scr[SCR_PTS + 3 * cnt + 1] = c[1]; scr[SCR_PTS + 3 * cnt + 2] = c[2] - d * (T)0.5;
      scr[SCR_DIST + cnt] = d;
      cnt++;
    }
  }
  return cnt;
}

template <class T>
MM_HD void make_tangent(const T* n, T* t1) {  // mju_makeFrame rule (A3)
  T t[3] = {0, 0, 0};
  if (n[1] < (T)0.5 && n[1] > (T)-0.5) t[1] = 1; else t[2] = 1;
  T pr = dot3(n, t);
  for (int k = 0; k < 3; k++) t[k] -= pr * n[k];
  T inv = (T)1 / tsqrt(dot3(t, t));
  for (int k = 0; k < 3; k++) t1[k] = t[k] * inv;
}

// Separating-axis test of the oriented bounding boxes (local AABB of the hull / box / cylinder, carried
// by the body frame).  Returns false only when the boxes, inflated by 1e-6, are disjoint.
template <class T>
MM_HDL bool obb_overlap(const Scratch<T>& s, const GeomDev<T>& gm, int a, int b, const T* ident) {
  T ca[3], cb[3];
  geom_bcenter(s, gm, a, ca);
  geom_bcenter(s, gm, b, cb);
  const T* Ra = gm.body[a] < 0 ? ident : s.bR[gm.body[a]];
  const T* Rb = gm.body[b] < 0 ? ident : s.bR[gm.body[b]];
  T ha[3], hb[3];
  for (int k = 0; k < 3; k++) { ha[k] = gm.size[a][k] + (T)1e-6; hb[k] = gm.size[b][k] + (T)1e-6; }
  if (gm.type[a] == GT_CYL) { ha[2] = ha[1]; ha[1] = ha[0]; }
  if (gm.type[b] == GT_CYL) { hb[2] = hb[1]; hb[1] = hb[0]; }
  // C = Ra^T Rb, t = Ra^T (cb - ca)
  T C[3][3], Q[3][3], d[3] = {cb[0] - ca[0], cb[1] - ca[1], cb[2] - ca[2]}, t[3];
  for (int i = 0; i < 3; i++) {
    t[i] = Ra[i] * d[0] + Ra[3 + i] * d[1] + Ra[6 + i] * d[2];
    for (int j = 0; j < 3; j++) {
      C[i][j] = Ra[i] * Rb[j] + Ra[3 + i] * Rb[3 + j] + Ra[6 + i] * Rb[6 + j];
      Q[i][j] = tabs(C[i][j]) + (T)1e-6;
    }
  }
  for (int i = 0; i < 3; i++)
    if (tabs(t[i]) > ha[i] + hb[0] * Q[i][0] + hb[1] * Q[i][1] + hb[2] * Q[i][2]) return false;
  for (int j = 0; j < 3; j++)
    if (tabs(t[0] * C[0][j] + t[1] * C[1][j] + t[2] * C[2][j]) > hb[j] + ha[0] * Q[0][j] + ha[1] * Q[1][j] + ha[2] * Q[2][j])
      return false;
  for (int i = 0; i < 3; i++) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    for (int j = 0; j < 3; j++) {
      int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      T ra = ha[i1] * Q[i2][j] + ha[i2] * Q[i1][j];
      T rb = hb[j1] * Q[i][j2] + hb[j2] * Q[i][j1];
      if (tabs(t[i2] * C[i1][j] - t[i1] * C[i2][j]) > ra + rb) return false;
    }
  }
  return true;
}

// plane z = 0 vs convex hull: deepest vertex, one contact
template <class T>
MM_HDL int plane_hull(const T* gpos, const T* R, const T* V, int nvert, T* nrm, T* scr) {
  nrm[0] = 0; nrm[1] = 0; nrm[2] = 1;
  T nl[3] = {R[6], R[7], R[8]};  // R^T n
  int best = 0;
  T bv = (T)1e30;
#pragma unroll 2
  for (int i = 0; i < nvert; i++) {
    T v = V[3 * i] * nl[0] + V[3 * i + 1] * nl[1] + V[3 * i + 2] * nl[2];
    if (v < bv) { bv = v; best = i; }
  }
  T c[3];
  rot(c, R, V + 3 * best);
  for (int k = 0; k < 3; k++) c[k] += gpos[k];
  T d = c[2];
  if (d < 0) {
    scr[SCR_PTS] = c[0]; scr[SCR_PTS + 1] = c[1]; scr[SCR_PTS + 2] = c[2] - d * (T)0.5;
    scr[SCR_DIST] = d;
    return 1;
  }
  return 0;
}

// shape of geom `gi` for the general convex test; body poses come as plain arrays (shared scratch or a global context)
template <class T, int G>
MM_HD void fill_shape(const Grp<G>& g, const T (*bpos)[3], const T (*bR)[9], const GeomDev<T>& gm, int gi, const T* ident, Shape<T>& sh) {
  sh.type = gm.type[gi];
  int body = gm.body[gi];
  if (body < 0) { sh.pos[0] = gm.pos[gi][0]; sh.pos[1] = gm.pos[gi][1]; sh.pos[2] = gm.pos[gi][2]; sh.R = ident; }
  else {
    const T* R = bR[body];
    T v[3];
    rot(v, R, gm.pos[gi]);
    for (int k = 0; k < 3; k++) sh.pos[k] = bpos[body][k] + v[k];
    sh.R = R;
  }
  sh.size[0] = gm.size[gi][0]; sh.size[1] = gm.size[gi][1]; sh.size[2] = gm.size[gi][2];
  sh.verts = &gm.hull[gm.vadr[gi]][0];
  sh.nvert = gm.vnum[gi];
  if (G == 32 && MM_HULL_REGS && sh.type == GT_HULL) {
#pragma unroll
    for (int k = 0; k < SHAPE_LV; k++) {
      int i = g.lane + 32 * k;
      if (i >= sh.nvert) i = 0;
      sh.lv[k][0] = sh.verts[3 * i]; sh.lv[k][1] = sh.verts[3 * i + 1]; sh.lv[k][2] = sh.verts[3 * i + 2];
    }
  }
}

// impedance / regulariser of a contact (A4): default solref (0.02, 1), solimp (0.9, 0.95, 0.001, 0.5, 2)
template <class T>
MM_HD void store_contact(Work<T>& w, int c, const T* pos, const T* nrm, const T* t1, T dist, T tran, T mu, int meta) {
  for (int d = 0; d < 3; d++) { w.cpos[d * MAXCON + c] = pos[d]; w.cn[d * MAXCON + c] = nrm[d]; w.ct1[d * MAXCON + c] = t1[d]; }
  w.cdist[c] = dist;
  T x = tabs(dist) / (T)0.001, imp;
  if (x >= 1) imp = (T)0.95;
  else { T y = x <= (T)0.5 ? 2 * x * x : 1 - 2 * (1 - x) * (1 - x); imp = (T)0.9 + y * (T)0.05; }
  T R0 = (1 - imp) / imp * tran * (1 + mu * mu);
  if (R0 < (T)MINVAL_D) R0 = (T)MINVAL_D;
  w.cD[c] = (T)1 / (2 * mu * mu * R0);
  w.cmeta[c] = meta;
}

// meta word of a contact between geoms a and b
template <class T>
MM_HD int contact_meta(const GeomDev<T>& gm, int a, int b) {
  int ca = gm.cls[a], cb = gm.cls[b];
  int cube = gm.cube[a] || gm.cube[b];
  // robot geom against an obstacle geom (table / bins; the floor does not count, gym_env.py:137-152,341-350)
  int robobs = (ca >= 1 && ca <= 9 && gm.obst[b]) || (cb >= 1 && cb <= 9 && gm.obst[a]);
  return (ca << META_KEY_SHIFT) | (cb << (META_KEY_SHIFT + 4)) | (cube << META_DIM4_BIT) | (robobs << META_ROBOBS_BIT);
}

// Broad phase: two levels over the 780 candidates -> w.surv[0, s.nsurv) in candidate order
template <class T, int G>
MM_HDX void broad_phase(const Grp<G>& g, Scratch<T>& s, const ModelDev<T>& md, Work<T>& w) {
  MM_IN_SHARED(&s);
  MM_IN_GLOBAL(&md);
  MM_IN_GLOBAL(w.surv);
  const T ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  const GeomDev<T>& gm = *md.geom;
  long long tb0 = MM_T0(s);
  // bounding-sphere centres of all geoms, once per forward pass (the H region is free during collision; the
  // narrow-phase scratch takes it over after the broad phase)
  T* bcen = s.H;
  static_assert(3 * NGEOM <= NV * NV, "bounding-sphere centres do not fit the H region");
  for (int gi = g.lane; gi < NGEOM; gi += G) geom_bcenter(s, gm, gi, bcen + 3 * gi);
  g.sync();
  // first level: ordered compaction of the surviving candidates
  int nsurv = 0;
  for (int base = 0; base < NPAIRC; base += G) {
    int ci = base + g.lane;
    int keep = 0;
    if (ci < NPAIRC) {
      // (kind, radius and the geom pair are all indexed by the candidate: three independent, coalesced loads)
      const int kind = gm.pairkind[ci];
      const T rs = gm.pairrs[ci];
      const int a = gm.pair[ci][0], b = gm.pair[ci][1];
      const T* cb = bcen + 3 * b;
      if (kind == 0) keep = !(cb[2] > rs);
      else if (kind == 1) {  // axis-aligned static box vs bounding sphere
        T d2 = 0;
        for (int k = 0; k < 3; k++) {
          T d = tabs(cb[k] - gm.pos[a][k]) - gm.size[a][k];
          if (d > 0) d2 += d * d;
        }
        keep = !(d2 > rs * rs);
      } else if (kind == 2) {
        const T* ca = bcen + 3 * a;
        T r[3] = {cb[0] - ca[0], cb[1] - ca[1], cb[2] - ca[2]};
        keep = !(dot3(r, r) > rs * rs);
      }
    }
    int tot;
    int off = g.scan_flag(keep, &tot);
    if (keep && nsurv + off < MAXSURV) w.surv[nsurv + off] = ci;
    nsurv += tot;
  }
  if (nsurv > MAXSURV) { nsurv = MAXSURV; if (g.lane == 0) s.overflow |= 1; }
  g.sync();
  // second level (conservative) on the compacted list, so that the lanes stay busy: the oriented bounding boxes
  // of the two geoms must overlap.  In-place ordered compaction (writes never pass the reads of later chunks).
  {
    int kept = 0;
    for (int base = 0; base < nsurv; base += G) {
      int si = base + g.lane;
      int keep = 0, ci = 0;
      if (si < nsurv) {
        ci = w.surv[si];
        int a = gm.pair[ci][0], b = gm.pair[ci][1];
        keep = gm.type[a] == GT_PLANE ? 1 : (int)obb_overlap(s, gm, a, b, ident);
      }
      int tot;
      int off = g.scan_flag(keep, &tot);
      g.sync();
      if (keep) w.surv[kept + off] = ci;
      kept += tot;
      g.sync();
    }
    nsurv = kept;
  }
  if (g.lane == 0) s.nsurv = nsurv;
  g.sync();
  MM_TICK(s, g, 1, tb0);
}

// Narrow phase 1: box / plane pairs, one pair per lane, NARROW_LANES lanes at a time (their clip polygons live in
// the shared H region, which is free during collision); ordered compaction of the contacts -> s.ncon (raw count), s.nbox
template <class T, int G>
MM_HDX void narrow_box(const Grp<G>& g, Scratch<T>& s, const ModelDev<T>& md, Work<T>& w) {
  MM_IN_SHARED(&s);
  MM_IN_GLOBAL(&md);
  MM_IN_GLOBAL(w.cpos); MM_IN_GLOBAL(w.cn); MM_IN_GLOBAL(w.ct1); MM_IN_GLOBAL(w.cdist); MM_IN_GLOBAL(w.cD);
  MM_IN_GLOBAL(w.cmeta); MM_IN_GLOBAL(w.surv); MM_IN_GLOBAL(w.Jv);
  const T ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  const GeomDev<T>& gm = *md.geom;
  long long tc0 = MM_T0(s);
  const int nsurv = s.nsurv;
  constexpr int KB = G < NARROW_LANES ? G : NARROW_LANES;
  // the survivors this phase handles, compacted in candidate order (a closed gripper interleaves 25 pad-pad box pairs
  // with hull pairs: without this every pass of KB survivors pays a whole box-box test for a few busy lanes); the list
  // lives in the row arrays, which are idle during collision
  int* blist = reinterpret_cast<int*>(w.Jv);
  int nb = 0;
  for (int base = 0; base < nsurv; base += G) {
    int si = base + g.lane;
    int mine = 0, ci = 0;
    if (si < nsurv) {
      ci = w.surv[si];
      mine = gm.pairflags[ci] & 1;
    }
    int tot;
    int off = g.scan_flag(mine, &tot);
    if (mine) blist[nb + off] = ci;
    nb += tot;
  }
  g.sync();
  int ncon = 0;
  for (int base = 0; base < nb; base += KB) {
    int si = base + g.lane;
    int cnt = 0, a = 0, b = 0;
    T nrm[3] = {0, 0, 1};
    T* scr = s.H + (g.lane < KB ? g.lane : 0) * NARROW_SCR;
    if (g.lane < KB && si < nb) {
      int ci = blist[si];
      a = gm.pair[ci][0]; b = gm.pair[ci][1];
      int ta = gm.type[a], tb = gm.type[b];
      if (tb == GT_BOX && (ta == GT_PLANE || ta == GT_BOX)) {
        T pb[3], pa[3];
        BoxRef<T> Bb;
        Bb.R = geom_pose(s, gm, b, ident, pb); Bb.c = pb; Bb.s = gm.size[b];
        if (ta == GT_PLANE) cnt = plane_box(Bb, nrm, scr);
        else {
          BoxRef<T> Ba;
          Ba.R = geom_pose(s, gm, a, ident, pa); Ba.c = pa; Ba.s = gm.size[a];
          cnt = box_box(Ba, Bb, nrm, scr);
        }
      } else if (ta == GT_PLANE && tb == GT_HULL) {
        T pb[3];
        const T* Rb = geom_pose(s, gm, b, ident, pb);
        cnt = plane_hull(pb, Rb, &gm.hull[gm.vadr[b]][0], gm.vnum[b], nrm, scr);
      }
    }
    int tot;
    int off = g.scan_excl(cnt, &tot);
    if (cnt > 0) {
      T t1[3];
      make_tangent(nrm, t1);
      int cube = gm.cube[a] || gm.cube[b];
      T tran = gm.invw[a] + gm.invw[b];
      int meta = contact_meta(gm, a, b);
      for (int k = 0; k < cnt; k++) {
        int c = ncon + off + k;
        if (c >= MAXCON) break;
        store_contact(w, c, scr + SCR_PTS + 3 * k, nrm, t1, scr[SCR_DIST + k], tran, cube ? (T)2 : (T)1, meta);
      }
    }
    ncon += tot;
    g.sync();
  }
  if (g.lane == 0) { s.ncon = ncon; s.nbox = ncon < MAXCON ? ncon : MAXCON; }
  g.sync();
  MM_TICK(s, g, 2, tc0);
}

// Convex candidates of the env (survivors with a mesh hull or a cylinder), in candidate order.  Returns their number;
// item k (candidate index) goes to emit(k, ci).
template <class T, int G, class Emit>
MM_HD int list_convex(const Grp<G>& g, const Scratch<T>& s, const GeomDev<T>& gm, const Work<T>& w, Emit emit) {
  const int nsurv = s.nsurv;
  int n = 0;
  for (int base = 0; base < nsurv; base += G) {
    int si = base + g.lane;
    int is = 0, ci = 0;
    if (si < nsurv) {
      ci = w.surv[si];
      is = (gm.pairflags[ci] >> 1) & 1;
    }
    int tot;
    int off = g.scan_flag(is, &tot);
    if (is) emit(n + off, ci);
    n += tot;
  }
  return n;
}

// Narrow phase 2 of ONE geom pair with a mesh hull or a cylinder: GJK + EPA by the whole group (support scans, face
// searches and face creation are spread over the lanes).  Body poses from bpos / bR; result (same on every lane) -> out.
template <class T, int G>
MM_HDN void convex_pair(const Grp<G>& g, const T (*bpos)[3], const T (*bR)[9], const GeomDev<T>& gm, int ci,
                        const EpaMem<T>& em, CvxRes<T>* out) {
  const T ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  int a = gm.pair[ci][0], b = gm.pair[ci][1];
  Shape<T> s1, s2;
  fill_shape<T, G>(g, bpos, bR, gm, a, ident, s1);
  fill_shape<T, G>(g, bpos, bR, gm, b, ident, s2);
  SP<T> sx[4];
  T pos[3] = {0, 0, 0}, pn[3] = {0, 0, 1}, depth = 0;
  bool hit = gjk<T, G>(g, s1, s2, sx);
  if (hit) hit = epa<T, G>(g, s1, s2, sx, em, pos, pn, &depth);
  if (g.lane == 0) {
    for (int k = 0; k < 3; k++) { out->pos[k] = pos[k]; out->nrm[k] = pn[k]; }
    out->depth = depth; out->hit = hit ? 1 : 0; out->ci = ci;
  }
  g.sync();
}

// The convex results w.cvx[0, s.ncvx) (candidate order) are appended to the box contacts; then the box contacts
// [0, nbox) and the convex contacts [nbox, ncon) - each ordered by body-pair key - are merged (stable) so that every
// body pair owns ONE contiguous run of contacts and therefore one pair slot; finally the pair tables.
template <class T, int G>
MM_HDX void assemble_contacts(const Grp<G>& g, Scratch<T>& s, const ModelDev<T>& md, Work<T>& w) {
  MM_IN_SHARED(&s);
  MM_IN_GLOBAL(&md);
  MM_IN_GLOBAL(w.cpos); MM_IN_GLOBAL(w.cn); MM_IN_GLOBAL(w.ct1); MM_IN_GLOBAL(w.cdist); MM_IN_GLOBAL(w.cD); MM_IN_GLOBAL(w.aref);
  MM_IN_GLOBAL(w.cmeta);
  const GeomDev<T>& gm = *md.geom;
  long long tx0 = MM_T0(s);
  int ncon = s.ncon;
  const int nbox = s.nbox, ncvx = s.ncvx;
  for (int base = 0; base < ncvx; base += G) {
    int k = base + g.lane;
    int hit = 0;
    if (k < ncvx) hit = w.cvx[k].hit;
    int tot;
    int off = g.scan_flag(hit, &tot);
    if (hit && ncon + off < MAXCON) {
      const CvxRes<T>& r = w.cvx[k];
      int a = gm.pair[r.ci][0], b = gm.pair[r.ci][1];
      T pos[3] = {r.pos[0], r.pos[1], r.pos[2]}, pn[3] = {r.nrm[0], r.nrm[1], r.nrm[2]}, t1[3];
      make_tangent(pn, t1);
      int cube = gm.cube[a] || gm.cube[b];
      store_contact(w, ncon + off, pos, pn, t1, -r.depth, gm.invw[a] + gm.invw[b], cube ? (T)2 : (T)1, contact_meta(gm, a, b));
    }
    ncon += tot;
  }
  if (ncon > MAXCON) { ncon = MAXCON; if (g.lane == 0) s.overflow |= 2; }
  g.sync();
  if (ncon > nbox && nbox > 0) {
    for (int base = 0; base < ncon; base += G) {
      int c = base + g.lane;
      int dst = 0, m = 0;
      T v[11];
      if (c < ncon) {
        m = w.cmeta[c];
        int key = sort_key(m);
        if (c < nbox) {  // box contact: stays before convex contacts of the same key
          int before = 0;
          for (int j = nbox; j < ncon; j++) before += sort_key(w.cmeta[j]) < key;
          dst = c + before;
        } else {
          int before = 0;
          for (int j = 0; j < nbox; j++) before += sort_key(w.cmeta[j]) <= key;
          dst = (c - nbox) + before;
        }
        for (int d = 0; d < 3; d++) { v[d] = w.cpos[d * MAXCON + c]; v[3 + d] = w.cn[d * MAXCON + c]; v[6 + d] = w.ct1[d * MAXCON + c]; }
        v[9] = w.cdist[c]; v[10] = w.cD[c];
      }
      // destinations of this chunk may hit sources of later chunks: stage through the row arrays (aref, Jaref, Jv
      // are contiguous, 18 * MAXCON reals, and not in use before make_constraints)
      if (c < ncon) {
        for (int d = 0; d < 11; d++) w.aref[d * MAXCON + dst] = v[d];
        reinterpret_cast<int*>(w.aref + 12 * MAXCON)[dst] = m;
      }
    }
    g.sync();
    for (int c = g.lane; c < ncon; c += G) {
      for (int d = 0; d < 3; d++) {
        w.cpos[d * MAXCON + c] = w.aref[d * MAXCON + c]; w.cn[d * MAXCON + c] = w.aref[(3 + d) * MAXCON + c];
        w.ct1[d * MAXCON + c] = w.aref[(6 + d) * MAXCON + c];
      }
      w.cdist[c] = w.aref[9 * MAXCON + c]; w.cD[c] = w.aref[10 * MAXCON + c];
      w.cmeta[c] = reinterpret_cast<int*>(w.aref + 12 * MAXCON)[c];
    }
    g.sync();
  }
  MM_TICK(s, g, 3, tx0);
  // pair slots: contacts are ordered by (classA, classB); a new slot starts where the key changes.  ONE pass over the
  // meta words: the key of the left neighbour comes by shuffle (by a load for the first lane of a warp), and
  // the pair tables are written to shared memory on the assumption that at most MAXPAIR_S body pairs touch; only when
  // there are more (rare) a second pass fills the tables of the global spill instead (same layout).
  if (g.lane == 0) {
    s.pairK = s.pairK_s; s.pairW = s.pairW_s; s.pairF = s.pairF_s;
    s.pairkey = s.pairkey_s; s.pairmd = s.pairmd_s; s.pairmb = s.pairmb_s;
  }
  int npair = 0;
  {
    for (int base = 0; base < ncon; base += G) {
      int c = base + g.lane;
      int key = -1, m = 0, left = -1;
      if (c < ncon) {
        m = w.cmeta[c];
        key = meta_key(m);
        // (first lane of a warp: the neighbour's word comes from memory - its key bits are never rewritten)
        if (g.wlane() == 0 && c > 0) left = meta_key(w.cmeta[c - 1]);
      }
      int prev = g.wshfl_up(key, 1);
      if (g.wlane() == 0) prev = left;
      int head = c < ncon && prev != key;
      int tot;
      int off = g.scan_flag(head, &tot);
      if (c < ncon) {
        int slot = npair + off + head - 1;
        if (slot >= MAXPAIR) slot = MAXPAIR - 1;  // unreachable: MAXPAIR covers every (class, class) key of the model
        w.cmeta[c] = (m & ~127) | slot;
        if (head && npair + off < MAXPAIR_S) {
          int mA = dofmask(key & 15), mB = dofmask((key >> 4) & 15);
          s.pairkey_s[npair + off] = key; s.pairmd_s[npair + off] = mA ^ mB; s.pairmb_s[npair + off] = mB;
        }
      }
      npair += tot;
    }
  }
  if (npair > MM_PAIR_SPILL_AT) {  // (MAXPAIR_S; a test build lowers it to send every env through the spill)
    if (g.lane == 0) {
      s.pairK = reinterpret_cast<T (*)[21]>(w.pairbig);
      s.pairW = reinterpret_cast<T (*)[6]>(w.pairbig + 21 * MAXPAIR);
      s.pairF = reinterpret_cast<T (*)[6]>(w.pairbig + 27 * MAXPAIR);
      s.pairkey = w.pairbig_i; s.pairmd = w.pairbig_i + MAXPAIR; s.pairmb = w.pairbig_i + 2 * MAXPAIR;
    }
    g.sync();
    for (int base = 0; base < ncon; base += G) {
      int c = base + g.lane;
      int slot = -1, key = 0, left = -1;
      if (c < ncon) {
        int m = w.cmeta[c];
        slot = meta_slot(m); key = meta_key(m);
        if (g.wlane() == 0 && c > 0) left = meta_slot(w.cmeta[c - 1]);
      }
      int prev = g.wshfl_up(slot, 1);
      if (g.wlane() == 0) prev = left;
      if (c < ncon && prev != slot && slot < MAXPAIR) {
        int mA = dofmask(key & 15), mB = dofmask((key >> 4) & 15);
        s.pairkey[slot] = key; s.pairmd[slot] = mA ^ mB; s.pairmb[slot] = mB;
      }
    }
  }
  if (npair > MAXPAIR) { npair = MAXPAIR; if (g.lane == 0) s.overflow |= 4; }
  if (g.lane == 0) { s.ncon = ncon; s.npair = npair; }
  g.sync();
}

// Fused collision of one env by its own group (reset / engine-level ops / host emulation): the convex pairs are
// tested one after the other by the whole group; results are identical to the staged path (mm_stage.h), where a
// batch-wide kernel tests every queued pair with its own warp.
template <class T, int G>
MM_HDX void collide(const Grp<G>& g, Scratch<T>& s, const ModelDev<T>& md, Work<T>& w) {
  const GeomDev<T>& gm = *md.geom;
  broad_phase<T, G>(g, s, md, w);
  narrow_box<T, G>(g, s, md, w);
  // candidate indices first (the EPA polytope takes over the shared H region below)
  CvxRes<T>* res = w.cvx;
  int ncvx = list_convex<T, G>(g, s, gm, w, [&](int k, int ci) { res[k].ci = ci; });
  if (g.lane == 0) s.ncvx = ncvx;
  g.sync();
  // EPA polytope: faces, their index words, horizon edges and canonical vertex ids live in SHARED memory (the H /
  // tmp6 / pair-block region, free during collision); vertices stay in the global workspace.
  static_assert(sizeof(s.H) + sizeof(s.tmp6) + sizeof(s.pairK_s) + sizeof(s.pairW_s) + sizeof(s.pairF_s) + (ROWS_S ? sizeof(s.rows_s) : 0) + 5 * sizeof(s.qacc) >=
                    EPA_MAXF * 4 * sizeof(T) + EPA_INTS * sizeof(int), "EPA workspace does not fit the shared scratch");
  EpaMem<T> em;
  em.vert = w.epa.vert;
  em.face = s.H;
  em.fidx = reinterpret_cast<int*>(s.H + EPA_MAXF * 4);
  em.edge = em.fidx + EPA_MAXF;
  em.canon = em.edge + EPA_MAXE;
  em.ecan = em.canon + EPA_MAXV;
  long long tx0 = MM_T0(s);
  for (int k = 0; k < ncvx; k++) convex_pair<T, G>(g, s.bpos, s.bR, gm, res[k].ci, em, res + k);
  MM_TICK(s, g, 3, tx0);
  assemble_contacts<T, G>(g, s, md, w);
}

// ------------------------------------------------------------------------------------------------
// matrix-free Jacobian products through body-pair twists / wrenches
// ------------------------------------------------------------------------------------------------
template <class T, int G>
MM_HDS void pair_twists(const Grp<G>& g, Scratch<T>& s, const T* x) {
  int np = s.npair;
  for (int idx = g.lane; idx < np * 6; idx += G) {
    int p = idx / 6, c = idx % 6;
    unsigned md = (unsigned)s.pairmd[p];  // dofs shared by both bodies cancel
    int mB = s.pairmb[p];
    T acc = 0;
    while (md) {  // ascending dof order
      int i = tctz(md);
      md &= md - 1;
      T sg = ((mB >> i) & 1) ? (T)1 : (T)-1;
      acc += sg * x[i] * S_comp(s, i, c);
    }
    s.pairW[p][c] = acc;
  }
  g.sync();
}

template <class T>
struct ConGeom { T pos[3], n[3], t1[3], t2[3]; };
template <class T>
MM_HD void load_con(const Work<T>& w, int c, ConGeom<T>& q) {
  for (int d = 0; d < 3; d++) { q.pos[d] = w.cpos[d * MAXCON + c]; q.n[d] = w.cn[d * MAXCON + c]; q.t1[d] = w.ct1[d * MAXCON + c]; }
  cross3(q.t2, q.n, q.t1);
}

// rows of J * x of one contact from the relative twist W of its body pair: out[0..3] (and [4], [5] for a contact with
// torsional friction; zeros otherwise); returns whether the contact has six rows
template <class T>
MM_HD bool con_rows(const ConGeom<T>& q, const T* W, int m, T* out) {
  T u[3];
  cross3(u, W, q.pos);
  for (int d = 0; d < 3; d++) u[d] += W[3 + d];
  T un = dot3(q.n, u), u1 = dot3(q.t1, u), u2 = dot3(q.t2, u);
  const bool dim4 = meta_dim4(m);
  T mu = dim4 ? (T)2 : (T)1;
  out[0] = un + mu * u1; out[1] = un - mu * u1;
  out[2] = un + mu * u2; out[3] = un - mu * u2;
  out[4] = 0; out[5] = 0;
  if (dim4) {
    T u3 = dot3(q.n, W);  // torsional friction coefficient of cube contacts = 1.0
    out[4] = un + u3; out[5] = un - u3;
  }
  return dim4;
}

// Reference accelerations and warm-start selection in ONE pass over the contacts: aref from J qvel, then the rows
// J x - aref at x0 = qacc_smooth (kept in Jv) and at x1 = qacc_warmstart (kept in Jaref) and this lane's share of both
// constraint costs.  Same arithmetic per row and the
// same order of the cost terms as two mulJ passes each followed by a cost loop, but the contact geometry, the reference
// accelerations and the pair tables are fetched once, the two evaluations overlap instruction by instruction, and half
// of the barriers are gone.  The relative twists of x1 are parked in the pair-force table (idle until the first
// constraint update clears it).
template <class T, int G>
MM_HDS void warm_rows(const Grp<G>& g, Scratch<T>& s, const ModelDev<T>& md, const Work<T>& w, const T* x0, const T* x1, T* cost0, T* cost1) {
  int np = s.npair;
  for (int idx = g.lane; idx < np * 6; idx += G) {
    int p = idx / 6, c = idx % 6;
    unsigned dm = (unsigned)s.pairmd[p];  // dofs shared by both bodies cancel
    int mB = s.pairmb[p];
    T a0 = 0, a1 = 0, av = 0;
    while (dm) {  // ascending dof order
      int i = tctz(dm);
      dm &= dm - 1;
      T sc = (((mB >> i) & 1) ? (T)1 : (T)-1);
      T Sc = S_comp(s, i, c);
      a0 += sc * x0[i] * Sc;
      a1 += sc * x1[i] * Sc;
      av += sc * s.qvel[i] * Sc;
    }
    s.pairW[p][c] = a0;
    s.pairF[p][c] = a1;
    s.pairK[p][c] = av;  // (relative velocity of the pair: head of its block, which the first constraint update clears)
  }
  g.sync();
  // contact aref_r = -B (J qvel)_r - K imp dist (A4: default solref (0.02, 1), solimp (0.9, 0.95, 0.001, 0.5, 2)), made
  // here, in registers, from the same fetch of the contact as the two candidates' rows
  const T h = md.timestep;
  const T tc = tmax((T)0.02, 2 * h), dmax = (T)0.95;
  const T Kc = (T)1 / (dmax * dmax * tc * tc), Bc = (T)2 / (dmax * tc);
  T c0 = 0, c1 = 0;
  int ncon = s.ncon;
  for (int c = g.lane; c < ncon; c += G) {
    ConGeom<T> q;
    load_con(w, c, q);
    int m = w.cmeta[c];
    T D = w.cD[c];
    T dist = w.cdist[c];
    const bool dim4 = meta_dim4(m);
    T ar[6];
    {
      con_rows(q, s.pairK[meta_slot(m)], m, ar);
      T x = tabs(dist) / (T)0.001, imp;
      if (x >= 1) imp = (T)0.95;
      else { T y = x <= (T)0.5 ? 2 * x * x : 1 - 2 * (1 - x) * (1 - x); imp = (T)0.9 + y * (T)0.05; }
#pragma unroll
      for (int r = 0; r < 6; r++) ar[r] = (r < 4 || dim4) ? -Bc * ar[r] - Kc * imp * dist : (T)0;
#ifndef __CUDA_ARCH__
      // (host build only: the row array keeps a copy for the engine known-answer tests; the kernels never read it)
      for (int r = 0; r < (dim4 ? 6 : 4); r++) w.aref[c * 6 + r] = ar[r];
#endif
    }
#pragma unroll
    for (int which = 0; which < 2; which++) {
      const T* W = which ? s.pairF[meta_slot(m)] : s.pairW[meta_slot(m)];
      T* out = which ? w.Jaref : w.Jv;
      T row[6];
      con_rows(q, W, m, row);
      for (int r = 0; r < 6; r++) {
        if (r >= 4 && !dim4) break;
        T ja = row[r] - ar[r];
        out[c * 6 + r] = ja;
        const T Da = ja < 0 ? D : (T)0;
        if (which) c1 += (T)0.5 * Da * ja * ja; else c0 += (T)0.5 * Da * ja * ja;  // (term by term: the order of the sum is part of the result)
      }
    }
  }
  for (int k = g.lane; k < s.nspec; k += G) {
    int d = s.specdof[k];
#pragma unroll
    for (int which = 0; which < 2; which++) {
      const T* x = which ? x1 : x0;
      T ja = (d < 0 ? x[7] - x[8] : ((d & 256) ? -x[d & 255] : x[d & 255])) - s.specAref[k];
      if (which) s.specJaref[k] = ja; else s.specJv[k] = ja;
      const T Da = (d < 0 || ja < 0) ? s.specD[k] : (T)0;
      if (which) c1 += (T)0.5 * Da * ja * ja; else c0 += (T)0.5 * Da * ja * ja;
    }
  }
  g.sync();
  *cost0 = c0;
  *cost1 = c1;
}

// ------------------------------------------------------------------------------------------------
// constraint rows (A4): per-contact D is set in collide(); here the reference accelerations
// ------------------------------------------------------------------------------------------------
template <class T>
MM_HD T impedance_generic(const T* solimp, T pos) {
  T dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  T x = tabs(pos) / width;
  if (x >= 1) return dmax;
  if (x <= 0) return dmin;
  // power = 2 for every solimp of this model (panda.xml:260-262 and the MuJoCo default); fill_model() refuses any other
  // value, so the general x^p branch (a 500-instruction pow expansion on the device) does not exist here
  (void)power;
  T y = x <= mid ? x * x / mid : 1 - (1 - x) * (1 - x) / (1 - mid);
  return dmin + y * (dmax - dmin);
}

template <class T, int G>
MM_HDX void make_constraints(const Grp<G>& g, Scratch<T>& s, const ModelDev<T>& md, Work<T>& w) {
  MM_IN_SHARED(&s);
  MM_IN_GLOBAL(&md);
  MM_IN_GLOBAL(w.cpos); MM_IN_GLOBAL(w.cn); MM_IN_GLOBAL(w.ct1); MM_IN_GLOBAL(w.cdist); MM_IN_GLOBAL(w.cD); MM_AREF_GLOBAL(w.aref);
  MM_ROWS_GLOBAL(w.Jaref); MM_ROWS_GLOBAL(w.Jv); MM_IN_GLOBAL(w.cmeta); MM_IN_GLOBAL(w.surv);
  const T h = md.timestep;
  if (g.lane == 0) {
    int n = 0;
    // joint equality f1 = f2 : solref (0.005, 1) -> refsafe, solimp (0.95, 0.99, 0.001, 0.5, 2)
    {
      T pos = s.qpos[7] - s.qpos[8], vel = s.qvel[7] - s.qvel[8];
      T tc = tmax(md.eq_solref[0], 2 * h), dr = md.eq_solref[1], dmax = md.eq_solimp[1];
      T K = (T)1 / (dmax * dmax * tc * tc * dr * dr), B = (T)2 / (dmax * tc);
      T imp = impedance_generic(md.eq_solimp, pos);
      T R = tmax((T)MINVAL_D, (1 - imp) / imp * md.eq_invw);
      s.specdof[n] = -1; s.specD[n] = (T)1 / R; s.specAref[n] = -B * vel - K * imp * pos;
      n++;
    }
    const T dsolimp[5] = {(T)0.9, (T)0.95, (T)0.001, (T)0.5, (T)2};
    const T tc = tmax((T)0.02, 2 * h), dmax = (T)0.95;
    const T K = (T)1 / (dmax * dmax * tc * tc), B = (T)2 / (dmax * tc);
    for (int i = 0; i < NROB; i++) {
      T q = s.qpos[i];
      T dlo = q - md.jnt_lo[i], dhi = md.jnt_hi[i] - q;
      for (int side = 0; side < 2; side++) {
        T dist = side == 0 ? dlo : dhi;
        if (dist < 0 && n < MAXSPEC) {
          T imp = impedance_generic(dsolimp, dist);
          T R = tmax((T)MINVAL_D, (1 - imp) / imp * md.dof_invw[i]);
          T vel = side == 0 ? s.qvel[i] : -s.qvel[i];
          s.specdof[n] = i | (side ? 256 : 0); s.specD[n] = (T)1 / R; s.specAref[n] = -B * vel - K * imp * dist;
          n++;
        }
      }
    }
    s.nspec = n;
  }
  g.sync();  // (the contacts' reference accelerations are made in warm_rows, from the fetch that also serves the warm start)
}

// ------------------------------------------------------------------------------------------------
// Newton solver (A6)
// ------------------------------------------------------------------------------------------------
// Constraint update at the current Jaref: active set, cost, qfrc_constraint (s.fc); when `buildK`
// also the per-pair 6x6 blocks K_p = sum_active D y y^T.  Returns constraint cost; *changed is set
// when any active bit differs from the stored one.
template <class T, int G>
MM_HDL T update_constraint(const Grp<G>& g, Scratch<T>& s, Work<T>& w, bool buildK, int* changed) {
  int np = s.npair;
  {  // (rows are contiguous: flat index, no division)
    T* pf = &s.pairF[0][0];
    T* pk = &s.pairK[0][0];
#pragma unroll 1
    for (int idx = g.lane; idx < np * 6; idx += G) pf[idx] = 0;
    if (buildK) {
#pragma unroll 1
      for (int idx = g.lane; idx < np * 21; idx += G) pk[idx] = 0;
    }
  }
  g.sync();
  T cost = 0;
  int chg = 0;
  int ncon = s.ncon;
  for (int base = 0; base < ncon; base += G) {
    int c = base + g.lane;
    bool valid = c < ncon;
    T F[6] = {0, 0, 0, 0, 0, 0};
    T Kc[21];
    if (buildK) for (int k = 0; k < 21; k++) Kc[k] = 0;
    int slot = -1 - g.lane;  // invalid lanes never merge
    if (valid) {
      ConGeom<T> q;
      load_con(w, c, q);
      int m = w.cmeta[c];
      slot = meta_slot(m);
      T D = w.cD[c];
      int dim4 = meta_dim4(m), nr = dim4 ? 6 : 4;
      T mu = dim4 ? (T)2 : (T)1;
      int bits = 0;
      T pxn[3], px1[3], px2[3];
      cross3(pxn, q.pos, q.n);
      cross3(px1, q.pos, q.t1);
      cross3(px2, q.pos, q.t2);
#pragma unroll 1  // (rolled on purpose: six unrolled copies of the row body are 8 KB of instruction cache)
      for (int r = 0; r < nr; r++) {
        T ja = w.Jaref[c * 6 + r];
        // (no branch on the row's activity: an inactive row runs the same code with D = 0 - its contributions are
        // exact zeros - instead of costing an instruction-fetch bubble per row)
        const bool act = ja < 0;
        const T Da = act ? D : (T)0;
        bits |= (act ? 1 : 0) << r;
        T f = -Da * ja;
        cost += (T)0.5 * Da * ja * ja;
        // y_r = [pos x d + tau ; d]
        T y[6];
        T sg = (r & 1) ? (T)-1 : (T)1;
        if (r < 2) for (int d = 0; d < 3; d++) { y[d] = pxn[d] + sg * mu * px1[d]; y[3 + d] = q.n[d] + sg * mu * q.t1[d]; }
        else if (r < 4) for (int d = 0; d < 3; d++) { y[d] = pxn[d] + sg * mu * px2[d]; y[3 + d] = q.n[d] + sg * mu * q.t2[d]; }
        else for (int d = 0; d < 3; d++) { y[d] = pxn[d] + sg * q.n[d]; y[3 + d] = q.n[d]; }
        for (int d = 0; d < 6; d++) F[d] += f * y[d];
        if (buildK) {
          int k = 0;
          for (int a = 0; a < 6; a++) { T da = Da * y[a]; for (int b = 0; b <= a; b++) Kc[k++] += da * y[b]; }
        }
      }
      int old = (m >> META_ACT_SHIFT) & 63;
      if (old != bits) { chg = 1; w.cmeta[c] = (m & ~(63 << META_ACT_SHIFT)) | (bits << META_ACT_SHIFT); }
    }
    // segmented inclusive scan over the lanes of a warp (keys are non-decreasing); segment tails commit.  A group that
    // spans several warps (Grp<128>) commits warp after warp, so that a run cut by a warp boundary adds up in order.
    constexpr int WL = Grp<G>::WL;
    if (WL > 1) {
#pragma unroll 1
      for (int o = 1; o < WL; o <<= 1) {
        int ks = g.wshfl_up(slot, o);
        bool take = g.wlane() >= o && ks == slot;
        if (!g.wany(take)) break;  // slots are sorted: no run is longer than o
        for (int d = 0; d < 6; d++) { T t = g.wshfl_up(F[d], o); if (take) F[d] += t; }
        if (buildK) for (int k = 0; k < 21; k++) { T t = g.wshfl_up(Kc[k], o); if (take) Kc[k] += t; }
      }
    }
    int nxt = g.wshfl_down(slot, 1);
    bool tail = valid && (WL == 1 || g.wlane() == WL - 1 || nxt != slot);
    for (int wq = 0; wq < Grp<G>::nwarps(); wq++) {
      if (tail && g.warp() == wq) {
        for (int d = 0; d < 6; d++) s.pairF[slot][d] += F[d];
        if (buildK) for (int k = 0; k < 21; k++) s.pairK[slot][k] += Kc[k];
      }
      g.sync();
    }
  }
  // special rows
  for (int k = g.lane; k < s.nspec; k += G) {
    T ja = s.specJaref[k];
    const T Da = (s.specdof[k] < 0 || ja < 0) ? s.specD[k] : (T)0;
    cost += (T)0.5 * Da * ja * ja;
  }
  cost = g.sum(cost);
  *changed = g.any(chg);
  // qfrc_constraint
  // (branch-free on purpose: a pair / row that does not involve dof i contributes with coefficient 0 - a data-dependent
  // branch per pair cost more in instruction-fetch bubbles than the six idle multiply-adds do)
  for (int i = g.lane; i < NV; i += G) {
    T acc = 0;
    T Si[6];
    S_get(s, i, Si);
    for (int p = 0; p < np; p++) {
      const int in = (s.pairmd[p] >> i) & 1, pos = (s.pairmb[p] >> i) & 1;
      const T c = in ? (pos ? (T)1 : (T)-1) : (T)0;
      acc += c * dot6(Si, s.pairF[p]);
    }
    for (int k = 0; k < s.nspec; k++) {
      const int d = s.specdof[k];
      const T ja = s.specJaref[k];
      const T f = -s.specD[k] * ja;
      const T ceq = i == 7 ? (T)1 : (i == 8 ? (T)-1 : (T)0);
      const T clim = ((d & 255) == i && ja < 0) ? ((d & 256) ? (T)-1 : (T)1) : (T)0;
      acc += (d < 0 ? ceq : clim) * f;
    }
    s.fc[i] = acc;
  }
  g.sync();
  return cost;
}

// (row, column) of the e-th entry of a packed lower triangle, row-major, up to 27 rows
template <class T>
MM_HD void tri_rc(int e, int* r, int* c) {
  int i = (int)((sqrtf((float)(8 * e + 1)) - 1.0f) * 0.5f);
  while ((i + 1) * (i + 2) / 2 <= e) i++;
  while (i * (i + 1) / 2 > e) i--;
  *r = i; *c = e - i * (i + 1) / 2;
}

// which cube blocks are independent this forward pass (from the touching body pairs)
template <class T, int G>
MM_HDN void analyse_coupling(const Grp<G>& g, Scratch<T>& s) {
  int busy = 0;  // cubes that touch the robot or another cube
  for (int p = 0; p < s.npair; p++) {
    int key = s.pairkey[p], ca = key & 15, cb = (key >> 4) & 15;
    bool ra = ca >= 1 && ca <= 9, rb = cb >= 1 && cb <= 9;
    int qa = ca >= CLS_CUBE0 ? ca - CLS_CUBE0 : -1, qb = cb >= CLS_CUBE0 ? cb - CLS_CUBE0 : -1;
    if (ra && qb >= 0) busy |= 1 << qb;
    if (rb && qa >= 0) busy |= 1 << qa;
    if (qa >= 0 && qb >= 0) busy |= (1 << qa) | (1 << qb);
  }
  if (g.lane == 0) {
    int n = 0;
    for (int i = 0; i < NROB; i++) s.il[n++] = (signed char)i;
    for (int c = 0; c < 3; c++)
      if ((busy >> c) & 1) for (int k = 0; k < 6; k++) s.il[n++] = (signed char)(9 + 6 * c + k);
    s.n_il = n;
    s.lone = (~busy) & 7;
  }
  g.sync();
}

// H = M + J^T D J over the active rows, assembled from the per-pair blocks; then factor in place
template <class T, int G>
MM_HDS void build_factor_H(const Grp<G>& g, Scratch<T>& s, const ModelDev<T>& md) {
  // H = M on the entries that will be read: the robot block, and of every cube its own block (lone cube) or its whole
  // row (cube coupled with the robot / another cube); constant divisors, no triangular index arithmetic
  // (three short loops: the robot block; the 6x6 diagonal block of every cube; and - only while some cube is coupled -
  // the part of a coupled cube's rows left of its block.  One loop over all 81 + 3 x 6 x 27 entries took 18 turns per
  // lane, and 62 % of the envs need 7.)
  for (int e = g.lane; e < NROB * NROB; e += G) {
    int i = e / NROB, j = e - i * NROB;
    if (j <= i) s.H[i * NV + j] = s.Mr[e];
  }
  for (int e = g.lane; e < 3 * 36; e += G) {
    int c = e / 36, r = e - c * 36, k = r / 6, l = r - k * 6;
    if (l <= k) s.H[(NROB + 6 * c + k) * NV + NROB + 6 * c + l] = k == l ? (k < 3 ? md.cube_mass : md.cube_inertia) : (T)0;
  }
  if (s.lone != 7) {
    for (int e = g.lane; e < 3 * 6 * (NROB + 12); e += G) {
      int c = e / (6 * (NROB + 12)), r = e - c * 6 * (NROB + 12), k = r / (NROB + 12), j = r - k * (NROB + 12);
      if (((s.lone >> c) & 1) || j >= NROB + 6 * c) continue;
      s.H[(NROB + 6 * c + k) * NV + j] = (T)0;
    }
  }
  g.sync();
  if (g.lane == 0) {
    for (int k = 0; k < s.nspec; k++) {
      int d = s.specdof[k];
      T D = s.specD[k];
      if (d < 0) { s.H[7 * NV + 7] += D; s.H[8 * NV + 8] += D; s.H[8 * NV + 7] -= D; }
      else if (s.specJaref[k] < 0) s.H[(d & 255) * (NV + 1)] += D;
    }
  }
  for (int p = 0; p < s.npair; p++) {
    int mB = s.pairmb[p];
    int mm_ = s.pairmd[p];  // dofs shared by both bodies cancel (sigma = 0)
    const T* K = s.pairK[p];
    g.sync();
    // dofs of this pair in ascending order (rank = number of lower set bits), and u_k = sigma K S
    for (int j = g.lane; j < NV; j += G) {
      if (!((mm_ >> j) & 1)) continue;
      int k = tpopc(mm_ & ((1 << j) - 1));
      s.dl[k] = (signed char)j;
      T S[6];
      S_get(s, j, S);
      T sg = ((mB >> j) & 1) ? (T)1 : (T)-1;
      // symmetric packed K (lower, row-major): K[a][b] = K[a*(a+1)/2 + b], b <= a
      for (int a = 0; a < 6; a++) {
        T acc = 0;
        for (int b = 0; b < 6; b++) acc += (a >= b ? K[a * (a + 1) / 2 + b] : K[b * (b + 1) / 2 + a]) * S[b];
        s.tmp6[k][a] = sg * acc;
      }
    }
    g.sync();
    int nd = tpopc(mm_);
    for (int e = g.lane; e < nd * (nd + 1) / 2; e += G) {
      int a, b;
      tri_rc<T>(e, &a, &b);
      int i = s.dl[a], j = s.dl[b];
      T sg = ((mB >> i) & 1) ? (T)1 : (T)-1;
      s.H[i * NV + j] += sg * S_dot(s, i, s.tmp6[b]);
    }
  }
  g.sync();
  // factor: independent cube blocks on single lanes, the coupled part cooperatively
  for (int c = (G == 1 ? 0 : g.lane); c < 3; c += (G == 1 ? 1 : G))
    if ((s.lone >> c) & 1) chol6_local(s.H + (9 + 6 * c) * (NV + 1));
  chol_factor_list<T, G>(g, s.H, s.il, s.n_il);
}

// Rows of the Newton direction, exact line search along it and the move of the rows, in one function: J * search for
// every contact, then Newton iterations on the one-dimensional cost (bracketed; mj_solver's rule as the oracle states
// it), then Jaref += step * Jv.  Returns the step; *flat = the slope at 0 is not negative (then, and with a zero step,
// nothing moves).  Every evaluation of the line search needs the derivative and curvature of the cost at alpha, i.e. a
// pass over all rows: the rows of this lane's FIRST contact (Jaref, Jv, D - all an env with at most G contacts has per
// lane) are computed / fetched once and stay in registers from the row product to the move, so such an env reads its
// contact list once per Newton iteration here and an evaluation touches no memory beyond the special rows.  Rows past
// a contact's dimension are held as zeros and contribute exact zeros (no case split per row); further contacts of the
// lane go through the row arrays.
template <class T, int G>
MM_HDS T dir_search(const Grp<G>& g, Scratch<T>& s, const Work<T>& w, T qg1, T qg2, T gtol, bool* flat_out) {
  pair_twists<T, G>(g, s, s.search);
  const int ncon = s.ncon;
  T ja0[6], jv0[6], D0 = 0;
  bool dim40 = false;
#pragma unroll
  for (int r = 0; r < 6; r++) { ja0[r] = 0; jv0[r] = 0; }
  for (int c = g.lane; c < ncon; c += G) {
    ConGeom<T> q;
    load_con(w, c, q);
    int m = w.cmeta[c];
    T row[6];
    const bool dim4 = con_rows(q, s.pairW[meta_slot(m)], m, row);
    if (c == g.lane) {
      D0 = w.cD[c];
      dim40 = dim4;
#pragma unroll
      for (int r = 0; r < 6; r++)
        if (r < 4 || dim4) { jv0[r] = row[r]; ja0[r] = w.Jaref[c * 6 + r]; }
    } else {
      for (int r = 0; r < (dim4 ? 6 : 4); r++) w.Jv[c * 6 + r] = row[r];
    }
  }
  for (int k = g.lane; k < s.nspec; k += G) {
    int d = s.specdof[k];
    const T* x = s.search;
    s.specJv[k] = d < 0 ? x[7] - x[8] : ((d & 256) ? -x[d & 255] : x[d & 255]);
  }
  g.sync();
  T lo = 0, hi = -1, d1 = 0, d2 = 1, a = 0;
  bool flat = false;
#pragma unroll 1
  for (int it = -1; it < 50; it++) {  // it = -1: slope at alpha = 0
    if (it >= 0) {
      T an = a - d1 / d2;
      if (hi > 0 && !(an > lo && an < hi)) an = (T)0.5 * (lo + hi);
      a = an;
    }
    // derivative / curvature of the cost at step a
    T a1 = 0, a2 = 0;
#pragma unroll
    for (int r = 0; r < 6; r++) {
      T jv = jv0[r];
      T x = ja0[r] + a * jv;
      const T Da = x < 0 ? D0 : (T)0;  // (select instead of a branch per row: no instruction-fetch bubble)
      a1 += Da * jv * x; a2 += Da * jv * jv;
    }
    for (int c = g.lane + G; c < ncon; c += G) {
      T D = w.cD[c];
      int nr = meta_dim4(w.cmeta[c]) ? 6 : 4;
      for (int r = 0; r < nr; r++) {
        T jv = w.Jv[c * 6 + r];
        T x = w.Jaref[c * 6 + r] + a * jv;
        const T Da = x < 0 ? D : (T)0;
        a1 += Da * jv * x; a2 += Da * jv * jv;
      }
    }
    for (int k = g.lane; k < s.nspec; k += G) {
      T jv = s.specJv[k], x = s.specJaref[k] + a * jv;
      const T Da = (s.specdof[k] < 0 || x < 0) ? s.specD[k] : (T)0;
      a1 += Da * jv * x; a2 += Da * jv * jv;
    }
    a1 = g.sum(a1);
    a2 = g.sum(a2);
    d1 = a1 + qg1 + 2 * a * qg2;
    d2 = a2 + 2 * qg2;
    if (it < 0) { if (d1 >= 0) { flat = true; break; } continue; }
    if (tabs(d1) < gtol) break;
    if (d1 < 0) lo = a; else hi = a;
  }
  *flat_out = flat;
  if (!flat && a != 0) {  // the move of the rows
    if (g.lane < ncon) {
#pragma unroll
      for (int r = 0; r < 6; r++)
        if (r < 4 || dim40) w.Jaref[g.lane * 6 + r] = ja0[r] + a * jv0[r];
    }
    for (int c = g.lane + G; c < ncon; c += G)
      for (int r = 0; r < (meta_dim4(w.cmeta[c]) ? 6 : 4); r++) w.Jaref[c * 6 + r] += a * w.Jv[c * 6 + r];
    for (int k = g.lane; k < s.nspec; k += G) s.specJaref[k] += a * s.specJv[k];
  }
  return a;
}

template <class T, int G>
MM_HDX void solve(const Grp<G>& g, Scratch<T>& s, const ModelDev<T>& md, Work<T>& w) {
  MM_IN_SHARED(&s);
  MM_IN_GLOBAL(&md);
  MM_IN_GLOBAL(w.cpos); MM_IN_GLOBAL(w.cn); MM_IN_GLOBAL(w.ct1); MM_IN_GLOBAL(w.cdist); MM_IN_GLOBAL(w.cD); MM_AREF_GLOBAL(w.aref);
  MM_ROWS_GLOBAL(w.Jaref); MM_ROWS_GLOBAL(w.Jv); MM_IN_GLOBAL(w.cmeta); MM_IN_GLOBAL(w.surv);
  analyse_coupling<T, G>(g, s);
  const T scale_inv = md.meaninertia * (T)NV;
  const T scale = (T)1 / scale_inv;
  const T tol = (T)1e-8;
  int ncon = s.ncon;
  int changed;
  // warmstart selection: cost at qacc_smooth (which = 0, rows kept in Jv) vs cost at qacc_warmstart
  // (which = 1, rows kept in Jaref)
  T cost_sm = 0, cost_ws = 0;
  {
    // qacc_warmstart arrives from the global state (parked in `search`, free until the Newton loop)
    for (int i = g.lane; i < NV; i += G) s.search[i] = (T)s.warm_g[i];
    g.sync();
    warm_rows<T, G>(g, s, md, w, s.as, s.search, &cost_sm, &cost_ws);
    mulM<T, G>(g, s, md, s.search, s.Ma);
    g.sync();
    for (int i = g.lane; i < NV; i += G) cost_ws += (T)0.5 * (s.Ma[i] - s.fs[i]) * (s.search[i] - s.as[i]);
    cost_sm = g.sum(cost_sm);
    cost_ws = g.sum(cost_ws);
    g.sync();
  }
  if (cost_ws < cost_sm) {
    for (int i = g.lane; i < NV; i += G) s.qacc[i] = s.search[i];
  } else {
    for (int i = g.lane; i < NV; i += G) s.qacc[i] = s.as[i];
    for (int c = g.lane; c < ncon; c += G) for (int r = 0; r < 6; r++) w.Jaref[c * 6 + r] = w.Jv[c * 6 + r];
    for (int k = g.lane; k < s.nspec; k += G) s.specJaref[k] = s.specJv[k];
    g.sync();
    mulM<T, G>(g, s, md, s.qacc, s.Ma);
  }
  g.sync();
  // Newton iterations.  One evaluation site per iteration: constraint update (active set, cost, forces),
  // convergence test of the previous move, then - only when the active set changed - the per-pair
  // blocks and the factorisation of H.
  int specbits = -1;
  int iter = 0;
  T cost = 0, a = 0;
  bool first = true, finished = false;
  while (true) {
    // (A) constraint update at the current point: active set, cost, forces; convergence test of the last move
    if (!finished) {
      T oldcost = cost;
      int chg;
      cost = update_constraint<T, G>(g, s, w, first, &chg);  // the first pass also builds the per-pair blocks
      changed = chg;
      int sb = 0;
#pragma unroll 1
      for (int k = 0; k < s.nspec; k++) if (s.specdof[k] < 0 || s.specJaref[k] < 0) sb |= 1 << k;
      if (sb != specbits) { changed = 1; specbits = sb; }
      T ga = 0, gn = 0;
      for (int i = g.lane; i < NV; i += G) {
        ga += (T)0.5 * (s.Ma[i] - s.fs[i]) * (s.qacc[i] - s.as[i]);
        T gr = s.Ma[i] - s.fs[i] - s.fc[i];
        gn += gr * gr;
      }
      cost += g.sum(ga);
      gn = tsqrt(g.sum(gn));
      if (!first) {
        iter++;
        T improvement = scale * (oldcost - cost), gradient = scale * gn;
#if defined(MM_TRACE) && !defined(__CUDA_ARCH__)
        printf("  it %d alpha %.6e cost %.12e impr %.3e grad %.3e changed %d\n", iter, (double)a, (double)cost, (double)improvement, (double)gradient, changed);
#endif
        if (improvement < tol || gradient < tol || iter >= 100) finished = true;
      }
    }
    if (finished) break;
    // (B) per-pair blocks + factorisation of H, only when the active set changed
    if (!finished && (first || changed)) {
      int dummy;
      // (`rebuild` is always true here; reading it from the scratch keeps ONE copy of update_constraint in the kernel
      // instead of a second one specialised for the constant - 20 KB of instruction footprint)
      const bool rebuild = s.n_il > 0;
      if (!first) update_constraint<T, G>(g, s, w, rebuild, &dummy);
      build_factor_H<T, G>(g, s, md);
    }
    first = false;
    // (C) Newton direction and the quantities of the line search
    T sn = 0, qg1 = 0, qg2 = 0;
    if (!finished) {
      for (int i = g.lane; i < NV; i += G) s.search[i] = -(s.Ma[i] - s.fs[i] - s.fc[i]);
      g.sync();
      solve_H<T, G>(g, s, s.search);
      mulM<T, G>(g, s, md, s.search, s.Mv);
      for (int i = g.lane; i < NV; i += G) {
        sn += s.search[i] * s.search[i];
        qg1 += s.search[i] * (s.Ma[i] - s.fs[i]);
        qg2 += (T)0.5 * s.search[i] * s.Mv[i];
      }
      sn = tsqrt(g.sum(sn)); qg1 = g.sum(qg1); qg2 = g.sum(qg2);
      if (sn < (T)MINVAL_D) finished = true;
    }
    // (D) rows of the direction, exact line search and the move
    if (!finished) {
      T gtol = tol * (T)0.01 * sn * scale_inv;
      bool flat;
      a = dir_search<T, G>(g, s, w, qg1, qg2, gtol, &flat);  // (rows of the direction, line search, move of the rows)
      if (flat || a == 0) finished = true;
      else {
        for (int i = g.lane; i < NV; i += G) { s.qacc[i] += a * s.search[i]; s.Ma[i] += a * s.Mv[i]; }
        g.sync();
      }
    }
  }
  if (g.lane == 0) s.niter = iter;
  for (int i = g.lane; i < NV; i += G) s.warm_g[i] = (double)s.qacc[i];
  g.sync();
}

// full forward at the current (qpos, qvel, ctrl): everything mj_forward computes that the path needs
template <class T, int G>
MM_HDN void forward(const Grp<G>& g, Scratch<T>& s, const ModelDev<T>& md, Work<T>& w) {
  long long t0 = MM_T0(s);
  fk<T, G>(g, s, md);
  dyn_smooth<T, G>(g, s, md);
  MM_TICK(s, g, 0, t0);
  t0 = MM_T0(s);
  collide<T, G>(g, s, md, w);
  t0 = MM_T0(s);
  make_constraints<T, G>(g, s, md, w);
  MM_TICK(s, g, 4, t0);
  t0 = MM_T0(s);
  solve<T, G>(g, s, md, w);
  MM_TICK(s, g, 5, t0);

}

// ------------------------------------------------------------------------------------------------
// implicitfast (A7)
// ------------------------------------------------------------------------------------------------
template <class T, int G>
MM_HDX void integrate(const Grp<G>& g, Scratch<T>& s, const ModelDev<T>& md) {
  MM_IN_SHARED(&s);
  MM_IN_GLOBAL(&md);
  const T h = md.timestep;
  T* MH = s.H;
  T* acc = s.Mv;
  for (int e = g.lane; e < NROB * NROB; e += G) {
    int i = e / NROB, j = e % NROB;
    T v = s.Mr[e];
    if (i == j) {
      v += h * md.damping[i];
      if (i < NARM && !s.actsat[i]) v -= h * md.act_b2[i];
    }
    if (i >= 7 && j >= 7 && !s.actsat[7]) v -= h * md.act_b2[7] * (T)0.25;
    MH[e] = v;
  }
  for (int i = g.lane; i < NV; i += G) {
    T f = s.fs[i] + s.fc[i];
    acc[i] = i < NROB ? f : f / (((i - 9) % 6) < 3 ? md.cube_mass : md.cube_inertia);
  }
  g.sync();
  chol_factor<T, G>(g, MH, NROB);
  chol_solve<T, G>(g, MH, NROB, acc);
  for (int i = g.lane; i < NV; i += G) s.qvel[i] += h * acc[i];
  g.sync();
  for (int i = g.lane; i < NROB + 3; i += G) {
    if (i < NROB) s.qpos[i] += h * s.qvel[i];
    else {
      int j = i - NROB;
      T* q = s.qpos + 9 + 7 * j;
      const T* v = s.qvel + 9 + 6 * j;
      for (int a = 0; a < 3; a++) q[a] += h * v[a];
      T wn = tsqrt(dot3(v + 3, v + 3));
      if (wn * h >= (T)MINVAL_D) {
        T sn, cs;
        tsincos(wn * h * (T)0.5, &sn, &cs);
        T inv = sn / wn;
        T dq[4] = {cs, v[3] * inv, v[4] * inv, v[5] * inv}, *qq = q + 3, r[4];
        r[0] = qq[0] * dq[0] - qq[1] * dq[1] - qq[2] * dq[2] - qq[3] * dq[3];
        r[1] = qq[0] * dq[1] + qq[1] * dq[0] + qq[2] * dq[3] - qq[3] * dq[2];
        r[2] = qq[0] * dq[2] - qq[1] * dq[3] + qq[2] * dq[0] + qq[3] * dq[1];
        r[3] = qq[0] * dq[3] + qq[1] * dq[2] - qq[2] * dq[1] + qq[3] * dq[0];
        T nn = (T)1 / tsqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
        for (int a = 0; a < 4; a++) qq[a] = r[a] * nn;
      }
    }
  }
  g.sync();
}

// ------------------------------------------------------------------------------------------------
// DLS IK (controller.py:87-137) on the kinematics left by the LAST position stage (staleness quirk,
// SURVEY 3.3) and the CURRENT joint angles.  Writes ctrl[0..6].
// ------------------------------------------------------------------------------------------------
template <class T>
MM_HDN void orientation_error(const T* Rc, T* out) {
  // R_err = TARGET_ORI * Rc^T with TARGET_ORI = [[0,1,0],[1,0,0],[0,0,-1]]  (controller.py:12-18, 21-43)
  T E[9];
  for (int j = 0; j < 3; j++) { E[j] = Rc[3 * j + 1]; E[3 + j] = Rc[3 * j]; E[6 + j] = -Rc[3 * j + 2]; }
  T tv = tclamp((E[0] + E[4] + E[8] - 1) / 2, (T)-1, (T)1);
  T ang = tacos(tv);
  if (ang < (T)1e-6) { out[0] = out[1] = out[2] = 0; return; }
  T sc = ang / (2 * tsin(ang));
  out[0] = (E[7] - E[5]) * sc; out[1] = (E[2] - E[6]) * sc; out[2] = (E[3] - E[1]) * sc;
}

template <class T, int G>
MM_HDX void ik(const Grp<G>& g, Scratch<T>& s, const ModelDev<T>& md) {
  MM_IN_SHARED(&s);
  MM_IN_GLOBAL(&md);
  // spread over the lanes; scratch in the (stale) H region: J [6][7] at 0, z [7] at 42, e / b [6] at 49,
  // A = J J^T + damping (6x6, row stride NV) at 60, dq [7] at 60 + 6 NV
  T* J = s.H;
  T* z = s.H + 42;
  T* b = s.H + 49;
  T* A = s.H + 60;
  T* dq = s.H + 60 + 6 * NV;
  const T* ee = s.bpos[DB_HAND];
  for (int e = g.lane; e < 42; e += G) {
    int d = e / NARM, i = e % NARM;
    const T* S = s.S[i];
    T v;
    if (d < 3) {
      int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
      v = S[d1] * ee[d2] - S[d2] * ee[d1] + S[3 + d];  // (a x ee + p x a)_d
    } else v = S[d - 3];
    J[e] = v;
  }
  for (int i = g.lane; i < NARM; i += G) z[i] = (T)0.5 * (md.home[i] - s.qpos[i]);
  if (g.lane == 0) {
    T er[6];
    for (int d = 0; d < 3; d++) er[d] = s.target[d] - ee[d];
    orientation_error(s.bR[DB_HAND], er + 3);
    for (int d = 0; d < 6; d++) b[d] = er[d];
  }
  g.sync();
  for (int e = g.lane; e < 21 + 6; e += G) {
    if (e < 21) {
      int i, j;
      tri_rc<T>(e, &i, &j);
      T a = 0;
      for (int c = 0; c < NARM; c++) a += J[i * NARM + c] * J[j * NARM + c];
      A[i * NV + j] = a + (i == j ? (T)1e-3 : (T)0);
    } else {
      int i = e - 21;
      T jz = 0;
      for (int c = 0; c < NARM; c++) jz += J[i * NARM + c] * z[c];
      dq[i] = jz;  // J z, subtracted from the error below
    }
  }
  g.sync();
  if (g.lane == 0) {
    for (int i = 0; i < 6; i++) b[i] -= dq[i];
    chol6_local(A);
    solve6_local(A, b);
  }
  g.sync();
  T mine = 0;
  for (int c = g.lane; c < NARM; c += G) {
    T a = z[c];
    for (int i = 0; i < 6; i++) a += J[i * NARM + c] * b[i];
    dq[c] = a;
    mine += a * a;
  }
  T nn = tsqrt(g.sum(mine));
  T sc = nn > (T)5 ? (T)5 / nn : (T)1;
  for (int c = g.lane; c < NARM; c += G) s.ctrl[c] = tclamp(s.qpos[c] + dq[c] * sc, md.jnt_lo[c], md.jnt_hi[c]);
  g.sync();
}

}  // namespace mm
