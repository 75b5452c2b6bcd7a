// Per-precision model tables the kernels read (global memory, read-only, L1 / L2 resident).
// Filled on the host from the generated double tables (model_gen.h / model_dev_gen.h).
#pragma once
#include "mm_group.h"

namespace mm {

constexpr int NQ = 30, NV = 27, NU = 8, NARM = 7, NROB = 9;
constexpr int NDB = 13;      // dynamic bodies: link1..7, hand, left/right finger, 3 cubes
constexpr int NGEOM = 47;    // collision geoms in the oracle's numbering (tools/modelc.py)
constexpr int NPAIRC = 780;  // candidate geom pairs after the static filters
constexpr int NHULLV = 1234; // convex-hull vertices of the collision meshes
enum { GT_PLANE = 0, GT_CYL = 5, GT_BOX = 6, GT_HULL = 7 };
constexpr int DB_HAND = 7, DB_LF = 8, DB_RF = 9, DB_CUBE0 = 10;
constexpr int CLS_CUBE0 = 10;

// Collision geometry (read-only, global memory; the same for every env so it stays in L1/L2).
// Geom frames are aligned with their body frame: world pose = (bpos + bR * pos, bR); static geoms
// (body < 0) carry world positions.
template <class T>
struct GeomDev {
  T size[NGEOM][3], pos[NGEOM][3], bc[NGEOM][3], rbound[NGEOM], invw[NGEOM];
  int type[NGEOM], body[NGEOM], cls[NGEOM], cube[NGEOM], obst[NGEOM], vadr[NGEOM], vnum[NGEOM];
  short pair[NPAIRC][2];  // sorted by (class of geom1, class of geom2): contacts of a body pair are contiguous
  // first broad-phase level, one record per candidate (so that the test needs no load that depends on another load):
  // kind 0 plane vs bounded geom, 1 static axis-aligned box vs bounding sphere, 2 sphere vs sphere, 3 never kept
  // (plane vs cylinder); rsum = bounding radius of geom b (kinds 0, 1) or rbound[a] + rbound[b] (kind 2)
  unsigned char pairkind[NPAIRC];
  unsigned char pairflags[NPAIRC];  // bit 0: box / plane narrow phase (plane-box, box-box, plane-hull), bit 1: convex stage (GJK / EPA)
  T pairrs[NPAIRC];
  T hull[NHULLV][3];
};

template <class T>
struct ModelDev {
  const GeomDev<T>* geom;
  T link_pos[10][3];
  T link_R[10][9];
  T jnt_lo[NROB], jnt_hi[NROB], armature[NROB], damping[NROB], dof_invw[NROB];
  T ib_mass[NROB], ib_com[NROB][3], ib_inertia[NROB][6];
  T act_gain[NU], act_b1[NU], act_b2[NU], ctrl_lo[NU], ctrl_hi[NU], frc_lo[NU], frc_hi[NU];
  T cube_mass, cube_inertia;
  T eq_solref[2], eq_solimp[5], eq_invw;
  T key_qpos[NQ], key_ctrl[NU], home[NARM];
  T timestep, gravity_z, meaninertia;
};

}  // namespace mm

#ifdef MM_MODEL_HOST_FILL
#include <cstdio>
#include <cstdlib>
#ifndef MM_CONST
#define MM_CONST static const
#endif
#include "model_dev_gen.h"
#include "model_gen.h"
namespace mm {
template <class T>
static void fill_model(ModelDev<T>& m) {
  m.geom = nullptr;
  for (int i = 0; i < 10; i++) {
    for (int k = 0; k < 3; k++) m.link_pos[i][k] = (T)mmd_link_pos[i][k];
    for (int k = 0; k < 9; k++) m.link_R[i][k] = (T)mmd_link_R[i][k];
  }
  for (int j = 0; j < NROB; j++) {
    m.jnt_lo[j] = (T)mm_jnt_range[j][0];
    m.jnt_hi[j] = (T)mm_jnt_range[j][1];
    m.armature[j] = (T)mm_jnt_armature[j];
    m.damping[j] = (T)mm_jnt_damping[j];
    m.dof_invw[j] = (T)mm_dof_invweight0[j];
    m.ib_mass[j] = (T)mmd_ib_mass[j];
    for (int k = 0; k < 3; k++) m.ib_com[j][k] = (T)mmd_ib_com[j][k];
    for (int k = 0; k < 6; k++) m.ib_inertia[j][k] = (T)mmd_ib_inertia[j][k];
  }
  for (int a = 0; a < NU; a++) {
    m.act_gain[a] = (T)mm_act_gain[a];
    m.act_b1[a] = (T)mm_act_bias[a][1];
    m.act_b2[a] = (T)mm_act_bias[a][2];
    m.ctrl_lo[a] = (T)mm_act_ctrlrange[a][0];
    m.ctrl_hi[a] = (T)mm_act_ctrlrange[a][1];
    m.frc_lo[a] = (T)mm_act_forcerange[a][0];
    m.frc_hi[a] = (T)mm_act_forcerange[a][1];
  }
  m.cube_mass = (T)mm_body_mass[16];
  m.cube_inertia = (T)mm_body_inertia[16][0];
  for (int k = 0; k < 2; k++) m.eq_solref[k] = (T)mm_eq_solref[k];
  for (int k = 0; k < 5; k++) m.eq_solimp[k] = (T)mm_eq_solimp[k];
  if (mm_eq_solimp[4] != 2.0) { std::fprintf(stderr, "mm_model: solimp power %g is not supported (the kernels implement power 2)\n", mm_eq_solimp[4]); std::abort(); }
  m.eq_invw = (T)(mm_dof_invweight0[MM_EQ_DOF1] + mm_dof_invweight0[MM_EQ_DOF2]);
  for (int k = 0; k < NQ; k++) m.key_qpos[k] = (T)mm_key_qpos[k];
  for (int k = 0; k < NU; k++) m.key_ctrl[k] = (T)mm_key_ctrl[k];
  static const double home[7] = {1.5708, -0.2, 0.0, -2.1, 0.0, 1.8, 0.785};
  for (int k = 0; k < NARM; k++) m.home[k] = (T)home[k];
  m.timestep = (T)MM_TIMESTEP;
  m.gravity_z = (T)mm_gravity[2];
  m.meaninertia = (T)MM_MEANINERTIA;
}

template <class T>
static void fill_geom(GeomDev<T>& g) {
  for (int i = 0; i < NGEOM; i++) {
    for (int k = 0; k < 3; k++) { g.size[i][k] = (T)mmd_g_size[i][k]; g.pos[i][k] = (T)mmd_g_pos[i][k]; g.bc[i][k] = (T)mmd_g_bc[i][k]; }
    g.rbound[i] = (T)mmd_g_rbound[i]; g.invw[i] = (T)mmd_g_invw[i];
    g.type[i] = mmd_g_type[i]; g.body[i] = mmd_g_body[i]; g.cls[i] = mmd_g_class[i]; g.cube[i] = mmd_g_cube[i];
    g.obst[i] = mmd_g_obst[i]; g.vadr[i] = mmd_g_vadr[i]; g.vnum[i] = mmd_g_vnum[i];
    if (mmd_g_vnum[i] > 160) { std::fprintf(stderr, "mm_model: hull %d has %d vertices (the support scan holds at most 160 in registers)\n", i, mmd_g_vnum[i]); std::abort(); }
  }
  for (int c = 0; c < NPAIRC; c++) {
    int a = mmd_pair[c][0], b = mmd_pair[c][1];
    g.pair[c][0] = (short)a; g.pair[c][1] = (short)b;
    if (g.type[a] == GT_PLANE) { g.pairkind[c] = g.type[b] == GT_CYL ? 3 : 0; g.pairrs[c] = g.rbound[b]; }
    else if (g.type[a] == GT_BOX && g.body[a] < 0) { g.pairkind[c] = 1; g.pairrs[c] = g.rbound[b]; }
    else { g.pairkind[c] = 2; g.pairrs[c] = g.rbound[b] + g.rbound[a]; }
    {
      int ta = g.type[a], tb = g.type[b];
      int narrow = (tb == GT_BOX && (ta == GT_PLANE || ta == GT_BOX)) || (ta == GT_PLANE && tb == GT_HULL);
      int convex = !(ta == GT_PLANE || (ta == GT_BOX && tb == GT_BOX));
      g.pairflags[c] = (unsigned char)(narrow | (convex << 1));
    }
  }
  for (int v = 0; v < NHULLV; v++) for (int k = 0; k < 3; k++) g.hull[v][k] = (T)mm_hull[v][k];
}
}  // namespace mm
#endif
