// Per-precision model table the kernels read (one copy per CTA in shared memory).
// Filled on the host from the generated double tables (model_gen.h / model_dev_gen.h).
#pragma once
#include "mm_group.h"

namespace mm {

constexpr int NQ = 30, NV = 27, NU = 8, NARM = 7, NROB = 9;
constexpr int NDB = 13;      // dynamic bodies: link1..7, hand, left/right finger, 3 cubes
constexpr int NBOX = 29;     // 10 pads, tabletop, 15 bin boxes, 3 cubes (oracle geom order)
constexpr int PLANE_ID = 29;
constexpr int NCAND = 279;
constexpr int DB_HAND = 7, DB_LF = 8, DB_RF = 9, DB_CUBE0 = 10;
constexpr int CLS_CUBE0 = 10;

template <class T>
struct ModelDev {
  T link_pos[10][3];
  T link_R[10][9];
  T jnt_lo[NROB], jnt_hi[NROB], armature[NROB], damping[NROB], dof_invw[NROB];
  T ib_mass[NROB], ib_com[NROB][3], ib_inertia[NROB][6];
  T act_gain[NU], act_b1[NU], act_b2[NU], ctrl_lo[NU], ctrl_hi[NU], frc_lo[NU], frc_hi[NU];
  T box_size[NBOX][3], box_pos[NBOX][3], box_invw[NBOX], box_rbound[NBOX];
  T cube_mass, cube_inertia;
  T eq_solref[2], eq_solimp[5], eq_invw;
  T key_qpos[NQ], key_ctrl[NU], home[NARM];
  T timestep, gravity_z, meaninertia;
  int box_body[NBOX], box_class[NBOX], box_cube[NBOX];
  short cand[NCAND][2];
};

}  // namespace mm

#ifdef MM_MODEL_HOST_FILL
#ifndef MM_CONST
#define MM_CONST static const
#endif
#include "model_dev_gen.h"
#include "model_gen.h"
namespace mm {
template <class T>
static void fill_model(ModelDev<T>& m) {
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
  for (int b = 0; b < NBOX; b++) {
    double r2 = 0;
    for (int k = 0; k < 3; k++) {
      m.box_size[b][k] = (T)mmd_box_size[b][k];
      m.box_pos[b][k] = (T)mmd_box_pos[b][k];
      r2 += mmd_box_size[b][k] * mmd_box_size[b][k];
    }
    m.box_rbound[b] = (T)std::sqrt(r2);
    m.box_invw[b] = (T)mmd_box_invw[b];
    m.box_body[b] = mmd_box_body[b];
    m.box_class[b] = mmd_box_class[b];
    m.box_cube[b] = mmd_box_cube[b];
  }
  m.cube_mass = (T)mm_body_mass[16];
  m.cube_inertia = (T)mm_body_inertia[16][0];
  for (int k = 0; k < 2; k++) m.eq_solref[k] = (T)mm_eq_solref[k];
  for (int k = 0; k < 5; k++) m.eq_solimp[k] = (T)mm_eq_solimp[k];
  m.eq_invw = (T)(mm_dof_invweight0[MM_EQ_DOF1] + mm_dof_invweight0[MM_EQ_DOF2]);
  for (int k = 0; k < NQ; k++) m.key_qpos[k] = (T)mm_key_qpos[k];
  for (int k = 0; k < NU; k++) m.key_ctrl[k] = (T)mm_key_ctrl[k];
  static const double home[7] = {1.5708, -0.2, 0.0, -2.1, 0.0, 1.8, 0.785};
  for (int k = 0; k < NARM; k++) m.home[k] = (T)home[k];
  m.timestep = (T)MM_TIMESTEP;
  m.gravity_z = (T)mm_gravity[2];
  m.meaninertia = (T)MM_MEANINERTIA;
  for (int c = 0; c < NCAND; c++) { m.cand[c][0] = (short)mmd_cand[c][0]; m.cand[c][1] = (short)mmd_cand[c][1]; }
}
}  // namespace mm
#endif
