// ORACLE - TEST INFRASTRUCTURE ONLY.  Not a product path: only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build, link or call anything in oracle/.
//
// CPU (FP64) restatement of the third-party engine the reference calls on its step hot path:
// MuJoCo 3.5.0 (`mujoco>=3.5.0`, /root/reference/pyproject.toml:11, uv.lock:985-986), restricted to
// the reference's own scene (mujoco_manip/data/pick_and_place_scene.xml + franka_emika_panda/panda.xml).
// MuJoCo's C source is NOT vendored in /root/reference and the wheel is not installable here, so this
// follows MuJoCo's published algorithm ("Computation" chapter + SURVEY.md Appendix A) and is anchored
// on the reference's call sites:
//   mj_step               mujoco_manip/env.py:119-121, gym_env.py:558
//   mj_forward            mujoco_manip/env.py:117,161, gym_env.py:560
//   mj_jac                mujoco_manip/controller.py:101-108
//   mj_resetDataKeyframe  mujoco_manip/env.py:116
// PARITY UNPINNED against real MuJoCo trajectories (no MuJoCo in this image; the reference ships no
// golden trajectories).  Everything above the engine (IK, FSM, decode, reward, RNG) IS pinned: the
// reference's own Python runs on top of this engine through oracle/fake_mujoco.py (tests/golden).
#pragma once
#include <vector>

#include "../mujoco_manip_b200/csrc/model_gen.h"

namespace orc {

constexpr int NQ = MM_NQ, NV = MM_NV, NU = MM_NU, NBODY = MM_NBODY, NGEOM = MM_NGEOM, NJNT = MM_NJNT;
constexpr double MINVAL = 1e-15;

struct Contact {
  double dist;
  double pos[3];
  double frame[9];  // row 0 = normal (geom1 -> geom2), rows 1,2 = tangents
  double friction[5];
  double solref[2];
  double solimp[5];
  int dim;
  int geom1, geom2;
  int efc_address;
};

struct Stats {  // work counters exported for the roofline FLOP model
  long long substeps, ncon, nefc, newton_iters, ls_evals, narrow_tests, ccd_tests;
  int max_ncon, max_nefc, max_newton;
  long long forwards, nnzJ, nnzJ2;  // forward passes; nonzeros of the constraint Jacobian and sum of squared row nonzeros
};

struct Data {
  // state
  double qpos[NQ], qvel[NV], ctrl[NU], qacc_warmstart[NV], time;
  // position-dependent
  double xpos[NBODY][3], xquat[NBODY][4], xmat[NBODY][9], xipos[NBODY][3];
  double xanchor[NJNT][3], xaxis[NJNT][3];
  double geom_xpos[NGEOM][3], geom_xmat[NGEOM][9];
  double cam_xpos[3][3], cam_xmat[3][9];
  double M[NV * NV];
  // forces
  double qfrc_bias[NV], qfrc_passive[NV], qfrc_actuator[NV], qfrc_smooth[NV], qacc_smooth[NV];
  double qfrc_constraint[NV], qacc[NV], actuator_force[NU];
  int act_saturated[NU];
  // constraints
  std::vector<Contact> contact;
  int ncon, nefc;
  std::vector<double> efc_J, efc_pos, efc_D, efc_R, efc_aref, efc_force, efc_vel;
  std::vector<int> efc_type;  // 0 equality, 1 limit, 2 contact
  int solver_niter;
  Stats stats;
  int flags;  // bit0: disable mesh/cylinder (general convex) collisions
};

void reset_keyframe(Data& d);
void forward(Data& d);
void step(Data& d);
void jac(const Data& d, double* jacp, double* jacr, const double point[3], int body);  // 3 x NV each
void kinematics_only(Data& d);

}  // namespace orc
