// ORACLE - TEST INFRASTRUCTURE ONLY (see engine.h).
// CPU restatement of the reference's Python layer above the engine, each function citing the
// reference file:line it follows.  Validated in this container against the reference's own Python
// running on oracle/fake_mujoco.py (tools/make_golden.py -> tests/golden/).
#pragma once
#include "engine.h"

namespace orc {

enum ActionMode { ABS_POS = 0, EE_POS_QUAT_G = 1, EE_POS_ROT6D_G = 2, EE_POS_QUAT_G_REL = 3, EE_POS_ROT6D_G_REL = 4 };
enum RewardType { DENSE = 0, SPARSE = 1, STAGED = 2 };
constexpr int BODY_HAND = 9, BODY_TABLE = 12, BODY_BIN0 = 13, BODY_OBJ0 = 16;
constexpr int OBS_STATE_DIM = 53, OBS_FULL_DIM = 85;

struct Env {
  Data d;
  int action_mode = EE_POS_QUAT_G_REL, reward_type = DENSE, max_episode_steps = 500;
  // episode (gym_env.py:111-133)
  int step_count = 0, obj_idx = 0, bin_idx = 0;
  double init_pos[3], init_R[9];
  float tgt_obj_kp[2], tgt_bin_kp[2];
  bool has_grasped = false, has_lifted = false, above_target = false, has_placed = false, hwm_set = false;
  double hwm[5] = {0, 0, 0, 0, 0};
  // FSM (pick_and_place.py:100-105); single-task list (obj_idx, bin_idx)
  int fsm_state = 1, task_index = 0, settle_counter = 0, gripper_open = 1, has_target = 0;
  double target[3] = {0, 0, 0}, transit_end[3] = {0, 0, 0};
};

void env_reset(Env& e, const double* obj_xy /*6 or null*/, int obj_idx, int bin_idx,
               const double* yaw_cs = nullptr /*6: (cos, sin)(theta/2) per cube, randomization.py:55-62*/);
void decode_action(const Env& e, const float* action, double target[3], float* gripper);
void ik_compute(const Data& d, const double target[3], double q_target[7]);
bool ik_reached(const Data& d, const double target[3]);
void env_step(Env& e, const float* action, float* obs /*85*/, double* reward, int* terminated, int* truncated,
              int* success, float* reward_components /*6 or null*/);
void env_obs(const Env& e, float* obs /*85*/);
void fsm_reset(Env& e);
void fsm_plan(Env& e, int n_steps);
void fsm_action(const Env& e, float action[4]);
void orientation_error(const double* R_cur, const double* R_tgt, double out[3]);

}  // namespace orc
