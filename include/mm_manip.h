/* mm_manip.h - C ABI of the B200 batched pick-and-place step.
 *
 * Drop-in boundary for the hot path the reference reaches through the `mujoco` Python module
 * (libmujoco C API).  Each entry point names the reference call it replaces; N environments are
 * stepped per call instead of one MjData.
 *
 * Conventions
 *  - every array pointer is a DEVICE pointer owned by the caller (PyTorch on the Python side),
 *    env-major and contiguous; state is stored in FP64 whatever the compute precision;
 *  - every call takes a cudaStream_t (as void*), enqueues work and returns without synchronising
 *    (except the *_host convenience call, which is the end-to-end path and synchronises the stream);
 *  - return value 0 = ok, negative = error; mm_last_error() gives a thread-local message;
 *  - nothing is allocated after mm_create; a handle is bound to one device and is not thread-safe
 *    (same rule as one MjData per thread in the reference).
 */
#ifndef MM_MANIP_H
#define MM_MANIP_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MM_NQ 30
#define MM_NV 27
#define MM_NU 8
#define MM_OBS_DIM 85        /* 53 state floats + 32 keypoint floats */
#define MM_ACTION_STRIDE 10  /* actions are [N,10] float32, unused tail ignored */

/* action modes: mujoco_manip/gym_env.py:30-36 */
enum { MM_ABS_POS = 0, MM_EE_POS_QUAT_G = 1, MM_EE_POS_ROT6D_G = 2, MM_EE_POS_QUAT_G_REL = 3, MM_EE_POS_ROT6D_G_REL = 4 };
/* reward types: mujoco_manip/gym_env.py:436-470 */
enum { MM_REWARD_DENSE = 0, MM_REWARD_SPARSE = 1, MM_REWARD_STAGED = 2 };

typedef struct mm_handle mm_handle;

typedef struct {
  int32_t num_envs;           /* envs owned by this process / GPU */
  int32_t device;             /* CUDA device ordinal */
  int32_t precision;          /* 0 = FP64 arithmetic (parity mode), 1 = FP32 arithmetic */
  int32_t group;              /* lanes cooperating on one env: 8, 16 or 32 */
  int32_t reward_type;        /* MM_REWARD_* */
  int32_t max_episode_steps;  /* truncation limit (gym_env.py:70, constants.py:27) */
} mm_config;

/* Per-env persistent state = what one MjData + the Python objects around it hold in the reference
 * (SURVEY.md Appendix B).  All [N, k] env-major. */
typedef struct {
  double* qpos;        /* [N,30] MjData.qpos                                   */
  double* qvel;        /* [N,27] MjData.qvel                                   */
  double* ctrl;        /* [N,8]  MjData.ctrl  (robot.py:65-79)                 */
  double* warm;        /* [N,27] MjData.qacc_warmstart                         */
  double* tinit;       /* [N,12] initial EE pose pos(3)+R(9) (gym_env.py:245)  */
  double* eepose;      /* [N,12] EE pose after the last forward (robot.py:50-58) */
  double* fsm_f;       /* [N,6]  FSM target + transit_end (pick_and_place.py:103-104) */
  double* hwm;         /* [N,5]  staged reward high-water marks (gym_env.py:133) */
  double* kin;         /* [N,18] qpos of the last position stage (arm+fingers 9, cube xyz 9): what data.xpos /
                                 mj_jac describe between mj_step calls (SURVEY 3.3 staleness quirk) */
  int32_t* step_count; /* [N]    gym_env.py:111                                */
  int32_t* task;       /* [N,2]  object index, bin index (constants.py:3-4)    */
  int32_t* fsm_i;      /* [N,5]  state 1..11, task_index, settle_counter, gripper_open, has_target */
  int32_t* fsm_tasks;  /* [N,20] the FSM's own task list (pick_and_place.py:91, independent of `task`): count, then up to 9
                                 (object index, bin index) pairs; mm_reset arms it with the env's single task */
  int32_t* flags;      /* [N]    staged stickies bits0-3, hwm-set bit4 (gym_env.py:129-132) */
  int32_t* diag;       /* [N,4]  ncon, newton iters (last forward), overflow bits, non-finite resets */
} mm_state;

typedef struct {
  float* obs;                 /* [N,85] packed observation, layout in INTEGRATION.md */
  float* reward;              /* [N]   */
  uint8_t* terminated;        /* [N]   */
  uint8_t* truncated;         /* [N]   */
  uint8_t* success;           /* [N]   info["success"]                          */
  float* reward_components;   /* [N,6] info["reward_components"] or NULL        */
} mm_step_out;

/* replaces MjModel.from_xml_path + MjData(model) (env.py:53-67,97) for N envs */
int mm_create(const mm_config* cfg, mm_handle** out);
void mm_destroy(mm_handle* h);
const char* mm_last_error(void);
/* bytes of device workspace mm_create allocates for this config */
size_t mm_workspace_bytes(const mm_config* cfg);

/* replaces reset_to_keyframe = mj_resetDataKeyframe + mj_forward (env.py:104-117), the qpos write of
 * randomize_object_positions + mj_forward (randomization.py:52-65, env.py:160-161) and the episode
 * bookkeeping of PickPlaceGymEnv.reset (gym_env.py:477-534).
 *   mask    [N] uint8 or NULL (NULL = all envs)
 *   obj_xy  [N,6] double (x,y of the three cubes) or NULL (keyframe placement)
 *   task    [N,2] int32 (object index, bin index)
 *   obs     [N,85] float or NULL */
int mm_reset(mm_handle* h, const mm_state* st, const uint8_t* mask, const double* obj_xy, const int32_t* task,
             float* obs, void* stream);

/* replaces PickPlaceGymEnv.step (gym_env.py:536-581): decode_action, 16 x (IKController.compute =
 * mj_jac + DLS solve, controller.py:87-137; set ctrl; mj_step, env.py:119-121), mj_forward
 * (gym_env.py:560), reward / termination and the state observation.  actions: [N,10] float32. */
int mm_step(mm_handle* h, const mm_state* st, const float* actions, int action_mode, const mm_step_out* out,
            void* stream);

/* Same step with HOST buffers (pinned recommended): copies actions H2D, steps, copies obs / reward /
 * flags D2H and synchronises the stream.  This is the end-to-end path bench.py times as `e2e`. */
int mm_step_host(mm_handle* h, const mm_state* st, const float* h_actions, int action_mode, float* h_obs,
                 float* h_reward, uint8_t* h_terminated, uint8_t* h_truncated, uint8_t* h_success, void* stream);

/* mm_step_host without the closing synchronisation: the copies and the step are enqueued on `stream` and the call
 * returns; the host buffers are valid once the caller has synchronised the stream.  Lets the caller enqueue the
 * episode bookkeeping of the step (mm_post_step, mm_sample_episode, mm_reset - they follow the copies in stream
 * order) BEFORE waiting, so that the wait covers everything at once (PickPlaceVecEnv.step_host). */
int mm_step_host_async(mm_handle* h, const mm_state* st, const float* h_actions, int action_mode, float* h_obs,
                       float* h_reward, uint8_t* h_terminated, uint8_t* h_truncated, uint8_t* h_success, void* stream);

/* The device buffers mm_step_host stages its results in (owned by the handle): lets the caller run mm_post_step /
 * mm_reset on the results of a host-buffer step without another copy. */
int mm_host_staging(mm_handle* h, mm_step_out* out);

/* Engine-level calls for the physics-step-level loops of the reference (main.py:65-91,
 * pick_and_place.py:279-304, tests/test_controller.py): a combination of
 *   MM_OP_IK        IKController.compute(target) + set_arm_ctrl (controller.py:87-137, robot.py:65-71) on the
 *                   kinematics of the last position stage; target [N,3] double, world frame
 *   MM_OP_FORWARD   mj_forward (env.py:117)
 *   MM_OP_INTEGRATE with MM_OP_FORWARD: mj_step (env.py:119-121) */
enum { MM_OP_IK = 1, MM_OP_FORWARD = 2, MM_OP_INTEGRATE = 4 };
int mm_ops(mm_handle* h, const mm_state* st, int ops, const double* target, void* stream);

/* replaces PickAndPlaceTask.plan(n_steps) (pick_and_place.py:167-277) and the abs_pos action built
 * from it (scripts/generate_dataset.py:145-148).  actions_out: [N,10] float32 or NULL. */
int mm_fsm_plan(mm_handle* h, const mm_state* st, int n_steps, float* actions_out, void* stream);

/* replaces the per-env numpy Generator draws of reset (randomization.py:70-98 rejection sampler, then the
 * task draw of gym_env.py:516) for the vectorised env: Philox4x32-10 keyed by `seed`, counter =
 * (env_id_offset + env, episode_index[env], block) - see csrc/mm_rng.h; oracle/philox.py is the CPU
 * statement.  obj_xy [N,6] double out; task_draw [N] int32 out in [0,npool) or NULL; attempts [N] int32
 * out or NULL (0 = all 1000 attempts rejected, the case where the reference raises RuntimeError). */
int mm_sample_placements(mm_handle* h, uint64_t seed, int64_t env_id_offset, const int64_t* episode_index, double x_lo,
                         double x_hi, double y_lo, double y_hi, double min_separation, int32_t npool, double* obj_xy,
                         int32_t* task_draw, int32_t* attempts, void* stream);

/* replaces the `randomize_yaw=True` branch of randomize_object_positions (randomization.py:55-62): the cube
 * quaternions written with the placement become (cos(theta/2), 0, 0, sin(theta/2)).
 *   yaw_cs  [N,6] double, device: (cos, sin)(theta/2) of the three cubes, or NULL to switch the option off.
 * The pointer is kept (caller-owned) and read by every later mm_reset that is given obj_xy. */
int mm_set_placement_yaw(mm_handle* h, const double* yaw_cs);

/* Philox draw of theta = uniform(0, 2 pi) for the three cubes of every env (csrc/mm_rng.h: words 0,1 of block
 * 4(o+1)+3 of the (seed, env, episode) stream); theta [N,3] double out or NULL, yaw_cs [N,6] double out. */
int mm_sample_yaw(mm_handle* h, uint64_t seed, int64_t env_id_offset, const int64_t* episode_index, double* theta,
                  double* yaw_cs, void* stream);

/* One call for everything the vectorised env draws at a reset (the reference's order inside reset(): placement, then
 * the task, gym_env.py:496-517; yaw with the placement, randomization.py:55-62), for the envs where mask[e] != 0 (NULL =
 * all): the same Philox stream as mm_sample_placements / mm_sample_yaw, written in place, no host round trip.
 *   pool [npool,2] int32 (object, bin); task_mode 0 = pool[0], 1 = pool[global env id % npool], 2 = Philox draw
 *   randomize_xy / randomize_yaw: which arrays are drawn; task may be NULL; advance != 0: episode_index[e] += 1 after
 *   the draw; stats ([8] double or NULL): stats[6] counts placements whose 1000 attempts were all rejected. */
int mm_sample_episode(mm_handle* h, uint64_t seed, int64_t env_id_offset, int64_t* episode_index, const uint8_t* mask,
                      double x_lo, double x_hi, double y_lo, double y_hi, double min_separation, const int32_t* pool,
                      int32_t npool, int32_t task_mode, int32_t randomize_xy, int32_t randomize_yaw, int32_t advance,
                      double* obj_xy, int32_t* task, int32_t* attempts, double* yaw_theta, double* yaw_cs, double* stats,
                      void* stream);

/* Bookkeeping of a vectorised env after mm_step (no reference counterpart: the reference steps one env and leaves
 * auto-reset to the caller; this is gymnasium's vector-env convention on the device): ep_return [N] += reward;
 * reset_mask [N] (or NULL) = terminated | truncated; final_obs [N,85] (or NULL) = copy of out->obs; stats ([8] double
 * or NULL) accumulates episodes, successes, sum of lengths, sum of returns, non-finite state resets (diag[:,3], cleared),
 * finished episodes that hit a workspace overflow (diag[:,2]); with auto_reset the returns of finished envs restart. */
int mm_post_step(mm_handle* h, const mm_state* st, const mm_step_out* out, double* ep_return, uint8_t* reset_mask,
                 float* final_obs, double* stats, int32_t auto_reset, void* stream);

/* Measurement helper (no reference counterpart): FMA throughput of the CUDA cores in TFLOP/s (FP32 or FP64),
 * best of 5 timed launches - the measured denominator of the roofline bench.py reports. */
int mm_measure_fma_peak(int device, int fp64, double* tflops);

/* replaces get_actions (scripts/generate_dataset.py:56-80): the four action encodings of an expert abs_pos action
 * [N,10] (x, y, z, gripper): pose (target, TARGET_ORI) in the world frame and relative to the initial EE pose.
 * encodings [N,36] float32 = pos_quat_g (8) | pos_rot6d_g (10) | pos_quat_g_rel (8) | pos_rot6d_g_rel (10). */
int mm_expert_actions(mm_handle* h, const mm_state* st, const float* abs_actions, float* encodings, void* stream);

/* Load-aware scheduling of mm_step (no reference counterpart; results do not depend on it).  By default (batches of more
 * than 32 envs, MM_BALANCE=0 switches it off) every mm_step starts with a library kernel that sorts the envs by the busy
 * time of their previous step (heaviest first, dealt round-robin over the chunks of the launch plan): shorter tails of the
 * stage launches, and neighbouring warps run envs of similar cost.  `work` ([N] int32, device, or NULL = library-owned)
 * receives each env's busy time of the step (SM cycles / 256) and feeds that sort; `order` ([N] int32 device
 * permutation, or NULL = the library's own schedule) overrides it: which env each execution slot processes. */
int mm_set_schedule(mm_handle* h, const int32_t* order, int32_t* work);

/* Profiling aid: when `cycles` ([N,9] int64, device) is non-NULL every mm_step ADDS the SM clock cycles each env spent
 * in the stage kernels: [0] total, [1] stage A (IK, kinematics, dynamics, broad + box narrow phase), [2] convex stage
 * (GJK / EPA of the env's queued pairs, summed over the warps that ran them), [3] stage C (contact assembly, constraint
 * rows, Newton solver, integration); [4..8] unused.  NULL switches it off. */
int mm_set_cycle_buffer(mm_handle* h, long long* cycles);

/* Measurement helper: with `on`, every stage launch of mm_step is bracketed by CUDA events on the stream it is launched
 * on; mm_stage_times synchronises the device, returns the summed device time (ms) of the launches since the last call -
 * ms[0] stage A (IK, kinematics, dynamics, broad + box narrow phase), ms[1] convex stage (GJK / EPA queue), ms[2] stage C
 * (contact assembly, constraint rows, Newton solver, integration), ms[3] stage C of contact-rich envs (when that plan is
 * on) - and their counts.  Stages of different chunks overlap on different streams, so the sums can exceed wall time. */
int mm_stage_timing(mm_handle* h, int32_t on);
int mm_stage_times(mm_handle* h, double* ms, long long* launches);

/* number of kernels this handle has launched so far */
int mm_launch_count(mm_handle* h, long long* out);

#ifdef __cplusplus
}
#endif
#endif
