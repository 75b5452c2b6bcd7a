"""Behavioural tests of the single-env API on the CUDA path, following the reference's own test-suite
(tests/test_gym_env.py, test_controller.py, test_pick_and_place.py, test_randomization.py: same
scenarios and tolerances), plus scripted-expert success parity with the oracle."""
import os

import numpy as np
import pytest

from hostlib import GOLDEN, reltol

pytestmark = pytest.mark.gpu

IDENT8 = np.array([0, 0, 0, 0, 0, 0, 1, 1], dtype=np.float32)
IDENT10 = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0, 1], dtype=np.float32)


def make(cuda_device, **kw):
    from mujoco_manip_b200 import PickPlaceGymEnv

    kw.setdefault("max_episode_steps", 50)
    return PickPlaceGymEnv(device=str(cuda_device), **kw)


def test_constructor_errors_and_spaces(cuda_device):
    from mujoco_manip_b200 import PickPlaceGymEnv

    with pytest.raises(ValueError):
        PickPlaceGymEnv(action_mode="joint_torque", device=str(cuda_device))
    dims = {"abs_pos": 4, "ee_pos_quat_g": 8, "ee_pos_rot6d_g": 10, "ee_pos_quat_g_rel": 8, "ee_pos_rot6d_g_rel": 10}
    for mode, d in dims.items():
        env = make(cuda_device, action_mode=mode)
        assert env.action_space.shape == (d,) and env.action_mode == mode
        assert env.action_space.low[-1] == 0.0 and env.action_space.high[-1] == 1.0
        if mode == "abs_pos":
            np.testing.assert_allclose(env.action_space.low, [-0.5, 0.0, 0.24, 0.0])
            np.testing.assert_allclose(env.action_space.high, [0.5, 0.8, 0.60, 1.0])
        else:
            assert np.isinf(env.action_space.low[0]) and np.isinf(env.action_space.high[0])
        env.close()


def test_observation_keys_shapes_dtypes(cuda_device):
    env = make(cuda_device)
    obs, info = env.reset(seed=0)
    assert info == {}
    exp = {"state": (11,), "state.ee.pos_quat_g": (8,), "state.ee.pos_rot6d_g": (10,), "state.ee.pos_quat_g_rel": (8,),
           "state.ee.pos_rot6d_g_rel": (10,), "target_bin_onehot": (3,), "target_obj_onehot": (3,),
           "keypoints_overhead": (7, 2), "keypoints_wrist": (7, 2), "target_obj_keypoints_overhead": (2,),
           "target_bin_keypoints_overhead": (2,)}
    assert set(obs.keys()) == set(env.observation_space.spaces.keys())
    for k, shp in exp.items():
        assert obs[k].shape == shp and obs[k].dtype == np.float32, k
        assert env.observation_space.spaces[k].shape == shp
    assert obs["image_overhead"].shape == (224, 224, 3) and obs["image_overhead"].dtype == np.uint8
    # relative pose of the initial EE is the identity
    np.testing.assert_allclose(obs["state.ee.pos_quat_g_rel"][:7], [0, 0, 0, 0, 0, 0, 1], atol=1e-6)
    # [DERIVED] keyframe keypoints
    np.testing.assert_allclose(obs["keypoints_overhead"][0], [0.60406095, 0.8121828], atol=1e-6)
    np.testing.assert_allclose(obs["keypoints_overhead"][6], [0.50000143, 0.8897216], atol=1e-6)
    env.close()


def test_seeded_reset_is_deterministic_and_task_sampling(cuda_device):
    env = make(cuda_device, tasks="all", randomize_objects=True)
    o1, _ = env.reset(seed=42)
    t1 = (env.obj_name, env.bin_name)
    p1 = env.pick_place_env.get_body_pos("obj_red")
    o2, _ = env.reset(seed=42)
    assert (env.obj_name, env.bin_name) == t1
    for k in o1:
        np.testing.assert_array_equal(o1[k], o2[k])
    # golden: default_rng(42) accepted placement of the red cube
    np.testing.assert_allclose(p1, [0.0575460, 0.3340858, 0.26], atol=1e-6)
    env.reset(seed=43)
    assert not np.allclose(env.pick_place_env.get_body_pos("obj_red"), p1)
    env.reset(options={"task": ("obj_blue", "bin_green")})
    assert (env.obj_name, env.bin_name) == ("obj_blue", "bin_green")
    obs, _ = env.reset(options={"task": ("obj_green", "bin_blue")})
    np.testing.assert_array_equal(obs["target_obj_onehot"], [0, 1, 0])
    np.testing.assert_array_equal(obs["target_bin_onehot"], [0, 0, 1])
    env.close()


def test_randomization_properties(cuda_device):
    env = make(cuda_device, randomize_objects=True)
    for s in range(20):
        env.reset(seed=s)
        P = np.array([env.pick_place_env.get_body_pos(n) for n in ("obj_red", "obj_green", "obj_blue")])
        assert np.all(np.abs(P[:, 2] - 0.26) < 0.01)
        assert np.all((P[:, 0] >= -0.2) & (P[:, 0] <= 0.2) & (P[:, 1] >= 0.30) & (P[:, 1] <= 0.45))
        for i in range(3):
            for j in range(i + 1, 3):
                assert np.linalg.norm(P[i, :2] - P[j, :2]) >= 0.08 - 1e-9
    with pytest.raises(ValueError):
        env.pick_place_env.get_body_pos("no_such_body")
    with pytest.raises(ValueError):
        env.pick_place_env.reset_to_keyframe("no_such_key")
    env.close()


@pytest.mark.parametrize("mode,ident", [("ee_pos_quat_g_rel", IDENT8), ("ee_pos_rot6d_g_rel", IDENT10)])
def test_identity_action_holds_and_translation_is_reached(cuda_device, mode, ident):
    env = make(cuda_device, action_mode=mode, task=("obj_red", "bin_red"))
    env.reset(seed=0)
    p0 = env.robot.ee_pos
    for _ in range(5):
        out = env.step(ident)
    assert len(out) == 5 and isinstance(out[1], float) and isinstance(out[2], bool) and isinstance(out[3], bool)
    assert np.linalg.norm(env.robot.ee_pos - p0) < 0.05
    a = ident.copy()
    a[2] = -0.1  # along the initial EE z axis (pointing down): 10 cm lower in the world
    target, g = env.decode_action(a)
    for _ in range(20):
        env.step(a)
    assert np.linalg.norm(env.robot.ee_pos - target) < 0.05
    assert env.step_count == 25
    env.close()


def test_gripper_command_and_truncation_and_sparse_reward(cuda_device):
    env = make(cuda_device, action_mode="abs_pos", reward_type="sparse", task=("obj_red", "bin_red"), max_episode_steps=4)
    env.reset(seed=0)
    obs, r, te, tr, info = env.step(np.array([0.0, 0.45, 0.5, 1.0], dtype=np.float32))
    assert env.robot.gripper_ctrl == 255.0 and obs["state"][3] == 1.0 and r in (0.0, 1.0) and not tr
    obs, r, te, tr, info = env.step(np.array([0.0, 0.45, 0.5, 0.0], dtype=np.float32))
    assert env.robot.gripper_ctrl == 0.0 and obs["state"][3] == 0.0
    obs, r, te, tr, info = env.step(np.array([0.0, 0.45, 0.5, 0.5], dtype=np.float32))  # strictly > 0.5 opens
    assert env.robot.gripper_ctrl == 0.0
    obs, r, te, tr, info = env.step(np.array([0.0, 0.45, 0.5, 0.0], dtype=np.float32))
    assert tr and not te and info["success"] is False
    obs, _ = env.reset()
    assert env.step_count == 0
    env.close()


def test_quat_and_rot6d_modes_agree_and_absolute_modes(cuda_device):
    from mujoco_manip_b200 import pose_utils as P

    e8 = make(cuda_device, action_mode="ee_pos_quat_g_rel", task=("obj_red", "bin_red"))
    e10 = make(cuda_device, action_mode="ee_pos_rot6d_g_rel", task=("obj_red", "bin_red"))
    e8.reset(seed=0)
    e10.reset(seed=0)
    a8, a10 = IDENT8.copy(), IDENT10.copy()
    a8[:3] = a10[:3] = [0.05, -0.03, -0.06]
    for _ in range(15):
        e8.step(a8)
        e10.step(a10)
    np.testing.assert_allclose(e8.robot.ee_pos, e10.robot.ee_pos, atol=0.01)
    # absolute SE(3) mode reaches a world target
    ea = make(cuda_device, action_mode="ee_pos_quat_g", task=("obj_red", "bin_red"))
    ea.reset(seed=0)
    from mujoco_manip_b200.gym_env import TARGET_ORI

    T = P.pos_rotmat_to_se3([0.1, 0.45, 0.45], TARGET_ORI)
    act = P.se3_to_pos_quat_g(T, 1.0)
    tgt, g = ea.decode_action(act)
    np.testing.assert_allclose(tgt, [0.1, 0.45, 0.45], atol=1e-6)
    for _ in range(20):
        ea.step(act)
    assert np.linalg.norm(ea.robot.ee_pos - tgt) < 0.05
    np.testing.assert_allclose(ea.robot.ee_xmat, TARGET_ORI, atol=0.1)
    np.testing.assert_allclose(ea.robot.ee_xmat @ ea.robot.ee_xmat.T, np.eye(3), atol=1e-6)
    for e in (e8, e10, ea):
        e.close()


def test_decode_action_matches_relative_frame(cuda_device):
    from mujoco_manip_b200 import pose_utils as P

    env = make(cuda_device, action_mode="ee_pos_rot6d_g_rel")
    env.reset(seed=1)
    T0 = env.initial_ee_se3
    Tw = P.pos_rotmat_to_se3([0.2, 0.5, 0.4], T0[:3, :3])
    a = P.se3_to_pos_rot6d_g(np.linalg.inv(T0) @ Tw, 0.0)
    tgt, g = env.decode_action(a)
    np.testing.assert_allclose(tgt, [0.2, 0.5, 0.4], atol=1e-5)
    assert g == 0.0
    env.close()


def test_target_keypoints_frozen_and_reset_hygiene(cuda_device):
    env = make(cuda_device, action_mode="abs_pos", task=("obj_red", "bin_blue"), reward_type="staged")
    obs0, _ = env.reset(seed=3)
    kp_o, kp_b = obs0["target_obj_keypoints_overhead"].copy(), obs0["target_bin_keypoints_overhead"].copy()
    assert np.all((kp_o >= 0) & (kp_o <= 1) & (kp_b >= 0) & (kp_b <= 1))
    last = -1.0
    for t in range(10):
        obs, r, te, tr, info = env.step(np.array([-0.15, 0.45, 0.40, 1.0], dtype=np.float32))
        np.testing.assert_array_equal(obs["target_obj_keypoints_overhead"], kp_o)
        np.testing.assert_array_equal(obs["target_bin_keypoints_overhead"], kp_b)
        assert 0.0 <= r <= 1.0 and r >= last - 1e-12  # staged reward: monotone, in [0, 1]
        last = r
        assert info["reward_components"].shape == (6,) and abs(info["reward_components"][0] - r) < 1e-5
    env.reset(seed=3)
    assert env.step_count == 0 and not env._has_grasped and env._reward_hwm is None
    env.close()


def test_staged_reward_collision_terminates(cuda_device):
    """Reference tests/test_gym_env.py:868-876: driving the hand to z = 0.10 hits the table -> -1, terminate."""
    env = make(cuda_device, action_mode="abs_pos", task=("obj_red", "bin_red"), reward_type="staged", max_episode_steps=100)
    env.reset(seed=0)
    hit = False
    for _ in range(60):
        obs, r, te, tr, info = env.step(np.array([0.3, 0.3, 0.10, 1.0], dtype=np.float32))
        if r < 0:
            hit = True
            assert r == -1.0 and te and info["success"] is False
            break
    assert hit
    env.close()


def test_controller_and_physics_level_loop(cuda_device):
    """tests/test_controller.py: compute() is finite and inside the joint limits; 200 x (compute, set ctrl,
    mj_step) brings the EE within 3 cm of a target with the hand pointing down."""
    from mujoco_manip_b200.gym_env import TARGET_ORI

    env = make(cuda_device, action_mode="abs_pos", task=("obj_red", "bin_red"))
    env.reset(seed=0)
    lo = np.array([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973])
    hi = np.array([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973])
    ctrl_before = env.pick_place_env._vec.state["ctrl"][0, :7].cpu().numpy().copy()
    q = env.controller.compute(np.array([0.2, 0.4, 0.45]))
    assert q.shape == (7,) and np.all(np.isfinite(q)) and np.all(q >= lo - 1e-6) and np.all(q <= hi + 1e-6)
    np.testing.assert_array_equal(env.pick_place_env._vec.state["ctrl"][0, :7].cpu().numpy(), ctrl_before)
    assert env.controller.reached(env.robot.ee_pos) and not env.controller.reached(np.array([0.5, 0.5, 0.9]))
    target = np.array([0.15, 0.45, 0.42])
    d0 = np.linalg.norm(env.robot.ee_pos - target)
    for _ in range(200):
        env.robot.set_arm_ctrl(env.controller.compute(target))
        env.pick_place_env.step()
    d1 = np.linalg.norm(env.robot.ee_pos - target)
    assert d1 < 0.03 and d1 < d0
    np.testing.assert_allclose(env.robot.ee_xmat, TARGET_ORI, atol=0.1)
    env.robot.close_gripper()
    assert env.robot.gripper_ctrl == 0.0
    env.robot.open_gripper()
    assert env.robot.gripper_ctrl == 255.0
    env.close()


def test_fsm_class_surface_and_expert_episode(cuda_device):
    """tests/test_pick_and_place.py: State -> Phase mapping, descriptions, plan() leaves qpos / ctrl alone, timers
    decrement by n_steps, and the plan(16) + env.step expert loop finishes with the cube in the bin."""
    from mujoco_manip_b200.pick_and_place import _STATE_TO_PHASE, Phase, PickAndPlaceTask, State

    assert [s.value for s in State] == list(range(1, 12))
    assert _STATE_TO_PHASE[State.CLOSE_GRIPPER] == Phase.GRASPING and _STATE_TO_PHASE[State.SETTLE_AT_BIN] == Phase.TRANSPORTING
    assert set(_STATE_TO_PHASE) == set(State)
    env = make(cuda_device, action_mode="abs_pos", task=("obj_red", "bin_red"), max_episode_steps=500)
    env.reset(seed=0)
    fsm = PickAndPlaceTask(env.pick_place_env, env.robot, env.controller, tasks=[("obj_red", "bin_red")])
    assert fsm.state == State.IDLE and fsm.phase_description == "idle" and fsm.target_pos is None and fsm.gripper_val == 1.0
    q0 = env.pick_place_env._vec.state["qpos"].clone()
    c0 = env.pick_place_env._vec.state["ctrl"].clone()
    fsm.plan(16)
    assert fsm.state == State.PRE_GRASP and fsm.phase_description == "approaching the red cube"
    import torch

    assert torch.equal(env.pick_place_env._vec.state["qpos"], q0) and torch.equal(env.pick_place_env._vec.state["ctrl"], c0)
    np.testing.assert_allclose(fsm.target_pos, [-0.15, 0.45, 0.44], atol=1e-9)
    seen, n = set(), 0
    while not fsm.is_done and n < 2000:
        fsm.plan(16)
        seen.add(fsm.phase)
        if fsm.state == State.CLOSE_GRIPPER and fsm.settle_counter == 150:
            assert fsm.gripper_val == 0.0
        tp = fsm.target_pos if fsm.target_pos is not None else env.robot.ee_pos
        obs, r, te, tr, info = env.step(np.array([*tp, fsm.gripper_val], dtype=np.float32))
        n += 1
    assert fsm.is_done and n < 2000 and len(seen) >= 6
    assert info["success"] is True
    p = env.pick_place_env.get_body_pos("obj_red")
    assert np.linalg.norm(p[:2] - [-0.3, 0.55]) < 0.05 and p[2] < 0.30
    assert fsm.phase_description == "idle"
    env.close()


@pytest.mark.parametrize("fname", ["fsm_multi3_seed5.npz", "fsm_multi2_cross_seed11.npz"])
def test_multi_task_fsm_matches_reference_fsm(cuda_device, fname):
    """PickAndPlaceTask with a task LIST (default TASKS has three pairs, pick_and_place.py:91): state, task index, timer
    and the status string of every plan(16) call equal the reference FSM's (golden produced by the reference's own
    class), and the FSM's list leaves the env's own task alone."""
    from mujoco_manip_b200.constants import BINS, OBJECTS
    from mujoco_manip_b200.pick_and_place import PickAndPlaceTask

    g = np.load(os.path.join(GOLDEN, fname))
    env = make(cuda_device, action_mode="abs_pos", task=("obj_red", "bin_red"), max_episode_steps=2000, randomize_objects=True)
    seed = int(fname.split("seed")[1].split(".")[0])
    env.reset(seed=seed)
    np.testing.assert_allclose(env.pick_place_env._vec.state["qpos"][0].cpu().numpy(), g["init_qpos"], atol=1e-12)
    tasks = [(OBJECTS[o], BINS[b]) for o, b in g["tasks"]]
    fsm = PickAndPlaceTask(env.pick_place_env, env.robot, env.controller, tasks=tasks)
    assert (env.obj_name, env.bin_name) == ("obj_red", "bin_red")
    for t in range(len(g["fsm_state"])):
        status = fsm.plan(16)
        assert status == str(g["status"][t]), (t, status, str(g["status"][t]))
        assert fsm.state.value == int(g["fsm_state"][t]) and fsm.task_index == int(g["task_index"][t]), t
        assert fsm.settle_counter == int(g["counter"][t])
        tp = fsm.target_pos if fsm.target_pos is not None else env.robot.ee_pos
        np.testing.assert_allclose(tp, g["target"][t], atol=1e-7)
        obs, r, te, tr, info = env.step(np.array([*tp, fsm.gripper_val], dtype=np.float32))
        assert reltol(env.pick_place_env._vec.state["qpos"][0].cpu().numpy(), g["qpos"][t], 1e-5) < 1e-5, t
        assert bool(info["success"]) == bool(g["success"][t])
    assert fsm.is_done and (env.obj_name, env.bin_name) == ("obj_red", "bin_red")
    env.close()


def test_expert_success_rate_matches_oracle(cuda_device, oracle_lib):
    """Scripted-FSM episodes with seeded placements (config 3 semantics: tasks cycle env % 9, placements from
    numpy PCG64 seeds spawned from SeedSequence(42)): per-episode success / length vs the CPU oracle;
    success rate within 1 percentage point (north_star)."""
    import torch
    from concurrent.futures import ThreadPoolExecutor

    from mujoco_manip_b200 import PickPlaceVecEnv
    from mujoco_manip_b200.constants import TASK_SETS, task_indices

    n = 192
    seeds = [int(c.generate_state(1)[0]) for c in np.random.SeedSequence(42).spawn(n)]
    env = PickPlaceVecEnv(n, device=cuda_device, tasks="all", action_mode="abs_pos", randomize_objects=True, rng="numpy",
                          auto_reset=False, task_assignment="cycle", max_episode_steps=2000)
    env.reset(seed=seeds)
    xy = env._obj_xy.cpu().numpy().reshape(n, 3, 2)
    tasks = env._task.cpu().numpy()
    assert [tuple(t) for t in tasks[:9]] == [task_indices(t) for t in TASK_SETS["all"]]
    done_len = torch.zeros(n, dtype=torch.int32, device=cuda_device)
    succ = torch.zeros(n, dtype=torch.bool, device=cuda_device)
    for t in range(400):
        running = env.fsm_state != 11  # the loop of generate_dataset.py:140 steps once more after the plan that reaches DONE
        if not bool(running.any()):
            break
        a = env.fsm_plan(16)
        obs, r, te, tr, info = env.step(a)
        done_len += running.to(torch.int32)
        succ = torch.where(running, info["success"], succ)

    def run(i):
        o = oracle_lib.OracleEnv(action_mode="abs_pos")
        s, length, hist = o.run_fsm_episode(xy[i], int(tasks[i, 0]), int(tasks[i, 1]), 2000)
        return s, length

    with ThreadPoolExecutor(16) as ex:
        ref = list(ex.map(run, range(n)))
    ref_s = np.array([r[0] for r in ref])
    ref_l = np.array([r[1] for r in ref])
    gs, gl = succ.cpu().numpy(), done_len.cpu().numpy()
    assert abs(gs.mean() - ref_s.mean()) <= 0.01, (gs.mean(), ref_s.mean())
    assert (gs == ref_s).mean() >= 0.98
    assert (gl == ref_l).mean() >= 0.95, (gl[:10], ref_l[:10])
    assert ref_s.mean() > 0.8


def test_expert_action_encodings_match_reference_get_actions(cuda_device):
    """mm_expert_actions vs the arithmetic of scripts/generate_dataset.py:56-80 (host pose utils, pinned on the
    reference's own vectors in tests/test_pose_utils.py)."""
    import torch

    from mujoco_manip_b200 import PickPlaceVecEnv
    from mujoco_manip_b200 import pose_utils as P
    from mujoco_manip_b200.features import expert_action_encodings, pack_rows
    from mujoco_manip_b200.gym_env import TARGET_ORI

    env = PickPlaceVecEnv(5, device=cuda_device, action_mode="abs_pos", reward_type="staged", rng="numpy", auto_reset=False)
    env.reset()
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.uniform([-0.3, 0.3, 0.3], [0.3, 0.65, 0.6], size=(5, 3)), rng.integers(0, 2, size=(5, 1))], axis=1)
    enc = expert_action_encodings(env, torch.from_numpy(a.astype(np.float32)).to(cuda_device)).cpu().numpy()
    T0 = env.initial_ee_se3.cpu().numpy()
    for i in range(5):
        g = float(np.float32(a[i, 3]))
        T = P.pos_rotmat_to_se3(a[i, :3].astype(np.float32).astype(np.float64), TARGET_ORI)
        Tr = np.linalg.inv(T0[i]) @ T
        ref = np.concatenate([P.se3_to_pos_quat_g(T, g), P.se3_to_pos_rot6d_g(T, g), P.se3_to_pos_quat_g(Tr, g),
                              P.se3_to_pos_rot6d_g(Tr, g)])
        np.testing.assert_allclose(enc[i], ref, atol=1e-6)
    pre = env.obs_packed.clone()
    obs, r, te, tr, info = env.step(torch.from_numpy(a.astype(np.float32)).to(cuda_device))
    row = pack_rows(pre, torch.from_numpy(enc).to(cuda_device), env.fsm_state, info["reward_components"])
    from mujoco_manip_b200.features import FEATURES

    for k, v in row.items():
        if FEATURES[k]["dtype"] == "float32":
            assert tuple(v.shape[1:]) == FEATURES[k]["shape"], k


def test_randomize_yaw_draw_order_and_quaternions(cuda_device):
    """randomize_yaw=True (randomization.py:19,55-62): positions are drawn first (rejection sampler), then one
    theta = uniform(0, 2 pi) per cube in order; the cube quaternion becomes (cos(theta/2), 0, 0, sin(theta/2))."""
    from mujoco_manip_b200 import PickPlaceVecEnv
    from mujoco_manip_b200.randomization import sample_separated_positions

    # engine-level call of the N = 1 facade, as PickPlaceEnv.randomize_objects(rng, randomize_yaw=True)
    env = make(cuda_device)
    env.reset(seed=0)
    out = env.pick_place_env.randomize_objects(np.random.default_rng(123), randomize_yaw=True)
    ref = np.random.default_rng(123)
    pos = sample_separated_positions(ref, 3, (-0.20, 0.20), (0.30, 0.45))
    theta = [ref.uniform(0, 2 * np.pi) for _ in range(3)]
    q = env.pick_place_env._vec.state["qpos"][0].cpu().numpy()
    for o, name in enumerate(("obj_red", "obj_green", "obj_blue")):
        np.testing.assert_array_equal(out[f"{name}_jnt"], [pos[o][0], pos[o][1], 0.26])
        np.testing.assert_allclose(q[9 + 7 * o: 12 + 7 * o], [pos[o][0], pos[o][1], 0.26], rtol=0, atol=0)
        np.testing.assert_allclose(q[12 + 7 * o: 16 + 7 * o], [np.cos(theta[o] / 2), 0, 0, np.sin(theta[o] / 2)], rtol=0, atol=1e-15)
        R = env.pick_place_env.get_body_xmat(name)
        np.testing.assert_allclose(R[:2, :2], [[np.cos(theta[o]), -np.sin(theta[o])], [np.sin(theta[o]), np.cos(theta[o])]], atol=1e-12)
    # vectorised env with per-env numpy generators: same draw order, then the task draw
    v = PickPlaceVecEnv(3, device=cuda_device, tasks="all", randomize_objects=True, randomize_yaw=True, rng="numpy", auto_reset=False)
    v.reset(seed=[7, 8, 9])
    qv = v.state["qpos"].cpu().numpy()
    for i, sd in enumerate((7, 8, 9)):
        ref = np.random.default_rng(sd)
        pos = sample_separated_positions(ref, 3, (-0.20, 0.20), (0.30, 0.45))
        theta = np.array([ref.uniform(0, 2 * np.pi) for _ in range(3)])
        task = int(ref.integers(9))
        np.testing.assert_array_equal(v.last_yaw[i].cpu().numpy(), theta)
        np.testing.assert_array_equal(qv[i, 9:30].reshape(3, 7)[:, :2], np.asarray(pos))
        np.testing.assert_allclose(qv[i, 9:30].reshape(3, 7)[:, 3], np.cos(theta / 2), rtol=0, atol=1e-15)
        np.testing.assert_allclose(qv[i, 9:30].reshape(3, 7)[:, 6], np.sin(theta / 2), rtol=0, atol=1e-15)
        assert tuple(v._task[i].cpu().numpy()) == tuple(v._pool_idx[task].cpu().numpy())
