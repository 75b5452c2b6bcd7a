"""The C++ oracle (oracle/) against the golden vectors recorded by running the UNMODIFIED reference
Python (IK, FSM, decode, reward, obs packing, reset order) on the oracle engine
(tools/make_golden.py).  This pins the oracle's restatement of everything above the engine boundary."""
import os

import numpy as np
import pytest

from hostlib import GOLDEN, MODES

RANDOM_FILES = [f"random50_{m}.npz" for m in MODES]


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.mark.parametrize("fname", RANDOM_FILES)
def test_random_rollout_matches_reference_python(oracle_lib, fname):
    g = _load(fname)
    mode = fname[len("random50_"):-4]
    env = oracle_lib.OracleEnv(action_mode=mode)
    obs0 = env.reset(None, 0, 0)
    np.testing.assert_allclose(obs0, g["obs0"], rtol=0, atol=1e-7)
    for t in range(g["action"].shape[0]):
        obs, r, te, tr, info = env.step(g["action"][t])
        # identical arithmetic (same engine, restated Python layer): agreement to rounding of the
        # numpy-vs-C++ linear algebra in the IK (np.linalg.inv vs Cholesky), amplified by the servo gains
        np.testing.assert_allclose(env.qpos, g["qpos"][t], rtol=0, atol=2e-9)
        np.testing.assert_allclose(env.qvel, g["qvel"][t], rtol=0, atol=2e-6)
        np.testing.assert_allclose(obs, g["obs"][t], rtol=0, atol=2e-6)
        assert abs(r - g["reward"][t]) < 1e-7
        assert te == bool(g["terminated"][t]) and tr == bool(g["truncated"][t])
        assert info["success"] == bool(g["success"][t])


@pytest.mark.parametrize("fname,mode,reward", [
    ("fsm_quat_rel_red_red.npz", "ee_pos_quat_g_rel", "dense"),
    ("fsm_abs_green_blue_seed42_staged.npz", "abs_pos", "staged"),
    ("fsm_rot6d_rel_blue_red_seed7.npz", "ee_pos_rot6d_g_rel", "dense"),
])
def test_fsm_episode_matches_reference_python(oracle_lib, fname, mode, reward):
    """Scripted expert episode: FSM state sequence bit-exact, trajectories to rounding."""
    g = _load(fname)
    env = oracle_lib.OracleEnv(action_mode="abs_pos", reward_type=reward)
    xy = None
    if not np.allclose(g["init_qpos"][9:11], [-0.15, 0.45]):
        q = g["init_qpos"]
        xy = np.array([q[9:11], q[16:18], q[23:25]])
    env.reset(xy, int(g["obj_idx"]), int(g["bin_idx"]))
    env.fsm_reset()
    n = g["fsm_state"].shape[0]
    for t in range(n):
        env.fsm_plan(16)
        f = env.fsm_get()
        assert f["state"] == int(g["fsm_state"][t]), f"FSM state differs at step {t}"
        assert f["counter"] == int(g["counter"][t])
        a = env.fsm_action()
        np.testing.assert_allclose(a[:3], g["target"][t].astype(np.float32), rtol=0, atol=1e-6)
        assert a[3] == np.float32(g["gripper"][t])
        obs, r, te, tr, info = env.step(a)
        np.testing.assert_allclose(env.qpos, g["qpos"][t], rtol=0, atol=5e-6)
        assert abs(r - g["reward"][t]) < 1e-4
        if reward == "staged":
            np.testing.assert_allclose(info["reward_components"], g["rc"][t], atol=1e-5)
    env.fsm_plan(16)
    assert env.fsm_get()["state"] == int(g["final_fsm_state"]) == 11


def test_stress_rollout_with_hull_contacts(oracle_lib):
    g = _load("stress30_abs_pos_staged.npz")
    env = oracle_lib.OracleEnv(action_mode="abs_pos", reward_type="staged")
    env.reset(None, 0, 0)
    for t in range(g["action"].shape[0]):
        obs, r, te, tr, info = env.step(g["action"][t])
        np.testing.assert_allclose(env.qpos, g["qpos"][t], rtol=0, atol=1e-7)
        assert abs(r - g["reward"][t]) < 1e-6
        assert te == bool(g["terminated"][t])
    assert g["reward"].min() == -1.0  # the rollout does hit the table (gym_env.py:429-430)


def test_reset_vectors(oracle_lib):
    g = _load("reset_seeds.npz")
    for i, seed in enumerate(g["seeds"]):
        xy, rng = oracle_lib.sample_placement(int(seed))
        np.testing.assert_array_equal(xy, g["obj_xy"][i])  # bit-exact RNG
        tidx = int(rng.integers(9))
        assert tuple(g["task_sets_all"][tidx]) == tuple(g["task"][i])
        env = oracle_lib.OracleEnv(action_mode="ee_pos_rot6d_g_rel")
        obs0 = env.reset(xy, int(g["task"][i][0]), int(g["task"][i][1]))
        np.testing.assert_array_equal(env.qpos, g["qpos"][i])
        np.testing.assert_allclose(obs0, g["obs0"][i], rtol=0, atol=1e-7)
    np.testing.assert_array_equal(oracle_lib.sample_placement(42)[0], g["sampler42"])
    ss = np.random.SeedSequence(42).spawn(8)
    np.testing.assert_array_equal(np.array([int(c.generate_state(1)[0]) for c in ss], dtype=np.uint64), g["episode_seeds42"])


def test_fk_known_answer(oracle_lib):
    """SURVEY 2.1 [DERIVED] known answer at the keyframe."""
    env = oracle_lib.OracleEnv()
    env.reset(None, 0, 0)
    np.testing.assert_allclose(env.xpos[9], [-1.78e-6, 0.485004798, 0.497767036], atol=2e-8)
