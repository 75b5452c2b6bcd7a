"""Parity tests proper: the CUDA path (through the C ABI, via PickPlaceVecEnv) against the CPU oracle
and the committed golden vectors.  Run on the B200 box with `-m gpu`."""
import os

import numpy as np
import pytest

from hostlib import GOLDEN, MODES, reltol

pytestmark = pytest.mark.gpu

# north_star: per-step qpos, qvel and EE pose within 1e-5 relative over the first 50 steps,
# |a-b| <= TOL * max(|b|, 1)
TOL = 1e-5


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


def _make(n, dev, **kw):
    from mujoco_manip_b200 import PickPlaceVecEnv

    kw.setdefault("auto_reset", False)
    kw.setdefault("rng", "numpy")
    return PickPlaceVecEnv(n, device=dev, **kw)


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("group", [32, 16, 8])
@pytest.mark.parametrize("mode", MODES)
def test_f64_random_rollout_vs_oracle(cuda_device, oracle_lib, mode, group):
    """50 seeded random-action steps, 4 envs: 2 copies of the golden action stream + 2 fresh streams."""
    import torch

    g = _load(f"random50_{mode}.npz")
    rng = np.random.default_rng(5)
    n = 4
    env = _make(n, cuda_device, task=("obj_red", "bin_red"), action_mode=mode, group=group)
    env.reset()
    orcs = [oracle_lib.OracleEnv(action_mode=mode, flags=0) for _ in range(n)]
    for o in orcs:
        o.reset(None, 0, 0)
    for t in range(50):
        acts = np.repeat(g["action"][t][None], n, axis=0)
        for k in (2, 3):
            acts[k, :3] += rng.uniform(-0.05, 0.05, size=3).astype(np.float32)
        obs, r, te, tr, info = env.step(torch.from_numpy(acts[:, : env.action_dim]).to(cuda_device))
        qpos, qvel, ee = _np(env.state["qpos"]), _np(env.state["qvel"]), _np(env.state["eepose"])
        for k in range(n):
            o_obs, o_r, o_te, o_tr, o_info = orcs[k].step(acts[k])
            assert reltol(qpos[k], orcs[k].qpos, TOL) < TOL, f"qpos env {k} step {t}"
            assert reltol(qvel[k], orcs[k].qvel, TOL) < TOL, f"qvel env {k} step {t}"
            assert reltol(ee[k][:3], orcs[k].xpos[9], TOL) < TOL
            assert reltol(ee[k][3:], orcs[k].xmat[9], TOL) < TOL
            np.testing.assert_allclose(_np(env.obs_packed)[k], o_obs, rtol=0, atol=2e-5)
            assert abs(float(r[k]) - o_r) < 1e-4
            assert bool(te[k]) == o_te and bool(tr[k]) == o_tr and bool(info["success"][k]) == o_info["success"]
    assert np.array_equal(qpos[0], qpos[1])  # identical inputs -> bit-identical envs


@pytest.mark.parametrize("fname,reward", [
    ("fsm_quat_rel_red_red.npz", "dense"),
    ("fsm_abs_green_blue_seed42_staged.npz", "staged"),
    ("fsm_rot6d_rel_blue_red_seed7.npz", "dense"),
])
def test_f64_fsm_episode_vs_oracle(cuda_device, oracle_lib, fname, reward):
    """Scripted expert episode (config 1): FSM state indices bit-exact vs the reference's own FSM
    (golden), trajectory within tolerance vs the oracle."""
    g = _load(fname)
    q = g["init_qpos"]
    xy = np.array([q[9:11], q[16:18], q[23:25]])
    oi, bi = int(g["obj_idx"]), int(g["bin_idx"])
    env = _make(2, cuda_device, action_mode="abs_pos", reward_type=reward)
    from mujoco_manip_b200.constants import BINS, OBJECTS

    env.reset(options={"task": (OBJECTS[oi], BINS[bi]), "obj_xy": np.stack([xy, xy])})
    orc = oracle_lib.OracleEnv(action_mode="abs_pos", reward_type=reward, flags=0)
    orc.reset(xy, oi, bi)
    orc.fsm_reset()
    n = g["fsm_state"].shape[0]
    done_at = None
    fsm_trace = []
    for t in range(n + 40):
        a = env.fsm_plan(16)
        orc.fsm_plan(16)
        f = orc.fsm_get()
        fs = _np(env.state["fsm_i"])
        assert int(fs[0, 0]) == f["state"] and int(fs[0, 2]) == f["counter"], f"FSM differs from oracle at {t}"
        if f["state"] == 11:
            done_at = t
            break
        fsm_trace.append(int(fs[0, 0]))
        np.testing.assert_allclose(_np(a)[0], orc.fsm_action(), rtol=0, atol=1e-6)
        obs, r, te, tr, info = env.step(a)
        o_obs, o_r, o_te, o_tr, o_info = orc.step(orc.fsm_action())
        assert reltol(_np(env.state["qpos"])[0], orc.qpos, TOL) < TOL, f"step {t}"
        if t < n:
            assert reltol(_np(env.state["qpos"])[0], g["qpos"][t], TOL) < TOL, f"golden step {t}"
        assert abs(float(r[0]) - o_r) < 1e-3
    assert done_at is not None
    # same phase sequence and episode length as the reference's own FSM run (golden; its last plan reaches DONE)
    assert [int(x) for x in g["fsm_state"]] == fsm_trace + [11]


def test_f32_tracks_oracle(cuda_device, oracle_lib):
    """FP32 arithmetic (throughput path) on the arm/gripper dofs; bound stated, not the 1e-5 bar."""
    import torch

    g = _load("random50_ee_pos_quat_g_rel.npz")
    env = _make(2, cuda_device, task=("obj_red", "bin_red"), precision="f32")
    env.reset()
    orc = oracle_lib.OracleEnv(flags=0)
    orc.reset(None, 0, 0)
    worst = 0.0
    for t in range(50):
        a = np.repeat(g["action"][t][None], 2, axis=0)
        env.step(torch.from_numpy(a[:, :8]).to(cuda_device))
        orc.step(g["action"][t])
        worst = max(worst, reltol(_np(env.state["qpos"])[0][:9], orc.qpos[:9], 1.0))
    assert worst < 5e-3, worst


def test_philox_bit_exact(cuda_device):
    from oracle import philox

    env = _make(64, cuda_device, randomize_objects=True, rng="philox", seed=42, env_id_offset=1000, tasks="all")
    env.reset()
    xy = _np(env._obj_xy).reshape(64, 3, 2)
    tk = _np(env._task)
    pool = _np(env._pool_idx)
    for i in range(64):
        exp, att = philox.place(42, 1000 + i, 0)
        assert np.array_equal(xy[i], exp), i
        assert int(_np(env.last_attempts)[i]) == att
        assert tuple(tk[i]) == tuple(pool[philox.task_draw(42, 1000 + i, 0, 9)])
    env.reset()  # second episode uses episode index 1
    xy2 = _np(env._obj_xy).reshape(64, 3, 2)
    assert np.array_equal(xy2[5], philox.place(42, 1005, 1)[0])


def test_sharding_invariance(cuda_device):
    """Two 'ranks' of 32 envs (env_id_offset 0 / 32) == one launch of 64 envs, bit for bit."""
    import torch

    kw = dict(randomize_objects=True, rng="philox", seed=7, tasks="cross", action_mode="abs_pos", precision="f32")
    full = _make(64, cuda_device, **kw)
    a_, b_ = _make(32, cuda_device, env_id_offset=0, **kw), _make(32, cuda_device, env_id_offset=32, **kw)
    for e in (full, a_, b_):
        e.reset()
    for t in range(3):
        acts = [e.fsm_plan(16).clone() for e in (full, a_, b_)]
        for e, a in zip((full, a_, b_), acts):
            e.step(a)
    q = _np(full.state["qpos"])
    assert np.array_equal(q[:32], _np(a_.state["qpos"])) and np.array_equal(q[32:], _np(b_.state["qpos"]))


def test_large_batch_properties(cuda_device):
    """Full-size properties (config 2 size): identical copies stay identical, unit quaternions,
    cubes stay on the table, finite state, step counts advance."""
    import torch

    n = 4096
    env = _make(n, cuda_device, task=("obj_red", "bin_red"), precision="f32")
    env.reset()
    gen = torch.Generator(device=cuda_device).manual_seed(1234)
    for t in range(6):
        a = torch.zeros((n, 8), device=cuda_device)
        a[:, :3] = (torch.rand((n // 2, 3), device=cuda_device, generator=gen) * 0.2 - 0.1).repeat(2, 1)
        a[:, 6] = 1.0
        a[:, 7] = (torch.rand(n // 2, device=cuda_device, generator=gen) > 0.5).float().repeat(2)
        env.step(a)
    q = env.state["qpos"]
    assert torch.isfinite(q).all() and torch.isfinite(env.state["qvel"]).all()
    assert torch.equal(q[: n // 2], q[n // 2:])
    for o in range(3):
        quat = q[:, 9 + 7 * o + 3: 9 + 7 * o + 7]
        assert torch.allclose(quat.norm(dim=1), torch.ones(n, dtype=quat.dtype, device=cuda_device), atol=1e-6)
        assert (q[:, 9 + 7 * o + 2] > 0.2).all()
    assert (env.state["step_count"] == 6).all()
    assert int(env.state["diag"][:, 2].max()) == 0  # no workspace overflow


def test_host_buffer_step_matches_device_step(cuda_device):
    """mm_step_host (the e2e path) == mm_step."""
    import ctypes as C

    import torch

    from mujoco_manip_b200 import _lib

    n = 8
    e1, e2 = _make(n, cuda_device, action_mode="abs_pos"), _make(n, cuda_device, action_mode="abs_pos")
    e1.reset()
    e2.reset()
    a = torch.tensor([[0.1, 0.45, 0.4, 1.0]] * n, dtype=torch.float32)
    e1.step(a.to(cuda_device))
    ha = torch.zeros((n, 10), dtype=torch.float32).pin_memory()
    ha[:, :4] = a
    hobs = torch.zeros((n, 85), dtype=torch.float32).pin_memory()
    hr = torch.zeros(n, dtype=torch.float32).pin_memory()
    _lib.check(e2._L.mm_step_host(e2._h, C.byref(e2._st), ha.data_ptr(), 0, hobs.data_ptr(), hr.data_ptr(), None, None, None,
                                  e2._stream()), "mm_step_host")
    assert torch.equal(hobs, e1.obs_packed.cpu())
    assert torch.equal(e1.state["qpos"], e2.state["qpos"])


def test_vec_env_step_host_matches_step_across_auto_resets(cuda_device):
    """PickPlaceVecEnv.step_host (mm_step_host_async + the episode bookkeeping enqueued behind it + ONE wait) returns, in the
    host buffers, what step() returns on the device - including the steps on which episodes end and envs restart."""
    import torch

    n = 16
    kw = dict(action_mode="abs_pos", max_episode_steps=4, auto_reset=True, rng="philox", randomize_objects=True, seed=5)
    e1, e2 = _make(n, cuda_device, **kw), _make(n, cuda_device, **kw)
    e1.reset()
    e2.reset()
    gen = torch.Generator().manual_seed(3)
    ha = torch.zeros((n, 10), dtype=torch.float32).pin_memory()
    hobs = torch.zeros((n, 85), dtype=torch.float32).pin_memory()
    hr = torch.zeros(n, dtype=torch.float32).pin_memory()
    hf = torch.zeros((3, n), dtype=torch.uint8).pin_memory()
    for t in range(11):
        a = torch.rand((n, 4), generator=gen) * torch.tensor([0.4, 0.4, 0.3, 1.0]) + torch.tensor([-0.1, -0.2, 0.25, 0.0])
        a[:, 3] = (a[:, 3] > 0.5).float()
        _, r1, term1, trunc1, info1 = e1.step(a.to(cuda_device))
        ha[:, :4] = a
        e2.step_host(ha, hobs, hr, hf)
        # the host observation is the step's own (pre-reset) observation: final_obs where an episode ended
        done = (term1 | trunc1).cpu()
        want = torch.where(done[:, None], info1["final_obs"].cpu(), e1.obs_packed.cpu())
        assert torch.equal(hobs, want), t
        assert torch.equal(hr, r1.cpu())
        assert torch.equal(hf[0].bool(), term1.cpu()) and torch.equal(hf[1].bool(), trunc1.cpu())
        assert torch.equal(e1.state["qpos"], e2.state["qpos"]), t
    assert torch.equal(e1.stats, e2.stats)


def test_autoreset_and_stats(cuda_device):
    import torch

    env = _make(16, cuda_device, action_mode="abs_pos", max_episode_steps=3, auto_reset=True, rng="philox",
                randomize_objects=True)
    env.reset()
    a = torch.tensor([[0.0, 0.45, 0.5, 1.0]] * 16, device=cuda_device)
    for t in range(3):
        obs, r, te, tr, info = env.step(a)
    assert bool(tr.all())
    assert (env.state["step_count"] == 0).all()  # already reset
    assert float(env.stats[0]) == 16.0 and float(env.stats[2]) == 48.0
    assert (env.episode_index == 2).all()


def test_f64_stress_rollout_with_table_collisions(cuda_device):
    """Hand driven into the table (link / hand / finger hull contacts): staged reward -1 and terminate,
    trajectory vs the golden file recorded from the reference's Python on the full oracle."""
    import torch

    g = _load("stress30_abs_pos_staged.npz")
    env = _make(2, cuda_device, task=("obj_red", "bin_red"), action_mode="abs_pos", reward_type="staged")
    env.reset()
    for t in range(g["action"].shape[0]):
        a = torch.from_numpy(np.repeat(g["action"][t][None, :4], 2, axis=0)).to(cuda_device)
        obs, r, te, tr, info = env.step(a)
        assert reltol(_np(env.state["qpos"])[0], g["qpos"][t], TOL) < TOL, t
        assert abs(float(r[0]) - g["reward"][t]) < 1e-5
        assert bool(te[0]) == bool(g["terminated"][t])
    assert int(env.state["diag"][:, 2].max()) == 0
    assert g["reward"].min() == -1.0


@pytest.mark.parametrize("mode", MODES)
def test_f64_random_rollout_vs_golden(cuda_device, mode):
    import torch

    g = _load(f"random50_{mode}.npz")
    env = _make(1, cuda_device, task=("obj_red", "bin_red"), action_mode=mode)
    env.reset()
    np.testing.assert_allclose(_np(env.obs_packed)[0], g["obs0"], rtol=0, atol=1e-6)
    for t in range(50):
        a = torch.from_numpy(g["action"][t][None, : env.action_dim]).to(cuda_device)
        obs, r, te, tr, info = env.step(a)
        assert reltol(_np(env.state["qpos"])[0], g["qpos"][t], TOL) < TOL, t
        assert reltol(_np(env.state["qvel"])[0], g["qvel"][t], TOL) < TOL, t
        assert reltol(_np(env.state["eepose"])[0][:3], g["ee_pos"][t], TOL) < TOL
        assert reltol(_np(env.state["eepose"])[0][3:], g["ee_R"][t], TOL) < TOL
        assert int(env.state["diag"][0, 0]) == int(g["ncon"][t])
        np.testing.assert_allclose(_np(env.obs_packed)[0], g["obs"][t], rtol=0, atol=2e-5)


def test_physics_level_loop_vs_oracle(cuda_device, oracle_lib):
    """Engine-level ops (mm_ops): compute -> set_arm_ctrl -> mj_step loop of main.py / tests/test_controller.py.
    The IK of tick k runs on the kinematics of the previous position stage; the oracle's
    Data keeps that staleness naturally, mm_ops through the `kin` state."""
    import torch

    from mujoco_manip_b200 import PickPlaceGymEnv

    env = PickPlaceGymEnv(action_mode="abs_pos", task=("obj_red", "bin_red"), device=str(cuda_device))
    env.reset(seed=0)
    orc = oracle_lib.OracleEnv(action_mode="abs_pos")
    orc.reset(None, 0, 0)
    target = np.array([0.12, 0.5, 0.4])
    for t in range(60):
        q = env.controller.compute(target)
        qo = orc.ik(target)
        np.testing.assert_allclose(q, qo, rtol=0, atol=1e-9, err_msg=f"IK differs at tick {t}")
        env.robot.set_arm_ctrl(q)
        orc.ctrl[:7] = qo
        if t == 30:
            env.robot.close_gripper()
            orc.ctrl[7] = 0.0
        env.pick_place_env.step()
        orc.mj_step()
        qpos = env.pick_place_env._vec.state["qpos"][0].cpu().numpy()
        assert reltol(qpos, orc.qpos, TOL) < 1e-8, t
        # robot.ee_pos is the pose of the last position stage (pre-integration), as data.xpos in the reference
        np.testing.assert_allclose(env.robot.ee_pos, orc.xpos[9], rtol=0, atol=1e-9)
    env.pick_place_env.forward()
    orc.mj_forward()
    np.testing.assert_allclose(env.robot.ee_pos, orc.xpos[9], rtol=0, atol=1e-9)
    env.close()


def test_state_round_trip_is_reproducible(cuda_device):
    """get_state / set_state (checkpoint of the sim): replaying from a snapshot gives bit-identical results,
    whatever the env-to-CTA schedule was."""
    import torch

    env = _make(37, cuda_device, action_mode="abs_pos", randomize_objects=True, rng="philox", seed=3, tasks="cross")
    env.reset()
    for _ in range(3):
        env.step(env.fsm_plan(16).clone())
    snap = env.get_state()
    acts = []
    for _ in range(4):
        a = env.fsm_plan(16).clone()
        acts.append(a)
        env.step(a)
    end1 = {k: v.clone() for k, v in env.state.items()}
    env.set_state(snap)
    env._work.zero_()  # different schedule on the replay
    for a in acts:
        env.fsm_plan(16)
        env.step(a)
    for k in ("qpos", "qvel", "ctrl", "warm", "fsm_i", "fsm_f", "step_count"):
        assert torch.equal(env.state[k], end1[k]), k


def test_reward_types_and_success_flags_vs_oracle(cuda_device, oracle_lib):
    """dense / sparse / staged rewards, terminated / success flags for three different tasks in one batch."""
    import torch

    tasks = [("obj_red", "bin_blue"), ("obj_green", "bin_red"), ("obj_blue", "bin_green")]
    for rtype in ("dense", "sparse", "staged"):
        env = _make(3, cuda_device, action_mode="abs_pos", reward_type=rtype)
        env.reset(options={"task": tasks})
        orcs = []
        for o, b in tasks:
            oc = oracle_lib.OracleEnv(action_mode="abs_pos", reward_type=rtype)
            oc.reset(None, ["obj_red", "obj_green", "obj_blue"].index(o), ["bin_red", "bin_green", "bin_blue"].index(b))
            oc.fsm_reset()
            orcs.append(oc)
        for t in range(25):
            a = env.fsm_plan(16).clone()
            obs, r, te, tr, info = env.step(a)
            for k, oc in enumerate(orcs):
                oc.fsm_plan(16)
                o_obs, o_r, o_te, o_tr, o_info = oc.step(oc.fsm_action())
                assert abs(float(r[k]) - o_r) < 1e-5, (rtype, t, k)
                assert bool(te[k]) == o_te and bool(info["success"][k]) == o_info["success"]
                if rtype == "staged":
                    np.testing.assert_allclose(info["reward_components"][k].cpu().numpy(), o_info["reward_components"], atol=1e-5)
                np.testing.assert_allclose(_np(env.obs_packed)[k], o_obs, rtol=0, atol=2e-5)
        env.close()


def test_philox_yaw_draw_and_placement(cuda_device):
    """randomize_yaw (randomization.py:55-62) on the device stream: theta bit-exact with oracle/philox.py,
    the cube quaternions (cos, 0, 0, sin)(theta/2) to the last bits (device cos/sin vs libm)."""
    from oracle import philox

    env = _make(32, cuda_device, randomize_objects=True, randomize_yaw=True, rng="philox", seed=9, env_id_offset=77)
    env.reset()
    th = _np(env.last_yaw)
    q = _np(env.state["qpos"])[:, 9:30].reshape(32, 3, 7)
    for i in range(32):
        exp = np.array([philox.yaw(9, 77 + i, 0, o) for o in range(3)])
        assert np.array_equal(th[i], exp), i
        assert np.all((exp >= 0) & (exp < 2 * np.pi))
        np.testing.assert_allclose(q[i, :, 3], np.cos(exp / 2), rtol=0, atol=4e-16)
        np.testing.assert_allclose(q[i, :, 6], np.sin(exp / 2), rtol=0, atol=4e-16)
        assert np.all(q[i, :, 4:6] == 0)
        assert np.array_equal(q[i, :, :2], philox.place(9, 77 + i, 0)[0])
    # option off again: identity quaternions
    env2 = _make(4, cuda_device, randomize_objects=True, rng="philox", seed=9, env_id_offset=77)
    env2.reset()
    assert np.all(_np(env2.state["qpos"])[:, 9:30].reshape(4, 3, 7)[:, :, 3] == 1.0)


def test_yawed_cubes_fsm_episode_vs_oracle(cuda_device, oracle_lib):
    """Scripted expert on cubes that are not axis-aligned, several envs with different yaw / placement:
    FSM states bit-exact, qpos within TOL at every step, success flags equal (a cube turned ~45 degrees can slip
    out of the expert's grasp - in the oracle and on the GPU alike)."""
    n = 6
    rng = np.random.default_rng(11)
    xy = np.stack([oracle_lib.sample_placement(100 + i)[0] for i in range(n)])
    yaw = rng.uniform(0, 2 * np.pi, size=(n, 3))
    tasks = [(i % 3, (i + 1) % 3) for i in range(n)]
    from mujoco_manip_b200.constants import BINS, OBJECTS

    env = _make(n, cuda_device, action_mode="abs_pos")
    env.reset(options={"obj_xy": xy, "obj_yaw": yaw, "task": [(OBJECTS[o], BINS[b]) for o, b in tasks]})
    orcs = []
    for i in range(n):
        o = oracle_lib.OracleEnv(action_mode="abs_pos", flags=0)
        o.reset(xy[i], tasks[i][0], tasks[i][1], yaw=yaw[i])
        o.fsm_reset()
        orcs.append(o)
    assert reltol(_np(env.state["qpos"]), np.stack([o.qpos for o in orcs]), TOL) < 1e-9
    done = np.zeros(n, dtype=bool)
    succ = np.zeros(n, dtype=bool)
    osucc = np.zeros(n, dtype=bool)
    for t in range(400):
        a = env.fsm_plan(16).clone()
        fs = _np(env.fsm_state)
        for i, o in enumerate(orcs):
            if not done[i]:
                o.fsm_plan(16)
                assert int(fs[i]) == o.fsm_get()["state"], (t, i)
        done |= fs == 11
        if done.all():
            break
        _, _, _, _, info = env.step(a)
        q = _np(env.state["qpos"])
        s = _np(info["success"])
        for i, o in enumerate(orcs):
            if not done[i]:
                osucc[i] = o.step(o.fsm_action())[4]["success"]
                assert reltol(q[i], o.qpos, TOL) < TOL, (t, i)
                succ[i] = s[i]
    assert done.all() and np.array_equal(succ, osucc) and succ.sum() >= n - 2


def test_large_batch_kernel_variant_vs_oracle(cuda_device, oracle_lib):
    """Batches of 8,192 envs and more run the 6-warp CTA shape of the step kernel (csrc/mm_launch.cuh); a ragged batch
    of 9,001 scripted-expert envs (Philox placements, cycled tasks, load-aware scheduling on) is compared with the
    oracle on a handful of its envs, first and last slots included."""
    n = 9001
    env = _make(n, cuda_device, action_mode="abs_pos", tasks="all", randomize_objects=True, rng="philox", seed=5,
                task_assignment="cycle")
    env.reset()
    xy = _np(env._obj_xy).reshape(n, 3, 2)
    tk = _np(env._task)
    idx = [0, 1, 4500, 8191, 8999, 9000]
    orcs = []
    for i in idx:
        o = oracle_lib.OracleEnv(action_mode="abs_pos", flags=0)
        o.reset(xy[i], int(tk[i, 0]), int(tk[i, 1]))
        o.fsm_reset()
        orcs.append(o)
    for t in range(45):
        a = env.fsm_plan(16).clone()
        fs = _np(env.fsm_state)
        env.step(a)
        q = _np(env.state["qpos"][idx])
        for k, o in enumerate(orcs):
            o.fsm_plan(16)
            assert int(fs[idx[k]]) == o.fsm_get()["state"], (t, k)
            o.step(o.fsm_action())
            assert reltol(q[k], o.qpos, TOL) < TOL, (t, k)
    assert int(env.state["diag"][:, 2].max()) == 0
    assert bool((env.state["step_count"] == 45).all())


def test_env_on_second_gpu_without_set_device(oracle_lib):
    """a handle is bound to its own device - every C entry point selects it and restores the caller's; an env
    on cuda:1 created and stepped while the current device is 0 must match the oracle.  Skipped on one-GPU boxes."""
    import torch

    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    env = _make(3, torch.device("cuda:1"), task=("obj_red", "bin_red"), action_mode="abs_pos")
    env.reset()
    orc = oracle_lib.OracleEnv(action_mode="abs_pos")
    orc.reset(None, 0, 0)
    a = np.array([0.1, 0.5, 0.45, 1.0], dtype=np.float32)
    for _ in range(3):
        env.step(torch.from_numpy(np.repeat(a[None], 3, axis=0)).to("cuda:1"))
        orc.step(a)
    assert torch.cuda.current_device() == 0
    assert reltol(_np(env.state["qpos"])[0], orc.qpos, TOL) < TOL
    env.close()
