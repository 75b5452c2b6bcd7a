#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference package on an engine, and compare two such sets.

    python tools/make_golden.py                          # here: reference (read-only, /root/reference) on the CPU oracle
                                                         # engine through oracle/fake_mujoco.py -> tests/golden/
    python tools/make_golden.py --engine mujoco --reference /path/to/mujoco-manip --out /tmp/golden_mujoco --compare
                                                         # on a machine WITH MuJoCo 3.5: same rollouts on the real engine,
                                                         # then the deviation of tests/golden/ (oracle engine) from them

What the committed fixtures pin: everything the reference implements in Python above the engine boundary -
action decode, DLS IK, the pick-and-place FSM, rewards, observation packing, reset/RNG order - as
executed by the reference's own code.  The engine underneath is the oracle restatement (MuJoCo is
not installable in the build container), so engine-level trajectories remain "parity unpinned" against real
MuJoCo until someone runs the second command; its report (largest relative deviation of qpos / qvel / EE pose per
file and the first step above 1e-5, FSM-state disagreements) is what would pin the oracle.
Rendering is stubbed in both modes (no GL needed; image observations are off the hot path).

The first form runs here only (the GPU box has no /root/reference).
"""
import argparse
import os
import sys

sys.dont_write_bytecode = True  # never write __pycache__ into the reference tree
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

_ap = argparse.ArgumentParser()
_ap.add_argument("--engine", default="oracle", choices=["oracle", "mujoco"])
_ap.add_argument("--reference", default="/root/reference", help="checkout of the reference package")
_ap.add_argument("--out", default=None, help="output directory (default tests/golden for the oracle engine)")
_ap.add_argument("--compare", action="store_true", help="after generating, compare tests/golden/ with the new set")
_ap.add_argument("--compare-only", nargs=2, metavar=("A", "B"), help="only compare two existing directories")
ARGS = _ap.parse_args()

if ARGS.compare_only is None:
    if ARGS.engine == "oracle":
        from oracle import fake_mujoco  # noqa: E402

        fake_mujoco.install(ARGS.reference)
    else:
        import mujoco  # noqa: E402  (the real engine)

        class _NoRenderer:  # physics does not depend on the renderer; avoids the GL context
            def __init__(self, model, height=224, width=224):
                self._h, self._w = height, width

            def update_scene(self, data, camera=None):
                pass

            def render(self):
                return np.zeros((self._h, self._w, 3), dtype=np.uint8)

            def close(self):
                pass

        mujoco.Renderer = _NoRenderer
        sys.path.insert(0, ARGS.reference)

    from mujoco_manip import pose_utils as P  # noqa: E402
    from mujoco_manip.constants import ACTION_REPEAT, TASK_SETS  # noqa: E402
    from mujoco_manip.controller import TARGET_ORI  # noqa: E402
    from mujoco_manip.gym_env import PickPlaceGymEnv  # noqa: E402
    from mujoco_manip.pick_and_place import PickAndPlaceTask  # noqa: E402
    from mujoco_manip.randomization import _sample_separated_positions  # noqa: E402

GOLDEN = os.path.join(REPO, "tests", "golden")
OUT = ARGS.out or (GOLDEN if ARGS.engine == "oracle" else os.path.join(REPO, "tests", "golden_mujoco"))
OBS_KEYS = ["state", "state.ee.pos_quat_g", "state.ee.pos_rot6d_g", "state.ee.pos_quat_g_rel",
            "state.ee.pos_rot6d_g_rel", "target_bin_onehot", "target_obj_onehot"]
KP_KEYS = ["keypoints_overhead", "keypoints_wrist", "target_obj_keypoints_overhead", "target_bin_keypoints_overhead"]


def pack_obs(obs):
    return np.concatenate([np.asarray(obs[k], dtype=np.float32).ravel() for k in OBS_KEYS + KP_KEYS])


def get_actions(target_pos, g, init_inv):  # same arithmetic as scripts/generate_dataset.py:56-80
    T = P.pos_rotmat_to_se3(target_pos, TARGET_ORI)
    Tr = init_inv @ T
    return {"abs_pos": np.array([*target_pos, g], dtype=np.float32),
            "ee_pos_quat_g": P.se3_to_pos_quat_g(T, g), "ee_pos_rot6d_g": P.se3_to_pos_rot6d_g(T, g),
            "ee_pos_quat_g_rel": P.se3_to_pos_quat_g(Tr, g), "ee_pos_rot6d_g_rel": P.se3_to_pos_rot6d_g(Tr, g)}


def snapshot(env):
    d = env.pick_place_env.data
    return dict(qpos=d.qpos.copy(), qvel=d.qvel.copy(), ctrl=d.ctrl.copy(), warm=d.qacc_warmstart.copy(),
                ee_pos=env.robot.ee_pos, ee_R=env.robot.ee_xmat.ravel())


def fsm_episode(mode, task, randomize, seed, reward_type="dense", max_steps=400):
    env = PickPlaceGymEnv(task=None, tasks="all", action_mode=mode, reward_type=reward_type,
                          randomize_objects=randomize)
    obs, _ = env.reset(seed=seed, options={"task": task})
    fsm = PickAndPlaceTask(env.pick_place_env, env.robot, env.controller, tasks=[task])
    init_inv = np.linalg.inv(env.initial_ee_se3)
    rec = {k: [] for k in ("action", "obs", "reward", "terminated", "truncated", "success", "fsm_state", "target",
                           "gripper", "qpos", "qvel", "ctrl", "warm", "ee_pos", "ee_R", "rc", "counter")}
    init = snapshot(env)
    obs0 = pack_obs(obs)
    n = 0
    while not fsm.is_done and n < max_steps:
        fsm.plan(n_steps=ACTION_REPEAT)
        tp = fsm.target_pos if fsm.target_pos is not None else env.robot.ee_pos
        a = get_actions(tp, fsm.gripper_val, init_inv)[mode]
        rec["fsm_state"].append(fsm.state.value)
        rec["counter"].append(fsm.settle_counter)
        rec["target"].append(np.array(tp, dtype=np.float64))
        rec["gripper"].append(fsm.gripper_val)
        pad = np.zeros(10, dtype=np.float32)
        pad[: a.size] = a
        rec["action"].append(pad)
        obs, r, te, tr, info = env.step(a)
        s = snapshot(env)
        for k in ("qpos", "qvel", "ctrl", "warm", "ee_pos", "ee_R"):
            rec[k].append(s[k])
        rec["obs"].append(pack_obs(obs))
        rec["reward"].append(r)
        rec["terminated"].append(te)
        rec["truncated"].append(tr)
        rec["success"].append(info["success"])
        rec["rc"].append(info.get("reward_components", np.zeros(6, dtype=np.float32)))
        n += 1
    out = {k: np.array(v) for k, v in rec.items()}
    out.update(init_qpos=init["qpos"], init_ee_pos=init["ee_pos"], init_ee_R=init["ee_R"], obs0=obs0,
               T_init=env.initial_ee_se3, final_fsm_state=fsm.state.value,
               obj_idx=["obj_red", "obj_green", "obj_blue"].index(task[0]),
               bin_idx=["bin_red", "bin_green", "bin_blue"].index(task[1]))
    env.close()
    return out


def fsm_multi_episode(tasks, seed, max_steps=900):
    """The reference FSM run over a LIST of tasks at gym-step level (plan(16) -> abs_pos action -> env.step), as
    tests/test_pick_and_place.py:274-289 and scripts/generate_dataset.py:140-196 drive it.  The env's own task (reward,
    one-hots) stays (obj_red, bin_red): the FSM's list is independent of it (pick_and_place.py:91 vs gym_env.py:511-517)."""
    env = PickPlaceGymEnv(task=("obj_red", "bin_red"), action_mode="abs_pos", randomize_objects=True, max_episode_steps=2000)
    env.reset(seed=seed)
    fsm = PickAndPlaceTask(env.pick_place_env, env.robot, env.controller, tasks=list(tasks))
    rec = {k: [] for k in ("fsm_state", "task_index", "counter", "target", "gripper", "qpos", "status", "reward", "success")}
    init = snapshot(env)
    n = 0
    while not fsm.is_done and n < max_steps:
        status = fsm.plan(n_steps=ACTION_REPEAT)
        tp = fsm.target_pos if fsm.target_pos is not None else env.robot.ee_pos
        rec["status"].append(status)
        rec["fsm_state"].append(fsm.state.value)
        rec["task_index"].append(fsm.task_index)
        rec["counter"].append(fsm.settle_counter)
        rec["target"].append(np.array(tp, dtype=np.float64))
        rec["gripper"].append(fsm.gripper_val)
        obs, r, te, tr, info = env.step(np.array([*tp, fsm.gripper_val], dtype=np.float32))
        rec["qpos"].append(env.pick_place_env.data.qpos.copy())
        rec["reward"].append(r)
        rec["success"].append(info["success"])
        n += 1
    out = {k: np.array(v) for k, v in rec.items()}
    names_o, names_b = ["obj_red", "obj_green", "obj_blue"], ["bin_red", "bin_green", "bin_blue"]
    out.update(init_qpos=init["qpos"], final_fsm_state=fsm.state.value,
               tasks=np.array([[names_o.index(o), names_b.index(b)] for o, b in tasks]))
    env.close()
    return out


def dataset_rows(seed=42, num_episodes=3, tasks="cross"):
    """Frames of the reference's own dataset generator (scripts/generate_dataset.py:83-198 `run_episode`, imported
    unmodified - hydra / omegaconf are stubbed, they only decorate its CLI) for the first episodes of a `tasks` run with
    randomize_objects=True, reward_type=staged: the row-level contract of the state-only dataset writer."""
    import importlib.util
    import types

    for name in ("hydra", "omegaconf"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.main = lambda *a, **k: (lambda f: f)
            m.DictConfig, m.OmegaConf = dict, object
            sys.modules[name] = m
    spec = importlib.util.spec_from_file_location("ref_generate_dataset", os.path.join(ARGS.reference, "scripts", "generate_dataset.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    from mujoco_manip.features import FEATURES

    keys = [k for k in FEATURES if "images" not in k]
    env = PickPlaceGymEnv(action_mode="abs_pos", reward_type="staged", randomize_objects=True)
    ss = np.random.SeedSequence(seed).spawn(num_episodes)
    ep_seeds = [int(c.generate_state(1)[0]) for c in ss]
    task_list = TASK_SETS[tasks]
    out = {"episode_seeds": np.array(ep_seeds, dtype=np.uint64), "keys": np.array(keys)}
    names_o, names_b = ["obj_red", "obj_green", "obj_blue"], ["bin_red", "bin_green", "bin_blue"]
    for ep in range(num_episodes):
        o, b = task_list[ep % len(task_list)]
        frames = gen.run_episode(env, o, b, set(keys), reward_type="staged", episode_seed=ep_seeds[ep])
        out[f"ep{ep}.task"] = np.array([names_o.index(o), names_b.index(b)])
        out[f"ep{ep}.task_string"] = np.array(frames[0]["task"])
        for k in keys:
            if k == "observation.phase_description":
                out[f"ep{ep}.{k}"] = np.array([f[k] for f in frames])
            else:
                out[f"ep{ep}.{k}"] = np.stack([np.asarray(f[k], dtype=np.float32).ravel() for f in frames])
    env.close()
    return out


def random_rollout(mode, seed, n_steps=50, reward_type="dense", stress=False):
    env = PickPlaceGymEnv(task=("obj_red", "bin_red"), action_mode=mode, reward_type=reward_type)
    obs, _ = env.reset(seed=seed)
    rng = np.random.default_rng(seed)
    T_init = env.initial_ee_se3
    init_inv = np.linalg.inv(T_init)
    rec = {k: [] for k in ("action", "obs", "reward", "terminated", "truncated", "success", "qpos", "qvel", "ctrl",
                           "warm", "ee_pos", "ee_R", "ncon")}
    obs0 = pack_obs(obs)
    for t in range(n_steps):
        if stress:  # drive the hand low: finger/hand hulls against table, bins and cubes
            w = np.array([rng.uniform(-0.3, 0.3), rng.uniform(0.30, 0.65), rng.uniform(0.20, 0.40)])
        else:
            w = np.array([rng.uniform(-0.3, 0.3), rng.uniform(0.30, 0.65), rng.uniform(0.30, 0.60)])
        g = float(rng.uniform() > 0.5)
        # random (unused) rotation to exercise the decoders
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        Rr = P.quat_xyzw_to_rotmat(q)
        T = P.pos_rotmat_to_se3(w, Rr)
        Tr = init_inv @ T
        a = {"abs_pos": np.array([*w, g], dtype=np.float32), "ee_pos_quat_g": P.se3_to_pos_quat_g(T, g),
             "ee_pos_rot6d_g": P.se3_to_pos_rot6d_g(T, g), "ee_pos_quat_g_rel": P.se3_to_pos_quat_g(Tr, g),
             "ee_pos_rot6d_g_rel": P.se3_to_pos_rot6d_g(Tr, g)}[mode]
        pad = np.zeros(10, dtype=np.float32)
        pad[: a.size] = a
        rec["action"].append(pad)
        obs, r, te, tr, info = env.step(a)
        s = snapshot(env)
        for k in ("qpos", "qvel", "ctrl", "warm", "ee_pos", "ee_R"):
            rec[k].append(s[k])
        rec["obs"].append(pack_obs(obs))
        rec["reward"].append(r)
        rec["terminated"].append(te)
        rec["truncated"].append(tr)
        rec["success"].append(info["success"])
        rec["ncon"].append(env.pick_place_env.data.ncon)
    out = {k: np.array(v) for k, v in rec.items()}
    out.update(obs0=obs0, T_init=T_init)
    env.close()
    return out


def pose_vectors(seed=0):
    rng = np.random.default_rng(seed)
    Rs, quats, r6, T8, T10, d8, d10 = [], [], [], [], [], [], []
    for i in range(64):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        R = P.quat_xyzw_to_rotmat(q)
        if i < 4:  # exercise each branch of the R->quat conversion
            R = [np.eye(3), np.diag([1.0, -1, -1]), np.diag([-1.0, 1, -1]), np.diag([-1.0, -1, 1])][i]
        p = rng.uniform(-1, 1, size=3)
        g = float(rng.uniform())
        T = P.pos_rotmat_to_se3(p, R)
        Rs.append(R)
        quats.append(P.rotmat_to_quat_xyzw(R))
        r6.append(P.rotmat_to_6d(R))
        v8, v10 = P.se3_to_pos_quat_g(T, g), P.se3_to_pos_rot6d_g(T, g)
        d8.append(v8)
        d10.append(v10)
        T8.append(P.se3_from_pos_quat_g(v8))
        a10 = v10.copy()
        a10[3:9] *= rng.uniform(0.5, 2.0)  # un-normalised 6D input
        a10[6:9] += 0.3 * a10[3:6]
        T10.append(np.concatenate([a10.astype(np.float64), P.se3_from_pos_rot6d_g(a10).ravel()]))
    return dict(R=np.array(Rs), quat=np.array(quats), rot6d=np.array(r6), dof8=np.array(d8), dof10=np.array(d10),
                T_from8=np.array(T8), in10_T_from10=np.array(T10))


def reset_vectors():
    seeds = [0, 1, 7, 42, 123, 2684470948, 4091952314, 233227757, 3276785861]
    xy, task, obs0, qpos = [], [], [], []
    env = PickPlaceGymEnv(tasks="all", randomize_objects=True, action_mode="ee_pos_rot6d_g_rel")
    for s in seeds:
        obs, _ = env.reset(seed=s)
        d = env.pick_place_env.data
        xy.append(np.array([d.qpos[9:11], d.qpos[16:18], d.qpos[23:25]]))
        task.append([["obj_red", "obj_green", "obj_blue"].index(env.obj_name),
                     ["bin_red", "bin_green", "bin_blue"].index(env.bin_name)])
        obs0.append(pack_obs(obs))
        qpos.append(d.qpos.copy())
    # raw sampler draws too
    rng = np.random.default_rng(42)
    first = np.array(_sample_separated_positions(rng, 3, (-0.20, 0.20), (0.30, 0.45), 0.08))
    ss = np.random.SeedSequence(42).spawn(8)
    ep_seeds = np.array([int(c.generate_state(1)[0]) for c in ss], dtype=np.uint64)
    env.close()
    return dict(seeds=np.array(seeds, dtype=np.uint64), obj_xy=np.array(xy), task=np.array(task), obs0=np.array(obs0),
                qpos=np.array(qpos), sampler42=first, episode_seeds42=ep_seeds,
                task_sets_all=np.array([[["obj_red", "obj_green", "obj_blue"].index(o), ["bin_red", "bin_green", "bin_blue"].index(b)]
                                        for o, b in TASK_SETS["all"]]),
                task_sets_cross=np.array([[["obj_red", "obj_green", "obj_blue"].index(o), ["bin_red", "bin_green", "bin_blue"].index(b)]
                                          for o, b in TASK_SETS["cross"]]))


def main():
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "pose_utils.npz"), **pose_vectors())
    np.savez_compressed(os.path.join(OUT, "reset_seeds.npz"), **reset_vectors())
    # config 1: the reference's own CPU-runnable case
    np.savez_compressed(os.path.join(OUT, "fsm_quat_rel_red_red.npz"),
                        **fsm_episode("ee_pos_quat_g_rel", ("obj_red", "bin_red"), False, None))
    np.savez_compressed(os.path.join(OUT, "fsm_abs_green_blue_seed42_staged.npz"),
                        **fsm_episode("abs_pos", ("obj_green", "bin_blue"), True, 42, reward_type="staged"))
    np.savez_compressed(os.path.join(OUT, "fsm_rot6d_rel_blue_red_seed7.npz"),
                        **fsm_episode("ee_pos_rot6d_g_rel", ("obj_blue", "bin_red"), True, 7))
    np.savez_compressed(os.path.join(OUT, "dataset_rows_seed42_cross.npz"), **dataset_rows())
    np.savez_compressed(os.path.join(OUT, "fsm_multi3_seed5.npz"),
                        **fsm_multi_episode([("obj_red", "bin_red"), ("obj_green", "bin_green"), ("obj_blue", "bin_blue")], 5))
    np.savez_compressed(os.path.join(OUT, "fsm_multi2_cross_seed11.npz"),
                        **fsm_multi_episode([("obj_blue", "bin_red"), ("obj_red", "bin_green")], 11))
    for mode in ("abs_pos", "ee_pos_quat_g", "ee_pos_rot6d_g", "ee_pos_quat_g_rel", "ee_pos_rot6d_g_rel"):
        np.savez_compressed(os.path.join(OUT, f"random50_{mode}.npz"), **random_rollout(mode, 1234, 50))
    np.savez_compressed(os.path.join(OUT, "stress30_abs_pos_staged.npz"),
                        **random_rollout("abs_pos", 99, 30, reward_type="staged", stress=True))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def compare_dirs(a, b, tol=1e-5):
    """Deviation of the trajectories in directory `a` from those in `b` (same file names): |x_a - x_b| <= tol * max(|x_b|, 1)
    is the bar of BASELINE.json's north_star.  Returns the number of files above the bar."""
    bad = 0
    for f in sorted(os.listdir(b)):
        if not f.endswith(".npz") or not os.path.exists(os.path.join(a, f)):
            continue
        A, B = np.load(os.path.join(a, f)), np.load(os.path.join(b, f))
        line, worst = [], 0.0
        for k in ("qpos", "qvel", "ee_pos", "ee_R", "obs", "reward"):
            if k not in A.files or k not in B.files:
                continue
            x, y = np.asarray(A[k], dtype=np.float64), np.asarray(B[k], dtype=np.float64)
            n = min(len(x), len(y))
            if n == 0:
                continue
            x, y = x[:n].reshape(n, -1), y[:n].reshape(n, -1)
            rel = (np.abs(x - y) / np.maximum(np.abs(y), 1.0)).max(axis=1)
            first = int(np.argmax(rel > tol)) if (rel > tol).any() else -1
            at50 = float(rel[: min(n, 50)].max())
            worst = max(worst, at50 if k in ("qpos", "qvel", "ee_pos", "ee_R") else 0.0)
            line.append(f"{k} max {rel.max():.2e} (first 50 steps {at50:.2e}, first step above {tol:g}: {first})")
        for k in ("fsm_state", "terminated", "truncated", "success"):
            if k in A.files and k in B.files:
                n = min(len(A[k]), len(B[k]))
                mism = int((np.asarray(A[k][:n]) != np.asarray(B[k][:n])).sum()) + abs(len(A[k]) - len(B[k]))
                if mism:
                    line.append(f"{k}: {mism} steps differ (lengths {len(A[k])} / {len(B[k])})")
        if f in ("pose_utils.npz", "reset_seeds.npz"):
            same = all(np.array_equal(A[k], B[k]) for k in B.files if k in A.files)
            line.append("identical" if same else "DIFFERENT")
            worst = 0.0 if same else 1.0
        ok = worst <= tol
        bad += 0 if ok else 1
        print(("ok   " if ok else "ABOVE") + f" {f}: " + "; ".join(line))
    return bad


if __name__ == "__main__":
    if ARGS.compare_only:
        sys.exit(1 if compare_dirs(*ARGS.compare_only) else 0)
    main()
    if ARGS.compare and os.path.abspath(OUT) != os.path.abspath(GOLDEN):
        print(f"\n== tests/golden (oracle engine) against {OUT} ({ARGS.engine} engine) ==")
        sys.exit(1 if compare_dirs(GOLDEN, OUT) else 0)
