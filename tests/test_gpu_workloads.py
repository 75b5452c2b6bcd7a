"""Parity of the CUDA path on the workloads BASELINE.json names, at their real shapes (`-m gpu`):
  configs[1]  500-step random-action rollouts (long horizon: pile-ups with 100+ contacts appear after ~20 steps)
  configs[3]  ee_pos_rot6d_g_rel + device Philox placement at >= 8,192 envs (sampled envs against the oracle)
  configs[4]  half abs_pos + half ee_pos_rot6d_g, tasks = cross, scripted-FSM driven
  config 1    a scripted-FSM episode driven through ee_pos_quat_g_rel actions (scripts/generate_dataset.py:56-80)
The oracle (oracle/) is the checker; the CUDA path goes through the C ABI (PickPlaceVecEnv)."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from hostlib import GOLDEN, reltol

pytestmark = pytest.mark.gpu
TOL = 1e-5  # north_star: |a - b| <= 1e-5 * max(|b|, 1) over the first 50 steps


def _np(t):
    return t.detach().cpu().numpy()


def _make(n, dev, **kw):
    from mujoco_manip_b200 import PickPlaceVecEnv

    kw.setdefault("auto_reset", False)
    return PickPlaceVecEnv(n, device=dev, **kw)


def _dev(qpos, orcs, idx=None):
    idx = range(len(orcs)) if idx is None else idx
    return np.array([np.max(np.abs(qpos[i] - o.qpos) / np.maximum(np.abs(o.qpos), 1.0)) for i, o in zip(idx, orcs)])


def test_500_step_random_rollout_vs_oracle(cuda_device, oracle_lib):
    """configs[1] action distribution, 64 FP64 envs, a full 500-step episode against the oracle.

    Contact dynamics are chaotic: both sides carry ~1e-13 of rounding noise per step and, at a contact make / break in
    a pile-up, that noise decides a discrete event, after which the two trajectories are different (both valid)
    rollouts.  The test therefore runs a CONTROL next to the comparison - the oracle against itself with the arm
    joints of every env perturbed by 1e-13 rad after the reset - which measures that sensitivity on the same actions.
    Bar: (1) an env that is within 1e-5 (relative) has the oracle's contact count, at every step; (2) before its first
    excursion above 1e-5 every env tracks the oracle to 1e-5 by construction - and the number of envs that ever made
    such an excursion is, at steps 50, 200 and 500, not larger than the control's (+ 3 envs of slack, it is a different
    noise realisation); (3) at least 85 % of the envs are within 1e-5 over the first 50 steps; (4) no workspace
    overflow, no non-finite reset, and the rollout reaches contact-rich states (>= 60 contacts in some env)."""
    import torch

    n, steps = 64, 500
    env = _make(n, cuda_device, task=("obj_red", "bin_red"), action_mode="ee_pos_quat_g_rel", max_episode_steps=10000, rng="numpy")
    env.reset()
    orcs = [oracle_lib.OracleEnv(action_mode="ee_pos_quat_g_rel", max_episode_steps=10000) for _ in range(n)]
    ctrl = [oracle_lib.OracleEnv(action_mode="ee_pos_quat_g_rel", max_episode_steps=10000) for _ in range(n)]
    for o, c in zip(orcs, ctrl):
        o.reset(None, 0, 0)
        c.reset(None, 0, 0)
        c.qpos[:7] += 1e-13
    rng = np.random.default_rng(1234)
    T0 = _np(env.state["tinit"])[0]
    p0, R0 = T0[:3], T0[3:].reshape(3, 3)
    lo, hi = np.array([-0.3, 0.30, 0.30]), np.array([0.3, 0.65, 0.60])
    ok, ok_ctrl = np.ones(n, bool), np.ones(n, bool)
    pool = ThreadPoolExecutor(min(16, os.cpu_count() or 1))
    ncon_max, marks = 0, {}
    for t in range(steps):
        w = lo + (hi - lo) * rng.uniform(size=(n, 3))
        a = np.zeros((n, 8), dtype=np.float32)
        a[:, :3] = (w - p0) @ R0
        q = rng.normal(size=(n, 4))
        a[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
        a[:, 7] = rng.uniform(size=n) > 0.5
        env.step(torch.from_numpy(a).to(cuda_device))
        list(pool.map(lambda k: (orcs[k].step(a[k]), ctrl[k].step(a[k])), range(n)))
        qpos, diag = _np(env.state["qpos"]), _np(env.state["diag"])
        dev = _dev(qpos, orcs)
        dev_ctrl = np.array([np.max(np.abs(c.qpos - o.qpos) / np.maximum(np.abs(o.qpos), 1.0)) for c, o in zip(ctrl, orcs)])
        same_ncon = np.array([int(diag[k, 0]) == orcs[k].ncon for k in range(n)])
        ncon_max = max(ncon_max, int(diag[:, 0].max()))
        ok &= dev <= TOL
        ok_ctrl &= dev_ctrl <= TOL
        assert same_ncon[ok].all(), f"step {t}: an env within tolerance has a different contact count"
        if t + 1 in (50, 200, 500):
            marks[t + 1] = (int(ok.sum()), int(ok_ctrl.sum()))
    diag = _np(env.state["diag"])
    assert int(diag[:, 2].max()) == 0 and int(diag[:, 3].max()) == 0, "workspace overflow / non-finite reset"
    print("500-step rollout, envs within 1e-5 of the oracle (CUDA | 1e-13-perturbed oracle control): " +
          ", ".join(f"step {k}: {v[0]} | {v[1]} of {n}" for k, v in marks.items()) + f"; largest contact count {ncon_max}")
    for k, (mine, control) in marks.items():
        assert mine >= control - 3, f"step {k}: {mine} envs within tolerance, control {control}"
    assert marks[50][0] >= 0.85 * n
    assert ncon_max >= 60, "the rollout was supposed to reach contact-rich states"
    env.close()


def test_config3_rot6d_rel_philox_large_batch(cuda_device, oracle_lib):
    """configs[3] shape: 8,192 envs, ee_pos_rot6d_g_rel, randomized objects from the device Philox stream; 24 sampled envs
    (first, last, strided) against the oracle for 30 steps, placements bit-exact with the CPU Philox statement.  Every
    sampled env within 1e-5 with the oracle's contact count for 15 steps; up to step 30 at most two of them may have
    gone through a chaotic contact event (see test_500_step_random_rollout_vs_oracle for the control experiment)."""
    import torch

    from oracle import philox

    n, steps, off = 8192, 30, 65536 * 3
    env = _make(n, cuda_device, task=("obj_red", "bin_red"), action_mode="ee_pos_rot6d_g_rel", randomize_objects=True,
                rng="philox", seed=42, env_id_offset=off)
    env.reset()
    sample = sorted(set([0, 1, n - 1, n - 2] + list(range(7, n, n // 20))))
    xy = _np(env._obj_xy).reshape(n, 3, 2)
    orcs = []
    for i in sample:
        exp, att = philox.place(42, off + i, 0)
        assert np.array_equal(xy[i], exp) and att > 0, i
        o = oracle_lib.OracleEnv(action_mode="ee_pos_rot6d_g_rel")
        o.reset(exp, 0, 0)
        orcs.append(o)
    good = np.ones(len(sample), bool)
    gen = torch.Generator(device=cuda_device).manual_seed(7)
    T0 = env.state["tinit"][0]
    p0, R0 = T0[:3], T0[3:].reshape(3, 3)
    lo = torch.tensor([-0.3, 0.30, 0.30], device=cuda_device, dtype=torch.float64)
    hi = torch.tensor([0.3, 0.65, 0.60], device=cuda_device, dtype=torch.float64)
    for t in range(steps):
        w = lo + (hi - lo) * torch.rand((n, 3), device=cuda_device, dtype=torch.float64, generator=gen)
        a = torch.zeros((n, 10), device=cuda_device, dtype=torch.float32)
        a[:, :3] = ((w - p0) @ R0).float()
        a[:, 3:9] = torch.randn((n, 6), device=cuda_device, generator=gen)
        a[:, 9] = (torch.rand(n, device=cuda_device, generator=gen) > 0.5).float()
        obs, r, te, tr, info = env.step(a)
        ah = _np(a)
        qpos, diag = _np(env.state["qpos"]), _np(env.state["diag"])
        for k, (i, o) in enumerate(zip(sample, orcs)):
            o_obs, o_r, o_te, o_tr, o_info = o.step(ah[i])
            if not good[k]:
                continue
            if reltol(qpos[i], o.qpos, TOL) >= TOL:
                assert t >= 15, (t, i)
                good[k] = False
                continue
            assert int(diag[i, 0]) == o.ncon, (t, i)
            assert abs(float(r[i]) - o_r) < 1e-4
    assert good.sum() >= len(sample) - 2, good
    q = env.state["qpos"]
    assert torch.isfinite(q).all() and int(env.state["diag"][:, 2].max()) == 0 and int(env.state["diag"][:, 3].max()) == 0
    env.close()


def test_config4_mixed_modes_cross_tasks_fsm(cuda_device, oracle_lib):
    """configs[4] shape (scaled to 2 x 1,024 envs): half the envs take abs_pos actions, half ee_pos_rot6d_g (absolute EE
    pose), tasks = cross cycled by global env id, Philox placements, scripted-FSM expert; sampled envs against the
    oracle's FSM + engine until the grasp / lift phases are over (90 steps): FSM state and timer identical, qpos 1e-5."""
    from mujoco_manip_b200.constants import TASK_SETS, task_indices
    from mujoco_manip_b200.features import expert_action_encodings
    from oracle import philox

    half, steps = 1024, 90
    envs, orcs, samples = [], [], []
    for k, mode in enumerate(("abs_pos", "ee_pos_rot6d_g")):
        e = _make(half, cuda_device, tasks="cross", action_mode=mode, randomize_objects=True, rng="philox", seed=42,
                  env_id_offset=k * half, task_assignment="cycle", max_episode_steps=2000)
        e.reset()
        envs.append(e)
        smp = [0, 5, half // 2 + 1, half - 1]
        tk = _np(e._task)
        oo = []
        for i in smp:
            gid = k * half + i
            assert tuple(tk[i]) == task_indices(TASK_SETS["cross"][gid % 6])
            xy, _ = philox.place(42, gid, 0)
            o = oracle_lib.OracleEnv(action_mode=mode, max_episode_steps=2000)
            o.reset(xy, int(tk[i][0]), int(tk[i][1]))
            o.fsm_reset()
            oo.append(o)
        orcs.append(oo)
        samples.append(smp)
    seen = set()
    for t in range(steps):
        for k, e in enumerate(envs):
            a = e.fsm_plan(16).clone()
            act = a if e.action_mode == "abs_pos" else expert_action_encodings(e, a)[:, 8:18]
            e.step(act)
            ah, fs, qpos = _np(act), _np(e.state["fsm_i"]), _np(e.state["qpos"])
            for i, o in zip(samples[k], orcs[k]):
                o.fsm_plan(16)
                f = o.fsm_get()
                assert int(fs[i, 0]) == f["state"] and int(fs[i, 2]) == f["counter"], (t, k, i)
                seen.add(f["state"])
                o.step(ah[i])  # the oracle decodes the very action the device env was given
                assert reltol(qpos[i], o.qpos, TOL) < TOL, (t, k, i)
    assert {4, 5, 6} <= seen  # grasp, lift and transport were part of the comparison
    for e in envs:
        e.close()


def test_fsm_episode_through_quat_rel_actions(cuda_device):
    """Config 1 as the dataset generator runs it (scripts/generate_dataset.py:56-80, 140-196): the FSM's abs target is
    encoded as an ee_pos_quat_g_rel action and stepped in THAT mode.  The golden holds the reference's own actions."""
    from mujoco_manip_b200.features import expert_action_encodings

    g = np.load(os.path.join(GOLDEN, "fsm_quat_rel_red_red.npz"))
    env = _make(2, cuda_device, task=("obj_red", "bin_red"), action_mode="ee_pos_quat_g_rel", rng="numpy")
    env.reset()
    n = g["fsm_state"].shape[0]
    for t in range(n):
        a = env.fsm_plan(16).clone()
        assert int(env.fsm_state[0]) == int(g["fsm_state"][t]), t
        rel = expert_action_encodings(env, a)[:, 18:26]  # pos_quat_g_rel
        np.testing.assert_allclose(_np(rel)[0], g["action"][t][:8], rtol=0, atol=2e-6)
        obs, r, te, tr, info = env.step(rel)
        assert reltol(_np(env.state["qpos"])[0], g["qpos"][t], TOL) < TOL, t
        np.testing.assert_allclose(_np(env.obs_packed)[0], g["obs"][t], rtol=0, atol=2e-5)
        assert abs(float(r[0]) - g["reward"][t]) < 1e-4
    env.fsm_plan(16)
    assert int(env.fsm_state[0]) == 11
    env.close()


@pytest.mark.parametrize("switches", [{"MM_HEAVY": "8"}, {"MM_FUSE_CA": "1"}, {"MM_HEAVY": "8", "MM_FUSE_CA": "1", "MM_STREAMS": "1"},
                                      {"MM_GRAPH": "0"}, {"MM_GRAPH": "0", "MM_ISSUE": "0", "MM_STREAMS": "3", "MM_CHUNK": "1"},
                                      {"MM_STREAMS": "3", "MM_CHUNK": "1"}])
def test_optional_launch_plans_match_the_golden(cuda_device, monkeypatch, switches):
    """The launch-plan switches of mm_create (contact-rich envs by a CTA each = Grp<128>; stage A fused behind stage C;
    direct launches instead of the captured CUDA graph; chunk-major issue order)
    change the schedule, never the physics: the table-collision stress rollout (hull contacts, 70+ contacts) and a
    scripted-FSM episode stay on the golden trajectories recorded from the reference's Python on the oracle engine."""
    import torch

    for k, v in switches.items():
        monkeypatch.setenv(k, v)
    g = np.load(os.path.join(GOLDEN, "stress30_abs_pos_staged.npz"))
    env = _make(3, cuda_device, task=("obj_red", "bin_red"), action_mode="abs_pos", reward_type="staged", rng="numpy")
    env.reset()
    for t in range(g["action"].shape[0]):
        a = torch.from_numpy(np.repeat(g["action"][t][None, :4], 3, axis=0)).to(cuda_device)
        obs, r, te, tr, info = env.step(a)
        assert reltol(_np(env.state["qpos"])[0], g["qpos"][t], TOL) < TOL, t
        assert int(_np(env.state["diag"])[0, 0]) == int(g["ncon"][t])
        assert abs(float(r[0]) - g["reward"][t]) < 1e-5
    assert torch.equal(env.state["qpos"][0], env.state["qpos"][2])
    env.close()
    g = np.load(os.path.join(GOLDEN, "fsm_abs_green_blue_seed42_staged.npz"))
    q = g["init_qpos"]
    env = _make(2, cuda_device, action_mode="abs_pos", reward_type="staged", rng="numpy")
    from mujoco_manip_b200.constants import BINS, OBJECTS

    xy = np.array([q[9:11], q[16:18], q[23:25]])
    env.reset(options={"task": (OBJECTS[int(g["obj_idx"])], BINS[int(g["bin_idx"])]), "obj_xy": np.stack([xy, xy])})
    for t in range(g["fsm_state"].shape[0]):
        a = env.fsm_plan(16).clone()
        assert int(env.fsm_state[0]) == int(g["fsm_state"][t]), t
        env.step(a)
        assert reltol(_np(env.state["qpos"])[0], g["qpos"][t], TOL) < TOL, t
    env.close()


def test_dataset_writer_rows_match_the_reference_generator(cuda_device, tmp_path):
    """f3: ExpertDatasetWriter against the frames of the reference's own `run_episode` (scripts/generate_dataset.py:83-198,
    golden recorded by running it unmodified on the oracle engine): same episode seeds and task cycle, same number of
    frames per episode, every state feature row by row (float32 observations: 2e-5), phase strings and task strings equal."""
    from mujoco_manip_b200 import dataset

    g = np.load(os.path.join(GOLDEN, "dataset_rows_seed42_cross.npz"))
    w = dataset.ExpertDatasetWriter(str(tmp_path), num_envs=2, device=str(cuda_device), tasks="cross", reward_type="staged",
                                    randomize_objects=True, seed=42)
    meta = w.generate(3)
    assert meta["episode_seeds"] == [int(x) for x in g["episode_seeds"]] and meta["shards"] == 1
    for ep in range(3):
        rows = dataset.read_episode(str(tmp_path), ep)
        n = g[f"ep{ep}.observation.state"].shape[0]
        assert len(rows["frame_index"]) == n == meta["episode_lengths"][ep], (ep, len(rows["frame_index"]), n)
        assert rows["task"][0] == str(g[f"ep{ep}.task_string"])
        assert rows["observation.phase_description"] == [str(x) for x in g[f"ep{ep}.observation.phase_description"]]
        for k in [str(x) for x in g["keys"]]:
            if k == "observation.phase_description":
                continue
            np.testing.assert_allclose(rows[k], g[f"ep{ep}.{k}"], rtol=0, atol=2e-5, err_msg=f"episode {ep} {k}")


def test_results_do_not_depend_on_the_schedule(cuda_device, monkeypatch):
    """The library orders the envs of every step by the busy time of their previous step (k_schedule) and cuts the batch
    into chunks on several streams: 600 envs in three ragged chunks (256 + 256 + 88) with the schedule on, replayed
    with the schedule off in one chunk - every state array and observation must be bit-identical over 25 random-action
    steps (contacts, hull pairs and the convex queue included), and the schedule must have been a real permutation."""
    import torch

    n, steps = 600, 25
    gen = torch.Generator(device="cpu").manual_seed(5)
    acts = []
    for _ in range(steps):
        a = torch.zeros(n, 8)
        a[:, :3] = (torch.rand(n, 3, generator=gen) - 0.5) * torch.tensor([0.6, 0.5, 0.5])
        q = torch.randn(n, 4, generator=gen)
        a[:, 3:7] = q / q.norm(dim=1, keepdim=True)
        a[:, 7] = (torch.rand(n, generator=gen) > 0.5).float()
        acts.append(a.to(cuda_device))

    def run(env_vars):
        for k, v in env_vars.items():
            monkeypatch.setenv(k, v)
        env = _make(n, cuda_device, action_mode="ee_pos_quat_g_rel", randomize_objects=True, rng="philox", seed=11, tasks="all")
        env.reset()
        obs = []
        for a in acts:
            env.step(a)
            obs.append(env.obs_packed.clone())
        out = {k: v.clone() for k, v in env.state.items()}, torch.stack(obs), env._work.clone()
        env.close()
        for k in env_vars:
            monkeypatch.delenv(k)
        return out

    st1, obs1, work1 = run({"MM_CHUNK": "256", "MM_STREAMS": "3"})
    st0, obs0, _ = run({"MM_BALANCE": "0", "MM_STREAMS": "1"})
    assert int((work1 > 0).sum()) == n and float(work1.float().std()) > 0  # busy times were recorded and differ
    assert torch.equal(obs1, obs0)
    for k in ("qpos", "qvel", "ctrl", "warm", "step_count", "diag"):
        assert torch.equal(st1[k], st0[k]), k
    assert int(st1["diag"][:, 0].max()) >= 30  # pile-ups were part of it


def test_step_graph_cache_eviction(cuda_device, monkeypatch):
    """mm_step keeps one captured CUDA graph per distinct argument set (64 at most, least recently used out): 70 action
    buffers, visited twice, give the same trajectory as direct launches."""
    import ctypes as C

    import torch

    from mujoco_manip_b200 import _lib

    n = 40
    gen = torch.Generator(device="cpu").manual_seed(9)
    bufs = []
    for _ in range(70):
        a = torch.zeros(n, _lib.ACTION_STRIDE)
        a[:, :3] = (torch.rand(n, 3, generator=gen) - 0.5) * 0.2
        a[:, 3] = 1.0
        a[:, 7] = (torch.rand(n, generator=gen) > 0.5).float()
        bufs.append(a.to(cuda_device))

    def run(graph):
        monkeypatch.setenv("MM_GRAPH", graph)
        env = _make(n, cuda_device, action_mode="ee_pos_quat_g_rel")
        env.reset()
        mode = _lib.ACTION_MODES.index("ee_pos_quat_g_rel")
        for rep in range(2):
            for a in bufs:  # the raw entry point with the caller's own buffers: every buffer is a new argument set
                _lib.check(env._L.mm_step(env._h, C.byref(env._st), a.data_ptr(), mode, C.byref(env._out), env._stream()), "mm_step")
        torch.cuda.synchronize()
        out = {k: v.clone() for k, v in env.state.items()}, env.launch_count()
        env.close()
        return out

    (st1, l1), (st0, l0) = run("1"), run("0")
    assert l1 == l0  # the replay accounts for every kernel of the captured step
    for k in ("qpos", "qvel", "ctrl", "warm", "step_count"):
        assert torch.equal(st1[k], st0[k]), k
