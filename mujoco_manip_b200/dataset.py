"""State-only expert dataset generation: the workload of the reference's scripts/generate_dataset.py (configs[2] of
BASELINE.json) on the batched CUDA path.

The reference runs one episode at a time (`run_episode`, scripts/generate_dataset.py:83-198): reset with the episode's
seed and task, then `while not fsm.is_done: plan(16) -> frame from the PRE-step observation + the four encodings of the
expert action + phase description -> step -> next.reward from info`.  Here N episodes run side by side on one
PickPlaceVecEnv (scripted FSM, action encodings, staged reward and observation packing are library kernels) and the
rows are cut per episode on the host.  Same row content, feature names, dtypes and shapes as `FEATURES`
(mujoco_manip/features.py:10-106) minus the two image features; same episode seeds (`SeedSequence(seed).spawn(n)`,
:263-268), same task cycle (:276-277), same `metadata.json` (:302-313).  The LeRobot container format itself is out of
scope: shards are Parquet files (pyarrow) with one row per frame plus episode_index / frame_index / timestamp / task.
"""
from __future__ import annotations

import json
import os

import numpy as np

from .constants import ACTION_REPEAT, BINS, CONTROL_FPS, OBJECTS, SPAWN_X_RANGE, SPAWN_Y_RANGE, TASK_SETS
from .features import _ACT_FEATURES, _OBS_FEATURES, FEATURES, phase_description

STATE_FEATURES = [k for k in FEATURES if "images" not in k]  # everything the state-only path produces


def make_task_string(obj_name: str, bin_name: str) -> str:
    """scripts/generate_dataset.py:41-53"""
    return f"Pick {obj_name.replace('obj_', '')} object and place in {bin_name.replace('bin_', '')} bin"


def episode_seeds(seed: int, num_episodes: int) -> list[int]:
    """scripts/generate_dataset.py:263-268"""
    return [int(cs.generate_state(1)[0]) for cs in np.random.SeedSequence(seed).spawn(num_episodes)]


def assemble_episode_rows(obs, enc, state_after, running, rc, task, feature_keys) -> dict:
    """Rows of ONE episode from its env's per-step records (host arrays): obs [T,85] pre-step packed observation,
    enc [T,36] action encodings, state_after [T] FSM state after plan(), running [T] bool (the FSM was not DONE before
    the plan), rc [T,6] reward components or None, task (obj_name, bin_name)."""
    keep = np.flatnonzero(running)
    rows = {}
    for k, (a, b) in _OBS_FEATURES.items():
        if k in feature_keys:
            rows[k] = np.ascontiguousarray(obs[keep, a:b], dtype=np.float32)
    for k, (a, b) in _ACT_FEATURES.items():
        if k in feature_keys:
            rows[k] = np.ascontiguousarray(enc[keep, a:b], dtype=np.float32)
    if "observation.phase_description" in feature_keys:
        rows["observation.phase_description"] = [phase_description(int(s), task[0], task[1]) for s in state_after[keep]]
    if "next.reward" in feature_keys and rc is not None:
        rows["next.reward"] = np.ascontiguousarray(rc[keep], dtype=np.float32)
    rows["task"] = [make_task_string(*task)] * len(keep)
    rows["frame_index"] = np.arange(len(keep), dtype=np.int64)
    rows["timestamp"] = (np.arange(len(keep)) / CONTROL_FPS).astype(np.float32)
    return rows


def write_parquet(path: str, rows: dict) -> None:
    import pyarrow as pa
    import pyarrow.parquet as pq

    cols = {}
    for k, v in rows.items():
        if isinstance(v, np.ndarray) and v.ndim == 2:
            cols[k] = pa.FixedSizeListArray.from_arrays(pa.array(v.reshape(-1)), v.shape[1])
        else:
            cols[k] = pa.array(v)
    pq.write_table(pa.table(cols), path)


class ExpertDatasetWriter:
    """Generates `num_episodes` expert episodes (arguments as configs/generate.yaml) and writes them under `root`."""

    def __init__(self, root: str, num_envs: int = 256, device: str = "cuda:0", task=None, tasks="all", reward_type: str = "staged",
                 randomize_objects: bool = False, seed: int = 0, spawn_x_range=SPAWN_X_RANGE, spawn_y_range=SPAWN_Y_RANGE,
                 features=None, max_frames: int = 2000, episodes_per_shard: int = 256):
        if task is not None:
            if len(tuple(task)) != 2:
                raise ValueError(f"task must be [obj, bin], got {task}")
            self.task_list = [tuple(task)]
        elif tasks in TASK_SETS:
            self.task_list = TASK_SETS[tasks]
        else:
            raise ValueError(f"Unknown task set '{tasks}'. Choose from: {list(TASK_SETS.keys())}")
        keys = STATE_FEATURES if features is None else [k for k in features if "images" not in k]
        unknown = [k for k in keys if k not in FEATURES]
        if unknown:
            raise ValueError(f"Unknown feature keys: {unknown}. Valid keys: {list(FEATURES.keys())}")
        if reward_type != "staged":
            keys = [k for k in keys if k != "next.reward"]
        self.feature_keys = set(keys)
        self.root, self.num_envs, self.device = root, int(num_envs), device
        self.reward_type, self.randomize_objects, self.seed = reward_type, bool(randomize_objects), int(seed)
        self.spawn_x_range, self.spawn_y_range = tuple(spawn_x_range), tuple(spawn_y_range)
        self.max_frames, self.episodes_per_shard = int(max_frames), int(episodes_per_shard)
        self.config = {"num_episodes": None, "task": list(task) if task is not None else None, "tasks": tasks,
                       "reward_type": reward_type, "randomize_objects": bool(randomize_objects), "seed": int(seed),
                       "spawn_x_range": list(self.spawn_x_range), "spawn_y_range": list(self.spawn_y_range),
                       "features": sorted(self.feature_keys)}

    def rollout_wave(self, env, tasks, seeds):
        """One wave of N side-by-side episodes -> per-env row dicts."""
        import torch

        from .features import expert_action_encodings

        n = env.num_envs
        opts = {"task": tasks}
        if seeds is not None:
            env.reset(seed=np.asarray(seeds, dtype=np.uint64), options=opts)
        else:
            env.reset(options=opts)
        rec = {k: [] for k in ("obs", "enc", "state", "running", "rc")}
        for _ in range(self.max_frames):
            before = env.fsm_state.clone()
            running = before != 11
            if not bool(running.any()):
                break
            rec["obs"].append(env.obs_packed.clone())  # PRE-step observation (generate_dataset.py:150-155)
            a = env.fsm_plan(ACTION_REPEAT).clone()
            rec["state"].append(env.fsm_state.clone())
            rec["enc"].append(expert_action_encodings(env, a))
            rec["running"].append(running)
            _, _, _, _, info = env.step(a)
            if "reward_components" in info:
                rec["rc"].append(info["reward_components"].clone())
        host = {k: (torch.stack(v).cpu().numpy() if v else None) for k, v in rec.items()}
        out = []
        for i in range(n):
            out.append(assemble_episode_rows(host["obs"][:, i], host["enc"][:, i], host["state"][:, i], host["running"][:, i],
                                             None if host["rc"] is None else host["rc"][:, i], tasks[i], self.feature_keys))
        return out

    def generate(self, num_episodes: int) -> dict:
        from .vec_env import PickPlaceVecEnv

        os.makedirs(os.path.join(self.root, "data"), exist_ok=True)
        seeds = episode_seeds(self.seed, num_episodes) if self.randomize_objects else None
        n = min(self.num_envs, num_episodes)
        env = PickPlaceVecEnv(n, device=self.device, tasks="all", action_mode="abs_pos", reward_type=self.reward_type,
                              randomize_objects=self.randomize_objects, spawn_x_range=self.spawn_x_range,
                              spawn_y_range=self.spawn_y_range, rng="numpy", auto_reset=False, max_episode_steps=self.max_frames)
        shard, shard_eps, nshard, total_frames, lengths = [], 0, 0, 0, []

        def flush():
            nonlocal shard, shard_eps, nshard
            if not shard:
                return
            merged = {}
            for k in shard[0]:
                vals = [r[k] for r in shard]
                merged[k] = np.concatenate(vals) if isinstance(vals[0], np.ndarray) else [x for v in vals for x in v]
            write_parquet(os.path.join(self.root, "data", f"chunk-{nshard:05d}.parquet"), merged)
            shard, shard_eps, nshard = [], 0, nshard + 1

        for first in range(0, num_episodes, n):
            eps = [min(first + i, num_episodes - 1) for i in range(n)]  # a short last wave repeats its final episode
            tasks = [self.task_list[e % len(self.task_list)] for e in eps]
            rows = self.rollout_wave(env, tasks, None if seeds is None else [seeds[e] for e in eps])
            for i, e in enumerate(eps):
                if first + i >= num_episodes:
                    break
                r = rows[i]
                r["episode_index"] = np.full(len(r["frame_index"]), e, dtype=np.int64)
                r["index"] = np.arange(total_frames, total_frames + len(r["frame_index"]), dtype=np.int64)
                total_frames += len(r["frame_index"])
                lengths.append(len(r["frame_index"]))
                shard.append(r)
                shard_eps += 1
                if shard_eps >= self.episodes_per_shard:
                    flush()
        flush()
        env.close()
        meta = dict(self.config, num_episodes=num_episodes, fps=CONTROL_FPS, robot_type="franka_panda", total_frames=total_frames,
                    episode_lengths=lengths, shards=nshard,
                    feature_shapes={k: list(FEATURES[k]["shape"]) for k in sorted(self.feature_keys)})
        if seeds is not None:
            meta["episode_seeds"] = seeds  # O(1) replay of any episode (generate_dataset.py:302-308)
        with open(os.path.join(self.root, "metadata.json"), "w") as f:
            json.dump(meta, f, indent=2)
        return meta


def read_episode(root: str, episode_index: int) -> dict:
    """Rows of one episode back from the shards (numpy arrays / lists), for tests and replay."""
    import pyarrow.parquet as pq

    ddir = os.path.join(root, "data")
    for name in sorted(os.listdir(ddir)):
        t = pq.read_table(os.path.join(ddir, name))
        ep = t.column("episode_index").to_numpy()
        sel = np.flatnonzero(ep == episode_index)
        if len(sel) == 0:
            continue
        out = {}
        for k in t.column_names:
            col = t.column(k).take(sel).combine_chunks()
            if str(col.type).startswith("fixed_size_list"):
                w = col.type.list_size
                out[k] = col.flatten().to_numpy(zero_copy_only=False).reshape(-1, w)
            elif str(col.type) in ("string", "large_string"):
                out[k] = col.to_pylist()
            else:
                out[k] = col.to_numpy(zero_copy_only=False)
        return out
    raise KeyError(f"episode {episode_index} not found under {root}")
