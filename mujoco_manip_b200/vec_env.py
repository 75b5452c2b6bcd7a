"""PickPlaceVecEnv: N pick-and-place environments stepped on one B200, torch CUDA tensors in and out.

Vectorised counterpart of the reference's PickPlaceGymEnv (mujoco_manip/gym_env.py:39-602): same
constructor keywords, action modes, task sets, reward types and reset/step semantics per env, but
every array carries a leading env axis and lives on the GPU.  All arithmetic is done by the CUDA
library behind include/mm_manip.h; torch only owns the memory and the stream.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .constants import (ACTION_REPEAT, BINS, MAX_EPISODE_STEPS, OBJECTS, SPAWN_X_RANGE, SPAWN_Y_RANGE, TASK_SETS,
                        task_indices)
from .constants import MIN_OBJ_SEPARATION
from .randomization import sample_separated_positions

# slices of the packed [N,85] observation (layout written by csrc/mm_env.h write_obs)
OBS_SLICES = {
    "state": (0, 11, None),
    "state.ee.pos_quat_g": (11, 19, None),
    "state.ee.pos_rot6d_g": (19, 29, None),
    "state.ee.pos_quat_g_rel": (29, 37, None),
    "state.ee.pos_rot6d_g_rel": (37, 47, None),
    "target_bin_onehot": (47, 50, None),
    "target_obj_onehot": (50, 53, None),
    "keypoints_overhead": (53, 67, (7, 2)),
    "keypoints_wrist": (67, 81, (7, 2)),
    "target_obj_keypoints_overhead": (81, 83, None),
    "target_bin_keypoints_overhead": (83, 85, None),
}


def split_obs(packed: torch.Tensor) -> dict[str, torch.Tensor]:
    """Views of the packed observation under the reference's observation keys (gym_env.py:325-339)."""
    out = {}
    for k, (a, b, shp) in OBS_SLICES.items():
        v = packed[:, a:b]
        out[k] = v.reshape(packed.shape[0], *shp) if shp else v
    return out


class PickPlaceVecEnv:
    """N independent pick-and-place environments on one GPU.

    Args mirror PickPlaceGymEnv (gym_env.py:62-75).  Extra arguments:
        num_envs: environments owned by this process / GPU.
        device: CUDA device.
        seed: base seed of the per-env random streams.
        rng: "philox" (device counter-based sampler keyed by (seed, global env id, episode); no host
            sync on auto-reset) or "numpy" (host numpy PCG64 per env, bit-exact with the reference's
            `reset(seed=...)` draw order; used for parity and dataset replay).
        env_id_offset: global id of env 0 (rank * num_envs when sharded over GPUs) - random streams
            and the task cycle depend on the global id only, so results do not depend on the sharding.
        precision: "f64" (parity arithmetic) or "f32" (throughput arithmetic); state is stored in FP64.
        group: lanes cooperating on one env (8, 16 or 32).
        auto_reset: reset finished envs inside `step` (same-step mode: the returned observation of a
            finished env is the first observation of its next episode; `info["final_obs"]` keeps the last).
    """

    def __init__(self, num_envs: int, device: str | torch.device = "cuda:0", task=None, tasks="all",
                 action_mode: str = "ee_pos_quat_g_rel", reward_type: str = "dense",
                 max_episode_steps: int = MAX_EPISODE_STEPS, randomize_objects: bool = False,
                 spawn_x_range=SPAWN_X_RANGE, spawn_y_range=SPAWN_Y_RANGE, randomize_yaw: bool = False, seed: int = 0,
                 rng: str = "philox",
                 env_id_offset: int = 0, precision: str = "f64", group: int = 32, auto_reset: bool = True,
                 task_assignment: str = "random", load_balance: bool = False):
        if action_mode not in _lib.ACTION_MODES:
            raise ValueError(f"action_mode must be one of {_lib.ACTION_MODES}, got '{action_mode}'")
        if reward_type not in _lib.REWARD_TYPES:
            raise ValueError(f"reward_type must be one of {_lib.REWARD_TYPES}, got '{reward_type}'")
        if rng not in ("philox", "numpy"):
            raise ValueError("rng must be 'philox' or 'numpy'")
        if precision not in ("f64", "f32"):
            raise ValueError("precision must be 'f64' or 'f32'")
        if task_assignment not in ("random", "cycle"):
            raise ValueError("task_assignment must be 'random' or 'cycle'")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("PickPlaceVecEnv needs a CUDA device (there is no CPU path)")
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device visible (this package has no CPU path)")
        self.num_envs = int(num_envs)
        self.action_mode = action_mode
        self.action_dim = _lib.ACTION_DIMS[action_mode]
        self.reward_type = reward_type
        self.max_episode_steps = int(max_episode_steps)
        self.randomize_objects = bool(randomize_objects)
        self.randomize_yaw = bool(randomize_yaw)  # randomization.py:19,55-62 (only with randomize_objects)
        self.spawn_x_range = tuple(float(v) for v in spawn_x_range)
        self.spawn_y_range = tuple(float(v) for v in spawn_y_range)
        self.seed = int(seed)
        self.rng_kind = rng
        self.env_id_offset = int(env_id_offset)
        self.auto_reset = bool(auto_reset)
        self.task_assignment = task_assignment
        self._fixed_task = task
        self._task_pool = TASK_SETS[tasks] if isinstance(tasks, str) else list(tasks)
        pool = [self._fixed_task] if self._fixed_task is not None else self._task_pool
        self._pool_idx = torch.tensor([task_indices(t) for t in pool], dtype=torch.int32, device=self.device)

        self._L = _lib.lib()
        cfg = _lib.MMConfig(self.num_envs, self.device.index if self.device.index is not None else torch.cuda.current_device(), 0 if precision == "f64" else 1, int(group),
                            _lib.REWARD_TYPES.index(reward_type), self.max_episode_steps)
        self._cfg = cfg
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._L.mm_create(C.byref(cfg), C.byref(h)), "mm_create")
        self._h = h

        n, dev = self.num_envs, self.device
        self.state: dict[str, torch.Tensor] = {}
        for name, width, is_double in _lib.STATE_FIELDS:
            self.state[name] = torch.zeros((n, width), dtype=torch.float64 if is_double else torch.int32, device=dev)
        self._st = _lib.MMState(*[self.state[name].data_ptr() for name, _, _ in _lib.STATE_FIELDS])
        self._obs = torch.zeros((n, _lib.OBS_DIM), dtype=torch.float32, device=dev)
        self._final_obs = torch.zeros_like(self._obs)
        self._reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self._flags = torch.zeros((3, n), dtype=torch.uint8, device=dev)
        self._rc = torch.zeros((n, 6), dtype=torch.float32, device=dev) if reward_type == "staged" else None
        self._out = _lib.MMStepOut(self._obs.data_ptr(), self._reward.data_ptr(), self._flags[0].data_ptr(),
                                   self._flags[1].data_ptr(), self._flags[2].data_ptr(),
                                   self._rc.data_ptr() if self._rc is not None else None)
        self._actions = torch.zeros((n, _lib.ACTION_STRIDE), dtype=torch.float32, device=dev)
        self._fsm_actions = torch.zeros((n, _lib.ACTION_STRIDE), dtype=torch.float32, device=dev)
        self._obj_xy = torch.zeros((n, 6), dtype=torch.float64, device=dev)
        self._yaw_cs = torch.zeros((n, 6), dtype=torch.float64, device=dev)
        self._yaw_cs[:, 0::2] = 1.0
        self.last_yaw = torch.zeros((n, 3), dtype=torch.float64, device=dev)
        self._task = torch.zeros((n, 2), dtype=torch.int32, device=dev)
        self._mask = torch.ones(n, dtype=torch.uint8, device=dev)
        self._gid = torch.arange(n, dtype=torch.int64, device=dev) + self.env_id_offset
        self.episode_index = torch.zeros(n, dtype=torch.int64, device=dev)
        # episode statistics (device resident; see stats.py for the cross-GPU gather)
        # episodes, successes, sum of lengths, sum of returns, non-finite state resets, episodes that hit a workspace
        # overflow, placements whose 1000 attempts all failed (the reference raises there), spare  (mm_post_step)
        self.stats = torch.zeros(8, dtype=torch.float64, device=dev)
        self.last_attempts = torch.zeros(n, dtype=torch.int32, device=dev)
        self._ep_return = torch.zeros(n, dtype=torch.float64, device=dev)
        self._np_rngs: list | None = None
        self._out_host = None
        self._closed = False
        # load-aware scheduling lives in the library (k_schedule at the head of every mm_step; MM_BALANCE=0 switches it off):
        # `load_balance` is accepted for compatibility and has no effect
        self.load_balance = bool(load_balance)
        self._work = torch.zeros(n, dtype=torch.int32, device=dev)  # busy time of every env's last step (SM cycles / 256)
        _lib.check(self._L.mm_set_schedule(self._h, None, self._work.data_ptr()), "mm_set_schedule")

    # ------------------------------------------------------------------------------------------
    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def obs_packed(self) -> torch.Tensor:
        return self._obs

    def _draw(self, mask: torch.Tensor | None, seeds, task_override):
        """Fill self._obj_xy / self._task for the envs selected by `mask` (None = all)."""
        n = self.num_envs
        if self.rng_kind == "numpy":
            sel = np.arange(n) if mask is None else np.flatnonzero(mask.cpu().numpy())
            if self._np_rngs is None:
                self._np_rngs = [np.random.default_rng([self.seed, self.env_id_offset + i]) for i in range(n)]
            xy = self._obj_xy.cpu().numpy()
            yaw = self.last_yaw.cpu().numpy()
            tk = self._task.cpu().numpy()
            pool = self._pool_idx.cpu().numpy()
            for i in sel:
                if seeds is not None:
                    self._np_rngs[i] = np.random.default_rng(int(seeds[i]))
                g = self._np_rngs[i]
                if self.randomize_objects:  # placement draws come first (gym_env.py:496-501)
                    xy[i] = np.asarray(sample_separated_positions(g, 3, self.spawn_x_range, self.spawn_y_range)).ravel()
                    if self.randomize_yaw:  # one theta per cube after the accepted placement (randomization.py:55-56)
                        yaw[i] = [g.uniform(0, 2 * np.pi) for _ in range(3)]
                if task_override is not None:
                    tk[i] = task_override[i]
                elif self._fixed_task is not None:
                    tk[i] = pool[0]
                elif self.task_assignment == "cycle":
                    tk[i] = pool[(self.env_id_offset + i) % len(pool)]
                else:  # gym_env.py:516
                    tk[i] = pool[int(g.integers(len(pool)))]
            self._obj_xy.copy_(torch.from_numpy(xy))
            if self.randomize_yaw:
                self.last_yaw.copy_(torch.from_numpy(yaw))
                cs = np.stack([np.cos(yaw / 2), np.sin(yaw / 2)], axis=2).reshape(n, 6)
                self._yaw_cs.copy_(torch.from_numpy(cs))
            self._task.copy_(torch.from_numpy(tk))
        else:
            self._sample_episode(mask, advance=False)
            if task_override is not None:
                tk = torch.as_tensor(np.asarray(task_override), dtype=torch.int32, device=self.device)
                if mask is None:
                    self._task.copy_(tk)
                else:
                    m = mask.bool()
                    self._task[m] = tk[m]

    def _sample_episode(self, mask, advance: bool):
        """Device draw (mm_sample_episode) of placement / task / yaw for the masked envs, written in place."""
        mode = 0 if self._fixed_task is not None else (1 if self.task_assignment == "cycle" else 2)
        m = None if mask is None else mask
        _lib.check(self._L.mm_sample_episode(
            self._h, C.c_uint64(self.seed & 0xFFFFFFFFFFFFFFFF), self.env_id_offset, self.episode_index.data_ptr(),
            None if m is None else m.data_ptr(), self.spawn_x_range[0], self.spawn_x_range[1], self.spawn_y_range[0],
            self.spawn_y_range[1], float(MIN_OBJ_SEPARATION), self._pool_idx.data_ptr(), len(self._pool_idx), mode,
            1 if self.randomize_objects else 0, 1 if (self.randomize_objects and self.randomize_yaw) else 0,
            1 if advance else 0, self._obj_xy.data_ptr(), self._task.data_ptr(), self.last_attempts.data_ptr(),
            self.last_yaw.data_ptr(), self._yaw_cs.data_ptr(), self.stats.data_ptr(), self._stream()), "mm_sample_episode")

    def reset(self, *, seed=None, options: dict | None = None, mask: torch.Tensor | None = None, _from_step: bool = False):
        """Reset all envs (or those where `mask` is non-zero).

        seed: None, an int (new base seed) or a sequence of N per-env seeds (rng="numpy": each env is
            reseeded with `np.random.default_rng(seed_i)` exactly as `PickPlaceGymEnv.reset(seed=seed_i)`).
        options: {"task": (obj, bin)} for all envs or {"task": [N pairs]} per env (gym_env.py:511-512),
            {"obj_xy": [N,3,2]} to place the cubes explicitly, {"obj_yaw": [N,3]} (with obj_xy or
            randomize_objects) to set their yaw angles.
        """
        seeds = None
        if seed is not None:
            if np.ndim(seed) == 0:
                self.seed = int(seed)
                self._np_rngs = None
                if mask is None:
                    self.episode_index.zero_()
            else:
                seeds = np.asarray(seed, dtype=np.uint64)
                if seeds.shape != (self.num_envs,):
                    raise ValueError("per-env seeds must have shape (num_envs,)")
                if self.rng_kind != "numpy":
                    raise ValueError("per-env seeds need rng='numpy'")
        task_override = None
        if options and "task" in options:
            t = options["task"]
            if isinstance(t[0], str):
                task_override = np.tile(np.asarray(task_indices(t), dtype=np.int32), (self.num_envs, 1))
            else:
                task_override = np.asarray([task_indices(x) for x in t], dtype=np.int32)
        if mask is not None:
            mask = mask.to(torch.uint8).contiguous()
        self._draw(mask, seeds, task_override)
        use_xy = self.randomize_objects
        if options and "obj_xy" in options:
            xy = torch.as_tensor(np.asarray(options["obj_xy"], dtype=np.float64).reshape(self.num_envs, 6),
                                 device=self.device)
            self._obj_xy.copy_(xy)
            use_xy = True
        use_yaw = self.randomize_objects and self.randomize_yaw
        if options and "obj_yaw" in options:
            th = np.asarray(options["obj_yaw"], dtype=np.float64).reshape(self.num_envs, 3)
            self.last_yaw.copy_(torch.from_numpy(th))
            self._yaw_cs.copy_(torch.from_numpy(np.stack([np.cos(th / 2), np.sin(th / 2)], axis=2).reshape(-1, 6)))
            use_yaw = True
        _lib.check(self._L.mm_set_placement_yaw(self._h, self._yaw_cs.data_ptr() if use_yaw else None),
                   "mm_set_placement_yaw")
        m = mask
        _lib.check(self._L.mm_reset(self._h, C.byref(self._st), None if m is None else m.data_ptr(),
                                    self._obj_xy.data_ptr() if use_xy else None, self._task.data_ptr(),
                                    self._obs.data_ptr(), self._stream()), "mm_reset")
        if mask is None:
            self._ep_return.zero_()
            self.episode_index += 1
        else:
            self._ep_return.masked_fill_(mask.bool(), 0.0)
            self.episode_index += mask.to(torch.int64)
        return split_obs(self._obs), {}

    def step(self, actions: torch.Tensor):
        """One control step for every env (gym_env.py:536-581).  actions: [N, action_dim] float32 CUDA."""
        a = actions
        if a.device != self.device or a.dtype != torch.float32:
            a = a.to(device=self.device, dtype=torch.float32)
        if a.dim() != 2 or a.shape[0] != self.num_envs or a.shape[1] < self.action_dim:
            raise ValueError(f"actions must be [{self.num_envs}, {self.action_dim}]")
        if a.shape[1] == _lib.ACTION_STRIDE and a.is_contiguous():
            buf = a
        else:
            self._actions[:, : a.shape[1]].copy_(a)
            buf = self._actions
        _lib.check(self._L.mm_step(self._h, C.byref(self._st), buf.data_ptr(), _lib.ACTION_MODES.index(self.action_mode),
                                   C.byref(self._out), self._stream()), "mm_step")
        return self._post_step_autoreset()

    def step_host(self, h_actions: torch.Tensor, h_obs: torch.Tensor, h_reward: torch.Tensor, h_flags: torch.Tensor):
        """One control step with HOST buffers (pinned memory recommended), the end-to-end path of the C ABI
        (`mm_step_host`): actions [N,10] float32 go to the device, the step runs, observation [N,85], reward [N] and
        flags [3,N] uint8 (terminated, truncated, success) come back, the stream is synchronised; finished envs are
        then reset on the device (statistics, Philox draw, reset kernels) when auto_reset is on."""
        # (enqueue everything - copies, step, episode bookkeeping - then wait once: the bookkeeping kernels follow the
        # device-to-host copies in stream order, and the host thread does not sit between the step and them)
        _lib.check(self._L.mm_step_host_async(self._h, C.byref(self._st), h_actions.data_ptr(), _lib.ACTION_MODES.index(self.action_mode),
                                              h_obs.data_ptr(), h_reward.data_ptr(), h_flags[0].data_ptr(), h_flags[1].data_ptr(),
                                              h_flags[2].data_ptr(), self._stream()), "mm_step_host_async")
        if self._out_host is None:
            self._out_host = _lib.MMStepOut()
            _lib.check(self._L.mm_host_staging(self._h, C.byref(self._out_host)), "mm_host_staging")
        self._post_step_autoreset(self._out_host)
        torch.cuda.current_stream(self.device).synchronize()

    def _post_step_autoreset(self, out=None):
        """Bookkeeping after the mm_step launch: episode statistics and (optionally) the reset of finished envs - all in
        library kernels (mm_post_step, mm_sample_episode, mm_reset), no host round trip.  Returns the 5-tuple of `step`."""
        fb = self._flags.view(torch.bool)  # (zero-copy: the kernels write 0 / 1 bytes; valid until the next step)
        terminated, truncated, success = fb[0], fb[1], fb[2]
        info = {"success": success}
        if self._rc is not None:
            info["reward_components"] = self._rc
        fast = self.auto_reset and self.rng_kind == "philox"
        out = self._out if out is None else out
        _lib.check(self._L.mm_post_step(self._h, C.byref(self._st), C.byref(out), self._ep_return.data_ptr(),
                                        self._mask.data_ptr(), self._final_obs.data_ptr() if self.auto_reset else None,
                                        self.stats.data_ptr(), 1 if self.auto_reset else 0, self._stream()), "mm_post_step")
        reward = self._reward
        if self.auto_reset:
            info["final_obs"] = self._final_obs
            reward = reward.clone()
            if fast:
                self._sample_episode(self._mask, advance=True)
                use_xy = self.randomize_objects
                _lib.check(self._L.mm_set_placement_yaw(
                    self._h, self._yaw_cs.data_ptr() if (use_xy and self.randomize_yaw) else None), "mm_set_placement_yaw")
                _lib.check(self._L.mm_reset(self._h, C.byref(self._st), self._mask.data_ptr(),
                                            self._obj_xy.data_ptr() if use_xy else None, self._task.data_ptr(),
                                            out.obs, self._stream()), "mm_reset")
            else:  # host numpy generators (reference-exact draw order): the mask has to visit the host
                self.reset(mask=self._mask.clone(), _from_step=True)
        return split_obs(self._obs), reward, terminated, truncated, info

    def fsm_plan(self, n_steps: int = ACTION_REPEAT) -> torch.Tensor:
        """Advance every env's scripted FSM by one plan(n_steps) call (pick_and_place.py:167-277) and
        return the abs_pos actions [N,4] it commands (scripts/generate_dataset.py:145-148)."""
        _lib.check(self._L.mm_fsm_plan(self._h, C.byref(self._st), int(n_steps), self._fsm_actions.data_ptr(),
                                       self._stream()), "mm_fsm_plan")
        return self._fsm_actions[:, :4]

    @property
    def fsm_state(self) -> torch.Tensor:
        return self.state["fsm_i"][:, 0]

    @property
    def initial_ee_se3(self) -> torch.Tensor:
        """[N,4,4] initial EE poses (gym_env.py:472-475)."""
        T = torch.zeros((self.num_envs, 4, 4), dtype=torch.float64, device=self.device)
        T[:, :3, 3] = self.state["tinit"][:, :3]
        T[:, :3, :3] = self.state["tinit"][:, 3:].reshape(-1, 3, 3)
        T[:, 3, 3] = 1.0
        return T

    def get_state(self) -> dict[str, torch.Tensor]:
        out = {k: v.clone() for k, v in self.state.items()}
        out["episode_index"] = self.episode_index.clone()
        return out

    def set_state(self, st: dict[str, torch.Tensor]) -> None:
        for k, v in self.state.items():
            v.copy_(st[k])
        if "episode_index" in st:
            self.episode_index.copy_(st["episode_index"])

    def launch_count(self) -> int:
        out = C.c_longlong()
        _lib.check(self._L.mm_launch_count(self._h, C.byref(out)), "mm_launch_count")
        return out.value

    def task_names(self):
        t = self._task.cpu().numpy()
        return [(OBJECTS[o], BINS[b]) for o, b in t]

    def close(self) -> None:
        if not self._closed and self._h:
            torch.cuda.synchronize(self.device)
            self._L.mm_destroy(self._h)
            self._closed = True

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
