"""PickPlaceGymEnv: the reference's single-env Gymnasium-style API (mujoco_manip/gym_env.py:39-602)
on top of the CUDA step path (an N = 1 PickPlaceVecEnv).  Same constructor keywords, action modes,
task sets, reward types, reset / step return types, properties and error behaviour; observations are
the state / keypoint keys (rendering is off the hot path: image keys hold zeros)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, pose_utils
from .constants import (BINS, IMAGE_SIZE, KEYPOINT_BODIES, MAX_EPISODE_STEPS, OBJECTS, SPAWN_X_RANGE, SPAWN_Y_RANGE,
                        TASK_SETS, check_scene_xml)
from .randomization import sample_separated_positions
from .spaces import Box, Dict, Env

ACTION_MODES = _lib.ACTION_MODES
_HOME_QPOS = np.array([1.5708, -0.2, 0.0, -2.1, 0.0, 1.8, 0.785])
TARGET_ORI = np.array([[0.0, 1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, -1.0]])
_BIN_POS = {"bin_red": (-0.3, 0.55, 0.24), "bin_green": (0.0, 0.65, 0.24), "bin_blue": (0.3, 0.55, 0.24)}
_STATIC_POS = {"world": (0, 0, 0), "link0": (0, 0, 0), "table": (0, 0.45, 0.0), **_BIN_POS}


class PandaRobot:
    """robot.py:6-84 on the device state of env 0."""

    NUM_ARM_JOINTS = 7
    GRIPPER_OPEN = 255.0
    GRIPPER_CLOSED = 0.0
    EE_BODY_NAME = "hand"
    BODY_NAMES = frozenset({"link0", "link1", "link2", "link3", "link4", "link5", "link6", "link7", "hand",
                            "left_finger", "right_finger"})

    def __init__(self, vec):
        self._vec = vec

    @property
    def ee_pos(self) -> np.ndarray:
        return self._vec.state["eepose"][0, :3].cpu().numpy().copy()

    @property
    def ee_xmat(self) -> np.ndarray:
        return self._vec.state["eepose"][0, 3:].cpu().numpy().reshape(3, 3).copy()

    @property
    def arm_qpos(self) -> np.ndarray:
        return self._vec.state["qpos"][0, :7].cpu().numpy().copy()

    def set_arm_ctrl(self, targets) -> None:
        import torch

        self._vec.state["ctrl"][0, :7] = torch.as_tensor(np.asarray(targets, dtype=np.float64), device=self._vec.device)

    def open_gripper(self) -> None:
        self._vec.state["ctrl"][0, 7] = self.GRIPPER_OPEN

    def close_gripper(self) -> None:
        self._vec.state["ctrl"][0, 7] = self.GRIPPER_CLOSED

    @property
    def gripper_ctrl(self) -> float:
        return float(self._vec.state["ctrl"][0, 7])


class IKController:
    """controller.py:46-145: `compute` runs the device DLS IK (MM_OP_IK) on the kinematics of the last
    position stage; `reached` is the 2 cm test on the cached EE position."""

    def __init__(self, vec, robot, pos_tolerance: float = 0.02):
        self._vec, self.robot, self.pos_tolerance = vec, robot, pos_tolerance
        self.damping, self.max_dq, self.pos_gain, self.ori_gain, self.nullspace_gain = 1e-3, 5.0, 1.0, 1.0, 0.5

    def compute(self, target_pos) -> np.ndarray:
        import torch

        v = self._vec
        keep = v.state["ctrl"][0, :7].clone()
        t = torch.as_tensor(np.asarray(target_pos, dtype=np.float64).reshape(1, 3), device=v.device)
        _lib.check(v._L.mm_ops(v._h, C.byref(v._st), 1, t.data_ptr(), v._stream()), "mm_ops")
        q = v.state["ctrl"][0, :7].cpu().numpy().copy()
        v.state["ctrl"][0, :7] = keep  # compute() has no side effect in the reference
        return q

    def reached(self, target_pos) -> bool:
        return bool(np.linalg.norm(self.robot.ee_pos - np.asarray(target_pos)) < self.pos_tolerance)


class PickPlaceEnv:
    """env.py:72-176 on the device state of env 0 (no viewer)."""

    def __init__(self, vec):
        self._vec = vec
        self.viewer = None

    def _body_pose(self, name):
        if name in OBJECTS:
            q = self._vec.state["qpos"][0].cpu().numpy()
            a = 9 + 7 * OBJECTS.index(name)
            w, x, y, z = q[a + 3: a + 7]
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                          [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                          [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
            return q[a: a + 3].copy(), R
        if name == "hand":
            e = self._vec.state["eepose"][0].cpu().numpy()
            return e[:3].copy(), e[3:].reshape(3, 3).copy()
        if name in _STATIC_POS:
            return np.array(_STATIC_POS[name], dtype=np.float64), np.eye(3)
        return None

    def get_body_pos(self, name: str) -> np.ndarray:
        p = self._body_pose(name)
        if p is None:
            raise ValueError(f"Body '{name}' not found")
        return p[0]

    def get_body_xmat(self, name: str) -> np.ndarray:
        p = self._body_pose(name)
        if p is None:
            raise ValueError(f"Body '{name}' not found")
        return p[1]

    def reset_to_keyframe(self, name: str = "scene_start") -> None:
        if name != "scene_start":
            raise ValueError(f"Keyframe '{name}' not found")
        self._vec.reset(options={"task": self._vec.task_names()[0]})

    def randomize_objects(self, rng, x_range=SPAWN_X_RANGE, y_range=SPAWN_Y_RANGE, **kw):
        import torch

        pos = sample_separated_positions(rng, 3, x_range, y_range, kw.get("min_separation", 0.08))
        v = self._vec
        for o, (x, y) in enumerate(pos):
            quat = [1.0, 0.0, 0.0, 0.0]
            if kw.get("randomize_yaw", False):  # randomization.py:55-62
                theta = rng.uniform(0, 2 * np.pi)
                quat = [np.cos(theta / 2), 0.0, 0.0, np.sin(theta / 2)]
            v.state["qpos"][0, 9 + 7 * o: 16 + 7 * o] = torch.tensor([x, y, kw.get("obj_z", 0.26)] + quat,
                                                                   dtype=torch.float64, device=v.device)
        _lib.check(v._L.mm_ops(v._h, C.byref(v._st), 2, None, v._stream()), "mm_ops")  # mj_forward
        return {f"{n}_jnt": np.array([x, y, kw.get("obj_z", 0.26)]) for n, (x, y) in zip(OBJECTS, pos)}

    def step(self) -> None:
        """One physics step (mj_step, env.py:119-121)."""
        v = self._vec
        _lib.check(v._L.mm_ops(v._h, C.byref(v._st), 6, None, v._stream()), "mm_ops")

    def forward(self) -> None:
        v = self._vec
        _lib.check(v._L.mm_ops(v._h, C.byref(v._st), 2, None, v._stream()), "mm_ops")

    def sync(self) -> None:
        pass

    def is_running(self) -> bool:
        return True


class PickPlaceGymEnv(Env):
    metadata = {"render_modes": ["rgb_array", "human"], "render_fps": 30}

    def __init__(self, xml_path: str | None = None, task=None, tasks="all", action_mode: str = "ee_pos_quat_g_rel",
                 reward_type: str = "dense", image_size: int = IMAGE_SIZE, render_mode: str = "rgb_array",
                 max_episode_steps: int = MAX_EPISODE_STEPS, randomize_objects: bool = False,
                 spawn_x_range=SPAWN_X_RANGE, spawn_y_range=SPAWN_Y_RANGE, device: str = "cuda:0",
                 precision: str = "f64"):
        if action_mode not in ACTION_MODES:
            raise ValueError(f"action_mode must be one of {ACTION_MODES}, got '{action_mode}'")
        check_scene_xml(xml_path)  # the bundled scene (or None) is accepted, gym_env.py:64
        from .vec_env import PickPlaceVecEnv

        self._fixed_task = task
        self._task_pool = TASK_SETS[tasks] if isinstance(tasks, str) else tasks
        self._action_mode = action_mode
        self._reward_type = reward_type
        self._image_size = image_size
        self.render_mode = render_mode
        self._max_episode_steps = max_episode_steps
        self._randomize_objects = randomize_objects
        self._spawn_x_range = tuple(spawn_x_range)
        self._spawn_y_range = tuple(spawn_y_range)
        self._vec = PickPlaceVecEnv(1, device=device, task=task, tasks=tasks, action_mode=action_mode,
                                    reward_type=reward_type, max_episode_steps=max_episode_steps, rng="numpy",
                                    auto_reset=False, precision=precision)
        self._robot = PandaRobot(self._vec)
        self._env = PickPlaceEnv(self._vec)
        self._env._vec = self._vec
        self._controller = IKController(self._vec, self._robot)
        self._obj_name = ""
        self._bin_name = ""
        self._initial_ee_se3 = None
        if action_mode == "abs_pos":
            self.action_space = Box(low=np.array([-0.5, 0.0, 0.24, 0.0], dtype=np.float32),
                                    high=np.array([0.5, 0.8, 0.60, 1.0], dtype=np.float32))
        else:
            n = _lib.ACTION_DIMS[action_mode]
            low, high = np.full(n, -np.inf, dtype=np.float32), np.full(n, np.inf, dtype=np.float32)
            low[n - 1], high[n - 1] = 0.0, 1.0
            self.action_space = Box(low=low, high=high)
        nk = len(KEYPOINT_BODIES)
        f32 = np.float32
        self.observation_space = Dict({
            "image_overhead": Box(0, 255, (image_size, image_size, 3), dtype=np.uint8),
            "image_wrist": Box(0, 255, (image_size, image_size, 3), dtype=np.uint8),
            "state": Box(-np.inf, np.inf, (11,), dtype=f32),
            "state.ee.pos_quat_g": Box(-np.inf, np.inf, (8,), dtype=f32),
            "state.ee.pos_rot6d_g": Box(-np.inf, np.inf, (10,), dtype=f32),
            "state.ee.pos_quat_g_rel": Box(-np.inf, np.inf, (8,), dtype=f32),
            "state.ee.pos_rot6d_g_rel": Box(-np.inf, np.inf, (10,), dtype=f32),
            "target_bin_onehot": Box(0.0, 1.0, (3,), dtype=f32),
            "target_obj_onehot": Box(0.0, 1.0, (3,), dtype=f32),
            "keypoints_overhead": Box(0.0, 1.0, (nk, 2), dtype=f32),
            "keypoints_wrist": Box(0.0, 1.0, (nk, 2), dtype=f32),
            "target_obj_keypoints_overhead": Box(0.0, 1.0, (2,), dtype=f32),
            "target_bin_keypoints_overhead": Box(0.0, 1.0, (2,), dtype=f32),
        })

    # -- properties of the reference class (gym_env.py:210-243, 472-475) ---------------------------
    action_mode = property(lambda self: self._action_mode)
    pick_place_env = property(lambda self: self._env)
    robot = property(lambda self: self._robot)
    controller = property(lambda self: self._controller)
    obj_name = property(lambda self: self._obj_name)
    bin_name = property(lambda self: self._bin_name)

    @property
    def step_count(self) -> int:
        return int(self._vec.state["step_count"][0, 0])

    @property
    def _step_count(self) -> int:
        return self.step_count

    @property
    def initial_ee_se3(self) -> np.ndarray:
        return self._initial_ee_se3.copy()

    # staged-reward stickies (gym_env.py:129-133) live in the device flags word
    def _flag(self, bit):
        return bool(int(self._vec.state["flags"][0, 0]) & bit)

    _has_grasped = property(lambda self: self._flag(1))
    _has_lifted = property(lambda self: self._flag(2))
    _above_target = property(lambda self: self._flag(4))
    _has_placed = property(lambda self: self._flag(8))

    @property
    def _reward_hwm(self):
        return self._vec.state["hwm"][0].cpu().numpy().copy() if self._flag(16) else None

    def decode_action(self, action):
        """World-frame EE target and gripper command of an action (gym_env.py:252-281)."""
        action = np.asarray(action)
        m = self._action_mode
        if m == "abs_pos":
            return action[:3], action[3]
        T = pose_utils.se3_from_pos_quat_g(action) if "quat" in m else pose_utils.se3_from_pos_rot6d_g(action)
        g = action[7] if "quat" in m else action[9]
        if m.endswith("_rel"):
            T = self._initial_ee_se3 @ T
        return T[:3, 3], g

    def _obs(self) -> dict:
        packed = self._vec.obs_packed[0].cpu().numpy()
        from .vec_env import OBS_SLICES

        obs = {"image_overhead": np.zeros((self._image_size, self._image_size, 3), dtype=np.uint8),
               "image_wrist": np.zeros((self._image_size, self._image_size, 3), dtype=np.uint8)}
        for k, (a, b, shp) in OBS_SLICES.items():
            v = packed[a:b].copy()
            obs[k] = v.reshape(shp) if shp else v
        return obs

    _get_obs = _obs

    def reset(self, *, seed=None, options=None):
        super().reset(seed=seed)
        opts = {}
        if self._randomize_objects:  # placement draws first, then the task draw (gym_env.py:496-517)
            xy = sample_separated_positions(self.np_random, 3, self._spawn_x_range, self._spawn_y_range)
            opts["obj_xy"] = np.asarray(xy, dtype=np.float64).reshape(1, 3, 2)
        if options and "task" in options:
            self._obj_name, self._bin_name = options["task"]
        elif self._fixed_task is not None:
            self._obj_name, self._bin_name = self._fixed_task
        else:
            idx = self.np_random.integers(len(self._task_pool))
            self._obj_name, self._bin_name = self._task_pool[idx]
        if self._obj_name not in OBJECTS or self._bin_name not in BINS:
            raise ValueError(f"unknown task ({self._obj_name}, {self._bin_name})")
        opts["task"] = (self._obj_name, self._bin_name)
        self._vec.reset(options=opts)
        self._initial_ee_se3 = self._vec.initial_ee_se3[0].cpu().numpy()
        return self._obs(), {}

    def step(self, action):
        import torch

        a = np.zeros((1, _lib.ACTION_STRIDE), dtype=np.float32)
        act = np.asarray(action, dtype=np.float32).ravel()
        a[0, : act.size] = act
        _, r, te, tr, info = self._vec.step(torch.from_numpy(a).to(self._vec.device))
        out = {"success": bool(info["success"][0])}
        if self._reward_type == "staged":
            out["reward_components"] = info["reward_components"][0].cpu().numpy().copy()
        return self._obs(), float(r[0]), bool(te[0]), bool(tr[0]), out

    def _compute_reward(self):
        """(reward, success) at the current state, dense / sparse (gym_env.py:436-470); the staged reward
        is stateful and is produced by `step`."""
        obj, b, ee = self._env.get_body_pos(self._obj_name), self._env.get_body_pos(self._bin_name), self._robot.ee_pos
        success = bool(np.linalg.norm(obj[:2] - b[:2]) < 0.05 and obj[2] < b[2] + 0.06)
        if self._reward_type == "sparse":
            return (1.0 if success else 0.0), success
        r = -float(np.linalg.norm(ee - obj))
        if obj[2] > 0.30:
            r += 2.0 - float(np.linalg.norm(obj - b))
        if success:
            r += 10.0
        return r, success

    def render(self):
        if self.render_mode == "rgb_array":
            return np.zeros((self._image_size, self._image_size, 3), dtype=np.uint8)
        return None

    def close(self) -> None:
        self._vec.close()
