"""Dataset feature schema of the state-only expert rollouts (same keys, shapes, dtypes and dimension names as
the reference's mujoco_manip/features.py:10-236) and the packing of a vectorised step into those rows.

Images are not produced (rendering is off the hot path): the two image features are listed for schema
compatibility and omitted from packed rows; `observation.phase_description` is packed as the FSM state index,
which `phase_description()` turns into the reference's strings on the host."""
from __future__ import annotations

from .constants import IMAGE_SIZE

_POSE8 = ["x", "y", "z", "qx", "qy", "qz", "qw", "gripper"]
_POSE10 = ["x", "y", "z", "r11", "r12", "r13", "r21", "r22", "r23", "gripper"]
_KP = [f"{n}_{c}" for n in ("red", "green", "blue", "bin_red", "bin_green", "bin_blue", "hand") for c in ("u", "v")]


def _f32(n):
    return {"dtype": "float32", "shape": (n,), "names": None}


FEATURES = {
    "observation.images.overhead": {"dtype": "image", "shape": (IMAGE_SIZE, IMAGE_SIZE, 3), "names": ["height", "width", "channels"]},
    "observation.images.wrist": {"dtype": "image", "shape": (IMAGE_SIZE, IMAGE_SIZE, 3), "names": ["height", "width", "channels"]},
    "observation.state": _f32(11),
    "observation.state.ee.pos_quat_g": _f32(8),
    "observation.state.ee.pos_rot6d_g": _f32(10),
    "observation.state.ee.pos_quat_g_rel": _f32(8),
    "observation.state.ee.pos_rot6d_g_rel": _f32(10),
    "action.ee.pos_quat_g": _f32(8),
    "action.ee.pos_rot6d_g": _f32(10),
    "action.ee.pos_quat_g_rel": _f32(8),
    "action.ee.pos_rot6d_g_rel": _f32(10),
    "observation.target_bin_onehot": _f32(3),
    "observation.target_obj_onehot": _f32(3),
    "observation.keypoints_overhead": _f32(14),
    "observation.keypoints_wrist": _f32(14),
    "observation.target_obj_keypoints_overhead": _f32(2),
    "observation.target_bin_keypoints_overhead": _f32(2),
    "observation.phase_description": {"dtype": "string", "shape": (1,), "names": None},
    "next.reward": _f32(6),
}

DIM_NAMES: dict[str, list[str]] = {
    "observation.state": ["ee_x", "ee_y", "ee_z", "gripper"] + [f"q{i}" for i in range(7)],
    "observation.state.ee.pos_quat_g": list(_POSE8),
    "observation.state.ee.pos_rot6d_g": list(_POSE10),
    "observation.state.ee.pos_quat_g_rel": list(_POSE8),
    "observation.state.ee.pos_rot6d_g_rel": list(_POSE10),
    "action.ee.pos_quat_g": list(_POSE8),
    "action.ee.pos_rot6d_g": list(_POSE10),
    "action.ee.pos_quat_g_rel": list(_POSE8),
    "action.ee.pos_rot6d_g_rel": list(_POSE10),
    "observation.target_bin_onehot": ["red", "green", "blue"],
    "observation.target_obj_onehot": ["red", "green", "blue"],
    "observation.keypoints_overhead": list(_KP),
    "observation.keypoints_wrist": list(_KP),
    "observation.target_obj_keypoints_overhead": ["u", "v"],
    "observation.target_bin_keypoints_overhead": ["u", "v"],
    "next.reward": ["total", "reach_obj", "pick_obj", "reach_target", "place_obj", "reach_home"],
}

# packed observation slices -> feature keys
_OBS_FEATURES = {
    "observation.state": (0, 11), "observation.state.ee.pos_quat_g": (11, 19), "observation.state.ee.pos_rot6d_g": (19, 29),
    "observation.state.ee.pos_quat_g_rel": (29, 37), "observation.state.ee.pos_rot6d_g_rel": (37, 47),
    "observation.target_bin_onehot": (47, 50), "observation.target_obj_onehot": (50, 53),
    "observation.keypoints_overhead": (53, 67), "observation.keypoints_wrist": (67, 81),
    "observation.target_obj_keypoints_overhead": (81, 83), "observation.target_bin_keypoints_overhead": (83, 85),
}
_ACT_FEATURES = {"action.ee.pos_quat_g": (0, 8), "action.ee.pos_rot6d_g": (8, 18), "action.ee.pos_quat_g_rel": (18, 26),
                 "action.ee.pos_rot6d_g_rel": (26, 36)}


def expert_action_encodings(env, abs_actions):
    """[N,36] CUDA tensor with the four encodings of the expert's abs_pos actions (generate_dataset.py:56-80)."""
    import ctypes as C

    import torch

    from . import _lib

    a = torch.zeros((env.num_envs, _lib.ACTION_STRIDE), dtype=torch.float32, device=env.device)
    a[:, : abs_actions.shape[1]] = abs_actions
    out = torch.empty((env.num_envs, 36), dtype=torch.float32, device=env.device)
    _lib.check(env._L.mm_expert_actions(env._h, C.byref(env._st), a.data_ptr(), out.data_ptr(), env._stream()),
               "mm_expert_actions")
    return out


def pack_rows(pre_step_obs_packed, action_encodings, fsm_state, reward_components):
    """One dataset row per env as a dict of CUDA tensors keyed like FEATURES (pre-step observation, the four
    action encodings, FSM state index for the phase string, next.reward)."""
    row = {k: pre_step_obs_packed[:, a:b] for k, (a, b) in _OBS_FEATURES.items()}
    row.update({k: action_encodings[:, a:b] for k, (a, b) in _ACT_FEATURES.items()})
    row["observation.phase_description"] = fsm_state
    row["next.reward"] = reward_components
    return row


def phase_description(fsm_state: int, obj_name: str, bin_name: str) -> str:
    """The reference's phase string for an FSM state index (pick_and_place.py:127-149)."""
    oc, bc = obj_name.replace("obj_", ""), bin_name.replace("bin_", "")
    if fsm_state in (1, 11):
        return "idle"
    if fsm_state == 10:
        return "retreating to neutral position"
    return {2: f"approaching the {oc} cube", 3: f"grasping the {oc} cube", 4: f"grasping the {oc} cube",
            5: f"lifting the {oc} cube", 6: f"transporting the {oc} cube to the {bc} bin",
            7: f"transporting the {oc} cube to the {bc} bin", 8: f"placing the {oc} cube in the {bc} bin",
            9: f"placing the {oc} cube in the {bc} bin"}[fsm_state]
