"""B200-native batched implementation of the mujoco-manip step hot path.

Public surface (mirrors the reference package mujoco_manip/__init__.py + gym_env.py):
    PickPlaceVecEnv  - N envs on one GPU, torch CUDA tensors
    PickPlaceGymEnv  - the reference's single-env Gymnasium-style API on top of the same CUDA path
"""
from .constants import (ACTION_REPEAT, ALL_TASKS, BINS, CROSS_TASKS, MATCH_TASKS, MAX_EPISODE_STEPS, OBJECTS, TASK_SETS)

__all__ = ["PickPlaceVecEnv", "PickPlaceGymEnv", "ACTION_REPEAT", "ALL_TASKS", "BINS", "CROSS_TASKS", "MATCH_TASKS",
           "MAX_EPISODE_STEPS", "OBJECTS", "TASK_SETS"]


def __getattr__(name):  # torch is imported only when an env class is requested
    if name == "PickPlaceVecEnv":
        from .vec_env import PickPlaceVecEnv

        return PickPlaceVecEnv
    if name == "PickPlaceGymEnv":
        from .gym_env import PickPlaceGymEnv

        return PickPlaceGymEnv
    raise AttributeError(name)
