#!/usr/bin/env python3
"""Smallest meaningful case for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_case.py`:
14 envs (not a multiple of the CTA size: exercises the padding warps), contact-rich actions (hand pushed into the
table and cubes: box, hull and EPA paths), randomized placements, auto-reset, scripted FSM and the engine-level ops."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C  # noqa: E402

import torch  # noqa: E402

from mujoco_manip_b200 import PickPlaceVecEnv, _lib  # noqa: E402

dev = torch.device("cuda:0")
env = PickPlaceVecEnv(14, device=dev, tasks="all", action_mode="abs_pos", randomize_objects=True, max_episode_steps=3,
                      reward_type="staged", seed=5)
env.reset()
a = torch.tensor([[-0.15, 0.45, 0.27, 0.0]] * 14, device=dev)
a[::2, 0] = 0.15
a[1::3, 2] = 0.12
for t in range(4):
    env.step(a)
env.step(env.fsm_plan(16))
tgt = torch.tensor([[0.1, 0.5, 0.4]] * 14, dtype=torch.float64, device=dev)
_lib.check(env._L.mm_ops(env._h, C.byref(env._st), 7, tgt.data_ptr(), env._stream()), "mm_ops")
torch.cuda.synchronize()
print("sanitize case done: max ncon", int(env.state["diag"][:, 0].max()), "overflow", int(env.state["diag"][:, 2].max()))
