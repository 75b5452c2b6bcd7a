"""Object placement on reset.

Two generators, both reproducing the reference's rejection sampler (mujoco_manip/randomization.py:70-98:
per attempt three x draws then three y draws, accept when every pairwise XY distance is >= 0.08 m,
at most 1000 attempts):

* `sample_separated_positions` - host numpy Generator; with `np.random.default_rng(seed)` it is
  bit-exact with `PickPlaceGymEnv.reset(seed=seed)` of the reference.  Used by the N=1 env and by
  `PickPlaceVecEnv(rng="numpy")`.
* `philox_placements` - the CUDA library's counter-based sampler (csrc/mm_rng.h) keyed by
  (seed, global env id, episode index); no host round trip.  oracle/philox.py states it on the CPU.
"""
from __future__ import annotations

import numpy as np

from .constants import MAX_REJECTION_ATTEMPTS, MIN_OBJ_SEPARATION

OBJ_JOINT_NAMES = ("obj_red_jnt", "obj_green_jnt", "obj_blue_jnt")


def all_separated(positions, min_sep: float) -> bool:
    pts = list(positions)
    lim = min_sep * min_sep
    for a in range(len(pts)):
        for b in range(a + 1, len(pts)):
            dx = pts[a][0] - pts[b][0]
            dy = pts[a][1] - pts[b][1]
            if dx * dx + dy * dy < lim:
                return False
    return True


def sample_separated_positions(rng: np.random.Generator, n: int, x_range, y_range,
                               min_separation: float = MIN_OBJ_SEPARATION) -> list[tuple[float, float]]:
    """n XY positions with pairwise distance >= min_separation; RuntimeError when 1000 attempts fail
    (randomization.py:84-87)."""
    for _ in range(MAX_REJECTION_ATTEMPTS):
        xs = rng.uniform(x_range[0], x_range[1], size=n)
        ys = rng.uniform(y_range[0], y_range[1], size=n)
        cand = list(zip(xs.tolist(), ys.tolist()))
        if all_separated(cand, min_separation):
            return cand
    raise RuntimeError(f"Failed to sample {n} positions with min_separation={min_separation} "
                       f"in {MAX_REJECTION_ATTEMPTS} attempts")


def philox_placements(env, min_separation: float = MIN_OBJ_SEPARATION):
    """Device draw for every env of a PickPlaceVecEnv: returns (obj_xy [N,6] f64, task_draw [N] i32,
    attempts [N] i32) CUDA tensors for the env's current (seed, global ids, episode indices)."""
    import ctypes as C

    import torch

    from . import _lib

    n, dev = env.num_envs, env.device  # (the env itself draws through mm_sample_episode; this is the stand-alone call)
    xy = torch.empty((n, 6), dtype=torch.float64, device=dev)
    draw = torch.empty(n, dtype=torch.int32, device=dev)
    att = torch.empty(n, dtype=torch.int32, device=dev)
    _lib.check(env._L.mm_sample_placements(env._h, C.c_uint64(env.seed & 0xFFFFFFFFFFFFFFFF), env.env_id_offset,
                                           env.episode_index.data_ptr(), env.spawn_x_range[0], env.spawn_x_range[1],
                                           env.spawn_y_range[0], env.spawn_y_range[1], float(min_separation),
                                           len(env._pool_idx), xy.data_ptr(), draw.data_ptr(), att.data_ptr(),
                                           env._stream()), "mm_sample_placements")
    return xy, draw, att
