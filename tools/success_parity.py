#!/usr/bin/env python3
"""Scripted-FSM success-rate parity over N seeded episodes (north_star: within 1 percentage point over 1000):
CUDA path (FP64) vs the CPU oracle on identical placements / tasks (BASELINE.json configs[2] semantics:
tasks cycle env % 9, episode seeds spawned from SeedSequence(42), placements by numpy PCG64).
Run on the GPU box:  python tools/success_parity.py [--episodes 1000]  -> one JSON line."""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mujoco_manip_b200 import PickPlaceVecEnv  # noqa: E402
from oracle import oracle  # noqa: E402  (checker)

ap = argparse.ArgumentParser()
ap.add_argument("--episodes", type=int, default=1000)
ap.add_argument("--precision", default="f64")
a = ap.parse_args()
n = a.episodes
dev = torch.device("cuda:0")
seeds = [int(c.generate_state(1)[0]) for c in np.random.SeedSequence(42).spawn(n)]
env = PickPlaceVecEnv(n, device=dev, tasks="all", action_mode="abs_pos", randomize_objects=True, rng="numpy", auto_reset=False,
                      task_assignment="cycle", max_episode_steps=2000, precision=a.precision)
env.reset(seed=seeds)
xy = env._obj_xy.cpu().numpy().reshape(n, 3, 2)
tasks = env._task.cpu().numpy()
length = torch.zeros(n, dtype=torch.int32, device=dev)
succ = torch.zeros(n, dtype=torch.bool, device=dev)
t0 = time.time()
for t in range(600):
    running = env.fsm_state != 11
    if not bool(running.any()):
        break
    act = env.fsm_plan(16)
    obs, r, te, tr, info = env.step(act)
    length += running.to(torch.int32)
    succ = torch.where(running, info["success"], succ)
torch.cuda.synchronize()
t_gpu = time.time() - t0


def run(i):
    o = oracle.OracleEnv(action_mode="abs_pos")
    s, ln, _ = o.run_fsm_episode(xy[i], int(tasks[i, 0]), int(tasks[i, 1]), 2000)
    return s, ln


t0 = time.time()
with ThreadPoolExecutor(os.cpu_count() or 8) as ex:
    ref = list(ex.map(run, range(n)))
t_cpu = time.time() - t0
rs, rl = np.array([x[0] for x in ref]), np.array([x[1] for x in ref])
gs, gl = succ.cpu().numpy(), length.cpu().numpy()
print(json.dumps({"episodes": n, "precision": a.precision, "success_rate_cuda": float(gs.mean()), "success_rate_oracle": float(rs.mean()),
                  "abs_diff_pp": 100 * abs(float(gs.mean()) - float(rs.mean())), "per_episode_success_agreement": float((gs == rs).mean()),
                  "per_episode_length_agreement": float((gl == rl).mean()), "mean_length": float(gl.mean()),
                  "cuda_seconds": t_gpu, "oracle_seconds": t_cpu, "oracle_threads": os.cpu_count()}))
